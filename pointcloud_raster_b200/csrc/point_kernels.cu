// point_kernels.cu — launcher of the Point kernels + the run-aggregating instantiations
// (templates in point_kernels_impl.cuh; the non-aggregating ones live in point_kernels_noagg.cu so
// the two halves compile in parallel).
#include "point_kernels_impl.cuh"

namespace pcrb {

cudaError_t point_dispatch_agg(cudaStream_t s, int variant, const uint8_t* mask, const double* x, const double* y,
                               const ChannelPtrs& ch, size_t n, uint32_t* state, const GridParams& g,
                               const PassLayout& L, uint32_t* touched, int sm_count)
{
    return point_impl::dispatch_add<true>(s, variant, mask, x, y, ch, n, state, g, L, touched, sm_count);
}

cudaError_t launch_point_accumulate(cudaStream_t s, int variant, bool warp_aggregate, const uint8_t* mask,
                                    const double* x, const double* y, const ChannelPtrs& ch,
                                    size_t n, uint32_t* state, const GridParams& g,
                                    const PassLayout& L, uint32_t* touched, int sm_count)
{
    if (n == 0) return cudaSuccess;
    return warp_aggregate ? point_dispatch_agg(s, variant, mask, x, y, ch, n, state, g, L, touched, sm_count)
                          : point_dispatch_noagg(s, variant, mask, x, y, ch, n, state, g, L, touched, sm_count);
}

}  // namespace pcrb
