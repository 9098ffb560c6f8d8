// point_kernels_noagg.cu — Point kernels without warp run aggregation (warp_aggregate = 2).
#include "point_kernels_impl.cuh"

namespace pcrb {

cudaError_t point_dispatch_noagg(cudaStream_t s, int variant, const uint8_t* mask, const double* x, const double* y,
                                 const ChannelPtrs& ch, size_t n, uint32_t* state, const GridParams& g,
                                 const PassLayout& L, uint32_t* touched, int sm_count)
{
    return point_impl::dispatch_add<false>(s, variant, mask, x, y, ch, n, state, g, L, touched, sm_count);
}

}  // namespace pcrb
