// exact_kernels.cu — deterministic mode 2: Point accumulate and finalize over the exact fixed-point
// state of exact_acc.cuh (the Line / Gaussian kernels of glyph_kernels.cu take the same state).
// Semantics per reducer as include/pcr/ops/builtin_ops.h:10-103 of the reference; only the ORDER
// dependence of the float sums is gone.
#include "exact_acc.cuh"
#include "kernels.cuh"

namespace pcrb {

namespace {

constexpr int kThreads = 256;

template <bool EXACT>
__global__ void __launch_bounds__(kThreads)
k_point_exact(const uint8_t* __restrict__ mask, const double* __restrict__ xs, const double* __restrict__ ys,
              const __grid_constant__ ChannelPtrs ch, size_t n, const __grid_constant__ XAcc xa,
              const __grid_constant__ GridParams g, const __grid_constant__ PassLayout L,
              uint32_t* __restrict__ touched)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (i >= n || (mask != nullptr && mask[i] == 0)) return;
    int col, row;
    if (!route_cell<EXACT>(g, ldg_stream_d(xs + i), ldg_stream_d(ys + i), col, row)) return;
    const size_t cell = static_cast<size_t>(row) * g.width + col;
    float v[kMaxChan];
#pragma unroll
    for (int c = 0; c < kMaxChan; ++c) v[c] = c < L.n_chan ? ldg_stream_f(ch.p[c] + i) : 0.0f;
    auto pick = [&](int src) { return src == 0 ? v[0] : src == 1 ? v[1] : src == 2 ? v[2] : v[3]; };
    for (int j = 0; j < L.n_add; ++j) xacc_add(xa, j, cell, L.add_src[j] < 0 ? 1.0f : pick(L.add_src[j]));
    for (int j = 0; j < L.n_max; ++j) {
        const float m = pick(L.max_src[j]);
        if (m == m) red_max(xa.ext + static_cast<size_t>(j) * xa.cells + cell, f32_ordered(m));
    }
    for (int j = 0; j < L.n_min; ++j) {
        const float m = pick(L.min_src[j]);
        if (m == m) red_min(xa.ext + static_cast<size_t>(L.n_max + j) * xa.cells + cell, f32_ordered(m));
    }
    const int t = tile_of(g, col, row);
    if (touched[t] == 0) touched[t] = 1;
}

__global__ void __launch_bounds__(kThreads)
k_exact_init_ext(int32_t* __restrict__ ext, size_t cells, int n_max, int n_min)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
    const size_t total = cells * static_cast<size_t>(n_max + n_min);
    for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < total; i += stride)
        ext[i] = f32_ordered(i / cells < static_cast<size_t>(n_max) ? -FLT_MAX : FLT_MAX);
}

// one thread per cell: round every additive word once, then Op::finalize per band
__global__ void __launch_bounds__(kThreads)
k_finalize_exact(const __grid_constant__ XAcc xa, size_t cell0, size_t count, float* __restrict__ out, size_t band_stride,
                 const __grid_constant__ GridParams g, const __grid_constant__ PassLayout L,
                 const __grid_constant__ FinalizeProgram fp, const uint32_t* __restrict__ touched)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (i >= count) return;
    const size_t cell = cell0 + i;
    bool live;
    if (g.tiles_x * g.tiles_y == 1) live = touched[0] != 0;
    else {
        const unsigned row = static_cast<unsigned>(cell) / static_cast<unsigned>(g.width);
        const unsigned col = static_cast<unsigned>(cell) - row * static_cast<unsigned>(g.width);
        live = touched[tile_of(g, static_cast<int>(col), static_cast<int>(row))] != 0;
    }
    float addw[kMaxAdd] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < L.n_add; ++j) {
        long long limbs[kXLimbs];
#pragma unroll
        for (int k = 0; k < kXLimbs; ++k) limbs[k] = xa.limbs[(static_cast<size_t>(j) * kXLimbs + k) * xa.cells + cell];
        addw[j] = xacc_round(limbs, xa.flags[static_cast<size_t>(j) * xa.cells + cell]);
    }
    const float nan = __int_as_float(0x7fc00000);
    auto add_word = [&](int w) { return w == 0 ? addw[0] : w == 1 ? addw[1] : w == 2 ? addw[2] : addw[3]; };
    for (int b = 0; b < fp.n; ++b) {
        float o = nan;
        if (live) {
            const int kind = fp.kind[b];
            if (kind == FIN_SUM) o = add_word(fp.word_a[b]);
            else if (kind == FIN_COUNT) { const float a = add_word(fp.word_a[b]); o = a > 0.0f ? a : nan; }
            else if (kind == FIN_RATIO) {
                const float a = add_word(fp.word_a[b]), d = add_word(fp.word_b[b]);
                o = d > 0.0f ? __fdiv_rn(a, d) : nan;
            } else {
                const float m = ordered_f32(xa.ext[static_cast<size_t>(fp.word_a[b] - L.n_add) * xa.cells + cell]);
                o = (m == (kind == FIN_MAX ? -FLT_MAX : FLT_MAX)) ? nan : m;
            }
        }
        out[static_cast<size_t>(fp.band[b]) * band_stride + cell] = o;
    }
}

}  // namespace

size_t xacc_limb_bytes(size_t cells, const PassLayout& L) { return cells * static_cast<size_t>(L.n_add) * kXLimbs * sizeof(long long); }
size_t xacc_flag_bytes(size_t cells, const PassLayout& L) { return cells * static_cast<size_t>(L.n_add) * sizeof(uint32_t); }
size_t xacc_ext_bytes(size_t cells, const PassLayout& L) { return cells * static_cast<size_t>(L.n_max + L.n_min) * sizeof(int32_t); }

cudaError_t launch_exact_init(cudaStream_t s, const XAcc& xa, const PassLayout& L)
{
    cudaError_t e = cudaSuccess;
    if (L.n_add) {
        e = cudaMemsetAsync(xa.limbs, 0, xacc_limb_bytes(xa.cells, L), s);
        if (e == cudaSuccess) e = cudaMemsetAsync(xa.flags, 0, xacc_flag_bytes(xa.cells, L), s);
    }
    if (e == cudaSuccess && L.n_max + L.n_min > 0) {
        k_exact_init_ext<<<148 * 8, kThreads, 0, s>>>(xa.ext, xa.cells, L.n_max, L.n_min);
        e = cudaGetLastError();
    }
    return e;
}

cudaError_t launch_point_exact(cudaStream_t s, const uint8_t* mask, const double* x, const double* y,
                               const ChannelPtrs& ch, size_t n, const XAcc& xa, const GridParams& g,
                               const PassLayout& L, uint32_t* touched)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = static_cast<unsigned>((n + kThreads - 1) / kThreads);
    if (g.exact_x && g.exact_y) k_point_exact<true><<<grid, kThreads, 0, s>>>(mask, x, y, ch, n, xa, g, L, touched);
    else k_point_exact<false><<<grid, kThreads, 0, s>>>(mask, x, y, ch, n, xa, g, L, touched);
    return cudaGetLastError();
}

cudaError_t launch_finalize_exact(cudaStream_t s, const XAcc& xa, size_t cell0, size_t count, float* out,
                                  size_t band_stride, const GridParams& g, const PassLayout& L,
                                  const FinalizeProgram& fp, const uint32_t* touched)
{
    if (count == 0) return cudaSuccess;
    const unsigned grid = static_cast<unsigned>((count + kThreads - 1) / kThreads);
    k_finalize_exact<<<grid, kThreads, 0, s>>>(xa, cell0, count, out, band_stride, g, L, fp, touched);
    return cudaGetLastError();
}

}  // namespace pcrb
