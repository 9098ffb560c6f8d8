// state_kernels.cu — accumulator-record identity fill, and the fused
// merge + finalize kernel, sm_100a.
//
// Reference: init_state_kernel / merge_state_kernel / finalize_kernel templates
// (src/engine/grid_merge.cu:16-45) and the CPU finalize loop the reference
// pipeline actually runs after a D2H of every tile
// (src/engine/pipeline.cpp:1204-1281, src/ops/reduction_registry.cpp:139-156).
// Here it is one streaming pass over the records in HBM: merge the partial
// states of all ranks in rank order (Op::merge), apply Op::finalize per band,
// apply the touched-tile NaN rule, write band-major float32 rasters.
#include "kernels.cuh"

namespace pcrb {

namespace {

constexpr int kThreads = 256;

struct RecordIdentity { uint32_t w[8]; };

template <int W>
__global__ void __launch_bounds__(kThreads)
k_init_state(uint32_t* __restrict__ state, size_t cells, const __grid_constant__ RecordIdentity id)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
    for (size_t c = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; c < cells; c += stride) {
        if constexpr (W == 1) state[c] = id.w[0];
        if constexpr (W == 2) reinterpret_cast<uint2*>(state)[c] = make_uint2(id.w[0], id.w[1]);
        if constexpr (W == 4) reinterpret_cast<uint4*>(state)[c] = make_uint4(id.w[0], id.w[1], id.w[2], id.w[3]);
        if constexpr (W == 8) {
            reinterpret_cast<uint4*>(state)[2 * c]     = make_uint4(id.w[0], id.w[1], id.w[2], id.w[3]);
            reinterpret_cast<uint4*>(state)[2 * c + 1] = make_uint4(id.w[4], id.w[5], id.w[6], id.w[7]);
        }
    }
}

template <int W>
__device__ __forceinline__ void load_record(const uint32_t* __restrict__ base, size_t cell, uint32_t (&r)[8])
{
    if constexpr (W == 1) r[0] = base[cell];
    if constexpr (W == 2) { const uint2 t = reinterpret_cast<const uint2*>(base)[cell]; r[0] = t.x; r[1] = t.y; }
    if constexpr (W == 4) {
        const uint4 t = reinterpret_cast<const uint4*>(base)[cell];
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
    }
    if constexpr (W == 8) {
        const uint4 t = reinterpret_cast<const uint4*>(base)[2 * cell];
        const uint4 u = reinterpret_cast<const uint4*>(base)[2 * cell + 1];
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w; r[4] = u.x; r[5] = u.y; r[6] = u.z; r[7] = u.w;
    }
}

// One thread per cell of [cell0, cell0+count).  parts.part[k] points at the
// record of cell `part_cell0` of rank k's partial state.
template <int W>
__global__ void __launch_bounds__(kThreads)
k_finalize(const __grid_constant__ StateParts parts, size_t part_cell0, size_t cell0, size_t count,
           float* __restrict__ out, size_t band_stride, const __grid_constant__ GridParams g,
           const __grid_constant__ PassLayout L, const __grid_constant__ FinalizeProgram fp,
           const uint32_t* __restrict__ touched)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (i >= count) return;
    const size_t cell = cell0 + i;
    const int row = static_cast<int>(cell / static_cast<size_t>(g.width));
    const int col = static_cast<int>(cell - static_cast<size_t>(row) * g.width);
    const bool live = touched[tile_of(g, col, row)] != 0;   // TileManager::tile_has_state

    uint32_t r[8];
    load_record<W>(parts.part[0], cell - part_cell0, r);
    for (int k = 1; k < parts.n; ++k) {                      // Op::merge, fixed rank order
        uint32_t q[8];
        load_record<W>(parts.part[k], cell - part_cell0, q);
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if (j < L.n_add)
                r[j] = __float_as_uint(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(q[j])));
            else if (j < L.n_add + L.n_max)
                r[j] = static_cast<uint32_t>(max(static_cast<int32_t>(r[j]), static_cast<int32_t>(q[j])));
            else if (j < L.n_add + L.n_max + L.n_min)
                r[j] = static_cast<uint32_t>(min(static_cast<int32_t>(r[j]), static_cast<int32_t>(q[j])));
        }
    }

    const float nan = __int_as_float(0x7fc00000);
    for (int b = 0; b < fp.n; ++b) {
        float o = nan;
        if (live) {
            uint32_t wa = 0, wb = 0;
#pragma unroll
            for (int j = 0; j < W; ++j) {           // register select, no local memory
                if (j == fp.word_a[b]) wa = r[j];
                if (j == fp.word_b[b]) wb = r[j];
            }
            const float a = __uint_as_float(wa);
            switch (fp.kind[b]) {
            case FIN_SUM:   o = a; break;                                           // SumOp::finalize
            case FIN_COUNT: o = a > 0.0f ? a : nan; break;                          // CountOp::finalize
            case FIN_RATIO: { const float d = __uint_as_float(wb);                  // Average / WeightedAverage
                              o = d > 0.0f ? __fdiv_rn(a, d) : nan; } break;
            case FIN_MAX:   { const float m = ordered_f32(static_cast<int32_t>(wa)); // MaxOp::finalize
                              o = (m == -FLT_MAX) ? nan : m; } break;
            case FIN_MIN:   { const float m = ordered_f32(static_cast<int32_t>(wa));
                              o = (m == FLT_MAX) ? nan : m; } break;
            }
        }
        out[static_cast<size_t>(fp.band[b]) * band_stride + cell] = o;
    }
}

}  // namespace

cudaError_t launch_init_state(cudaStream_t s, uint32_t* state, size_t cells, const PassLayout& L)
{
    if (cells == 0) return cudaSuccess;
    if (L.n_max == 0 && L.n_min == 0)
        return cudaMemsetAsync(state, 0, cells * L.width * sizeof(uint32_t), s);
    RecordIdentity id{};
    for (int j = 0; j < 8; ++j) id.w[j] = 0;
    for (int j = 0; j < L.n_max; ++j) id.w[L.n_add + j] = static_cast<uint32_t>(f32_ordered(-FLT_MAX));
    for (int j = 0; j < L.n_min; ++j) id.w[L.n_add + L.n_max + j] = static_cast<uint32_t>(f32_ordered(FLT_MAX));
    size_t blocks = (cells + kThreads - 1) / kThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const unsigned grid = static_cast<unsigned>(blocks);
    switch (L.width) {
    case 1: k_init_state<1><<<grid, kThreads, 0, s>>>(state, cells, id); break;
    case 2: k_init_state<2><<<grid, kThreads, 0, s>>>(state, cells, id); break;
    case 4: k_init_state<4><<<grid, kThreads, 0, s>>>(state, cells, id); break;
    case 8: k_init_state<8><<<grid, kThreads, 0, s>>>(state, cells, id); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_finalize(cudaStream_t s, const StateParts& parts, size_t part_cell0,
                            size_t cell0, size_t count, float* out, size_t band_stride,
                            const GridParams& g, const PassLayout& L, const FinalizeProgram& fp,
                            const uint32_t* touched)
{
    if (count == 0) return cudaSuccess;
    const unsigned grid = static_cast<unsigned>((count + kThreads - 1) / kThreads);
    switch (L.width) {
    case 1: k_finalize<1><<<grid, kThreads, 0, s>>>(parts, part_cell0, cell0, count, out, band_stride, g, L, fp, touched); break;
    case 2: k_finalize<2><<<grid, kThreads, 0, s>>>(parts, part_cell0, cell0, count, out, band_stride, g, L, fp, touched); break;
    case 4: k_finalize<4><<<grid, kThreads, 0, s>>>(parts, part_cell0, cell0, count, out, band_stride, g, L, fp, touched); break;
    case 8: k_finalize<8><<<grid, kThreads, 0, s>>>(parts, part_cell0, cell0, count, out, band_stride, g, L, fp, touched); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace pcrb
