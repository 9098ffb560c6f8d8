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

#include <algorithm>
#include <cstdlib>

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

// ld.global.cg: cache at L2 only.  A part may be PEER memory (NVLink): the L1 must not serve a
// line left over from an earlier finalize of the same addresses.
template <int W>
__device__ __forceinline__ void load_record(const uint32_t* __restrict__ base, size_t cell, uint32_t (&r)[8])
{
    if constexpr (W == 1) r[0] = __ldcg(base + cell);
    if constexpr (W == 2) { const uint2 t = __ldcg(reinterpret_cast<const uint2*>(base) + cell); r[0] = t.x; r[1] = t.y; }
    if constexpr (W == 4) {
        const uint4 t = __ldcg(reinterpret_cast<const uint4*>(base) + cell);
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
    }
    if constexpr (W == 8) {
        const uint4 t = __ldcg(reinterpret_cast<const uint4*>(base) + 2 * cell);
        const uint4 u = __ldcg(reinterpret_cast<const uint4*>(base) + 2 * cell + 1);
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w; r[4] = u.x; r[5] = u.y; r[6] = u.z; r[7] = u.w;
    }
}

template <int W>
__device__ __forceinline__ void store_record(uint32_t* __restrict__ base, size_t cell, const uint32_t (&r)[8])
{
    if constexpr (W == 1) base[cell] = r[0];
    if constexpr (W == 2) reinterpret_cast<uint2*>(base)[cell] = make_uint2(r[0], r[1]);
    if constexpr (W == 4) reinterpret_cast<uint4*>(base)[cell] = make_uint4(r[0], r[1], r[2], r[3]);
    if constexpr (W == 8) {
        reinterpret_cast<uint4*>(base)[2 * cell]     = make_uint4(r[0], r[1], r[2], r[3]);
        reinterpret_cast<uint4*>(base)[2 * cell + 1] = make_uint4(r[4], r[5], r[6], r[7]);
    }
}

// One thread per cell of [cell0, cell0+count).  parts.part[k] points at the
// record of cell `part_cell0` of rank k's partial state.
// Merged record of one cell -> all bands of the pass.  The record words are parked in shared memory
// ([word][thread], conflict-free) so that a band picks its words with one LDS at a warp-uniform row
// instead of a chain of register selects; the band program itself is warp-uniform.
template <int W>
__device__ __forceinline__ void finalize_cell(const StateParts& parts, size_t part_cell0, size_t cell,
                                              const OutTargets& outs, size_t band_stride,
                                              const PassLayout& L, const FinalizeProgram& fp, bool live,
                                              uint32_t (*s_words)[kThreads], uint32_t* accum = nullptr)
{
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
    if (accum) store_record<W>(accum, cell - part_cell0, r);      // the owner keeps the running merge
#pragma unroll
    for (int j = 0; j < W; ++j) s_words[j][threadIdx.x] = r[j];   // own column only: no sync needed

    const float nan = __int_as_float(0x7fc00000);
    for (int b = 0; b < fp.n; ++b) {
        float o = nan;
        if (live) {
            const int kind = fp.kind[b];
            const uint32_t wa = s_words[fp.word_a[b]][threadIdx.x];
            const float a = __uint_as_float(wa);
            if (kind == FIN_SUM) {                                               // SumOp::finalize
                o = a;
            } else if (kind == FIN_COUNT) {                                      // CountOp::finalize
                o = a > 0.0f ? a : nan;
            } else if (kind == FIN_RATIO) {                                      // Average / WeightedAverage
                const float d = __uint_as_float(s_words[fp.word_b[b]][threadIdx.x]);
                o = d > 0.0f ? __fdiv_rn(a, d) : nan;
            } else {                                                             // MaxOp / MinOp::finalize
                const float m = ordered_f32(static_cast<int32_t>(wa));
                o = (m == (kind == FIN_MAX ? -FLT_MAX : FLT_MAX)) ? nan : m;
            }
        }
        const size_t at = static_cast<size_t>(fp.band[b]) * band_stride + cell;
        for (int t = 0; t < outs.n; ++t) outs.out[t][at] = o;
    }
}

template <int W>
__global__ void __launch_bounds__(kThreads)
k_finalize(const __grid_constant__ StateParts parts, size_t part_cell0, size_t cell0, size_t count,
           const __grid_constant__ OutTargets outs, size_t band_stride, const __grid_constant__ GridParams g,
           const __grid_constant__ PassLayout L, const __grid_constant__ FinalizeProgram fp,
           const uint32_t* __restrict__ touched)
{
    __shared__ uint32_t s_words[W][kThreads];
    const size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (i >= count) return;
    const size_t cell = cell0 + i;
    bool live;                                               // TileManager::tile_has_state
    if (g.tiles_x * g.tiles_y == 1) {
        live = touched[0] != 0;
    } else {                                                 // cells < 2^32 (engine check): 32-bit divides
        const unsigned row = static_cast<unsigned>(cell) / static_cast<unsigned>(g.width);
        const unsigned col = static_cast<unsigned>(cell) - row * static_cast<unsigned>(g.width);
        live = touched[tile_of(g, static_cast<int>(col), static_cast<int>(row))] != 0;
    }
    finalize_cell<W>(parts, part_cell0, cell, outs, band_stride, L, fp, live, s_words);
}

// Single-GPU finalize, four consecutive cells per thread: one 128-bit store per band instead of
// four 32-bit ones (the kernel is store-issue bound: W record words in, one word per band out).
// Requires cell0, count and band_stride to be multiples of 4.
template <int W>
__global__ void __launch_bounds__(kThreads)
k_finalize_v4(const uint32_t* __restrict__ part, size_t part_cell0, size_t cell0, size_t groups,
              float* __restrict__ out, size_t band_stride, const __grid_constant__ GridParams g,
              const __grid_constant__ PassLayout L, const __grid_constant__ FinalizeProgram fp,
              const uint32_t* __restrict__ touched)
{
    __shared__ uint32_t s_words[W * 4][kThreads];          // [word][cell of the group][thread]
    const size_t gi = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (gi >= groups) return;
    const size_t cell = cell0 + 4 * gi;
    bool live[4];
    if (g.tiles_x * g.tiles_y == 1) {
        live[0] = live[1] = live[2] = live[3] = touched[0] != 0;
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const unsigned at = static_cast<unsigned>(cell) + c;
            const unsigned row = at / static_cast<unsigned>(g.width);
            const unsigned col = at - row * static_cast<unsigned>(g.width);
            live[c] = touched[tile_of(g, static_cast<int>(col), static_cast<int>(row))] != 0;
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t r[8];
        load_record<W>(part, cell + c - part_cell0, r);
#pragma unroll
        for (int j = 0; j < W; ++j) s_words[j * 4 + c][threadIdx.x] = r[j];   // own column only: no sync needed
    }
    const float nan = __int_as_float(0x7fc00000);
    for (int b = 0; b < fp.n; ++b) {
        const int kind = fp.kind[b];
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            o[c] = nan;
            if (!live[c]) continue;
            const uint32_t wa = s_words[fp.word_a[b] * 4 + c][threadIdx.x];
            const float a = __uint_as_float(wa);
            if (kind == FIN_SUM) {
                o[c] = a;
            } else if (kind == FIN_COUNT) {
                o[c] = a > 0.0f ? a : nan;
            } else if (kind == FIN_RATIO) {
                const float d = __uint_as_float(s_words[fp.word_b[b] * 4 + c][threadIdx.x]);
                o[c] = d > 0.0f ? __fdiv_rn(a, d) : nan;
            } else {
                const float m = ordered_f32(static_cast<int32_t>(wa));
                o[c] = (m == (kind == FIN_MAX ? -FLT_MAX : FLT_MAX)) ? nan : m;
            }
        }
        *reinterpret_cast<float4*>(out + static_cast<size_t>(fp.band[b]) * band_stride + cell) =
            make_float4(o[0], o[1], o[2], o[3]);
    }
}

// Point filter -> byte mask (evaluate_predicate, src/engine/filter.cpp:37-58; predicates are AND-ed)
__global__ void __launch_bounds__(kThreads)
k_filter_mask(const __grid_constant__ FilterProgram fp, size_t n, uint8_t* __restrict__ mask,
              unsigned long long* __restrict__ survivors)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    bool keep = i < n;
    for (int k = 0; k < fp.n && keep; ++k) {
        const float v = fp.chan[k][i];
        bool in_set = false;
        switch (fp.op[k]) {
        case 0: keep = v == fp.value[k]; break;
        case 1: keep = v != fp.value[k]; break;
        case 2: keep = v <  fp.value[k]; break;
        case 3: keep = v <= fp.value[k]; break;
        case 4: keep = v >  fp.value[k]; break;
        case 5: keep = v >= fp.value[k]; break;
        case 6:
        case 7:
            for (int j = 0; j < fp.set_size[k]; ++j) in_set = in_set || (fp.set[k][j] == v);
            keep = (fp.op[k] == 6) ? in_set : !in_set;
            break;
        default: keep = false;
        }
    }
    if (i < n) mask[i] = keep ? 1 : 0;
    const unsigned kept = __popc(__ballot_sync(0xffffffffu, keep));
    if ((threadIdx.x & 31) == 0 && kept) atomicAdd(survivors, static_cast<unsigned long long>(kept));
}

// ---- peer flags: system-scope release/acquire ----
__device__ __forceinline__ void wait_flag(const uint32_t* slot, uint32_t epoch)
{
    uint32_t v;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
    } while (static_cast<int32_t>(v - epoch) < 0);
}

// Peer-memory finalize: wait + touched-tile OR + merge + finalize + completion signal in ONE
// launch.  Every CTA waits (polling this rank's own flag array) until all ranks' pushed slices
// have landed in the local combine buffer, merges them with the local partial state in rank
// order, finalizes, and stores the bands into this rank's and the peers' arrays (posted NVLink
// writes); the last CTA to finish releases the "done" flag on every rank.

// Push: thread per cell of the WHOLE grid; every cell's record is copied to the rank that owns its row
// (NVLink stores for foreign slices, a local copy for this rank's own slice).
template <int W>
__global__ void __launch_bounds__(kThreads)
k_push_slices(uint32_t* __restrict__ state, uint32_t* __restrict__ touched, int n_tiles,
              const __grid_constant__ GridParams g, const __grid_constant__ PushTargets pt,
              const __grid_constant__ PeerSync ps, int push_touched, int signal, int reset,
              const __grid_constant__ RecordIdentity id)
{
    const PeerFlags& pf = ps.pf;
    // The peers' combine buffers may still be being read by their previous finalize: wait for
    // their "done" flags of the previous epoch (phase 1) before overwriting them.  Nothing else
    // of the previous finalize is waited for here, so the ingest in between ran unhindered.
    if (ps.epoch > 1 && !ps.waited) {
        if (threadIdx.x < pf.n) wait_flag(pf.flags[pf.rank] + kMaxParts + threadIdx.x, ps.epoch - 1);
        __syncthreads();
    }
    // Persistent grid-stride loop (cells < 2^32, engine check): one system fence per CTA at the end
    // instead of one per 256 cells, and no CTA turnover while the posted NVLink stores drain.
    const unsigned cells = static_cast<unsigned>(g.width) * static_cast<unsigned>(g.height);
    const unsigned per_owner = static_cast<unsigned>(pt.rows_per) * static_cast<unsigned>(g.width);
    const unsigned stride = gridDim.x * kThreads;
    for (unsigned cell = blockIdx.x * kThreads + threadIdx.x; cell < cells; cell += stride) {
        // every cell's record goes to the rank that owns its row: NVLink stores for foreign slices, a
        // local copy for my own (after this kernel nothing of this finalize reads the live state, so
        // the next ingest may overlap the merge)
        const unsigned owner = cell / per_owner;
        const unsigned local = cell - owner * per_owner;
        uint32_t* dst = pt.combined[owner] + (static_cast<size_t>(pf.rank) * pt.max_slice_cells + local) * W;
        uint32_t* src = state + static_cast<size_t>(cell) * W;
        if constexpr (W == 1) dst[0] = src[0];
        if constexpr (W == 2) *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
        if constexpr (W == 4) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
        if constexpr (W == 8) {
            reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(src)[0];
            reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(src)[1];
        }
        if (reset) store_record<W>(state, cell, id.w);   // delta buffer: back to the identity behind the copy
    }
    if (push_touched && blockIdx.x == 0) {
        for (int t = threadIdx.x; t < n_tiles; t += kThreads) {
            const uint32_t v = touched[t];
            for (int k = 0; k < pf.n; ++k) pt.touched_stage[k][static_cast<size_t>(pf.rank) * n_tiles + t] = v;
            if (reset) touched[t] = 0;
        }
    }
    if (signal) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned int done = atomicAdd(ps.done_counter, 1u);
            if (done == gridDim.x - 1) {
                *ps.done_counter = 0;
                for (int k = 0; k < pf.n; ++k) {
                    uint32_t* slot = pf.flags[k] + pf.rank;                       // phase 0
                    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(slot), "r"(ps.epoch) : "memory");
                }
            }
        }
    }
}

template <int W>
__global__ void __launch_bounds__(kThreads)
k_finalize_peer(const __grid_constant__ StateParts parts, size_t part_cell0, size_t cell0, size_t count,
                const __grid_constant__ OutTargets outs, size_t band_stride,
                const __grid_constant__ GridParams g, const __grid_constant__ PassLayout L,
                const __grid_constant__ FinalizeProgram fp, const __grid_constant__ PeerSync ps,
                uint32_t* __restrict__ accum)
{
    __shared__ uint32_t s_words[W][kThreads];
    const PeerFlags& pf = ps.pf;
    if (!ps.waited) {
        if (threadIdx.x < pf.n) wait_flag(pf.flags[pf.rank] + threadIdx.x, ps.epoch);
        __syncthreads();
    }

    // persistent grid-stride loop: one system fence per CTA behind all of its remote band stores
    const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
    for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < count; i += stride) {
        const size_t cell = cell0 + i;
        const unsigned row = static_cast<unsigned>(cell) / static_cast<unsigned>(g.width);
        const unsigned col = static_cast<unsigned>(cell) - row * static_cast<unsigned>(g.width);
        const int t = (g.tiles_x * g.tiles_y == 1) ? 0 : tile_of(g, static_cast<int>(col), static_cast<int>(row));
        uint32_t live = 0;
        for (int k = 0; k < ps.pt.n; ++k) live |= __ldcg(ps.pt.touched[k] + t);
        finalize_cell<W>(parts, part_cell0, cell, outs, band_stride, L, fp, live != 0, s_words, accum);
    }

    if (ps.signal_end) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            const unsigned int done = atomicAdd(ps.done_counter, 1u);
            if (done == gridDim.x - 1) {
                *ps.done_counter = 0;
                for (int k = 0; k < pf.n; ++k) {
                    uint32_t* slot = pf.flags[k] + kMaxParts + pf.rank;       // phase 1
                    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(slot), "r"(ps.epoch) : "memory");
                }
            }
        }
    }
}

__global__ void k_peer_signal(const __grid_constant__ PeerFlags pf, int phase, uint32_t epoch)
{
    const int k = threadIdx.x;
    if (k >= pf.n) return;
    __threadfence_system();          // everything this stream did so far is visible before the flag
    uint32_t* slot = pf.flags[k] + phase * kMaxParts + pf.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(slot), "r"(epoch) : "memory");
}

__global__ void k_push_touched(const uint32_t* __restrict__ touched, int n_tiles, const __grid_constant__ PushTargets pt,
                               const __grid_constant__ PeerFlags pf, uint32_t epoch, const __grid_constant__ PartCounts pc)
{
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
        const uint32_t v = touched[t];
        for (int k = 0; k < pf.n; ++k) pt.touched_stage[k][static_cast<size_t>(pf.rank) * n_tiles + t] = v;
    }
    // pages I used in my slice of every owner's pool -> that owner; my counters start the next epoch from zero
    for (int j = threadIdx.x; j < pc.n_pass * pf.n; j += blockDim.x) {
        const int i = j / pf.n, k = j - i * pf.n;
        uint32_t* c = pc.local[i] + k * 4;
        *pc.owner_count[i][k] = c[0];
        if (c[1]) *pc.my_overflow[i] = 1u;
        c[0] = 0; c[1] = 0;
    }
    __syncthreads();
    if (threadIdx.x < pf.n) {
        __threadfence_system();      // the stores above, and everything earlier kernels of this stream posted
        uint32_t* slot = pf.flags[threadIdx.x] + pf.rank;                     // phase 0
        asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(slot), "r"(epoch) : "memory");
    }
}

__global__ void k_peer_wait(const __grid_constant__ PeerFlags pf, int phase, uint32_t epoch)
{
    const int k = threadIdx.x;
    if (k < pf.n) wait_flag(pf.flags[pf.rank] + phase * kMaxParts + k, epoch);
}

__global__ void k_peer_wait_merge_touched(const __grid_constant__ PeerFlags pf, uint32_t epoch,
                                          const __grid_constant__ PeerTouched pt,
                                          uint32_t* __restrict__ merged, int n_tiles, int cumulative)
{
    if (threadIdx.x < pf.n) wait_flag(pf.flags[pf.rank] + threadIdx.x, epoch);
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
        uint32_t v = cumulative ? merged[t] : 0u;
        for (int k = 0; k < pt.n; ++k) v |= __ldcg(pt.touched[k] + t);
        merged[t] = v;
    }
}

}  // namespace

static RecordIdentity identity_of(const PassLayout& L)
{
    RecordIdentity id{};
    for (int j = 0; j < 8; ++j) id.w[j] = 0;
    for (int j = 0; j < L.n_max; ++j) id.w[L.n_add + j] = static_cast<uint32_t>(f32_ordered(-FLT_MAX));
    for (int j = 0; j < L.n_min; ++j) id.w[L.n_add + L.n_max + j] = static_cast<uint32_t>(f32_ordered(FLT_MAX));
    return id;
}

cudaError_t launch_init_state(cudaStream_t s, uint32_t* state, size_t cells, const PassLayout& L)
{
    if (cells == 0) return cudaSuccess;
    if (L.n_max == 0 && L.n_min == 0)
        return cudaMemsetAsync(state, 0, cells * L.width * sizeof(uint32_t), s);
    const RecordIdentity id = identity_of(L);
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
                            size_t cell0, size_t count, const OutTargets& out, size_t band_stride,
                            const GridParams& g, const PassLayout& L, const FinalizeProgram& fp,
                            const uint32_t* touched)
{
    if (count == 0) return cudaSuccess;
    if (parts.n == 1 && out.n == 1 && cell0 % 4 == 0 && count % 4 == 0 && band_stride % 4 == 0 &&
        (cell0 - part_cell0) % 4 == 0 && L.width <= 4) {
        const size_t groups = count / 4;
        const unsigned vgrid = static_cast<unsigned>((groups + kThreads - 1) / kThreads);
        switch (L.width) {
        case 1: k_finalize_v4<1><<<vgrid, kThreads, 0, s>>>(parts.part[0], part_cell0, cell0, groups, out.out[0], band_stride, g, L, fp, touched); break;
        case 2: k_finalize_v4<2><<<vgrid, kThreads, 0, s>>>(parts.part[0], part_cell0, cell0, groups, out.out[0], band_stride, g, L, fp, touched); break;
        case 4: k_finalize_v4<4><<<vgrid, kThreads, 0, s>>>(parts.part[0], part_cell0, cell0, groups, out.out[0], band_stride, g, L, fp, touched); break;
        default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
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

namespace pcrb {

cudaError_t launch_filter_mask(cudaStream_t s, const FilterProgram& fp, size_t n, uint8_t* mask,
                               unsigned long long* survivors)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = static_cast<unsigned>((n + kThreads - 1) / kThreads);
    k_filter_mask<<<grid, kThreads, 0, s>>>(fp, n, mask, survivors);
    return cudaGetLastError();
}

// CTAs per SM of the push and merge kernels.  They run beside the ingest kernel of the next step: a grid
// that fills every SM slot (8 per SM) makes the three kernels take turns instead of overlapping (N=2, config 2:
// 88.8 us per step with 8, 85-86 us with 2-4, 116 us with 1 where the push itself becomes the bottleneck).
static int comm_ctas_per_sm()
{
    static const int v = [] { const char* e = std::getenv("PCR_COMM_CTAS_PER_SM"); const int k = e ? std::atoi(e) : 0; return k > 0 ? k : 3; }();
    return v;
}

cudaError_t launch_push_slices(cudaStream_t s, uint32_t* state, uint32_t* touched, int n_tiles,
                               const GridParams& g, const PassLayout& L, const PushTargets& pt, const PeerSync& ps,
                               bool push_touched, bool signal, bool reset, int sm_count)
{
    const RecordIdentity id = identity_of(L);
    const size_t cells = static_cast<size_t>(g.width) * g.height;
    const unsigned grid = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>((cells + kThreads - 1) / kThreads,
                                                                                     static_cast<size_t>(sm_count) * comm_ctas_per_sm())));
    switch (L.width) {
    case 1: k_push_slices<1><<<grid, kThreads, 0, s>>>(state, touched, n_tiles, g, pt, ps, push_touched, signal, reset, id); break;
    case 2: k_push_slices<2><<<grid, kThreads, 0, s>>>(state, touched, n_tiles, g, pt, ps, push_touched, signal, reset, id); break;
    case 4: k_push_slices<4><<<grid, kThreads, 0, s>>>(state, touched, n_tiles, g, pt, ps, push_touched, signal, reset, id); break;
    case 8: k_push_slices<8><<<grid, kThreads, 0, s>>>(state, touched, n_tiles, g, pt, ps, push_touched, signal, reset, id); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_finalize_peer(cudaStream_t s, const StateParts& parts, size_t part_cell0, size_t cell0,
                                 size_t count, const OutTargets& out, size_t band_stride,
                                 const GridParams& g, const PassLayout& L, const FinalizeProgram& fp,
                                 const PeerSync& ps, uint32_t* accum, int sm_count)
{
    // count may be 0 (a rank that owns no rows): the handshake still has to happen
    const unsigned grid = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>((count + kThreads - 1) / kThreads,
                                                                                     static_cast<size_t>(sm_count) * comm_ctas_per_sm())));
    switch (L.width) {
    case 1: k_finalize_peer<1><<<grid, kThreads, 0, s>>>(parts, part_cell0, cell0, count, out, band_stride, g, L, fp, ps, accum); break;
    case 2: k_finalize_peer<2><<<grid, kThreads, 0, s>>>(parts, part_cell0, cell0, count, out, band_stride, g, L, fp, ps, accum); break;
    case 4: k_finalize_peer<4><<<grid, kThreads, 0, s>>>(parts, part_cell0, cell0, count, out, band_stride, g, L, fp, ps, accum); break;
    case 8: k_finalize_peer<8><<<grid, kThreads, 0, s>>>(parts, part_cell0, cell0, count, out, band_stride, g, L, fp, ps, accum); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_push_touched(cudaStream_t s, const uint32_t* touched, int n_tiles, const PushTargets& pt,
                                const PeerFlags& pf, uint32_t epoch, const PartCounts& pc)
{
    k_push_touched<<<1, 256, 0, s>>>(touched, n_tiles, pt, pf, epoch, pc);
    return cudaGetLastError();
}

cudaError_t launch_peer_signal(cudaStream_t s, const PeerFlags& pf, int phase, uint32_t epoch)
{
    k_peer_signal<<<1, 32, 0, s>>>(pf, phase, epoch);
    return cudaGetLastError();
}

cudaError_t launch_peer_wait(cudaStream_t s, const PeerFlags& pf, int phase, uint32_t epoch)
{
    k_peer_wait<<<1, 32, 0, s>>>(pf, phase, epoch);
    return cudaGetLastError();
}

cudaError_t launch_peer_wait_merge_touched(cudaStream_t s, const PeerFlags& pf, uint32_t epoch,
                                           const PeerTouched& pt, uint32_t* merged, int n_tiles, bool cumulative)
{
    k_peer_wait_merge_touched<<<1, 256, 0, s>>>(pf, epoch, pt, merged, n_tiles, cumulative);
    return cudaGetLastError();
}

}  // namespace pcrb
