// point_kernels_impl.cuh — Point glyph: fused route + accumulate, sm_100a.
//
// Replaces the reference chain  kernel_assign -> build_sort_keys -> cub radix sort
// -> 5x apply_perm -> global_to_local -> kernel_accumulate_{sum,count,max,min,average}
// (src/engine/tile_router_kernels.cu:34-132,169-293; accumulator_kernels.cu:31-133)
// with ONE pass over the points for ALL reductions of the pipeline: x,y,value are
// read once (20 B/point), the cell is routed with the CPU rule (common.cuh), and
// every state word of the cell's record is updated by red.global (vector add for
// the additive words, s32 max/min on the ordered-float map).
//
// Two variants:
//   POINT_DIRECT  one thread = kUnroll points, streaming LDGs issued up front.
//   POINT_TMA     persistent CTAs, one elected thread stages point tiles into shared
//                 memory with cp.async.bulk (TMA, UBLKCP) behind mbarriers, consumer
//                 warps route+accumulate from smem.
#pragma once
#include "kernels.cuh"

namespace pcrb {

namespace point_impl {

constexpr int kThreads = 256;
constexpr int kUnroll  = 4;

template <int NADD, int NMAX, int NMIN>
struct RecordShape {
    static constexpr int total = NADD + NMAX + NMIN;
    static constexpr int width = total <= 1 ? 1 : total <= 2 ? 2 : total <= 4 ? 4 : 8;
};

__device__ __forceinline__ float pick(const float (&v)[kMaxChan], int src)
{
    return src == 0 ? v[0] : src == 1 ? v[1] : src == 2 ? v[2] : v[3];
}

// One routed point -> its record.  `cell` is the global cell index.
template <int NADD, int NMAX, int NMIN>
__device__ __forceinline__ void update_record(uint32_t* __restrict__ state, size_t cell,
                                              const float (&add)[kMaxAdd],
                                              const float (&mx)[kMaxExt],
                                              const float (&mn)[kMaxExt])
{
    constexpr int W = RecordShape<NADD, NMAX, NMIN>::width;
    uint32_t* rec = state + cell * W;
    red_add_words<NADD>(reinterpret_cast<float*>(rec), add);
#pragma unroll
    for (int j = 0; j < NMAX; ++j) {
        // fmaxf(acc, NaN) == acc on the CPU path (builtin_ops.h:26): NaN never lands.
        if (mx[j] == mx[j]) red_max(reinterpret_cast<int32_t*>(rec) + NADD + j, f32_ordered(mx[j]));
    }
#pragma unroll
    for (int j = 0; j < NMIN; ++j) {
        if (mn[j] == mn[j]) red_min(reinterpret_cast<int32_t*>(rec) + NADD + NMAX + j, f32_ordered(mn[j]));
    }
}

// Run-length aggregation across the warp: lanes whose points fall in the same
// cell as their left neighbour form a run; the run is reduced with 5 shuffles
// and only its head lane issues the reds.  Spatially ordered clouds (LiDAR scan
// order) otherwise serialise on the L2 atomic unit of one address.
template <int NADD, int NMAX, int NMIN>
__device__ __forceinline__ void aggregate_runs(unsigned heads, int lane, float (&add)[kMaxAdd],
                                               float (&mx)[kMaxExt], float (&mn)[kMaxExt])
{
    const unsigned after = (lane == 31) ? 0u : (heads >> (lane + 1));
    const int run_end = after ? lane + __ffs(after) - 1 : 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const bool take = lane + d <= run_end;
#pragma unroll
        for (int j = 0; j < NADD; ++j) {
            const float t = __shfl_down_sync(0xffffffffu, add[j], d);
            if (take) add[j] += t;
        }
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            const float t = __shfl_down_sync(0xffffffffu, mx[j], d);
            // NaN-ignoring max, like fmaxf
            if (take && !(mx[j] >= t)) mx[j] = (t == t) ? t : mx[j];
        }
#pragma unroll
        for (int j = 0; j < NMIN; ++j) {
            const float t = __shfl_down_sync(0xffffffffu, mn[j], d);
            if (take && !(mn[j] <= t)) mn[j] = (t == t) ? t : mn[j];
        }
    }
}

// Shared tail of both variants: given one point per lane (warp-converged call),
// route it and fold it into the state.
template <int NADD, int NMAX, int NMIN, bool AGG, bool EXACT>
__device__ __forceinline__ void fold_point(const GridParams& g, const PassLayout& L,
                                           uint32_t* __restrict__ state,
                                           uint32_t* __restrict__ touched, bool live, double x,
                                           double y, const float (&v)[kMaxChan], bool& any_valid)
{
    int col = 0, row = 0;
    const bool ok = live && route_cell<EXACT>(g, x, y, col, row);
    const size_t cell = static_cast<size_t>(row) * g.width + col;

    float add[kMaxAdd], mx[kMaxExt], mn[kMaxExt];
#pragma unroll
    for (int j = 0; j < kMaxAdd; ++j) add[j] = (j < NADD) ? (L.add_src[j] < 0 ? 1.0f : pick(v, L.add_src[j])) : 0.0f;
#pragma unroll
    for (int j = 0; j < kMaxExt; ++j) mx[j] = (j < NMAX) ? pick(v, L.max_src[j]) : 0.0f;
#pragma unroll
    for (int j = 0; j < kMaxExt; ++j) mn[j] = (j < NMIN) ? pick(v, L.min_src[j]) : 0.0f;

    bool issue = ok;
    if constexpr (AGG) {
        const int lane = threadIdx.x & 31;
        // invalid lanes get a key no cell can have, distinct from their neighbours'
        const unsigned long long key = ok ? static_cast<unsigned long long>(cell)
                                          : (0x8000000000000000ull | lane);
        const unsigned long long left = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (key != left);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        if (heads != 0xffffffffu) {          // some run is longer than one lane
            aggregate_runs<NADD, NMAX, NMIN>(heads, lane, add, mx, mn);
            issue = ok && head;
        }
    }
    if (issue) update_record<NADD, NMAX, NMIN>(state, cell, add, mx, mn);

    if (ok) {
        any_valid = true;
        if (g.tiles_x * g.tiles_y > 1) {     // touched-tile rule, tile_manager.cpp:437-444
            const int t = tile_of(g, col, row);
            if (touched[t] == 0) touched[t] = 1;
        }
    }
}

// ---------------------------------------------------------------------------
// POINT_DIRECT
// ---------------------------------------------------------------------------
template <int NADD, int NMAX, int NMIN, bool AGG, bool EXACT>
__global__ void __launch_bounds__(kThreads)
k_point_direct(const uint8_t* __restrict__ mask, const double* __restrict__ xs, const double* __restrict__ ys,
               const __grid_constant__ ChannelPtrs ch, size_t n,
               uint32_t* __restrict__ state, const __grid_constant__ GridParams g,
               const __grid_constant__ PassLayout L, uint32_t* __restrict__ touched)
{
    const size_t base = static_cast<size_t>(blockIdx.x) * (kThreads * kUnroll) + threadIdx.x;

    double x[kUnroll], y[kUnroll];
    float  v[kUnroll][kMaxChan];
    bool   live[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const size_t i = base + static_cast<size_t>(u) * kThreads;
        live[u] = i < n && (mask == nullptr || mask[i] != 0);
        x[u] = live[u] ? ldg_stream_d(xs + i) : 0.0;
        y[u] = live[u] ? ldg_stream_d(ys + i) : 0.0;
#pragma unroll
        for (int c = 0; c < kMaxChan; ++c)
            v[u][c] = (live[u] && c < L.n_chan) ? ldg_stream_f(ch.p[c] + i) : 0.0f;
    }

    bool any_valid = false;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
        fold_point<NADD, NMAX, NMIN, AGG, EXACT>(g, L, state, touched, live[u], x[u], y[u], v[u], any_valid);

    if (g.tiles_x * g.tiles_y == 1) {
        if (__any_sync(0xffffffffu, any_valid) && (threadIdx.x & 31) == 0 && touched[0] == 0)
            touched[0] = 1;
    }
}

// ---------------------------------------------------------------------------
// POINT_TMA — persistent, TMA-staged point tiles
// ---------------------------------------------------------------------------
// Shared-memory ring of kStages stages; one stage holds kTile points as SoA
// (x f64 | y f64 | up to kMaxChan f32 channels).  Thread 0 is the producer: it
// waits for a stage to drain (empty mbarrier), arms the full mbarrier with the
// byte count and issues one cp.async.bulk per array.  All kThreads threads are
// consumers: they wait on the full mbarrier, fold kTile/kThreads points each and
// arrive on the empty mbarrier.
constexpr int kTile       = 1024;
constexpr int kStages     = 6;
constexpr int kTmaThreads = 512;

__host__ __device__ constexpr size_t stage_bytes(int n_chan)
{
    return static_cast<size_t>(kTile) * (16 + 4 * n_chan);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// n must be a multiple of 4 (bulk-copy sizes are multiples of 16 B) and all
// arrays 16-byte aligned; the launcher peels the remainder off to POINT_DIRECT.
template <int NADD, int NMAX, int NMIN, bool AGG>
__global__ void __launch_bounds__(kTmaThreads, 1)
k_point_tma(const uint8_t* __restrict__ mask, const double* __restrict__ xs, const double* __restrict__ ys,
            const __grid_constant__ ChannelPtrs ch, size_t n, uint32_t* __restrict__ state,
            const __grid_constant__ GridParams g, const __grid_constant__ PassLayout L,
            uint32_t* __restrict__ touched)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kStages];
    __shared__ uint64_t empty_bar[kStages];

    const size_t n_tiles = (n + kTile - 1) / kTile;
    const int n_chan = L.n_chan;
    const size_t sbytes = stage_bytes(n_chan);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kTmaThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](size_t tile, int s) {
        const size_t p0 = tile * kTile;
        const uint32_t cnt = static_cast<uint32_t>(min(static_cast<size_t>(kTile), n - p0));
        unsigned char* st = smem_raw + s * sbytes;
        mbar_expect_tx(&full_bar[s], cnt * (16u + 4u * n_chan));
        tma_load_1d(st, xs + p0, cnt * 8u, &full_bar[s]);
        tma_load_1d(st + kTile * 8, ys + p0, cnt * 8u, &full_bar[s]);
        for (int c = 0; c < n_chan; ++c)
            tma_load_1d(st + kTile * 16 + c * kTile * 4, ch.p[c] + p0, cnt * 4u, &full_bar[s]);
    };

    // prologue: fill the ring
    if (threadIdx.x == 0) {
        size_t t = blockIdx.x;
        for (int s = 0; s < kStages && t < n_tiles; ++s, t += gridDim.x) issue(t, s);
    }

    bool any_valid = false;
    int stage = 0;
    uint32_t phase = 0;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(&full_bar[stage], phase);
        const size_t p0 = tile * kTile;
        const int cnt = static_cast<int>(min(static_cast<size_t>(kTile), n - p0));
        const unsigned char* st = smem_raw + stage * sbytes;
        const double* sx = reinterpret_cast<const double*>(st);
        const double* sy = reinterpret_cast<const double*>(st + kTile * 8);
        const float*  sv = reinterpret_cast<const float*>(st + kTile * 16);
#pragma unroll
        for (int k = 0; k < kTile / kTmaThreads; ++k) {
            const int i = threadIdx.x + k * kTmaThreads;
            const bool live = i < cnt && (mask == nullptr || mask[p0 + i] != 0);
            float v[kMaxChan];
#pragma unroll
            for (int c = 0; c < kMaxChan; ++c) v[c] = (live && c < n_chan) ? sv[c * kTile + i] : 0.0f;
            const double px = live ? sx[i] : 0.0;
            const double py = live ? sy[i] : 0.0;
            fold_point<NADD, NMAX, NMIN, AGG, false>(g, L, state, touched, live, px, py, v, any_valid);
        }
        mbar_arrive(&empty_bar[stage]);
        if (threadIdx.x == 0) {
            // refill this stage with the tile kStages rounds ahead
            const size_t t2 = tile + static_cast<size_t>(kStages) * gridDim.x;
            if (t2 < n_tiles) {
                mbar_wait(&empty_bar[stage], phase);
                issue(t2, stage);
            }
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
    }

    if (g.tiles_x * g.tiles_y == 1) {
        if (__any_sync(0xffffffffu, any_valid) && (threadIdx.x & 31) == 0 && touched[0] == 0)
            touched[0] = 1;
    }
}

template <int NADD, int NMAX, int NMIN, bool AGG>
cudaError_t launch_shape(cudaStream_t s, int variant, const uint8_t* mask, const double* x, const double* y,
                         const ChannelPtrs& ch, size_t n, uint32_t* state, const GridParams& g,
                         const PassLayout& L, uint32_t* touched, int sm_count)
{
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0 &&
                         [&] { for (int c = 0; c < L.n_chan; ++c)
                                   if (reinterpret_cast<uintptr_t>(ch.p[c]) & 15u) return false;
                               return true; }();
    size_t done = 0;
    if (variant == POINT_TMA && aligned && n >= static_cast<size_t>(kTile)) {
        auto kern = k_point_tma<NADD, NMAX, NMIN, AGG>;
        const size_t smem = stage_bytes(L.n_chan) * kStages;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        const size_t n_main = n & ~static_cast<size_t>(3);
        const size_t n_tiles = (n_main + kTile - 1) / kTile;
        const unsigned grid = static_cast<unsigned>(n_tiles < static_cast<size_t>(sm_count) ? n_tiles : sm_count);
        kern<<<grid, kTmaThreads, smem, s>>>(mask, x, y, ch, n_main, state, g, L, touched);
        done = n_main;
    }
    if (done < n) {
        ChannelPtrs ch2 = ch;
        for (int c = 0; c < L.n_chan; ++c) ch2.p[c] = ch.p[c] + done;
        const size_t rest = n - done;
        const size_t per_block = static_cast<size_t>(kThreads) * kUnroll;
        const unsigned grid = static_cast<unsigned>((rest + per_block - 1) / per_block);
        if (g.exact_x && g.exact_y)
            k_point_direct<NADD, NMAX, NMIN, AGG, true><<<grid, kThreads, 0, s>>>(mask ? mask + done : nullptr, x + done, y + done, ch2, rest, state, g, L, touched);
        else
            k_point_direct<NADD, NMAX, NMIN, AGG, false><<<grid, kThreads, 0, s>>>(mask ? mask + done : nullptr, x + done, y + done, ch2, rest, state, g, L, touched);
    }
    return cudaGetLastError();
}

template <int NADD, int NMAX, bool AGG>
cudaError_t dispatch_min(int n_min, cudaStream_t s, int variant, const uint8_t* mask, const double* x, const double* y,
                         const ChannelPtrs& ch, size_t n, uint32_t* state, const GridParams& g,
                         const PassLayout& L, uint32_t* touched, int sm)
{
    switch (n_min) {
    case 0: if constexpr (NADD + NMAX > 0) return launch_shape<NADD, NMAX, 0, AGG>(s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
            else return cudaErrorInvalidValue;
    case 1: return launch_shape<NADD, NMAX, 1, AGG>(s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
    case 2: return launch_shape<NADD, NMAX, 2, AGG>(s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
    }
    return cudaErrorInvalidValue;
}

template <int NADD, bool AGG>
cudaError_t dispatch_max(int n_max, int n_min, cudaStream_t s, int variant, const uint8_t* mask, const double* x,
                         const double* y, const ChannelPtrs& ch, size_t n, uint32_t* state,
                         const GridParams& g, const PassLayout& L, uint32_t* touched, int sm)
{
    switch (n_max) {
    case 0: return dispatch_min<NADD, 0, AGG>(n_min, s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
    case 1: return dispatch_min<NADD, 1, AGG>(n_min, s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
    case 2: return dispatch_min<NADD, 2, AGG>(n_min, s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
    }
    return cudaErrorInvalidValue;
}

template <bool AGG>
cudaError_t dispatch_add(cudaStream_t s, int variant, const uint8_t* mask, const double* x, const double* y,
                         const ChannelPtrs& ch, size_t n, uint32_t* state, const GridParams& g,
                         const PassLayout& L, uint32_t* touched, int sm)
{
    switch (L.n_add) {
    case 0: return dispatch_max<0, AGG>(L.n_max, L.n_min, s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
    case 1: return dispatch_max<1, AGG>(L.n_max, L.n_min, s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
    case 2: return dispatch_max<2, AGG>(L.n_max, L.n_min, s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
    case 3: return dispatch_max<3, AGG>(L.n_max, L.n_min, s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
    case 4: return dispatch_max<4, AGG>(L.n_max, L.n_min, s, variant, mask, x, y, ch, n, state, g, L, touched, sm);
    }
    return cudaErrorInvalidValue;
}

}  // namespace point_impl

// one translation unit per AGG value (compile time): see point_kernels.cu / point_kernels_noagg.cu
cudaError_t point_dispatch_agg(cudaStream_t s, int variant, const uint8_t* mask, const double* x, const double* y,
                               const ChannelPtrs& ch, size_t n, uint32_t* state, const GridParams& g,
                               const PassLayout& L, uint32_t* touched, int sm_count);
cudaError_t point_dispatch_noagg(cudaStream_t s, int variant, const uint8_t* mask, const double* x, const double* y,
                                 const ChannelPtrs& ch, size_t n, uint32_t* state, const GridParams& g,
                                 const PassLayout& L, uint32_t* touched, int sm_count);

}  // namespace pcrb
