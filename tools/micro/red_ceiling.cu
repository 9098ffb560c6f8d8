// Microbenchmark (measurement aid, not product): what does the hardware allow for the accumulate
// step of the Point glyph, independent of our kernel?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a red_ceiling.cu -o red_ceiling && ./red_ceiling
// All variants process P "points" whose cell is a hash of the point index (no coordinate loads unless
// stated), records are 16 B (the bench's [sum, count, max, pad]).
//   red_v2_max      red.add.v2.f32 + red.max.s32 per point       (what k_point_direct issues, config 2)
//   red_v2          red.add.v2.f32 only                          (Point Average)
//   red_v4          red.add.v4.f32 only
//   red_max         red.max.s32 only
//   loads_only      3 streaming loads per point (x, y f64, v f32), no REDs   (the HBM side alone)
//   loads+red       both (cell still hashed, so the loads only add their traffic)
//   big_uniform     red_v2_max over a 6.4 GB record array (20000^2 grid), uniform cells
//   big_window<k>   same array, but consecutive runs of points stay inside a window of 2^k cells
//                   (what a coarse binning pass in front of the kernel would produce)
//   smem_atomics    tile-local accumulation: 2 atomicAdd(float) + 1 atomicMax(int) per point into a
//                   4096-cell shared-memory tile per CTA (the "shared-memory tile accumulation" variant)
//   match_any       __match_any_sync throughput (the ranking primitive of an atomic-free smem binning)
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ void red_add2(float* p, float a, float b)
{
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d)
{
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_max(int* p, int v)
{
    asm volatile("red.relaxed.gpu.global.max.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ldg_d(const double* p)
{
    double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}
__device__ __forceinline__ float ldg_f(const float* p)
{
    float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}

constexpr int kThreads = 256, kUnroll = 4;

// MODE bits: 1 = v2 add, 2 = max, 4 = v4 add, 8 = loads
template <int MODE>
__global__ void __launch_bounds__(kThreads)
k_red(uint32_t* __restrict__ state, size_t cells, size_t n, const double* __restrict__ xs,
      const double* __restrict__ ys, const float* __restrict__ vs, int window_log2, size_t pts_per_window,
      float* __restrict__ sink)
{
    const size_t base = (size_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
    double x[kUnroll], y[kUnroll]; float v[kUnroll];
    float keep = 0.f;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const size_t i = base + (size_t)u * kThreads;
        if constexpr (MODE & 8) {
            x[u] = i < n ? ldg_d(xs + i) : 0.0; y[u] = i < n ? ldg_d(ys + i) : 0.0; v[u] = i < n ? ldg_f(vs + i) : 0.f;
        } else { x[u] = 0; y[u] = 0; v[u] = 0.5f; }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const size_t i = base + (size_t)u * kThreads;
        if (i >= n) continue;
        const uint32_t h = hash32((uint32_t)i * 2654435761u + 12345u);
        size_t cell;
        if (window_log2 > 0) {
            const size_t win = i / pts_per_window;                       // consecutive points share a window
            const size_t nwin = cells >> window_log2;
            cell = ((size_t)(hash32((uint32_t)win) % nwin) << window_log2) + (h & ((1u << window_log2) - 1));
        } else cell = h % cells;
        uint32_t* rec = state + cell * 4;
        float val = v[u];
        if constexpr (MODE & 8) keep += (float)(x[u] + y[u]);
        if constexpr (MODE & 1) red_add2((float*)rec, val, 1.0f);
        if constexpr (MODE & 4) red_add4((float*)rec, val, 1.0f, val, 1.0f);
        if constexpr (MODE & 2) red_max((int*)rec + 2, __float_as_int(val) + (int)(h & 1023));
    }
    if constexpr (MODE & 8) if (keep == 1.2345f) *sink = keep;
}

// shared-memory tile accumulation: every CTA owns a 4096-cell tile (sum, count, max = 48 KB) and folds
// `per_cta` points whose cells are random inside the tile; flush with coalesced plain stores.
__global__ void __launch_bounds__(kThreads)
k_smem_atomics(uint32_t* __restrict__ state, size_t n_per_cta, int flush)
{
    __shared__ float s_sum[4096]; __shared__ float s_cnt[4096]; __shared__ int s_max[4096];
    for (int c = threadIdx.x; c < 4096; c += kThreads) { s_sum[c] = 0.f; s_cnt[c] = 0.f; s_max[c] = -2147483647; }
    __syncthreads();
    for (size_t i = threadIdx.x; i < n_per_cta; i += kThreads) {
        const uint32_t h = hash32((uint32_t)(i + blockIdx.x * n_per_cta) * 2654435761u + 777u);
        const int c = h & 4095;
        const float val = (h >> 12) * (1.0f / 1048576.0f);
        atomicAdd(&s_sum[c], val);
        atomicAdd(&s_cnt[c], 1.0f);
        atomicMax(&s_max[c], __float_as_int(val));
    }
    __syncthreads();
    if (flush) {
        uint4* out = (uint4*)state + (size_t)blockIdx.x * 4096;
        for (int c = threadIdx.x; c < 4096; c += kThreads)
            out[c] = make_uint4(__float_as_uint(s_sum[c]), __float_as_uint(s_cnt[c]), (uint32_t)s_max[c], 0u);
    }
}

// same tile accumulation without float atomics: count via integer ATOMS (returns the rank),
// values scattered to a per-cell slot list... approximated here by one integer atomicAdd + 2 STS per
// point (the cost structure of a counting-sort based fold)
__global__ void __launch_bounds__(kThreads)
k_smem_int_rank(uint32_t* __restrict__ state, size_t n_per_cta)
{
    __shared__ int s_cnt[4096]; __shared__ float s_slots[4096 * 2];
    for (int c = threadIdx.x; c < 4096; c += kThreads) s_cnt[c] = 0;
    __syncthreads();
    for (size_t i = threadIdx.x; i < n_per_cta; i += kThreads) {
        const uint32_t h = hash32((uint32_t)(i + blockIdx.x * n_per_cta) * 2654435761u + 777u);
        const int c = h & 4095;
        const float val = (h >> 12) * (1.0f / 1048576.0f);
        const int r = atomicAdd(&s_cnt[c], 1);
        s_slots[(c * 2 + (r & 1))] = val;
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt[5] == -1) state[0] = __float_as_uint(s_slots[3]);
}

__global__ void __launch_bounds__(kThreads)
k_match_any(uint32_t* __restrict__ out, int iters)
{
    uint32_t acc = 0;
    uint32_t h = hash32(threadIdx.x + blockIdx.x * kThreads);
    for (int i = 0; i < iters; ++i) {
        h = h * 1664525u + 1013904223u;
        acc += __match_any_sync(0xffffffffu, h >> 24);      // 256 buckets
    }
    if (acc == 0x12345u) out[0] = acc;
}

template <typename F>
static float time_ms(F&& launch, int reps = 20)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    std::vector<float> t;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); t.push_back(ms);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

int main(int argc, char** argv)
{
    const size_t P = 5'000'000, cells = 1'000'000;
    uint32_t* state; CK(cudaMalloc(&state, cells * 16)); CK(cudaMemset(state, 0, cells * 16));
    // 4 rotated point sets so that the loads always come from HBM (4 x 100 MB > L2)
    double *xs[4], *ys[4]; float* vs[4];
    for (int r = 0; r < 4; ++r) {
        CK(cudaMalloc(&xs[r], P * 8)); CK(cudaMalloc(&ys[r], P * 8)); CK(cudaMalloc(&vs[r], P * 4));
        CK(cudaMemset(xs[r], 0, P * 8)); CK(cudaMemset(ys[r], 0, P * 8)); CK(cudaMemset(vs[r], 0, P * 4));
    }
    float* sink; CK(cudaMalloc(&sink, 4));
    const unsigned grid = (unsigned)((P + kThreads * kUnroll - 1) / (kThreads * kUnroll));
    int rot = 0;
    printf("{\"points\": %zu, \"cells\": %zu, \"record_bytes\": 16,\n", P, cells);
    auto report = [&](const char* name, float ms, size_t pts, double reds_per_pt) {
        printf(" \"%s\": {\"us\": %.2f, \"gpts_per_s\": %.2f, \"greds_per_s\": %.2f, \"alg_gbs_at_20B\": %.1f},\n", name, ms * 1e3,
               pts / (ms * 1e-3) / 1e9, pts * reds_per_pt / (ms * 1e-3) / 1e9, pts * 20.0 / (ms * 1e-3) / 1e9);
    };
#define RUN(MODE, name, reds) report(name, time_ms([&] { k_red<MODE><<<grid, kThreads>>>(state, cells, P, xs[rot & 3], ys[rot & 3], vs[rot & 3], 0, 1, sink); ++rot; }), P, reds)
    RUN(3, "red_v2_max", 2);
    RUN(1, "red_v2", 1);
    RUN(4, "red_v4", 1);
    RUN(2, "red_max", 1);
    RUN(8, "loads_only", 0);
    RUN(11, "loads+red_v2_max", 2);
    RUN(9, "loads+red_v2", 1);
    CK(cudaFree(state));

    // the 20000^2 grid: 400M records = 6.4 GB
    const size_t big_cells = 400'000'000, BP = 50'000'000;
    CK(cudaMalloc(&state, big_cells * 16)); CK(cudaMemset(state, 0, big_cells * 16));
    const unsigned bgrid = (unsigned)((BP + kThreads * kUnroll - 1) / (kThreads * kUnroll));
    report("big_uniform", time_ms([&] { k_red<3><<<bgrid, kThreads>>>(state, big_cells, BP, nullptr, nullptr, nullptr, 0, 1, sink); }, 5), BP, 2);
    for (int wl : {24, 22, 21, 20, 19}) {
        // points per window = the share a uniform cloud of BP points would put there
        const size_t ppw = std::max<size_t>(1, (size_t)((double)BP * (double)(1u << wl) / (double)big_cells));
        char nm[64]; snprintf(nm, sizeof nm, "big_window_2^%d_cells(%zu MB)_ppw%zu", wl, ((size_t)16 << wl) >> 20, ppw);
        report(nm, time_ms([&] { k_red<3><<<bgrid, kThreads>>>(state, big_cells, BP, nullptr, nullptr, nullptr, wl, ppw, sink); }, 5), BP, 2);
        const size_t ppw8 = ppw * 8;   // denser cloud: 8x the points per window visit
        snprintf(nm, sizeof nm, "big_window_2^%d_cells_ppw%zu", wl, ppw8);
        report(nm, time_ms([&] { k_red<3><<<bgrid, kThreads>>>(state, big_cells, BP, nullptr, nullptr, nullptr, wl, ppw8, sink); }, 5), BP, 2);
    }
    // shared-memory tile accumulation: 1184 CTAs (8 per SM) x 4224 points = 5M points
    {
        const size_t per = 4224; const unsigned g = 1184;
        report("smem_atomics_f32x2+max", time_ms([&] { k_smem_atomics<<<g, kThreads>>>(state, per, 1); }), per * g, 3);
        report("smem_int_rank+sts", time_ms([&] { k_smem_int_rank<<<g, kThreads>>>(state, per); }), per * g, 1);
    }
    {
        const int iters = 2048; const unsigned g = 148 * 8;
        const float ms = time_ms([&] { k_match_any<<<g, kThreads>>>(state, iters); });
        printf(" \"match_any\": {\"us\": %.2f, \"lane_ops_per_s_G\": %.2f, \"cycles_per_warp_instr_per_sm_at_1.9GHz\": %.2f},\n", ms * 1e3,
               (double)g * kThreads * iters / (ms * 1e-3) / 1e9, ms * 1e-3 * 1.9e9 / ((double)g * (kThreads / 32) * iters / 148.0));
    }
    printf(" \"end\": 0}\n");
    return 0;
}
