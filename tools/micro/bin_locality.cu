// Microbenchmark (measurement aid, not product): does the L2 keep the records of a bin while the
// persistent page-walking accumulate kernel folds the bin's entries?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a bin_locality.cu -o bin_locality
// Entries are generated in bin order: entry i belongs to bin i / per_bin, its cell is uniform inside the bin.
// The kernel is the structure of k_bin_accumulate: 8 CTAs per SM, CTA c folds pages c, c+G, c+2G, ...
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__global__ void k_gen(uint32_t* cell, float* val, size_t n, size_t per_bin, int shift)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t h = hash32((uint32_t)i * 2654435761u + 99u);
        cell[i] = (uint32_t)((i / per_bin) << shift) + (h & ((1u << shift) - 1));
        val[i] = (h >> 8) * (1.0f / 16777216.0f);
    }
}
// POL: 0 = plain REDs; 1 = REDs with L2::evict_last; 2 = REDs evict_last + entry loads L2::evict_first;
//      3 = REDs evict_last, entry loads plain ld.cs; 4 = fractional evict_last 0.5
template <int POL>
__global__ void __launch_bounds__(256) k_acc(const uint32_t* __restrict__ cell, const float* __restrict__ val, size_t n,
                                             uint32_t* __restrict__ state, size_t per_bin, int shift)
{
    const size_t npages = (n + 4095) / 4096;
    uint64_t pol_last, pol_first;
    if (POL == 4) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_unchanged.b64 %0, 0.5;" : "=l"(pol_last));
    else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    for (size_t pg = blockIdx.x; pg < npages; pg += gridDim.x) {
        const size_t e0 = pg * 4096;
#pragma unroll 4
        for (uint32_t t = threadIdx.x; t < 4096; t += 256) {
            if (e0 + t >= n) break;
            uint32_t c; float v;
            if (POL == 5) {
                const size_t i = e0 + t;
                const uint32_t h = hash32((uint32_t)i * 2654435761u + 99u);
                c = (uint32_t)((i / per_bin) << shift) + (h & ((1u << shift) - 1));
                v = (h >> 8) * (1.0f / 16777216.0f);
            } else if (POL == 2) {
                asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(c) : "l"(cell + e0 + t), "l"(pol_first));
                asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(val + e0 + t), "l"(pol_first));
            } else { c = __ldcs(cell + e0 + t); v = __ldcs(val + e0 + t); }
            uint32_t* rec = state + (size_t)c * 4;
            if (POL == 0 || POL == 5) {
                asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" :: "l"(rec), "f"(v), "f"(1.0f) : "memory");
                asm volatile("red.relaxed.gpu.global.max.s32 [%0], %1;" :: "l"(rec + 2), "r"(__float_as_int(v)) : "memory");
            } else {
                asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" :: "l"(rec), "f"(v), "f"(1.0f), "l"(pol_last) : "memory");
                asm volatile("red.relaxed.gpu.global.max.L2::cache_hint.s32 [%0], %1, %2;" :: "l"(rec + 2), "r"(__float_as_int(v)), "l"(pol_last) : "memory");
            }
        }
    }
}
template <int RANDOM_ORDER>
__global__ void __launch_bounds__(256) k_acc_np(size_t n, uint32_t* __restrict__ state, size_t per_bin, int shift, size_t nbins)
{
    const size_t base = (size_t)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const size_t i = base + (size_t)u * 256;
        if (i >= n) continue;
        const uint32_t h = hash32((uint32_t)i * 2654435761u + 99u);
        size_t bin = i / per_bin;
        if (RANDOM_ORDER) bin = hash32((uint32_t)bin * 7919u + 13u) % nbins;
        const uint32_t c = (uint32_t)(bin << shift) + (h & ((1u << shift) - 1));
        const float v = (h >> 8) * (1.0f / 16777216.0f);
        uint32_t* rec = state + (size_t)c * 4;
        asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};" :: "l"(rec), "f"(v), "f"(1.0f) : "memory");
        asm volatile("red.relaxed.gpu.global.max.s32 [%0], %1;" :: "l"(rec + 2), "r"(__float_as_int(v)) : "memory");
    }
}
int main(int argc, char** argv)
{
    const size_t n = argc > 1 ? (size_t)atof(argv[1]) : 400'000'000;
    const size_t cells = 400'000'000;
    uint32_t *cell, *state; float* val;
    CK(cudaMalloc(&cell, n * 4)); CK(cudaMalloc(&val, n * 4)); CK(cudaMalloc(&state, cells * 16));
    CK(cudaMemset(state, 0, cells * 16));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int shift : {20}) {
        const size_t nbins = cells >> shift;
        const size_t per_bin = (n + nbins - 1) / nbins;
        k_gen<<<148 * 8, 256>>>(cell, val, n, per_bin, shift);
        for (int ro = 0; ro < 2; ++ro) {
            float best = 1e9f;
            for (int r = 0; r < 3; ++r) {
                CK(cudaEventRecord(a));
                if (ro) k_acc_np<1><<<(unsigned)((n + 1023) / 1024), 256>>>(n, state, per_bin, shift, nbins);
                else k_acc_np<0><<<(unsigned)((n + 1023) / 1024), 256>>>(n, state, per_bin, shift, nbins);
                CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
                float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = std::min(best, ms);
            }
            printf("entries %zu bin 2^%d non-persistent, no loads, %s bin order: %.3f ms = %.1f Gpts/s\n", n, shift, ro ? "hashed" : "sequential", best, n / (best * 1e-3) / 1e9);
        }
        for (int pol : {0, 5}) {
            float best = 1e9f;
            for (int r = 0; r < 3; ++r) {
                CK(cudaEventRecord(a));
                if (pol == 0) k_acc<0><<<148 * 8, 256>>>(cell, val, n, state, per_bin, shift);
                if (pol == 1) k_acc<1><<<148 * 8, 256>>>(cell, val, n, state, per_bin, shift);
                if (pol == 2) k_acc<2><<<148 * 8, 256>>>(cell, val, n, state, per_bin, shift);
                if (pol == 3) k_acc<3><<<148 * 8, 256>>>(cell, val, n, state, per_bin, shift);
                if (pol == 4) k_acc<4><<<148 * 8, 256>>>(cell, val, n, state, per_bin, shift);
                if (pol == 5) k_acc<5><<<148 * 8, 256>>>(cell, val, n, state, per_bin, shift);
                CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
                float ms; CK(cudaEventElapsedTime(&ms, a, b)); best = std::min(best, ms);
            }
            printf("entries %zu bin 2^%d cells (%zu MB of records, %zu entries/bin) policy=%d: %.3f ms = %.1f Gpts/s\n", n, shift,
                   ((size_t)16 << shift) >> 20, per_bin, pol, best, n / (best * 1e-3) / 1e9);
        }
    }
    return 0;
}
