// Microbenchmark (measurement aid, not product): ways to move the multi-GPU finalize's record
// slices (1..16 MB) from GPU 0 into GPU 1's memory — the message sizes where NVLink latency,
// not bandwidth, decides.  nvcc -O3 -arch=sm_100a p2p_push.cu -o p2p_push
//   thread16   one 16-B load + store per thread, thread per element (what k_push_slices does)
//   strided    grid-stride loop, 16 B per iteration
//   ilp4       4 independent 16-B loads, then 4 stores, per iteration
//   bulk       cp.async.bulk global->shared, then shared->peer global, 16 KB tiles, 2 stages
//   memcpy     cudaMemcpyAsync (copy engine)
//   memcpy x4  the same bytes as 4 concurrent copies on 4 streams
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_thread16(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
__global__ void k_strided(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
__global__ void k_ilp4(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = src[i];
}

constexpr int kTile = 16384;
__global__ void __launch_bounds__(32) k_bulk(const char* __restrict__ src, char* __restrict__ dst, size_t bytes)
{
    extern __shared__ __align__(128) char smem[];
    __shared__ __align__(8) unsigned long long bar[2];
    const unsigned b0 = (unsigned)__cvta_generic_to_shared(&bar[0]), b1 = (unsigned)__cvta_generic_to_shared(&bar[1]);
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" :: "r"(b0));
        asm volatile("mbarrier.init.shared.b64 [%0], 1;" :: "r"(b1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    const size_t ntiles = (bytes + kTile - 1) / kTile;
    unsigned phase[2] = {0, 0};
    int it = 0;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int st = it & 1;
        const unsigned bar_a = st ? b1 : b0, sm = s0 + st * kTile;
        const unsigned len = (unsigned)min((size_t)kTile, bytes - t * kTile);
        if (it >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that used this stage
        asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;" :: "r"(bar_a), "r"(len) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(sm), "l"(src + t * kTile), "r"(len), "r"(bar_a) : "memory");
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar_a), "r"(phase[st]) : "memory");
        phase[st] ^= 1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(dst + t * kTile), "r"(sm), "r"(len) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    CK(cudaSetDevice(1)); cudaDeviceEnablePeerAccess(0, 0);
    CK(cudaSetDevice(0)); cudaDeviceEnablePeerAccess(1, 0);
    const size_t maxb = 16u << 20;
    char *a0, *a1;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&a1, maxb));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&a0, maxb)); CK(cudaMemset(a0, 1, maxb));
    cudaStream_t s[4]; for (auto& x : s) CK(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, ej[4]; cudaEventCreate(&e0); cudaEventCreate(&e1); for (auto& x : ej) cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kTile));
    const char* names[] = {"thread16", "strided x4/SM", "strided x16/SM", "ilp4 x8/SM", "bulk 1/SM", "bulk 4/SM", "memcpy", "memcpy x4"};
    for (size_t bytes : {size_t(1) << 20, size_t(2) << 20, size_t(8) << 20, size_t(14) << 20}) {
        const size_t nv = bytes / 16;
        for (int mode = 0; mode < 8; ++mode) {
            float best = 1e9f;
            for (int it = 0; it < 8; ++it) {
                cudaEventRecord(e0, s[0]);
                switch (mode) {
                case 0: k_thread16<<<(unsigned)((nv + 255) / 256), 256, 0, s[0]>>>((const uint4*)a0, (uint4*)a1, nv); break;
                case 1: k_strided<<<148 * 4, 256, 0, s[0]>>>((const uint4*)a0, (uint4*)a1, nv); break;
                case 2: k_strided<<<148 * 16, 256, 0, s[0]>>>((const uint4*)a0, (uint4*)a1, nv); break;
                case 3: k_ilp4<<<148 * 8, 256, 0, s[0]>>>((const uint4*)a0, (uint4*)a1, nv); break;
                case 4: k_bulk<<<148, 32, 2 * kTile, s[0]>>>(a0, a1, bytes); break;
                case 5: k_bulk<<<148 * 4, 32, 2 * kTile, s[0]>>>(a0, a1, bytes); break;
                case 6: cudaMemcpyAsync(a1, a0, bytes, cudaMemcpyDeviceToDevice, s[0]); break;
                case 7:
                    for (int k = 1; k < 4; ++k) { cudaEventRecord(ej[0], s[0]); cudaStreamWaitEvent(s[k], ej[0], 0); }
                    for (int k = 0; k < 4; ++k) cudaMemcpyAsync(a1 + k * (bytes / 4), a0 + k * (bytes / 4), bytes / 4, cudaMemcpyDeviceToDevice, s[k]);
                    for (int k = 1; k < 4; ++k) { cudaEventRecord(ej[k], s[k]); cudaStreamWaitEvent(s[0], ej[k], 0); }
                    break;
                }
                cudaEventRecord(e1, s[0]); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 1 && ms < best) best = ms;
            }
            printf("%3zu MB %-15s: %7.1f us  %7.1f GB/s\n", bytes >> 20, names[mode], best * 1e3, bytes / best / 1e6);
        }
    }
    // both directions at once (every rank pushes while it is being pushed to)
    {
        char *b0, *b1;
        CK(cudaSetDevice(0)); CK(cudaMalloc(&b0, maxb));
        CK(cudaSetDevice(1)); CK(cudaMalloc(&b1, maxb)); CK(cudaMemset(b1, 2, maxb));
        cudaStream_t t1; CK(cudaStreamCreateWithFlags(&t1, cudaStreamNonBlocking));
        cudaEvent_t f0, f1; cudaEventCreate(&f0); cudaEventCreate(&f1);
        for (size_t bytes : {size_t(2) << 20, size_t(8) << 20, size_t(14) << 20}) {
            const size_t nv = bytes / 16;
            float best0 = 1e9f, best1 = 1e9f;
            for (int it = 0; it < 8; ++it) {
                CK(cudaSetDevice(0)); cudaEventRecord(e0, s[0]);
                k_thread16<<<(unsigned)((nv + 255) / 256), 256, 0, s[0]>>>((const uint4*)a0, (uint4*)a1, nv);
                cudaEventRecord(e1, s[0]);
                CK(cudaSetDevice(1)); cudaEventRecord(f0, t1);
                k_thread16<<<(unsigned)((nv + 255) / 256), 256, 0, t1>>>((const uint4*)b1, (uint4*)b0, nv);
                cudaEventRecord(f1, t1);
                CK(cudaEventSynchronize(e1)); CK(cudaEventSynchronize(f1));
                float m0, m1; cudaEventElapsedTime(&m0, e0, e1); cudaEventElapsedTime(&m1, f0, f1);
                if (it > 1) { best0 = fminf(best0, m0); best1 = fminf(best1, m1); }
            }
            printf("%3zu MB both directions: 0->1 %7.1f us, 1->0 %7.1f us\n", bytes >> 20, best0 * 1e3, best1 * 1e3);
        }
        CK(cudaSetDevice(0));
    }
    return 0;
}
