// Microbenchmark (measurement aid, not product): NVLink throughput of the binning kernel's copy-out as a function of
// the RUN length — every CTA issues cp.async.bulk shared->global stores of `run` bytes out of a double-buffered 32 KB
// staging area, to destinations scattered over a 2 GB pool on the peer GPU (or on the local one).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a p2p_runs.cu -o p2p_runs
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(512, 2) k_runs(char* __restrict__ dst, size_t pool_bytes, uint32_t run, int chunks)
{
    extern __shared__ __align__(128) char stage[];                    // [2][32 KB]
    const uint32_t per_chunk = 32768 / run;                            // runs per chunk
    uint32_t h = blockIdx.x * 2654435761u + 12345u;
    for (int c = 0; c < chunks; ++c) {
        char* buf = stage + (c & 1) * 32768;
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) reinterpret_cast<uint4*>(buf)[i] = make_uint4(c, i, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        for (uint32_t r = threadIdx.x; r < per_chunk; r += blockDim.x) {
            uint32_t x = h + (c * per_chunk + r) * 0x9E3779B9u; x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15;
            const size_t off = (static_cast<size_t>(x) * 4096) % (pool_bytes - 65536) / 16 * 16;   // scattered, 16-byte aligned
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst + off), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(buf + r * run))), "r"(run) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
    int n = 0; CK(cudaGetDeviceCount(&n));
    const size_t pool = size_t(2) << 30;
    char *local, *peer = nullptr;
    CK(cudaSetDevice(0)); CK(cudaMalloc(&local, pool));
    if (n >= 2) { cudaDeviceEnablePeerAccess(1, 0); CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, pool)); cudaDeviceEnablePeerAccess(0, 0); CK(cudaSetDevice(0)); }
    CK(cudaFuncSetAttribute(k_runs, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int chunks = 64;                                             // 296 CTAs x 64 x 32 KB = 620 MB per launch
    for (char* dst : {local, peer}) {
        if (!dst) continue;
        for (uint32_t run : {128u, 256u, 352u, 704u, 1024u, 4096u, 32768u}) {
            if (32768 % run) { /* 352, 704 do not divide 32 KB: fine, the tail of a chunk is unused */ }
            float best = 1e9f;
            for (int it = 0; it < 5; ++it) {
                cudaEventRecord(a);
                k_runs<<<296, 512, 65536>>>(dst, pool, run, chunks);
                cudaEventRecord(b); CK(cudaEventSynchronize(b));
                float ms; cudaEventElapsedTime(&ms, a, b); if (it > 0 && ms < best) best = ms;
            }
            const double bytes = 296.0 * chunks * (32768 / run) * run;
            printf("%s run %6u B: %8.1f us  %7.1f GB/s\n", dst == local ? "local" : "peer ", run, best * 1e3, bytes / best / 1e6);
        }
    }
    return 0;
}
