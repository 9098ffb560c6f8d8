// Microbenchmark (measurement aid, not product): P2P bandwidth between GPU 0 and GPU 1 for the
// message sizes of the multi-GPU finalize (8 MB .. 256 MB): kernel pull (LDG.128 from peer),
// kernel push (STG.128 to peer), cudaMemcpyPeerAsync.  nvcc -O3 -arch=sm_100a p2p_bw.cu -o p2p_bw
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_copy(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

int main()
{
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    int can01 = 0, can10 = 0;
    cudaDeviceCanAccessPeer(&can01, 0, 1); cudaDeviceCanAccessPeer(&can10, 1, 0);
    printf("canAccessPeer 0->1 %d, 1->0 %d\n", can01, can10);
    CK(cudaSetDevice(1)); cudaDeviceEnablePeerAccess(0, 0);
    CK(cudaSetDevice(0)); cudaDeviceEnablePeerAccess(1, 0);
    const size_t maxb = 256u << 20;
    uint4 *a0, *b0, *a1;
    CK(cudaSetDevice(0)); CK(cudaMalloc(&a0, maxb)); CK(cudaMalloc(&b0, maxb));
    CK(cudaSetDevice(1)); CK(cudaMalloc(&a1, maxb));
    CK(cudaSetDevice(0));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t bytes : {size_t(1) << 20, size_t(8) << 20, size_t(16) << 20, size_t(64) << 20, size_t(256) << 20}) {
        const size_t nv = bytes / 16;
        for (int mode = 0; mode < 4; ++mode) {
            for (int grid : {148 * 2, 148 * 8, 148 * 32}) {
                if (mode == 3 && grid != 148 * 2) continue;
                float best = 1e9f;
                for (int it = 0; it < 6; ++it) {
                    cudaEventRecord(e0, s);
                    if (mode == 0) k_copy<<<grid, 256, 0, s>>>(a0, b0, nv);            // local
                    if (mode == 1) k_copy<<<grid, 256, 0, s>>>(a1, b0, nv);            // pull from peer
                    if (mode == 2) k_copy<<<grid, 256, 0, s>>>(a0, a1, nv);            // push to peer
                    if (mode == 3) cudaMemcpyPeerAsync(a1, 1, a0, 0, bytes, s);
                    cudaEventRecord(e1, s); CK(cudaEventSynchronize(e1));
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (it > 0 && ms < best) best = ms;
                }
                const char* names[] = {"local", "pull", "push", "memcpyPeer"};
                printf("%4zu MB %-10s grid %5d : %8.1f us  %7.1f GB/s\n", bytes >> 20, names[mode], grid, best * 1e3, bytes / best / 1e6);
            }
        }
    }
    return 0;
}
