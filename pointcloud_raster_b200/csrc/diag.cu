// diag.cu — measurement aids exported through the C-ABI (not on the product path).
//
// pcr_diag_red_ceiling: what does this GPU allow for the accumulate step of the Point glyph,
// independent of our kernel?  `points` synthetic points whose cell is a hash of the point index are
// folded into `cells` 16-byte records [sum, count, max, pad] with exactly the reduction instructions
// k_point_direct issues (red.global.add.v2.f32 + red.global.max.s32), optionally behind the same three
// streaming loads per point (x, y f64 + v f32 = 20 B).  bench.py reports the result next to the HBM
// roofline fraction: the kernel is bound by the rate at which L2 resolves scattered reductions.
#include <algorithm>
#include <cstdint>
#include <new>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/pcr_b200.h"
#include "common.cuh"

namespace {

constexpr int kThreads = 256, kUnroll = 4;

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// MODE bits: 1 = vector add (sum, count), 2 = max, 8 = streaming loads of x, y, v
template <int MODE>
__global__ void __launch_bounds__(kThreads)
k_red_probe(uint32_t* __restrict__ state, uint32_t cells, size_t n, const double* __restrict__ xs,
            const double* __restrict__ ys, const float* __restrict__ vs, float* __restrict__ sink)
{
    const size_t base = static_cast<size_t>(blockIdx.x) * (kThreads * kUnroll) + threadIdx.x;
    double x[kUnroll], y[kUnroll];
    float v[kUnroll];
    float keep = 0.f;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const size_t i = base + static_cast<size_t>(u) * kThreads;
        if constexpr ((MODE & 8) != 0) {
            x[u] = i < n ? pcrb::ldg_stream_d(xs + i) : 0.0;
            y[u] = i < n ? pcrb::ldg_stream_d(ys + i) : 0.0;
            v[u] = i < n ? pcrb::ldg_stream_f(vs + i) : 0.f;
        } else { x[u] = 0; y[u] = 0; v[u] = 0.5f; }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const size_t i = base + static_cast<size_t>(u) * kThreads;
        if (i >= n) continue;
        const uint32_t h = hash32(static_cast<uint32_t>(i) * 2654435761u + 12345u);
        uint32_t* rec = state + static_cast<size_t>(h % cells) * 4;
        if constexpr ((MODE & 8) != 0) keep += static_cast<float>(x[u] + y[u]);
        if constexpr ((MODE & 1) != 0) pcrb::red_add2(reinterpret_cast<float*>(rec), v[u], 1.0f);
        if constexpr ((MODE & 2) != 0) pcrb::red_max(reinterpret_cast<int32_t*>(rec) + 2, __float_as_int(v[u]) + static_cast<int>(h & 1023));
    }
    if constexpr ((MODE & 8) != 0) if (keep == 1.2345f) *sink = keep;
}

}  // namespace

extern "C" int pcr_diag_red_ceiling(int device, uint64_t points, uint64_t cells, int with_loads, int reps,
                                    double* median_us)
{
    if (!median_us || points == 0 || cells == 0 || cells > 0xffffffffull || reps < 1) return PCR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return PCR_CUDA_ERROR;
    uint32_t* state = nullptr;
    double *xs[4] = {}, *ys[4] = {};
    float* vs[4] = {};
    float* sink = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
    int rc = PCR_OK;
    auto ok = [&](cudaError_t e) { if (e != cudaSuccess && rc == PCR_OK) rc = PCR_CUDA_ERROR; return e == cudaSuccess; };
    ok(cudaMalloc(&state, cells * 16)) && ok(cudaMemset(state, 0, cells * 16));
    ok(cudaMalloc(&sink, 4));
    // four rotated point sets, so that the loads of a repetition never find their input in L2
    const int sets = with_loads ? 4 : 0;
    for (int r = 0; r < sets && rc == PCR_OK; ++r) {
        ok(cudaMalloc(&xs[r], points * 8)) && ok(cudaMemset(xs[r], 0, points * 8));
        ok(cudaMalloc(&ys[r], points * 8)) && ok(cudaMemset(ys[r], 0, points * 8));
        ok(cudaMalloc(&vs[r], points * 4)) && ok(cudaMemset(vs[r], 0, points * 4));
    }
    ok(cudaEventCreate(&a)); ok(cudaEventCreate(&b));
    if (rc == PCR_OK) {
        const unsigned grid = static_cast<unsigned>((points + kThreads * kUnroll - 1) / (kThreads * kUnroll));
        std::vector<float> t;
        try { t.reserve(reps); } catch (const std::bad_alloc&) { rc = PCR_OUT_OF_MEMORY; }
        for (int i = 0; i < reps + 3 && rc == PCR_OK; ++i) {
            const int r = i & 3;
            ok(cudaEventRecord(a));
            if (with_loads) k_red_probe<11><<<grid, kThreads>>>(state, static_cast<uint32_t>(cells), points, xs[r], ys[r], vs[r], sink);
            else            k_red_probe<3><<<grid, kThreads>>>(state, static_cast<uint32_t>(cells), points, nullptr, nullptr, nullptr, sink);
            ok(cudaEventRecord(b));
            ok(cudaEventSynchronize(b));
            float ms = 0.f;
            ok(cudaEventElapsedTime(&ms, a, b));
            if (i >= 3) t.push_back(ms);             // 3 warm-up launches
        }
        if (rc == PCR_OK) {
            std::sort(t.begin(), t.end());
            *median_us = t[t.size() / 2] * 1e3;
        }
    }
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    for (int r = 0; r < 4; ++r) { cudaFree(xs[r]); cudaFree(ys[r]); cudaFree(vs[r]); }
    cudaFree(state); cudaFree(sink);
    return rc;
}
