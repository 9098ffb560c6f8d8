// det_kernels.cu — deterministic Point path: sort by cell, then an in-order
// segmented reduce.  Bit-reproducible run to run (no atomics anywhere).
//
// The reference sorts too (64-bit (tile,cell) keys through all 64 bits, 5 co-sorted
// arrays, then one-thread-per-point atomics: tile_router_kernels.cu:169-293,
// accumulator_kernels.cu:31-133) but never exploits the order.  Here the key is
// the bare cell index with only ceil(log2(cells+1)) bits sorted, the payload is
// the point index, the radix sort is stable, and each cell's run is folded in
// original point order into ONE thread's registers (long runs with the warp's help
// for the loads) — so the float sums are a fixed left-to-right fold per cell per
// ingest, and the record is updated with a plain read-modify-write.
#include "engine.h"

#include <cub/device/device_radix_sort.cuh>

namespace pcrb {

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
k_det_keys(const uint8_t* __restrict__ mask, const double* __restrict__ xs, const double* __restrict__ ys, size_t n,
           const __grid_constant__ GridParams g, uint32_t* __restrict__ keys,
           uint32_t* __restrict__ idx, uint32_t* __restrict__ touched)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (i >= n) return;
    int col, row;
    const bool ok = route_cell(g, xs[i], ys[i], col, row) && (mask == nullptr || mask[i] != 0);
    const uint32_t invalid = static_cast<uint32_t>(static_cast<size_t>(g.width) * g.height);
    keys[i] = ok ? static_cast<uint32_t>(static_cast<size_t>(row) * g.width + col) : invalid;
    idx[i] = static_cast<uint32_t>(i);
    if (ok) {
        const int t = tile_of(g, col, row);
        if (touched[t] == 0) touched[t] = 1;
    }
}

// Thread i owns the run starting at sorted position i (if i is a run head) and folds its first kDetSerial
// entries itself.  A longer run (hot cells: scan lines, config 5's points clipped onto the bbox edge) is then
// finished by the whole warp, 32 entries per step: the gathers idx[k] -> value run in parallel and one step
// ahead, the additions stay strictly in sorted (= original point) order in the head's registers, so the
// result is bit for bit what a single thread walking the run would produce — without its chain of
// dependent global loads (hundreds of cycles per entry).
constexpr int kDetSerial = 32;

template <int W>
__global__ void __launch_bounds__(kThreads)
k_det_reduce(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ idx, size_t n,
             uint32_t invalid, const __grid_constant__ ChannelPtrs ch,
             uint32_t* __restrict__ state, const __grid_constant__ PassLayout L)
{
    constexpr unsigned kFull = 0xffffffffu;
    const size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t key = invalid;
    bool head = false;
    if (i < n) {
        key = keys[i];
        head = key != invalid && (i == 0 || keys[i - 1] != key);
    }

    float add[kMaxAdd];
    int32_t mx[kMaxExt], mn[kMaxExt];
    auto fold = [&](const float (&v)[kMaxChan]) {
        auto pick = [&](int src) { return src == 0 ? v[0] : src == 1 ? v[1] : src == 2 ? v[2] : v[3]; };
#pragma unroll
        for (int j = 0; j < kMaxAdd; ++j)
            if (j < L.n_add) add[j] = __fadd_rn(add[j], L.add_src[j] < 0 ? 1.0f : pick(L.add_src[j]));
#pragma unroll
        for (int j = 0; j < kMaxExt; ++j)
            if (j < L.n_max) { const float m = pick(L.max_src[j]); if (m == m) mx[j] = max(mx[j], f32_ordered(m)); }
#pragma unroll
        for (int j = 0; j < kMaxExt; ++j)
            if (j < L.n_min) { const float m = pick(L.min_src[j]); if (m == m) mn[j] = min(mn[j], f32_ordered(m)); }
    };
    auto gather = [&](size_t k, float (&v)[kMaxChan]) {
        const uint32_t p = idx[k];
#pragma unroll
        for (int c = 0; c < kMaxChan; ++c) v[c] = (c < L.n_chan) ? ch.p[c][p] : 0.0f;
    };

    uint32_t* rec = state + static_cast<size_t>(head ? key : 0) * W;
    size_t k = i;
    bool more = false;
    if (head) {
#pragma unroll
        for (int j = 0; j < kMaxAdd; ++j) add[j] = (j < L.n_add) ? __uint_as_float(rec[j]) : 0.0f;
#pragma unroll
        for (int j = 0; j < kMaxExt; ++j) mx[j] = (j < L.n_max) ? static_cast<int32_t>(rec[L.n_add + j]) : 0;
#pragma unroll
        for (int j = 0; j < kMaxExt; ++j) mn[j] = (j < L.n_min) ? static_cast<int32_t>(rec[L.n_add + L.n_max + j]) : 0;
        const size_t stop = min(n, i + static_cast<size_t>(kDetSerial));
        for (; k < stop && keys[k] == key; ++k) {
            float v[kMaxChan];
            gather(k, v);
            fold(v);
        }
        more = k < n && keys[k] == key;
    }

    // ---- long runs: one after another, by the whole warp (every lane of the warp gets here) ----
    for (unsigned todo = __ballot_sync(kFull, more); todo; todo &= todo - 1) {
        const int h = __ffs(todo) - 1;
        const uint32_t hkey = __shfl_sync(kFull, key, h);
        size_t pos = static_cast<size_t>(__shfl_sync(kFull, static_cast<unsigned long long>(k), h));
        float v[kMaxChan] = {0.f, 0.f, 0.f, 0.f}, vn[kMaxChan] = {0.f, 0.f, 0.f, 0.f};
        bool ok = pos + lane < n && keys[pos + lane] == hkey;
        if (ok) gather(pos + lane, v);
        for (;;) {
            // the run is contiguous in the sorted keys: the lanes that still see it form a prefix
            const int cnt = __popc(__ballot_sync(kFull, ok));
            const size_t next = pos + 32;
            const bool okn = cnt == 32 && next + lane < n && keys[next + lane] == hkey;
            if (okn) gather(next + lane, vn);                   // in flight while this step is folded
            for (int j = 0; j < cnt; ++j) {
                float vj[kMaxChan];
#pragma unroll
                for (int c = 0; c < kMaxChan; ++c) vj[c] = (c < L.n_chan) ? __shfl_sync(kFull, v[c], j) : 0.0f;
                if (lane == h) fold(vj);
            }
            if (cnt < 32) break;
            pos = next;
            ok = okn;
#pragma unroll
            for (int c = 0; c < kMaxChan; ++c) v[c] = vn[c];
        }
    }

    if (head) {
#pragma unroll
        for (int j = 0; j < kMaxAdd; ++j) if (j < L.n_add) rec[j] = __float_as_uint(add[j]);
#pragma unroll
        for (int j = 0; j < kMaxExt; ++j) if (j < L.n_max) rec[L.n_add + j] = static_cast<uint32_t>(mx[j]);
#pragma unroll
        for (int j = 0; j < kMaxExt; ++j) if (j < L.n_min) rec[L.n_add + L.n_max + j] = static_cast<uint32_t>(mn[j]);
    }
}

}  // namespace

size_t det_sort_temp_bytes(size_t n, int key_bits)
{
    size_t bytes = 0;
    cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, static_cast<int64_t>(n), 0, key_bits);
    return bytes;
}

cudaError_t det_build_keys(cudaStream_t s, const uint8_t* mask, const double* x, const double* y, size_t n,
                           const GridParams& g, uint32_t* keys, uint32_t* idx, uint32_t* touched)
{
    const unsigned grid = static_cast<unsigned>((n + kThreads - 1) / kThreads);
    k_det_keys<<<grid, kThreads, 0, s>>>(mask, x, y, n, g, keys, idx, touched);
    return cudaGetLastError();
}

cudaError_t det_sort(cudaStream_t s, void* tmp, size_t tmp_bytes, uint32_t*& keys, uint32_t*& keys_alt,
                     uint32_t*& idx, uint32_t*& idx_alt, size_t n, int key_bits)
{
    cub::DoubleBuffer<uint32_t> k(keys, keys_alt), v(idx, idx_alt);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k, v, static_cast<int64_t>(n), 0,
                                                    key_bits, s);
    if (e != cudaSuccess) return e;
    if (k.Current() != keys) std::swap(keys, keys_alt);
    if (v.Current() != idx) std::swap(idx, idx_alt);
    return cudaSuccess;
}

cudaError_t det_point_reduce(cudaStream_t s, const uint32_t* keys, const uint32_t* idx, size_t n,
                             const ChannelPtrs& ch, uint32_t* state, const PassLayout& L,
                             uint32_t invalid_key)
{
    const unsigned grid = static_cast<unsigned>((n + kThreads - 1) / kThreads);
    switch (L.width) {
    case 1: k_det_reduce<1><<<grid, kThreads, 0, s>>>(keys, idx, n, invalid_key, ch, state, L); break;
    case 2: k_det_reduce<2><<<grid, kThreads, 0, s>>>(keys, idx, n, invalid_key, ch, state, L); break;
    case 4: k_det_reduce<4><<<grid, kThreads, 0, s>>>(keys, idx, n, invalid_key, ch, state, L); break;
    case 8: k_det_reduce<8><<<grid, kThreads, 0, s>>>(keys, idx, n, invalid_key, ch, state, L); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace pcrb
