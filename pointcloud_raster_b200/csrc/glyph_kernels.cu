// glyph_kernels.cu — Line (Bresenham) and Gaussian glyph splats, sm_100a.
//
// Semantics follow the reference's CPU glyph code (src/engine/glyph_kernels.cu:79-281),
// NOT its CUDA kernels (which round Line endpoints in f32, :425-492): fractional
// cell position by multiply-with-reciprocal in f64, f32 footprint parameters,
// footprint clipped to the reference tile that holds the point's routed cell
// (src/engine/pipeline.cpp:699-709).  Every arithmetic step that decides a cell
// index uses _rn intrinsics so ptxas cannot contract it into an FMA the x86-64
// reference build does not have.
//
// All reductions sharing the glyph are fused: the record's additive words are
// [ sum(v_c * w) for each distinct value channel c ..., sum(w) ] and go out as one
// vector red per painted cell.
#include "kernels.cuh"
#include "exact_acc.cuh"

namespace pcrb {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxProfile = 161;      // separable fast path of the Gaussian scatter: footprints up to r = 80

// libstdc++ std::min/std::max (NaN behaviour differs from fminf/fmaxf), as called
// at glyph_kernels.cu:131,228-229 of the reference.
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }

// glibc cosf/sinf are correctly rounded for all but a vanishing fraction of
// inputs; evaluating in f64 and rounding once reproduces that.
__device__ __forceinline__ float cos_f32(float a) { return static_cast<float>(cos(static_cast<double>(a))); }
__device__ __forceinline__ float sin_f32(float a) { return static_cast<float>(sin(static_cast<double>(a))); }

struct ClipRect { int c0, r0, c1, r1; };   // [c0,c1) x [r0,r1), global cells

// GridConfig::tile_cell_range (src/core/grid_config.cpp:81-91) of the tile holding (col,row).
__device__ __forceinline__ ClipRect clip_of(const GridParams& g, int col, int row)
{
    ClipRect t;
    t.c0 = (col / g.tile_w) * g.tile_w;
    t.r0 = (row / g.tile_h) * g.tile_h;
    t.c1 = min(t.c0 + g.tile_w, g.width);
    t.r1 = min(t.r0 + g.tile_h, g.height);
    return t;
}

template <int NADD, bool XACC = false>
__device__ __forceinline__ void paint(uint32_t* __restrict__ state, const GridParams& g, int cx,
                                      int cy, const PassLayout& L, const float (&v)[kMaxChan],
                                      float w, const XAcc* xa = nullptr)
{
    constexpr int W = NADD <= 1 ? 1 : NADD <= 2 ? 2 : 4;
    float a[kMaxAdd];
#pragma unroll
    for (int j = 0; j < kMaxAdd; ++j) {
        if (j < NADD) {
            const int src = L.add_src[j];
            const float val = src == 0 ? v[0] : src == 1 ? v[1] : src == 2 ? v[2] : v[3];
            a[j] = src < 0 ? w : __fmul_rn(val, w);
        } else {
            a[j] = 0.0f;
        }
    }
    if constexpr (XACC) {                          // deterministic mode 2: exact, order-independent
#pragma unroll
        for (int j = 0; j < NADD; ++j) xacc_add(*xa, j, static_cast<size_t>(cy) * g.width + cx, a[j]);
    } else {
        float* rec = reinterpret_cast<float*>(state) + (static_cast<size_t>(cy) * g.width + cx) * W;
        red_add_words<NADD>(rec, a);
    }
}

__device__ __forceinline__ void mark_touched(const GridParams& g, uint32_t* touched, int col, int row)
{
    const int t = tile_of(g, col, row);
    if (touched[t] == 0) touched[t] = 1;
}

// ---------------------------------------------------------------------------
// Line: one thread per point, Bresenham walk.
// ---------------------------------------------------------------------------
template <int NADD, bool XACC>
__global__ void __launch_bounds__(kThreads)
k_line(const uint8_t* __restrict__ mask, const double* __restrict__ xs, const double* __restrict__ ys,
       const __grid_constant__ ChannelPtrs ch, const __grid_constant__ GlyphParams gp, size_t n,
       uint32_t* __restrict__ state, const __grid_constant__ GridParams g,
       const __grid_constant__ PassLayout L, uint32_t* __restrict__ touched, const __grid_constant__ XAcc xa)
{
    const size_t p = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    if (p >= n || (mask != nullptr && mask[p] == 0)) return;
    const double wx = xs[p], wy = ys[p];
    int col, row;
    if (!route_cell(g, wx, wy, col, row)) return;
    mark_touched(g, touched, col, row);
    const ClipRect clip = clip_of(g, col, row);

    float v[kMaxChan];
#pragma unroll
    for (int c = 0; c < kMaxChan; ++c) v[c] = (c < L.n_chan) ? ch.p[c][p] : 0.0f;

    const double fcx = __dmul_rn(__dsub_rn(wx, g.min_x), g.inv_csx);
    const double fcy = __dmul_rn(__dsub_rn(wy, g.max_y), g.inv_csy);

    const float dir = gp.direction   ? gp.direction[p]   : gp.default_direction;
    const float hl  = gp.half_length ? gp.half_length[p] : gp.default_half_length;
    const float hx = std_min(__fmul_rn(hl, static_cast<float>(g.inv_csx)), gp.max_radius_cells);
    const float hy = std_min(__fmul_rn(hl, static_cast<float>(g.inv_csy)), gp.max_radius_cells);
    const float px = __fmul_rn(hx, cos_f32(dir));
    const float py = __fmul_rn(hy, sin_f32(dir));

    // f64 add/sub, std::round = half away from zero
    const int ix0 = static_cast<int>(round(__dsub_rn(fcx, static_cast<double>(px))));
    const int iy0 = static_cast<int>(round(__dsub_rn(fcy, static_cast<double>(py))));
    const int ix1 = static_cast<int>(round(__dadd_rn(fcx, static_cast<double>(px))));
    const int iy1 = static_cast<int>(round(__dadd_rn(fcy, static_cast<double>(py))));

    const int adx = abs(ix1 - ix0), ady = abs(iy1 - iy0);
    const int stepx = ix0 < ix1 ? 1 : -1, stepy = iy0 < iy1 ? 1 : -1;
    int err = adx - ady, cx = ix0, cy = iy0;
    const int max_steps = 2 * (adx + ady) + 2;
    // Two-word records (sum, weight): two horizontally adjacent cells whose first record is 16-byte aligned
    // leave as ONE red.global.add.v4.f32 instead of two v2 — every cell of a line receives the same pair
    // {v * 1, 1}.  The kernel sits on the per-SM reduction issue rate (225 G REDs/s), so fewer REDs is the lever;
    // x-major lines pair up about 40 % of their cells.  Same adds per cell, hence identical results.
    constexpr bool kPair = (NADD == 2) && !XACC;
    bool pending = false;
    int pcx = 0, pcy = 0;
    float pa[kMaxAdd] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (kPair) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int src = L.add_src[j];
            pa[j] = src < 0 ? 1.0f : __fmul_rn(src == 0 ? v[0] : src == 1 ? v[1] : src == 2 ? v[2] : v[3], 1.0f);
        }
    }
    auto flush = [&]() {
        if (pending) red_add2(reinterpret_cast<float*>(state) + (static_cast<size_t>(pcy) * g.width + pcx) * 2, pa[0], pa[1]);
        pending = false;
    };
    for (int s = 0; s <= max_steps; ++s) {
        if (cx >= clip.c0 && cx < clip.c1 && cy >= clip.r0 && cy < clip.r1) {
            if constexpr (kPair) {
                bool merged = false;
                if (pending && cy == pcy && (cx == pcx + 1 || cx == pcx - 1)) {
                    const int lo = min(cx, pcx);
                    const size_t cell = static_cast<size_t>(cy) * g.width + lo;
                    if ((cell & 1) == 0) {
                        red_add4(reinterpret_cast<float*>(state) + cell * 2, pa[0], pa[1], pa[0], pa[1]);
                        pending = false;
                        merged = true;
                    }
                }
                if (!merged) { flush(); pending = true; pcx = cx; pcy = cy; }
            } else {
                paint<NADD, XACC>(state, g, cx, cy, L, v, 1.0f, &xa);
            }
        }
        if (cx == ix1 && cy == iy1) break;
        const int e2 = 2 * err;
        if (e2 > -ady) { err -= ady; cx += stepx; }
        if (e2 <  adx) { err += adx; cy += stepy; }
    }
    if constexpr (kPair) flush();
}

// ---------------------------------------------------------------------------
// Gaussian: one warp per point, lanes sweep the (2r+1)^2 footprint row-major so
// the reds of a warp land on consecutive records.
// ---------------------------------------------------------------------------
template <int NADD, bool XACC>
__global__ void __launch_bounds__(kThreads)
k_gaussian_warp(const uint8_t* __restrict__ mask, const double* __restrict__ xs, const double* __restrict__ ys,
                const __grid_constant__ ChannelPtrs ch, const __grid_constant__ GlyphParams gp,
                size_t n, uint32_t* __restrict__ state, const __grid_constant__ GridParams g,
                const __grid_constant__ PassLayout L, uint32_t* __restrict__ touched, const __grid_constant__ XAcc xa)
{
    __shared__ float s_profile[kThreads / 32][2][kMaxProfile];     // per warp: wx[], wy[]
    const int lane = threadIdx.x & 31;
    const bool unrotated = gp.rotation == nullptr && gp.default_rotation == 0.0f;
    const size_t warps_total = static_cast<size_t>(gridDim.x) * (kThreads / 32);
    for (size_t p = static_cast<size_t>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5); p < n;
         p += warps_total) {
        if (mask != nullptr && mask[p] == 0) continue;      // warp-uniform
        const double wx = xs[p], wy = ys[p];
        int col, row;
        if (!route_cell(g, wx, wy, col, row)) continue;     // warp-uniform
        if (lane == 0) mark_touched(g, touched, col, row);
        const ClipRect clip = clip_of(g, col, row);

        float v[kMaxChan];
#pragma unroll
        for (int c = 0; c < kMaxChan; ++c) v[c] = (c < L.n_chan) ? ch.p[c][p] : 0.0f;

        const double fcx = __dmul_rn(__dsub_rn(wx, g.min_x), g.inv_csx);
        const double fcy = __dmul_rn(__dsub_rn(wy, g.max_y), g.inv_csy);
        const double flx = floor(fcx), fly = floor(fcy);
        const float subx = static_cast<float>(__dsub_rn(fcx, flx));
        const float suby = static_cast<float>(__dsub_rn(fcy, fly));

        const float sxc = gp.sigma_x ? gp.sigma_x[p] : 0.0f;
        const float syc = gp.sigma_y ? gp.sigma_y[p] : 0.0f;
        const float sxw = (gp.sigma_x && sxc > 0.0f) ? sxc : gp.default_sigma_x;
        const float syw = (gp.sigma_y && syc > 0.0f) ? syc : gp.default_sigma_y;
        const float sx = __fmul_rn(sxw, static_cast<float>(g.inv_csx));
        const float sy = __fmul_rn(syw, static_cast<float>(g.inv_csy));   // < 0 for north-up

        const float rot = gp.rotation ? gp.rotation[p] : gp.default_rotation;
        const float cr = cos_f32(-rot);
        const float sr = sin_f32(-rot);
        const float nsr = -sr;

        const float R = std_min(__fmul_rn(3.0f, std_max(sx, sy)), gp.max_radius_cells);
        const int r = static_cast<int>(ceilf(R));
        const int icx = static_cast<int>(flx), icy = static_cast<int>(fly);
        if (r < 0) continue;

        const int side = 2 * r + 1;
        const int total = side * side;

        // Separable fast path.  Without rotation w(dx,dy) = exp(-ax/2 - ay/2) exactly (cos(-0) = 1,
        // sin(-0) = -0), so w = wx[dx] * wy[dy] up to ~5 ulp; that is only safe where the `w < 1e-6`
        // cut provably never fires (exponent bound E < 13, i.e. w > 2.2e-6) and the values are finite.
        // The two 1-D profiles (one IEEE divide + one expf per entry) are built by the warp in shared
        // memory; the (2r+1)^2 cells then cost two LDS and two multiplies each instead of two divides,
        // the rotation and an expf.  ncu (profiles/r01_other_kernels_ncu_full.json) showed the plain
        // per-cell evaluation to be issue-bound (92 % issue slots, L2 atomics 27 %).
        bool fast = unrotated && side <= kMaxProfile;
        if (fast) {
            const float m = static_cast<float>(r + 1);
            const float ex = m / sx, ey = m / sy;
            fast = (0.5f * (ex * ex + ey * ey)) < 13.0f;
#pragma unroll
            for (int c = 0; c < kMaxChan; ++c) fast = fast && (c >= L.n_chan || fabsf(v[c]) <= 3.0e38f);
        }
        if (fast) {
            float* wxs = s_profile[threadIdx.x >> 5][0];
            float* wys = s_profile[threadIdx.x >> 5][1];
            __syncwarp();
            for (int i = lane; i < side; i += 32) {
                const float d = static_cast<float>(i - r);
                const float tx = __fdiv_rn(__fsub_rn(d, subx), sx), ty = __fdiv_rn(__fsub_rn(d, suby), sy);
                wxs[i] = expf(__fmul_rn(-0.5f, __fmul_rn(tx, tx)));
                wys[i] = expf(__fmul_rn(-0.5f, __fmul_rn(ty, ty)));
            }
            __syncwarp();
            int dx = -r + lane, dy = -r;
            while (dx > r) { dx -= side; ++dy; }
            for (int idx = lane; idx < total; idx += 32) {
                const int gc = icx + dx, gr = icy + dy;
                if (gc >= clip.c0 && gc < clip.c1 && gr >= clip.r0 && gr < clip.r1) {
                    const float w = __fmul_rn(wxs[dx + r], wys[dy + r]);
                    if (!(w < 1e-6f)) paint<NADD, XACC>(state, g, gc, gr, L, v, w, &xa);
                }
                dx += 32;
                while (dx > r) { dx -= side; ++dy; }
            }
            continue;
        }

        int dx = -r + lane, dy = -r;
        while (dx > r) { dx -= side; ++dy; }
        for (int idx = lane; idx < total; idx += 32) {
            const int gc = icx + dx, gr = icy + dy;
            if (gc >= clip.c0 && gc < clip.c1 && gr >= clip.r0 && gr < clip.r1) {
                const float ox = __fsub_rn(static_cast<float>(dx), subx);
                const float oy = __fsub_rn(static_cast<float>(dy), suby);
                const float rx = __fadd_rn(__fmul_rn(ox, cr), __fmul_rn(oy, nsr));
                const float ry = __fadd_rn(__fmul_rn(ox, sr), __fmul_rn(oy, cr));
                const float qx = __fdiv_rn(rx, sx), qy = __fdiv_rn(ry, sy);
                const float e = __fmul_rn(-0.5f, __fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)));
                const float w = expf(e);
                if (!(w < 1e-6f)) paint<NADD, XACC>(state, g, gc, gr, L, v, w, &xa);
            }
            dx += 32;
            while (dx > r) { dx -= side; ++dy; }
        }
    }
}

template <typename F>
cudaError_t dispatch_nadd(int n_add, F&& f)
{
    switch (n_add) {
    case 1: return f(std::integral_constant<int, 1>{});
    case 2: return f(std::integral_constant<int, 2>{});
    case 3: return f(std::integral_constant<int, 3>{});
    case 4: return f(std::integral_constant<int, 4>{});
    }
    return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_line_accumulate(cudaStream_t s, const uint8_t* mask, const double* x, const double* y,
                                   const ChannelPtrs& ch, const GlyphParams& gp, size_t n,
                                   uint32_t* state, const GridParams& g, const PassLayout& L,
                                   uint32_t* touched, const XAcc* exact)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = static_cast<unsigned>((n + kThreads - 1) / kThreads);
    return dispatch_nadd(L.n_add, [&](auto na) {
        if (exact) k_line<decltype(na)::value, true><<<grid, kThreads, 0, s>>>(mask, x, y, ch, gp, n, state, g, L, touched, *exact);
        else k_line<decltype(na)::value, false><<<grid, kThreads, 0, s>>>(mask, x, y, ch, gp, n, state, g, L, touched, XAcc{});
        return cudaGetLastError();
    });
}

cudaError_t launch_gaussian_accumulate(cudaStream_t s, const uint8_t* mask, const double* x, const double* y,
                                       const ChannelPtrs& ch, const GlyphParams& gp, size_t n,
                                       uint32_t* state, const GridParams& g, const PassLayout& L,
                                       uint32_t* touched, const XAcc* exact)
{
    if (n == 0) return cudaSuccess;
    const size_t warps_per_block = kThreads / 32;
    size_t blocks = (n + warps_per_block - 1) / warps_per_block;
    if (blocks > 148 * 64) blocks = 148 * 64;
    return dispatch_nadd(L.n_add, [&](auto na) {
        if (exact)
            k_gaussian_warp<decltype(na)::value, true><<<static_cast<unsigned>(blocks), kThreads, 0, s>>>(
                mask, x, y, ch, gp, n, state, g, L, touched, *exact);
        else
            k_gaussian_warp<decltype(na)::value, false><<<static_cast<unsigned>(blocks), kThreads, 0, s>>>(
                mask, x, y, ch, gp, n, state, g, L, touched, XAcc{});
        return cudaGetLastError();
    });
}

}  // namespace pcrb
