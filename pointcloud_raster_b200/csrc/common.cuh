// common.cuh — device-side building blocks shared by every sm_100a kernel of the
// ingest/finalize path: grid geometry, the bit-exact point->cell routing rule,
// the accumulator record layout and the reduction primitives (red.global.*).
//
// Reference citations are relative to the reference repository root.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace pcrb {

// ---------------------------------------------------------------------------
// Grid geometry handed to kernels by value (__grid_constant__).
// ---------------------------------------------------------------------------
struct GridParams {
    double min_x, max_x, min_y, max_y;   // BBox, inclusive on all four edges
    double csx, csy;                     // cell sizes (csy < 0 = north-up)
    double inv_csx, inv_csy;             // 1.0/cs, as the glyph code computes it
    int    width, height;                // cells
    int    tile_w, tile_h;               // reference tile size (cells)
    int    tiles_x, tiles_y;
    int    exact_x, exact_y;             // |cs| is a power of two: d/cs == d*inv bit-exactly
    int    tile_w_shift, tile_h_shift;   // log2 of the tile size when it is a power of two (default 4096), else -1
};

// GridConfig::world_to_cell (src/core/grid_config.cpp:24-43) + BBox::contains
// (src/core/types.cpp:41-43): inclusive bounds test (NaN fails), IEEE f64
// subtract, DIVIDE, floor, clamp.  When the cell size is a power of two the
// quotient d/cs and the product d*(1/cs) are the same real number, hence the
// same double: the multiply is then bit-exact and 10x cheaper than DDIV.
// _rn intrinsics keep ptxas from contracting anything into an FMA.
// EXACT = both cell sizes are powers of two (compile-time variant without the DDIV path: fewer
// registers, higher occupancy for the streaming Point kernel).
template <bool EXACT = false>
__device__ __forceinline__ bool route_cell(const GridParams& g, double x, double y,
                                           int& col, int& row)
{
    const bool inside = (x >= g.min_x) && (x <= g.max_x) && (y >= g.min_y) && (y <= g.max_y);
    const double dx = __dsub_rn(x, g.min_x);
    const double dy = __dsub_rn(y, g.max_y);
    double qx, qy;
    if constexpr (EXACT) {
        qx = __dmul_rn(dx, g.inv_csx);
        qy = __dmul_rn(dy, g.inv_csy);
    } else {
        qx = g.exact_x ? __dmul_rn(dx, g.inv_csx) : __ddiv_rn(dx, g.csx);
        qy = g.exact_y ? __dmul_rn(dy, g.inv_csy) : __ddiv_rn(dy, g.csy);
    }
    int c = __double2int_rd(qx);          // floor + convert (cvt.rmi.s32.f64)
    int r = __double2int_rd(qy);
    c = max(0, min(c, g.width - 1));
    r = max(0, min(r, g.height - 1));
    col = c;
    row = r;
    return inside;
}

// TileRouter::assign tile id, src/engine/tile_router.cpp:112-121.
__device__ __forceinline__ int tile_of(const GridParams& g, int col, int row)
{
    // two integer divisions cost ~40 instructions per point; the reference's default tiles are 4096 x 4096
    const int tr = g.tile_h_shift >= 0 ? (row >> g.tile_h_shift) : (row / g.tile_h);
    const int tc = g.tile_w_shift >= 0 ? (col >> g.tile_w_shift) : (col / g.tile_w);
    return tr * g.tiles_x + tc;
}

// ---------------------------------------------------------------------------
// Accumulator records.
//
// The reference keeps each reduction's state band-sequential per tile
// (include/pcr/ops/builtin_ops.h:143-176).  Here all reductions that share a
// glyph are fused into ONE pass over the points, and their state words are
// interleaved into one record per cell:
//     [ add words (f32) | max words (ordered s32) | min words (ordered s32) | pad ]
// record width W in {1,2,4,8} 32-bit words, so that the additive words go out as
// one red.global.add.v2/v4.f32 and a random cell costs one 32-byte sector.
// Max/Min use the order-preserving float->s32 map so a single red.max/min.s32
// replaces the reference's CAS loop (src/engine/accumulator_kernels.cu:57-98).
// ---------------------------------------------------------------------------
constexpr int kMaxAdd = 4;      // additive words per pass
constexpr int kMaxExt = 2;      // max words, and min words, per pass
constexpr int kMaxChan = 4;     // distinct f32 value channels read per pass

__host__ __device__ __forceinline__ int32_t f32_ordered(float f)
{
#ifdef __CUDA_ARCH__
    const int32_t b = __float_as_int(f);
#else
    union { float f; int32_t i; } u; u.f = f; const int32_t b = u.i;
#endif
    return b >= 0 ? b : (b ^ 0x7fffffff);
}
__host__ __device__ __forceinline__ float ordered_f32(int32_t o)
{
    const int32_t b = o >= 0 ? o : (o ^ 0x7fffffff);
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    union { float f; int32_t i; } u; u.i = b; return u.f;
#endif
}

// red.global.* : fire-and-forget reductions resolved in L2 (no return value, so
// no round trip).  Vector forms need sm_90+ (PTX ISA 8.1).
__device__ __forceinline__ void red_add(float* p, float a)
{
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" :: "l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red_add2(float* p, float a, float b)
{
    asm volatile("red.relaxed.gpu.global.add.v2.f32 [%0], {%1, %2};"
                 :: "l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d)
{
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_max(int32_t* p, int32_t v)
{
    asm volatile("red.relaxed.gpu.global.max.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_min(int32_t* p, int32_t v)
{
    asm volatile("red.relaxed.gpu.global.min.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Additive part of a record: NADD words starting at rec[0].
template <int NADD>
__device__ __forceinline__ void red_add_words(float* rec, const float (&a)[kMaxAdd])
{
    if constexpr (NADD == 1) red_add(rec, a[0]);
    if constexpr (NADD == 2) red_add2(rec, a[0], a[1]);
    if constexpr (NADD == 3) { red_add2(rec, a[0], a[1]); red_add(rec + 2, a[2]); }
    if constexpr (NADD == 4) red_add4(rec, a[0], a[1], a[2], a[3]);
}

// Streaming loads: points are read exactly once, keep them out of L1 (the
// .L2::evict_first qualifier is only legal on 256-bit loads, so no L2 hint here).
__device__ __forceinline__ double2 ldg_stream_d2(const double* p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldg_stream_d(const double* p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg_stream_f2(const float* p)
{
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];"
                 : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream_f(const float* p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------
// Pass descriptors (host fills, kernels take by value).
// ---------------------------------------------------------------------------
struct PassLayout {
    int n_add, n_max, n_min;     // words of each kind
    int width;                   // record width in words: 1, 2, 4 or 8
    int n_chan;                  // distinct channels read
    int8_t add_src[kMaxAdd];     // channel index feeding add word j, or -1 = weight/count word
    int8_t max_src[kMaxExt];
    int8_t min_src[kMaxExt];
};

struct ChannelPtrs {
    const float* p[kMaxChan];
};

// Finalize program: how each output band is derived from a record
// (Op::finalize, include/pcr/ops/builtin_ops.h:16,29,42,55,68-70,99-101).
enum FinKind : int { FIN_SUM = 0, FIN_COUNT = 1, FIN_RATIO = 2, FIN_MAX = 3, FIN_MIN = 4 };
constexpr int kMaxBandsPerPass = 16;
struct FinalizeProgram {
    int n;                              // bands produced by this pass
    int kind[kMaxBandsPerPass];
    int word_a[kMaxBandsPerPass];       // value / max / min / count word
    int word_b[kMaxBandsPerPass];       // denominator word for FIN_RATIO
    int band[kMaxBandsPerPass];         // output band index
};

}  // namespace pcrb
