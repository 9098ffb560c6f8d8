// kernels.cuh — host-callable launchers of the sm_100a kernels (internal C++ API
// between engine.cu and the kernel translation units).
#pragma once

#include "common.cuh"

namespace pcrb {

// Per-point glyph parameters (GlyphSpec, include/pcr/engine/glyph.h:19-43);
// channel pointers are device pointers for the current chunk, nullptr => default.
struct GlyphParams {
    const float* direction;   float default_direction;
    const float* half_length; float default_half_length;
    const float* sigma_x;     float default_sigma_x;
    const float* sigma_y;     float default_sigma_y;
    const float* rotation;    float default_rotation;
    float        max_radius_cells;
};

// Up to 8 partial states merged in slot order by the finalize kernel
// (multi-GPU combine; 1 part on a single GPU).
constexpr int kMaxParts = 8;
struct StateParts {
    const uint32_t* part[kMaxParts];
    int n;
};

enum PointKernelVariant { POINT_DIRECT = 1, POINT_TMA = 2 };

// state identity fill (init_state_kernel<Op>, src/engine/grid_merge.cu:16-23)
cudaError_t launch_init_state(cudaStream_t s, uint32_t* state, size_t cells, const PassLayout& L);

// Point glyph: fused route + accumulate for every reduction of the pass
// (kernel_assign + kernel_accumulate_*, tile_router_kernels.cu:34-61,
//  accumulator_kernels.cu:31-133 — with the CPU routing rule).
cudaError_t launch_point_accumulate(cudaStream_t s, int variant, bool warp_aggregate,
                                    const double* x, const double* y, const ChannelPtrs& ch,
                                    size_t n, uint32_t* state, const GridParams& g,
                                    const PassLayout& L, uint32_t* touched, int sm_count);

// Line glyph (accumulate_glyph_line_cpu, src/engine/glyph_kernels.cu:188-281)
cudaError_t launch_line_accumulate(cudaStream_t s, const double* x, const double* y,
                                   const ChannelPtrs& ch, const GlyphParams& gp, size_t n,
                                   uint32_t* state, const GridParams& g, const PassLayout& L,
                                   uint32_t* touched);

// Gaussian glyph (accumulate_glyph_gaussian_cpu, src/engine/glyph_kernels.cu:79-183)
cudaError_t launch_gaussian_accumulate(cudaStream_t s, const double* x, const double* y,
                                       const ChannelPtrs& ch, const GlyphParams& gp, size_t n,
                                       uint32_t* state, const GridParams& g, const PassLayout& L,
                                       uint32_t* touched);

// merge (Op::merge over parts) + finalize (Op::finalize) + touched-tile NaN rule,
// for cells [cell0, cell0+count); writes out[band*band_stride + cell].
cudaError_t launch_finalize(cudaStream_t s, const StateParts& parts, size_t part_cell0,
                            size_t cell0, size_t count, float* out, size_t band_stride,
                            const GridParams& g, const PassLayout& L, const FinalizeProgram& fp,
                            const uint32_t* touched);

}  // namespace pcrb
