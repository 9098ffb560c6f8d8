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
    const uint32_t* part[kMaxParts + 1];   // peer mode: [owner-accumulated slice, rank 0's delta, rank 1's, ...]
    int n;
};

// Where a finalized slice is stored: the rank's own band array, plus (peer-memory path)
// the band arrays of other ranks mapped over NVLink.
struct OutTargets {
    float* out[kMaxParts];
    int n;
};

enum PointKernelVariant { POINT_DIRECT = 1, POINT_TMA = 2 };

// Point filter (FilterSpec, include/pcr/engine/filter.h:33-51): AND of up to kMaxPredicates
// comparisons on f32 channels, evaluated into a byte mask (1 = keep) in front of routing.
constexpr int kMaxPredicates = 8;
struct FilterProgram {
    int n;
    const float* chan[kMaxPredicates];      // channel arrays of the current chunk
    int op[kMaxPredicates];
    float value[kMaxPredicates];
    const float* set[kMaxPredicates];       // device arrays for InSet / NotInSet
    int set_size[kMaxPredicates];
};
// mask[i] = all predicates hold for point i; *survivors += number of kept points
cudaError_t launch_filter_mask(cudaStream_t s, const FilterProgram& fp, size_t n, uint8_t* mask,
                               unsigned long long* survivors);

// state identity fill (init_state_kernel<Op>, src/engine/grid_merge.cu:16-23)
cudaError_t launch_init_state(cudaStream_t s, uint32_t* state, size_t cells, const PassLayout& L);

// Point glyph: fused route + accumulate for every reduction of the pass
// (kernel_assign + kernel_accumulate_*, tile_router_kernels.cu:34-61,
//  accumulator_kernels.cu:31-133 — with the CPU routing rule).
cudaError_t launch_point_accumulate(cudaStream_t s, int variant, bool warp_aggregate, const uint8_t* mask,
                                    const double* x, const double* y, const ChannelPtrs& ch,
                                    size_t n, uint32_t* state, const GridParams& g,
                                    const PassLayout& L, uint32_t* touched, int sm_count);

// Line glyph (accumulate_glyph_line_cpu, src/engine/glyph_kernels.cu:188-281).
// `exact` (may be null): deterministic mode 2 — contributions go into the exact fixed-point state of
// exact_acc.cuh instead of float reductions on `state`.
struct XAcc;
cudaError_t launch_line_accumulate(cudaStream_t s, const uint8_t* mask, const double* x, const double* y,
                                   const ChannelPtrs& ch, const GlyphParams& gp, size_t n,
                                   uint32_t* state, const GridParams& g, const PassLayout& L,
                                   uint32_t* touched, const XAcc* exact = nullptr);

// Gaussian glyph (accumulate_glyph_gaussian_cpu, src/engine/glyph_kernels.cu:79-183)
cudaError_t launch_gaussian_accumulate(cudaStream_t s, const uint8_t* mask, const double* x, const double* y,
                                       const ChannelPtrs& ch, const GlyphParams& gp, size_t n,
                                       uint32_t* state, const GridParams& g, const PassLayout& L,
                                       uint32_t* touched, const XAcc* exact = nullptr);

// deterministic mode 2 (exact_kernels.cu)
size_t xacc_limb_bytes(size_t cells, const PassLayout& L);
size_t xacc_flag_bytes(size_t cells, const PassLayout& L);
size_t xacc_ext_bytes(size_t cells, const PassLayout& L);
cudaError_t launch_exact_init(cudaStream_t s, const XAcc& xa, const PassLayout& L);
cudaError_t launch_point_exact(cudaStream_t s, const uint8_t* mask, const double* x, const double* y,
                               const ChannelPtrs& ch, size_t n, const XAcc& xa, const GridParams& g,
                               const PassLayout& L, uint32_t* touched);
cudaError_t launch_finalize_exact(cudaStream_t s, const XAcc& xa, size_t cell0, size_t count, float* out,
                                  size_t band_stride, const GridParams& g, const PassLayout& L,
                                  const FinalizeProgram& fp, const uint32_t* touched);

// merge (Op::merge over parts) + finalize (Op::finalize) + touched-tile NaN rule,
// for cells [cell0, cell0+count); writes out[band*band_stride + cell].
cudaError_t launch_finalize(cudaStream_t s, const StateParts& parts, size_t part_cell0,
                            size_t cell0, size_t count, const OutTargets& out, size_t band_stride,
                            const GridParams& g, const PassLayout& L, const FinalizeProgram& fp,
                            const uint32_t* touched);

// ---- peer-memory combine (one process per GPU, buffers mapped with CUDA IPC) ----
// flags layout on every rank: flags[phase * kMaxParts + writer_rank], written by the peers.
struct PeerFlags {
    uint32_t* flags[kMaxParts];      // every rank's flag array (own one included)
    int n, rank;
};
struct PeerTouched {
    const uint32_t* touched[kMaxParts];
    int n;
};
// Peer handshake fused into the finalize kernel (N>1, peer-memory path).
struct PeerSync {
    PeerFlags pf;
    PeerTouched pt;
    uint32_t epoch;
    int signal_begin;                // this launch announces "my state is complete" (first pass)
    int signal_end;                  // the last CTA of this launch announces "I am done with your
                                     // state and your bands hold my slice" (last pass)
    unsigned int* done_counter;      // CTAs finished so far (self-resetting)
    int waited;                      // a k_peer_wait* launch ahead of this one on the stream already
                                     // observed the flags: the CTAs do not poll them again
};
// `accum` (may be null): the owner's accumulated slice; the merged record of every cell is stored back
// into it (parts.part[0] then normally points at the same memory), so that the ranks only ever ship the
// DELTA they accumulated since the previous finalize.
cudaError_t launch_finalize_peer(cudaStream_t s, const StateParts& parts, size_t part_cell0, size_t cell0,
                                 size_t count, const OutTargets& out, size_t band_stride,
                                 const GridParams& g, const PassLayout& L, const FinalizeProgram& fp,
                                 const PeerSync& ps, uint32_t* accum, int sm_count);

// Push phase of the peer-memory combine: every record of `state` that belongs to another rank's
// row slice is stored (posted NVLink writes, no round trip) into that rank's combine buffer, slot
// `rank`; the touched-tile flags go to every rank's touched staging, slot `rank`.  When
// `signal` is set the last CTA releases phase 0 ("my contribution has landed") on every rank.
struct PushTargets {
    uint32_t* combined[kMaxParts];   // rank k's combine buffer for this pass (kMaxParts slots of max_slice cells)
    uint32_t* touched_stage[kMaxParts];
    int rows_per;                    // rows per slice (ceil(height / world))
    size_t max_slice_cells;
};
// `reset`: the pushed state is a delta buffer — every record (and touched flag) is put back to the
// identity behind the copy, ready to take the ingests of the finalize after next.
cudaError_t launch_push_slices(cudaStream_t s, uint32_t* state, uint32_t* touched, int n_tiles,
                               const GridParams& g, const PassLayout& L, const PushTargets& pt,
                               const PeerSync& ps, bool push_touched, bool signal, bool reset, int sm_count);

// partition mode: my touched-tile flags into everyone's staging (slot = my rank), then phase 0 of `epoch`
// on every rank ("everything this stream wrote into your memory so far has landed")
// ... and, per binned pass, how many pages of my slice of every owner's pool I used (local counters -> the owners'
// src_count words; the local counters go back to zero; my overflow flags are folded into my own pool's flag)
struct PartCounts {
    int n_pass;
    uint32_t* local[4];                  // [world][4] local counters of pass i
    uint32_t* my_overflow[4];            // my own pool's overflow word
    uint32_t* owner_count[4][kMaxParts]; // where rank k keeps my page count
};
cudaError_t launch_push_touched(cudaStream_t s, const uint32_t* touched, int n_tiles, const PushTargets& pt,
                                const PeerFlags& pf, uint32_t epoch, const PartCounts& pc);
// store `epoch` into slot (phase, my rank) of every rank's flag array (system-scope release)
cudaError_t launch_peer_signal(cudaStream_t s, const PeerFlags& pf, int phase, uint32_t epoch);
// wait until every rank's slot of `phase` in MY flag array has reached `epoch`
cudaError_t launch_peer_wait(cudaStream_t s, const PeerFlags& pf, int phase, uint32_t epoch);
// wait (phase 0), then OR the touched-tile flags of all ranks into `merged`
// (`cumulative`: OR into `merged` instead of overwriting it — the ranks pushed touched-flag deltas)
cudaError_t launch_peer_wait_merge_touched(cudaStream_t s, const PeerFlags& pf, uint32_t epoch,
                                           const PeerTouched& pt, uint32_t* merged, int n_tiles, bool cumulative);

}  // namespace pcrb
