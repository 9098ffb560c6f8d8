// bin_kernels.cuh — interface of the tile-binning path (bin_kernels.cu).
#pragma once

#include "kernels.cuh"

namespace pcrb {

constexpr uint32_t kBinPageEntries = 4096;   // entries per page (= points per scatter chunk)
constexpr uint32_t kBinPadChunk = 2048;      // smallest scatter chunk: every bin's run of a chunk is padded to 16 bytes
constexpr int kMaxBins = 1024;
constexpr int kBinMaxChan = 2;               // value channels an entry can carry
inline int bin_entry_words(int n_chan) { return n_chan == 0 ? 1 : n_chan == 1 ? 2 : 4; }

// One rank's entry pool (device pointers; another rank's pool is the same struct over peer memory).
struct BinPool {
    // [pool_pages * kBinPageEntries] entries, array of structures: {cell u32} (no value channel),
    // {cell u32, value f32} (one channel) or {cell, v0, v1, pad} (two) — one 4 / 8 / 16-byte store per thread, so
    // a warp's copy-out is one contiguous 128 / 256 / 512-byte write (what the NVLink path wants:
    // the SoA layout this replaces moved 4-byte stores and ran at 175 GB/s per GPU at N = 8)
    uint32_t* ent;
    uint32_t* page_bin;                      // [pool_pages] bin a page belongs to
    uint32_t* page_fill;                     // [pool_pages] entries written to a page so far
    // Control words.  The SCATTER side allocates pages with atomicAdd(next_page): on one GPU next_page is word [0]
    // of the pool's own control block; in the tile-partitioned layout every source rank owns a fixed slice of
    // every owner's pool (pages [page_base, page_base + sub_pages)) and next_page is a counter in the SOURCE's
    // local memory — no remote atomic, whose round trip queues behind the posted stores on NVLink (measured: it
    // tripled the time of the exchange at 8 GPUs).  The FOLD side (the owner) reads how many pages each source
    // used from src_count[0..n_src) and uses ctrl[2] as its page cursor, ctrl[3] as the number of pages to fold.
    uint32_t* next_page;                     // scatter: allocation counter of this (source, owner) pair
    uint32_t* overflow;                      // set to 1 if the slice ran out (entries were dropped)
    uint32_t* ctrl;                          // owner's control block: [0] next (one GPU) [1] overflow [2] cursor [3] total
    uint32_t* src_count;                     // owner: pages used by each source, [n_src]
    uint32_t  page_base;                     // scatter: first page of my slice of this pool
    uint32_t  sub_pages;                     // pages per source slice (= pool_pages on one GPU)
    int       n_src;
    uint32_t  pool_pages;
};
constexpr int kBinCtrlWords = 4 + kMaxParts;  // ctrl[4 + r] = src_count[r] in the partitioned layout

// Where the scatter kernel sends a bin's entries.
struct BinTargets {
    BinPool pool[kMaxParts];                 // pool of the rank that owns the bin (pool[0] on one GPU)
    int bin_owner_shift;                     // < 0: everything is pool[0]
    uint32_t bins_per_owner;                 // owner = bin / bins_per_owner
    int shift;                               // bin = cell >> shift
    int nbins;
    uint32_t* open_page;                     // [scatter grid][nbins] page chains of this rank's scatter CTAs
    uint32_t* open_fill;
};

bool bin_supported(const PassLayout& L);
void bin_geometry(size_t cells, int record_words, int log2_req, int& shift, int& nbins);
void bin_owner_cells(size_t cells, int shift, int nbins, int world, int rank, uint32_t& bins_per_owner,
                     size_t& cell0, size_t& cell1);
uint64_t bin_capacity(uint64_t pages, int nbins, int n_chan);   // points that fit `pages` pages whatever their distribution
// scatter CTA = `threads` (256 or 512) threads staging 8 points each; 1024 threads per SM either way
constexpr int kBinThreadsLocal = 256;        // one GPU: 4 CTAs per SM (8.3 -> 6.9 ms per 1B points with the L2 prefetch)
constexpr int kBinThreadsPeer = 512;         // tile-partitioned layout: longer runs per bulk store over NVLink
                                             // (N=2, config 5: bin + exchange 4.99 ms with 512 threads, 6.25 ms with 256)
size_t bin_scatter_smem(int nbins, int n_chan, int threads);
unsigned bin_scatter_grid(int sm_count, int nbins, int n_chan, int threads);
// route n points once and append {cell, values} to the page chain of each point's bin
cudaError_t launch_bin_scatter(cudaStream_t s, const uint8_t* mask, const double* x, const double* y,
                               const ChannelPtrs& ch, size_t n, const GridParams& g, const PassLayout& L,
                               const BinTargets& bt, uint32_t* touched, unsigned grid, int threads);
// fold every pending entry of `pool` into `state` (record of global cell c at state[(c - cell_base) * W]),
// bin by bin, and empty the pool
cudaError_t launch_bin_flush(cudaStream_t s, const BinPool& pool, int nbins, uint32_t* bin_pages, uint32_t* bin_first,
                             uint32_t* order, uint32_t* state, size_t cell_base, const PassLayout& L,
                             uint32_t* open_page, size_t n_open, int sm_count);
cudaError_t launch_bin_reset(cudaStream_t s, const BinPool& pool, uint32_t* open_page, size_t n_open, int sm_count);

}  // namespace pcrb
