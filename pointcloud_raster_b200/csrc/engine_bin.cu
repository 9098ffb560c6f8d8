// engine_bin.cu — host side of the tile-binning path (kernels: bin_kernels.cu).
//
// When a Point pass runs on a grid whose records are far larger than L2 (BASELINE config 5: 20000^2
// cells, 6.4 GB of records), k_point_direct pays a DRAM read-modify-write of a 32-byte sector for
// every point (profiles/r02_c5_point_direct_ncu_full.json).  The binned path replaces
// "route + reduce now" by "route + append now, reduce bin by bin later": the role of the reference's
// TileRouter::sort / extract_batches (src/engine/tile_router.cpp:138-366) without a sort.
#include "engine.h"

#include <algorithm>
#include <cstring>

namespace pcrb {

#define CU_TRY(expr)                                                                     \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return Status::error(PCR_CUDA_ERROR, std::string("CUDA error: ") +           \
                                 cudaGetErrorString(_e) + " (" #expr ")");               \
    } while (0)
#define ST_TRY(expr) do { Status _s = (expr); if (!_s.ok()) return _s; } while (0)

// Decide whether pass `p` is binned and allocate its pool.  Called from alloc_state().
Status Engine::bin_setup(Pass& p)
{
    BinState& b = p.bin;
    b.on = false;
    if (p.glyph.type != PCR_GLYPH_POINT || deterministic_ || exact_ || !bin_supported(p.layout)) return Status::success();
    const size_t W = p.layout.width;
    const size_t state_bytes = cells_ * W * 4;
    const bool forced = point_kernel_knob_ == 3;
    // auto: only where the records cannot live in L2 (126 MB) anyway
    if (!forced && !(point_kernel_knob_ == 0 && state_bytes >= (size_t(256) << 20))) return Status::success();

    bin_geometry(cells_, static_cast<int>(W), bin_cells_log2_, b.shift, b.nbins);
    b.threads = kBinThreadsLocal;
    b.grid = bin_scatter_grid(sm_count_, b.nbins, p.layout.n_chan, b.threads);
    const size_t chains = static_cast<size_t>(b.grid) * b.nbins;

    // pool size: every chain ends in one partly filled page, all other pages are full, so T points need at
    // most T / P + chains pages whatever their distribution
    const size_t entry_bytes = 4 * static_cast<size_t>(bin_entry_words(p.layout.n_chan));
    size_t want = bin_pool_points_;
    if (want == 0) {
        size_t free_b = 0, total_b = 0;
        CU_TRY(cudaMemGetInfo(&free_b, &total_b));
        want = std::min<size_t>(size_t(1) << 31, free_b / 4 / entry_bytes);
    }
    want = std::max<size_t>(want, kBinPageEntries);
    // a run is padded to 16 bytes with null entries (at most align - 1 per bin and scatter chunk): size the
    // pool so that `want` POINTS fit whatever their distribution
    const size_t pad_align = 4 / static_cast<size_t>(bin_entry_words(p.layout.n_chan));
    const size_t max_pages = (size_t(1) << 32) / kBinPageEntries;                // entry indices are 32-bit
    size_t pages = 0;
    for (;;) {
        const size_t slots = (want + kBinPadChunk - 1) / kBinPadChunk * (kBinPadChunk + static_cast<size_t>(b.nbins) * (pad_align - 1));
        pages = (slots + kBinPageEntries - 1) / kBinPageEntries + chains + 1;
        if (pages < max_pages) break;
        if (bin_pool_points_ != 0 || want <= kBinPageEntries)
            return Status::error(PCR_INVALID_ARGUMENT, "pipeline: bin_pool_points too large");
        want -= want / 8;                                                        // auto: the largest pool that can be indexed
    }
    b.pool.pool_pages = static_cast<uint32_t>(pages);
    b.capacity = bin_capacity(pages - chains - 1, b.nbins, p.layout.n_chan);
    if (b.capacity < kBinPageEntries)
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: bin_pool_points too small");
    const size_t entries = pages * kBinPageEntries;
    CU_TRY(cudaMalloc(&b.pool.ent, entries * entry_bytes));
    CU_TRY(cudaMalloc(&b.pool.page_bin, pages * 4));
    CU_TRY(cudaMalloc(&b.pool.page_fill, pages * 4));
    CU_TRY(cudaMalloc(&b.pool.ctrl, kBinCtrlWords * sizeof(uint32_t)));
    CU_TRY(cudaMemsetAsync(b.pool.ctrl, 0, kBinCtrlWords * sizeof(uint32_t), compute_));
    b.pool.next_page = b.pool.ctrl;              // one GPU: the pool is one slice, its counter is ctrl[0]
    b.pool.overflow = b.pool.ctrl + 1;
    b.pool.src_count = b.pool.ctrl;
    b.pool.n_src = 1;
    b.pool.page_base = 0;
    b.pool.sub_pages = static_cast<uint32_t>(pages);
    CU_TRY(cudaMalloc(&b.bin_pages, static_cast<size_t>(b.nbins) * 4));
    CU_TRY(cudaMalloc(&b.bin_first, static_cast<size_t>(b.nbins) * 4));
    CU_TRY(cudaMemsetAsync(b.bin_pages, 0, static_cast<size_t>(b.nbins) * 4, compute_));
    CU_TRY(cudaMalloc(&b.order, pages * 4));
    CU_TRY(cudaMalloc(&b.open_page, chains * 4));
    CU_TRY(cudaMalloc(&b.open_fill, chains * 4));
    CU_TRY(cudaMemsetAsync(b.open_page, 0xFF, chains * 4, compute_));
    CU_TRY(cudaMemsetAsync(b.open_fill, 0, chains * 4, compute_));
    if (!h_overflow_) CU_TRY(cudaMallocHost(&h_overflow_, sizeof(uint32_t) * 16));
    b.pending = 0;
    b.on = true;
    return Status::success();
}

// bin = 2^shift consecutive cells; auto (log2_req <= 0): the records of a bin fill at most 64 MB; never more
// than kMaxBins bins
void bin_geometry(size_t cells, int record_words, int log2_req, int& shift, int& nbins)
{
    shift = log2_req;
    if (shift <= 0) {
        shift = 4;
        while ((size_t(2) << shift) * static_cast<size_t>(record_words) * 4 <= (size_t(64) << 20)) ++shift;
    }
    shift = std::max(4, std::min(shift, 31));
    while (((cells + (size_t(1) << shift) - 1) >> shift) > static_cast<size_t>(kMaxBins)) ++shift;
    nbins = static_cast<int>((cells + (size_t(1) << shift) - 1) >> shift);
}

// tile-partitioned layout: rank k owns bins [k * per, (k + 1) * per), per = ceil(nbins / world), i.e. the
// row-major cell range [cell0, cell1)
void bin_owner_cells(size_t cells, int shift, int nbins, int world, int rank, uint32_t& bins_per_owner,
                     size_t& cell0, size_t& cell1)
{
    bins_per_owner = static_cast<uint32_t>((nbins + world - 1) / world);
    const size_t bin0 = std::min<size_t>(nbins, static_cast<size_t>(rank) * bins_per_owner);
    const size_t bin1 = std::min<size_t>(nbins, static_cast<size_t>(rank + 1) * bins_per_owner);
    cell0 = std::min(cells, bin0 << shift);
    cell1 = std::min(cells, bin1 << shift);
}

uint64_t bin_capacity(uint64_t pages, int nbins, int n_chan)
{
    const uint64_t align = 4 / static_cast<uint64_t>(bin_entry_words(n_chan));
    const uint64_t slots = pages * kBinPageEntries;
    return slots / (kBinPadChunk + static_cast<uint64_t>(nbins) * (align - 1)) * kBinPadChunk;
}

void Engine::bin_free(Pass& p)
{
    BinState& b = p.bin;
    cudaFree(b.pool.ent);
    cudaFree(b.pool.page_bin); cudaFree(b.pool.page_fill); cudaFree(b.pool.ctrl); cudaFree(b.part_counters);
    cudaFree(b.bin_pages); cudaFree(b.bin_first); cudaFree(b.order); cudaFree(b.open_page); cudaFree(b.open_fill);
    b = BinState{};
}

// Append one device-resident chunk.  Folds the pool first when the chunk would not fit.
Status Engine::bin_append(Pass& p, const uint8_t* mask, const double* dx, const double* dy, const ChannelPtrs& ch, size_t n)
{
    BinState& b = p.bin;
    if (partition_) return part_append(p, mask, dx, dy, ch, n);
    size_t done = 0;
    while (done < n) {
        if (b.pending >= b.capacity) ST_TRY(bin_flush(p));
        const size_t cnt = std::min<size_t>(n - done, b.capacity - b.pending);
        if (cnt == 0) return Status::error(PCR_OUT_OF_MEMORY, "pipeline: the tile-binning entry pool cannot take any point");
        ChannelPtrs c2 = ch;
        for (int c = 0; c < p.layout.n_chan; ++c) c2.p[c] = ch.p[c] + done;
        BinTargets bt{};
        bt.pool[0] = b.pool;
        bt.bin_owner_shift = -1;
        bt.bins_per_owner = static_cast<uint32_t>(b.nbins);
        bt.shift = b.shift;
        bt.nbins = b.nbins;
        bt.open_page = b.open_page;
        bt.open_fill = b.open_fill;
        CU_TRY(launch_bin_scatter(compute_, mask ? mask + done : nullptr, dx + done, dy + done, c2, cnt, gp_, p.layout, bt,
                                  d_touched_, b.grid, b.threads));
        ++launches_;
        b.pending += cnt;
        done += cnt;
    }
    return Status::success();
}

Status Engine::bin_flush(Pass& p)
{
    BinState& b = p.bin;
    if (!b.on || b.pending == 0) return Status::success();
    prof_begin(PROF_ACC, compute_);
    CU_TRY(launch_bin_flush(compute_, b.pool, b.nbins, b.bin_pages, b.bin_first, b.order, p.d_state, 0, p.layout,
                            b.open_page, static_cast<size_t>(b.grid) * b.nbins, sm_count_));
    launches_ += 5;
    prof_end(compute_);
    b.pending = 0;
    return Status::success();
}

Status Engine::bin_flush_all()
{
    if (partition_) return Status::success();      // folded inside finalize_multi_part, behind the peers' flags
    for (Pass& p : passes_) ST_TRY(bin_flush(p));
    return Status::success();
}

}  // namespace pcrb
