// engine_part.cu — N>1, tile-partitioned grid with an all-to-all point exchange (SURVEY §8e G1,
// "partitioning B"; BASELINE north_star: "or the grid is tile-partitioned with an all-to-all point
// exchange when it exceeds HBM").
//
// Every rank owns a contiguous range of bins (bin = 2^shift consecutive cells, bin_kernels.cu) and keeps
// accumulator records ONLY for the cells of its bins.  There is no partial grid and no reduce:
//
//   ingest    k_bin_scatter routes a rank's points and appends each entry {cell, value} to a page in the
//             pool of the rank that OWNS the entry's bin — local memory for its own bins, NVLink peer
//             memory (posted stores, a warp writes 256 contiguous bytes; one remote atomic per 4096-entry page) for the others.
//             This is the all-to-all, fused into the binning kernel: 8 B per point and channel cross the
//             fabric instead of the 20 B of the raw point, and nothing waits for anything.
//   finalize  every rank announces "my entries have landed" (flag store, system scope), waits for the same
//             announcement from everybody, folds the pages of its pool into its records bin by bin
//             (k_bin_accumulate), finalizes its cell range into the bands, and announces "my pool is empty
//             again".  Bands stay distributed (comm_root_only = 2) or are stored into the peers' band arrays
//             by the finalize kernel.
//
// The reference has no counterpart (single GPU); its closest relative is the tile-local accumulation of
// TileRouter::extract_batches + Accumulator (src/engine/tile_router.cpp:253-366).
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

bool Engine::partition_wanted() const
{
    if (comm_layout_ == 1 || passes_.empty()) return false;
    for (const Pass& p : passes_)
        if (!p.bin.on) return false;                 // only tile-binned Point passes can be partitioned
    return true;                                     // auto (0) and forced (2)
}

// Called from comm_init once peer memory is mapped (peer_ok_).  Collective.
Status Engine::partition_setup()
{
    // 1. my pools' buffers -> everybody
    std::vector<void*> mine;
    for (Pass& p : passes_) {
        BinPool& q = p.bin.pool;
        mine.push_back(q.ent);
        mine.push_back(q.page_bin);
        mine.push_back(q.page_fill);
        mine.push_back(q.ctrl);
    }
    std::vector<std::vector<void*>> all;
    ST_TRY(ipc_exchange(mine, all));

    // 2. ownership: contiguous ranges of bins; my records cover only my cells
    size_t at = 0;
    for (Pass& p : passes_) {
        BinState& b = p.bin;
        bin_owner_cells(cells_, b.shift, b.nbins, world_, rank_, b.bins_per_owner, b.cell0, b.cell1);
        for (int k = 0; k < world_; ++k) {
            BinPool q = b.pool;                      // same geometry (pool_pages) on every rank: same free-memory rule
            const std::vector<void*>& h = all[k];
            size_t j = at;
            q.ent = static_cast<uint32_t*>(h[j++]);
            q.page_bin = static_cast<uint32_t*>(h[j++]);
            q.page_fill = static_cast<uint32_t*>(h[j++]);
            q.ctrl = static_cast<uint32_t*>(h[j++]);
            b.peer_pool[k] = q;
        }
        at += 4;
        // records: only my cells.  (create() allocated the whole grid before the world size was known.)
        const size_t W = p.layout.width;
        CU_TRY(cudaStreamSynchronize(compute_));
        cudaFree(p.d_delta[0]);
        p.d_delta[0] = nullptr;
        CU_TRY(cudaMalloc(&p.d_delta[0], std::max<size_t>(b.cell1 - b.cell0, 1) * W * 4));
        p.d_state = p.d_delta[0];
        CU_TRY(launch_init_state(compute_, p.d_state, b.cell1 - b.cell0, p.layout));
    }
    // the pools' geometry must agree (pool_pages is derived from each rank's free memory)
    {
        uint32_t* d = nullptr;
        CU_TRY(cudaMalloc(&d, passes_.size() * 4));
        std::vector<uint32_t> pages;
        for (Pass& p : passes_) pages.push_back(p.bin.pool.pool_pages);
        CU_TRY(cudaMemcpyAsync(d, pages.data(), pages.size() * 4, cudaMemcpyHostToDevice, compute_));
        ST_TRY(nccl_allreduce_min_u32(d, pages.size()));
        CU_TRY(cudaMemcpyAsync(pages.data(), d, pages.size() * 4, cudaMemcpyDeviceToHost, compute_));
        CU_TRY(cudaStreamSynchronize(compute_));
        cudaFree(d);
        for (size_t i = 0; i < passes_.size(); ++i) {
            BinState& b = passes_[i].bin;
            // the chains were sized for the single-GPU scatter grid (more, smaller CTAs): a prefix of them is used here
            b.threads = kBinThreadsPeer;
            b.grid = std::min(b.grid, bin_scatter_grid(sm_count_, b.nbins, passes_[i].layout.n_chan, b.threads));
            const uint64_t chains = static_cast<uint64_t>(b.grid) * b.nbins;
            // Every source rank owns a fixed slice of every owner's pool and allocates pages from it with a counter
            // in its OWN memory: no remote atomic anywhere.  A slice must take all the points a rank may send
            // between two finalizes even if every one of them lands on one owner: capacity = what one slice holds.
            const uint32_t slice = pages[i] / static_cast<uint32_t>(world_);
            if (slice <= chains + 1)
                return Status::error(PCR_OUT_OF_MEMORY, "pipeline: tile-partitioned layout: the entry pool is too small to be "
                                                        "sliced over the ranks; raise bin_pool_points");
            if (!b.part_counters) CU_TRY(cudaMalloc(&b.part_counters, static_cast<size_t>(world_) * 4 * sizeof(uint32_t)));
            CU_TRY(cudaMemsetAsync(b.part_counters, 0, static_cast<size_t>(world_) * 4 * sizeof(uint32_t), compute_));
            for (int k = 0; k < world_; ++k) {
                BinPool& q = b.peer_pool[k];
                q.pool_pages = pages[i];
                q.sub_pages = slice;
                q.page_base = static_cast<uint32_t>(rank_) * slice;
                q.next_page = b.part_counters + static_cast<size_t>(k) * 4;     // LOCAL counter for (me, owner k)
                q.overflow = q.next_page + 1;
                q.n_src = world_;
                q.src_count = q.ctrl + 4;                                        // owner k's per-source page counts (peer memory)
            }
            b.pool.pool_pages = pages[i];                                        // my own pool, as the fold sees it
            b.pool.sub_pages = slice;
            b.pool.n_src = world_;
            b.pool.src_count = b.pool.ctrl + 4;
            b.capacity = bin_capacity(static_cast<uint64_t>(slice) - chains - 1, b.nbins, passes_[i].layout.n_chan);
        }
        CU_TRY(cudaStreamSynchronize(compute_));
    }
    if (!e_delta_) CU_TRY(cudaEventCreateWithFlags(&e_delta_, cudaEventDisableTiming));
    partition_ = true;
    delta_mode_ = false;
    return Status::success();
}

// N>1 partition mode: append a device-resident chunk to the owners' pools.
Status Engine::part_append(Pass& p, const uint8_t* mask, const double* dx, const double* dy, const ChannelPtrs& ch, size_t n)
{
    BinState& b = p.bin;
    if (b.pending + n > b.capacity)
        return Status::error(PCR_OUT_OF_MEMORY,
                             "pipeline: tile-partitioned layout: this rank sent more points since the last finalize than "
                             "its share of the owners' entry pools (" + std::to_string(b.capacity) +
                             "); call finalize() more often or raise bin_pool_points");
    // the owners' pools were emptied by their last finalize: nobody may append before that has happened
    if (part_waited_epoch_ != epoch_) {
        PeerFlags pf{};
        pf.n = world_; pf.rank = rank_;
        for (int k = 0; k < world_; ++k) pf.flags[k] = peer_[k].flags;
        if (epoch_ > 0) { CU_TRY(launch_peer_wait(compute_, pf, 1, epoch_)); ++launches_; }
        part_waited_epoch_ = epoch_;
    }
    BinTargets bt{};
    for (int k = 0; k < world_; ++k) bt.pool[k] = b.peer_pool[k];
    bt.bin_owner_shift = 0;
    bt.bins_per_owner = b.bins_per_owner;
    bt.shift = b.shift;
    bt.nbins = b.nbins;
    bt.open_page = b.open_page;
    bt.open_fill = b.open_fill;
    CU_TRY(launch_bin_scatter(compute_, mask, dx, dy, ch, n, gp_, p.layout, bt, d_touched_, b.grid, b.threads));
    ++launches_;
    b.pending += n;
    return Status::success();
}

Status Engine::finalize_multi_part()
{
    ++epoch_;
    PeerFlags pf{};
    pf.n = world_; pf.rank = rank_;
    PeerTouched pt{};
    pt.n = world_;
    for (int k = 0; k < world_; ++k) {
        pf.flags[k] = peer_[k].flags;
        pt.touched[k] = d_touched_stage_ + static_cast<size_t>(k) * n_tiles_;
    }
    // a rank that ingested nothing in this epoch has not waited for the previous one yet
    if (part_waited_epoch_ != epoch_ - 1 && epoch_ > 1) { CU_TRY(launch_peer_wait(compute_, pf, 1, epoch_ - 1)); ++launches_; }
    part_waited_epoch_ = epoch_ - 1;

    // ---- my entries have landed everywhere; my touched-tile flags go to everybody ----
    prof_begin(PROF_PUSH, compute_);
    {
        PushTargets tg{};
        for (int k = 0; k < world_; ++k) tg.touched_stage[k] = peer_[k].touched_stage;
        PartCounts pc{};
        pc.n_pass = static_cast<int>(std::min<size_t>(passes_.size(), 4));
        for (int i = 0; i < pc.n_pass; ++i) {
            BinState& b = passes_[i].bin;
            pc.local[i] = b.part_counters;
            pc.my_overflow[i] = b.pool.ctrl + 1;
            for (int k = 0; k < world_; ++k) pc.owner_count[i][k] = b.peer_pool[k].ctrl + 4 + rank_;
        }
        CU_TRY(launch_push_touched(compute_, d_touched_, n_tiles_, tg, pf, epoch_, pc));
        ++launches_;
    }
    prof_end(compute_);

    // ---- everybody's entries have landed here: fold my pool, finalize my cells ----
    prof_begin(PROF_ACC, compute_);
    CU_TRY(launch_peer_wait_merge_touched(compute_, pf, epoch_, pt, d_touched_merged_, std::max(1, n_tiles_), true));
    ++launches_;
    for (Pass& p : passes_) {
        BinState& b = p.bin;
        CU_TRY(launch_bin_flush(compute_, b.pool, b.nbins, b.bin_pages, b.bin_first, b.order, p.d_state, b.cell0, p.layout,
                                b.open_page, static_cast<size_t>(b.grid) * b.nbins, sm_count_));
        launches_ += 5;
        b.pending = 0;
    }
    prof_end(compute_);

    prof_begin(PROF_FIN, compute_);
    for (size_t i = 0; i < reductions_.size(); ++i)
        if (reductions_[i].rejected)
            CU_TRY(cudaMemsetAsync(d_out_ + i * cells_, 0xFF, cells_ * sizeof(float), compute_));
    OutTargets outs{};
    outs.out[outs.n++] = d_out_;
    for (int k = 0; k < world_; ++k)
        if (k != rank_ && !(gather_root_only_ && k != 0) && !bands_distributed_) outs.out[outs.n++] = peer_[k].out;
    for (Pass& p : passes_) {
        BinState& b = p.bin;
        StateParts parts{};
        parts.n = 1;
        parts.part[0] = p.d_state;
        CU_TRY(launch_finalize(compute_, parts, b.cell0, b.cell0, b.cell1 - b.cell0, outs, cells_, gp_, p.layout, p.fin,
                               d_touched_merged_));
        ++launches_;
    }
    // my pool is empty again (k_bin_reset ran inside the flush) and my band stores are issued
    CU_TRY(launch_peer_signal(compute_, pf, 1, epoch_));
    ++launches_;
    prof_end(compute_);
    return Status::success();
}

}  // namespace pcrb
