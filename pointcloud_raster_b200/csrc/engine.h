// engine.h — C++ host side of the B200 ingest/finalize path (one GPU per Engine).
//
// Mirrors pcr::Pipeline::Impl (src/engine/pipeline.cpp:31-57,92-281) in role, not
// in structure: there is no TileRouter / TileManager / MemoryPool / Accumulator
// object chain.  State lives whole-grid in HBM as interleaved records for the life
// of the pipeline; all reductions sharing a glyph are planned into one fused pass;
// host clouds stream through a pinned staging ring on a copy stream while kernels
// run on the compute stream (the replacement for to_device_async + Hybrid mode).
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/pcr_b200.h"
#include "kernels.cuh"
#include "bin_kernels.cuh"
#include "exact_acc.cuh"

namespace pcrb {

struct Status {
    int code = PCR_OK;
    std::string message;
    bool ok() const { return code == PCR_OK; }
    static Status success() { return {}; }
    static Status error(int c, std::string m) { return {c, std::move(m)}; }
};

struct GlyphSpecHost {
    int type = PCR_GLYPH_POINT;
    std::string direction_channel, half_length_channel, sigma_x_channel, sigma_y_channel,
        rotation_channel;
    float default_direction = 0.f, default_half_length = 1.f, default_sigma_x = 1.f,
          default_sigma_y = 1.f, default_rotation = 0.f, max_radius_cells = 32.f;
    bool same_footprint(const GlyphSpecHost& o) const;
};

struct ReductionHost {
    std::string value_channel;
    int type = PCR_SUM;
    std::string band_name;
    GlyphSpecHost glyph;
    bool rejected = false;   // glyph + Max/Min: NotImplemented at ingest (pipeline.cpp:500-508)
};

struct FilterPredHost {
    std::string channel;
    int op = 0;
    float value = 0.f;
    std::vector<float> set;
};

// Tile binning of a Point pass on a grid whose records do not fit L2 (bin_kernels.cu): entries are
// appended to per-bin page chains at ingest time and folded into the records bin by bin at finalize
// (or when the pool is full).
struct BinState {
    bool on = false;
    int shift = 0, nbins = 0;
    BinPool pool{};
    uint32_t *bin_pages = nullptr, *bin_first = nullptr, *order = nullptr;
    uint32_t *open_page = nullptr, *open_fill = nullptr;   // [grid][nbins] chains of this rank's scatter CTAs
    unsigned grid = 0;
    int threads = 0;                // scatter CTA size (kBinThreadsLocal; kBinThreadsPeer in the tile-partitioned layout)
    uint64_t pending = 0;           // points appended since the last flush (upper bound of entries)
    uint64_t capacity = 0;          // points the pool takes whatever their distribution over the bins
    // N>1, tile-partitioned layout (engine_part.cu): owner of a bin = bin / bins_per_owner; this rank holds
    // records for cells [cell0, cell1) only and appends to the owners' pools over peer memory
    uint32_t bins_per_owner = 0;
    size_t cell0 = 0, cell1 = 0;
    BinPool peer_pool[kMaxParts] = {};
    uint32_t* part_counters = nullptr;   // [world][4] local: {pages I took from my slice of owner k's pool, overflow, -, -}
};

// One fused pass: every reduction in it shares the glyph footprint.
struct Pass {
    GlyphSpecHost glyph;
    PassLayout layout{};
    FinalizeProgram fin{};
    std::vector<std::string> channels;   // distinct value channels, index = ChannelPtrs slot
    uint32_t* d_state = nullptr;         // cells * layout.width words: what the accumulate kernels update
    uint32_t* d_combined = nullptr;      // multi-GPU: received peer slices
    // multi-GPU peer mode ("delta epochs"): d_state alternates between two full-grid DELTA buffers — the
    // ingests since the previous finalize — while the push of the other one is still in flight; the rank
    // that owns a row slice keeps the running merge of everybody's deltas in d_owned.
    uint32_t* d_delta[2] = {nullptr, nullptr};
    uint32_t* d_owned = nullptr;         // my row slice, accumulated over all finalizes so far
    BinState bin;
    // deterministic mode 2: exact fixed-point state (exact_acc.cuh); xa_sum = all ranks' states added up
    XAcc xa{}, xa_sum{};
};

// Worker pool for the pageable -> pinned staging copies of one host ingest.
// The ingest is cut into pieces (one segment range of one ring chunk each); workers
// pull pieces in order from an atomic cursor and copy them with non-temporal stores.
// The ingesting thread never copies (unless the pool has no workers): it only gates
// which chunks may be overwritten (their slot's previous DMA finished) and issues the
// H2D + kernels of a chunk once all of its pieces are staged.  One wake-up per ingest.
class CopyPool {
public:
    struct Piece { char* dst; const char* src; size_t bytes; uint32_t chunk; };

    explicit CopyPool(int threads);
    ~CopyPool();
    int workers() const { return static_cast<int>(workers_.size()); }

    // pieces must be ordered by chunk; pieces_in_chunk[k] = number of pieces of chunk k
    void begin(std::vector<Piece>&& pieces, const std::vector<uint32_t>& pieces_in_chunk, bool wake);
    void set_writable(size_t chunks) { writable_.store(chunks, std::memory_order_release); }
    bool staged(size_t chunk) const
    {
        return staged_[chunk].load(std::memory_order_acquire) == pieces_in_chunk_[chunk];
    }
    bool help() { return drain(false); }   // the caller copies one ready piece; false if none was
    void end();                             // every piece is staged: release the workers
    void abort();                           // error path: hand out no more pieces, then end()

    static void stage_copy(void* dst, const void* src, size_t bytes);

private:
    void worker();
    bool drain(bool until_done);
    std::vector<std::thread> workers_;
    std::vector<Piece> pieces_;
    std::vector<uint32_t> pieces_in_chunk_;
    std::unique_ptr<std::atomic<uint32_t>[]> staged_;
    size_t staged_capacity_ = 0;
    std::atomic<size_t> cursor_{0};
    std::atomic<size_t> writable_{0};
    std::mutex mu_;
    std::condition_variable cv_start_, cv_done_;
    uint64_t generation_ = 0;
    int active_ = 0;
    bool job_open_ = false;
    bool stop_ = false;
};

struct NcclApi;   // dlopen'ed subset of NCCL (engine_ext.cu)

// device scratch of the Gaussian gather path (gauss_gather.cu)
struct GaussScratch {
    uint32_t *keys = nullptr, *keys_alt = nullptr, *idx = nullptr, *idx_alt = nullptr;
    void* sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    uint32_t* records = nullptr;
    int* aux = nullptr;              // {max footprint radius, tile counter}
    size_t capacity = 0, record_bytes = 0;
};
bool gauss_gather_supported(const PassLayout& L);
size_t gauss_record_bytes(const PassLayout& L);
size_t gauss_sort_temp_bytes(size_t n, const GridParams& g);
cudaError_t launch_gaussian_gather(cudaStream_t s, const uint8_t* mask, const double* x, const double* y, const ChannelPtrs& ch,
                                   const GlyphParams& gp, size_t n, uint32_t* state, const GridParams& g,
                                   const PassLayout& L, uint32_t* touched, GaussScratch& sc, int sm_count, bool bin_mma);
bool gauss_binmma_supported(const PassLayout& L, float max_radius_cells, bool rotated);

class Engine {
public:
    static Status create(const pcr_pipeline_desc& desc, Engine** out);
    ~Engine();

    Status validate() const;
    Status ingest(const double* x, const double* y, size_t n, const pcr_channel_view* chans,
                  int nchans, int location);
    Status finalize(bool to_host);
    Status result_band(int band, const float** data, int* rows, int* cols, bool device);
    Status band_name(int band, std::string& out) const;
    Status stats(pcr_progress& out);
    void set_progress(pcr_progress_fn fn, void* user) { progress_fn_ = fn; progress_user_ = user; }
    Status reset();
    Status save_state(const std::string& dir);
    Status load_state(const std::string& dir);
    Status synchronize();

    Status profile_enable(int period);      // 0 = off, N = sample every Nth kernel group of each kind
    Status profile_reset();
    Status profile_read(pcr_profile& out);
    Status timer_begin();
    Status timer_end(double& ms);

    Status comm_init(const void* id128, int rank, int world);
    Status comm_barrier();
    void owned_cells(uint64_t& c0, uint64_t& c1) const;

private:
    Engine() = default;
    Status init(const pcr_pipeline_desc& desc);
    Status plan();
    Status alloc_state();
    Status init_state();
    Status run_passes(const double* dx, const double* dy, size_t n,
                      const std::vector<const float*>& chan_ptrs);
    Status run_passes_deterministic(const double* dx, const double* dy, size_t n,
                                    const std::vector<const float*>& chan_ptrs, const uint8_t* mask);
    Status ingest_device(const double* x, const double* y, size_t n,
                         const std::vector<const float*>& chan_ptrs);
    Status ingest_host(const double* x, const double* y, size_t n,
                       const std::vector<const float*>& chan_ptrs, bool pinned);
    Status bin_setup(Pass& p);            // decide + allocate the entry pool of a Point pass
    Status bin_append(Pass& p, const uint8_t* mask, const double* dx, const double* dy, const ChannelPtrs& ch, size_t n);
    Status bin_flush(Pass& p);            // fold the pending entries into the records (compute stream)
    Status bin_flush_all();
    void bin_free(Pass& p);
    Status finalize_single();
    Status finalize_multi();
    Status finalize_multi_nccl();
    Status finalize_multi_peer();
    Status finalize_multi_exact();
    Status alloc_exact(Pass& p);
    Status run_passes_exact(const double* dx, const double* dy, size_t n, const std::vector<const float*>& cp,
                            const uint8_t* mask);
    Status peer_map();               // exchange CUDA IPC handles, map every rank's buffers
    void peer_unmap();
    void peer_close_handles();
    Status ipc_exchange(const std::vector<void*>& mine, std::vector<std::vector<void*>>& all);
    Status nccl_allreduce_min_u32(uint32_t* d, size_t count);
    bool partition_wanted() const;
    Status partition_setup();
    Status part_append(Pass& p, const uint8_t* mask, const double* dx, const double* dy, const ChannelPtrs& ch, size_t n);
    Status finalize_multi_part();
    int channel_slot(const std::string& name);

    // profiling helpers
    enum ProfKind { PROF_ACC = 0, PROF_SORT = 1, PROF_FIN = 2, PROF_INIT = 3, PROF_PUSH = 4, PROF_KINDS = 5 };
    void prof_begin(ProfKind k, cudaStream_t s);
    void prof_end(cudaStream_t s);
    Status prof_collect();

    // ---- configuration ----
    pcr_grid_desc grid_{};
    GridParams gp_{};
    std::vector<ReductionHost> reductions_;
    int exec_mode_ = PCR_EXEC_AUTO;
    int device_ = 0;
    int sm_count_ = 148;
    bool deterministic_ = false;     // mode 1: sort by cell, in-order segmented reduce (bit-reproducible run to run)
    bool exact_ = false;             // mode 2: exact fixed-point accumulation (independent of order, chunking, sharding)
    bool async_device_ingest_ = false;
    int point_variant_ = POINT_DIRECT;
    int point_kernel_knob_ = 0;       // pcr_pipeline_desc::point_kernel as given (3 = force the binned path)
    int bin_cells_log2_ = 0;          // 0 = auto (records of one bin <= 64 MB)
    uint64_t bin_pool_points_ = 0;    // 0 = auto (from free HBM)
    uint32_t* h_overflow_ = nullptr;  // pinned read-back of the pools' overflow flags
    bool warp_aggregate_ = true;
    size_t cells_ = 0;
    int n_tiles_ = 0;

    // ---- point filter (FilterSpec) ----
    std::vector<FilterPredHost> filter_;
    float* d_filter_sets_ = nullptr;            // all InSet value lists, concatenated
    std::vector<size_t> filter_set_off_;
    uint8_t* d_mask_ = nullptr;
    size_t mask_capacity_ = 0;
    unsigned long long* d_survivors_ = nullptr; // running count of points that passed the filter
    Status build_mask(size_t n, const std::vector<const float*>& chan_ptrs, const uint8_t** mask);

    // ---- plan / state ----
    std::vector<Pass> passes_;
    std::vector<std::string> all_channels_;   // every channel any pass reads (value + glyph)
    uint32_t* d_touched_ = nullptr;           // what the accumulate kernels set (= d_touched_buf_[cur_])
    uint32_t* d_touched_buf_[2] = {nullptr, nullptr};
    bool delta_mode_ = false;                 // peer mode: states and touched flags are per-finalize deltas
    int cur_ = 0;
    cudaEvent_t e_delta_ = nullptr;           // compute stream: the delta of this epoch is complete
    uint32_t* d_touched_all_ = nullptr;       // multi-GPU merged flags
    uint32_t* d_touched_merged_ = nullptr;   // peer mode: OR of every rank's touched-tile flags
    float* d_out_ = nullptr;                  // [bands][cells]
    float* h_out_ = nullptr;                  // pinned, [bands][cells]
    bool finalized_ = false;

    // ---- streams / ring ----
    cudaStream_t compute_ = nullptr, copy_ = nullptr;
    // multi-GPU finalize: the merge/finalize kernels run on their own stream behind the push, so the
    // next ingest's kernels (compute stream) overlap the wait for the peers
    cudaStream_t fin_ = nullptr;
    cudaStream_t push_ = nullptr;     // N>1 delta epochs: the slice push, between the ingest and the merge streams
    cudaEvent_t e_pushed_ = nullptr, e_fin_ = nullptr;
    bool fin_pending_ = false;
    int band_copy_ = 0;               // pcr_pipeline_desc::comm_band_copy
    Status join_fin();                // make the compute stream wait for the last merge/finalize
    // Ingest ring: pinned staging buffers (one staged chunk each) and device buffers (one kernel
    // group each: several staged chunks, or one direct-DMA chunk out of pinned caller memory).
    struct HostSlot {
        char* h = nullptr;
        cudaEvent_t h2d_done = nullptr;
        bool in_flight = false;      // h2d_done guards a DMA that may still be reading `h`
    };
    struct DevSlot {
        char* d = nullptr;
        cudaEvent_t filled = nullptr, kernel_done = nullptr;
        bool used = false;           // kernel_done has been recorded at least once
    };
    std::vector<HostSlot> host_ring_;
    std::vector<DevSlot> dev_ring_;
    size_t host_pos_ = 0, dev_pos_ = 0;   // next slots; carry over from one ingest to the next
    size_t slot_points_ = 0;         // staged (pageable) chunk
    size_t direct_points_ = 0;       // chunk of a direct DMA out of pinned caller memory
    size_t group_chunks_ = 1;        // staged chunks per kernel group
    size_t slot_bytes_ = 0;
    CopyPool* pool_ = nullptr;
    int staging_threads_ = 0;
    bool staging_auto_ = true;       // not set by the caller: re-derived from the local world size at comm_init
    Status ensure_ring();

    // deterministic-mode scratch
    void* d_sort_tmp_ = nullptr; size_t sort_tmp_bytes_ = 0;
    uint32_t *d_keys_ = nullptr, *d_keys_alt_ = nullptr, *d_idx_ = nullptr, *d_idx_alt_ = nullptr;
    size_t sort_capacity_ = 0;
    Status ensure_sort_scratch(size_t n);
    GaussScratch gs_;
    int gaussian_variant_ = 0;       // 0 auto, 1 scatter (warp per point), 2 gather
    Status ensure_gauss_scratch(size_t n, size_t record_bytes);
    bool use_gather(const Pass& p) const;
    bool use_bin_mma(const Pass& p) const;

    // ---- stats / progress ----
    uint64_t collections_ = 0, points_ = 0;
    std::chrono::steady_clock::time_point t0_;
    pcr_progress_fn progress_fn_ = nullptr;
    void* progress_user_ = nullptr;

    // ---- profiling ----
    int prof_on_ = 0;                       // sampling period, 0 = off
    uint64_t prof_seq_[8] = {};             // kernel groups seen per kind
    struct ProfSpan { cudaEvent_t a, b; ProfKind k; };
    std::vector<ProfSpan> prof_open_;
    std::vector<cudaEvent_t> prof_free_;
    ProfSpan prof_cur_{};
    double prof_ms_[PROF_KINDS] = {};
    uint64_t prof_n_[PROF_KINDS] = {};
    uint64_t prof_h2d_ = 0, prof_d2h_ = 0, prof_points_ = 0, launches_ = 0;
    cudaEvent_t timer_a_ = nullptr, timer_b_ = nullptr;

    // ---- multi-GPU ----
    NcclApi* nccl_ = nullptr;
    void* comm_ = nullptr;
    int rank_ = 0, world_ = 1;
    // peer-memory path: every rank's state / touched / bands / flags mapped into this process
    bool peer_ok_ = false;
    bool partition_ = false;          // tile-partitioned grid + point exchange instead of partial grids + merge
    uint32_t part_waited_epoch_ = 0;  // last epoch whose "pool is empty again" flags this rank has awaited
    std::vector<void*> ipc_opened_;   // peer mappings opened by ipc_exchange
    int comm_layout_ = 0;             // 0 auto, 1 replicated partial grids, 2 tile-partitioned
    int comm_mode_ = 0;               // 0 auto (peer memory when every pair has P2P), 1 NCCL, 2 peer
    bool gather_root_only_ = false;
    bool bands_distributed_ = false;  // N>1: every rank keeps only its own row slice of the bands
    struct PeerBuffers {
        std::vector<uint32_t*> combined;   // one per pass: kMaxParts slots of max_slice cells
        uint32_t* touched_stage = nullptr; // [world][n_tiles]
        float* out = nullptr;
        uint32_t* flags = nullptr;
    };
    uint32_t* d_touched_stage_ = nullptr;
    PeerBuffers peer_[kMaxParts];
    uint32_t* d_flags_ = nullptr;     // [2 phases][kMaxParts]
    uint32_t epoch_ = 0;
    uint32_t waited_epoch_ = 0;        // last epoch whose phase-1 ("everyone is done") flags were awaited
    Status peer_quiesce();            // enqueue that wait if it is still owed
};

// deterministic path (det_kernels.cu)
size_t det_sort_temp_bytes(size_t n, int key_bits);
cudaError_t det_build_keys(cudaStream_t s, const uint8_t* mask, const double* x, const double* y, size_t n,
                           const GridParams& g, uint32_t* keys, uint32_t* idx, uint32_t* touched);
cudaError_t det_sort(cudaStream_t s, void* tmp, size_t tmp_bytes, uint32_t*& keys, uint32_t*& keys_alt,
                     uint32_t*& idx, uint32_t*& idx_alt, size_t n, int key_bits);
cudaError_t det_point_reduce(cudaStream_t s, const uint32_t* keys, const uint32_t* idx, size_t n,
                             const ChannelPtrs& ch, uint32_t* state, const PassLayout& L,
                             uint32_t invalid_key);
Status comm_unique_id(void* id128);
int default_staging_threads(int local_ranks);
// rows [row0,row1) owned by `rank`: ceil(height/world) rows each, the last slices may be short or empty
inline void slice_rows(int height, int world, int rank, int& row0, int& row1)
{
    const long per = (static_cast<long>(height) + world - 1) / world;
    row0 = static_cast<int>(std::min<long>(height, per * rank));
    row1 = static_cast<int>(std::min<long>(height, per * (rank + 1)));
}
void engine_comm_destroy(NcclApi* api, void* comm);

}  // namespace pcrb
