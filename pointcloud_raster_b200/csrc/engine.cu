// engine.cu — host orchestration of the B200 ingest/finalize path.
// See engine.h for the design; reference counterparts are cited per function.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <emmintrin.h>
#include <sched.h>

namespace pcrb {

#define CU_TRY(expr)                                                                     \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess)                                                           \
            return Status::error(PCR_CUDA_ERROR, std::string("CUDA error: ") +           \
                                 cudaGetErrorString(_e) + " (" #expr ")");               \
    } while (0)

#define ST_TRY(expr)                        \
    do {                                    \
        Status _s = (expr);                 \
        if (!_s.ok()) return _s;            \
    } while (0)

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Host copy threads of the pageable staging path: up to 12 (measured optimum on a 16-core host for one
// GPU), but never more than this rank's share of the cores this process may run on — N ranks with 12
// spinning workers each on a 32-core host made the pageable path slower in aggregate at N=8 than at N=1.
int default_staging_threads(int local_ranks)
{
    const unsigned cores = std::max(1u, std::thread::hardware_concurrency());
    unsigned mine = cores;                                   // cores this process may run on
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) mine = static_cast<unsigned>(std::max(1, CPU_COUNT(&set)));
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) local_ranks = std::max(local_ranks, std::atoi(e));
    const unsigned share = std::min(mine, std::max(1u, cores / static_cast<unsigned>(std::max(1, local_ranks))));
    return static_cast<int>(std::max(1u, std::min(12u, share > 1 ? share - 1 : 1u)));   // one core stays with the ingesting thread
}

// ---------------------------------------------------------------------------
// CopyPool
// ---------------------------------------------------------------------------
// Staging copy pageable -> pinned.  The pinned slot is only ever read by the DMA
// engine, so it is written with non-temporal stores: no read-for-ownership of
// the destination lines and no cache pollution (3 instead of 4 DRAM transfers per
// staged byte, counting the DMA read).  Measured on the B200 host (16 cores), API scope,
// 5M-point Point Average: 1.49 -> 1.95 Gpts/s with 8 copy threads, 2.04 with 12.
void CopyPool::stage_copy(void* dst, const void* src, size_t bytes)
{
    if (bytes < 4096) { std::memcpy(dst, src, bytes); return; }
    char* d = static_cast<char*>(dst);
    const char* s = static_cast<const char*>(src);
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    if (head) { std::memcpy(d, s, head); d += head; s += head; bytes -= head; }
    const size_t body = bytes & ~size_t(63);
    for (size_t i = 0; i < body; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
    }
    _mm_sfence();
    if (bytes > body) std::memcpy(d + body, s + body, bytes - body);
}

CopyPool::CopyPool(int threads)
{
    for (int i = 0; i < std::max(0, threads); ++i) workers_.emplace_back(&CopyPool::worker, this);
}

CopyPool::~CopyPool()
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
        ++generation_;
    }
    cv_start_.notify_all();
    for (auto& t : workers_) t.join();
}

void CopyPool::worker()
{
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
        cv_start_.wait(lk, [&] { return stop_ || (job_open_ && generation_ != seen); });
        if (stop_) return;
        seen = generation_;
        ++active_;                       // joined this job; end() waits for us
        lk.unlock();
        drain(true);
        lk.lock();
        if (--active_ == 0) cv_done_.notify_all();
    }
}

// Pull pieces from the cursor.  A piece whose chunk is not writable yet (its ring slot is
// still being read by an earlier DMA) is waited for: briefly spinning, then yielding.
bool CopyPool::drain(bool until_done)
{
    const size_t total = pieces_.size();
    bool copied = false;
    for (;;) {
        size_t i = cursor_.load(std::memory_order_relaxed);
        if (i >= total) return copied;
        if (!until_done && pieces_[i].chunk >= writable_.load(std::memory_order_acquire)) return copied;
        if (!cursor_.compare_exchange_weak(i, i + 1, std::memory_order_relaxed)) continue;
        const Piece& pc = pieces_[i];
        for (int spins = 0; pc.chunk >= writable_.load(std::memory_order_acquire); ++spins) {
            if (spins < 256) _mm_pause();
            else std::this_thread::yield();
        }
        stage_copy(pc.dst, pc.src, pc.bytes);
        staged_[pc.chunk].fetch_add(1, std::memory_order_release);
        copied = true;
        if (!until_done) return copied;
    }
}

void CopyPool::begin(std::vector<Piece>&& pieces, const std::vector<uint32_t>& pieces_in_chunk, bool wake)
{
    pieces_ = std::move(pieces);
    pieces_in_chunk_ = pieces_in_chunk;
    if (staged_capacity_ < pieces_in_chunk_.size()) {
        staged_capacity_ = std::max<size_t>(64, 2 * pieces_in_chunk_.size());
        staged_.reset(new std::atomic<uint32_t>[staged_capacity_]);
    }
    for (size_t k = 0; k < pieces_in_chunk_.size(); ++k) staged_[k].store(0, std::memory_order_relaxed);
    cursor_.store(0, std::memory_order_relaxed);
    writable_.store(0, std::memory_order_relaxed);
    if (workers_.empty() || !wake) return;
    {
        std::lock_guard<std::mutex> lk(mu_);     // publishes the job state above to the workers
        job_open_ = true;
        ++generation_;
    }
    cv_start_.notify_all();
}

void CopyPool::abort()
{
    cursor_.store(pieces_.size(), std::memory_order_relaxed);
    writable_.store(~size_t(0), std::memory_order_release);
    end();
}

// Called once every piece is staged: workers that joined leave at once, late wakers never join.
void CopyPool::end()
{
    if (!workers_.empty()) {
        std::unique_lock<std::mutex> lk(mu_);     // (no-op for a job the workers were never woken for)
        job_open_ = false;
        cv_done_.wait(lk, [&] { return active_ == 0; });
    }
    pieces_.clear();
}

// ---------------------------------------------------------------------------
// Planning
// ---------------------------------------------------------------------------
bool GlyphSpecHost::same_footprint(const GlyphSpecHost& o) const
{
    if (type != o.type) return false;
    if (type == PCR_GLYPH_POINT) return true;
    auto feq = [](float a, float b) { return std::memcmp(&a, &b, sizeof(float)) == 0; };
    if (!feq(max_radius_cells, o.max_radius_cells)) return false;
    if (type == PCR_GLYPH_LINE)
        return direction_channel == o.direction_channel &&
               half_length_channel == o.half_length_channel &&
               feq(default_direction, o.default_direction) &&
               feq(default_half_length, o.default_half_length);
    return sigma_x_channel == o.sigma_x_channel && sigma_y_channel == o.sigma_y_channel &&
           rotation_channel == o.rotation_channel && feq(default_sigma_x, o.default_sigma_x) &&
           feq(default_sigma_y, o.default_sigma_y) && feq(default_rotation, o.default_rotation);
}

static bool power_of_two(double v)
{
    int e;
    return std::isfinite(v) && v != 0.0 && std::frexp(std::fabs(v), &e) == 0.5;
}

static int record_width(int words) { return words <= 1 ? 1 : words <= 2 ? 2 : words <= 4 ? 4 : 8; }

namespace {
// Tries to place reduction `r` (index `band`) into pass `p`; false if it does not fit.
bool place(Pass& p, const ReductionHost& r, int band)
{
    PassLayout L = p.layout;
    std::vector<std::string> chans = p.channels;
    FinalizeProgram fin = p.fin;
    if (fin.n >= kMaxBandsPerPass) return false;

    auto chan_of = [&](const std::string& name) -> int {
        for (size_t i = 0; i < chans.size(); ++i)
            if (chans[i] == name) return static_cast<int>(i);
        if (static_cast<int>(chans.size()) >= kMaxChan) return -2;
        chans.push_back(name);
        return static_cast<int>(chans.size()) - 1;
    };
    auto add_word = [&](int src) -> int {   // src = channel slot, or -1 for the count/weight word
        for (int j = 0; j < L.n_add; ++j)
            if (L.add_src[j] == src) return j;
        if (L.n_add >= kMaxAdd) return -2;
        L.add_src[L.n_add] = static_cast<int8_t>(src);
        return L.n_add++;
    };
    auto ext_word = [&](int8_t* srcs, int& n, int src) -> int {
        for (int j = 0; j < n; ++j)
            if (srcs[j] == src) return j;
        if (n >= kMaxExt) return -2;
        srcs[n] = static_cast<int8_t>(src);
        return n++;
    };

    // Words are recorded by (kind, index-within-kind); absolute word offsets are
    // fixed up when the pass is sealed, because n_add / n_max may still grow.
    int kind = 0, a = -1, b = -1;
    const bool needs_value = r.type != PCR_COUNT;
    int c = -1;
    if (needs_value) {
        c = chan_of(r.value_channel);
        if (c == -2) return false;
    }
    switch (r.type) {
    case PCR_SUM:     kind = FIN_SUM;   a = add_word(c); break;
    case PCR_COUNT:   kind = FIN_COUNT; a = add_word(-1); break;
    case PCR_AVERAGE:
    case PCR_WEIGHTED_AVERAGE:
                      kind = FIN_RATIO; a = add_word(c); b = add_word(-1); break;
    case PCR_MAX:     kind = FIN_MAX;   a = ext_word(L.max_src, L.n_max, c); break;
    case PCR_MIN:     kind = FIN_MIN;   a = ext_word(L.min_src, L.n_min, c); break;
    default: return false;
    }
    if (a == -2 || b == -2) return false;
    if (L.n_add + L.n_max + L.n_min > 8) return false;

    fin.kind[fin.n] = kind;
    fin.word_a[fin.n] = a;    // index within its kind for now
    fin.word_b[fin.n] = b;
    fin.band[fin.n] = band;
    ++fin.n;
    L.n_chan = static_cast<int>(chans.size());
    p.layout = L;
    p.channels = chans;
    p.fin = fin;
    return true;
}

void seal(Pass& p)
{
    PassLayout& L = p.layout;
    L.width = record_width(L.n_add + L.n_max + L.n_min);
    for (int i = 0; i < p.fin.n; ++i) {
        if (p.fin.kind[i] == FIN_MAX) p.fin.word_a[i] += L.n_add;
        if (p.fin.kind[i] == FIN_MIN) p.fin.word_a[i] += L.n_add + L.n_max;
    }
}
}  // namespace

Status Engine::plan()
{
    passes_.clear();
    std::vector<bool> placed(reductions_.size(), false);
    for (size_t i = 0; i < reductions_.size(); ++i) {
        ReductionHost& r = reductions_[i];
        if (r.glyph.type != PCR_GLYPH_POINT && (r.type == PCR_MAX || r.type == PCR_MIN)) {
            r.rejected = true;
            placed[i] = true;
        }
    }
    for (size_t i = 0; i < reductions_.size(); ++i) {
        if (placed[i]) continue;
        Pass p;
        p.glyph = reductions_[i].glyph;
        std::memset(&p.layout, 0, sizeof(p.layout));
        std::memset(&p.fin, 0, sizeof(p.fin));
        for (size_t j = i; j < reductions_.size(); ++j) {
            if (placed[j] || !reductions_[j].glyph.same_footprint(p.glyph)) continue;
            if (place(p, reductions_[j], static_cast<int>(j))) placed[j] = true;
        }
        seal(p);
        passes_.push_back(std::move(p));
    }
    // every channel any pass reads, value channels first
    all_channels_.clear();
    auto want = [&](const std::string& name) {
        if (name.empty()) return;
        if (std::find(all_channels_.begin(), all_channels_.end(), name) == all_channels_.end())
            all_channels_.push_back(name);
    };
    for (const Pass& p : passes_) for (const auto& c : p.channels) want(c);
    for (const Pass& p : passes_) {
        if (p.glyph.type == PCR_GLYPH_LINE) { want(p.glyph.direction_channel); want(p.glyph.half_length_channel); }
        if (p.glyph.type == PCR_GLYPH_GAUSSIAN) { want(p.glyph.sigma_x_channel); want(p.glyph.sigma_y_channel); want(p.glyph.rotation_channel); }
    }
    for (const FilterPredHost& f : filter_) want(f.channel);
    return Status::success();
}

int Engine::channel_slot(const std::string& name)
{
    for (size_t i = 0; i < all_channels_.size(); ++i)
        if (all_channels_[i] == name) return static_cast<int>(i);
    return -1;
}

// ---------------------------------------------------------------------------
// create / destroy   (Pipeline::create + Impl::initialize, pipeline.cpp:92-281,1294-1304)
// ---------------------------------------------------------------------------
Status Engine::create(const pcr_pipeline_desc& desc, Engine** out)
{
    *out = nullptr;
    Engine* e = new Engine();
    Status s = e->init(desc);
    if (!s.ok()) { delete e; return s; }
    *out = e;
    return s;
}

Status Engine::init(const pcr_pipeline_desc& d)
{
    t0_ = std::chrono::steady_clock::now();
    grid_ = d.grid;
    exec_mode_ = d.exec_mode;
    device_ = d.cuda_device_id;
    deterministic_ = d.deterministic == 1;
    exact_ = d.deterministic == 2;
    if (d.deterministic < 0 || d.deterministic > 2)
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: deterministic must be 0, 1 or 2");
    async_device_ingest_ = d.async_ingest != 0;
    point_variant_ = d.point_kernel == 2 ? POINT_TMA : d.point_kernel == 1 ? POINT_DIRECT : POINT_DIRECT;
    point_kernel_knob_ = d.point_kernel;
    bin_cells_log2_ = d.bin_cells_log2;
    bin_pool_points_ = d.bin_pool_points;
    comm_layout_ = d.comm_layout;
    warp_aggregate_ = d.warp_aggregate != 2;
    gaussian_variant_ = d.gaussian_kernel;
    comm_mode_ = d.comm_mode;
    gather_root_only_ = d.comm_root_only == 1;
    bands_distributed_ = d.comm_root_only == 2;
    band_copy_ = d.comm_band_copy;
    // Ring chunk sizes (points), measured on B200 / PCIe Gen5 x16 with 5M-point ingests: staged (pageable)
    // chunks want to be small so that staging, DMA and kernels overlap early (256 Ki: 2.11 Gpts/s, 2 Mi: 1.85);
    // direct DMA out of pinned caller memory wants few large copies (256 Ki: 2.16, 2 Mi: 2.37).
    slot_points_ = d.ring_slot_points ? static_cast<size_t>(d.ring_slot_points) : (size_t(1) << 18);
    slot_points_ = align_up(slot_points_, 1024);
    direct_points_ = d.ring_slot_points ? slot_points_ : (size_t(1) << 21);
    const int depth = d.ring_depth > 0 ? d.ring_depth : 3;
    host_ring_.resize(depth);
    dev_ring_.resize(std::max(depth, 2));
    staging_auto_ = d.staging_threads <= 0;
    staging_threads_ = d.staging_threads > 0 ? d.staging_threads : default_staging_threads(1);

    if (exec_mode_ == PCR_EXEC_CPU)
        return Status::error(PCR_NOT_IMPLEMENTED,
                             "pipeline: ExecutionMode.CPU is not available on the B200 path "
                             "(GPU-only; there is no CPU fallback)");
    if (exec_mode_ != PCR_EXEC_GPU && exec_mode_ != PCR_EXEC_AUTO && exec_mode_ != PCR_EXEC_HYBRID)
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: unknown execution mode");

    // reductions (registry check: src/ops/reduction_registry.cpp:174-186 -> "unknown reduction type")
    for (int i = 0; i < d.num_reductions; ++i) {
        const pcr_reduction_desc& r = d.reductions[i];
        ReductionHost h;
        h.value_channel = r.value_channel ? r.value_channel : "";
        h.type = r.type;
        if (h.type < PCR_SUM || h.type > PCR_COUNT)
            return Status::error(PCR_INVALID_ARGUMENT, "pipeline: unknown reduction type");
        h.band_name = (r.output_band_name && r.output_band_name[0])
                          ? r.output_band_name
                          : h.value_channel + "_" + std::to_string(h.type);   // pipeline.cpp:1178-1180
        auto str = [](const char* s) { return std::string(s ? s : ""); };
        h.glyph.type = r.glyph.type;
        if (h.glyph.type < PCR_GLYPH_POINT || h.glyph.type > PCR_GLYPH_GAUSSIAN)
            return Status::error(PCR_NOT_IMPLEMENTED, "glyph: unknown glyph type");
        h.glyph.direction_channel = str(r.glyph.direction_channel);
        h.glyph.half_length_channel = str(r.glyph.half_length_channel);
        h.glyph.sigma_x_channel = str(r.glyph.sigma_x_channel);
        h.glyph.sigma_y_channel = str(r.glyph.sigma_y_channel);
        h.glyph.rotation_channel = str(r.glyph.rotation_channel);
        h.glyph.default_direction = r.glyph.default_direction;
        h.glyph.default_half_length = r.glyph.default_half_length;
        h.glyph.default_sigma_x = r.glyph.default_sigma_x;
        h.glyph.default_sigma_y = r.glyph.default_sigma_y;
        h.glyph.default_rotation = r.glyph.default_rotation;
        h.glyph.max_radius_cells = r.glyph.max_radius_cells;
        reductions_.push_back(std::move(h));
    }

    if (d.num_predicates < 0 || d.num_predicates > kMaxPredicates || (d.num_predicates > 0 && !d.filter))
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: a filter takes at most 8 predicates");
    for (int i = 0; i < d.num_predicates; ++i) {
        FilterPredHost f;
        f.channel = d.filter[i].channel_name ? d.filter[i].channel_name : "";
        f.op = d.filter[i].op;
        f.value = d.filter[i].value;
        if (f.op < PCR_CMP_EQUAL || f.op > PCR_CMP_NOT_IN_SET)
            return Status::error(PCR_INVALID_ARGUMENT, "pipeline: unknown filter comparison");
        if (d.filter[i].value_set && d.filter[i].value_set_size > 0)
            f.set.assign(d.filter[i].value_set, d.filter[i].value_set + d.filter[i].value_set_size);
        filter_.push_back(std::move(f));
    }

    if (grid_.width <= 0 || grid_.height <= 0)
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: grid dimensions must be positive");
    if (grid_.tile_width <= 0 || grid_.tile_height <= 0)
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: tile dimensions must be positive");
    cells_ = static_cast<size_t>(grid_.width) * static_cast<size_t>(grid_.height);
    // the reference stores the global cell index as u32 (tile_router.cpp:111-112)
    if (cells_ >= (size_t(1) << 32))
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: grid has 2^32 or more cells");

    gp_.min_x = grid_.min_x; gp_.max_x = grid_.max_x; gp_.min_y = grid_.min_y; gp_.max_y = grid_.max_y;
    gp_.csx = grid_.cell_size_x; gp_.csy = grid_.cell_size_y;
    gp_.inv_csx = 1.0 / grid_.cell_size_x; gp_.inv_csy = 1.0 / grid_.cell_size_y;
    gp_.width = grid_.width; gp_.height = grid_.height;
    gp_.tile_w = grid_.tile_width; gp_.tile_h = grid_.tile_height;
    gp_.tiles_x = (grid_.width + grid_.tile_width - 1) / grid_.tile_width;
    gp_.tiles_y = (grid_.height + grid_.tile_height - 1) / grid_.tile_height;
    auto log2_exact = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
    gp_.tile_w_shift = log2_exact(grid_.tile_width);
    gp_.tile_h_shift = log2_exact(grid_.tile_height);
    gp_.exact_x = power_of_two(grid_.cell_size_x);
    gp_.exact_y = power_of_two(grid_.cell_size_y);
    n_tiles_ = gp_.tiles_x * gp_.tiles_y;

    // device (no fallback: gpu_fallback_to_cpu is never honoured on this path)
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count <= 0)
        return Status::error(PCR_CUDA_ERROR,
                             "No CUDA-capable GPU detected - the B200 path has no CPU fallback"
                             " (gpu_fallback_to_cpu is not honoured)");
    if (device_ < 0 || device_ >= count)
        return Status::error(PCR_CUDA_ERROR, "pipeline: cuda_device_id out of range");
    CU_TRY(cudaSetDevice(device_));
    cudaDeviceProp prop{};
    CU_TRY(cudaGetDeviceProperties(&prop, device_));
    sm_count_ = prop.multiProcessorCount;
    if (prop.major < 10)
        return Status::error(PCR_CUDA_ERROR, std::string("pipeline: device '") + prop.name +
                             "' is not sm_100-class; this library carries sm_100a code only");
    CU_TRY(cudaStreamCreateWithFlags(&compute_, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&copy_, cudaStreamNonBlocking));
    {   // highest priority: once the peers' data is there, the short merge should not queue behind the
        // CTAs of a concurrently running ingest kernel
        int lo = 0, hi = 0;
        CU_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU_TRY(cudaStreamCreateWithPriority(&fin_, cudaStreamNonBlocking, hi));
    }
    CU_TRY(cudaEventCreateWithFlags(&e_pushed_, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&e_fin_, cudaEventDisableTiming));

    ST_TRY(plan());
    if (deterministic_)
        for (const Pass& p : passes_)
            if (p.glyph.type == PCR_GLYPH_LINE || (p.glyph.type == PCR_GLYPH_GAUSSIAN && !use_gather(p)))
                return Status::error(PCR_NOT_IMPLEMENTED,
                                     "pipeline: deterministic mode covers the Point glyph and the Gaussian "
                                     "gather kernel, not the Line glyph / Gaussian scatter kernel");
    ST_TRY(alloc_state());
    ST_TRY(init_state());
    CU_TRY(cudaStreamSynchronize(compute_));
    return Status::success();
}

Status Engine::alloc_state()
{
    for (Pass& p : passes_) {
        if (exact_) { ST_TRY(alloc_exact(p)); continue; }          // mode 2 keeps no float records at all
        CU_TRY(cudaMalloc(&p.d_delta[0], cells_ * p.layout.width * sizeof(uint32_t)));
        p.d_state = p.d_delta[0];
    }
    CU_TRY(cudaMalloc(&d_touched_buf_[0], std::max(1, n_tiles_) * sizeof(uint32_t)));
    d_touched_ = d_touched_buf_[0];
    const size_t out_bytes = std::max<size_t>(1, reductions_.size()) * cells_ * sizeof(float);
    CU_TRY(cudaMalloc(&d_out_, out_bytes));
    if (!filter_.empty()) {
        std::vector<float> all;
        for (const FilterPredHost& f : filter_) { filter_set_off_.push_back(all.size()); all.insert(all.end(), f.set.begin(), f.set.end()); }
        CU_TRY(cudaMalloc(&d_filter_sets_, std::max<size_t>(all.size(), 1) * sizeof(float)));
        if (!all.empty()) CU_TRY(cudaMemcpy(d_filter_sets_, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU_TRY(cudaMalloc(&d_survivors_, sizeof(unsigned long long)));
        CU_TRY(cudaMemset(d_survivors_, 0, sizeof(unsigned long long)));
    }
    for (Pass& p : passes_) ST_TRY(bin_setup(p));      // last: the entry pools take their share of what is left
    return Status::success();
}

Status Engine::init_state()
{
    prof_begin(PROF_INIT, compute_);
    for (Pass& p : passes_) {
        if (exact_) { CU_TRY(launch_exact_init(compute_, p.xa, p.layout)); ++launches_; continue; }
        const size_t n_rec = partition_ ? p.bin.cell1 - p.bin.cell0 : cells_;   // partitioned: my bins' cells only
        for (uint32_t* d : p.d_delta)
            if (d) { CU_TRY(launch_init_state(compute_, d, n_rec, p.layout)); ++launches_; }
        if (p.d_owned) {
            int r0, r1;
            slice_rows(grid_.height, world_, rank_, r0, r1);
            CU_TRY(launch_init_state(compute_, p.d_owned, static_cast<size_t>(r1 - r0) * grid_.width, p.layout));
            ++launches_;
        }
    }
    for (uint32_t* t : d_touched_buf_)
        if (t) CU_TRY(cudaMemsetAsync(t, 0, std::max(1, n_tiles_) * sizeof(uint32_t), compute_));
    if (d_touched_merged_) CU_TRY(cudaMemsetAsync(d_touched_merged_, 0, std::max(1, n_tiles_) * sizeof(uint32_t), compute_));
    prof_end(compute_);
    return Status::success();
}

Status Engine::reset()
{
    CU_TRY(cudaSetDevice(device_));
    ST_TRY(synchronize());
    // N>1: collective — the peers' push / scatter kernels write into this rank's buffers
    if (world_ > 1) ST_TRY(comm_barrier());
    ST_TRY(init_state());
    for (Pass& p : passes_)
        if (p.bin.on) {
            CU_TRY(launch_bin_reset(compute_, p.bin.pool, p.bin.open_page, static_cast<size_t>(p.bin.grid) * p.bin.nbins, sm_count_));
            if (p.bin.part_counters)
                CU_TRY(cudaMemsetAsync(p.bin.part_counters, 0, static_cast<size_t>(world_) * 4 * sizeof(uint32_t), compute_));
            p.bin.pending = 0;
        }
    if (d_survivors_) CU_TRY(cudaMemsetAsync(d_survivors_, 0, sizeof(unsigned long long), compute_));
    collections_ = 0;
    points_ = 0;
    finalized_ = false;
    ST_TRY(synchronize());
    if (world_ > 1) ST_TRY(comm_barrier());
    return Status::success();
}

// Cells whose finalized bands this rank produces: everything on one GPU, its row slice (replicated partial
// grids) or the cells of its bins (tile-partitioned layout) on several.
void Engine::owned_cells(uint64_t& c0, uint64_t& c1) const
{
    c0 = 0; c1 = cells_;
    if (world_ <= 1) return;
    if (partition_ && !passes_.empty()) { c0 = passes_[0].bin.cell0; c1 = passes_[0].bin.cell1; return; }
    int r0, r1;
    slice_rows(grid_.height, world_, rank_, r0, r1);
    c0 = static_cast<uint64_t>(r0) * grid_.width;
    c1 = static_cast<uint64_t>(r1) * grid_.width;
}

Status Engine::synchronize()
{
    CU_TRY(cudaSetDevice(device_));
    ST_TRY(peer_quiesce());
    CU_TRY(cudaStreamSynchronize(copy_));
    CU_TRY(cudaStreamSynchronize(compute_));
    if (push_) CU_TRY(cudaStreamSynchronize(push_));
    CU_TRY(cudaStreamSynchronize(fin_));
    fin_pending_ = false;
    return Status::success();
}

Status Engine::join_fin()
{
    if (fin_pending_) CU_TRY(cudaStreamWaitEvent(compute_, e_fin_, 0));
    fin_pending_ = false;
    return Status::success();
}

Engine::~Engine()
{
    if (compute_ || copy_) {
        cudaSetDevice(device_);
        if (copy_) cudaStreamSynchronize(copy_);
        if (compute_) cudaStreamSynchronize(compute_);
        if (push_) cudaStreamSynchronize(push_);
        if (fin_) cudaStreamSynchronize(fin_);
    }
    peer_unmap();
    if (comm_) engine_comm_destroy(nccl_, comm_);
    for (Pass& p : passes_) {
        cudaFree(p.d_delta[0]); cudaFree(p.d_delta[1]); cudaFree(p.d_owned); cudaFree(p.d_combined); bin_free(p);
        cudaFree(p.xa.limbs); cudaFree(p.xa.flags); cudaFree(p.xa.ext);
        cudaFree(p.xa_sum.limbs); cudaFree(p.xa_sum.flags); cudaFree(p.xa_sum.ext);
    }
    if (h_overflow_) cudaFreeHost(h_overflow_);
    cudaFree(d_touched_buf_[0]);
    cudaFree(d_touched_buf_[1]);
    if (e_delta_) cudaEventDestroy(e_delta_);
    cudaFree(d_touched_all_);
    cudaFree(d_touched_merged_);
    cudaFree(d_flags_);
    cudaFree(d_filter_sets_); cudaFree(d_mask_); cudaFree(d_survivors_);
    cudaFree(d_touched_stage_);
    cudaFree(d_out_);
    if (h_out_) cudaFreeHost(h_out_);
    for (HostSlot& s : host_ring_) {
        if (s.h) cudaFreeHost(s.h);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
    }
    for (DevSlot& s : dev_ring_) {
        cudaFree(s.d);
        if (s.filled) cudaEventDestroy(s.filled);
        if (s.kernel_done) cudaEventDestroy(s.kernel_done);
    }
    cudaFree(gs_.keys); cudaFree(gs_.keys_alt); cudaFree(gs_.idx); cudaFree(gs_.idx_alt);
    cudaFree(gs_.sort_tmp); cudaFree(gs_.records); cudaFree(gs_.aux);
    cudaFree(d_sort_tmp_); cudaFree(d_keys_); cudaFree(d_keys_alt_); cudaFree(d_idx_); cudaFree(d_idx_alt_);
    for (auto& sp : prof_open_) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto& ev : prof_free_) cudaEventDestroy(ev);
    if (timer_a_) { cudaEventDestroy(timer_a_); cudaEventDestroy(timer_b_); }
    delete pool_;
    if (e_pushed_) cudaEventDestroy(e_pushed_);
    if (e_fin_) cudaEventDestroy(e_fin_);
    if (compute_) cudaStreamDestroy(compute_);
    if (copy_) cudaStreamDestroy(copy_);
    if (fin_) cudaStreamDestroy(fin_);
    if (push_) cudaStreamDestroy(push_);
}

// Pipeline::validate, src/engine/pipeline.cpp:1306-1338
Status Engine::validate() const
{
    if (grid_.width <= 0 || grid_.height <= 0)
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: grid dimensions must be positive");
    if (grid_.tile_width <= 0 || grid_.tile_height <= 0)
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: tile dimensions must be positive");
    if (reductions_.empty())
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: at least one reduction must be specified");
    for (const auto& r : reductions_) {
        if (r.value_channel.empty())
            return Status::error(PCR_INVALID_ARGUMENT, "pipeline: value_channel must be specified");
        if (r.type < PCR_SUM || r.type > PCR_COUNT)
            return Status::error(PCR_INVALID_ARGUMENT, "pipeline: unknown reduction type");
    }
    return Status::success();
}

// ---------------------------------------------------------------------------
// profiling
// ---------------------------------------------------------------------------
void Engine::prof_begin(ProfKind k, cudaStream_t s)
{
    prof_cur_ = {nullptr, nullptr, k};
    if (!prof_on_ || (prof_seq_[k]++ % static_cast<uint64_t>(prof_on_)) != 0) return;   // not sampled
    auto get = [&]() {
        cudaEvent_t e;
        if (!prof_free_.empty()) { e = prof_free_.back(); prof_free_.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    };
    prof_cur_ = {get(), get(), k};
    cudaEventRecord(prof_cur_.a, s);
}

void Engine::prof_end(cudaStream_t s)
{
    if (!prof_cur_.a) return;
    cudaEventRecord(prof_cur_.b, s);
    prof_open_.push_back(prof_cur_);
    ++prof_n_[prof_cur_.k];
    prof_cur_.a = nullptr;
}

Status Engine::prof_collect()
{
    for (auto& sp : prof_open_) {
        CU_TRY(cudaEventSynchronize(sp.b));
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, sp.a, sp.b));
        prof_ms_[sp.k] += ms;
        prof_free_.push_back(sp.a);
        prof_free_.push_back(sp.b);
    }
    prof_open_.clear();
    return Status::success();
}

Status Engine::profile_enable(int period)
{
    CU_TRY(cudaSetDevice(device_));
    prof_on_ = period;
    for (auto& q : prof_seq_) q = 0;
    return Status::success();
}

Status Engine::profile_reset()
{
    CU_TRY(cudaSetDevice(device_));
    ST_TRY(prof_collect());
    for (int k = 0; k < PROF_KINDS; ++k) { prof_ms_[k] = 0; prof_n_[k] = 0; }
    prof_h2d_ = prof_d2h_ = prof_points_ = launches_ = 0;
    return Status::success();
}

Status Engine::profile_read(pcr_profile& o)
{
    CU_TRY(cudaSetDevice(device_));
    ST_TRY(prof_collect());
    o.accumulate_ms = prof_ms_[PROF_ACC];  o.accumulate_launches = prof_n_[PROF_ACC];
    o.sort_ms = prof_ms_[PROF_SORT];       o.sort_launches = prof_n_[PROF_SORT];
    o.finalize_ms = prof_ms_[PROF_FIN];    o.finalize_launches = prof_n_[PROF_FIN];
    o.init_ms = prof_ms_[PROF_INIT];       o.init_launches = prof_n_[PROF_INIT];
    o.h2d_bytes = prof_h2d_; o.d2h_bytes = prof_d2h_; o.points = prof_points_;
    o.kernel_launches = launches_;
    o.push_ms = prof_ms_[PROF_PUSH];       o.push_launches = prof_n_[PROF_PUSH];
    return Status::success();
}

Status Engine::timer_begin()
{
    CU_TRY(cudaSetDevice(device_));
    if (!timer_a_) { CU_TRY(cudaEventCreate(&timer_a_)); CU_TRY(cudaEventCreate(&timer_b_)); }
    CU_TRY(cudaEventRecord(timer_a_, compute_));
    return Status::success();
}

Status Engine::timer_end(double& ms)
{
    CU_TRY(cudaSetDevice(device_));
    if (!timer_a_) return Status::error(PCR_INVALID_ARGUMENT, "pipeline: timer_end without timer_begin");
    CU_TRY(cudaStreamSynchronize(copy_));
    ST_TRY(join_fin());
    CU_TRY(cudaEventRecord(timer_b_, compute_));
    CU_TRY(cudaEventSynchronize(timer_b_));
    float f = 0.f;
    CU_TRY(cudaEventElapsedTime(&f, timer_a_, timer_b_));
    ms = f;
    return Status::success();
}

// ---------------------------------------------------------------------------
// ingest   (Pipeline::ingest -> Impl::process_cloud, pipeline.cpp:283-770)
// ---------------------------------------------------------------------------
Status Engine::ingest(const double* x, const double* y, size_t n, const pcr_channel_view* chans,
                      int nchans, int location)
{
    if (n == 0) return Status::success();                       // pipeline.cpp:284-287
    if (!x || !y) return Status::error(PCR_INVALID_ARGUMENT, "pipeline: null coordinate arrays");
    CU_TRY(cudaSetDevice(device_));

    auto find = [&](const std::string& name) -> const pcr_channel_view* {
        for (int i = 0; i < nchans; ++i)
            if (chans[i].name && name == chans[i].name) return &chans[i];
        return nullptr;
    };
    // filter channels (filter_points_cpu, src/engine/filter.cpp:99-118)
    for (const FilterPredHost& f : filter_) {
        const pcr_channel_view* v = find(f.channel);
        if (!v || !v->data)
            return Status::error(PCR_INVALID_ARGUMENT, "filter_points: channel not found: " + f.channel);
        if (v->dtype != PCR_F32)
            return Status::error(PCR_INVALID_ARGUMENT, "filter_points: only Float32 channels supported for filtering");
    }
    // value channels must exist and be Float32 (pipeline.cpp:365-378); checked for
    // every reduction BEFORE any work, so a failing ingest leaves the state untouched
    // (the reference would already have folded the reductions listed earlier).
    for (const auto& r : reductions_) {
        const pcr_channel_view* v = find(r.value_channel);
        if (!v || !v->data)
            return Status::error(PCR_INVALID_ARGUMENT, "pipeline: value channel not found: " + r.value_channel);
        if (v->dtype != PCR_F32)
            return Status::error(PCR_INVALID_ARGUMENT, "pipeline: value channel must be Float32");
    }
    for (const auto& r : reductions_)
        if (r.rejected)
            return Status::error(PCR_NOT_IMPLEMENTED,
                                 "pipeline: glyph splatting only supports WeightedAverage, Average, "
                                 "Sum, or Count reduction types");

    // pointer per planned channel; glyph channels that are absent or not Float32
    // silently fall back to the default (pipeline.cpp:551-560)
    std::vector<const float*> ptrs(all_channels_.size(), nullptr);
    for (size_t i = 0; i < all_channels_.size(); ++i) {
        const pcr_channel_view* v = find(all_channels_[i]);
        if (v && v->data && v->dtype == PCR_F32) ptrs[i] = static_cast<const float*>(v->data);
    }

    if (location == PCR_MEM_DEVICE) {
        // a cloud on another GPU would fault inside the kernels: refuse it here
        auto on_my_device = [&](const void* p) {
            cudaPointerAttributes a{};
            return cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice &&
                   a.device == device_;
        };
        bool ok = on_my_device(x) && on_my_device(y);
        for (const float* p : ptrs) ok = ok && (!p || on_my_device(p));
        if (!ok) {
            cudaGetLastError();
            return Status::error(PCR_INVALID_ARGUMENT,
                                 "pipeline: Device cloud does not live on cuda_device_id " + std::to_string(device_));
        }
    }
    Status s;
    if (location == PCR_MEM_DEVICE) s = ingest_device(x, y, n, ptrs);
    else if (location == PCR_MEM_HOST || location == PCR_MEM_HOST_PINNED)
        s = ingest_host(x, y, n, ptrs, location == PCR_MEM_HOST_PINNED);
    else return Status::error(PCR_INVALID_ARGUMENT, "pipeline: unknown memory location");
    ST_TRY(s);

    if (filter_.empty()) points_ += n;   // every ingested point, in-grid or not (pipeline.cpp:749); with a filter
                                         // the survivors are counted on the device (points_processed += filtered_count)
    ++collections_;
    finalized_ = false;

    if (progress_fn_) {
        pcr_progress info{};
        ST_TRY(stats(info));
        if (!progress_fn_(&info, progress_user_))
            return Status::error(PCR_INVALID_ARGUMENT, "pipeline: cancelled by user");   // pipeline.cpp:762-766
    }
    return Status::success();
}

bool Engine::use_gather(const Pass& p) const
{
    if (p.glyph.type != PCR_GLYPH_GAUSSIAN || !gauss_gather_supported(p.layout)) return false;
    if (exact_) return false;          // mode 2: the scatter kernel's per-cell contributions are summed exactly
    if (gaussian_variant_ == 1) return false;
    if (gaussian_variant_ == 2 || gaussian_variant_ == 3) return true;
    // auto: the sorted paths (per-bin GEMM, tile gather) need footprints wide enough to amortise their per-point
    // tables.  Measured on B200, 5M points, 1000x1000 (kernel scope): sigma=16 (r=32) per-bin GEMM 5.4 ms, tile
    // gather 11.7 ms, scatter 58 ms; sigma=4 (r=12) 3.2 / 5.0 / 11.1 ms; below ~r=4 the scatter's few REDs per
    // point win.  The radius cap is the only bound known at plan time.
    return deterministic_ || p.glyph.max_radius_cells >= 8.0f;
}

// Behind the same keys / sort / records front end: the per-bin GEMM (k_gauss_binmma) where it applies and
// bit-reproducibility was not asked for, else the tile gather.  gaussian_kernel = 2 forces the tile gather.
bool Engine::use_bin_mma(const Pass& p) const
{
    if (deterministic_ || gaussian_variant_ == 2) return false;
    const bool rotated = !p.glyph.rotation_channel.empty() || p.glyph.default_rotation != 0.0f;
    return gauss_binmma_supported(p.layout, p.glyph.max_radius_cells, rotated);
}

Status Engine::ensure_gauss_scratch(size_t n, size_t record_bytes)
{
    if (n <= gs_.capacity && record_bytes <= gs_.record_bytes) return Status::success();
    CU_TRY(cudaStreamSynchronize(compute_));
    cudaFree(gs_.keys); cudaFree(gs_.keys_alt); cudaFree(gs_.idx); cudaFree(gs_.idx_alt);
    cudaFree(gs_.sort_tmp); cudaFree(gs_.records); cudaFree(gs_.aux);
    gs_ = GaussScratch{};
    const size_t cap = std::max(n, std::min(slot_points_, size_t(1) << 22));
    const size_t rb = std::max<size_t>(record_bytes, 64);
    gs_.sort_tmp_bytes = gauss_sort_temp_bytes(cap, gp_);
    CU_TRY(cudaMalloc(&gs_.sort_tmp, std::max<size_t>(gs_.sort_tmp_bytes, 16)));
    CU_TRY(cudaMalloc(&gs_.keys, cap * 4));
    CU_TRY(cudaMalloc(&gs_.keys_alt, cap * 4));
    CU_TRY(cudaMalloc(&gs_.idx, cap * 4));
    CU_TRY(cudaMalloc(&gs_.idx_alt, cap * 4));
    CU_TRY(cudaMalloc(&gs_.records, cap * rb));
    CU_TRY(cudaMalloc(&gs_.aux, 2 * sizeof(int)));
    gs_.capacity = cap;
    gs_.record_bytes = rb;
    return Status::success();
}

// FilterSpec -> byte mask for this chunk (nullptr when the pipeline has no filter).
Status Engine::build_mask(size_t n, const std::vector<const float*>& cp, const uint8_t** mask)
{
    *mask = nullptr;
    if (filter_.empty()) return Status::success();
    if (n > mask_capacity_) {
        CU_TRY(cudaStreamSynchronize(compute_));
        cudaFree(d_mask_);
        d_mask_ = nullptr;
        mask_capacity_ = 0;
        const size_t cap = std::max(n, slot_points_);
        CU_TRY(cudaMalloc(&d_mask_, cap));
        mask_capacity_ = cap;
    }
    FilterProgram fp{};
    fp.n = static_cast<int>(filter_.size());
    for (int k = 0; k < fp.n; ++k) {
        fp.chan[k] = cp[channel_slot(filter_[k].channel)];
        fp.op[k] = filter_[k].op;
        fp.value[k] = filter_[k].value;
        fp.set[k] = d_filter_sets_ + filter_set_off_[k];
        fp.set_size[k] = static_cast<int>(filter_[k].set.size());
    }
    CU_TRY(launch_filter_mask(compute_, fp, n, d_mask_, d_survivors_));
    ++launches_;
    *mask = d_mask_;
    return Status::success();
}

// Launches every pass over one device-resident chunk on the compute stream.
Status Engine::run_passes(const double* dx, const double* dy, size_t n,
                          const std::vector<const float*>& cp)
{
    const uint8_t* mask = nullptr;
    ST_TRY(build_mask(n, cp, &mask));
    if (exact_) return run_passes_exact(dx, dy, n, cp, mask);
    if (deterministic_) ST_TRY(run_passes_deterministic(dx, dy, n, cp, mask));
    // tile-binned Point passes only append entries now; their reductions run bin by bin at finalize
    bool any_binned = false;
    for (Pass& p : passes_) any_binned = any_binned || p.bin.on;
    if (any_binned) {
        prof_begin(PROF_SORT, compute_);
        for (Pass& p : passes_) {
            if (!p.bin.on) continue;
            ChannelPtrs ch{};
            for (size_t c = 0; c < p.channels.size(); ++c) ch.p[c] = cp[channel_slot(p.channels[c])];
            ST_TRY(bin_append(p, mask, dx, dy, ch, n));
        }
        prof_end(compute_);
    }
    prof_begin(PROF_ACC, compute_);
    for (Pass& p : passes_) {
        if (p.bin.on) continue;
        if (deterministic_ && p.glyph.type == PCR_GLYPH_POINT) continue;   // folded by the sort path above
        ChannelPtrs ch{};
        for (size_t c = 0; c < p.channels.size(); ++c) ch.p[c] = cp[channel_slot(p.channels[c])];
        if (p.glyph.type == PCR_GLYPH_POINT) {
            CU_TRY(launch_point_accumulate(compute_, point_variant_, warp_aggregate_, mask, dx, dy, ch, n,
                                           p.d_state, gp_, p.layout, d_touched_, sm_count_));
            ++launches_;
        } else {
            GlyphParams g{};
            auto opt = [&](const std::string& name) -> const float* {
                const int s = name.empty() ? -1 : channel_slot(name);
                return s < 0 ? nullptr : cp[s];
            };
            g.direction = opt(p.glyph.direction_channel);     g.default_direction = p.glyph.default_direction;
            g.half_length = opt(p.glyph.half_length_channel); g.default_half_length = p.glyph.default_half_length;
            g.sigma_x = opt(p.glyph.sigma_x_channel);         g.default_sigma_x = p.glyph.default_sigma_x;
            g.sigma_y = opt(p.glyph.sigma_y_channel);         g.default_sigma_y = p.glyph.default_sigma_y;
            g.rotation = opt(p.glyph.rotation_channel);       g.default_rotation = p.glyph.default_rotation;
            g.max_radius_cells = p.glyph.max_radius_cells;
            if (p.glyph.type == PCR_GLYPH_LINE)
                CU_TRY(launch_line_accumulate(compute_, mask, dx, dy, ch, g, n, p.d_state, gp_, p.layout, d_touched_));
            else if (use_gather(p)) {
                // bounded sub-chunks keep the sort / record scratch small
                const size_t kSub = size_t(1) << 22;
                for (size_t q0 = 0; q0 < n; q0 += kSub) {
                    const size_t cnt = std::min(kSub, n - q0);
                    ST_TRY(ensure_gauss_scratch(cnt, gauss_record_bytes(p.layout)));
                    ChannelPtrs c2 = ch;
                    for (size_t c = 0; c < p.channels.size(); ++c) c2.p[c] = ch.p[c] + q0;
                    GlyphParams g2 = g;
                    if (g2.sigma_x) g2.sigma_x += q0;
                    if (g2.sigma_y) g2.sigma_y += q0;
                    if (g2.rotation) g2.rotation += q0;
                    CU_TRY(launch_gaussian_gather(compute_, mask ? mask + q0 : nullptr, dx + q0, dy + q0, c2, g2, cnt, p.d_state, gp_, p.layout,
                                                  d_touched_, gs_, sm_count_, use_bin_mma(p)));
                    launches_ += 4;   // keys, sort, records, gather / per-bin GEMM
                }
            } else
                CU_TRY(launch_gaussian_accumulate(compute_, mask, dx, dy, ch, g, n, p.d_state, gp_, p.layout, d_touched_));
            ++launches_;
        }
    }
    prof_end(compute_);
    prof_points_ += n;
    return Status::success();
}

Status Engine::ingest_device(const double* x, const double* y, size_t n,
                             const std::vector<const float*>& cp)
{
    ST_TRY(run_passes(x, y, n, cp));
    if (!async_device_ingest_) CU_TRY(cudaStreamSynchronize(compute_));
    return Status::success();
}

// Segment layout of a chunk buffer holding `cap` points: x | y | channel 0 | channel 1 | ...,
// each segment 256-B aligned.
struct ChunkLayout {
    size_t seg64, seg32, bytes;
    ChunkLayout(size_t cap, size_t nch)
        : seg64(align_up(cap * 8, 256)), seg32(align_up(cap * 4, 256)), bytes(2 * seg64 + nch * seg32) {}
};

Status Engine::ensure_ring()
{
    if (dev_ring_[0].d) return Status::success();
    // The sort-based passes (Gaussian gather, deterministic Point) are markedly more efficient on larger
    // batches than a staged chunk (5M-point Gaussian sigma=4, API scope: 645 Mpts/s per 256 Ki chunk, 788 per
    // 1 Mi), so their kernels run once per group of staged chunks.
    bool sorted_pass = deterministic_;
    for (const Pass& p : passes_) sorted_pass = sorted_pass || use_gather(p);
    group_chunks_ = sorted_pass ? 4 : 1;
    const size_t nch = all_channels_.size();
    slot_bytes_ = ChunkLayout(slot_points_, nch).bytes;
    const size_t dev_bytes = std::max(ChunkLayout(group_chunks_ * slot_points_, nch).bytes,
                                      ChunkLayout(direct_points_, nch).bytes);
    for (HostSlot& s : host_ring_) {
        CU_TRY(cudaMallocHost(&s.h, slot_bytes_));
        CU_TRY(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
    }
    for (DevSlot& s : dev_ring_) {
        CU_TRY(cudaMalloc(&s.d, dev_bytes));
        CU_TRY(cudaEventCreateWithFlags(&s.filled, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&s.kernel_done, cudaEventDisableTiming));
    }
    if (!pool_) pool_ = new CopyPool(staging_threads_ > 1 ? staging_threads_ : 0);   // 1 = the ingesting thread copies
    return Status::success();
}

// The CUDA-stream ingest ring.  Chunk k of the cloud is staged into pinned buffer (pos+k)%depth by
// the copy pool (skipped when the caller's memory is already pinned), shipped with cudaMemcpyAsync
// on the copy stream into the device buffer of its kernel group, and the group is consumed by the
// pass kernels on the compute stream.  The ingesting thread never blocks on a kernel: re-use of a
// device buffer is a stream-side wait on its kernel_done event, re-use of a pinned buffer is gated
// by polling its h2d_done event.  Replaces PointCloud::to_device_async + cudaStreamSynchronize
// (pipeline.cpp:299-327) and Hybrid mode (pipeline.cpp:785-1152).
Status Engine::ingest_host(const double* x, const double* y, size_t n,
                           const std::vector<const float*>& cp, bool pinned)
{
    if (n == 0) return Status::success();
    ST_TRY(ensure_ring());
    const size_t nch = all_channels_.size();
    const size_t chunk = pinned ? direct_points_ : slot_points_;
    const size_t group = pinned ? 1 : group_chunks_;              // chunks per kernel group
    const ChunkLayout hl(slot_points_, nch), dl(group * chunk, nch);
    const size_t hdepth = host_ring_.size(), ddepth = dev_ring_.size();
    const size_t nchunks = (n + chunk - 1) / chunk;
    size_t present = 0;
    for (size_t c = 0; c < nch; ++c) present += cp[c] ? 1 : 0;

    // H2D of chunk k and, when it completes its group, the group's kernels.
    // `staged` = the data sits in the chunk's pinned buffer.
    auto issue = [&](size_t k, bool staged) -> Status {
        const size_t p0 = k * chunk, cnt = std::min(chunk, n - p0);
        const size_t grp = k / group, sub = k % group;
        DevSlot& ds = dev_ring_[(dev_pos_ + grp) % ddepth];
        if (sub == 0 && ds.used) CU_TRY(cudaStreamWaitEvent(copy_, ds.kernel_done, 0));   // buffer drained
        char* dx = ds.d + sub * chunk * 8;
        char* dy = ds.d + dl.seg64 + sub * chunk * 8;
        if (staged) {
            HostSlot& hs = host_ring_[(host_pos_ + k) % hdepth];
            if (group == 1 && present == nch && cnt == chunk) {
                // same layout on both sides, every segment present and full: one contiguous H2D
                CU_TRY(cudaMemcpyAsync(ds.d, hs.h, slot_bytes_, cudaMemcpyHostToDevice, copy_));
            } else {
                CU_TRY(cudaMemcpyAsync(dx, hs.h, cnt * 8, cudaMemcpyHostToDevice, copy_));
                CU_TRY(cudaMemcpyAsync(dy, hs.h + hl.seg64, cnt * 8, cudaMemcpyHostToDevice, copy_));
                for (size_t c = 0; c < nch; ++c)
                    if (cp[c])
                        CU_TRY(cudaMemcpyAsync(ds.d + 2 * dl.seg64 + c * dl.seg32 + sub * chunk * 4,
                                               hs.h + 2 * hl.seg64 + c * hl.seg32, cnt * 4,
                                               cudaMemcpyHostToDevice, copy_));
            }
            CU_TRY(cudaEventRecord(hs.h2d_done, copy_));
            hs.in_flight = true;
        } else {
            CU_TRY(cudaMemcpyAsync(dx, x + p0, cnt * 8, cudaMemcpyHostToDevice, copy_));
            CU_TRY(cudaMemcpyAsync(dy, y + p0, cnt * 8, cudaMemcpyHostToDevice, copy_));
            for (size_t c = 0; c < nch; ++c)
                if (cp[c])
                    CU_TRY(cudaMemcpyAsync(ds.d + 2 * dl.seg64 + c * dl.seg32 + sub * chunk * 4, cp[c] + p0,
                                           cnt * 4, cudaMemcpyHostToDevice, copy_));
        }
        prof_h2d_ += cnt * (16 + 4 * present);
        if (sub + 1 < group && k + 1 < nchunks) return Status::success();     // group still filling

        const size_t g0 = grp * group * chunk, gcnt = std::min(group * chunk, n - g0);
        std::vector<const float*> dptr(nch, nullptr);
        for (size_t c = 0; c < nch; ++c)
            if (cp[c]) dptr[c] = reinterpret_cast<const float*>(ds.d + 2 * dl.seg64 + c * dl.seg32);
        CU_TRY(cudaEventRecord(ds.filled, copy_));
        CU_TRY(cudaStreamWaitEvent(compute_, ds.filled, 0));
        ST_TRY(run_passes(reinterpret_cast<const double*>(ds.d), reinterpret_cast<const double*>(ds.d + dl.seg64),
                          gcnt, dptr));
        CU_TRY(cudaEventRecord(ds.kernel_done, compute_));
        ds.used = true;
        return Status::success();
    };
    const size_t ngroups = (nchunks + group - 1) / group;

    if (pinned) {
        for (size_t k = 0; k < nchunks; ++k) ST_TRY(issue(k, false));
        dev_pos_ = (dev_pos_ + ngroups) % ddepth;
        // the caller may reuse its buffers when we return: the DMA out of them must be complete
        CU_TRY(cudaStreamSynchronize(copy_));
        return Status::success();
    }

    // pageable: pieces of <= 64 Ki points of one segment of one chunk, in chunk order
    constexpr size_t kPiecePoints = size_t(1) << 16;
    std::vector<CopyPool::Piece> pieces;
    std::vector<uint32_t> per_chunk(nchunks, 0);
    pieces.reserve(nchunks * (2 + present) * ((chunk + kPiecePoints - 1) / kPiecePoints));
    for (size_t k = 0; k < nchunks; ++k) {
        const size_t p0 = k * chunk, cnt = std::min(chunk, n - p0);
        HostSlot& s = host_ring_[(host_pos_ + k) % hdepth];
        auto cut = [&](char* dst, const char* src, size_t elem) {
            for (size_t q = 0; q < cnt; q += kPiecePoints) {
                const size_t m = std::min(kPiecePoints, cnt - q);
                pieces.push_back({dst + q * elem, src + q * elem, m * elem, static_cast<uint32_t>(k)});
                ++per_chunk[k];
            }
        };
        cut(s.h, reinterpret_cast<const char*>(x + p0), 8);
        cut(s.h + hl.seg64, reinterpret_cast<const char*>(y + p0), 8);
        for (size_t c = 0; c < nch; ++c)
            if (cp[c]) cut(s.h + 2 * hl.seg64 + c * hl.seg32, reinterpret_cast<const char*>(cp[c] + p0), 4);
    }
    // a small cloud is not worth a wake-up: the ingesting thread stages it itself
    const bool self_serve = pool_->workers() == 0 || n * (16 + 4 * present) < (size_t(1) << 20);
    pool_->begin(std::move(pieces), per_chunk, !self_serve);
    struct Closer {       // on an error path: stop handing out pieces, release spinning workers
        CopyPool* p; bool done = false;
        ~Closer() { if (!done) p->abort(); }
    } closer{pool_};

    size_t issued = 0, writable = 0;
    while (issued < nchunks) {
        // chunk w may be staged once the DMA that last read its pinned buffer is finished
        while (writable < nchunks && writable < issued + hdepth) {
            HostSlot& s = host_ring_[(host_pos_ + writable) % hdepth];
            if (s.in_flight) {
                const cudaError_t q = cudaEventQuery(s.h2d_done);
                if (q == cudaErrorNotReady) break;
                CU_TRY(q);
                s.in_flight = false;
            }
            pool_->set_writable(++writable);
        }
        if (pool_->staged(issued)) {
            ST_TRY(issue(issued, true));
            ++issued;
        } else if (!self_serve || !pool_->help()) {
            _mm_pause();
        }
    }
    closer.done = true;
    pool_->end();
    host_pos_ = (host_pos_ + nchunks) % hdepth;
    dev_pos_ = (dev_pos_ + ngroups) % ddepth;
    return Status::success();
}

// ---------------------------------------------------------------------------
// finalize   (Pipeline::finalize -> Impl::finalize_result, pipeline.cpp:1154-1286)
// ---------------------------------------------------------------------------
Status Engine::finalize(bool to_host)
{
    CU_TRY(cudaSetDevice(device_));
    ST_TRY(bin_flush_all());
    if (world_ > 1) ST_TRY(finalize_multi());
    else ST_TRY(finalize_single());
    if (to_host || !async_device_ingest_) {
        ST_TRY(peer_quiesce());   // peers' band stores must have landed
        ST_TRY(join_fin());
    }
    if (to_host) {
        const size_t bytes = reductions_.size() * cells_ * sizeof(float);
        if (!h_out_) CU_TRY(cudaMallocHost(&h_out_, std::max<size_t>(bytes, 4)));
        CU_TRY(cudaMemcpyAsync(h_out_, d_out_, bytes, cudaMemcpyDeviceToHost, compute_));
        prof_d2h_ += bytes;
    }
    const bool syncing = to_host || !async_device_ingest_;
    int n_ovf = 0;
    if (syncing && h_overflow_)
        for (Pass& p : passes_)
            if (p.bin.on && n_ovf < 16)
                CU_TRY(cudaMemcpyAsync(h_overflow_ + n_ovf++, p.bin.pool.overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, compute_));
    if (syncing) CU_TRY(cudaStreamSynchronize(compute_));
    for (int i = 0; i < n_ovf; ++i)
        if (h_overflow_[i])
            return Status::error(PCR_OUT_OF_MEMORY, "pipeline: the tile-binning entry pool overflowed (points were dropped)");
    finalized_ = true;
    return Status::success();
}

Status Engine::finalize_single()
{
    prof_begin(PROF_FIN, compute_);
    for (size_t i = 0; i < reductions_.size(); ++i)
        if (reductions_[i].rejected)   // never accumulated: all NaN (0xFFFFFFFF is a NaN)
            CU_TRY(cudaMemsetAsync(d_out_ + i * cells_, 0xFF, cells_ * sizeof(float), compute_));
    for (Pass& p : passes_) {
        if (exact_) {
            CU_TRY(launch_finalize_exact(compute_, p.xa, 0, cells_, d_out_, cells_, gp_, p.layout, p.fin, d_touched_));
            ++launches_;
            continue;
        }
        StateParts parts{};
        parts.part[0] = p.d_state;
        parts.n = 1;
        OutTargets outs{};
        outs.out[0] = d_out_;
        outs.n = 1;
        CU_TRY(launch_finalize(compute_, parts, 0, 0, cells_, outs, cells_, gp_, p.layout, p.fin, d_touched_));
        ++launches_;
    }
    prof_end(compute_);
    return Status::success();
}

Status Engine::result_band(int band, const float** data, int* rows, int* cols, bool device)
{
    if (!finalized_)
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: result requested before finalize()");
    if (band < 0 || band >= static_cast<int>(reductions_.size()))
        return Status::error(PCR_INVALID_ARGUMENT, "Invalid band index or data type");
    if (!device && !h_out_)
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: bands were finalized on the device only");
    *data = (device ? d_out_ : h_out_) + static_cast<size_t>(band) * cells_;
    *rows = grid_.height;
    *cols = grid_.width;
    return Status::success();
}

Status Engine::band_name(int band, std::string& out) const
{
    if (band < 0 || band >= static_cast<int>(reductions_.size()))
        return Status::error(PCR_INVALID_ARGUMENT, "Invalid band index or data type");
    out = reductions_[band].band_name;
    return Status::success();
}

// Pipeline::stats, pipeline.cpp:1388-1401
Status Engine::stats(pcr_progress& o)
{
    CU_TRY(cudaSetDevice(device_));
    o.collections_processed = collections_;
    o.collections_total = 0;
    o.points_processed = points_;
    // both read-backs ride one stream sync (the progress callback calls this after every ingest)
    unsigned long long kept = 0;
    if (d_survivors_)
        CU_TRY(cudaMemcpyAsync(&kept, d_survivors_, sizeof kept, cudaMemcpyDeviceToHost, compute_));
    // delta mode: the tiles touched so far = those of the delta in progress, of a push still in flight,
    // and everything the ranks have merged at earlier finalizes
    const size_t nt = std::max(1, n_tiles_);
    const uint32_t* srcs[3] = {d_touched_buf_[0], d_touched_buf_[1], d_touched_merged_};
    std::vector<uint32_t> t(3 * nt, 0);
    if (delta_mode_) ST_TRY(join_fin());
    for (int k = 0; k < 3; ++k)
        if (srcs[k]) CU_TRY(cudaMemcpyAsync(t.data() + k * nt, srcs[k], nt * sizeof(uint32_t), cudaMemcpyDeviceToHost, compute_));
    CU_TRY(cudaStreamSynchronize(compute_));
    o.points_processed += kept;
    uint64_t active = 0;
    for (int i = 0; i < n_tiles_; ++i) active += (t[i] | t[nt + i] | t[2 * nt + i]) ? 1 : 0;
    o.tiles_active = active;
    o.elapsed_seconds = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0_).count();
    return Status::success();
}

}  // namespace pcrb
