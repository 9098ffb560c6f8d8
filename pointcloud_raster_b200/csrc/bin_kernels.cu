// bin_kernels.cu — tile binning for the Point glyph on grids whose records do not fit L2, sm_100a.
//
// Reference counterpart: TileRouter::sort + extract_batches (src/engine/tile_router.cpp:138-366,
// tile_router_kernels.cu:169-293) — the reference sorts every cloud by (tile, cell) so that a tile's
// points are contiguous, then accumulates tile by tile.  Here the idea is kept and the sort is not:
//
//   * the grid's cell index space is cut into BINS of 2^shift consecutive cells whose records
//     (<= 64 MB) stay resident in L2 while they are being updated;
//   * k_bin_scatter routes each point ONCE (the CPU rule, common.cuh) and appends a compact entry
//     {cell u32, value f32 ...} to the page chain of its bin.  A CTA stages 4096 points in shared
//     memory, counting-sorts them by bin there (shared-memory integer atomics return the rank), and
//     copies every bin's run out with coalesced stores into pages it owns exclusively — no global
//     atomics per point, no inter-CTA waiting, one global atomic per 4096-entry page;
//   * nothing is accumulated at ingest time.  Entries pile up (8 B per point and channel, pool sized
//     from free HBM) until finalize — or until the pool is full — and only then k_bin_accumulate walks
//     the pages bin by bin and issues the same reductions k_point_direct would have issued.  Deferring
//     maximises the number of points per record sector per visit: every sector of a bin is fetched from
//     HBM once and written back once per flush instead of once per point (profiles/r02_c5_*: the direct
//     kernel moves 130 B of DRAM traffic per 20-B point on the 20000^2 grid).
//
// The same pages can live in ANOTHER rank's pool (peer memory): that is the tile-partitioned layout
// with an all-to-all point exchange (engine_ext.cu, partition mode).
#include "point_kernels_impl.cuh"
#include "bin_kernels.cuh"

namespace pcrb {

namespace {

using point_impl::RecordShape;
using point_impl::pick;
using point_impl::update_record;

// A scatter CTA has THREADS threads and stages THREADS * kBinPts points per chunk; 32 warps per SM either way
// (<= 64 registers): 2 CTAs of 512 threads or 4 CTAs of 256, whose load -> sort -> store phases hide each other.
constexpr int kBinPts = 8;
constexpr int kBinThreadsMax = 512;
static_assert(kBinThreadsMax * kBinPts <= static_cast<int>(kBinPageEntries), "a chunk's run must fit two pages");
static_assert(kBinPadChunk <= 256 * kBinPts, "capacity formulas assume the smallest chunk");
constexpr uint32_t kNoPage = 0xffffffffu;

// entry = {cell, values...} as one vector (see BinPool::ent)
template <int NCH> struct EntryOf;
template <> struct EntryOf<0> { using type = uint32_t; static constexpr int words = 1; };
template <> struct EntryOf<1> { using type = uint2;    static constexpr int words = 2; };
template <> struct EntryOf<2> { using type = uint4;    static constexpr int words = 4; };
constexpr uint32_t kNullCell = 0xffffffffu;                 // padding entry: skipped by the fold
template <int NCH>
__device__ __forceinline__ typename EntryOf<NCH>::type make_null_entry()
{
    if constexpr (NCH == 0) return kNullCell;
    else if constexpr (NCH == 1) return make_uint2(kNullCell, 0u);
    else return make_uint4(kNullCell, 0u, 0u, 0u);
}
// shared memory -> global memory through the bulk-copy engine (TMA; SASS UBLKCP.G.S); 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_store(void* dst_global, const void* src_shared, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(dst_global), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(src_shared))), "r"(bytes) : "memory");
}
template <int NCH>
__device__ __forceinline__ typename EntryOf<NCH>::type make_entry(uint32_t cell, const float* v)
{
    if constexpr (NCH == 0) return cell;
    else if constexpr (NCH == 1) return make_uint2(cell, __float_as_uint(v[0]));
    else return make_uint4(cell, __float_as_uint(v[0]), __float_as_uint(v[1]), 0u);
}

// Exclusive scan of the counts s_hist[0..nbins), each rounded up to a multiple of `align` entries (so that
// every bin's run starts 16-byte aligned in the staging buffer), into s_prefix; nbins <= 4 * kBinThreads.
template <int kBinThreads>
__device__ __forceinline__ uint32_t block_scan_bins(const uint32_t* s_hist, uint32_t* s_prefix, int nbins,
                                                    uint32_t* s_warp, uint32_t align)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int b = tid * 4 + k;
        v[k] = b < nbins ? ((s_hist[b] + align - 1) & ~(align - 1)) : 0u;
        sum += v[k];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kBinThreads / 32 ? s_warp[lane] : 0u;
#pragma unroll
        for (int d = 1; d < kBinThreads / 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t;
        }
        if (lane < kBinThreads / 32) s_warp[lane] = w;     // inclusive over warps
    }
    __syncthreads();
    uint32_t run = inc - sum + (warp ? s_warp[warp - 1] : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int b = tid * 4 + k;
        if (b < nbins) s_prefix[b] = run;
        run += v[k];
    }
    return s_warp[kBinThreads / 32 - 1];                    // total
}

// One CTA = a persistent worker with its own page chain per bin (open_page / open_fill rows in global
// memory survive from launch to launch, so a chain's only partly filled page is its last one).
template <int NCH, bool EXACT, int kBinThreads>
__global__ void __launch_bounds__(kBinThreads, 1024 / kBinThreads)
k_bin_scatter(const uint8_t* __restrict__ mask, const double* __restrict__ xs, const double* __restrict__ ys,
              const __grid_constant__ ChannelPtrs ch, size_t n, const __grid_constant__ GridParams g,
              const __grid_constant__ BinTargets bt, uint32_t* __restrict__ touched, int prefetch_next)
{
    constexpr int kBinChunk = kBinThreads * kBinPts;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbins = bt.nbins;
    uint32_t* s_hist   = reinterpret_cast<uint32_t*>(smem_raw);          // [nbins]  points of this chunk per bin
    uint32_t* s_prefix = s_hist + nbins;                                  // [nbins]  exclusive scan
    // per bin, one 128-bit load in the sort step: {prefix, ranks served by the first page, entry index of
    // rank 0 in the first page, entry index of rank 0 in the second page minus len_a | owner << 28 ... }
    const int nb4 = (2 * nbins + 3) & ~3;                                 // keep the table 16-byte aligned
    uint4*    s_tab    = reinterpret_cast<uint4*>(s_hist + nb4);          // [nbins]  {prefix, len_a, seg_a, seg_b - len_a}
    uint32_t* s_pool   = reinterpret_cast<uint32_t*>(s_tab + nbins);      // [nbins]  which pool (owner rank) the bin lives in
    uint32_t* s_open_page = s_pool + nbins;                               // [nbins]  this CTA's page chains, kept in shared
    uint32_t* s_open_fill = s_open_page + nbins;                          // [nbins]  memory for the life of the launch
    uint32_t* s_warp   = s_open_fill + nbins;                             // [32]
    using Entry = typename EntryOf<NCH>::type;
    constexpr uint32_t kAlign = 16 / sizeof(Entry);                       // entries per 16 bytes: runs start and end on it
    const uint32_t stride = (kBinChunk + static_cast<uint32_t>(nbins) * (kAlign - 1) + 3u) & ~3u;   // one staging buffer, 16-byte multiple
    Entry* st_base = reinterpret_cast<Entry*>(s_hist + ((nb4 + 7 * nbins + 32 + 3) & ~3));   // [2][stride], 16-byte aligned

    const int tid = threadIdx.x;
    uint32_t* my_open_page = bt.open_page + static_cast<size_t>(blockIdx.x) * nbins;
    uint32_t* my_open_fill = bt.open_fill + static_cast<size_t>(blockIdx.x) * nbins;
    const size_t nchunks = (n + kBinChunk - 1) / kBinChunk;
    const bool multi_tile = g.tiles_x * g.tiles_y > 1;
    const bool few_tiles = g.tiles_x * g.tiles_y <= 32;    // touched tiles collected in a register, flagged once at the end
    uint32_t tile_bits = 0;
    bool any_valid = false;
    for (int b = tid; b < nbins; b += kBinThreads) { s_open_page[b] = my_open_page[b]; s_open_fill[b] = my_open_fill[b]; }

    int buf = 0;
    for (size_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x, buf ^= 1) {
        Entry* st_ent = st_base + static_cast<size_t>(buf) * stride;      // {cell, values} of this chunk in bin order
        for (int b = tid; b < nbins; b += kBinThreads) s_hist[b] = 0;
        __syncthreads();

        // ---- route; rank every point inside its bin with a shared-memory atomic ----
        uint32_t cell[kBinPts], key[kBinPts];       // key = bin << 13 | rank  (rank < 4096), or ~0 = dropped
        float val[kBinPts][NCH > 0 ? NCH : 1];
        const size_t base = chunk * kBinChunk + tid;
        // FULL: the whole chunk lies inside the cloud and no filter mask is set (every chunk but the last one of an
        // unfiltered ingest) — no per-point bounds or mask tests
        auto route_chunk = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
            for (int q = 0; q < kBinPts; q += 4) {
                double x[4], y[4];
                bool live[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const size_t i = base + static_cast<size_t>(q + u) * kBinThreads;
                    live[u] = FULL || (i < n && (mask == nullptr || mask[i] != 0));
                    x[u] = live[u] ? ldg_stream_d(xs + i) : 0.0;
                    y[u] = live[u] ? ldg_stream_d(ys + i) : 0.0;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) val[q + u][c] = live[u] ? ldg_stream_f(ch.p[c] + i) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    int col = 0, row = 0;
                    const bool ok = route_cell<EXACT>(g, x[u], y[u], col, row) && live[u];
                    const uint32_t c = static_cast<uint32_t>(row) * static_cast<uint32_t>(g.width) + static_cast<uint32_t>(col);
                    cell[q + u] = c;
                    key[q + u] = 0xffffffffu;
                    if (ok) {
                        const uint32_t bin = c >> bt.shift;
                        key[q + u] = (bin << 13) | atomicAdd(&s_hist[bin], 1u);
                        any_valid = true;
                        if (multi_tile) {                    // touched-tile rule, tile_manager.cpp:437-444
                            const int t = tile_of(g, col, row);
                            if (few_tiles) tile_bits |= 1u << t;
                            else if (touched[t] == 0) touched[t] = 1;
                        }
                    }
                }
            }
        };
        if (mask == nullptr && (chunk + 1) * kBinChunk <= n) route_chunk(std::true_type{});
        else route_chunk(std::false_type{});
        if (prefetch_next) {
            // pull the next chunk of this CTA into L2 while this one is sorted and stored (one 128-byte line per
            // instruction: x and y are kBinChunk / 16 lines each, a value channel kBinChunk / 32)
            const size_t nb = (chunk + gridDim.x) * kBinChunk;
            constexpr int kLines = kBinChunk / 16;            // = kBinThreads / 2
            const int l = tid < kLines ? tid : tid - kLines;
            const size_t i = nb + static_cast<size_t>(l) * 16;
            if (i < n) {
                asm volatile("prefetch.global.L2 [%0];" :: "l"((tid < kLines ? xs : ys) + i));
                if (NCH > 0 && (l & 1) == 0) {
                    const int c = tid < kLines ? 0 : 1;
                    if (c < NCH) asm volatile("prefetch.global.L2 [%0];" :: "l"(ch.p[c] + i));
                }
            }
        }
        __syncthreads();

        // ---- where does each bin's run go?  (page chain of this CTA; at most two pages per run) ----
        block_scan_bins<kBinThreads>(s_hist, s_prefix, nbins, s_warp, kAlign);
        // this staging buffer was the source of the bulk stores issued two chunks ago: they must have read it
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();                                   // (also: s_prefix[b] is read by another thread than its writer)
        for (int b = tid; b < nbins; b += kBinThreads) {
            const uint32_t cnt = s_hist[b];
            if (cnt == 0) continue;
            const uint32_t cnt2 = (cnt + kAlign - 1) & ~(kAlign - 1);      // with the null entries that pad the run
            const int owner = bt.bin_owner_shift >= 0 ? static_cast<int>(static_cast<uint32_t>(b) / bt.bins_per_owner) : 0;
            const BinPool& pool = bt.pool[owner];
            uint32_t page = s_open_page[b], fill = s_open_fill[b];
            auto alloc = [&]() -> uint32_t {
                uint32_t idx = atomicAdd(pool.next_page, 1u);           // always local memory (see BinPool)
                if (idx >= pool.sub_pages) { *pool.overflow = 1u; idx = pool.sub_pages - 1; }
                const uint32_t pg = pool.page_base + idx;
                pool.page_bin[pg] = static_cast<uint32_t>(b);           // (a posted NVLink store when the pool is a peer's)
                return pg;
            };
            if (page == kNoPage || fill == kBinPageEntries) { page = alloc(); fill = 0; }
            const uint32_t len_a = min(cnt2, kBinPageEntries - fill);      // fill, cnt2 and the page size are multiples of kAlign
            const uint32_t seg_a = page * kBinPageEntries + fill;
            uint32_t seg_b = 0;
            pool.page_fill[page] = fill + len_a;
            fill += len_a;
            if (cnt2 > len_a) {
                page = alloc();
                fill = cnt2 - len_a;
                seg_b = page * kBinPageEntries;
                pool.page_fill[page] = fill;
            }
            const uint32_t off = s_prefix[b];
            s_tab[b] = make_uint4(off, len_a, seg_a, seg_b);
            s_pool[b] = static_cast<uint32_t>(owner) | (cnt2 << 8);
            for (uint32_t k = cnt; k < cnt2; ++k) st_ent[off + k] = make_null_entry<NCH>();
            s_open_page[b] = page;
            s_open_fill[b] = fill;
        }
        __syncthreads();

        // ---- counting sort into the staging buffer ----
#pragma unroll
        for (int k = 0; k < kBinPts; ++k) {
            if (key[k] == 0xffffffffu) continue;
            const uint32_t bin = key[k] >> 13, rank = key[k] & 8191u;
            st_ent[s_tab[bin].x + rank] = make_entry<NCH>(cell[k], val[k]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the bulk-copy engine
        __syncthreads();

        // ---- copy out: every run leaves as one or two bulk stores (TMA, UBLKCP): shared memory -> the owner's page,
        //      local HBM or NVLink peer memory, in large packets and without occupying any thread ----
        for (int b = tid; b < nbins; b += kBinThreads) {
            if (s_hist[b] == 0) continue;
            const uint4 tb = s_tab[b];
            const uint32_t pw = s_pool[b];
            const BinPool& pool = bt.pool[pw & 0xffu];
            const uint32_t cnt2 = pw >> 8;
            Entry* dst = reinterpret_cast<Entry*>(pool.ent);
            bulk_store(dst + tb.z, st_ent + tb.x, tb.y * static_cast<uint32_t>(sizeof(Entry)));
            if (cnt2 > tb.y) bulk_store(dst + tb.w, st_ent + tb.x + tb.y, (cnt2 - tb.y) * static_cast<uint32_t>(sizeof(Entry)));
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");              // every store has landed before the kernel ends
    for (int b = tid; b < nbins; b += kBinThreads) { my_open_page[b] = s_open_page[b]; my_open_fill[b] = s_open_fill[b]; }
    if (!multi_tile) {
        if (__any_sync(0xffffffffu, any_valid) && (tid & 31) == 0 && touched[0] == 0) touched[0] = 1;
    } else if (few_tiles) {
        tile_bits = __reduce_or_sync(0xffffffffu, tile_bits);
        if ((tid & 31) == 0)
            for (uint32_t m = tile_bits; m; m &= m - 1) { const int t = __ffs(m) - 1; if (touched[t] == 0) touched[t] = 1; }
    }
}

// ---- flush: order the pages by bin, then fold them ----
// page p of the pool is in use iff it lies in the used prefix of its source's slice
__device__ __forceinline__ bool page_used(const BinPool& pool, uint32_t p)
{
    const uint32_t r = p / pool.sub_pages;
    return r < static_cast<uint32_t>(pool.n_src) && p - r * pool.sub_pages < min(pool.src_count[r], pool.sub_pages);
}

__global__ void k_bin_page_count(const __grid_constant__ BinPool pool, uint32_t* __restrict__ bin_pages)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < pool.pool_pages; p += gridDim.x * blockDim.x)
        if (page_used(pool, p)) atomicAdd(&bin_pages[pool.page_bin[p]], 1u);
}

// single CTA: exclusive scan of bin_pages[0..nbins) -> bin_first; bin_pages becomes the running cursor;
// ctrl[3] = pages to fold
__global__ void k_bin_page_scan(const __grid_constant__ BinPool pool, uint32_t* __restrict__ bin_pages,
                                uint32_t* __restrict__ bin_first, int nbins)
{
    __shared__ uint32_t s[kMaxBins];
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) s[b] = bin_pages[b];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 0; b < nbins; ++b) { const uint32_t c = s[b]; s[b] = run; run += c; }
        pool.ctrl[3] = run;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) { bin_first[b] = s[b]; bin_pages[b] = 0; }
}

__global__ void k_bin_page_order(const __grid_constant__ BinPool pool, const uint32_t* __restrict__ bin_first,
                                 uint32_t* __restrict__ bin_cursor, uint32_t* __restrict__ order)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < pool.pool_pages; p += gridDim.x * blockDim.x) {
        if (!page_used(pool, p)) continue;
        const uint32_t b = pool.page_bin[p];
        order[bin_first[b] + atomicAdd(&bin_cursor[b], 1u)] = p;
    }
}

// CTA per page, pages in bin order: at any time the resident CTAs work on neighbouring bins, whose
// records stay in L2.  Same reductions per entry as k_point_direct (update_record).
constexpr int kAccThreads = 256;
template <int NADD, int NMAX, int NMIN, int NCH>
__global__ void __launch_bounds__(kAccThreads)
k_bin_accumulate(const __grid_constant__ BinPool pool, const uint32_t* __restrict__ order,
                 uint32_t* __restrict__ state, size_t cell_base, const __grid_constant__ PassLayout L)
{
    // Pages are handed out through a global cursor, NOT by a static stride: persistent CTAs drift apart
    // (tools/micro/bin_locality.cu: 25 Gpts/s with the static stride, 72 Gpts/s when the pages in flight are
    // always the contiguous frontier of the bin-ordered list), and only a compact frontier keeps the
    // records of the bins being folded resident in L2.
    const uint32_t np = pool.ctrl[3];
    __shared__ uint32_t s_next;
    for (;;) {
        if (threadIdx.x == 0) s_next = atomicAdd(pool.ctrl + 2, 1u);
        __syncthreads();
        const uint32_t i = s_next;
        __syncthreads();
        if (i >= np) break;
        const uint32_t page = order[i];
        const uint32_t cnt = min(pool.page_fill[page], kBinPageEntries);
        const uint32_t e0 = page * kBinPageEntries;
#pragma unroll 4
        for (uint32_t t = threadIdx.x; t < cnt; t += kAccThreads) {
            const auto ent = __ldcs(reinterpret_cast<const typename EntryOf<NCH>::type*>(pool.ent) + e0 + t);
            uint32_t cell;
            float v[kMaxChan] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (NCH == 0) cell = ent;
            else if constexpr (NCH == 1) { cell = ent.x; v[0] = __uint_as_float(ent.y); }
            else { cell = ent.x; v[0] = __uint_as_float(ent.y); v[1] = __uint_as_float(ent.z); }
            if (cell == kNullCell) continue;                   // alignment padding of a run
            float add[kMaxAdd], mx[kMaxExt], mn[kMaxExt];
#pragma unroll
            for (int j = 0; j < kMaxAdd; ++j) add[j] = (j < NADD) ? (L.add_src[j] < 0 ? 1.0f : pick(v, L.add_src[j])) : 0.0f;
#pragma unroll
            for (int j = 0; j < kMaxExt; ++j) mx[j] = (j < NMAX) ? pick(v, L.max_src[j]) : 0.0f;
#pragma unroll
            for (int j = 0; j < kMaxExt; ++j) mn[j] = (j < NMIN) ? pick(v, L.min_src[j]) : 0.0f;
            update_record<NADD, NMAX, NMIN>(state, static_cast<size_t>(cell) - cell_base, add, mx, mn);
        }
    }
}

// everything back to "empty pool"; the CTA chains forget their open pages
__global__ void k_bin_reset(const __grid_constant__ BinPool pool, uint32_t* __restrict__ open_page, size_t n_open)
{
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i == 0) { pool.ctrl[0] = 0; pool.ctrl[2] = 0; pool.ctrl[3] = 0; }   // pages handed out (one GPU), fold cursor, fold total
    if (i < static_cast<size_t>(pool.n_src)) pool.src_count[i] = 0;
    for (size_t k = i; k < n_open; k += static_cast<size_t>(gridDim.x) * blockDim.x) open_page[k] = kNoPage;
}

template <int NADD, int NMAX, int NMIN>
cudaError_t acc_nch(cudaStream_t s, unsigned grid, const BinPool& pool, const uint32_t* order, uint32_t* state,
                    size_t cell_base, const PassLayout& L)
{
    switch (L.n_chan) {
    case 0: k_bin_accumulate<NADD, NMAX, NMIN, 0><<<grid, kAccThreads, 0, s>>>(pool, order, state, cell_base, L); break;
    case 1: k_bin_accumulate<NADD, NMAX, NMIN, 1><<<grid, kAccThreads, 0, s>>>(pool, order, state, cell_base, L); break;
    case 2: k_bin_accumulate<NADD, NMAX, NMIN, 2><<<grid, kAccThreads, 0, s>>>(pool, order, state, cell_base, L); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int NADD, int NMAX>
cudaError_t acc_min(cudaStream_t s, unsigned grid, const BinPool& pool, const uint32_t* order, uint32_t* state,
                    size_t cell_base, const PassLayout& L)
{
    switch (L.n_min) {
    case 0: if constexpr (NADD + NMAX > 0) return acc_nch<NADD, NMAX, 0>(s, grid, pool, order, state, cell_base, L);
            else return cudaErrorInvalidValue;
    case 1: return acc_nch<NADD, NMAX, 1>(s, grid, pool, order, state, cell_base, L);
    case 2: return acc_nch<NADD, NMAX, 2>(s, grid, pool, order, state, cell_base, L);
    }
    return cudaErrorInvalidValue;
}

template <int NADD>
cudaError_t acc_max(cudaStream_t s, unsigned grid, const BinPool& pool, const uint32_t* order, uint32_t* state,
                    size_t cell_base, const PassLayout& L)
{
    switch (L.n_max) {
    case 0: return acc_min<NADD, 0>(s, grid, pool, order, state, cell_base, L);
    case 1: return acc_min<NADD, 1>(s, grid, pool, order, state, cell_base, L);
    case 2: return acc_min<NADD, 2>(s, grid, pool, order, state, cell_base, L);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

bool bin_supported(const PassLayout& L) { return L.n_chan <= kBinMaxChan; }

size_t bin_scatter_smem(int nbins, int n_chan, int threads)
{
    const size_t kBinChunk = static_cast<size_t>(threads) * kBinPts;
    const size_t nb4 = (2 * static_cast<size_t>(nbins) + 3) & ~size_t(3);
    const size_t ew = bin_entry_words(n_chan), align = 4 / ew;              // entries per 16 bytes
    const size_t stride = (kBinChunk + static_cast<size_t>(nbins) * (align - 1) + 3) & ~size_t(3);
    return ((nb4 + 7 * static_cast<size_t>(nbins) + 32 + 3) & ~size_t(3)) * 4 + 2 * stride * ew * 4;
}

unsigned bin_scatter_grid(int sm_count, int nbins, int n_chan, int threads)
{
    // CTAs per SM by shared memory (227 KB usable), at most 1024 threads (64 registers each)
    const size_t per = bin_scatter_smem(nbins, n_chan, threads);
    const size_t by_threads = 1024 / static_cast<size_t>(threads);
    const unsigned by_smem = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(by_threads, (220 * 1024) / per)));
    return static_cast<unsigned>(sm_count) * by_smem;
}

cudaError_t launch_bin_scatter(cudaStream_t s, const uint8_t* mask, const double* x, const double* y,
                               const ChannelPtrs& ch, size_t n, const GridParams& g, const PassLayout& L,
                               const BinTargets& bt, uint32_t* touched, unsigned grid, int threads)
{
    if (n == 0) return cudaSuccess;
    if (threads != 256 && threads != 512) return cudaErrorInvalidValue;
    const size_t smem = bin_scatter_smem(bt.nbins, L.n_chan, threads);
    const bool exact = g.exact_x && g.exact_y;
    const int pf = 1;       // L2 prefetch of the CTA's next chunk: 8.3 -> 7.4 ms per 1B points (profiles/r02s3_bin_sweep.txt)
#define PCR_BIN_LAUNCH_T(NCH, EX, T)                                                                        \
    do {                                                                                                    \
        auto kern = k_bin_scatter<NCH, EX, T>;                                                              \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                             static_cast<int>(smem));                                      \
        if (e != cudaSuccess) return e;                                                                     \
        kern<<<grid, T, smem, s>>>(mask, x, y, ch, n, g, bt, touched, pf);                                  \
    } while (0)
#define PCR_BIN_LAUNCH(NCH, EX)                                                                             \
    do { if (threads == 256) PCR_BIN_LAUNCH_T(NCH, EX, 256); else PCR_BIN_LAUNCH_T(NCH, EX, 512); } while (0)
    switch (L.n_chan * 2 + (exact ? 1 : 0)) {
    case 0: PCR_BIN_LAUNCH(0, false); break;
    case 1: PCR_BIN_LAUNCH(0, true); break;
    case 2: PCR_BIN_LAUNCH(1, false); break;
    case 3: PCR_BIN_LAUNCH(1, true); break;
    case 4: PCR_BIN_LAUNCH(2, false); break;
    case 5: PCR_BIN_LAUNCH(2, true); break;
    default: return cudaErrorInvalidValue;
    }
#undef PCR_BIN_LAUNCH
#undef PCR_BIN_LAUNCH_T
    return cudaGetLastError();
}

cudaError_t launch_bin_flush(cudaStream_t s, const BinPool& pool, int nbins, uint32_t* bin_pages, uint32_t* bin_first,
                             uint32_t* order, uint32_t* state, size_t cell_base, const PassLayout& L,
                             uint32_t* open_page, size_t n_open, int sm_count)
{
    const unsigned g1 = static_cast<unsigned>(sm_count) * 4;
    k_bin_page_count<<<g1, 256, 0, s>>>(pool, bin_pages);
    k_bin_page_scan<<<1, 256, 0, s>>>(pool, bin_pages, bin_first, nbins);
    k_bin_page_order<<<g1, 256, 0, s>>>(pool, bin_first, bin_pages, order);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const unsigned ga = static_cast<unsigned>(sm_count) * 8;
    switch (L.n_add) {
    case 0: e = acc_max<0>(s, ga, pool, order, state, cell_base, L); break;
    case 1: e = acc_max<1>(s, ga, pool, order, state, cell_base, L); break;
    case 2: e = acc_max<2>(s, ga, pool, order, state, cell_base, L); break;
    case 3: e = acc_max<3>(s, ga, pool, order, state, cell_base, L); break;
    case 4: e = acc_max<4>(s, ga, pool, order, state, cell_base, L); break;
    default: e = cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    // bin_pages was the scatter cursor of k_bin_page_order: zero it for the next flush
    e = cudaMemsetAsync(bin_pages, 0, static_cast<size_t>(nbins) * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    k_bin_reset<<<g1, 256, 0, s>>>(pool, open_page, n_open);
    return cudaGetLastError();
}

cudaError_t launch_bin_reset(cudaStream_t s, const BinPool& pool, uint32_t* open_page, size_t n_open, int sm_count)
{
    k_bin_reset<<<static_cast<unsigned>(sm_count) * 4, 256, 0, s>>>(pool, open_page, n_open);
    return cudaGetLastError();
}

}  // namespace pcrb
