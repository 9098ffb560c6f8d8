// engine_ext.cu — the two optional legs of the Engine:
//   * deterministic mode (sort -> in-order segmented reduce), and
//   * the multi-GPU combine at finalize (NCCL transport over NVLink + the fused
//     merge/finalize kernel).
#include "engine.h"

#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
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

// ---------------------------------------------------------------------------
// Deterministic mode
// ---------------------------------------------------------------------------
static int bits_for(size_t max_value)
{
    int b = 1;
    while ((max_value >> b) != 0) ++b;
    return b;
}

Status Engine::ensure_sort_scratch(size_t n)
{
    if (n <= sort_capacity_) return Status::success();
    CU_TRY(cudaStreamSynchronize(compute_));
    cudaFree(d_sort_tmp_); cudaFree(d_keys_); cudaFree(d_keys_alt_); cudaFree(d_idx_); cudaFree(d_idx_alt_);
    d_sort_tmp_ = nullptr; d_keys_ = d_keys_alt_ = d_idx_ = d_idx_alt_ = nullptr;
    sort_capacity_ = 0;
    const size_t cap = std::max(n, slot_points_);
    sort_tmp_bytes_ = det_sort_temp_bytes(cap, bits_for(cells_));
    CU_TRY(cudaMalloc(&d_sort_tmp_, std::max<size_t>(sort_tmp_bytes_, 16)));
    CU_TRY(cudaMalloc(&d_keys_, cap * 4));
    CU_TRY(cudaMalloc(&d_keys_alt_, cap * 4));
    CU_TRY(cudaMalloc(&d_idx_, cap * 4));
    CU_TRY(cudaMalloc(&d_idx_alt_, cap * 4));
    sort_capacity_ = cap;
    return Status::success();
}

Status Engine::run_passes_deterministic(const double* dx, const double* dy, size_t n,
                                        const std::vector<const float*>& cp, const uint8_t* mask)
{
    // point indices are u32 payloads: split very large device clouds
    const size_t kMaxChunk = size_t(1) << 30;
    for (size_t p0 = 0; p0 < n; p0 += kMaxChunk) {
        const size_t cnt = std::min(kMaxChunk, n - p0);
        ST_TRY(ensure_sort_scratch(cnt));
        const int key_bits = bits_for(cells_);   // keys are cell ids, `cells_` marks invalid points
        prof_begin(PROF_SORT, compute_);
        CU_TRY(det_build_keys(compute_, mask ? mask + p0 : nullptr, dx + p0, dy + p0, cnt, gp_, d_keys_, d_idx_, d_touched_));
        CU_TRY(det_sort(compute_, d_sort_tmp_, sort_tmp_bytes_, d_keys_, d_keys_alt_, d_idx_, d_idx_alt_,
                        cnt, key_bits));
        prof_end(compute_);
        launches_ += 2;   // key build + radix sort (CUB launches several kernels; counted as one)
        prof_begin(PROF_ACC, compute_);
        for (Pass& p : passes_) {
            if (p.glyph.type != PCR_GLYPH_POINT) continue;      // glyph passes run in run_passes
            ChannelPtrs ch{};
            for (size_t c = 0; c < p.channels.size(); ++c) ch.p[c] = cp[channel_slot(p.channels[c])] + p0;
            CU_TRY(det_point_reduce(compute_, d_keys_, d_idx_, cnt, ch, p.d_state, p.layout,
                                    static_cast<uint32_t>(cells_)));
            ++launches_;
        }
        prof_end(compute_);
        prof_points_ += cnt;
    }
    return Status::success();
}

// ---------------------------------------------------------------------------
// NCCL (dlopen'ed: the library must load and run on a single GPU without it)
// ---------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    typedef struct { char internal[128]; } UniqueId;
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void** comm, int nranks, UniqueId id, int rank) = nullptr;
    int (*CommDestroy)(void* comm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int dtype, int peer, void* comm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int dtype, int peer, void* comm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int dtype, int op, void* comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int dtype, void* comm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    // ncclDataType_t / ncclRedOp_t values (nccl.h)
    static constexpr int kUint32 = 3, kFloat32 = 7, kUint8 = 1;
    static constexpr int kSum = 0, kMax = 2;
};

static NcclApi* nccl_load(std::string& err)
{
    static NcclApi api;
    static bool tried = false;
    if (api.handle) return &api;
    if (tried) { err = "NCCL library not available"; return nullptr; }
    tried = true;
    // libnccl.so.2 resolves to the copy already mapped into the process (e.g. the
    // one bundled with PyTorch) when there is one, else to the system library.
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.handle) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return nullptr; }
    auto sym = [&](const char* n) { return dlsym(api.handle, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.GroupStart || !api.GroupEnd ||
        !api.Send || !api.Recv || !api.AllReduce || !api.AllGather || !api.GetErrorString) {
        err = "libnccl.so.2 lacks a required symbol";
        api.handle = nullptr;
        return nullptr;
    }
    return &api;
}

#define NC_TRY(expr)                                                                      \
    do {                                                                                  \
        int _r = (expr);                                                                  \
        if (_r != 0)                                                                      \
            return Status::error(PCR_CUDA_ERROR, std::string("NCCL error: ") +            \
                                 nccl_->GetErrorString(_r) + " (" #expr ")");             \
    } while (0)

Status comm_unique_id(void* id128)
{
    std::string err;
    NcclApi* api = nccl_load(err);
    if (!api) return Status::error(PCR_CUDA_ERROR, err);
    NcclApi::UniqueId id;
    const int r = api->GetUniqueId(&id);
    if (r != 0) return Status::error(PCR_CUDA_ERROR, std::string("NCCL error: ") + api->GetErrorString(r));
    std::memcpy(id128, id.internal, 128);
    return Status::success();
}

void engine_comm_destroy(NcclApi* api, void* comm)
{
    if (api && comm) api->CommDestroy(comm);
}

Status Engine::comm_init(const void* id128, int rank, int world)
{
    if (world < 1 || world > kMaxParts || rank < 0 || rank >= world)
        return Status::error(PCR_INVALID_ARGUMENT, "pipeline: world size must be 1..8 and 0 <= rank < world");
    if (world == 1) { rank_ = 0; world_ = 1; return Status::success(); }
    std::string err;
    nccl_ = nccl_load(err);
    if (!nccl_) return Status::error(PCR_CUDA_ERROR, err);
    CU_TRY(cudaSetDevice(device_));
    NcclApi::UniqueId id;
    std::memcpy(id.internal, id128, 128);
    NC_TRY(nccl_->CommInitRank(&comm_, world, id, rank));
    rank_ = rank;
    world_ = world;
    if (staging_auto_ && !pool_) staging_threads_ = default_staging_threads(world);   // ranks share one host
    CU_TRY(cudaMalloc(&d_touched_all_, std::max(1, n_tiles_) * sizeof(uint32_t)));
    if (exact_) {            // mode 2 combines with integer all-reduces (any order gives the same bits): no peer mapping
        for (Pass& p : passes_) {
            p.xa_sum = p.xa;
            p.xa_sum.limbs = nullptr; p.xa_sum.flags = nullptr; p.xa_sum.ext = nullptr;
            if (p.layout.n_add) {
                CU_TRY(cudaMalloc(&p.xa_sum.limbs, xacc_limb_bytes(cells_, p.layout)));
                CU_TRY(cudaMalloc(&p.xa_sum.flags, xacc_flag_bytes(cells_, p.layout)));
            }
            if (p.layout.n_max + p.layout.n_min) CU_TRY(cudaMalloc(&p.xa_sum.ext, xacc_ext_bytes(cells_, p.layout)));
        }
        return Status::success();
    }
    if (comm_mode_ != 1) {
        Status s = peer_map();
        if (!s.ok() && (comm_mode_ == 2 || comm_layout_ == 2)) return s;     // peer memory was demanded
    }
    if (comm_layout_ == 2 && !partition_)
        return Status::error(PCR_INVALID_ARGUMENT,
                             "pipeline: the tile-partitioned layout (comm_layout = 2) needs peer memory between all GPUs and "
                             "tile-binned Point passes only (point_kernel = 3, or a grid whose records exceed 256 MB)");
    return Status::success();
}

// ---------------------------------------------------------------------------
// Peer-memory path: map every rank's combine / touched-staging / band / flag buffers with CUDA IPC so
// that the push kernel stores partial records straight into their owner's memory over NVLink and the
// merge+finalize kernel stores the finalized slice straight into the peers' band arrays — no staging
// buffer, no separate collective.  NCCL is used once, to all-gather the 64-byte IPC handles.
// ---------------------------------------------------------------------------
Status Engine::peer_map()
{
    peer_ok_ = false;
    const bool part = partition_wanted();      // tile-partitioned layout: no partial grids, no combine buffers
    int* d_flag = nullptr;
    CU_TRY(cudaMalloc(&d_flag, sizeof(int) * std::max(world_, 2)));
    // every rank must reach every collective below whatever happens locally: failures are recorded and
    // agreed on by an all-reduce(min) instead of returning early (a rank that left would hang the others)
    auto agree = [&](int mine, int& all) -> Status {
        CU_TRY(cudaMemcpyAsync(d_flag, &mine, sizeof(int), cudaMemcpyHostToDevice, compute_));
        NC_TRY(nccl_->AllReduce(d_flag, d_flag, 1, NcclApi::kUint32, /*ncclMin*/ 3, comm_, compute_));
        CU_TRY(cudaMemcpyAsync(&all, d_flag, sizeof(int), cudaMemcpyDeviceToHost, compute_));
        CU_TRY(cudaStreamSynchronize(compute_));
        return Status::success();
    };
    struct FreeFlag { int* p; ~FreeFlag() { cudaFree(p); } } free_flag{d_flag};

    // 1. can this rank reach every other device?
    int can_all = 1;
    {
        std::vector<int> devs(world_, 0);
        CU_TRY(cudaMemcpyAsync(d_flag + rank_, &device_, sizeof(int), cudaMemcpyHostToDevice, compute_));
        NC_TRY(nccl_->AllGather(d_flag + rank_, d_flag, sizeof(int), NcclApi::kUint8, comm_, compute_));
        CU_TRY(cudaMemcpyAsync(devs.data(), d_flag, sizeof(int) * world_, cudaMemcpyDeviceToHost, compute_));
        CU_TRY(cudaStreamSynchronize(compute_));
        for (int k = 0; k < world_; ++k) {
            if (k == rank_) continue;
            int ok = 0;
            if (devs[k] == device_ || cudaDeviceCanAccessPeer(&ok, device_, devs[k]) != cudaSuccess || !ok) can_all = 0;
        }
        cudaGetLastError();
    }
    int all_can = 0;
    ST_TRY(agree(can_all, all_can));
    if (!all_can)
        return Status::error(PCR_CUDA_ERROR, "pipeline: peer-memory combine needs P2P access between all ranks' GPUs");

    // 2. local buffers: combine buffers (the push targets), the second delta buffer, the owner slice,
    //    touched staging; one ok flag for all of them
    const size_t rows_per = (static_cast<size_t>(grid_.height) + world_ - 1) / world_;
    const size_t max_slice = rows_per * static_cast<size_t>(grid_.width);
    const size_t nt = std::max(1, n_tiles_);
    int r0, r1;
    slice_rows(grid_.height, world_, rank_, r0, r1);
    const size_t my_cells = static_cast<size_t>(r1 - r0) * grid_.width;
    auto alloc_local = [&]() -> Status {
        if (!d_flags_) {
            CU_TRY(cudaMalloc(&d_flags_, (2 * kMaxParts + 4) * sizeof(uint32_t)));   // + the done counter
            CU_TRY(cudaMemsetAsync(d_flags_, 0, (2 * kMaxParts + 4) * sizeof(uint32_t), compute_));
        }
        for (Pass& p : passes_) {
            if (part) break;
            const size_t W = p.layout.width;
            if (!p.d_combined) CU_TRY(cudaMalloc(&p.d_combined, 2 * static_cast<size_t>(world_) * max_slice * W * 4));   // [epoch parity][rank][cell]
            if (!p.d_delta[1]) {
                CU_TRY(cudaMalloc(&p.d_delta[1], cells_ * W * 4));
                CU_TRY(launch_init_state(compute_, p.d_delta[1], cells_, p.layout));
            }
            if (!p.d_owned) {
                CU_TRY(cudaMalloc(&p.d_owned, std::max<size_t>(my_cells, 1) * W * 4));
                CU_TRY(launch_init_state(compute_, p.d_owned, my_cells, p.layout));
            }
        }
        if (!d_touched_buf_[1]) {
            CU_TRY(cudaMalloc(&d_touched_buf_[1], nt * 4));
            CU_TRY(cudaMemsetAsync(d_touched_buf_[1], 0, nt * 4, compute_));
        }
        if (!d_touched_merged_) {
            CU_TRY(cudaMalloc(&d_touched_merged_, nt * 4));
            CU_TRY(cudaMemsetAsync(d_touched_merged_, 0, nt * 4, compute_));
        }
        if (!d_touched_stage_) {
            CU_TRY(cudaMalloc(&d_touched_stage_, 2 * static_cast<size_t>(world_) * nt * 4));                             // [epoch parity][rank][tile]
            CU_TRY(cudaMemsetAsync(d_touched_stage_, 0, 2 * static_cast<size_t>(world_) * nt * 4, compute_));
        }
        if (!e_delta_) CU_TRY(cudaEventCreateWithFlags(&e_delta_, cudaEventDisableTiming));
        if (!push_) {
            int lo = 0, hi = 0;
            CU_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CU_TRY(cudaStreamCreateWithPriority(&push_, cudaStreamNonBlocking, hi));
        }
        return Status::success();
    };
    Status local = alloc_local();
    if (!local.ok()) cudaGetLastError();

    // 3. handles: [pass combine buffers..., touched staging, out, flags]
    const size_t n_comb = part ? 0 : passes_.size();
    const size_t n_buf = n_comb + 3;
    std::vector<cudaIpcMemHandle_t> mine(n_buf);
    if (local.ok()) {
        auto get = [&]() -> Status {
            for (size_t i = 0; i < n_comb; ++i) CU_TRY(cudaIpcGetMemHandle(&mine[i], passes_[i].d_combined));
            CU_TRY(cudaIpcGetMemHandle(&mine[n_comb], d_touched_stage_));
            CU_TRY(cudaIpcGetMemHandle(&mine[n_comb + 1], d_out_));
            CU_TRY(cudaIpcGetMemHandle(&mine[n_comb + 2], d_flags_));
            return Status::success();
        };
        local = get();
    }
    int all_local = 0;
    ST_TRY(agree(local.ok() ? 1 : 0, all_local));
    if (!all_local)
        return local.ok() ? Status::error(PCR_CUDA_ERROR, "pipeline: a peer rank could not allocate its combine buffers") : local;

    const size_t bytes = n_buf * sizeof(cudaIpcMemHandle_t);
    std::vector<cudaIpcMemHandle_t> all(n_buf * world_);
    unsigned char* d_h = nullptr;
    CU_TRY(cudaMalloc(&d_h, bytes * world_));
    CU_TRY(cudaMemcpyAsync(d_h + bytes * rank_, mine.data(), bytes, cudaMemcpyHostToDevice, compute_));
    NC_TRY(nccl_->AllGather(d_h + bytes * rank_, d_h, bytes, NcclApi::kUint8, comm_, compute_));
    CU_TRY(cudaMemcpyAsync(all.data(), d_h, bytes * world_, cudaMemcpyDeviceToHost, compute_));
    CU_TRY(cudaStreamSynchronize(compute_));
    cudaFree(d_h);

    // 4. open the peers' buffers; on any failure everyone closes what it opened and falls back together
    auto open_all = [&]() -> Status {
        for (int k = 0; k < world_; ++k) {
            PeerBuffers& pb = peer_[k];
            pb = PeerBuffers{};
            pb.combined.assign(passes_.size(), nullptr);
            if (k == rank_) {
                for (size_t i = 0; i < passes_.size(); ++i) pb.combined[i] = passes_[i].d_combined;
                pb.touched_stage = d_touched_stage_; pb.out = d_out_; pb.flags = d_flags_;
                continue;
            }
            const cudaIpcMemHandle_t* h = &all[n_buf * k];
            auto open = [&](const cudaIpcMemHandle_t& hh, void** out) {
                return cudaIpcOpenMemHandle(out, hh, cudaIpcMemLazyEnablePeerAccess);
            };
            for (size_t i = 0; i < n_comb; ++i) CU_TRY(open(h[i], reinterpret_cast<void**>(&pb.combined[i])));
            CU_TRY(open(h[n_comb], reinterpret_cast<void**>(&pb.touched_stage)));
            CU_TRY(open(h[n_comb + 1], reinterpret_cast<void**>(&pb.out)));
            CU_TRY(open(h[n_comb + 2], reinterpret_cast<void**>(&pb.flags)));
        }
        return Status::success();
    };
    Status opened = open_all();
    if (!opened.ok()) cudaGetLastError();
    int all_open = 0;
    ST_TRY(agree(opened.ok() ? 1 : 0, all_open));     // also: nobody signals into a flag array that is not zeroed yet
    if (!all_open) {
        peer_close_handles();
        return opened.ok() ? Status::error(PCR_CUDA_ERROR, "pipeline: a peer rank could not map the combine buffers") : opened;
    }
    peer_ok_ = true;
    delta_mode_ = !part;
    if (part) return partition_setup();
    return Status::success();
}

// Collective: every rank hands in the same number of device allocations (nullptr allowed, at the same
// positions on every rank) and gets everybody's back, mapped into its address space (its own as they are).
Status Engine::ipc_exchange(const std::vector<void*>& mine, std::vector<std::vector<void*>>& all)
{
    const size_t n = mine.size();
    std::vector<cudaIpcMemHandle_t> h(n);
    Status local = Status::success();
    for (size_t i = 0; i < n && local.ok(); ++i) {
        std::memset(&h[i], 0, sizeof h[i]);
        if (!mine[i]) continue;
        const cudaError_t e = cudaIpcGetMemHandle(&h[i], mine[i]);
        if (e != cudaSuccess) local = Status::error(PCR_CUDA_ERROR, std::string("CUDA error: ") + cudaGetErrorString(e) + " (cudaIpcGetMemHandle)");
    }
    const size_t bytes = n * sizeof(cudaIpcMemHandle_t);
    std::vector<cudaIpcMemHandle_t> got(n * world_);
    unsigned char* d_h = nullptr;
    CU_TRY(cudaMalloc(&d_h, std::max<size_t>(bytes, 16) * world_));
    CU_TRY(cudaMemcpyAsync(d_h + bytes * rank_, h.data(), bytes, cudaMemcpyHostToDevice, compute_));
    NC_TRY(nccl_->AllGather(d_h + bytes * rank_, d_h, bytes, NcclApi::kUint8, comm_, compute_));
    CU_TRY(cudaMemcpyAsync(got.data(), d_h, bytes * world_, cudaMemcpyDeviceToHost, compute_));
    CU_TRY(cudaStreamSynchronize(compute_));
    cudaFree(d_h);
    all.assign(world_, std::vector<void*>(n, nullptr));
    for (int k = 0; k < world_ && local.ok(); ++k)
        for (size_t i = 0; i < n && local.ok(); ++i) {
            if (!mine[i]) continue;
            if (k == rank_) { all[k][i] = mine[i]; continue; }
            const cudaError_t e = cudaIpcOpenMemHandle(&all[k][i], got[n * k + i], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) local = Status::error(PCR_CUDA_ERROR, std::string("CUDA error: ") + cudaGetErrorString(e) + " (cudaIpcOpenMemHandle)");
            else ipc_opened_.push_back(all[k][i]);
        }
    if (!local.ok()) cudaGetLastError();
    // agree on the outcome: a rank that failed must not leave the others with half a mapping
    uint32_t* d_ok = nullptr;
    CU_TRY(cudaMalloc(&d_ok, 4));
    const uint32_t mine_ok = local.ok() ? 1u : 0u;
    uint32_t all_ok = 0;
    CU_TRY(cudaMemcpyAsync(d_ok, &mine_ok, 4, cudaMemcpyHostToDevice, compute_));
    ST_TRY(nccl_allreduce_min_u32(d_ok, 1));
    CU_TRY(cudaMemcpyAsync(&all_ok, d_ok, 4, cudaMemcpyDeviceToHost, compute_));
    CU_TRY(cudaStreamSynchronize(compute_));
    cudaFree(d_ok);
    if (!all_ok) return local.ok() ? Status::error(PCR_CUDA_ERROR, "pipeline: a peer rank could not map the entry pools") : local;
    return Status::success();
}

Status Engine::nccl_allreduce_min_u32(uint32_t* d, size_t count)
{
    NC_TRY(nccl_->AllReduce(d, d, count, NcclApi::kUint32, /*ncclMin*/ 3, comm_, compute_));
    return Status::success();
}

void Engine::peer_close_handles()
{
    for (void* q : ipc_opened_) cudaIpcCloseMemHandle(q);
    ipc_opened_.clear();
    for (int k = 0; k < world_; ++k) {
        if (k == rank_) continue;
        for (uint32_t* p : peer_[k].combined) if (p) cudaIpcCloseMemHandle(p);
        if (peer_[k].touched_stage) cudaIpcCloseMemHandle(peer_[k].touched_stage);
        if (peer_[k].out) cudaIpcCloseMemHandle(peer_[k].out);
        if (peer_[k].flags) cudaIpcCloseMemHandle(peer_[k].flags);
        peer_[k] = PeerBuffers{};
    }
    cudaGetLastError();
}

void Engine::peer_unmap()
{
    if (!peer_ok_) return;
    cudaSetDevice(device_);
    if (comm_ && nccl_) {   // every rank must be done with everyone's memory before it is unmapped/freed
        nccl_->AllReduce(d_touched_all_, d_touched_all_, 1, NcclApi::kUint32, NcclApi::kMax, comm_, compute_);
        cudaStreamSynchronize(compute_);
    }
    peer_close_handles();
    if (comm_ && nccl_) {   // ... and every mapping must be closed before the owners free the memory
        nccl_->AllReduce(d_touched_all_, d_touched_all_, 1, NcclApi::kUint32, NcclApi::kMax, comm_, compute_);
        cudaStreamSynchronize(compute_);
    }
    peer_ok_ = false;
}

// N>1 finalize over peer memory, "delta epochs".  No host sync anywhere; three streams.
//
// What a rank accumulates between two finalizes is a DELTA (records start from the identity); the rank
// that owns a row slice keeps the running merge of everybody's deltas (d_owned).  Because Op::merge is
// a commutative monoid (builtin_ops.h:15,28,41,54,67,95-97) the result is the same state a single
// cumulative grid would hold.  Delta buffers, combine buffers and touched staging are all double-buffered
// by epoch parity, so NOTHING of a finalize sits on the ingest stream and push(e+1) runs under merge(e):
//
// compute stream  [ingest kernels of epoch e] -> event "delta e complete" -> [ingest kernels of e+1 ...]
// push stream     behind that event:
//   k_peer_wait     the peers are done reading the combine buffers of this parity (phase 1 of epoch e-2)
//   k_push_slices   persistent copy kernel: every record of the delta goes into the combine buffer of the
//                   rank that owns its row (posted NVLink writes; a local copy for my own slice) and is put
//                   back to the identity behind the copy; my touched-tile flags go into everyone's staging;
//                   the last CTA releases phase 0 on every rank
// merge stream    (highest priority) behind my own push:
//   k_peer_wait_merge_touched   every rank's push has landed here; OR the touched flags into the merged set
//   k_finalize_peer persistent kernel: owned = merge(owned, delta of rank 0, 1, ...) in rank order, finalize,
//                   store the bands into my array and the peers' arrays; the last CTA releases phase 1
//   (k_peer_wait)   peer_quiesce(): phase 1 from everyone, before the host reads the bands.
// The compute stream waits for the others only where it must: before it accumulates into the delta buffer
// whose push was enqueued one finalize earlier, before a D2H of the bands, in synchronize() and timer_end().
Status Engine::finalize_multi_peer()
{
    ++epoch_;
    const int par = static_cast<int>(epoch_ & 1u);
    PeerSync ps{};
    ps.pf.n = world_; ps.pf.rank = rank_;
    ps.pt.n = world_;
    ps.epoch = epoch_;
    const size_t nt = std::max(1, n_tiles_);
    for (int k = 0; k < world_; ++k) {
        ps.pf.flags[k] = peer_[k].flags;
        ps.pt.touched[k] = d_touched_stage_ + (static_cast<size_t>(par) * world_ + k) * nt;   // local staging of this parity
    }
    const size_t rows_per = (static_cast<size_t>(grid_.height) + world_ - 1) / world_;
    const size_t max_slice = rows_per * static_cast<size_t>(grid_.width);
    int r0, r1;
    slice_rows(grid_.height, world_, rank_, r0, r1);
    const size_t my0 = static_cast<size_t>(r0) * grid_.width, my_cells = static_cast<size_t>(r1 - r0) * grid_.width;

    // ---- compute stream: the delta of this epoch is complete; flip to the other buffer ----
    const int old = cur_;
    CU_TRY(cudaEventRecord(e_delta_, compute_));
    // the other buffer was pushed (and reset) by the previous finalize: its push must be done before the
    // next ingest accumulates into it (in steady state that event is long past)
    if (epoch_ > 1) CU_TRY(cudaStreamWaitEvent(compute_, e_pushed_, 0));
    cur_ ^= 1;
    for (Pass& p : passes_) p.d_state = p.d_delta[cur_];
    d_touched_ = d_touched_buf_[cur_];

    // ---- push stream ----
    CU_TRY(cudaStreamWaitEvent(push_, e_delta_, 0));
    ps.waited = 1;
    ps.done_counter = reinterpret_cast<unsigned int*>(d_flags_ + 2 * kMaxParts);
    prof_begin(PROF_PUSH, push_);
    // The combine buffers of this parity were last read by the peers' merge of epoch e-2.
    if (epoch_ > 2) { CU_TRY(launch_peer_wait(push_, ps.pf, 1, epoch_ - 2)); ++launches_; }
    if (passes_.empty()) CU_TRY(launch_peer_signal(push_, ps.pf, 0, epoch_));
    for (size_t i = 0; i < passes_.size(); ++i) {
        const size_t W = passes_[i].layout.width;
        PushTargets pt{};
        for (int k = 0; k < world_; ++k) {
            pt.combined[k] = peer_[k].combined[i] + static_cast<size_t>(par) * world_ * max_slice * W;
            pt.touched_stage[k] = peer_[k].touched_stage + static_cast<size_t>(par) * world_ * nt;
        }
        pt.rows_per = static_cast<int>(rows_per);
        pt.max_slice_cells = max_slice;
        CU_TRY(launch_push_slices(push_, passes_[i].d_delta[old], d_touched_buf_[old], n_tiles_, gp_, passes_[i].layout, pt, ps,
                                  i + 1 == passes_.size(), i + 1 == passes_.size(), true, sm_count_));
        ++launches_;
    }
    prof_end(push_);
    CU_TRY(cudaEventRecord(e_pushed_, push_));

    // ---- merge stream ----
    CU_TRY(cudaStreamWaitEvent(fin_, e_pushed_, 0));
    ps.done_counter = reinterpret_cast<unsigned int*>(d_flags_ + 2 * kMaxParts + 1);     // the two kernels may overlap
    prof_begin(PROF_FIN, fin_);
    // One small kernel holds the stream until every peer's push has landed (the merge kernel's own CTAs
    // would all spin on the same flags and occupy the SMs the next ingest wants) and ORs the ranks'
    // touched-tile flags into the merged set.
    CU_TRY(launch_peer_wait_merge_touched(fin_, ps.pf, epoch_, ps.pt, d_touched_merged_, std::max(1, n_tiles_), true));
    ++launches_;
    ps.pt.n = 1;
    ps.pt.touched[0] = d_touched_merged_;
    for (size_t i = 0; i < reductions_.size(); ++i)
        if (reductions_[i].rejected)
            CU_TRY(cudaMemsetAsync(d_out_ + i * cells_, 0xFF, cells_ * sizeof(float), fin_));
    // Where do my slice's bands go besides my own array?  By default the merge kernel stores them straight
    // into the peers' arrays (one NVLink latency, no extra launch, overlapped with the merge itself).
    // comm_band_copy = 2 writes them locally and ships one copy-engine transfer per band and peer instead
    // (measured on config 5 in round 1: no gain at 8 GPUs, a loss at 2), so it is opt-in.
    std::vector<int> targets;
    for (int k = 0; k < world_; ++k)
        if (k != rank_ && !(gather_root_only_ && k != 0) && !bands_distributed_) targets.push_back(k);
    const bool bulk_bands = band_copy_ == 2;
    OutTargets outs{};
    outs.out[outs.n++] = d_out_;
    if (!bulk_bands)
        for (int k : targets) outs.out[outs.n++] = peer_[k].out;
    if (passes_.empty() && !bulk_bands) CU_TRY(launch_peer_signal(fin_, ps.pf, 1, epoch_));
    for (size_t i = 0; i < passes_.size(); ++i) {
        Pass& p = passes_[i];
        const size_t W = p.layout.width;
        StateParts parts{};
        parts.n = world_ + 1;
        parts.part[0] = p.d_owned;
        for (int k = 0; k < world_; ++k)
            parts.part[k + 1] = p.d_combined + (static_cast<size_t>(par) * world_ + k) * max_slice * W;
        ps.signal_begin = 0;
        ps.signal_end = !bulk_bands && i + 1 == passes_.size();
        CU_TRY(launch_finalize_peer(fin_, parts, my0, my0, my_cells, outs, cells_, gp_, p.layout, p.fin, ps, p.d_owned, sm_count_));
        ++launches_;
    }
    if (bulk_bands) {
        for (int k : targets)
            for (size_t b = 0; b < reductions_.size(); ++b)
                if (!reductions_[b].rejected)
                    CU_TRY(cudaMemcpyAsync(peer_[k].out + b * cells_ + my0, d_out_ + b * cells_ + my0,
                                           my_cells * sizeof(float), cudaMemcpyDeviceToDevice, fin_));
        CU_TRY(launch_peer_signal(fin_, ps.pf, 1, epoch_));      // "done", behind the copies in stream order
        ++launches_;
    }
    prof_end(fin_);
    CU_TRY(cudaEventRecord(e_fin_, fin_));
    fin_pending_ = true;
    // Not awaited here: the peers' "done" flags of this epoch.  They are awaited by the push after next
    // (before it overwrites the peers' combine buffers of this parity) and by peer_quiesce() before the
    // host looks at the bands.
    return Status::success();
}

Status Engine::peer_quiesce()
{
    if (!peer_ok_ || waited_epoch_ == epoch_) return Status::success();
    PeerFlags pf{};
    pf.n = world_; pf.rank = rank_;
    for (int k = 0; k < world_; ++k) pf.flags[k] = peer_[k].flags;
    CU_TRY(launch_peer_wait(fin_, pf, 1, epoch_));
    CU_TRY(cudaEventRecord(e_fin_, fin_));
    fin_pending_ = true;
    ++launches_;
    waited_epoch_ = epoch_;
    return Status::success();
}

// Deterministic mode 2 on N ranks: the exact states are integers, so ANY reduction order gives the same
// bits — NCCL's all-reduce (ring, tree or NVLS in-switch) is used as it is: sum over the int64 limbs, max
// over the flag bytes and the max words, min over the min words.  Every rank then finalizes the whole grid.
Status Engine::finalize_multi_exact()
{
    constexpr int kInt32 = 2, kInt64 = 4, kMin = 3;
    const size_t nt = std::max(1, n_tiles_);
    NC_TRY(nccl_->AllReduce(d_touched_, d_touched_all_, nt, NcclApi::kUint32, NcclApi::kMax, comm_, compute_));
    prof_begin(PROF_PUSH, compute_);
    for (Pass& p : passes_) {
        const PassLayout& L = p.layout;
        if (L.n_add) {
            NC_TRY(nccl_->AllReduce(p.xa.limbs, p.xa_sum.limbs, cells_ * L.n_add * kXLimbs, kInt64, NcclApi::kSum, comm_, compute_));
            NC_TRY(nccl_->AllReduce(p.xa.flags, p.xa_sum.flags, xacc_flag_bytes(cells_, L), NcclApi::kUint8, NcclApi::kMax, comm_, compute_));
        }
        if (L.n_max) NC_TRY(nccl_->AllReduce(p.xa.ext, p.xa_sum.ext, cells_ * L.n_max, kInt32, NcclApi::kMax, comm_, compute_));
        if (L.n_min) NC_TRY(nccl_->AllReduce(p.xa.ext + cells_ * L.n_max, p.xa_sum.ext + cells_ * L.n_max, cells_ * L.n_min, kInt32, kMin, comm_, compute_));
    }
    prof_end(compute_);
    prof_begin(PROF_FIN, compute_);
    for (size_t i = 0; i < reductions_.size(); ++i)
        if (reductions_[i].rejected)
            CU_TRY(cudaMemsetAsync(d_out_ + i * cells_, 0xFF, cells_ * sizeof(float), compute_));
    for (Pass& p : passes_) {
        CU_TRY(launch_finalize_exact(compute_, p.xa_sum, 0, cells_, d_out_, cells_, gp_, p.layout, p.fin, d_touched_all_));
        ++launches_;
    }
    prof_end(compute_);
    return Status::success();
}

Status Engine::alloc_exact(Pass& p)
{
    p.xa.cells = cells_;
    const PassLayout& L = p.layout;
    if (L.n_add) {
        CU_TRY(cudaMalloc(&p.xa.limbs, xacc_limb_bytes(cells_, L)));
        CU_TRY(cudaMalloc(&p.xa.flags, xacc_flag_bytes(cells_, L)));
    }
    if (L.n_max + L.n_min) CU_TRY(cudaMalloc(&p.xa.ext, xacc_ext_bytes(cells_, L)));
    return Status::success();
}

Status Engine::run_passes_exact(const double* dx, const double* dy, size_t n, const std::vector<const float*>& cp,
                                const uint8_t* mask)
{
    prof_begin(PROF_ACC, compute_);
    for (Pass& p : passes_) {
        ChannelPtrs ch{};
        for (size_t c = 0; c < p.channels.size(); ++c) ch.p[c] = cp[channel_slot(p.channels[c])];
        if (p.glyph.type == PCR_GLYPH_POINT) {
            CU_TRY(launch_point_exact(compute_, mask, dx, dy, ch, n, p.xa, gp_, p.layout, d_touched_));
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
                CU_TRY(launch_line_accumulate(compute_, mask, dx, dy, ch, g, n, nullptr, gp_, p.layout, d_touched_, &p.xa));
            else
                CU_TRY(launch_gaussian_accumulate(compute_, mask, dx, dy, ch, g, n, nullptr, gp_, p.layout, d_touched_, &p.xa));
        }
        ++launches_;
    }
    prof_end(compute_);
    prof_points_ += n;
    return Status::success();
}

Status Engine::finalize_multi()
{
    if (exact_) return finalize_multi_exact();
    if (partition_) return finalize_multi_part();
    return peer_ok_ ? finalize_multi_peer() : finalize_multi_nccl();
}

Status Engine::comm_barrier()
{
    if (world_ == 1) return synchronize();
    CU_TRY(cudaSetDevice(device_));
    NC_TRY(nccl_->AllReduce(d_touched_all_, d_touched_all_, 1, NcclApi::kUint32, NcclApi::kMax, comm_, compute_));
    CU_TRY(cudaStreamSynchronize(compute_));
    return Status::success();
}

// Multi-GPU finalize.  Rank k owns the row slice [row0_k, row1_k).  Every rank
// ships its partial records of slice j to rank j (NCCL send/recv = an all-to-all
// over NVLink, same volume as a reduce-scatter but layout-agnostic: records mix
// f32 sums with ordered-s32 max/min words, which no single ncclRedOp covers);
// the owner merges the `world` parts in rank order inside the finalize kernel
// (Op::merge, builtin_ops.h:15,28,41,54,67,95-97 — fixed order, so the float sums
// do not depend on arrival order); the finalized slices are then exchanged so
// every rank ends with complete bands.
Status Engine::finalize_multi_nccl()
{
    const size_t rows_per = (static_cast<size_t>(grid_.height) + world_ - 1) / world_;
    auto row0 = [&](int k) -> size_t { int a, b; slice_rows(grid_.height, world_, std::min(k, world_ - 1), a, b);
                                       return static_cast<size_t>(k >= world_ ? b : a); };
    auto slice_cells = [&](int k) { return (row0(k + 1) - row0(k)) * static_cast<size_t>(grid_.width); };
    const size_t max_slice = rows_per * static_cast<size_t>(grid_.width);
    const size_t my0 = row0(rank_) * grid_.width, my_cells = slice_cells(rank_);

    NC_TRY(nccl_->AllReduce(d_touched_, d_touched_all_, std::max(1, n_tiles_), NcclApi::kUint32,
                            NcclApi::kMax, comm_, compute_));

    for (size_t i = 0; i < reductions_.size(); ++i)
        if (reductions_[i].rejected)
            CU_TRY(cudaMemsetAsync(d_out_ + i * cells_, 0xFF, cells_ * sizeof(float), compute_));

    for (Pass& p : passes_) {
        const size_t W = p.layout.width;
        if (!p.d_combined) CU_TRY(cudaMalloc(&p.d_combined, static_cast<size_t>(world_) * max_slice * W * 4));
        NC_TRY(nccl_->GroupStart());
        for (int k = 0; k < world_; ++k) {
            if (k == rank_) continue;
            if (slice_cells(k))
                NC_TRY(nccl_->Send(p.d_state + row0(k) * grid_.width * W, slice_cells(k) * W, NcclApi::kUint32, k, comm_, compute_));
            if (my_cells)
                NC_TRY(nccl_->Recv(p.d_combined + static_cast<size_t>(k) * max_slice * W, my_cells * W, NcclApi::kUint32, k, comm_, compute_));
        }
        NC_TRY(nccl_->GroupEnd());

        StateParts parts{};
        parts.n = world_;
        for (int k = 0; k < world_; ++k)
            parts.part[k] = (k == rank_) ? p.d_state + my0 * W : p.d_combined + static_cast<size_t>(k) * max_slice * W;
        prof_begin(PROF_FIN, compute_);
        OutTargets outs{};
        outs.out[0] = d_out_;
        outs.n = 1;
        CU_TRY(launch_finalize(compute_, parts, my0, my0, my_cells, outs, cells_, gp_, p.layout, p.fin, d_touched_all_));
        ++launches_;
        prof_end(compute_);
    }

    // all-gather of the finalized slices, band by band
    NC_TRY(nccl_->GroupStart());
    for (size_t b = 0; b < reductions_.size(); ++b) {
        if (reductions_[b].rejected) continue;
        float* band = d_out_ + b * cells_;
        for (int k = 0; k < world_; ++k) {
            if (k == rank_) continue;
            if (my_cells) NC_TRY(nccl_->Send(band + my0, my_cells, NcclApi::kFloat32, k, comm_, compute_));
            if (slice_cells(k)) NC_TRY(nccl_->Recv(band + row0(k) * grid_.width, slice_cells(k), NcclApi::kFloat32, k, comm_, compute_));
        }
    }
    NC_TRY(nccl_->GroupEnd());
    return Status::success();
}

}  // namespace pcrb
