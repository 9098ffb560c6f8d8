// abi.cu — the extern "C" surface declared in include/pcr_b200.h.  Thin: argument
// checks, handle casts, Status -> (code, thread-local message).
#include "engine.h"

#include <cmath>
#include <cstring>
#include <new>
#include <stdexcept>

using pcrb::Engine;
using pcrb::Status;

namespace {
thread_local std::string g_last_error;

int finish(const Status& s)
{
    if (!s.ok()) g_last_error = s.message;
    return s.code;
}
int fail(int code, const char* msg)
{
    g_last_error = msg;
    return code;
}
// No C++ exception may cross the extern "C" boundary (a bad_alloc unwinding through ctypes / a C caller
// aborts the host process): every entry that reaches the engine runs under this guard.
template <typename F>
int guarded(F&& f) noexcept
{
    try {
        return f();
    } catch (const std::bad_alloc&) {
        return fail(PCR_OUT_OF_MEMORY, "out of host memory");
    } catch (const std::length_error&) {
        return fail(PCR_OUT_OF_MEMORY, "out of host memory (allocation size overflow)");
    } catch (const std::exception& e) {
        g_last_error = std::string("internal error: ") + e.what();
        return PCR_IO_ERROR;
    } catch (...) {
        return fail(PCR_IO_ERROR, "internal error: unknown exception");
    }
}
#define GUARD(expr) guarded([&]() -> int { return (expr); })
Engine* eng(pcr_pipeline* p) { return reinterpret_cast<Engine*>(p); }
const Engine* eng(const pcr_pipeline* p) { return reinterpret_cast<const Engine*>(p); }
}  // namespace

extern "C" {

const char* pcr_last_error(void) { return g_last_error.c_str(); }
const char* pcr_version(void) { return "pcr-b200 0.1.0 (sm_100a)"; }

int pcr_device_count(void)
{
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int pcr_device_name(int device, char* buf, size_t buflen)
{
    if (!buf || !buflen) return fail(PCR_INVALID_ARGUMENT, "pcr_device_name: null buffer");
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        std::snprintf(buf, buflen, "Unknown GPU");
        return fail(PCR_CUDA_ERROR, "pcr_device_name: cudaGetDeviceProperties failed");
    }
    std::snprintf(buf, buflen, "%s", prop.name);
    return PCR_OK;
}

int pcr_device_mem_info(int device, uint64_t* free_bytes, uint64_t* total_bytes)
{
    size_t f = 0, t = 0;
    if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess)
        return fail(PCR_CUDA_ERROR, "pcr_device_mem_info: no usable device");
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return PCR_OK;
}

// GridConfig::compute_dimensions, src/core/grid_config.cpp:7-22
int pcr_grid_compute_dimensions(pcr_grid_desc* g)
{
    if (!g) return fail(PCR_INVALID_ARGUMENT, "null grid");
    if (!(g->max_x >= g->min_x && g->max_y >= g->min_y)) {
        g->width = g->height = 0;
        return PCR_OK;
    }
    g->width = static_cast<int32_t>(std::ceil((g->max_x - g->min_x) / std::fabs(g->cell_size_x)));
    g->height = static_cast<int32_t>(std::ceil((g->max_y - g->min_y) / std::fabs(g->cell_size_y)));
    return PCR_OK;
}

// GridConfig::world_to_cell, src/core/grid_config.cpp:24-43 (host twin of route_cell)
int pcr_grid_world_to_cell(const pcr_grid_desc* g, double wx, double wy, int32_t* col, int32_t* row)
{
    if (!g || !col || !row) return 0;
    if (!(wx >= g->min_x && wx <= g->max_x && wy >= g->min_y && wy <= g->max_y)) return 0;
    int32_t c = static_cast<int32_t>(std::floor((wx - g->min_x) / g->cell_size_x));
    int32_t r = static_cast<int32_t>(std::floor((wy - g->max_y) / g->cell_size_y));
    c = std::max(0, std::min(c, g->width - 1));
    r = std::max(0, std::min(r, g->height - 1));
    *col = c;
    *row = r;
    return 1;
}

int pcr_mem_alloc(int location, int device, size_t bytes, void** out)
{
    if (!out) return fail(PCR_INVALID_ARGUMENT, "pcr_mem_alloc: null out");
    *out = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaSuccess;
    switch (location) {
    case PCR_MEM_HOST:
        *out = std::malloc(bytes);
        return *out ? PCR_OK : fail(PCR_OUT_OF_MEMORY, "Failed to allocate coordinate arrays");
    case PCR_MEM_HOST_PINNED:
        e = cudaMallocHost(out, bytes);
        break;
    case PCR_MEM_DEVICE:
        e = cudaSetDevice(device);
        if (e == cudaSuccess) e = cudaMalloc(out, bytes);
        break;
    default:
        return fail(PCR_INVALID_ARGUMENT, "pcr_mem_alloc: unknown memory location");
    }
    if (e != cudaSuccess) {
        g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e);
        return PCR_CUDA_ERROR;
    }
    return PCR_OK;
}

int pcr_mem_free(int location, int device, void* ptr)
{
    if (!ptr) return PCR_OK;
    switch (location) {
    case PCR_MEM_HOST: std::free(ptr); return PCR_OK;
    case PCR_MEM_HOST_PINNED: cudaFreeHost(ptr); return PCR_OK;
    case PCR_MEM_DEVICE: cudaSetDevice(device); cudaFree(ptr); return PCR_OK;
    }
    return fail(PCR_INVALID_ARGUMENT, "pcr_mem_free: unknown memory location");
}

int pcr_mem_copy(void* dst, int dst_location, const void* src, int src_location, size_t bytes, int device)
{
    if (bytes == 0) return PCR_OK;
    if (!dst || !src) return fail(PCR_INVALID_ARGUMENT, "pcr_mem_copy: null pointer");
    if (dst_location != PCR_MEM_DEVICE && src_location != PCR_MEM_DEVICE) {
        std::memcpy(dst, src, bytes);
        return PCR_OK;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMemcpy(dst, src, bytes, cudaMemcpyDefault);
    if (e != cudaSuccess) {
        g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e);
        return PCR_CUDA_ERROR;
    }
    return PCR_OK;
}

int pcr_pipeline_create(const pcr_pipeline_desc* desc, pcr_pipeline** out)
{
    if (!out) return fail(PCR_INVALID_ARGUMENT, "pcr_pipeline_create: null out");
    *out = nullptr;
    if (!desc) return fail(PCR_INVALID_ARGUMENT, "pcr_pipeline_create: null desc");
    if (desc->num_reductions < 0 || (desc->num_reductions > 0 && !desc->reductions))
        return fail(PCR_INVALID_ARGUMENT, "pcr_pipeline_create: bad reductions array");
    return guarded([&]() -> int {
        Engine* e = nullptr;
        Status s = Engine::create(*desc, &e);
        *out = reinterpret_cast<pcr_pipeline*>(e);
        return finish(s);
    });
}

void pcr_pipeline_destroy(pcr_pipeline* p)
{
    try { delete eng(p); } catch (...) {}
}

#define NEED(p) if (!(p)) return fail(PCR_INVALID_ARGUMENT, "null pipeline handle")

int pcr_pipeline_validate(const pcr_pipeline* p) { NEED(p); return GUARD(finish(eng(p)->validate())); }

int pcr_pipeline_ingest(pcr_pipeline* p, const double* x, const double* y, size_t count,
                        const pcr_channel_view* channels, int32_t num_channels, int32_t location)
{
    NEED(p);
    if (num_channels < 0 || (num_channels > 0 && !channels))
        return fail(PCR_INVALID_ARGUMENT, "pcr_pipeline_ingest: bad channel array");
    return GUARD(finish(eng(p)->ingest(x, y, count, channels, num_channels, location)));
}

int pcr_pipeline_finalize(pcr_pipeline* p) { NEED(p); return GUARD(finish(eng(p)->finalize(true))); }
int pcr_pipeline_finalize_device(pcr_pipeline* p) { NEED(p); return GUARD(finish(eng(p)->finalize(false))); }

int pcr_pipeline_result_band(pcr_pipeline* p, int32_t band, const float** data, int32_t* rows, int32_t* cols)
{
    NEED(p);
    if (!data || !rows || !cols) return fail(PCR_INVALID_ARGUMENT, "null output pointer");
    return GUARD(finish(eng(p)->result_band(band, data, rows, cols, false)));
}

int pcr_pipeline_result_band_device(pcr_pipeline* p, int32_t band, const float** data, int32_t* rows, int32_t* cols)
{
    NEED(p);
    if (!data || !rows || !cols) return fail(PCR_INVALID_ARGUMENT, "null output pointer");
    return GUARD(finish(eng(p)->result_band(band, data, rows, cols, true)));
}

int pcr_pipeline_band_name(const pcr_pipeline* p, int32_t band, char* buf, size_t buflen)
{
    NEED(p);
    if (!buf || !buflen) return fail(PCR_INVALID_ARGUMENT, "null buffer");
    return guarded([&]() -> int {
        std::string name;
        Status s = eng(p)->band_name(band, name);
        if (s.ok()) std::snprintf(buf, buflen, "%s", name.c_str());
        return finish(s);
    });
}

int pcr_pipeline_stats(const pcr_pipeline* p, pcr_progress* out)
{
    NEED(p);
    if (!out) return fail(PCR_INVALID_ARGUMENT, "null output pointer");
    return GUARD(finish(const_cast<Engine*>(eng(p))->stats(*out)));
}

int pcr_pipeline_set_progress_callback(pcr_pipeline* p, pcr_progress_fn fn, void* user)
{
    NEED(p);
    eng(p)->set_progress(fn, user);
    return PCR_OK;
}

int pcr_pipeline_save_state(pcr_pipeline* p, const char* dir)
{
    NEED(p);
    if (!dir || !dir[0]) return fail(PCR_INVALID_ARGUMENT, "pcr_pipeline_save_state: empty directory");
    return GUARD(finish(eng(p)->save_state(dir)));
}
int pcr_pipeline_load_state(pcr_pipeline* p, const char* dir)
{
    NEED(p);
    if (!dir || !dir[0]) return fail(PCR_INVALID_ARGUMENT, "pcr_pipeline_load_state: empty directory");
    return GUARD(finish(eng(p)->load_state(dir)));
}
int pcr_pipeline_reset(pcr_pipeline* p) { NEED(p); return GUARD(finish(eng(p)->reset())); }
int pcr_pipeline_synchronize(pcr_pipeline* p) { NEED(p); return GUARD(finish(eng(p)->synchronize())); }

int pcr_pipeline_profile_enable(pcr_pipeline* p, int32_t on) { NEED(p); return GUARD(finish(eng(p)->profile_enable(on < 0 ? 0 : on))); }
int pcr_pipeline_profile_reset(pcr_pipeline* p) { NEED(p); return GUARD(finish(eng(p)->profile_reset())); }
int pcr_pipeline_profile_read(pcr_pipeline* p, pcr_profile* out)
{
    NEED(p);
    if (!out) return fail(PCR_INVALID_ARGUMENT, "null output pointer");
    return GUARD(finish(eng(p)->profile_read(*out)));
}

int pcr_pipeline_timer_begin(pcr_pipeline* p) { NEED(p); return GUARD(finish(eng(p)->timer_begin())); }
int pcr_pipeline_timer_end(pcr_pipeline* p, double* elapsed_ms)
{
    NEED(p);
    if (!elapsed_ms) return fail(PCR_INVALID_ARGUMENT, "null output pointer");
    return GUARD(finish(eng(p)->timer_end(*elapsed_ms)));
}

int pcr_comm_unique_id(void* id128)
{
    if (!id128) return fail(PCR_INVALID_ARGUMENT, "null id buffer");
    return GUARD(finish(pcrb::comm_unique_id(id128)));
}

int pcr_comm_slice_rows(int32_t height, int32_t world_size, int32_t rank, int32_t* row0, int32_t* row1)
{
    if (!row0 || !row1 || world_size < 1 || rank < 0 || rank >= world_size || height < 0)
        return fail(PCR_INVALID_ARGUMENT, "pcr_comm_slice_rows: bad arguments");
    pcrb::slice_rows(height, world_size, rank, *row0, *row1);
    return PCR_OK;
}

int pcr_pipeline_comm_init(pcr_pipeline* p, const void* id128, int32_t rank, int32_t world_size)
{
    NEED(p);
    if (!id128 && world_size > 1) return fail(PCR_INVALID_ARGUMENT, "null id buffer");
    return GUARD(finish(eng(p)->comm_init(id128, rank, world_size)));
}

int pcr_comm_partition_cells(uint64_t cells, int32_t record_words, int32_t bin_cells_log2, int32_t world_size, int32_t rank,
                             int32_t* bin_shift, int32_t* num_bins, uint64_t* cell0, uint64_t* cell1)
{
    if (!bin_shift || !num_bins || !cell0 || !cell1 || world_size < 1 || rank < 0 || rank >= world_size || cells == 0 ||
        cells >= (uint64_t(1) << 32) || record_words < 1 || record_words > 8)
        return fail(PCR_INVALID_ARGUMENT, "pcr_comm_partition_cells: bad arguments");
    int shift = 0, nbins = 0;
    pcrb::bin_geometry(static_cast<size_t>(cells), record_words, bin_cells_log2, shift, nbins);
    uint32_t per = 0;
    size_t c0 = 0, c1 = 0;
    pcrb::bin_owner_cells(static_cast<size_t>(cells), shift, nbins, world_size, rank, per, c0, c1);
    *bin_shift = shift; *num_bins = nbins; *cell0 = c0; *cell1 = c1;
    return PCR_OK;
}

int pcr_pipeline_owned_cells(const pcr_pipeline* p, uint64_t* cell0, uint64_t* cell1)
{
    NEED(p);
    if (!cell0 || !cell1) return fail(PCR_INVALID_ARGUMENT, "null output pointer");
    eng(p)->owned_cells(*cell0, *cell1);
    return PCR_OK;
}

int pcr_pipeline_comm_barrier(pcr_pipeline* p) { NEED(p); return GUARD(finish(eng(p)->comm_barrier())); }

}  // extern "C"
