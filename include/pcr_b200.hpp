// pcr_b200.hpp — the reference's C++ API for the ingest/finalize path,
// header-only over the C-ABI of pcr_b200.h.  Link with -lpcr_b200; no CUDA headers needed.
//
// Same names, argument meaning and error behaviour as the reference's public headers, so a
// C++ caller (or the reference's own tests/cpp/test_pipeline.cpp) compiles against this file
// by switching one include:
//     pcr::Pipeline / PipelineConfig / ReductionSpec / ExecutionMode / ProgressInfo
//                                           include/pcr/engine/pipeline.h:20-145
//     pcr::GlyphSpec / GlyphType            include/pcr/engine/glyph.h:7-33
//     pcr::FilterSpec / CompareOp           include/pcr/engine/filter.h:19-51
//     pcr::PointCloud                       include/pcr/core/point_cloud.h:29-103
//     pcr::Grid / BandDesc                  include/pcr/core/grid.h:14-96
//     pcr::GridConfig                       include/pcr/core/grid_config.h:10-62
//     pcr::BBox / CRS / Status / enums      include/pcr/core/types.h:20-135
//     pcr::write_geotiff / GeoTiffOptions   include/pcr/io/grid_io.h
// Everything below the API is the B200 engine: there is no CPU execution path.
// ExecutionMode::CPU makes Pipeline::create return nullptr (reason on stderr), and
// gpu_fallback_to_cpu is accepted and never honoured.
#ifndef PCR_B200_HPP
#define PCR_B200_HPP

#include "pcr_b200.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <functional>
#include <limits>
#include <memory>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace pcr {

// ---- enums (numeric values are the C-ABI's) --------------------------------------------
enum class DataType : uint8_t { Float32, Float64, Int32, UInt32, Int16, UInt16, UInt8 };
enum class ReductionType : uint8_t {
    Sum, Max, Min, Average, WeightedAverage, Count, Median, Percentile, MostRecent, PriorityMerge, Custom
};
enum class MemoryLocation : uint8_t { Host, HostPinned, Device };
enum class ExecutionMode : uint8_t { CPU, GPU, Auto, Hybrid };
enum class StatusCode : uint8_t { Ok, InvalidArgument, OutOfMemory, CudaError, IoError, CrsError, NotImplemented };
enum class CompareOp : uint8_t { Equal, NotEqual, Less, LessEqual, Greater, GreaterEqual, InSet, NotInSet };
enum class GlyphType : uint8_t { Point, Line, Gaussian };

inline size_t data_type_size(DataType t)
{
    switch (t) {
    case DataType::Float64: return 8;
    case DataType::Int16: case DataType::UInt16: return 2;
    case DataType::UInt8: return 1;
    default: return 4;
    }
}

struct Status {
    StatusCode  code = StatusCode::Ok;
    std::string message;
    bool ok() const { return code == StatusCode::Ok; }
    static Status success() { return {}; }
    static Status error(StatusCode c, const std::string& msg) { return {c, msg}; }
};

namespace detail {
inline Status from_rc(int rc)
{
    if (rc == PCR_OK) return Status::success();
    const char* m = pcr_last_error();
    return Status::error(static_cast<StatusCode>(rc), m ? m : "");
}
}  // namespace detail

struct BBox {
    double min_x = std::numeric_limits<double>::max();
    double min_y = std::numeric_limits<double>::max();
    double max_x = std::numeric_limits<double>::lowest();
    double max_y = std::numeric_limits<double>::lowest();
    BBox() = default;
    BBox(double x0, double y0, double x1, double y1) : min_x(x0), min_y(y0), max_x(x1), max_y(y1) {}

    void expand(double x, double y)
    {
        min_x = std::min(min_x, x); min_y = std::min(min_y, y);
        max_x = std::max(max_x, x); max_y = std::max(max_y, y);
    }
    void expand(const BBox& o) { if (o.valid()) { expand(o.min_x, o.min_y); expand(o.max_x, o.max_y); } }
    bool contains(double x, double y) const { return x >= min_x && x <= max_x && y >= min_y && y <= max_y; }
    double width()  const { return max_x - min_x; }
    double height() const { return max_y - min_y; }
    bool   valid()  const { return max_x >= min_x && max_y >= min_y; }
};

// No PROJ on this path: a CRS is carried (EPSG code into the GeoTIFF GeoKeys), never transformed.
struct CRS {
    std::string wkt;
    int         epsg = 0;
    bool is_valid() const { return !wkt.empty() || epsg != 0; }
    static CRS from_epsg(int code) { CRS c; c.epsg = code; c.wkt = "EPSG:" + std::to_string(code); return c; }
    static CRS from_wkt(const std::string& w) { CRS c; c.wkt = w; return c; }
    bool equivalent_to(const CRS& o) const { return (epsg != 0 && epsg == o.epsg) || (!wkt.empty() && wkt == o.wkt); }
};

struct TileIndex {
    int row = 0, col = 0;
    bool operator==(const TileIndex& o) const { return row == o.row && col == o.col; }
    bool operator<(const TileIndex& o) const { return row != o.row ? row < o.row : col < o.col; }
};

// ---- GridConfig ---------------------------------------------------------------------------
struct GridConfig {
    BBox   bounds;
    CRS    crs;
    double cell_size_x = 1.0;
    double cell_size_y = -1.0;
    int    width = 0, height = 0;
    int    tile_width = 4096, tile_height = 4096;
    int    tiles_x = 0, tiles_y = 0;

    pcr_grid_desc desc() const
    {
        return { bounds.min_x, bounds.min_y, bounds.max_x, bounds.max_y, cell_size_x, cell_size_y,
                 width, height, tile_width, tile_height };
    }
    void compute_dimensions()
    {
        if (!bounds.valid()) { width = height = tiles_x = tiles_y = 0; return; }
        pcr_grid_desc d = desc();
        if (pcr_grid_compute_dimensions(&d) != PCR_OK) { width = height = tiles_x = tiles_y = 0; return; }
        width = d.width; height = d.height;
        tiles_x = (width + tile_width - 1) / tile_width;
        tiles_y = (height + tile_height - 1) / tile_height;
    }
    // The cell-index contract (CPU rule, src/core/grid_config.cpp:24-43) — evaluated by the library.
    bool world_to_cell(double wx, double wy, int& col, int& row) const
    {
        pcr_grid_desc d = desc();
        int32_t c = 0, r = 0;
        const bool ok = pcr_grid_world_to_cell(&d, wx, wy, &c, &r) != 0;
        col = c; row = r;
        return ok;
    }
    void cell_to_world(int col, int row, double& wx, double& wy) const
    {
        wx = bounds.min_x + (col + 0.5) * cell_size_x;
        wy = bounds.max_y + (row + 0.5) * cell_size_y;
    }
    TileIndex cell_to_tile(int col, int row) const { return { row / tile_height, col / tile_width }; }
    void tile_cell_range(TileIndex idx, int& col_start, int& row_start, int& col_count, int& row_count) const
    {
        col_start = idx.col * tile_width;
        row_start = idx.row * tile_height;
        col_count = std::min(tile_width, width - col_start);
        row_count = std::min(tile_height, height - row_start);
    }
    BBox tile_bounds(TileIndex idx) const
    {
        int c0, r0, cc, rc;
        tile_cell_range(idx, c0, r0, cc, rc);
        BBox b;
        b.min_x = bounds.min_x + c0 * cell_size_x;
        b.max_x = bounds.min_x + (c0 + cc) * cell_size_x;
        b.max_y = bounds.max_y + r0 * cell_size_y;
        b.min_y = bounds.max_y + (r0 + rc) * cell_size_y;
        return b;
    }
    int     total_tiles() const { return tiles_x * tiles_y; }
    int64_t total_cells() const { return static_cast<int64_t>(width) * height; }
    void gdal_geotransform(double gt[6]) const
    {
        gt[0] = bounds.min_x; gt[1] = cell_size_x; gt[2] = 0.0;
        gt[3] = bounds.max_y; gt[4] = 0.0;         gt[5] = cell_size_y;
    }
    Status validate() const
    {
        auto bad = [](const char* m) { return Status::error(StatusCode::InvalidArgument, m); };
        if (!bounds.valid()) return bad("Invalid bounds: max < min");
        if (cell_size_x == 0.0 || cell_size_y == 0.0) return bad("Cell size cannot be zero");
        if (tile_width <= 0 || tile_height <= 0) return bad("Tile dimensions must be positive");
        if (width <= 0 || height <= 0) return bad("Grid dimensions not computed or invalid. Call compute_dimensions()");
        if (!crs.is_valid()) return bad("CRS is not valid");
        return Status::success();
    }
};

// ---- PointCloud: SoA storage in host, pinned-host or device memory -----------------------------
struct ChannelDesc {
    std::string name;
    DataType    dtype = DataType::Float32;
    size_t      offset = 0;
};

class PointCloud {
public:
    ~PointCloud()
    {
        release(x_);
        release(y_);
        for (auto& c : channels_) release(c.data);
    }
    PointCloud(const PointCloud&) = delete;
    PointCloud& operator=(const PointCloud&) = delete;

    // `device` (new, defaulted) selects the GPU for Device / HostPinned storage.
    static std::unique_ptr<PointCloud> create(size_t capacity, MemoryLocation loc = MemoryLocation::Host,
                                              int device = 0)
    {
        std::unique_ptr<PointCloud> pc(new PointCloud(capacity, loc, device));
        if (capacity > 0 && (!pc->alloc(pc->x_, 8) || !pc->alloc(pc->y_, 8))) return nullptr;
        return pc;
    }

    Status add_channel(const std::string& name, DataType dtype = DataType::Float32)
    {
        if (has_channel(name)) return Status::error(StatusCode::InvalidArgument, "channel already exists: " + name);
        Channel c;
        c.desc.name = name; c.desc.dtype = dtype;
        c.elem = data_type_size(dtype);
        if (capacity_ > 0 && !alloc(c.data, c.elem))
            return Status::error(StatusCode::OutOfMemory, "failed to allocate channel: " + name);
        index_[name] = channels_.size();
        channels_.push_back(std::move(c));
        return Status::success();
    }
    bool has_channel(const std::string& name) const { return index_.count(name) != 0; }
    const ChannelDesc* channel(const std::string& name) const
    {
        auto it = index_.find(name);
        return it == index_.end() ? nullptr : &channels_[it->second].desc;
    }
    std::vector<std::string> channel_names() const
    {
        std::vector<std::string> out;
        for (auto& c : channels_) out.push_back(c.desc.name);
        return out;
    }

    double*       x()       { return static_cast<double*>(x_); }
    const double* x() const { return static_cast<const double*>(x_); }
    double*       y()       { return static_cast<double*>(y_); }
    const double* y() const { return static_cast<const double*>(y_); }
    void* channel_data(const std::string& name)
    {
        auto it = index_.find(name);
        return it == index_.end() ? nullptr : channels_[it->second].data;
    }
    const void* channel_data(const std::string& name) const { return const_cast<PointCloud*>(this)->channel_data(name); }
    float* channel_f32(const std::string& name)
    {
        const ChannelDesc* d = channel(name);
        return d && d->dtype == DataType::Float32 ? static_cast<float*>(channel_data(name)) : nullptr;
    }
    const float* channel_f32(const std::string& name) const { return const_cast<PointCloud*>(this)->channel_f32(name); }

    size_t         count()    const { return count_; }
    size_t         capacity() const { return capacity_; }
    MemoryLocation location() const { return loc_; }
    int            device()   const { return device_; }
    CRS            crs()      const { return crs_; }
    void           set_crs(const CRS& c) { crs_ = c; }

    Status resize(size_t n)
    {
        if (n > capacity_) return Status::error(StatusCode::InvalidArgument, "resize: new_count exceeds capacity");
        count_ = n;
        return Status::success();
    }

    // Copy to another memory space (device = -1 keeps this cloud's device).
    std::unique_ptr<PointCloud> to(MemoryLocation dst, int device = -1) const
    {
        const int dev = device < 0 ? device_ : device;
        auto out = create(capacity_, dst, dev);
        if (!out) return nullptr;
        out->count_ = count_;
        out->crs_ = crs_;
        auto copy = [&](void* d, const void* s, size_t elem) {
            return count_ == 0 ||
                   pcr_mem_copy(d, static_cast<int>(dst), s, static_cast<int>(loc_), count_ * elem, dev) == PCR_OK;
        };
        if (!copy(out->x_, x_, 8) || !copy(out->y_, y_, 8)) return nullptr;
        for (auto& c : channels_) {
            if (!out->add_channel(c.desc.name, c.desc.dtype).ok()) return nullptr;
            if (!copy(out->channels_.back().data, c.data, c.elem)) return nullptr;
        }
        return out;
    }

private:
    struct Channel { ChannelDesc desc; void* data = nullptr; size_t elem = 4; };

    PointCloud(size_t capacity, MemoryLocation loc, int device) : capacity_(capacity), loc_(loc), device_(device) {}
    bool alloc(void*& p, size_t elem)
    {
        return pcr_mem_alloc(static_cast<int>(loc_), device_, capacity_ * elem, &p) == PCR_OK;
    }
    void release(void*& p)
    {
        if (p) pcr_mem_free(static_cast<int>(loc_), device_, p);
        p = nullptr;
    }

    size_t capacity_ = 0, count_ = 0;
    MemoryLocation loc_ = MemoryLocation::Host;
    int device_ = 0;
    CRS crs_;
    void* x_ = nullptr;
    void* y_ = nullptr;
    std::vector<Channel> channels_;
    std::unordered_map<std::string, size_t> index_;
};

// ---- Grid: the finalized raster (host, band-sequential, row-major) -----------------------
struct BandDesc {
    std::string name;
    DataType    dtype = DataType::Float32;
    bool        is_state = false;
};

class Grid {
public:
    static std::unique_ptr<Grid> create(int cols, int rows, const std::vector<BandDesc>& bands,
                                        MemoryLocation loc = MemoryLocation::Host)
    {
        if (cols <= 0 || rows <= 0 || bands.empty() || loc == MemoryLocation::Device) return nullptr;
        std::unique_ptr<Grid> g(new Grid(cols, rows, bands));
        g->owned_.resize(bands.size());
        for (size_t i = 0; i < bands.size(); ++i) {
            g->owned_[i].assign(static_cast<size_t>(cols) * rows, 0.0f);
            g->data_[i] = g->owned_[i].data();
        }
        return g;
    }
    int      num_bands() const { return static_cast<int>(bands_.size()); }
    BandDesc band_desc(int i) const { return bands_.at(static_cast<size_t>(i)); }
    int      band_index(const std::string& name) const
    {
        for (size_t i = 0; i < bands_.size(); ++i) if (bands_[i].name == name) return static_cast<int>(i);
        return -1;
    }
    float* band_f32(int i) { return i >= 0 && i < num_bands() ? data_[static_cast<size_t>(i)] : nullptr; }
    const float* band_f32(int i) const { return const_cast<Grid*>(this)->band_f32(i); }
    float* band_f32(const std::string& name) { return band_f32(band_index(name)); }
    const float* band_f32(const std::string& name) const { return band_f32(band_index(name)); }
    int     cols() const { return cols_; }
    int     rows() const { return rows_; }
    int64_t cell_count() const { return static_cast<int64_t>(cols_) * rows_; }
    MemoryLocation location() const { return MemoryLocation::Host; }
    Status fill(float v)
    {
        for (int i = 0; i < num_bands(); ++i) fill_band(i, v);
        return Status::success();
    }
    Status fill_band(int i, float v)
    {
        float* p = band_f32(i);
        if (!p) return Status::error(StatusCode::InvalidArgument, "band index out of range");
        std::fill(p, p + cell_count(), v);
        return Status::success();
    }
    std::vector<uint8_t> valid_mask(int band = 0) const
    {
        std::vector<uint8_t> m(static_cast<size_t>(cell_count()), 0);
        if (const float* p = band_f32(band))
            for (size_t i = 0; i < m.size(); ++i) m[i] = p[i] == p[i];
        return m;
    }

private:
    friend class Pipeline;
    Grid(int cols, int rows, std::vector<BandDesc> bands)
        : cols_(cols), rows_(rows), bands_(std::move(bands)), data_(bands_.size(), nullptr) {}

    int cols_, rows_;
    std::vector<BandDesc> bands_;
    std::vector<float*> data_;                  // views: the pipeline's pinned result, or owned_
    std::vector<std::vector<float>> owned_;
};

// ---- GeoTIFF (GDAL-free writer in the library) ------------------------------------------
struct GeoTiffOptions {
    bool        cloud_optimized = false;
    std::string compress = "LZW";        // NONE, LZW, DEFLATE (ZSTD: NotImplemented)
    int         compress_level = 6;
    int         tile_width = 256, tile_height = 256;
    bool        bigtiff = true;
    std::string overview_resampling = "AVERAGE";
};

inline Status write_geotiff(const std::string& path, const Grid& grid, const GridConfig& config,
                            const GeoTiffOptions& o = GeoTiffOptions())
{
    if (grid.cols() != config.width || grid.rows() != config.height)
        return Status::error(StatusCode::InvalidArgument, "grid dimensions mismatch config");
    std::vector<const float*> bands;
    std::vector<std::string> names;
    std::vector<const char*> cnames;
    for (int i = 0; i < grid.num_bands(); ++i) { bands.push_back(grid.band_f32(i)); names.push_back(grid.band_desc(i).name); }
    for (auto& n : names) cnames.push_back(n.c_str());
    const pcr_grid_desc d = config.desc();
    const int rc = pcr_geotiff_write(path.c_str(), bands.data(), grid.num_bands(), &d, cnames.data(),
                                     config.crs.epsg, o.compress.c_str(), o.compress_level, o.tile_width,
                                     o.tile_height, o.bigtiff ? 1 : 0, o.cloud_optimized ? 1 : 0);
    if (rc == PCR_OK) return Status::success();
    const char* m = pcr_geotiff_last_error();
    return Status::error(static_cast<StatusCode>(rc), m ? m : "GeoTIFF error");
}

inline Status read_geotiff_info(const std::string& path, int& width, int& height, int& num_bands, CRS& crs, BBox& bounds)
{
    int32_t w = 0, h = 0, nb = 0, epsg = 0;
    double b[4] = { 0, 0, 0, 0 };
    const int rc = pcr_geotiff_read_info(path.c_str(), &w, &h, &nb, &epsg, b);
    if (rc != PCR_OK) {
        const char* m = pcr_geotiff_last_error();
        return Status::error(static_cast<StatusCode>(rc), m ? m : "GeoTIFF error");
    }
    width = w; height = h; num_bands = nb;
    crs = epsg ? CRS::from_epsg(epsg) : CRS();
    bounds = BBox(b[0], b[1], b[2], b[3]);
    return Status::success();
}

inline Status read_geotiff_band(const std::string& path, int band_index, float* data, int width, int height)
{
    const int rc = pcr_geotiff_read_band(path.c_str(), band_index, data, width, height);
    if (rc == PCR_OK) return Status::success();
    const char* m = pcr_geotiff_last_error();
    return Status::error(static_cast<StatusCode>(rc), m ? m : "GeoTIFF error");
}

// include/pcr/io/grid_io.h:44-70: tiles arrive one at a time; the file (and its overviews) is written at close()
class TiledGeoTiffWriter {
public:
    ~TiledGeoTiffWriter() { if (h_) pcr_geotiff_tiled_close(h_); }
    TiledGeoTiffWriter(const TiledGeoTiffWriter&) = delete;
    TiledGeoTiffWriter& operator=(const TiledGeoTiffWriter&) = delete;

    static std::unique_ptr<TiledGeoTiffWriter> open(const std::string& path, const GridConfig& config,
                                                    const std::vector<std::string>& band_names,
                                                    const GeoTiffOptions& o = GeoTiffOptions())
    {
        std::vector<const char*> names;
        for (const auto& n : band_names) names.push_back(n.c_str());
        const pcr_grid_desc d = config.desc();
        void* h = nullptr;
        const int rc = pcr_geotiff_tiled_open(path.c_str(), &d, names.data(), static_cast<int32_t>(names.size()),
                                              config.crs.epsg, o.compress.c_str(), o.compress_level, o.tile_width,
                                              o.tile_height, o.bigtiff ? 1 : 0, o.cloud_optimized ? 1 : 0, &h);
        if (rc != PCR_OK || !h) return nullptr;
        return std::unique_ptr<TiledGeoTiffWriter>(new TiledGeoTiffWriter(h));
    }

    Status write_tile(TileIndex tile, const float* data, int num_bands)
    {
        if (!h_) return Status::error(StatusCode::InvalidArgument, "writer not open");
        const int rc = pcr_geotiff_tiled_write_tile(h_, tile.row, tile.col, data, num_bands);
        if (rc == PCR_OK) return Status::success();
        const char* m = pcr_geotiff_last_error();
        return Status::error(static_cast<StatusCode>(rc), m ? m : "GeoTIFF error");
    }

    Status close()
    {
        if (!h_) return Status::error(StatusCode::InvalidArgument, "writer not open");
        void* h = h_;
        h_ = nullptr;
        const int rc = pcr_geotiff_tiled_close(h);
        if (rc == PCR_OK) return Status::success();
        const char* m = pcr_geotiff_last_error();
        return Status::error(static_cast<StatusCode>(rc), m ? m : "GeoTIFF error");
    }

private:
    explicit TiledGeoTiffWriter(void* h) : h_(h) {}
    void* h_ = nullptr;
};

// ---- specs ----------------------------------------------------------------------------------
struct FilterPredicate {
    std::string        channel_name;
    CompareOp          op = CompareOp::Equal;
    float              value = 0.0f;
    std::vector<float> value_set;
};

struct FilterSpec {
    std::vector<FilterPredicate> predicates;
    FilterSpec& add(const std::string& channel, CompareOp op, float value)
    {
        predicates.push_back({ channel, op, value, {} });
        return *this;
    }
    FilterSpec& add_in_set(const std::string& channel, const std::vector<float>& values)
    {
        predicates.push_back({ channel, CompareOp::InSet, 0.0f, values });
        return *this;
    }
    bool empty() const { return predicates.empty(); }
};

struct GlyphSpec {
    GlyphType   type = GlyphType::Point;
    std::string direction_channel;    float default_direction   = 0.0f;
    std::string half_length_channel;  float default_half_length = 1.0f;
    std::string sigma_x_channel;      float default_sigma_x     = 1.0f;
    std::string sigma_y_channel;      float default_sigma_y     = 1.0f;
    std::string rotation_channel;     float default_rotation    = 0.0f;
    float max_radius_cells = 32.0f;
    bool  normalize_weights = false;
};

struct ReductionSpec {
    std::string   value_channel;
    ReductionType type = ReductionType::Sum;
    std::string   weight_channel;       // carried, unused: WeightedAverage(Point) == Average upstream
    std::string   timestamp_channel;
    float         percentile = 0.5f;
    std::string   output_band_name;
    GlyphSpec     glyph;
};

struct PipelineConfig {
    GridConfig                 grid;
    std::vector<ReductionSpec> reductions;
    FilterSpec                 filter;
    CRS                        target_crs;
    bool                       auto_reproject = true;
    ExecutionMode              exec_mode = ExecutionMode::Auto;

    // Accepted for source compatibility, not used: state lives in HBM and chunking is the ingest ring's.
    size_t gpu_memory_budget = 0, host_cache_budget = 0, chunk_size = 0;
    size_t gpu_pool_size_bytes = 512u * 1024 * 1024;
    bool   use_cuda_streams = true;
    bool   gpu_require_strict = false;
    size_t cpu_threads = 0, hybrid_cpu_threads = 0;

    int  cuda_device_id = 0;
    bool gpu_fallback_to_cpu = true;    // never honoured: no device => create() fails

    std::string state_dir;              // read only when resume == true (explicit .pcrt reload)
    bool        resume = false;
    std::string output_path;            // GeoTIFF written by finalize() when non-empty
    bool        write_cog = false;

    // Additive knobs of the B200 path (pcr_pipeline_desc, pcr_b200.h).
    int      deterministic = 0;         // 0 off, 1 sort + in-order reduce, 2 exact fixed-point (order/sharding independent)
    int      ring_depth = 0;
    uint64_t ring_slot_points = 0;
    int      staging_threads = 0;
    int      point_kernel = 0, warp_aggregate = 0, gaussian_kernel = 0;
    int      comm_mode = 0;
    int      comm_root_only = 0;        // 0 = bands complete on every rank, 1 = on rank 0 only, 2 = distributed
    bool     async_ingest = false;
    int      comm_band_copy = 0;
    int      bin_cells_log2 = 0;        // tile binning (point_kernel = 3, or auto on large grids)
    uint64_t bin_pool_points = 0;
    int      comm_layout = 0;           // N>1: 0 auto, 1 replicated partial grids, 2 tile-partitioned grid
};

struct ProgressInfo {
    size_t collections_processed = 0;
    size_t collections_total = 0;
    size_t points_processed = 0;
    size_t tiles_active = 0;
    float  elapsed_seconds = 0.0f;
};

using ProgressCallback = std::function<bool(const ProgressInfo&)>;   // return false to cancel

// ---- Pipeline -----------------------------------------------------------------------------
class Pipeline {
public:
    ~Pipeline() { pcr_pipeline_destroy(h_); }
    Pipeline(const Pipeline&) = delete;
    Pipeline& operator=(const Pipeline&) = delete;

    // nullptr on failure with the reason on stderr (reference: src/engine/pipeline.cpp:1294-1304).
    static std::unique_ptr<Pipeline> create(const PipelineConfig& config)
    {
        std::unique_ptr<Pipeline> p(new Pipeline(config));
        const PipelineConfig& c = p->cfg_;              // the desc points into this copy's strings
        std::vector<pcr_reduction_desc> reds(c.reductions.size());
        for (size_t i = 0; i < reds.size(); ++i) {
            const ReductionSpec& r = c.reductions[i];
            const GlyphSpec& g = r.glyph;
            reds[i].value_channel = r.value_channel.c_str();
            reds[i].type = static_cast<int32_t>(r.type);
            reds[i].output_band_name = r.output_band_name.c_str();
            reds[i].glyph = { static_cast<int32_t>(g.type),
                              g.direction_channel.c_str(),   g.default_direction,
                              g.half_length_channel.c_str(), g.default_half_length,
                              g.sigma_x_channel.c_str(),     g.default_sigma_x,
                              g.sigma_y_channel.c_str(),     g.default_sigma_y,
                              g.rotation_channel.c_str(),    g.default_rotation,
                              g.max_radius_cells, g.normalize_weights ? 1 : 0 };
        }
        std::vector<pcr_filter_predicate> preds(c.filter.predicates.size());
        for (size_t i = 0; i < preds.size(); ++i) {
            const FilterPredicate& f = c.filter.predicates[i];
            preds[i] = { f.channel_name.c_str(), static_cast<int32_t>(f.op), f.value, f.value_set.data(),
                         static_cast<int32_t>(f.value_set.size()) };
        }
        pcr_pipeline_desc d{};
        d.grid = c.grid.desc();
        d.reductions = reds.data();
        d.num_reductions = static_cast<int32_t>(reds.size());
        d.exec_mode = static_cast<int32_t>(c.exec_mode);
        d.gpu_fallback_to_cpu = c.gpu_fallback_to_cpu ? 1 : 0;
        d.cuda_device_id = c.cuda_device_id;
        d.deterministic = c.deterministic;
        d.ring_depth = c.ring_depth;
        d.ring_slot_points = c.ring_slot_points;
        d.staging_threads = c.staging_threads;
        d.point_kernel = c.point_kernel;
        d.warp_aggregate = c.warp_aggregate;
        d.gaussian_kernel = c.gaussian_kernel;
        d.comm_mode = c.comm_mode;
        d.comm_root_only = c.comm_root_only;
        d.filter = preds.empty() ? nullptr : preds.data();
        d.num_predicates = static_cast<int32_t>(preds.size());
        d.async_ingest = c.async_ingest ? 1 : 0;
        d.comm_band_copy = c.comm_band_copy;
        d.bin_cells_log2 = c.bin_cells_log2;
        d.bin_pool_points = c.bin_pool_points;
        d.comm_layout = c.comm_layout;
        if (pcr_pipeline_create(&d, &p->h_) != PCR_OK || !p->h_) {
            const char* m = pcr_last_error();
            std::fprintf(stderr, "Pipeline::create failed: %s\n", m ? m : "");
            return nullptr;
        }
        // resume: continue from the tile-state files of an earlier run.  (Upstream never reads `resume`
        // and silently reloads ANY matching file in state_dir, tile_manager.cpp:272-302; here the flag is
        // explicit.)  A missing directory is a fresh start; an unreadable file is an error.
        if (c.resume && !c.state_dir.empty() && std::filesystem::is_directory(c.state_dir)) {
            const Status s = p->load_state(c.state_dir);
            if (!s.ok()) {
                std::fprintf(stderr, "Pipeline::create failed: %s\n", s.message.c_str());
                return nullptr;
            }
        }
        return p;
    }

    Status validate() const { return detail::from_rc(pcr_pipeline_validate(h_)); }

    // Points may be in Host, HostPinned or Device memory; may be called repeatedly.
    Status ingest(const PointCloud& cloud)
    {
        const std::vector<std::string> names = cloud.channel_names();
        std::vector<pcr_channel_view> views;
        views.reserve(names.size());
        for (const std::string& n : names)
            views.push_back({ n.c_str(), cloud.channel_data(n), static_cast<int32_t>(cloud.channel(n)->dtype) });
        return detail::from_rc(pcr_pipeline_ingest(h_, cloud.x(), cloud.y(), cloud.count(), views.data(),
                                                   static_cast<int32_t>(views.size()),
                                                   static_cast<int32_t>(cloud.location())));
    }

    Status finalize()
    {
        Status s = detail::from_rc(pcr_pipeline_finalize(h_));
        if (!s.ok()) return s;
        std::vector<BandDesc> bands(cfg_.reductions.size());
        for (size_t i = 0; i < bands.size(); ++i) {
            char name[512];
            s = detail::from_rc(pcr_pipeline_band_name(h_, static_cast<int32_t>(i), name, sizeof name));
            if (!s.ok()) return s;
            bands[i].name = name;
        }
        result_.reset(new Grid(cfg_.grid.width, cfg_.grid.height, bands));
        for (size_t i = 0; i < bands.size(); ++i) {         // zero-copy views of the pinned result
            const float* p = nullptr;
            int32_t rows = 0, cols = 0;
            s = detail::from_rc(pcr_pipeline_result_band(h_, static_cast<int32_t>(i), &p, &rows, &cols));
            if (!s.ok()) return s;
            result_->data_[i] = const_cast<float*>(p);
        }
        if (!cfg_.output_path.empty()) {
            GeoTiffOptions o;
            o.compress = cfg_.write_cog ? "DEFLATE" : "NONE";
            o.cloud_optimized = cfg_.write_cog;
            GridConfig gc = cfg_.grid;
            return write_geotiff(cfg_.output_path, *result_, gc, o);
        }
        return Status::success();
    }

    Status run(const std::vector<const PointCloud*>& clouds)
    {
        for (const PointCloud* c : clouds) {
            if (!c) return Status::error(StatusCode::InvalidArgument, "pipeline: null cloud pointer");
            Status s = ingest(*c);
            if (!s.ok()) return s;
        }
        return finalize();
    }

    void set_progress_callback(ProgressCallback cb)
    {
        cb_ = std::move(cb);
        if (cb_) pcr_pipeline_set_progress_callback(h_, &Pipeline::trampoline, this);
        else     pcr_pipeline_set_progress_callback(h_, nullptr, nullptr);
    }

    // Valid after finalize() until the next finalize()/reset(); nullptr before.
    const Grid* result() const { return result_.get(); }

    ProgressInfo stats() const
    {
        pcr_progress s{};
        pcr_pipeline_stats(h_, &s);
        return convert(s);
    }

    // ---- new-path extras (pcr_b200.h) ----
    Status finalize_device() { return detail::from_rc(pcr_pipeline_finalize_device(h_)); }
    Status result_band_device(int band, const float** data, int* rows, int* cols)
    {
        int32_t r = 0, c = 0;
        const Status s = detail::from_rc(pcr_pipeline_result_band_device(h_, band, data, &r, &c));
        if (rows) *rows = r;
        if (cols) *cols = c;
        return s;
    }
    Status reset() { result_.reset(); return detail::from_rc(pcr_pipeline_reset(h_)); }
    Status synchronize() { return detail::from_rc(pcr_pipeline_synchronize(h_)); }
    Status save_state(const std::string& dir) { return detail::from_rc(pcr_pipeline_save_state(h_, dir.c_str())); }
    Status load_state(const std::string& dir) { return detail::from_rc(pcr_pipeline_load_state(h_, dir.c_str())); }
    Status comm_init(const void* id128, int rank, int world_size)
    {
        return detail::from_rc(pcr_pipeline_comm_init(h_, id128, rank, world_size));
    }
    pcr_pipeline* handle() { return h_; }

private:
    explicit Pipeline(const PipelineConfig& c) : cfg_(c) {}
    static ProgressInfo convert(const pcr_progress& s)
    {
        ProgressInfo i;
        i.collections_processed = s.collections_processed;
        i.collections_total = s.collections_total;
        i.points_processed = s.points_processed;
        i.tiles_active = s.tiles_active;
        i.elapsed_seconds = s.elapsed_seconds;
        return i;
    }
    static int trampoline(const pcr_progress* info, void* user)
    {
        Pipeline* self = static_cast<Pipeline*>(user);
        return self->cb_ && !self->cb_(convert(*info)) ? 0 : 1;
    }

    pcr_pipeline*         h_ = nullptr;
    PipelineConfig        cfg_;
    std::unique_ptr<Grid> result_;
    ProgressCallback      cb_;
};

inline int         cuda_device_count() { return pcr_device_count(); }
inline std::string cuda_device_name(int device = 0)
{
    char buf[256] = { 0 };
    return pcr_device_name(device, buf, sizeof buf) == PCR_OK ? std::string(buf) : std::string();
}
inline bool cuda_get_memory_info(size_t* free_bytes, size_t* total_bytes, int device = 0)
{
    uint64_t f = 0, t = 0;
    if (pcr_device_mem_info(device, &f, &t) != PCR_OK) return false;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return true;
}

}  // namespace pcr

#endif  // PCR_B200_HPP
