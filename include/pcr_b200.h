/*
 * pcr_b200.h — C-ABI of libpcr_b200.so, the B200-native (sm_100a) replacement for
 * the ingest/finalize path of BigHippo123/pointcloud-raster ("PCR").
 *
 * This is the drop-in boundary: plain C, opaque handles, plain pointers and
 * sizes, int status codes.  No C++ types, no PyTorch types.  Each entry point
 * names the reference interface it replaces (paths relative to the reference
 * repository root).  INTEGRATION.md shows the binding a reference maintainer
 * would add on their side (pybind11 / ctypes).
 *
 * Status codes are the values of pcr::StatusCode (include/pcr/core/types.h:118-126).
 * On any non-zero return, pcr_last_error() holds the message the reference would
 * have put in Status::message (thread-local, valid until the next call on the
 * same thread).
 *
 * There is NO CPU fallback behind this ABI: exec_mode CPU is NotImplemented, and
 * a missing / unusable device is a CudaError regardless of gpu_fallback_to_cpu.
 */
#ifndef PCR_B200_H
#define PCR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums (numeric values identical to the reference's) ------------------ */

/* pcr::StatusCode, include/pcr/core/types.h:118-126 */
enum {
    PCR_OK = 0, PCR_INVALID_ARGUMENT = 1, PCR_OUT_OF_MEMORY = 2, PCR_CUDA_ERROR = 3,
    PCR_IO_ERROR = 4, PCR_CRS_ERROR = 5, PCR_NOT_IMPLEMENTED = 6
};
/* pcr::ReductionType, include/pcr/core/types.h:34-46 (only 0..5 are registered,
 * src/ops/reduction_registry.cpp:174-186) */
enum {
    PCR_SUM = 0, PCR_MAX = 1, PCR_MIN = 2, PCR_AVERAGE = 3, PCR_WEIGHTED_AVERAGE = 4,
    PCR_COUNT = 5, PCR_MEDIAN = 6, PCR_PERCENTILE = 7, PCR_MOST_RECENT = 8,
    PCR_PRIORITY_MERGE = 9, PCR_CUSTOM = 10
};
/* pcr::GlyphType, include/pcr/engine/glyph.h:10-14 */
enum { PCR_GLYPH_POINT = 0, PCR_GLYPH_LINE = 1, PCR_GLYPH_GAUSSIAN = 2 };
/* pcr::ExecutionMode, include/pcr/engine/pipeline.h:39-44 */
enum { PCR_EXEC_CPU = 0, PCR_EXEC_GPU = 1, PCR_EXEC_AUTO = 2, PCR_EXEC_HYBRID = 3 };
/* pcr::MemoryLocation, include/pcr/core/types.h:91-95 */
enum { PCR_MEM_HOST = 0, PCR_MEM_HOST_PINNED = 1, PCR_MEM_DEVICE = 2 };
/* pcr::DataType, include/pcr/core/types.h:18-26 */
enum { PCR_F32 = 0, PCR_F64 = 1, PCR_I32 = 2, PCR_U32 = 3, PCR_I16 = 4, PCR_U16 = 5, PCR_U8 = 6 };

/* ---- plain-data mirrors of the reference's config structs ----------------- */

/* Numeric fields of pcr::GridConfig (include/pcr/core/grid_config.h:17-44). */
typedef struct pcr_grid_desc {
    double  min_x, min_y, max_x, max_y;   /* GridConfig::bounds */
    double  cell_size_x, cell_size_y;     /* cell_size_y < 0 = north-up */
    int32_t width, height;                /* cells; see pcr_grid_compute_dimensions */
    int32_t tile_width, tile_height;      /* reference tiles (default 4096): they
                                             define the touched-tile NaN rule and the
                                             glyph clipping seams, nothing else */
} pcr_grid_desc;

/* pcr::GlyphSpec (include/pcr/engine/glyph.h:19-43).  NULL or "" channel = default. */
typedef struct pcr_glyph_desc {
    int32_t     type;
    const char *direction_channel;    float default_direction;
    const char *half_length_channel;  float default_half_length;
    const char *sigma_x_channel;      float default_sigma_x;
    const char *sigma_y_channel;      float default_sigma_y;
    const char *rotation_channel;     float default_rotation;
    float       max_radius_cells;
    int32_t     normalize_weights;    /* accepted and ignored, as upstream
                                         (src/engine/glyph_kernels.cu:169-179) */
} pcr_glyph_desc;

/* pcr::ReductionSpec (include/pcr/engine/pipeline.h:20-34).  weight_channel,
 * timestamp_channel and percentile are never read upstream and are not mirrored. */
typedef struct pcr_reduction_desc {
    const char    *value_channel;
    int32_t        type;
    const char    *output_band_name;  /* NULL/"" => "<value_channel>_<type int>" */
    pcr_glyph_desc glyph;
} pcr_reduction_desc;

/* pcr::FilterPredicate (include/pcr/engine/filter.h:33-38); CompareOp values of filter.h:20-29.
 * A point survives iff ALL predicates hold (evaluate_predicate, src/engine/filter.cpp:37-58). */
enum { PCR_CMP_EQUAL = 0, PCR_CMP_NOT_EQUAL = 1, PCR_CMP_LESS = 2, PCR_CMP_LESS_EQUAL = 3,
       PCR_CMP_GREATER = 4, PCR_CMP_GREATER_EQUAL = 5, PCR_CMP_IN_SET = 6, PCR_CMP_NOT_IN_SET = 7 };
typedef struct pcr_filter_predicate {
    const char  *channel_name;     /* Float32 channel of the cloud */
    int32_t      op;
    float        value;            /* scalar comparisons */
    const float *value_set;        /* InSet / NotInSet */
    int32_t      value_set_size;
} pcr_filter_predicate;

/* pcr::PipelineConfig (include/pcr/engine/pipeline.h:49-86), hot-path fields,
 * plus the additive knobs of the new path (all default 0). */
typedef struct pcr_pipeline_desc {
    pcr_grid_desc             grid;
    const pcr_reduction_desc *reductions;
    int32_t                   num_reductions;
    int32_t                   exec_mode;            /* CPU => NotImplemented; GPU/Auto/Hybrid => GPU */
    int32_t                   gpu_fallback_to_cpu;  /* never honoured: no device => CudaError */
    int32_t                   cuda_device_id;
    /* --- additive knobs --- */
    int32_t                   deterministic;        /* 1 = sort by cell + in-order segmented reduce: bit-reproducible
                                                       run to run for a fixed chunking and sharding (Point glyph,
                                                       Gaussian gather); 2 = exact fixed-point accumulation of every
                                                       float contribution, rounded once at finalize: bit-identical
                                                       whatever the point order, the ingest chunking and the number
                                                       of GPUs (all glyphs; ranks combine by integer all-reduce) */
    int32_t                   ring_depth;           /* host-ingest staging slots, 0 = default (3) */
    uint64_t                  ring_slot_points;     /* points per ring chunk, 0 = default (256 Ki staged
                                                       from pageable memory, 2 Mi direct from pinned) */
    int32_t                   staging_threads;      /* host copy threads, 0 = default */
    int32_t                   point_kernel;         /* 0 = auto (direct; tile-binned when the records exceed 256 MB),
                                                       1 = direct LDG, 2 = TMA-staged persistent, 3 = tile-binned:
                                                       entries {cell, value} are appended to per-bin page chains at
                                                       ingest and folded bin by bin at finalize */
    int32_t                   warp_aggregate;       /* 0 = auto (adaptive run aggregation), 2 = off */
    int32_t                   gaussian_kernel;      /* 0 = auto, 1 = scatter (warp per point, REDs),
                                                       2 = gather (tile-binned, atomic-free, deterministic),
                                                       3 = per-bin GEMM (unrotated footprints, radius cap <= 32
                                                       cells; what auto picks there unless deterministic) */
    int32_t                   comm_mode;            /* N>1 combine: 0 = auto (peer memory over NVLink when every
                                                       GPU pair has P2P access, else NCCL), 1 = NCCL, 2 = peer */
    int32_t                   comm_root_only;       /* N>1, where the finalized bands end up: 0 = complete on every
                                                       rank, 1 = complete on rank 0 only, 2 = distributed (every
                                                       rank holds just the row slice it owns, pcr_comm_slice_rows) */
    const pcr_filter_predicate *filter;             /* PipelineConfig::filter (N3): evaluated on the device,
                                                       fused in front of routing; NULL/0 = no filter */
    int32_t                   num_predicates;
    int32_t                   async_ingest;         /* 1 = device-resident ingests return without a
                                                       stream sync; buffers must stay valid until the
                                                       next finalize / synchronize */
    int32_t                   comm_band_copy;       /* N>1 peer mode, how a rank's finalized band slice reaches
                                                       the other ranks: 0/1 = stores from the merge kernel
                                                       (default), 2 = one copy-engine transfer per band and peer */
    int32_t                   bin_cells_log2;       /* tile binning: a bin = 2^k consecutive cells; 0 = auto
                                                       (the records of one bin <= 64 MB, at most 1024 bins) */
    uint64_t                  bin_pool_points;      /* tile binning: entries the pool holds before it is folded
                                                       early; 0 = auto (a quarter of the free HBM, <= 2^31) */
    int32_t                   comm_layout;          /* N>1: 0 = auto, 1 = replicated partial grids merged at finalize,
                                                       2 = tile-partitioned grid: every rank owns a contiguous range
                                                       of bins, points are exchanged (all-to-all over NVLink peer
                                                       memory, fused into the binning kernel), no reduce at finalize */
} pcr_pipeline_desc;

/* One named channel of a point cloud (pcr::PointCloud, include/pcr/core/point_cloud.h:29-103). */
typedef struct pcr_channel_view {
    const char *name;
    const void *data;     /* count elements; same memory location as x/y */
    int32_t     dtype;    /* PCR_F32 is the only dtype the pipeline reduces
                             (src/engine/pipeline.cpp:372-378) */
} pcr_channel_view;

/* pcr::ProgressInfo (include/pcr/engine/pipeline.h:91-99). */
typedef struct pcr_progress {
    uint64_t collections_processed;
    uint64_t collections_total;    /* always 0 (streaming), as upstream */
    uint64_t points_processed;     /* counts every ingested point, in-grid or not
                                      (src/engine/pipeline.cpp:749) */
    uint64_t tiles_active;         /* reference tiles touched so far */
    float    elapsed_seconds;
} pcr_progress;

/* Device-side timings, accumulated since the last pcr_pipeline_profile_reset
 * while profiling is enabled (CUDA events on the launching stream). */
typedef struct pcr_profile {
    double   accumulate_ms;     uint64_t accumulate_launches;  /* route+accumulate / glyph kernels */
    double   sort_ms;           uint64_t sort_launches;        /* deterministic mode: key build + radix sort */
    double   finalize_ms;       uint64_t finalize_launches;    /* finalize (+merge) kernels */
    double   init_ms;           uint64_t init_launches;        /* state identity fill */
    uint64_t h2d_bytes;         uint64_t d2h_bytes;            /* bytes moved by ingest ring / finalize */
    uint64_t points;                                           /* points fed to accumulate kernels */
    uint64_t kernel_launches;                                  /* every kernel of this library launched */
    double   push_ms;           uint64_t push_launches;        /* N>1, peer mode: slice push over NVLink */
} pcr_profile;

typedef struct pcr_pipeline pcr_pipeline;
typedef int (*pcr_progress_fn)(const pcr_progress *info, void *user);  /* return 0 to cancel */

/* ---- errors / devices ------------------------------------------------------ */
const char *pcr_last_error(void);
int  pcr_device_count(void);                                  /* cuda_device_count, types.h:157-169 */
int  pcr_device_name(int device, char *buf, size_t buflen);   /* cuda_device_name, types.h:171-184 */
int  pcr_device_mem_info(int device, uint64_t *free_bytes, uint64_t *total_bytes); /* types.h:186-201 */
const char *pcr_version(void);

/* ---- GridConfig host logic ------------------------------------------------- */
/* GridConfig::compute_dimensions, src/core/grid_config.cpp:7-22 */
int pcr_grid_compute_dimensions(pcr_grid_desc *grid);
/* GridConfig::world_to_cell, src/core/grid_config.cpp:24-43; returns 1 if inside */
int pcr_grid_world_to_cell(const pcr_grid_desc *grid, double wx, double wy,
                           int32_t *col, int32_t *row);

/* ---- memory for PointCloud buffers (PointCloud::create/to, point_cloud.cpp:37-90,404-512) */
int pcr_mem_alloc(int location, int device, size_t bytes, void **out);
int pcr_mem_free(int location, int device, void *ptr);
int pcr_mem_copy(void *dst, int dst_location, const void *src, int src_location,
                 size_t bytes, int device);

/* ---- Pipeline (include/pcr/engine/pipeline.h:105-145) ---------------------- */
/* Pipeline::create + Impl::initialize, src/engine/pipeline.cpp:92-281,1294-1304.
 * On failure *out is NULL (the reference returns nullptr) and the code/message say why. */
int pcr_pipeline_create(const pcr_pipeline_desc *desc, pcr_pipeline **out);
void pcr_pipeline_destroy(pcr_pipeline *p);
/* Pipeline::validate, src/engine/pipeline.cpp:1306-1338 */
int pcr_pipeline_validate(const pcr_pipeline *p);
/* Pipeline::ingest -> Impl::process_cloud, src/engine/pipeline.cpp:283-770,1340.
 * Borrowed pointers; `location` says where x/y/channels live (host pageable, host
 * pinned or device).  Returns once the caller may free or overwrite the buffers.
 * PCR_MEM_DEVICE buffers are read by kernels on the pipeline's own (non-blocking) stream: they must be
 * COMPLETE before the call — synchronize the stream that produced them (the legacy default stream
 * included); there is no implicit ordering against the caller's streams. */
int pcr_pipeline_ingest(pcr_pipeline *p, const double *x, const double *y, size_t count,
                        const pcr_channel_view *channels, int32_t num_channels,
                        int32_t location);
/* Pipeline::finalize -> Impl::finalize_result, src/engine/pipeline.cpp:1154-1286,1344.
 * Finalizes on the device and copies every band to host memory owned by the
 * pipeline.  May be called repeatedly; later ingests keep accumulating. */
int pcr_pipeline_finalize(pcr_pipeline *p);
/* Same, but leaves the finalized bands in HBM only (no D2H).  With async_ingest = 1 it
 * returns without synchronizing (stream-ordered); call pcr_pipeline_synchronize. */
int pcr_pipeline_finalize_device(pcr_pipeline *p);
/* Pipeline::result()->band_f32(i), include/pcr/core/grid.h:62-66: row-major
 * rows x cols float32, valid until the next finalize/destroy. */
int pcr_pipeline_result_band(pcr_pipeline *p, int32_t band, const float **data,
                             int32_t *rows, int32_t *cols);
int pcr_pipeline_result_band_device(pcr_pipeline *p, int32_t band, const float **device_data,
                                    int32_t *rows, int32_t *cols);
/* Band name: output_band_name or "<value_channel>_<type>", pipeline.cpp:1178-1180 */
int pcr_pipeline_band_name(const pcr_pipeline *p, int32_t band, char *buf, size_t buflen);
/* Pipeline::stats, src/engine/pipeline.cpp:1388-1401 */
int pcr_pipeline_stats(const pcr_pipeline *p, pcr_progress *out);
/* Pipeline::set_progress_callback, src/engine/pipeline.cpp:1380-1382.  Fires
 * synchronously on the caller's thread after each ingest; returning 0 makes that
 * ingest fail with "pipeline: cancelled by user" (pipeline.cpp:753-767). */
int pcr_pipeline_set_progress_callback(pcr_pipeline *p, pcr_progress_fn fn, void *user);
/* Drop all accumulated state (what deleting state_dir + re-creating does upstream). */
int pcr_pipeline_reset(pcr_pipeline *p);
/* Block until all device work of this pipeline is complete. */
int pcr_pipeline_synchronize(pcr_pipeline *p);

/* ---- tile-state checkpoints (.pcrt), SURVEY §8f N2 ------------------------------ */
/* File format of the reference (src/io/tile_state_io.cpp:14-95): 36-byte packed header
 * {magic "PCRT", version 1, tile_row, tile_col, cols, rows, state_floats, reduction u8, 7 reserved}
 * + state_floats * cols * rows float32, band-sequential, one file per reference tile named
 * tile_RRRR_CCCC.pcrt (tile_state_filename, :197-211).  save writes one file per TOUCHED tile per
 * reduction — directly into `dir` for a single-reduction pipeline (the reference's layout, so the
 * files interoperate with the reference's TileManager), into `dir`/band_<k>/ otherwise.  load reads
 * whatever matching files exist (header dims must match, as tile_manager.cpp:272-302 checks), makes
 * them the accumulated state of their tiles and marks those tiles touched. */
int pcr_pipeline_save_state(pcr_pipeline *p, const char *dir);
int pcr_pipeline_load_state(pcr_pipeline *p, const char *dir);

/* ---- GeoTIFF out (GDAL-free; host only) --------------------------------------- */
/* write_geotiff, src/io/grid_io.cpp:39-182 (called by Pipeline::finalize when output_path is set,
 * src/engine/pipeline.cpp:1350-1361): Float32 tiled (Big)TIFF, one plane per band, nodata NaN, band
 * descriptions, geotransform of GridConfig::gdal_geotransform, EPSG GeoKeys.  compress: "NONE", "LZW" (the
 * reference's default, include/pcr/io/grid_io.h:18) or "DEFLATE".  cloud_optimized != 0 adds the overview
 * pyramid of grid_io.cpp:155-176 (levels 2, 4, ... while min(width, height) / level >= 256; NaN-aware AVERAGE).
 * bands[b] = rows*cols row-major floats (host). */
int pcr_geotiff_write(const char *path, const float *const *bands, int32_t num_bands,
                      const pcr_grid_desc *grid, const char *const *band_names, int32_t epsg,
                      const char *compress, int32_t compress_level, int32_t tile_width,
                      int32_t tile_height, int32_t bigtiff, int32_t cloud_optimized);
/* TiledGeoTiffWriter, include/pcr/io/grid_io.h:44-70, src/io/grid_io.cpp:185-380: open, write the finalized
 * data of one reference tile at a time (band-sequential, tile_cols x tile_rows floats per band, tile geometry =
 * GridConfig::tile_cell_range of grid->tile_width/height), close (writes the file and its overviews).
 * Tiles never written stay NaN. */
int pcr_geotiff_tiled_open(const char *path, const pcr_grid_desc *grid, const char *const *band_names,
                           int32_t num_bands, int32_t epsg, const char *compress, int32_t compress_level,
                           int32_t tile_width, int32_t tile_height, int32_t bigtiff, int32_t cloud_optimized,
                           void **handle);
int pcr_geotiff_tiled_write_tile(void *handle, int32_t tile_row, int32_t tile_col, const float *data,
                                 int32_t num_bands);
int pcr_geotiff_tiled_close(void *handle);
/* read_geotiff_info, src/io/grid_io.cpp:395-445; bounds = {min_x, min_y, max_x, max_y} */
int pcr_geotiff_read_info(const char *path, int32_t *width, int32_t *height, int32_t *num_bands,
                          int32_t *epsg, double bounds[4]);
/* read_geotiff_band, src/io/grid_io.cpp:445-497: one band of the full-resolution image into `data`
 * (width*height floats); files of this writer's layout (tiled, band-separate, NONE/LZW/DEFLATE). */
int pcr_geotiff_read_band(const char *path, int32_t band_index, float *data, int32_t width, int32_t height);
const char *pcr_geotiff_last_error(void);

/* ---- profiling (new; feeds bench.py's roofline block) ---------------------- */
/* Device-side stopwatch: begin records a CUDA event on the pipeline's compute stream
 * (the stream every kernel, the finalize D2H and the waits on the copy stream are
 * ordered on); end records a second one, synchronizes and returns the elapsed ms. */
int pcr_pipeline_timer_begin(pcr_pipeline *p);
int pcr_pipeline_timer_end(pcr_pipeline *p, double *elapsed_ms);
/* Per-kernel-group CUDA events.  on = 0: off; 1: every kernel group; N > 1: every Nth group of each kind
 * (accumulate, sort, push, finalize, init) — two timed events around a 60 us kernel cost ~5 us of stream
 * time and keep the next kernel from launching under its tail, so a hot loop samples.  The `*_launches`
 * fields of pcr_profile count the SAMPLED groups (mean = *_ms / *_launches); `kernel_launches` counts all. */
int pcr_pipeline_profile_enable(pcr_pipeline *p, int32_t on);
int pcr_pipeline_profile_reset(pcr_pipeline *p);
int pcr_pipeline_profile_read(pcr_pipeline *p, pcr_profile *out);

/* Measurement aid (not on the product path): median time, in microseconds over `reps` launches, of folding
 * `points` synthetic points with hashed cells into `cells` 16-byte records [sum, count, max, pad] with the
 * reduction instructions of the Point kernel (one red.global.add.v2.f32 + one red.global.max.s32 per point),
 * optionally behind the kernel's three streaming loads per point (x, y f64 + value f32).  This is the rate at
 * which the GPU's L2 resolves scattered reductions — the ceiling bench.py quotes next to the HBM roofline. */
int pcr_diag_red_ceiling(int device, uint64_t points, uint64_t cells, int with_loads, int reps,
                         double *median_us);

/* ---- multi-GPU: one process per GPU, point shards, combine at finalize ----- */
/* The partial grid states of all ranks are merged with Op::merge semantics
 * (include/pcr/ops/builtin_ops.h:15,28,41,54,67,95-97; src/engine/grid_merge.cu:26-35)
 * over NCCL: each rank owns a row slice, receives every peer's partial slice,
 * merges them in rank order and finalizes the slice in one kernel; slices are then
 * gathered so that every rank's result bands are complete.
 * pcr_comm_unique_id: rank 0 fills 128 bytes (ncclUniqueId) and ships them to the
 * other ranks by any side channel (bench.py uses torch.distributed). */
int pcr_comm_unique_id(void *id128);
/* Row slice [row0, row1) of the grid that `rank` owns (merges and finalizes) at an N-rank finalize. */
int pcr_comm_slice_rows(int32_t height, int32_t world_size, int32_t rank, int32_t *row0, int32_t *row1);
int pcr_pipeline_comm_init(pcr_pipeline *p, const void *id128, int32_t rank, int32_t world_size);
int pcr_pipeline_comm_barrier(pcr_pipeline *p);
/* Tile-partitioned layout (comm_layout = 2), host arithmetic only: the bin geometry the engine derives for a grid of
 * `cells` cells with `record_words`-word records (bin = 2^bin_shift consecutive cells; bin_cells_log2 = 0: records of a
 * bin <= 64 MB, at most 1024 bins) and the row-major cell range [cell0, cell1) of the bins `rank` owns
 * (ceil(num_bins / world_size) consecutive bins per rank). */
int pcr_comm_partition_cells(uint64_t cells, int32_t record_words, int32_t bin_cells_log2, int32_t world_size, int32_t rank,
                             int32_t *bin_shift, int32_t *num_bins, uint64_t *cell0, uint64_t *cell1);
/* Row-major cell range [cell0, cell1) whose finalized bands THIS rank produces: the whole grid on one GPU,
 * the rank's row slice with replicated partial grids, the cells of its bins with the tile-partitioned layout.
 * With comm_root_only = 2 a rank's band arrays are valid for exactly this range. */
int pcr_pipeline_owned_cells(const pcr_pipeline *p, uint64_t *cell0, uint64_t *cell1);

#ifdef __cplusplus
}
#endif
#endif /* PCR_B200_H */
