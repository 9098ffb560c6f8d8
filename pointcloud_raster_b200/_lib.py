"""ctypes loader for libpcr_b200.so and declarations of the C-ABI in include/pcr_b200.h.

No fallback of any kind: a missing library raises ImportError naming the build command.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpcr_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C pointcloud_raster_b200/csrc`. There is no CPU or PyTorch fallback.")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


class GridDesc(C.Structure):
    _fields_ = [("min_x", C.c_double), ("min_y", C.c_double), ("max_x", C.c_double),
                ("max_y", C.c_double), ("cell_size_x", C.c_double), ("cell_size_y", C.c_double),
                ("width", C.c_int32), ("height", C.c_int32),
                ("tile_width", C.c_int32), ("tile_height", C.c_int32)]


class GlyphDesc(C.Structure):
    _fields_ = [("type", C.c_int32),
                ("direction_channel", C.c_char_p), ("default_direction", C.c_float),
                ("half_length_channel", C.c_char_p), ("default_half_length", C.c_float),
                ("sigma_x_channel", C.c_char_p), ("default_sigma_x", C.c_float),
                ("sigma_y_channel", C.c_char_p), ("default_sigma_y", C.c_float),
                ("rotation_channel", C.c_char_p), ("default_rotation", C.c_float),
                ("max_radius_cells", C.c_float), ("normalize_weights", C.c_int32)]


class ReductionDesc(C.Structure):
    _fields_ = [("value_channel", C.c_char_p), ("type", C.c_int32),
                ("output_band_name", C.c_char_p), ("glyph", GlyphDesc)]


class FilterPredicate(C.Structure):
    _fields_ = [("channel_name", C.c_char_p), ("op", C.c_int32), ("value", C.c_float),
                ("value_set", C.POINTER(C.c_float)), ("value_set_size", C.c_int32)]


class PipelineDesc(C.Structure):
    _fields_ = [("grid", GridDesc), ("reductions", C.POINTER(ReductionDesc)),
                ("num_reductions", C.c_int32), ("exec_mode", C.c_int32),
                ("gpu_fallback_to_cpu", C.c_int32), ("cuda_device_id", C.c_int32),
                ("deterministic", C.c_int32), ("ring_depth", C.c_int32),
                ("ring_slot_points", C.c_uint64), ("staging_threads", C.c_int32),
                ("point_kernel", C.c_int32), ("warp_aggregate", C.c_int32),
                ("gaussian_kernel", C.c_int32), ("comm_mode", C.c_int32), ("comm_root_only", C.c_int32),
                ("filter", C.POINTER(FilterPredicate)), ("num_predicates", C.c_int32),
                ("async_ingest", C.c_int32), ("comm_band_copy", C.c_int32),
                ("bin_cells_log2", C.c_int32), ("bin_pool_points", C.c_uint64), ("comm_layout", C.c_int32)]


class ChannelView(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int32)]


class Progress(C.Structure):
    _fields_ = [("collections_processed", C.c_uint64), ("collections_total", C.c_uint64),
                ("points_processed", C.c_uint64), ("tiles_active", C.c_uint64),
                ("elapsed_seconds", C.c_float)]


class Profile(C.Structure):
    _fields_ = [("accumulate_ms", C.c_double), ("accumulate_launches", C.c_uint64),
                ("sort_ms", C.c_double), ("sort_launches", C.c_uint64),
                ("finalize_ms", C.c_double), ("finalize_launches", C.c_uint64),
                ("init_ms", C.c_double), ("init_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("points", C.c_uint64),
                ("kernel_launches", C.c_uint64),
                ("push_ms", C.c_double), ("push_launches", C.c_uint64)]


PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.POINTER(Progress), C.c_void_p)

# Every symbol include/pcr_b200.h declares: (name, restype, argtypes).
SYMBOLS = [
    ("pcr_last_error", C.c_char_p, []),
    ("pcr_version", C.c_char_p, []),
    ("pcr_device_count", C.c_int, []),
    ("pcr_device_name", C.c_int, [C.c_int, C.c_char_p, C.c_size_t]),
    ("pcr_device_mem_info", C.c_int, [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("pcr_grid_compute_dimensions", C.c_int, [C.POINTER(GridDesc)]),
    ("pcr_grid_world_to_cell", C.c_int, [C.POINTER(GridDesc), C.c_double, C.c_double,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("pcr_mem_alloc", C.c_int, [C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    ("pcr_mem_free", C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    ("pcr_mem_copy", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.c_int]),
    ("pcr_pipeline_create", C.c_int, [C.POINTER(PipelineDesc), C.POINTER(C.c_void_p)]),
    ("pcr_pipeline_destroy", None, [C.c_void_p]),
    ("pcr_pipeline_validate", C.c_int, [C.c_void_p]),
    ("pcr_pipeline_ingest", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.POINTER(ChannelView), C.c_int32, C.c_int32]),
    ("pcr_pipeline_finalize", C.c_int, [C.c_void_p]),
    ("pcr_pipeline_finalize_device", C.c_int, [C.c_void_p]),
    ("pcr_pipeline_result_band", C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("pcr_pipeline_result_band_device", C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p),
                                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("pcr_pipeline_band_name", C.c_int, [C.c_void_p, C.c_int32, C.c_char_p, C.c_size_t]),
    ("pcr_pipeline_stats", C.c_int, [C.c_void_p, C.POINTER(Progress)]),
    ("pcr_pipeline_set_progress_callback", C.c_int, [C.c_void_p, PROGRESS_FN, C.c_void_p]),
    ("pcr_pipeline_save_state", C.c_int, [C.c_void_p, C.c_char_p]),
    ("pcr_pipeline_load_state", C.c_int, [C.c_void_p, C.c_char_p]),
    ("pcr_pipeline_reset", C.c_int, [C.c_void_p]),
    ("pcr_pipeline_synchronize", C.c_int, [C.c_void_p]),
    ("pcr_geotiff_write", C.c_int, [C.c_char_p, C.POINTER(C.c_void_p), C.c_int32, C.POINTER(GridDesc),
                                    C.POINTER(C.c_char_p), C.c_int32, C.c_char_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32]),
    ("pcr_geotiff_tiled_open", C.c_int, [C.c_char_p, C.POINTER(GridDesc), C.POINTER(C.c_char_p), C.c_int32, C.c_int32,
                                         C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.POINTER(C.c_void_p)]),
    ("pcr_geotiff_tiled_write_tile", C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.c_int32]),
    ("pcr_geotiff_tiled_close", C.c_int, [C.c_void_p]),
    ("pcr_geotiff_read_band", C.c_int, [C.c_char_p, C.c_int32, C.POINTER(C.c_float), C.c_int32, C.c_int32]),
    ("pcr_geotiff_read_info", C.c_int, [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    ("pcr_geotiff_last_error", C.c_char_p, []),
    ("pcr_pipeline_profile_enable", C.c_int, [C.c_void_p, C.c_int32]),
    ("pcr_pipeline_profile_reset", C.c_int, [C.c_void_p]),
    ("pcr_pipeline_profile_read", C.c_int, [C.c_void_p, C.POINTER(Profile)]),
    ("pcr_pipeline_timer_begin", C.c_int, [C.c_void_p]),
    ("pcr_pipeline_timer_end", C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    ("pcr_diag_red_ceiling", C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    ("pcr_comm_unique_id", C.c_int, [C.c_void_p]),
    ("pcr_comm_slice_rows", C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("pcr_pipeline_comm_init", C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    ("pcr_pipeline_comm_barrier", C.c_int, [C.c_void_p]),
    ("pcr_comm_partition_cells", C.c_int, [C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("pcr_pipeline_owned_cells", C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
]

for _name, _res, _args in SYMBOLS:
    _f = getattr(lib, _name)          # AttributeError here = header/library mismatch
    _f.restype = _res
    _f.argtypes = _args


def last_error() -> str:
    return (lib.pcr_last_error() or b"").decode("utf-8", "replace")


def check(code: int) -> None:
    """check_status of the reference bindings (python/bindings.cpp:22-26):
    any non-OK status becomes RuntimeError(message)."""
    if code != 0:
        raise RuntimeError(last_error())
