"""pcr — the reference's Python API surface, served by the B200 C-ABI.

Mirrors what ``python/pcr/__init__.py:17-66,73-181`` of the reference exports and
what ``python/bindings.cpp`` binds (same class, field and method names, same
argument meaning, same error behaviour: a non-OK status raises
``RuntimeError(message)``, ``Pipeline.create`` returns ``None`` on failure).
Everything that computes goes through ``libpcr_b200.so`` (``include/pcr_b200.h``);
nothing here falls back to numpy or PyTorch.

Differences that are part of the contract of the new path:
  * ``ExecutionMode.CPU`` makes ``Pipeline.create`` fail (no CPU path);
    ``GPU``/``Auto``/``Hybrid`` all run the same GPU path.
  * ``gpu_fallback_to_cpu`` is accepted and never honoured.
  * additive ``PipelineConfig`` knobs: ``cuda_device_id``, ``deterministic``,
    ``ring_depth``, ``ring_slot_points``, ``staging_threads``, ``point_kernel``,
    ``warp_aggregate``, ``gaussian_kernel``, ``comm_mode``, ``comm_root_only``, ``async_ingest``.
"""
from __future__ import annotations

import ctypes as C
import enum
import sys
import weakref

import numpy as np

from .. import _lib
from .._lib import lib, check

__version__ = "0.1.0"


# ---------------------------------------------------------------------------
# Enums (values identical to include/pcr/core/types.h, glyph.h, pipeline.h, filter.h)
# ---------------------------------------------------------------------------
class DataType(enum.IntEnum):
    Float32 = 0
    Float64 = 1
    Int32 = 2
    UInt32 = 3
    Int16 = 4
    UInt16 = 5
    UInt8 = 6


_NP_DTYPE = {DataType.Float32: np.float32, DataType.Float64: np.float64, DataType.Int32: np.int32,
             DataType.UInt32: np.uint32, DataType.Int16: np.int16, DataType.UInt16: np.uint16,
             DataType.UInt8: np.uint8}


class ReductionType(enum.IntEnum):
    Sum = 0
    Max = 1
    Min = 2
    Average = 3
    WeightedAverage = 4
    Count = 5
    Median = 6
    Percentile = 7
    MostRecent = 8
    PriorityMerge = 9
    Custom = 10


class MemoryLocation(enum.IntEnum):
    Host = 0
    HostPinned = 1
    Device = 2


class ExecutionMode(enum.IntEnum):
    CPU = 0
    GPU = 1
    Auto = 2
    Hybrid = 3


class StatusCode(enum.IntEnum):
    Ok = 0
    InvalidArgument = 1
    OutOfMemory = 2
    CudaError = 3
    IoError = 4
    CrsError = 5
    NotImplemented = 6


class CompareOp(enum.IntEnum):
    Equal = 0
    NotEqual = 1
    Less = 2
    LessEqual = 3
    Greater = 4
    GreaterEqual = 5
    InSet = 6
    NotInSet = 7


class PointCloudFormat(enum.IntEnum):
    PCR_Binary = 0
    CSV = 1
    LAS = 2
    LAZ = 3
    Auto = 4


class GlyphType(enum.IntEnum):
    Point = 0
    Line = 1
    Gaussian = 2


# ---------------------------------------------------------------------------
# Core value types (bindings.cpp:106-187)
# ---------------------------------------------------------------------------
_DBL_MAX = sys.float_info.max


class BBox:
    def __init__(self):
        self.min_x = _DBL_MAX
        self.min_y = _DBL_MAX
        self.max_x = -_DBL_MAX
        self.max_y = -_DBL_MAX

    def expand(self, *args):
        if len(args) == 1:
            o = args[0]
            if not o.valid():
                return
            self.expand(o.min_x, o.min_y)
            self.expand(o.max_x, o.max_y)
            return
        x, y = args
        self.min_x = min(self.min_x, x)
        self.min_y = min(self.min_y, y)
        self.max_x = max(self.max_x, x)
        self.max_y = max(self.max_y, y)

    def contains(self, x, y):
        return self.min_x <= x <= self.max_x and self.min_y <= y <= self.max_y

    def width(self):
        return self.max_x - self.min_x

    def height(self):
        return self.max_y - self.min_y

    def valid(self):
        return self.max_x >= self.min_x and self.max_y >= self.min_y

    def __repr__(self):
        return (f"BBox(min_x={self.min_x:f}, min_y={self.min_y:f}, "
                f"max_x={self.max_x:f}, max_y={self.max_y:f})")


class CRS:
    """Metadata only: the reference never uses the CRS for arithmetic on this path
    (reprojection is a stub upstream, src/engine/reprojection.cpp:1-11); PROJ-backed
    lookups are out of scope, so from_epsg/from_wkt just record their argument."""

    def __init__(self):
        self.wkt = ""
        self.epsg = 0

    def is_projected(self):
        return "PROJCS" in self.wkt or "PROJCRS" in self.wkt

    def is_geographic(self):
        return "GEOGCS" in self.wkt or "GEOGCRS" in self.wkt

    def is_valid(self):
        return bool(self.wkt) or self.epsg != 0

    @staticmethod
    def from_epsg(code):
        c = CRS()
        c.epsg = int(code)
        return c

    @staticmethod
    def from_wkt(wkt):
        c = CRS()
        c.wkt = str(wkt)
        return c

    def equivalent_to(self, other):
        if self.epsg != 0 and other.epsg != 0 and self.epsg == other.epsg:
            return True
        return bool(self.wkt) and self.wkt == other.wkt

    def __repr__(self):
        return f"CRS(epsg={self.epsg})" if self.epsg else f"CRS(wkt='{self.wkt[:50]}...')"


class NoDataPolicy:
    def __init__(self):
        self.value = float("nan")
        self.use_nan = True

    def sentinel(self):
        return float("nan") if self.use_nan else self.value


class TileIndex:
    def __init__(self, row=0, col=0):
        self.row = int(row)
        self.col = int(col)

    def __eq__(self, o):
        return self.row == o.row and self.col == o.col

    def __lt__(self, o):
        return self.row < o.row or (self.row == o.row and self.col < o.col)

    def __hash__(self):
        return hash((self.row, self.col))

    def __repr__(self):
        return f"TileIndex(row={self.row}, col={self.col})"


class Status:
    def __init__(self, code=StatusCode.Ok, message=""):
        self.code = code
        self.message = message

    def ok(self):
        return self.code == StatusCode.Ok

    __bool__ = ok

    @staticmethod
    def success():
        return Status()

    @staticmethod
    def error(code, msg):
        return Status(code, msg)

    def __repr__(self):
        return "Status(Ok)" if self.ok() else f"Status(code={int(self.code)}, message='{self.message}')"


class ChannelDesc:
    def __init__(self, name="", dtype=DataType.Float32, offset=0):
        self.name = name
        self.dtype = dtype
        self.offset = offset


class BandDesc:
    def __init__(self, name="", dtype=DataType.Float32, is_state=False):
        self.name = name
        self.dtype = dtype
        self.is_state = is_state


# ---------------------------------------------------------------------------
# GridConfig (bindings.cpp:189-229; src/core/grid_config.cpp)
# ---------------------------------------------------------------------------
class GridConfig:
    def __init__(self):
        self.bounds = BBox()
        self.crs = CRS()
        self.cell_size_x = 1.0
        self.cell_size_y = -1.0
        self.width = 0
        self.height = 0
        self.nodata = NoDataPolicy()
        self.tile_width = 4096
        self.tile_height = 4096
        self.tiles_x = 0
        self.tiles_y = 0

    def _desc(self) -> _lib.GridDesc:
        b = self.bounds
        return _lib.GridDesc(b.min_x, b.min_y, b.max_x, b.max_y, self.cell_size_x, self.cell_size_y,
                             int(self.width), int(self.height), int(self.tile_width),
                             int(self.tile_height))

    def compute_dimensions(self):
        if not self.bounds.valid():
            self.width = self.height = self.tiles_x = self.tiles_y = 0
            return
        d = self._desc()
        check(lib.pcr_grid_compute_dimensions(C.byref(d)))
        self.width, self.height = d.width, d.height
        self.tiles_x = (self.width + self.tile_width - 1) // self.tile_width
        self.tiles_y = (self.height + self.tile_height - 1) // self.tile_height

    def world_to_cell(self, wx, wy):
        d = self._desc()
        col, row = C.c_int32(0), C.c_int32(0)
        ok = lib.pcr_grid_world_to_cell(C.byref(d), float(wx), float(wy), C.byref(col), C.byref(row))
        return (col.value, row.value, bool(ok))

    def cell_to_world(self, col, row):
        return (self.bounds.min_x + (col + 0.5) * self.cell_size_x,
                self.bounds.max_y + (row + 0.5) * self.cell_size_y)

    def cell_to_tile(self, col, row):
        # C++ integer division truncates toward zero
        return TileIndex(int(row / self.tile_height), int(col / self.tile_width))

    def tile_cell_range(self, idx):
        col_start = idx.col * self.tile_width
        row_start = idx.row * self.tile_height
        return (col_start, row_start, min(self.tile_width, self.width - col_start),
                min(self.tile_height, self.height - row_start))

    def tile_bounds(self, idx):
        c0, r0, cc, rc = self.tile_cell_range(idx)
        b = BBox()
        b.min_x = self.bounds.min_x + c0 * self.cell_size_x
        b.max_x = self.bounds.min_x + (c0 + cc) * self.cell_size_x
        b.max_y = self.bounds.max_y + r0 * self.cell_size_y
        b.min_y = self.bounds.max_y + (r0 + rc) * self.cell_size_y
        return b

    def total_tiles(self):
        return self.tiles_x * self.tiles_y

    def total_cells(self):
        return int(self.width) * int(self.height)

    def gdal_geotransform(self):
        return [self.bounds.min_x, self.cell_size_x, 0.0, self.bounds.max_y, 0.0, self.cell_size_y]

    def validate(self):
        if not self.bounds.valid():
            raise RuntimeError("Invalid bounds: max < min")
        if self.cell_size_x == 0.0 or self.cell_size_y == 0.0:
            raise RuntimeError("Cell size cannot be zero")
        if self.tile_width <= 0 or self.tile_height <= 0:
            raise RuntimeError("Tile dimensions must be positive")
        if self.width <= 0 or self.height <= 0:
            raise RuntimeError("Grid dimensions not computed or invalid. Call compute_dimensions()")
        if not self.crs.is_valid():
            raise RuntimeError("CRS is not valid")

    def __repr__(self):
        return f"GridConfig(width={self.width}, height={self.height}, tiles={self.tiles_x}x{self.tiles_y})"


# ---------------------------------------------------------------------------
# Grid — host multi-band float raster (bindings.cpp:234-284; src/core/grid.cpp)
# ---------------------------------------------------------------------------
class Grid:
    def __init__(self, cols, rows, bands, arrays, owner=None):
        self._cols, self._rows = int(cols), int(rows)
        self._bands = list(bands)
        self._arrays = arrays          # list of (rows, cols) float32 arrays
        self._owner = owner            # keeps the Pipeline (and its pinned result memory) alive

    @staticmethod
    def create(cols, rows, bands, loc=MemoryLocation.Host):
        if loc != MemoryLocation.Host:
            return None                # Device grids are NotImplemented upstream too (grid.cpp:44-53)
        if cols <= 0 or rows <= 0 or not bands:
            return None
        arrays = [np.zeros((rows, cols), _NP_DTYPE[DataType(b.dtype)]) for b in bands]
        return Grid(cols, rows, bands, arrays)

    @staticmethod
    def create_for_tile(config, tile, bands, loc=MemoryLocation.Host):
        _, _, cc, rc = config.tile_cell_range(tile)
        return Grid.create(cc, rc, bands, loc)

    def num_bands(self):
        return len(self._bands)

    def band_desc(self, i):
        return self._bands[i]

    def band_index(self, name):
        for i, b in enumerate(self._bands):
            if b.name == name:
                return i
        return -1

    def cols(self):
        return self._cols

    def rows(self):
        return self._rows

    def cell_count(self):
        return self._cols * self._rows

    def location(self):
        return MemoryLocation.Host

    def fill(self, value):
        for a in self._arrays:
            a[...] = value

    def fill_band(self, i, value):
        if not 0 <= i < len(self._arrays):
            raise RuntimeError("Invalid band index")
        self._arrays[i][...] = value

    def band_array(self, i):
        if not 0 <= i < len(self._arrays) or self._arrays[i].dtype != np.float32:
            raise RuntimeError("Invalid band index or data type")
        return self._arrays[i]

    def set_band_array(self, i, arr):
        a = self.band_array(i)
        arr = np.asarray(arr, np.float32)
        if arr.shape != a.shape:
            raise RuntimeError("Array shape mismatch")
        a[...] = arr

    def __repr__(self):
        return f"Grid(cols={self._cols}, rows={self._rows}, bands={len(self._bands)})"


# ---------------------------------------------------------------------------
# PointCloud (bindings.cpp:289-392; src/core/point_cloud.cpp)
# ---------------------------------------------------------------------------
class _Buffer:
    """One allocation from pcr_mem_alloc (pinned host or device), freed on GC."""

    def __init__(self, loc, nbytes, device=0):
        self.loc, self.nbytes, self.device = int(loc), int(nbytes), int(device)
        p = C.c_void_p()
        check(lib.pcr_mem_alloc(self.loc, self.device, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def as_array(self, dtype, count):
        if self.loc == MemoryLocation.Device:
            raise RuntimeError("PointCloud is in Device memory; call to_host() first")
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        a = np.frombuffer(buf, dtype=dtype, count=count)
        a.flags.writeable = True
        return a

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                lib.pcr_mem_free(self.loc, self.device, self.ptr)
                self.ptr = None
        except Exception:
            pass


class PointCloud:
    """SoA cloud: x, y float64 + named channels.  Host clouds live in numpy arrays;
    HostPinned / Device clouds live in CUDA allocations made through the C-ABI."""

    def __init__(self, capacity, loc, device=0):
        self._capacity = int(capacity)
        self._count = 0
        self._loc = MemoryLocation(loc)
        self._device = int(device)
        self._crs = CRS()
        self._channels = {}     # name -> (ChannelDesc, storage)
        self._x = self._alloc(np.float64)
        self._y = self._alloc(np.float64)

    # -- storage ------------------------------------------------------------
    def _alloc(self, dtype):
        n = max(self._capacity, 1)
        if self._loc == MemoryLocation.Host:
            return np.empty(n, dtype)
        # +64 bytes so vector/bulk loads may run a few elements past `count`
        return _Buffer(self._loc, n * np.dtype(dtype).itemsize + 64, self._device)

    def _view(self, storage, dtype, count):
        if isinstance(storage, np.ndarray):
            return storage[:count]
        return storage.as_array(dtype, self._capacity)[:count]

    @staticmethod
    def _ptr(storage):
        return storage.ctypes.data if isinstance(storage, np.ndarray) else storage.ptr

    # -- reference API ------------------------------------------------------
    @staticmethod
    def create(capacity, loc=MemoryLocation.Host, device=0):
        """`device` (new, default 0) selects the GPU for Device clouds; it must be the
        pipeline's cuda_device_id."""
        try:
            return PointCloud(capacity, loc, device)
        except RuntimeError as e:
            print(f"PointCloud.create failed: {e}", file=sys.stderr)
            return None

    def add_channel(self, name, dtype=DataType.Float32):
        if name in self._channels:
            raise RuntimeError("Channel already exists: " + name)
        dtype = DataType(dtype)
        self._channels[name] = (ChannelDesc(name, dtype, 0), self._alloc(_NP_DTYPE[dtype]))

    def has_channel(self, name):
        return name in self._channels

    def channel(self, name):
        c = self._channels.get(name)
        return c[0] if c else None

    def channel_names(self):
        return list(self._channels.keys())

    def count(self):
        return self._count

    def capacity(self):
        return self._capacity

    def location(self):
        return self._loc

    def crs(self):
        return self._crs

    def set_crs(self, crs):
        self._crs = crs

    def resize(self, new_count):
        if new_count > self._capacity:
            raise RuntimeError("resize: new_count exceeds capacity")
        self._count = int(new_count)

    def x_array(self):
        return self._view(self._x, np.float64, self._count)

    def y_array(self):
        return self._view(self._y, np.float64, self._count)

    def channel_array_f32(self, name):
        c = self._channels.get(name)
        if not c or c[0].dtype != DataType.Float32:
            raise RuntimeError("Channel not found or wrong type: " + name)
        return self._view(c[1], np.float32, self._count)

    def _store(self, storage, dtype, arr, what):
        arr = np.ascontiguousarray(arr, dtype)
        n = arr.shape[0]
        if self._loc == MemoryLocation.Device:
            check(lib.pcr_mem_copy(self._ptr(storage), int(self._loc), arr.ctypes.data,
                                   int(MemoryLocation.Host), arr.nbytes, self._device))
        else:
            self._view(storage, dtype, self._capacity)[:n] = arr
        return n

    def set_x_array(self, arr):                # memcpy + resize(n), bindings.cpp:338-346
        arr = np.asarray(arr)
        if arr.shape[0] > self._capacity:
            raise RuntimeError("Array too large for capacity")
        self._count = self._store(self._x, np.float64, arr, "x")

    def set_y_array(self, arr):                # memcpy only, bindings.cpp:347-354
        arr = np.asarray(arr)
        if arr.shape[0] > self._capacity:
            raise RuntimeError("Array too large for capacity")
        self._store(self._y, np.float64, arr, "y")

    def set_channel_array_f32(self, name, arr):   # len <= count(), bindings.cpp:355-365
        c = self._channels.get(name)
        if not c or c[0].dtype != DataType.Float32:
            raise RuntimeError("Channel not found or wrong type: " + name)
        arr = np.asarray(arr)
        if arr.shape[0] > self._count:
            raise RuntimeError("Array size exceeds point count")
        self._store(c[1], np.float32, arr, name)

    def _to(self, dst, device=None):
        dev = self._device if device is None else int(device)
        out = PointCloud(max(self._capacity, 1), dst, dev)
        out._count = self._count
        out._crs = self._crs
        n = self._count

        def move(src, dst_storage, itemsize):
            check(lib.pcr_mem_copy(self._ptr(dst_storage), int(dst), self._ptr(src), int(self._loc),
                                   n * itemsize, dev))
        move(self._x, out._x, 8)
        move(self._y, out._y, 8)
        for name, (desc, storage) in self._channels.items():
            out.add_channel(name, desc.dtype)
            move(storage, out._channels[name][1], np.dtype(_NP_DTYPE[desc.dtype]).itemsize)
        return out

    def to_device(self, device=None):
        try:
            return self._to(MemoryLocation.Device, device)
        except RuntimeError as e:
            raise RuntimeError("Failed to transfer point cloud to Device memory. Possible causes: "
                               "CUDA out of memory, CUDA not initialized, or incompatible GPU "
                               f"configuration. ({e})")

    def to_host(self):
        return self._to(MemoryLocation.Host)

    def to_pinned(self):
        """New: host-pinned copy, which the ingest ring can DMA from directly."""
        return self._to(MemoryLocation.HostPinned)

    def __repr__(self):
        return (f"PointCloud(count={self._count}, capacity={self._capacity}, "
                f"channels={len(self._channels)})")


# ---------------------------------------------------------------------------
# Filter (bindings.cpp:397-412); evaluated on the device in front of routing
# ---------------------------------------------------------------------------
class FilterPredicate:
    def __init__(self):
        self.channel_name = ""
        self.op = CompareOp.Equal
        self.value = 0.0
        self.value_set = []


class FilterSpec:
    def __init__(self):
        self.predicates = []

    def add(self, channel, op, value):
        p = FilterPredicate()
        p.channel_name, p.op, p.value = channel, op, value
        self.predicates.append(p)
        return self

    def add_in_set(self, channel, values):
        p = FilterPredicate()
        p.channel_name, p.op, p.value_set = channel, CompareOp.InSet, list(values)
        self.predicates.append(p)
        return self

    def empty(self):
        return not self.predicates


# ---------------------------------------------------------------------------
# Glyph / Reduction / Pipeline config (bindings.cpp:414-471)
# ---------------------------------------------------------------------------
class GlyphSpec:
    def __init__(self):
        self.type = GlyphType.Point
        self.direction_channel = ""
        self.default_direction = 0.0
        self.half_length_channel = ""
        self.default_half_length = 1.0
        self.sigma_x_channel = ""
        self.default_sigma_x = 1.0
        self.sigma_y_channel = ""
        self.default_sigma_y = 1.0
        self.rotation_channel = ""
        self.default_rotation = 0.0
        self.max_radius_cells = 32.0
        self.normalize_weights = False

    def __repr__(self):
        return f"GlyphSpec(type={GlyphType(self.type).name})"


class ReductionSpec:
    def __init__(self):
        self.value_channel = ""
        self.type = ReductionType.Sum
        self.weight_channel = ""        # never read upstream
        self.timestamp_channel = ""
        self.percentile = 0.5
        self.output_band_name = ""
        self.glyph = GlyphSpec()


class PipelineConfig:
    def __init__(self):
        self.grid = GridConfig()
        self.reductions = []
        self.filter = FilterSpec()
        self.target_crs = CRS()
        self.auto_reproject = True
        self.exec_mode = ExecutionMode.Auto
        self.gpu_memory_budget = 0
        self.host_cache_budget = 0
        self.chunk_size = 0
        self.cpu_threads = 0
        self.gpu_fallback_to_cpu = True
        self.hybrid_cpu_threads = 0
        self.state_dir = ""
        self.resume = False
        self.output_path = ""
        self.write_cog = False
        # --- additive knobs of the B200 path (C++-only or new) ---
        self.cuda_device_id = 0
        self.deterministic = False
        self.ring_depth = 0
        self.ring_slot_points = 0
        self.staging_threads = 0
        self.point_kernel = 0
        self.warp_aggregate = 0
        self.gaussian_kernel = 0
        self.comm_mode = 0
        self.comm_root_only = False
        self.async_ingest = False
        self.comm_band_copy = 0
        self.bin_cells_log2 = 0
        self.bin_pool_points = 0
        self.comm_layout = 0


class ProgressInfo:
    def __init__(self):
        self.collections_processed = 0
        self.collections_total = 0
        self.points_processed = 0
        self.tiles_active = 0
        self.elapsed_seconds = 0.0

    @staticmethod
    def _from(p):
        o = ProgressInfo()
        o.collections_processed = int(p.collections_processed)
        o.collections_total = int(p.collections_total)
        o.points_processed = int(p.points_processed)
        o.tiles_active = int(p.tiles_active)
        o.elapsed_seconds = float(p.elapsed_seconds)
        return o

    def __repr__(self):
        return (f"ProgressInfo(points={self.points_processed}, tiles={self.tiles_active}, "
                f"elapsed={self.elapsed_seconds:f}s)")


def _b(s):
    return (s or "").encode("utf-8")


class Pipeline:
    """pcr::Pipeline (include/pcr/engine/pipeline.h:105-145) over the C-ABI."""

    def __init__(self, handle, cfg, keep):
        self._h = handle
        self._cfg = cfg
        self._keep = keep            # C strings / arrays referenced by the desc
        # The result Grid (and every band view) keeps the Pipeline alive — the pinned result memory belongs to
        # it, as with pybind11's reference_internal upstream — so the Pipeline must NOT hold the Grid strongly:
        # that cycle would leave the device memory of every finalized pipeline to the cyclic garbage collector.
        self._result_ref = None
        self._has_result = False
        self._cb = None

    @staticmethod
    def create(cfg):
        """Returns a Pipeline, or None on failure with the reason on stderr
        (the reference returns nullptr -> None, pipeline.cpp:1294-1304)."""
        keep = []
        n = len(cfg.reductions)
        reds = (_lib.ReductionDesc * max(n, 1))()
        for i, r in enumerate(cfg.reductions):
            g = r.glyph
            strs = [_b(r.value_channel), _b(r.output_band_name), _b(g.direction_channel),
                    _b(g.half_length_channel), _b(g.sigma_x_channel), _b(g.sigma_y_channel),
                    _b(g.rotation_channel)]
            keep.append(strs)
            reds[i].value_channel = strs[0]
            reds[i].type = int(r.type)
            reds[i].output_band_name = strs[1]
            reds[i].glyph = _lib.GlyphDesc(int(g.type), strs[2], g.default_direction, strs[3],
                                           g.default_half_length, strs[4], g.default_sigma_x,
                                           strs[5], g.default_sigma_y, strs[6], g.default_rotation,
                                           g.max_radius_cells, int(bool(g.normalize_weights)))
        keep.append(reds)
        desc = _lib.PipelineDesc()
        desc.grid = cfg.grid._desc()
        desc.reductions = reds
        desc.num_reductions = n
        desc.exec_mode = int(cfg.exec_mode)
        desc.gpu_fallback_to_cpu = int(bool(cfg.gpu_fallback_to_cpu))
        desc.cuda_device_id = int(getattr(cfg, "cuda_device_id", 0))
        desc.deterministic = int(getattr(cfg, "deterministic", 0))      # False/True or 0/1/2
        desc.ring_depth = int(getattr(cfg, "ring_depth", 0))
        desc.ring_slot_points = int(getattr(cfg, "ring_slot_points", 0))
        desc.staging_threads = int(getattr(cfg, "staging_threads", 0))
        desc.point_kernel = int(getattr(cfg, "point_kernel", 0))
        desc.warp_aggregate = int(getattr(cfg, "warp_aggregate", 0))
        desc.gaussian_kernel = int(getattr(cfg, "gaussian_kernel", 0))
        desc.comm_mode = int(getattr(cfg, "comm_mode", 0))
        desc.comm_root_only = int(getattr(cfg, "comm_root_only", 0))
        desc.async_ingest = int(bool(getattr(cfg, "async_ingest", False)))
        desc.comm_band_copy = int(getattr(cfg, "comm_band_copy", 0))
        desc.bin_cells_log2 = int(getattr(cfg, "bin_cells_log2", 0))
        desc.bin_pool_points = int(getattr(cfg, "bin_pool_points", 0))
        desc.comm_layout = int(getattr(cfg, "comm_layout", 0))
        preds = list(cfg.filter.predicates)
        if preds:
            arr = (_lib.FilterPredicate * len(preds))()
            for i, fp in enumerate(preds):
                name = _b(fp.channel_name)
                vs = (C.c_float * max(len(fp.value_set), 1))(*[float(v) for v in fp.value_set])
                keep += [name, vs]
                arr[i].channel_name = name
                arr[i].op = int(fp.op)
                arr[i].value = float(fp.value)
                arr[i].value_set = vs
                arr[i].value_set_size = len(fp.value_set)
            keep.append(arr)
            desc.filter = arr
            desc.num_predicates = len(preds)
        h = C.c_void_p()
        rc = lib.pcr_pipeline_create(C.byref(desc), C.byref(h))
        if rc != 0 or not h.value:
            print(f"Pipeline.create failed: {_lib.last_error()}", file=sys.stderr)
            return None
        p = Pipeline(h, cfg, keep)
        # resume: continue from the tile-state files of an earlier run.  (Upstream `resume` is never
        # read and ANY matching file in state_dir is silently reloaded, tile_manager.cpp:272-302; here
        # it takes the explicit flag.)
        if getattr(cfg, "resume", False) and cfg.state_dir:
            import os
            if os.path.isdir(cfg.state_dir):
                try:
                    p.load_state(cfg.state_dir)
                except RuntimeError as e:
                    print(f"Pipeline.create failed: {e}", file=sys.stderr)
                    return None
        return p

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                lib.pcr_pipeline_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def validate(self):
        check(lib.pcr_pipeline_validate(self._h))

    def prepare(self, cloud):
        """New: pre-marshal a cloud's C-ABI arguments once, for hot loops that ingest the same
        buffers repeatedly (bench.py); returns an opaque token for ingest_prepared()."""
        names = cloud.channel_names()
        views = (_lib.ChannelView * max(len(names), 1))()
        keep = [cloud]
        for i, name in enumerate(names):
            desc, storage = cloud._channels[name]
            nb = _b(name)
            keep.append(nb)
            views[i].name = nb
            views[i].data = PointCloud._ptr(storage)
            views[i].dtype = int(desc.dtype)
        return (C.c_void_p(PointCloud._ptr(cloud._x)), C.c_void_p(PointCloud._ptr(cloud._y)),
                C.c_size_t(cloud.count()), views, C.c_int32(len(names)), C.c_int32(int(cloud.location())), keep)

    def ingest_prepared(self, tok):
        check(lib.pcr_pipeline_ingest(self._h, tok[0], tok[1], tok[2], tok[3], tok[4], tok[5]))

    def ingest(self, cloud):
        names = cloud.channel_names()
        views = (_lib.ChannelView * max(len(names), 1))()
        keep = []
        for i, name in enumerate(names):
            desc, storage = cloud._channels[name]
            nb = _b(name)
            keep.append(nb)
            views[i].name = nb
            views[i].data = PointCloud._ptr(storage)
            views[i].dtype = int(desc.dtype)
        check(lib.pcr_pipeline_ingest(self._h, PointCloud._ptr(cloud._x), PointCloud._ptr(cloud._y),
                                      cloud.count(), views, len(names), int(cloud.location())))

    def ingest_arrays(self, x_ptr, y_ptr, count, channels, location=MemoryLocation.Device):
        """New: ingest borrowed raw buffers (addresses as ints) without a PointCloud object — e.g. arrays a
        caller already holds in HBM.  channels = {name: address of `count` float32 values}.  Device buffers
        must be complete (their producer stream synchronized) before the call."""
        names = list(channels)
        views = (_lib.ChannelView * max(len(names), 1))()
        keep = []
        for i, name in enumerate(names):
            nb = _b(name)
            keep.append(nb)
            views[i].name = nb
            views[i].data = int(channels[name])
            views[i].dtype = int(DataType.Float32)
        check(lib.pcr_pipeline_ingest(self._h, C.c_void_p(int(x_ptr)), C.c_void_p(int(y_ptr)), int(count), views,
                                      len(names), int(location)))

    def _wrap_result(self):
        bands, arrays = [], []
        for i in range(len(self._cfg.reductions)):
            p = C.c_void_p()
            rows, cols = C.c_int32(0), C.c_int32(0)
            check(lib.pcr_pipeline_result_band(self._h, i, C.byref(p), C.byref(rows), C.byref(cols)))
            nbytes = rows.value * cols.value * 4
            buf = (C.c_char * nbytes).from_address(p.value)
            buf._pcr_owner = self      # array -> buffer -> pipeline: the pinned memory outlives views
            arrays.append(np.frombuffer(buf, np.float32).reshape(rows.value, cols.value))
            name = C.create_string_buffer(512)
            check(lib.pcr_pipeline_band_name(self._h, i, name, 512))
            bands.append(BandDesc(name.value.decode(), DataType.Float32, False))
        g = self._cfg.grid
        grid = Grid(g.width, g.height, bands, arrays, owner=self)
        self._result_ref = weakref.ref(grid)
        self._has_result = True
        return grid

    def finalize(self):
        check(lib.pcr_pipeline_finalize(self._h))
        grid = self._wrap_result()
        if self._cfg.output_path:
            opts = GeoTiffOptions()
            opts.compress = "DEFLATE" if self._cfg.write_cog else "NONE"
            opts.cloud_optimized = bool(self._cfg.write_cog)
            write_geotiff(self._cfg.output_path, grid, self._cfg.grid, opts)

    def finalize_device(self):
        """New: finalize into HBM only (no D2H); see result_band_device_ptr()."""
        check(lib.pcr_pipeline_finalize_device(self._h))

    def result_band_device_ptr(self, band):
        p = C.c_void_p()
        rows, cols = C.c_int32(0), C.c_int32(0)
        check(lib.pcr_pipeline_result_band_device(self._h, band, C.byref(p), C.byref(rows), C.byref(cols)))
        return p.value, rows.value, cols.value

    def run(self, clouds):
        for c in clouds:
            if c is None:
                raise RuntimeError("pipeline: null cloud pointer")
            self.ingest(c)
        self.finalize()

    def set_progress_callback(self, fn):
        if fn is None:
            self._cb = None
            check(lib.pcr_pipeline_set_progress_callback(self._h, C.cast(None, _lib.PROGRESS_FN), None))
            return

        def tramp(info_ptr, _user):
            return 1 if fn(ProgressInfo._from(info_ptr.contents)) else 0
        self._cb = _lib.PROGRESS_FN(tramp)
        check(lib.pcr_pipeline_set_progress_callback(self._h, self._cb, None))

    def result(self):
        """Grid over the pipeline's pinned host result (None before the first finalize()); the Grid and its
        band views keep the pipeline alive, not the other way round."""
        if not self._has_result:
            return None
        grid = self._result_ref() if self._result_ref is not None else None
        return grid if grid is not None else self._wrap_result()

    def stats(self):
        p = _lib.Progress()
        check(lib.pcr_pipeline_stats(self._h, C.byref(p)))
        return ProgressInfo._from(p)

    # -- new-path extras ----------------------------------------------------
    def reset(self):
        check(lib.pcr_pipeline_reset(self._h))
        self._result_ref = None
        self._has_result = False

    def save_state(self, directory):
        """Write the accumulated state as reference-format .pcrt tile files (one per touched tile per
        reduction; single-reduction pipelines use the reference's own directory layout)."""
        check(lib.pcr_pipeline_save_state(self._h, str(directory).encode()))

    def load_state(self, directory):
        """Make the .pcrt files found in `directory` the accumulated state of their tiles."""
        check(lib.pcr_pipeline_load_state(self._h, str(directory).encode()))

    def synchronize(self):
        check(lib.pcr_pipeline_synchronize(self._h))

    def profile_enable(self, on=True):
        """on: False/0 = off, True/1 = time every kernel group, N > 1 = every Nth group of each kind."""
        check(lib.pcr_pipeline_profile_enable(self._h, int(on)))

    def profile_reset(self):
        check(lib.pcr_pipeline_profile_reset(self._h))

    def profile_read(self):
        p = _lib.Profile()
        check(lib.pcr_pipeline_profile_read(self._h, C.byref(p)))
        return {f: getattr(p, f) for f, _ in _lib.Profile._fields_}

    def timer_begin(self):
        check(lib.pcr_pipeline_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_double(0.0)
        check(lib.pcr_pipeline_timer_end(self._h, C.byref(ms)))
        return ms.value

    def comm_init(self, unique_id: bytes, rank: int, world_size: int):
        buf = C.create_string_buffer(bytes(unique_id), 128) if unique_id else None
        check(lib.pcr_pipeline_comm_init(self._h, buf, int(rank), int(world_size)))

    def comm_barrier(self):
        check(lib.pcr_pipeline_comm_barrier(self._h))

    def owned_cells(self):
        """Row-major cell range [cell0, cell1) whose finalized bands this rank produces (N>1; the whole grid on
        one GPU).  With comm_root_only = 2 the rank's band arrays are valid for exactly this range."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(lib.pcr_pipeline_owned_cells(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value


def diag_red_ceiling(points=5_000_000, cells=1_000_000, with_loads=False, reps=20, device=0) -> float:
    """Median microseconds the GPU needs for the Point kernel's reductions alone (measurement aid)."""
    us = C.c_double(0.0)
    check(lib.pcr_diag_red_ceiling(int(device), int(points), int(cells), int(bool(with_loads)), int(reps), C.byref(us)))
    return us.value


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib.pcr_comm_unique_id(buf))
    return buf.raw


def comm_slice_rows(height: int, world_size: int, rank: int):
    """Rows [row0, row1) that `rank` merges and finalizes at an N-rank finalize."""
    a, b = C.c_int32(0), C.c_int32(0)
    check(lib.pcr_comm_slice_rows(int(height), int(world_size), int(rank), C.byref(a), C.byref(b)))
    return a.value, b.value


def comm_partition_cells(cells: int, record_words: int, world_size: int, rank: int, bin_cells_log2: int = 0):
    """Tile-partitioned layout: (bin_shift, num_bins, cell0, cell1) — bin = 2^bin_shift consecutive cells, `rank` owns
    the row-major cells [cell0, cell1)."""
    sh, nb, a, b = C.c_int32(0), C.c_int32(0), C.c_uint64(0), C.c_uint64(0)
    check(lib.pcr_comm_partition_cells(int(cells), int(record_words), int(bin_cells_log2), int(world_size), int(rank),
                                       C.byref(sh), C.byref(nb), C.byref(a), C.byref(b)))
    return sh.value, nb.value, a.value, b.value


def device_count() -> int:
    return int(lib.pcr_device_count())


def device_mem_info(device=0):
    """(free_bytes, total_bytes) of a GPU — cuda_get_memory_info, include/pcr/core/types.h:186-201."""
    free, total = C.c_uint64(0), C.c_uint64(0)
    check(lib.pcr_device_mem_info(int(device), C.byref(free), C.byref(total)))
    return free.value, total.value


def device_name(device=0) -> str:
    buf = C.create_string_buffer(256)
    lib.pcr_device_name(int(device), buf, 256)
    return buf.value.decode()


# ---------------------------------------------------------------------------
# GeoTIFF / point-cloud file IO (SURVEY §8f "next" rows)
# ---------------------------------------------------------------------------
class GeoTiffOptions:
    def __init__(self):
        self.cloud_optimized = False
        self.compress = "LZW"
        self.compress_level = 6
        self.tile_width = 256
        self.tile_height = 256
        self.bigtiff = True
        self.overview_resampling = "AVERAGE"


def write_geotiff(path, grid, config, options=None):
    from .geotiff import write_geotiff as _w
    _w(path, grid, config, options or GeoTiffOptions())


def read_geotiff_info(path):
    from .geotiff import read_geotiff_info as _r
    return _r(path)


def read_geotiff_band(path, band_index, width, height):
    from .geotiff import read_geotiff_band as _r
    return _r(path, band_index, width, height)


def __getattr__(name):                      # TiledGeoTiffWriter lives in .geotiff (needs the classes above)
    if name == "TiledGeoTiffWriter":
        from .geotiff import TiledGeoTiffWriter
        return TiledGeoTiffWriter
    raise AttributeError(name)


class PointCloudInfo:
    def __init__(self):
        self.num_points = 0
        self.channels = []
        self.crs = CRS()
        self.bounds = BBox()


def read_point_cloud(path, format=PointCloudFormat.Auto):
    from .pointcloud_io import read_point_cloud as _f
    return _f(path, format)


def write_point_cloud(path, cloud, format=PointCloudFormat.PCR_Binary):
    from .pointcloud_io import write_point_cloud as _f
    _f(path, cloud, format)


def read_point_cloud_info(path, format=PointCloudFormat.Auto):
    from .pointcloud_io import read_point_cloud_info as _f
    return _f(path, format)


from .pointcloud_io import PointCloudReader  # noqa: E402


# ---------------------------------------------------------------------------
# Convenience helpers (python/pcr/__init__.py:73-181 of the reference)
# ---------------------------------------------------------------------------
def gaussian_splat_spec(value_channel, sigma_x_channel="", sigma_y_channel="", rotation_channel="",
                        default_sigma=1.0, default_sigma_x=None, default_sigma_y=None,
                        default_rotation=0.0, max_radius_cells=32.0, output_band_name=None):
    """ReductionSpec for Gaussian glyph splatting (WeightedAverage of value by footprint weight)."""
    spec = ReductionSpec()
    spec.value_channel = value_channel
    spec.type = ReductionType.WeightedAverage
    g = spec.glyph
    g.type = GlyphType.Gaussian
    g.sigma_x_channel, g.sigma_y_channel, g.rotation_channel = sigma_x_channel, sigma_y_channel, rotation_channel
    g.default_sigma_x = default_sigma if default_sigma_x is None else default_sigma_x
    g.default_sigma_y = default_sigma if default_sigma_y is None else default_sigma_y
    g.default_rotation = default_rotation
    g.max_radius_cells = max_radius_cells
    if output_band_name:
        spec.output_band_name = output_band_name
    return spec


def line_splat_spec(value_channel, direction_channel="", half_length_channel="", default_direction=0.0,
                    default_half_length=1.0, max_radius_cells=32.0, output_band_name=None):
    """ReductionSpec for Line glyph splatting (1-cell-wide Bresenham segment per point)."""
    spec = ReductionSpec()
    spec.value_channel = value_channel
    spec.type = ReductionType.WeightedAverage
    g = spec.glyph
    g.type = GlyphType.Line
    g.direction_channel, g.half_length_channel = direction_channel, half_length_channel
    g.default_direction, g.default_half_length = default_direction, default_half_length
    g.max_radius_cells = max_radius_cells
    if output_band_name:
        spec.output_band_name = output_band_name
    return spec


__all__ = [
    'DataType', 'ReductionType', 'MemoryLocation', 'ExecutionMode', 'StatusCode', 'CompareOp',
    'PointCloudFormat', 'GlyphType', 'BBox', 'CRS', 'NoDataPolicy', 'TileIndex', 'Status',
    'ChannelDesc', 'BandDesc', 'GridConfig', 'Grid', 'PointCloud', 'FilterPredicate', 'FilterSpec',
    'GlyphSpec', 'ReductionSpec', 'PipelineConfig', 'ProgressInfo', 'Pipeline',
    'gaussian_splat_spec', 'line_splat_spec', 'GeoTiffOptions', 'write_geotiff',
    'read_geotiff_info', 'read_geotiff_band', 'TiledGeoTiffWriter', 'PointCloudInfo', 'read_point_cloud', 'write_point_cloud',
    'read_point_cloud_info', 'PointCloudReader',
    'comm_unique_id', 'comm_slice_rows', 'comm_partition_cells', 'device_count', 'device_name',
]
