"""TEST INFRASTRUCTURE — Python face of the parity checkers.

Two checkers live here, neither is ever imported by the product package:

* ``Oracle``   — ctypes binding of ``oracle/libpcr_oracle.so`` (our plain-C
  restatement, ``oracle/pcr_oracle.c``).  Travels to the GPU box as a built .so.
* ``load_reference()`` — the UNMODIFIED reference, compiled by ``oracle/Makefile``
  into ``oracle/_ref/_pcr*.so`` (pybind11 module of /root/reference/python/bindings.cpp).
  Used to pin the oracle (tests/test_oracle.py), to generate tests/golden/
  (oracle/make_golden.py) and as the timed CPU baseline in bench.py.

Both take the same duck-typed inputs: a grid description and "spec-like" objects
carrying the reference's ReductionSpec / GlyphSpec field names
(/root/reference/include/pcr/engine/pipeline.h:20-34, glyph.h:19-43), so the
tests can hand the product's own ``pcr.ReductionSpec`` objects to either.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpcr_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")

SUM, MAX, MIN, AVERAGE, WEIGHTED_AVERAGE, COUNT = 0, 1, 2, 3, 4, 5
GLYPH_POINT, GLYPH_LINE, GLYPH_GAUSSIAN = 0, 1, 2


class _Grid(C.Structure):
    _fields_ = [("min_x", C.c_double), ("min_y", C.c_double),
                ("max_x", C.c_double), ("max_y", C.c_double),
                ("cell_size_x", C.c_double), ("cell_size_y", C.c_double),
                ("width", C.c_int32), ("height", C.c_int32),
                ("tile_width", C.c_int32), ("tile_height", C.c_int32)]


_FP = C.POINTER(C.c_float)


class _Reduction(C.Structure):
    _fields_ = [("type", C.c_int32), ("glyph", C.c_int32),
                ("value", _FP), ("direction", _FP), ("half_length", _FP),
                ("sigma_x", _FP), ("sigma_y", _FP), ("rotation", _FP),
                ("default_direction", C.c_float), ("default_half_length", C.c_float),
                ("default_sigma_x", C.c_float), ("default_sigma_y", C.c_float),
                ("default_rotation", C.c_float), ("max_radius_cells", C.c_float)]


def build(force: bool = False) -> str:
    """Compile oracle/pcr_oracle.c -> libpcr_oracle.so (gcc, seconds)."""
    src = os.path.join(HERE, "pcr_oracle.c")
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return LIB_PATH


def build_reference() -> bool:
    """Compile the reference into oracle/_ref when /root/reference is present."""
    if not os.path.isdir(os.environ.get("PCR_REFERENCE_DIR", "/root/reference")):
        return False
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return True


def _int(v):
    """Enum-or-int -> int (pybind enums, IntEnum and plain ints all work)."""
    return int(v.value) if hasattr(v, "value") and not isinstance(v, int) else int(v)


class GridDesc:
    """Plain description of a GridConfig (reference grid_config.h:17-75)."""

    def __init__(self, min_x, min_y, max_x, max_y, cell_size_x=1.0, cell_size_y=-1.0,
                 tile_width=4096, tile_height=4096, width=None, height=None):
        self.min_x, self.min_y, self.max_x, self.max_y = map(float, (min_x, min_y, max_x, max_y))
        self.cell_size_x, self.cell_size_y = float(cell_size_x), float(cell_size_y)
        self.tile_width, self.tile_height = int(tile_width), int(tile_height)
        self.width, self.height = width, height

    @classmethod
    def from_config(cls, gc):
        """From any GridConfig-like object (product pcr.GridConfig or reference)."""
        b = gc.bounds
        return cls(b.min_x, b.min_y, b.max_x, b.max_y, gc.cell_size_x, gc.cell_size_y,
                   gc.tile_width, gc.tile_height, gc.width, gc.height)


class Oracle:
    def __init__(self):
        build()
        self.lib = C.CDLL(LIB_PATH)
        L = self.lib
        L.orc_compute_dimensions.argtypes = [C.POINTER(_Grid)]
        L.orc_world_to_cell.argtypes = [C.POINTER(_Grid), C.c_double, C.c_double,
                                        C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_world_to_cell.restype = C.c_int
        L.orc_assign.argtypes = [C.POINTER(_Grid), C.c_void_p, C.c_void_p, C.c_size_t,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_state_floats.argtypes = [C.c_int]
        L.orc_state_floats.restype = C.c_int
        L.orc_state_init.argtypes = [C.c_int, C.c_void_p, C.c_int64]
        L.orc_state_merge.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_finalize.argtypes = [C.POINTER(_Grid), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_accumulate.argtypes = [C.POINTER(_Grid), C.POINTER(_Reduction), C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_accumulate.restype = C.c_int
        L.orc_accumulate_bounds.argtypes = [C.POINTER(_Grid), C.POINTER(_Reduction), C.c_void_p,
                                            C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_void_p]
        L.orc_accumulate_bounds.restype = C.c_int

    # -- geometry -----------------------------------------------------------
    def grid(self, gd: GridDesc) -> _Grid:
        g = _Grid(gd.min_x, gd.min_y, gd.max_x, gd.max_y, gd.cell_size_x, gd.cell_size_y,
                  0, 0, gd.tile_width, gd.tile_height)
        if gd.width is None or gd.height is None:
            self.lib.orc_compute_dimensions(C.byref(g))
        else:
            g.width, g.height = int(gd.width), int(gd.height)
        return g

    def compute_dimensions(self, gd: GridDesc):
        g = self.grid(GridDesc(gd.min_x, gd.min_y, gd.max_x, gd.max_y, gd.cell_size_x,
                               gd.cell_size_y, gd.tile_width, gd.tile_height))
        return g.width, g.height

    def world_to_cell(self, gd: GridDesc, wx, wy):
        g = self.grid(gd)
        c, r = C.c_int32(0), C.c_int32(0)
        ok = self.lib.orc_world_to_cell(C.byref(g), wx, wy, C.byref(c), C.byref(r))
        return c.value, r.value, bool(ok)

    def assign(self, gd: GridDesc, x, y):
        g = self.grid(gd)
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        n = x.size
        cell = np.zeros(n, np.uint32); tile = np.zeros(n, np.uint32); valid = np.zeros(n, np.uint8)
        self.lib.orc_assign(C.byref(g), x.ctypes.data, y.ctypes.data, n,
                            cell.ctypes.data, tile.ctypes.data, valid.ctypes.data)
        return cell, tile, valid

    # -- reducers -----------------------------------------------------------
    def _reduction(self, spec, chans, keep):
        """spec-like -> _Reduction; `keep` collects arrays that must stay alive."""
        gl = spec.glyph
        rd = _Reduction()
        rd.type = _int(spec.type)
        rd.glyph = _int(gl.type)

        def chan(name):
            if not name or name not in chans:
                return None
            a = np.ascontiguousarray(chans[name], np.float32)
            keep.append(a)
            return a.ctypes.data_as(_FP)

        if spec.value_channel not in chans:
            raise KeyError("pipeline: value channel not found: " + spec.value_channel)
        rd.value = chan(spec.value_channel)
        rd.direction = chan(gl.direction_channel)
        rd.half_length = chan(gl.half_length_channel)
        rd.sigma_x = chan(gl.sigma_x_channel)
        rd.sigma_y = chan(gl.sigma_y_channel)
        rd.rotation = chan(gl.rotation_channel)
        rd.default_direction = gl.default_direction
        rd.default_half_length = gl.default_half_length
        rd.default_sigma_x = gl.default_sigma_x
        rd.default_sigma_y = gl.default_sigma_y
        rd.default_rotation = gl.default_rotation
        rd.max_radius_cells = gl.max_radius_cells
        return rd

    def run(self, gd: GridDesc, clouds, specs, return_state=False):
        """Full ingest(+ingest...)+finalize.  `clouds` = list of (x, y, {name: f32 array}).
        Returns one (height,width) float32 band per spec."""
        g = self.grid(gd)
        cells = g.width * g.height
        ntiles = ((g.width + g.tile_width - 1) // g.tile_width) * \
                 ((g.height + g.tile_height - 1) // g.tile_height)
        touched = np.zeros(ntiles, np.uint8)
        states = []
        for s in specs:
            t = _int(s.type)
            st = np.empty(self.lib.orc_state_floats(t) * cells, np.float32)
            self.lib.orc_state_init(t, st.ctypes.data, cells)
            states.append(st)
        for (x, y, chans) in clouds:
            x = np.ascontiguousarray(x, np.float64)
            y = np.ascontiguousarray(y, np.float64)
            for s, st in zip(specs, states):
                keep = []
                rd = self._reduction(s, chans, keep)
                rc = self.lib.orc_accumulate(C.byref(g), C.byref(rd), x.ctypes.data, y.ctypes.data,
                                             x.size, st.ctypes.data, touched.ctypes.data)
                if rc != 0:
                    raise RuntimeError("pipeline: glyph splatting only supports WeightedAverage, "
                                       "Average, Sum, or Count reduction types")
        bands = []
        for s, st in zip(specs, states):
            out = np.empty((g.height, g.width), np.float32)
            self.lib.orc_finalize(C.byref(g), _int(s.type), st.ctypes.data, touched.ctypes.data,
                                  out.ctypes.data)
            bands.append(out)
        if return_state:
            return bands, states, touched
        return bands

    def bounds(self, gd: GridDesc, clouds, spec, want_weight=False):
        """(sum64, abs64, count) planes of the value (or weight) contributions."""
        g = self.grid(gd)
        cells = g.width * g.height
        s64 = np.zeros(cells, np.float64); a64 = np.zeros(cells, np.float64)
        cnt = np.zeros(cells, np.uint32)
        for (x, y, chans) in clouds:
            x = np.ascontiguousarray(x, np.float64)
            y = np.ascontiguousarray(y, np.float64)
            keep = []
            rd = self._reduction(spec, chans, keep)
            self.lib.orc_accumulate_bounds(C.byref(g), C.byref(rd), x.ctypes.data, y.ctypes.data,
                                           x.size, int(want_weight), s64.ctypes.data,
                                           a64.ctypes.data, cnt.ctypes.data)
        shp = (g.height, g.width)
        return s64.reshape(shp), a64.reshape(shp), cnt.reshape(shp)


# ---------------------------------------------------------------------------
# The reference itself (oracle/_ref)
# ---------------------------------------------------------------------------

def reference_available(gpu: bool = False) -> bool:
    d = os.path.join(REF_DIR, "gpu") if gpu else REF_DIR
    return os.path.isdir(d) and any(f.startswith("_pcr.") for f in os.listdir(d))


_ref_mod = {}


def load_reference(gpu: bool = False):
    """Import the reference's pybind11 module `_pcr` from oracle/_ref (CPU build) or
    oracle/_ref/gpu (its own CUDA mode).  Only one of the two can live in a process
    (same module name), which is fine: tests use the CPU build, bench's reference-GPU
    leg runs in its own process."""
    key = "gpu" if gpu else "cpu"
    if key in _ref_mod:
        return _ref_mod[key]
    if _ref_mod:
        raise RuntimeError("another build of the reference module is already loaded")
    d = os.path.join(REF_DIR, "gpu") if gpu else REF_DIR
    sys.path.insert(0, d)
    try:
        import _pcr  # noqa
    finally:
        sys.path.remove(d)
    _ref_mod[key] = _pcr
    return _pcr


def to_reference_spec(ref, spec):
    """Product/duck ReductionSpec -> reference ReductionSpec."""
    r = ref.ReductionSpec()
    r.value_channel = spec.value_channel
    r.type = ref.ReductionType(_int(spec.type))
    if getattr(spec, "output_band_name", ""):
        r.output_band_name = spec.output_band_name
    g, rg = spec.glyph, r.glyph
    rg.type = ref.GlyphType(_int(g.type))
    for f in ("direction_channel", "half_length_channel", "sigma_x_channel", "sigma_y_channel",
              "rotation_channel", "default_direction", "default_half_length", "default_sigma_x",
              "default_sigma_y", "default_rotation", "max_radius_cells", "normalize_weights"):
        setattr(rg, f, getattr(g, f))
    r.glyph = rg
    return r


def reference_grid(ref, gd: GridDesc):
    b = ref.BBox()
    b.min_x, b.min_y, b.max_x, b.max_y = gd.min_x, gd.min_y, gd.max_x, gd.max_y
    gc = ref.GridConfig()
    gc.bounds = b
    gc.cell_size_x, gc.cell_size_y = gd.cell_size_x, gd.cell_size_y
    gc.tile_width, gc.tile_height = gd.tile_width, gd.tile_height
    gc.compute_dimensions()
    if gd.width is not None and gd.height is not None:
        gc.width, gc.height = int(gd.width), int(gd.height)
        gc.tiles_x = (gc.width + gc.tile_width - 1) // gc.tile_width
        gc.tiles_y = (gc.height + gc.tile_height - 1) // gc.tile_height
    return gc


def reference_cloud(ref, x, y, chans):
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    c = ref.PointCloud.create(max(int(x.size), 1))
    c.set_x_array(x)
    c.set_y_array(y)
    for k, v in chans.items():
        c.add_channel(k, ref.DataType.Float32)
        c.set_channel_array_f32(k, np.ascontiguousarray(v, np.float32))
    return c


def reference_run(gd: GridDesc, clouds, specs, cpu_threads=1, exec_mode="CPU", gpu=False):
    """Run the unmodified reference Pipeline on the same inputs; returns bands.
    Always a fresh state_dir (the reference reloads stale .pcrt files otherwise,
    tile_manager.cpp:272-302)."""
    ref = load_reference(gpu)
    cfg = ref.PipelineConfig()
    cfg.grid = reference_grid(ref, gd)
    cfg.reductions = [to_reference_spec(ref, s) for s in specs]
    cfg.exec_mode = getattr(ref.ExecutionMode, exec_mode)
    cfg.cpu_threads = cpu_threads
    tmp = tempfile.mkdtemp(prefix="pcr_ref_state_",
                           dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    cfg.state_dir = tmp
    try:
        p = ref.Pipeline.create(cfg)
        if p is None:
            raise RuntimeError("reference Pipeline.create returned None")
        for (x, y, chans) in clouds:
            p.ingest(reference_cloud(ref, x, y, chans))
        p.finalize()
        g = p.result()
        return [np.array(g.band_array(i)) for i in range(len(specs))]
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
