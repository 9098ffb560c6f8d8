"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what
include/pcr_b200.h declares, the ctypes table mirrors it, the host-side GridConfig logic matches
the reference's known answers, and the product fails loudly without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pcr_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcr_[a-z0-9_]+)\s*\(", src)) - {"pcr_progress_fn"})


def test_library_exports_every_declared_symbol():
    from pointcloud_raster_b200 import _lib
    names = header_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in pcr_b200.h but not exported by libpcr_b200.so"


def test_ctypes_table_matches_header():
    from pointcloud_raster_b200 import _lib
    assert sorted(n for n, _, _ in _lib.SYMBOLS) == header_functions()


def test_struct_sizes_match_c_layout():
    """Compile a tiny C program against the header and compare sizeof() with the ctypes mirrors."""
    import subprocess, tempfile
    from pointcloud_raster_b200 import _lib
    prog = r'''
#include <stdio.h>
#include "pcr_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(pcr_grid_desc), sizeof(pcr_glyph_desc),
  sizeof(pcr_reduction_desc), sizeof(pcr_pipeline_desc), sizeof(pcr_channel_view), sizeof(pcr_progress),
  sizeof(pcr_profile)); return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"),
                               os.path.join(d, "t.c")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "t")]).split()))
    mirrors = [_lib.GridDesc, _lib.GlyphDesc, _lib.ReductionDesc, _lib.PipelineDesc, _lib.ChannelView,
               _lib.Progress, _lib.Profile]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_no_oracle_or_torch_in_product():
    """The product path must not import the oracle, numpy fallbacks for compute, or torch."""
    pkg = os.path.join(ROOT, "pointcloud_raster_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import torch" not in text and "#include <torch" not in text, f
                assert "import oracle" not in text and "pcr_oracle" not in text and "oracle/" not in text, f


def test_grid_config_known_answers(pcr):
    """tests/cpp/test_grid_config.cpp of the reference."""
    gc = pcr.GridConfig()
    gc.bounds.min_x, gc.bounds.min_y, gc.bounds.max_x, gc.bounds.max_y = 0.0, 0.0, 100.0, 100.0
    gc.compute_dimensions()
    assert (gc.width, gc.height, gc.tiles_x, gc.tiles_y) == (100, 100, 1, 1)        # :12-29
    assert gc.world_to_cell(50.0, 50.0) == (50, 50, True)                            # :81-90
    assert gc.world_to_cell(0.0, 100.0) == (0, 0, True)                              # :92-101
    assert gc.world_to_cell(-10.0, 50.0)[2] is False                                 # :103-110
    assert gc.world_to_cell(100.0, 0.0) == (99, 99, True)        # inclusive max edge clamps
    assert gc.world_to_cell(float("nan"), 1.0)[2] is False
    assert gc.cell_to_world(0, 0) == (0.5, 99.5)                                     # :112-121
    gc.bounds.max_x = gc.bounds.max_y = 100.5
    gc.compute_dimensions()
    assert (gc.width, gc.height) == (101, 101)                                       # :31-44
    gc.tile_width = gc.tile_height = 32
    gc.compute_dimensions()
    assert gc.total_tiles() == 16 and gc.tile_cell_range(pcr.TileIndex(3, 3)) == (96, 96, 5, 5)   # :187-210
    assert gc.cell_to_tile(40, 70).row == 2 and gc.cell_to_tile(40, 70).col == 1
    assert gc.gdal_geotransform() == [0.0, 1.0, 0.0, 100.5, 0.0, -1.0]               # :231-247
    bad = pcr.GridConfig()
    with pytest.raises(RuntimeError):
        bad.validate()


def test_world_to_cell_matches_oracle_on_random_grids(pcr, oracle):
    import oracle as orc
    rng = np.random.default_rng(0)
    for _ in range(20):
        gc = pcr.GridConfig()
        gc.bounds.min_x, gc.bounds.min_y = rng.uniform(-1e4, 1e4, 2)
        gc.bounds.max_x = gc.bounds.min_x + rng.uniform(1, 500)
        gc.bounds.max_y = gc.bounds.min_y + rng.uniform(1, 500)
        gc.cell_size_x = float(rng.choice([1.0, 0.3, 2.5, 0.125]))
        gc.cell_size_y = -float(rng.choice([1.0, 0.7, 2.0]))
        gc.compute_dimensions()
        gd = orc.GridDesc.from_config(gc)
        assert oracle.compute_dimensions(gd) == (gc.width, gc.height)
        for _ in range(200):
            wx = rng.uniform(gc.bounds.min_x - 1, gc.bounds.max_x + 1)
            wy = rng.uniform(gc.bounds.min_y - 1, gc.bounds.max_y + 1)
            a, b = gc.world_to_cell(wx, wy), oracle.world_to_cell(gd, wx, wy)
            assert a[2] == b[2] and (not a[2] or a == b)


def test_point_cloud_api(pcr):
    c = pcr.PointCloud.create(10)
    assert (c.count(), c.capacity(), c.location()) == (0, 10, pcr.MemoryLocation.Host)
    c.set_x_array(np.arange(4, dtype=np.float64))           # set_x_array resizes (bindings.cpp:338-346)
    c.set_y_array(np.arange(4, dtype=np.float64) * 2)
    assert c.count() == 4
    c.add_channel("v", pcr.DataType.Float32)
    c.set_channel_array_f32("v", np.ones(4, np.float32))
    assert c.has_channel("v") and c.channel_names() == ["v"] and c.channel("v").dtype == pcr.DataType.Float32
    np.testing.assert_array_equal(c.y_array(), [0, 2, 4, 6])
    np.testing.assert_array_equal(c.channel_array_f32("v"), np.ones(4, np.float32))
    with pytest.raises(RuntimeError, match="exceeds point count"):
        c.set_channel_array_f32("v", np.ones(5, np.float32))
    with pytest.raises(RuntimeError, match="too large"):
        c.set_x_array(np.zeros(11))
    with pytest.raises(RuntimeError, match="Channel not found"):
        c.channel_array_f32("nope")


def test_spec_helpers_match_reference_defaults(pcr):
    g = pcr.gaussian_splat_spec("z", sigma_x_channel="s", default_sigma=2.0, max_radius_cells=9.0)
    assert g.type == pcr.ReductionType.WeightedAverage and g.glyph.type == pcr.GlyphType.Gaussian
    assert (g.glyph.default_sigma_x, g.glyph.default_sigma_y, g.glyph.sigma_x_channel) == (2.0, 2.0, "s")
    l = pcr.line_splat_spec("z", "d", default_half_length=3.0)
    assert l.type == pcr.ReductionType.WeightedAverage and l.glyph.type == pcr.GlyphType.Line
    assert (l.glyph.direction_channel, l.glyph.default_half_length, l.glyph.max_radius_cells) == ("d", 3.0, 32.0)
    d = pcr.GlyphSpec()                                      # glyph.h:19-43 defaults
    assert (d.default_direction, d.default_half_length, d.default_sigma_x, d.max_radius_cells) == (0.0, 1.0, 1.0, 32.0)


def test_no_cpu_fallback(pcr, capsys):
    """Without a device (this container) create must FAIL, whatever gpu_fallback_to_cpu says;
    ExecutionMode.CPU is refused everywhere."""
    gc = pcr.GridConfig()
    gc.bounds.min_x = gc.bounds.min_y = 0.0
    gc.bounds.max_x = gc.bounds.max_y = 8.0
    gc.compute_dimensions()
    cfg = pcr.PipelineConfig()
    cfg.grid = gc
    s = pcr.ReductionSpec(); s.value_channel = "v"; s.type = pcr.ReductionType.Sum
    cfg.reductions = [s]
    cfg.exec_mode = pcr.ExecutionMode.CPU
    assert pcr.Pipeline.create(cfg) is None
    assert "no CPU fallback" in capsys.readouterr().err
    if pcr.device_count() == 0:
        for mode in (pcr.ExecutionMode.GPU, pcr.ExecutionMode.Auto, pcr.ExecutionMode.Hybrid):
            cfg.exec_mode = mode
            cfg.gpu_fallback_to_cpu = True
            assert pcr.Pipeline.create(cfg) is None
        assert "no CPU fallback" in capsys.readouterr().err
