#!/usr/bin/env python
"""Small driver for profiling one glyph configuration: python tools/run_glyph.py <config> <n> [kernel]
config in point_avg | line_hl16 | gauss_s4 | gauss_s16 (same specs as tools/glyph_bench.py)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
from pointcloud_raster_b200 import pcr
import glyph_bench as gb

name, n = sys.argv[1], int(sys.argv[2])
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
extra = dict(a.split("=") for a in sys.argv[4:])          # e.g. point_kernel=2 deterministic=1
x, y, ch = gb.arrays(n)
b = pcr.BBox(); b.min_x = b.min_y = 0.0; b.max_x = b.max_y = float(gb.GRID)
gc = pcr.GridConfig(); gc.bounds = b; gc.compute_dimensions()
cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = [gb.make_spec(pcr, name)]
cfg.exec_mode = pcr.ExecutionMode.GPU; cfg.gaussian_kernel = kernel
for k_, v_ in extra.items():
    setattr(cfg, k_, int(v_))
p = pcr.Pipeline.create(cfg)
c = pcr.PointCloud.create(n); c.set_x_array(x); c.set_y_array(y)
for k, v in ch.items():
    c.add_channel(k, pcr.DataType.Float32); c.set_channel_array_f32(k, v)
d = c.to_device()
p.profile_enable(True)
for i in range(3):
    p.profile_reset()
    t0 = time.perf_counter(); p.ingest(d); p.finalize(); dt = time.perf_counter() - t0
    pr = p.profile_read()
    print(f"{name} n={n} kernel={kernel} {extra}: sort {pr['sort_ms']:.3f} ms, wall {dt*1e3:.3f} ms, accumulate {pr['accumulate_ms']:.3f} ms, finalize {pr['finalize_ms']:.3f} ms")
