#!/usr/bin/env python
"""Small all-kernels workload for compute-sanitizer (memcheck / racecheck): every Point variant, Line,
Gaussian scatter + gather (plain, rotated, exact-path), deterministic sort path, finalize."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
from pointcloud_raster_b200 import pcr
from util import make_grid, spec, run_product

rng = np.random.default_rng(1)
n = 20000
gc = make_grid(pcr, 150, 97, tile=64)
x, y = rng.uniform(-2, 152, n), rng.uniform(-2, 99, n)
ch = {"value": rng.uniform(0, 1, n).astype(np.float32), "s": rng.uniform(0.3, 4, n).astype(np.float32),
      "rot": rng.uniform(-3, 3, n).astype(np.float32), "d": rng.uniform(0, 3, n).astype(np.float32)}
R = pcr.ReductionType
pt = [spec(pcr, "value", t) for t in (R.Sum, R.Max, R.Min, R.Average, R.Count)]
for knobs in ({}, {"point_kernel": 2}, {"deterministic": True}, {"warp_aggregate": 2}):
    for loc in (pcr.MemoryLocation.Host, pcr.MemoryLocation.Device):
        run_product(pcr, gc, [(x, y, ch)], pt, loc=loc, **knobs)
line = pcr.line_splat_spec("value", "d", default_half_length=5.0, max_radius_cells=8.0)
g1 = pcr.gaussian_splat_spec("value", "s", "s", max_radius_cells=12.0)
g2 = pcr.gaussian_splat_spec("value", "s", "s", "rot", max_radius_cells=9.0)
g3 = pcr.gaussian_splat_spec("value", default_sigma_x=3.0, default_sigma_y=0.4, max_radius_cells=10.0)
for k in (1, 2):
    run_product(pcr, gc, [(x, y, ch)], [line, g1, g2, g3], gaussian_kernel=k)
print("sanitize driver done")
