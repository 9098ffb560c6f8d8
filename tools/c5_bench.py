#!/usr/bin/env python
"""BASELINE config 5 at single-GPU scale: clustered LiDAR-like points, Average+Max+Count fused, on a
20000x20000 grid (6.4 GB of accumulator records, far beyond L2).  python tools/c5_bench.py [points_per_ingest] [ingests]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
from pointcloud_raster_b200 import pcr

W = 20000
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dist = sys.argv[3] if len(sys.argv) > 3 else "clustered"      # or "uniform": the worst case for record locality


def clustered(n, seed):
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        return rng.uniform(0, W, n), rng.uniform(0, W, n), rng.uniform(0, 1, n).astype(np.float32)
    K = 64
    cx, cy = rng.uniform(0, W, K), rng.uniform(0, W, K)
    sig = np.exp(rng.uniform(np.log(50), np.log(2000), K))
    which = rng.integers(0, K, n)
    x = np.clip(rng.normal(cx[which], sig[which]), 0, W)
    y = np.clip(rng.normal(cy[which], sig[which]), 0, W)
    return x, y, (which / K + rng.normal(0, 0.05, n)).astype(np.float32)


gc = pcr.GridConfig(); gc.bounds.min_x = gc.bounds.min_y = 0.0; gc.bounds.max_x = gc.bounds.max_y = float(W)
gc.compute_dimensions()
specs = []
for t in (pcr.ReductionType.Average, pcr.ReductionType.Max, pcr.ReductionType.Count):
    s = pcr.ReductionSpec(); s.value_channel = "value"; s.type = t; specs.append(s)
cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = pcr.ExecutionMode.GPU
cfg.async_ingest = True
p = pcr.Pipeline.create(cfg)
assert p is not None
clouds = []
for i in range(2):
    x, y, v = clustered(n, 42 + i)
    c = pcr.PointCloud.create(n); c.set_x_array(x); c.set_y_array(y)
    c.add_channel("value", pcr.DataType.Float32); c.set_channel_array_f32("value", v)
    clouds.append((c.to_device(), c))
p.profile_enable(True)
p.ingest(clouds[0][0]); p.synchronize(); p.profile_reset()
t0 = time.perf_counter()
for i in range(k):
    p.ingest(clouds[i % 2][0])
p.synchronize()
dt = time.perf_counter() - t0
pr = p.profile_read()
print(f"device-resident ingest: {k} x {n} {dist} points on {W}x{W}: {k*n/dt/1e6:.1f} Mpts/s "
      f"(accumulate {pr['accumulate_ms']/k:.2f} ms per ingest = {n*20/(pr['accumulate_ms']/k*1e-3)/1e9:.1f} GB/s algorithmic)")
t0 = time.perf_counter(); p.finalize_device(); p.synchronize(); t1 = time.perf_counter() - t0
t0 = time.perf_counter(); p.finalize(); t2 = time.perf_counter() - t0
print(f"finalize_device {t1*1e3:.1f} ms, finalize (with D2H of 3 x 1.6 GB bands) {t2*1e3:.1f} ms")
t0 = time.perf_counter(); p.ingest(clouds[0][1]); p.synchronize(); dt = time.perf_counter() - t0
print(f"host (pageable) ingest of {n} points: {n/dt/1e6:.1f} Mpts/s")
cnt = np.asarray(p.result().band_array(2))
print("count band sum", float(np.nansum(cnt.astype(np.float64))), "expected", (k + 1) * n)
