#!/usr/bin/env python
"""API-scope comparison on the GPU box: numpy host arrays -> ingest + finalize ->
host band, best-of-3 after one warm-up, pipeline created outside the timer — exactly how the
reference's scripts/benchmarks/benchmark_glyph_full.py:80-100 times itself — for

    ours      pointcloud_raster_b200 (this repo)
    ref_gpu   the reference's own CUDA mode, compiled unmodified for sm_100 (oracle/_ref/gpu)
    ref_cpu   the reference's CPU mode (oracle/_ref), all host cores

on the BASELINE.json glyph configs (5M points, 1000x1000 grid): Point Average, Line hl=16
(direction + half_length channels), Gaussian sigma=4 and sigma=16 (per-point sigma channel,
max_radius_cells=32).  Each implementation runs in its own process (the two reference builds
share the module name `_pcr`).  `python tools/glyph_bench.py all` prints one JSON table.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

GRID = 1000
CONFIGS = ["point_avg", "line_hl16", "gauss_s4", "gauss_s16"]


def arrays(n):
    rng = np.random.default_rng(42)
    x = rng.uniform(2, GRID - 2, n); y = rng.uniform(2, GRID - 2, n)
    ch = {"value": rng.uniform(0, 1, n).astype(np.float32),
          "direction": rng.uniform(0, np.pi, n).astype(np.float32),
          "half_length": np.full(n, 16.0, np.float32),
          "sigma4": np.full(n, 4.0, np.float32), "sigma16": np.full(n, 16.0, np.float32)}
    return x, y, ch


def make_spec(api, name):
    if name == "point_avg":
        s = api.ReductionSpec(); s.value_channel = "value"; s.type = api.ReductionType.Average
        return s
    s = api.ReductionSpec(); s.value_channel = "value"; s.type = api.ReductionType.WeightedAverage
    g = s.glyph
    if name == "line_hl16":
        g.type = api.GlyphType.Line; g.direction_channel = "direction"; g.half_length_channel = "half_length"
        g.max_radius_cells = 18.0
    else:
        sig = "sigma4" if name == "gauss_s4" else "sigma16"
        g.type = api.GlyphType.Gaussian; g.sigma_x_channel = sig; g.sigma_y_channel = sig
        g.max_radius_cells = 32.0
    s.glyph = g
    return s


def run(impl, n_override=None):
    import shutil, tempfile
    if impl == "ours":
        from pointcloud_raster_b200 import pcr as api
        mode = api.ExecutionMode.GPU
    else:
        import oracle as orc
        api = orc.load_reference(gpu=(impl == "ref_gpu"))
        mode = api.ExecutionMode.GPU if impl == "ref_gpu" else api.ExecutionMode.CPU
    out = {}
    for name in (os.environ.get("PCR_CONFIGS", ",".join(CONFIGS)).split(",")):
        n = 5_000_000
        if impl == "ref_cpu":
            n = {"point_avg": 5_000_000, "line_hl16": 2_000_000, "gauss_s4": 200_000, "gauss_s16": 20_000}[name]
        if n_override:
            n = min(n, n_override)
        x, y, ch = arrays(n)
        b = api.BBox(); b.min_x = b.min_y = 0.0; b.max_x = b.max_y = float(GRID)
        gc = api.GridConfig(); gc.bounds = b; gc.cell_size_x = 1.0; gc.cell_size_y = -1.0; gc.compute_dimensions()
        cfg = api.PipelineConfig(); cfg.grid = gc; cfg.reductions = [make_spec(api, name)]; cfg.exec_mode = mode
        tmp = tempfile.mkdtemp(prefix="pcr_state_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        cfg.state_dir = tmp
        if impl == "ref_gpu":
            cfg.gpu_fallback_to_cpu = False
        if impl == "ours":
            cfg.gaussian_kernel = int(os.environ.get("PCR_GAUSS_KERNEL", "0"))
            cfg.staging_threads = int(os.environ.get("PCR_STAGING_THREADS", "0"))
            cfg.ring_slot_points = int(os.environ.get("PCR_RING_SLOT", "0"))
            cfg.ring_depth = int(os.environ.get("PCR_RING_DEPTH", "0"))
        p = api.Pipeline.create(cfg)
        if p is None:
            out[name] = {"error": "create failed"}; continue
        c = api.PointCloud.create(n)
        c.set_x_array(x); c.set_y_array(y)
        for k, v in ch.items():
            c.add_channel(k, api.DataType.Float32); c.set_channel_array_f32(k, v)
        times = []
        try:
            for i in range(4 if impl != "ref_cpu" else 2):
                t0 = time.perf_counter()
                p.ingest(c); p.finalize()
                times.append(time.perf_counter() - t0)
            best = min(times[1:])
            band = np.array(p.result().band_array(0))
            out[name] = {"n": n, "best_s": round(best, 6), "mpts": round(n / best / 1e6, 3),
                         "nan_frac": round(float(np.isnan(band).mean()), 6),
                         "mean": round(float(np.nanmean(band)), 6)}
        except Exception as e:   # noqa
            out[name] = {"error": str(e)[:200]}
        del p
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps({"impl": impl, "cores": os.cpu_count(), "results": out}), flush=True)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what != "all":
        run(what, int(sys.argv[2]) if len(sys.argv) > 2 else None)
        return
    table = {}
    for impl in ("ours", "ref_gpu", "ref_cpu"):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), impl], capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        table[impl] = json.loads(line[-1]) if line else {"error": (r.stderr or r.stdout)[-500:]}
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
