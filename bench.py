#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 ingest/finalize path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1], the configuration the metric is quoted on):
    Point glyph, Sum + Count + Max of one value channel, 5,000,000 uniform points
    (default_rng(42), U(2, 998), value U(0,1)) on a 1000 x 1000 grid, cell 1 x -1.
A "step" = Pipeline.ingest(cloud) + Pipeline.finalize().  N > 1: weak scaling — every rank (one
process per GPU) ingests its own 5M-point shard; what a rank accumulated since the previous
finalize is pushed to the owners of the row slices over NVLink peer memory and merged there.

Printed JSON line (rank 0):
  value      Mpts/s with the clouds already resident in HBM (device PointCloud), results finalized into
             HBM; CUDA-event timed on the pipeline's stream, max over ranks.  R windows of exactly K steps
             are timed (each bracketed by barrier + synchronize); the MEDIAN window is reported, all windows
             are listed.  Four distinct clouds (400 MB > 126 MB L2) are rotated so no step finds its input in L2.
  e2e        same metric through the public API with HOST buffers: pinned host cloud -> ingest (H2D
             inside, through the staging ring) -> finalize (D2H of the bands inside).
  roofline   dominant kernel = fused route+accumulate; achieved = N * 20 B (x, y f64 + one f32 channel;
             SURVEY §8d M3) / its mean launch duration measured live by CUDA events inside the timed
             region; peak = MEASURED_PEAKS.json hbm_gbs.  l2_red_ceiling: the same reductions without our
             kernel around them (pcr_diag_red_ceiling), measured in the same process.
  count_check   after the timed loops the Count band must sum to exactly the points ingested (all ranks).
  parity_check  (N > 1) a small sharded pipeline — Point Sum/Max/Min/Average/Count, Line, Gaussian — whose
             merged bands on rank 0 are compared with the C oracle (checker only) under tests/util's bars.
  c5         BASELINE configs[4]: 1B clustered points, Average+Max+Count, 20000 x 20000 grid, the SAME
             cloud for every N (40 chunks of 25M points, chunk j generated from seed 42+j), sharded by chunk.
  per_glyph  (N = 1) Line hl=16, Gaussian sigma=4 / sigma=16: kernel scope, e2e, algorithmic bytes, cells/s.
  ref_gpu_baseline  (N = 1) the reference's own CUDA mode compiled for sm_100 (oracle/_ref/gpu), API scope.
  cpu_baseline  the reference's own CPU mode (oracle/_ref, compiled unmodified) on this box's host cores,
             same arrays; a reported baseline, not the target.
--impl reference runs only that CPU leg, as the driver's reference arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_POINTS = 5_000_000
GRID = 1000
BYTES_PER_POINT = 20          # x f64 + y f64 + value f32 (Sum/Count/Max fused on one channel)
N_ROTATE = 4                  # distinct device clouds: 4 * 100 MB > L2
PROF_EVERY = int(os.environ.get("PCR_PROF_EVERY", "5"))   # kernel-timing events on every 5th step (coprime to N_ROTATE)
N_WINDOWS = int(os.environ.get("PCR_BENCH_WINDOWS", "5"))
METRIC = "Mpts/s per glyph (Point/Line/Gauss), N=5M-1B, at 1/2/4/8 B200; % HBM peak"
WORKLOAD = "Point glyph Sum+Count+Max, 5M uniform points, 1000x1000 grid (BASELINE configs[1])"

C5_GRID = 20000
C5_TOTAL = int(float(os.environ.get("PCR_C5_POINTS", "1e9")))
C5_CHUNK = 25_000_000


def make_arrays(seed, n=N_POINTS):
    """scripts/benchmarks/benchmark_glyph_full.py:62-78 of the reference."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(2, GRID - 2, n)
    y = rng.uniform(2, GRID - 2, n)
    v = rng.uniform(0, 1, n).astype(np.float32)
    return x, y, v


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.path = tempfile.mktemp(prefix="pcr_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={device}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def bind_to_gpu_numa_node(device):
    """Best effort: run this rank (and first-touch its pinned staging memory) on the NUMA node the GPU
    hangs off, so H2D traffic does not cross the socket interconnect.  Returns a short description."""
    try:
        bus = subprocess.check_output(["nvidia-smi", f"--id={device}", "--query-gpu=pci.bus_id",
                                       "--format=csv,noheader"], text=True).strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus               # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "numa: single node"
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return f"numa node {node} ({len(ids)} cpus)"
    except Exception as e:            # noqa
        return f"numa: unbound ({type(e).__name__})"
    return "numa: unbound"


def square_grid(pcr, w):
    gc = pcr.GridConfig()
    gc.bounds.min_x = gc.bounds.min_y = 0.0
    gc.bounds.max_x = gc.bounds.max_y = float(w)
    gc.cell_size_x, gc.cell_size_y = 1.0, -1.0
    gc.compute_dimensions()
    return gc


def point_specs(pcr, types):
    specs = []
    for t in types:
        s = pcr.ReductionSpec()
        s.value_channel = "value"
        s.type = t
        specs.append(s)
    return specs


def build_pipeline(pcr, device, async_ingest, rank=0, world=1, unique_id=None, grid=GRID, specs=None, root_only=None):
    gc = square_grid(pcr, grid)
    if specs is None:
        specs = point_specs(pcr, (pcr.ReductionType.Sum, pcr.ReductionType.Count, pcr.ReductionType.Max))
    cfg = pcr.PipelineConfig()
    cfg.grid = gc
    cfg.reductions = specs
    cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.cuda_device_id = device
    cfg.async_ingest = async_ingest
    cfg.point_kernel = int(os.environ.get("PCR_POINT_KERNEL", "0"))
    cfg.warp_aggregate = int(os.environ.get("PCR_WARP_AGG", "0"))
    cfg.comm_mode = int(os.environ.get("PCR_COMM_MODE", "0"))
    cfg.ring_slot_points = int(os.environ.get("PCR_RING_SLOT", "0"))
    cfg.ring_depth = int(os.environ.get("PCR_RING_DEPTH", "0"))
    cfg.staging_threads = int(os.environ.get("PCR_STAGING_THREADS", "0"))
    cfg.bin_cells_log2 = int(os.environ.get("PCR_BIN_LOG2", "0"))
    cfg.bin_pool_points = int(float(os.environ.get("PCR_BIN_POOL", "0")))
    cfg.comm_layout = int(os.environ.get("PCR_COMM_LAYOUT", "0"))
    # N>1: the finished raster is assembled on rank 0 (the rank that would write the GeoTIFF);
    # the other ranks keep only their own row slice.  PCR_COMM_ROOT_ONLY=0 gives every rank all bands.
    cfg.comm_root_only = int(os.environ.get("PCR_COMM_ROOT_ONLY", "1")) if root_only is None else root_only
    p = pcr.Pipeline.create(cfg)
    if p is None:
        raise RuntimeError("Pipeline.create failed (no CPU fallback exists)")
    if world > 1:
        p.comm_init(unique_id, rank, world)
    return p, gc, specs


def make_cloud(pcr, x, y, chans, loc, device=0):
    if not isinstance(chans, dict):
        chans = {"value": chans}
    c = pcr.PointCloud.create(len(x), loc, device)
    c.set_x_array(x)
    c.set_y_array(y)
    for k, v in chans.items():
        c.add_channel(k, pcr.DataType.Float32)
        c.set_channel_array_f32(k, v)
    return c


class Dist:
    """torch.distributed as the side channel of an N>1 run (rendezvous, barrier, max over ranks, the 128-byte
    NCCL id of each pipeline); the data path is the library's own."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1:
            import faulthandler           # a rank that dies must not leave the others waiting for ever
            faulthandler.dump_traceback_later(int(os.environ.get("PCR_BENCH_WATCHDOG", "700")), exit=True)
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def new_comm_id(self, pcr):
        if self.dist is None:
            return None
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            idt = torch.frombuffer(bytearray(pcr.comm_unique_id()), dtype=torch.uint8).cuda()
        self.dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    def barrier(self):
        if self.dist is not None:
            import torch
            self.dist.barrier()
            torch.cuda.synchronize()

    def reduce(self, v, op="max"):
        if self.dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.dist is not None:
            self.barrier()
            self.dist.destroy_process_group()


def band_to_host(pcr, p, band, device, cell0=None, cell1=None):
    """D2H of (a row-major cell range of) a finalized band left in HBM by finalize_device(); flat array."""
    import ctypes as C
    from pointcloud_raster_b200._lib import lib
    ptr, rows, cols = p.result_band_device_ptr(band)
    cell0, cell1 = (0, rows * cols) if cell0 is None else (cell0, cell1)
    out = np.empty(cell1 - cell0, np.float32)
    if out.size:
        lib.pcr_mem_copy(C.c_void_p(out.ctypes.data), 0, C.c_void_p(ptr + cell0 * 4), 2, out.nbytes, device)
    return out


# ---------------------------------------------------------------------------
# legs of the product arm
# ---------------------------------------------------------------------------
def leg_headline(pcr, D, K, W):
    """Device-resident config 2: R windows of K steps; kernel times by CUDA events on every PROF_EVERY-th step."""
    p, gc, specs = build_pipeline(pcr, D.local, True, D.rank, D.world, D.new_comm_id(pcr))
    clouds, host0 = [], None
    for r in range(N_ROTATE):
        x, y, v = make_arrays(42 + 1000 * D.rank + r)
        if r == 0:
            host0 = (x, y, v)
        clouds.append(make_cloud(pcr, x, y, v, pcr.MemoryLocation.Device, D.local))
    toks = [p.prepare(c) for c in clouds]
    steps_done = 0

    def step(i):
        p.ingest_prepared(toks[i % N_ROTATE])
        p.finalize_device()

    for i in range(W):
        step(steps_done); steps_done += 1
    p.synchronize()
    # kernel times for the roofline come from CUDA events inside the timed region; every PROF_EVERY-th step is
    # instrumented (two timed events around a kernel cost ~5 us of stream time and keep the next launch from
    # starting under the kernel's tail: 78 us/step with every step instrumented, 67 us with none)
    p.profile_enable(PROF_EVERY)
    p.profile_reset()
    windows, walls = [], []
    for _ in range(N_WINDOWS):
        D.barrier()
        t0 = time.perf_counter()
        p.timer_begin()
        for _k in range(K):
            step(steps_done); steps_done += 1
        ms = p.timer_end()
        walls.append((time.perf_counter() - t0) * 1e3)
        D.barrier()
        windows.append(D.reduce(ms))
    prof = p.profile_read()
    p.profile_enable(False)
    # exact invariant of the whole run, N ranks included: every point is inside the grid, so the merged Count
    # band sums to the number of points ingested so far
    p.finalize()
    got = None
    if D.rank == 0:
        got = float(np.nansum(np.asarray(p.result().band_array(1)), dtype=np.float64))
    expect = float(N_POINTS) * D.world * steps_done
    count_check = {"expected": expect, "got": got, "ok": got == expect,
                   "what": "sum of the merged Count band after all warm-up + timed steps (rank 0)"}
    return p, host0, windows, walls, prof, count_check


def leg_e2e(pcr, D, host0, K):
    pe, _, _ = build_pipeline(pcr, D.local, False, D.rank, D.world, D.new_comm_id(pcr))
    x, y, v = host0
    pinned = make_cloud(pcr, x, y, v, pcr.MemoryLocation.HostPinned, D.local)
    pageable = make_cloud(pcr, x, y, v, pcr.MemoryLocation.Host)
    ke = max(3, min(K, 20))

    def run(cloud, steps):
        for _ in range(2):
            pe.ingest(cloud); pe.finalize()
        D.barrier()
        pe.timer_begin()
        t0 = time.perf_counter()
        for _ in range(steps):
            pe.ingest(cloud)
            pe.finalize()
        ms = pe.timer_end()
        wall = (time.perf_counter() - t0) * 1e3
        D.barrier()
        return D.reduce(max(ms, wall)) / steps      # host staging is part of e2e: take the larger clock

    e2e_ms = run(pinned, ke)
    pageable_ms = run(pageable, ke)
    D.barrier()
    del pe
    return e2e_ms, pageable_ms, ke


def leg_parity(pcr, D):
    """N>1: a small sharded pipeline checked against the C oracle on rank 0 (oracle = checker only)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import make_grid, spec, cloud as mk, compare_bands, grid_desc
    w, h = 300, 211
    gc = make_grid(pcr, w, h, tile=64)
    rng = np.random.default_rng(77)
    n = 400_000
    x, y = rng.uniform(-2, w + 2, n), rng.uniform(-2, h * 0.57, n)            # the north tiles stay untouched
    ch = {"value": rng.normal(0, 3, n).astype(np.float32), "hl": rng.uniform(0, 8, n).astype(np.float32)}
    R = pcr.ReductionType
    specs = [spec(pcr, "value", t) for t in (R.Sum, R.Max, R.Min, R.Average, R.Count)]
    specs.append(pcr.line_splat_spec("value", default_direction=0.3, half_length_channel="hl", max_radius_cells=9.0))
    specs.append(pcr.gaussian_splat_spec("value", default_sigma=1.5, max_radius_cells=5.0))
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.cuda_device_id = D.local
    cfg.comm_mode = int(os.environ.get("PCR_COMM_MODE", "0"))
    p = pcr.Pipeline.create(cfg)
    if p is None:
        raise RuntimeError("Pipeline.create failed")
    p.comm_init(D.new_comm_id(pcr), D.rank, D.world)
    lo, hi = D.rank * n // D.world, (D.rank + 1) * n // D.world
    edges = np.linspace(lo, hi, 4).astype(int)                                # three ingest+finalize rounds
    for a, b in zip(edges, edges[1:]):
        p.ingest(mk(pcr, x[a:b], y[a:b], {k: v[a:b] for k, v in ch.items()}))
        p.finalize()
    out = None
    if D.rank == 0:
        import oracle as orc
        got = [np.array(p.result().band_array(i)) for i in range(len(specs))]
        o = orc.Oracle()
        gd = grid_desc(gc)
        ref = o.run(gd, [(x, y, ch)], specs)
        ok, msg = True, ""
        try:
            compare_bands(o, gd, [(x, y, ch)], specs, ref, got, f"{D.world} GPUs", device_weights=True)
        except AssertionError as e:
            ok, msg = False, str(e)[:300]
        rel = 0.0
        for a, b in zip(got, ref):
            m = np.isfinite(a) & np.isfinite(b)
            if m.any():
                rel = max(rel, float(np.max(np.abs(a[m].astype(np.float64) - b[m]) / np.maximum(np.abs(b[m]), 1e-3))))
        out = {"ok": ok, "bands": len(specs), "max_rel": rel, "nan_masks_equal": all(
            np.array_equal(np.isnan(a), np.isnan(b)) for a, b in zip(got, ref)),
            "exact_bands": "Max, Min, Count bit-exact; Sum/Average/Line/Gaussian within tests/util.py bounds",
            "case": f"{n} points sharded over {D.world} ranks, 3 ingest+finalize rounds, {w}x{h} grid, 64-cell tiles",
            "checker": "oracle/libpcr_oracle.so (C restatement of the reference)"}
        if msg:
            out["error"] = msg
    p.comm_barrier()
    D.barrier()
    del p
    return out


def c5_chunk(torch, j, device):
    """Chunk j of the config-5 cloud, generated in HBM: K=64 cluster centres ~U(0,W), per-point N(centre, sigma_c)
    with sigma_c log-uniform in [50, 2000] cells, clipped to the bbox (mass exactly on max_x / min_y), value =
    cluster-id ramp + noise (after generate_gaussian_clusters, python/pcr/test_generators.py:560-632, streamed in
    chunks as scripts/benchmarks/benchmark_billion_points.py:166-218 does).  The cloud depends on j only."""
    cen = np.random.default_rng(5)
    K = 64
    cx = torch.tensor(cen.uniform(0, C5_GRID, K), dtype=torch.float64, device=device)
    cy = torch.tensor(cen.uniform(0, C5_GRID, K), dtype=torch.float64, device=device)
    sig = torch.tensor(np.exp(cen.uniform(np.log(50), np.log(2000), K)), dtype=torch.float64, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(42 + j)
    which = torch.randint(0, K, (C5_CHUNK,), generator=g, device=device)
    if os.environ.get("PCR_C5_DIST") == "uniform":      # development aid: no hot cells
        x = torch.rand(C5_CHUNK, generator=g, device=device, dtype=torch.float64) * C5_GRID
        y = torch.rand(C5_CHUNK, generator=g, device=device, dtype=torch.float64) * C5_GRID
        return x, y, torch.rand(C5_CHUNK, generator=g, device=device, dtype=torch.float32)
    x = (cx[which] + sig[which] * torch.randn(C5_CHUNK, generator=g, device=device, dtype=torch.float64)).clamp_(0.0, float(C5_GRID))
    y = (cy[which] + sig[which] * torch.randn(C5_CHUNK, generator=g, device=device, dtype=torch.float64)).clamp_(0.0, float(C5_GRID))
    v = (which.to(torch.float32) / K + 0.05 * torch.randn(C5_CHUNK, generator=g, device=device, dtype=torch.float32))
    return x.contiguous(), y.contiguous(), v.contiguous()


def leg_c5(pcr, D):
    import torch
    dev = torch.device("cuda", D.local)
    n_chunks = max(D.world, C5_TOTAL // C5_CHUNK)
    mine = [j for j in range(n_chunks) if j % D.world == D.rank]
    chunks = [c5_chunk(torch, j, dev) for j in mine]
    torch.cuda.synchronize(dev)
    R = pcr.ReductionType
    p, gc, specs = build_pipeline(pcr, D.local, True, D.rank, D.world, D.new_comm_id(pcr), grid=C5_GRID,
                                  specs=point_specs(pcr, (R.Average, R.Max, R.Count)), root_only=2)

    def ingest(c):
        p.ingest_arrays(c[0].data_ptr(), c[1].data_ptr(), C5_CHUNK, {"value": c[2].data_ptr()})

    def sync_all():
        p.synchronize()
        D.barrier()

    ingest(chunks[0]); p.finalize_device(); sync_all()        # warm-up: allocations, first touch, IPC
    p.reset(); sync_all()
    p.profile_enable(True); p.profile_reset()
    D.barrier()
    t0 = time.perf_counter()
    p.timer_begin()
    for c in chunks:
        ingest(c)
    p.finalize_device()
    ms = D.reduce(p.timer_end())
    wall = (time.perf_counter() - t0) * 1e3
    prof = p.profile_read()
    p.profile_enable(False)
    sync_all()
    # checks on the distributed bands: every rank sums the row slice it owns
    r0, r1 = p.owned_cells()
    cnt = band_to_host(pcr, p, 2, D.local, r0, r1)
    count_sum = D.reduce(float(np.nansum(cnt, dtype=np.float64)), "sum")
    cells_with_data = D.reduce(float(np.count_nonzero(~np.isnan(cnt))), "sum")
    del cnt
    mx = band_to_host(pcr, p, 1, D.local, r0, r1)
    max_checksum = D.reduce(float(np.nansum(mx, dtype=np.float64)), "sum")
    del mx
    tiles = p.stats().tiles_active
    points = C5_CHUNK * n_chunks
    # host-fed: this rank's first chunk from pinned host memory through the ingest ring (PCIe inside the timer)
    hx, hy, hv = (t.cpu().numpy() for t in chunks[0])
    pinned = make_cloud(pcr, hx, hy, hv, pcr.MemoryLocation.HostPinned, D.local)
    del hx, hy, hv
    p.ingest(pinned); p.finalize_device(); sync_all()
    t0 = time.perf_counter()
    p.ingest(pinned); p.finalize_device(); p.synchronize()
    host_ms = D.reduce((time.perf_counter() - t0) * 1e3)
    sync_all()
    peak, _ = measured_peak()
    out = {"workload": "BASELINE configs[4]: clustered LiDAR-like points, Average+Max+Count fused, 20000x20000 grid "
                       "(6.4 GB of 16-byte records), points sharded by 25M-point chunk over the ranks",
           "points": points, "chunks_per_rank": len(mine), "n_gpus": D.world,
           "ms": round(ms, 3), "wall_ms": round(wall, 3), "mpts_per_s": round(points / (ms * 1e-3) / 1e6, 1),
           "scaling": "strong (the same cloud for every N: chunk j depends on j only)",
           "step": "ingest of every chunk (device-resident) + one finalize_device()" + (
               "; tile-partitioned grid: every rank owns a contiguous range of bins and keeps records for those cells "
               "only, the binning kernel appends each point's 8-byte entry to its owner's pool over NVLink peer memory "
               "(the all-to-all), no reduce at finalize; bands stay distributed (every rank holds the cells it owns)"
               if D.world > 1 and int(os.environ.get("PCR_COMM_LAYOUT", "0")) != 1 else
               "; replicated partial grids merged at finalize, bands distributed" if D.world > 1 else ""),
           "rank0_accumulate_ms": round(prof["accumulate_ms"], 3), "rank0_sort_or_bin_ms": round(prof["sort_ms"], 3),
           "rank0_push_ms": round(prof.get("push_ms", 0.0), 3), "rank0_merge_finalize_ms": round(prof["finalize_ms"], 3),
           "hbm_frac_algorithmic": round(points * BYTES_PER_POINT / (ms * 1e-3) / 1e9 / (peak * D.world), 4),
           "count_band_sum": count_sum, "count_ok": count_sum == float(points),
           "max_band_checksum": round(max_checksum, 2), "cells_with_data": int(cells_with_data), "tiles_active_rank0": int(tiles),
           "host_fed": {"points_per_rank": C5_CHUNK, "ms": round(host_ms, 3),
                        "mpts_per_s": round(C5_CHUNK * D.world / (host_ms * 1e-3) / 1e6, 1),
                        "what": "one 25M-point chunk per rank from pinned host memory: ingest + finalize_device, wall clock"}}
    D.barrier()
    del p, chunks
    torch.cuda.empty_cache()
    return out


def leg_per_glyph(pcr, device):
    """Kernel scope + e2e for the glyph configs of BASELINE.json (configs[2], configs[3]) on one GPU."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import glyph_bench as gb
    x, y, ch = gb.arrays(N_POINTS)
    peak, _ = measured_peak()
    out = {}
    table = {"point_avg": (20, 1, ["value"]), "line_hl16": (28, None, ["value", "direction", "half_length"]),
             "gauss_s4": (24, 625, ["value", "sigma4"]), "gauss_s16": (24, 4225, ["value", "sigma16"])}
    for name, (bpp, cells_nominal, used) in table.items():
        spec = gb.make_spec(pcr, name)
        specs = [spec]
        if name == "line_hl16":                    # a Count band on the same glyph shares the weight word: its
            import copy                            # sum is the number of painted cells, at no extra cost
            cs = copy.deepcopy(spec); cs.type = pcr.ReductionType.Count
            specs.append(cs)
        chans = {k: ch[k] for k in used}
        p, gc, _ = build_pipeline(pcr, device, True, specs=specs)
        dcloud = make_cloud(pcr, x, y, chans, pcr.MemoryLocation.Device, device)
        tok = p.prepare(dcloud)
        for _ in range(2):
            p.ingest_prepared(tok); p.finalize_device()
        p.synchronize()
        p.profile_enable(True); p.profile_reset()
        steps = 5
        p.timer_begin()
        for _ in range(steps):
            p.ingest_prepared(tok); p.finalize_device()
        ms = p.timer_end() / steps
        prof = p.profile_read()
        p.profile_enable(False)
        acc = (prof["accumulate_ms"] + prof["sort_ms"]) / max(1, int(prof["accumulate_launches"]))
        painted = None
        if name == "line_hl16":
            p.reset()
            p.ingest_prepared(tok); p.finalize()
            painted = float(np.nansum(np.asarray(p.result().band_array(1)), dtype=np.float64))
        elif cells_nominal:
            painted = float(N_POINTS) * cells_nominal
        del p
        pe, _, _ = build_pipeline(pcr, device, False, specs=[spec])
        pinned = make_cloud(pcr, x, y, chans, pcr.MemoryLocation.HostPinned, device)
        pe.ingest(pinned); pe.finalize()
        t0 = time.perf_counter()
        for _ in range(3):
            pe.ingest(pinned); pe.finalize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / 3
        del pe, pinned, dcloud
        out[name] = {"kernel_scope_mpts": round(N_POINTS / (ms * 1e-3) / 1e6, 1), "ms_per_step": round(ms, 4),
                     "accumulate_ms": round(acc, 4),
                     "e2e_mpts": round(N_POINTS / (e2e_ms * 1e-3) / 1e6, 1), "e2e_ms": round(e2e_ms, 3),
                     "algorithmic_bytes_per_point": bpp,
                     "hbm_frac": round(N_POINTS * bpp / (acc * 1e-3) / 1e9 / peak, 4),
                     "cells_painted": painted,
                     "gcells_per_s": round(painted / (acc * 1e-3) / 1e9, 2) if painted else None}
    out["_scope"] = ("5M points, 1000x1000 grid; kernel scope = device-resident ingest + finalize_device, CUDA events; "
                     "e2e = pinned host cloud -> ingest -> finalize -> host band, wall clock; limiting unit: Point and Line "
                     "the L2 reduction-request rate (profiles/), Gaussian (per-bin GEMM, k_gauss_binmma) tensor pipe + table construction")
    return out


def leg_api_scope(impl, timeout=600):
    """tools/glyph_bench.py in its own process: pageable numpy in -> ingest+finalize -> host band, best of 3,
    exactly how the reference's benchmark_glyph_full.py:80-100 times itself."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "glyph_bench.py"), impl],
                           capture_output=True, text=True, timeout=timeout)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if not line:
            return {"unavailable": (r.stderr or r.stdout)[-300:]}
        res = json.loads(line[-1])["results"]
        return {k: (v.get("mpts") if "mpts" in v else v) for k, v in res.items()}
    except Exception as e:   # noqa
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def run_ours(args):
    D = Dist()
    rank, world, local = D.rank, D.world, D.local
    numa = bind_to_gpu_numa_node(local) if world > 1 else "n/a"
    from pointcloud_raster_b200 import pcr
    K, W = args.steps, args.warmup
    if args.only_c5:
        c5 = leg_c5(pcr, D)
        D.close()
        if rank == 0:
            print(json.dumps({"c5": c5}), flush=True)
        return
    sampler = ClockSampler(local) if rank == 0 else None

    p, host0, windows, walls, prof, count_check = leg_headline(pcr, D, K, W)
    e2e_ms, e2e_pageable_ms, ke = leg_e2e(pcr, D, host0, K)
    clocks = sampler.stop() if sampler else None
    D.barrier()
    del p

    parity = leg_parity(pcr, D) if world > 1 and not args.skip_parity else None
    c5 = None
    if not args.skip_c5:
        try:
            c5 = leg_c5(pcr, D)
        except Exception as e:   # noqa  (the headline numbers must survive a failure of this leg)
            c5 = {"error": f"{type(e).__name__}: {e}"[:300]}
    D.close()
    if rank != 0:
        return

    med = statistics.median(windows)
    ms_per_step = med / K
    total_points = N_POINTS * world
    value = total_points / (ms_per_step * 1e-3) / 1e6
    peak, peak_src = measured_peak()
    acc_launches = max(1, int(prof["accumulate_launches"]))
    acc_ms = prof["accumulate_ms"] / acc_launches
    achieved = N_POINTS * BYTES_PER_POINT / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "point_kernel_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    out = {
        "metric": METRIC, "value": round(value, 1), "unit": "Mpts/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 routing / f32 accumulate", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points_per_gpu": N_POINTS, "grid": [GRID, GRID],
                   "reductions": ["Sum", "Count", "Max"], "glyph": "Point",
                   "l2_policy": f"{N_ROTATE} distinct device clouds rotated (400 MB > L2), no step re-reads a resident input",
                   "step": "ingest(device cloud) + finalize_device()" + (
                       "; N>1: every finalize pushes the records accumulated since the previous one to the owners of the row "
                       "slices over NVLink peer memory (delta epochs, double-buffered), merged there, bands assembled on rank 0"
                       if world > 1 else ""),
                   "timer": f"CUDA events on the pipeline stream, max over ranks; {N_WINDOWS} windows of exactly {K} steps, "
                            "each bracketed by barrier + synchronize; value = median window",
                   "windows_ms_per_step": [round(w / K, 5) for w in windows],
                   "wall_ms_per_step_rank0": round(statistics.median(walls) / K, 5), "rank0_affinity": numa,
                   "kernel_timing": f"CUDA events around the kernels of every {PROF_EVERY}th step of the timed region"},
        "clocks": clocks,
        "e2e": {"value": round(total_points / (e2e_ms * 1e-3) / 1e6, 1), "unit": "Mpts/s",
                "h2d_bytes_per_step": N_POINTS * BYTES_PER_POINT, "d2h_bytes_per_step": GRID * GRID * 4 * 3,
                "ms_per_step": round(e2e_ms, 4), "steps": ke, "input": "pinned host PointCloud",
                "pageable_input_mpts": round(total_points / (e2e_pageable_ms * 1e-3) / 1e6, 1)},
        "gpu_launches": int(prof["kernel_launches"]),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                     "kernel": "k_point_direct<2,1,0,1,1> (fused route+accumulate)",
                     "algorithmic_bytes_per_launch": N_POINTS * BYTES_PER_POINT,
                     "mean_launch_ms": round(acc_ms, 5),
                     "finalize_mean_ms": round(prof["finalize_ms"] / max(1, int(prof["finalize_launches"])), 5)},
        "count_check": count_check,
    }
    if world > 1 and int(prof.get("push_launches", 0)):
        # N>1: push and merge/finalize run on the finalize stream, under the next step's ingest kernel
        out["roofline"]["push_mean_ms"] = round(prof["push_ms"] / int(prof["push_launches"]), 5)
    if world == 1:
        try:
            reds = pcr.diag_red_ceiling(N_POINTS, GRID * GRID, False, 20, local)
            both = pcr.diag_red_ceiling(N_POINTS, GRID * GRID, True, 20, local)
            out["roofline"]["l2_red_ceiling"] = {
                "reds_only_us": round(reds, 2), "loads_plus_reds_us": round(both, 2),
                "kernel_us": round(acc_ms * 1e3, 2), "frac_of_reds_only": round(reds / (acc_ms * 1e3), 4),
                "frac_of_loads_plus_reds": round(both / (acc_ms * 1e3), 4),
                "ceiling_gbs_at_20B": round(N_POINTS * BYTES_PER_POINT / (reds * 1e-6) / 1e9, 1),
                "what": "pcr_diag_red_ceiling, same process: 5M hashed cells into 1M 16-byte records with one "
                        "red.global.add.v2.f32 + one red.global.max.s32 per point (the kernel's own reductions), "
                        "without / with the kernel's three streaming loads per point"}
        except Exception as e:   # noqa
            out["roofline"]["l2_red_ceiling"] = {"error": str(e)[:200]}
    if parity is not None:
        out["parity_check"] = parity
    if c5 is not None:
        out["c5"] = c5
    if world == 1 and not args.skip_glyphs:
        try:
            out["per_glyph"] = leg_per_glyph(pcr, local)
        except Exception as e:   # noqa
            out["per_glyph"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        ours_api = leg_api_scope("ours")
        ref_gpu = leg_api_scope("ref_gpu")
        ratio = {}
        for k, v in ours_api.items():
            if isinstance(v, (int, float)) and isinstance(ref_gpu.get(k), (int, float)) and ref_gpu[k] > 0:
                ratio[k] = round(v / ref_gpu[k], 1)
        out["ref_gpu_baseline"] = {
            "kind": "the reference's own CUDA mode compiled unmodified for sm_100 (oracle/_ref/gpu), same B200, "
                    "API scope: pageable numpy arrays -> ingest + finalize -> host band, best of 3 after a warm-up "
                    "(benchmark_glyph_full.py:80-100), 5M points, 1000x1000",
            "reference_gpu_mpts": ref_gpu, "ours_api_scope_mpts": ours_api, "ratio": ratio}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(host0, budget_s=25.0)
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------
# CPU legs: the reference's own CPU mode (oracle/_ref), else the C oracle port
# ---------------------------------------------------------------------------
class _Spec:
    def __init__(self, t):
        import make_golden as mg
        self.value_channel, self.type, self.output_band_name = "value", t, ""
        self.glyph = mg.Glyph()


def _ref_pipeline(threads):
    import oracle as orc
    ref = orc.load_reference()
    gd = orc.GridDesc(0, 0, GRID, GRID)
    cfg = ref.PipelineConfig()
    cfg.grid = orc.reference_grid(ref, gd)
    cfg.reductions = [orc.to_reference_spec(ref, _Spec(t)) for t in (0, 5, 1)]
    cfg.exec_mode = ref.ExecutionMode.CPU
    cfg.cpu_threads = threads
    tmp = tempfile.mkdtemp(prefix="pcr_ref_state_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    cfg.state_dir = tmp
    p = ref.Pipeline.create(cfg)
    if p is None:
        raise RuntimeError("reference Pipeline.create failed")
    return ref, p, tmp


def _time_reference(arrays, n, threads, steps, warmup):
    """Mpts/s of the unmodified reference, CPU mode, ingest+finalize per step."""
    import shutil
    import oracle as orc
    ref, p, tmp = _ref_pipeline(threads)
    try:
        x, y, v = (a[:n] for a in arrays)
        cloud = orc.reference_cloud(ref, x, y, {"value": v})
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            p.ingest(cloud)
            p.finalize()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        return n / (sum(times) / len(times)) / 1e6
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def _time_oracle_port(arrays, n, steps, warmup):
    import oracle as orc
    o = orc.Oracle()
    gd = orc.GridDesc(0, 0, GRID, GRID)
    x, y, v = (a[:n] for a in arrays)
    specs = [_Spec(t) for t in (0, 5, 1)]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        o.run(gd, [(x, y, {"value": v})], specs)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return n / (sum(times) / len(times)) / 1e6


def cpu_baseline(arrays, budget_s=25.0):
    import oracle as orc
    cores = os.cpu_count() or 1
    if orc.reference_available():
        n = 1_000_000
        t0 = time.perf_counter()
        _time_reference(arrays, n, 0, 1, 0)                 # probe: how slow is this box?
        probe = time.perf_counter() - t0
        n = int(min(N_POINTS, max(1_000_000, n * (budget_s / 4.0) / max(probe, 1e-3))))
        all_cores = _time_reference(arrays, n, 0, 1, 1)
        one = _time_reference(arrays, n, 1, 1, 1)
        return {"value": round(all_cores, 3), "unit": "Mpts/s", "cores": cores, "kind": "reference",
                "sample": f"first {n} of the 5M points, Sum+Count+Max (3 reductions, routed 3x as upstream does), "
                          f"ingest+finalize, 1 warm-up + 1 timed, cpu_threads=0 (OpenMP default = {cores} threads)",
                "value_cpu_threads_1": round(one, 3)}
    n = N_POINTS
    return {"value": round(_time_oracle_port(arrays, n, 2, 1), 3), "unit": "Mpts/s", "cores": 1, "kind": "port",
            "sample": "all 5M points, C oracle (scalar, 1 thread), 1 warm-up + 2 timed"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as orc
    world = int(os.environ.get("WORLD_SIZE", "1"))
    arrays = make_arrays(42)
    K, W = args.steps, args.warmup
    cores = os.cpu_count() or 1
    n = N_POINTS                       # always the full 5M-point cloud of the product arm's config
    if orc.reference_available():
        val = _time_reference(arrays, n, 0, K, W)
        kind = "reference"
        sample = (f"each step = ingest+finalize of all {n} points through the unmodified "
                  f"reference (oracle/_ref), ExecutionMode.CPU, cpu_threads=0 ({cores} OpenMP threads)")
    else:
        val = _time_oracle_port(arrays, n, K, W)
        kind, cores = "port", 1
        sample = "each step = all 5M points through the C oracle port (oracle/_ref not built on this box)"
    ms = n / (val * 1e6) * 1e3
    out = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": "Mpts/s", "n_gpus": world,
           "steps": K, "warmup": W, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64 routing / f32 accumulate", "data": "synthetic",
           "config": {"workload": WORKLOAD, "points_per_gpu": n, "grid": [GRID, GRID],
                      "reductions": ["Sum", "Count", "Max"], "glyph": "Point"},
           "cpu_baseline": {"value": round(val, 3), "unit": "Mpts/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": round(val, 3), "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    if os.environ.get("PCR_BENCH_DEBUG"):        # where is a hung rank? dump all Python stacks and exit
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["PCR_BENCH_DEBUG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-c5", action="store_true")
    ap.add_argument("--skip-glyphs", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--only-c5", action="store_true", help="development aid: run the config-5 leg alone")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
