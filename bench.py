#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 ingest/finalize path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
    Point glyph, Sum + Count + Max of one value channel, 5,000,000 uniform points
    (default_rng(42), U(2, 998), value U(0,1)) on a 1000 x 1000 grid, cell 1 x -1.
A "step" = Pipeline.ingest(cloud) + Pipeline.finalize().  N > 1: weak scaling — every
rank (one process per GPU) ingests its own 5M-point shard, the partial grid states are
combined at finalize over NCCL (all-to-all of row slices + fused merge/finalize kernel).

Printed JSON line (rank 0):
  value      Mpts/s with the clouds already resident in HBM (device PointCloud), results
             finalized into HBM; CUDA-event timed on the pipeline's stream, max over ranks.
             Four distinct clouds (400 MB > 126 MB L2) are rotated so no step finds its
             input in L2.
  e2e        same metric through the public API with HOST buffers: pinned host cloud ->
             ingest (H2D inside, through the staging ring) -> finalize (D2H of the bands
             inside).
  roofline   dominant kernel = fused route+accumulate; achieved = N * 20 B (x, y f64 +
             one f32 channel) / its mean launch duration measured live by
             CUDA events inside the timed region; peak = MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the reference's own CPU mode (oracle/_ref, compiled unmodified) on this
             box's host cores, same arrays; a reported baseline, not the target.
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
METRIC = "Mpts/s per glyph (Point/Line/Gauss), N=5M-1B, at 1/2/4/8 B200; % HBM peak"
WORKLOAD = "Point glyph Sum+Count+Max, 5M uniform points, 1000x1000 grid (BASELINE configs[1])"


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


def build_pipeline(pcr, device, async_ingest, rank=0, world=1, unique_id=None):
    gc = pcr.GridConfig()
    gc.bounds.min_x = gc.bounds.min_y = 0.0
    gc.bounds.max_x = gc.bounds.max_y = float(GRID)
    gc.cell_size_x, gc.cell_size_y = 1.0, -1.0
    gc.compute_dimensions()
    specs = []
    for t in (pcr.ReductionType.Sum, pcr.ReductionType.Count, pcr.ReductionType.Max):
        s = pcr.ReductionSpec()
        s.value_channel = "value"
        s.type = t
        specs.append(s)
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
    # N>1: the finished raster is assembled on rank 0 (the rank that would write the GeoTIFF);
    # the other ranks keep only their own row slice.  PCR_COMM_ROOT_ONLY=0 gives every rank all bands.
    cfg.comm_root_only = bool(int(os.environ.get("PCR_COMM_ROOT_ONLY", "1")))
    p = pcr.Pipeline.create(cfg)
    if p is None:
        raise RuntimeError("Pipeline.create failed (no CPU fallback exists)")
    if world > 1:
        p.comm_init(unique_id, rank, world)
    return p, gc, specs


def make_cloud(pcr, x, y, v, loc, device=0):
    c = pcr.PointCloud.create(len(x), loc, device)
    c.set_x_array(x)
    c.set_y_array(y)
    c.add_channel("value", pcr.DataType.Float32)
    c.set_channel_array_f32("value", v)
    return c


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    numa = bind_to_gpu_numa_node(local) if world > 1 else "n/a"
    from pointcloud_raster_b200 import pcr
    if world > 1:
        import faulthandler           # a rank that dies must not leave the others waiting for ever
        faulthandler.dump_traceback_later(int(os.environ.get("PCR_BENCH_WATCHDOG", "900")), exit=True)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def new_comm_id():
        """A fresh 128-byte NCCL id per pipeline (one communicator each), made on rank 0 and
        broadcast — torch.distributed is only the side channel, the data path is the library's."""
        if dist is None:
            return None
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pcr.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K, W = args.steps, args.warmup
    # ---------------- device-resident leg (value, roofline) ----------------
    p, gc, specs = build_pipeline(pcr, local, True, rank, world, new_comm_id())
    clouds = []
    host_sets = []
    for r in range(N_ROTATE):
        x, y, v = make_arrays(42 + 1000 * rank + r)
        if r == 0:
            host_sets.append((x, y, v))
        clouds.append(make_cloud(pcr, x, y, v, pcr.MemoryLocation.Device, local))
    toks = [p.prepare(c) for c in clouds]

    def step(i):
        p.ingest_prepared(toks[i % N_ROTATE])
        p.finalize_device()

    for i in range(W):
        step(i)
    p.synchronize()
    # kernel times for the roofline come from CUDA events inside the timed region; every PROF_EVERY-th step is
    # instrumented (two timed events around a kernel cost ~5 us of stream time and keep the next launch from
    # starting under the kernel's tail: 78 us/step with every step instrumented, 67 us with none)
    p.profile_enable(PROF_EVERY)
    p.profile_reset()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.perf_counter()
    p.timer_begin()
    for i in range(K):
        step(W + i)
    ms_dev = p.timer_end()
    t_wall = (time.perf_counter() - t_wall0) * 1e3
    barrier()
    ms_dev = max_over_ranks(ms_dev)
    prof = p.profile_read()
    p.profile_enable(False)

    # keep the sampler alive through the e2e leg too so that it sees >= a few samples under load
    # ---------------- end-to-end leg (host buffers through the public API) ----------------
    pe, _, _ = build_pipeline(pcr, local, False, rank, world, new_comm_id())
    x, y, v = host_sets[0]
    pinned = make_cloud(pcr, x, y, v, pcr.MemoryLocation.HostPinned, local)
    pageable = make_cloud(pcr, x, y, v, pcr.MemoryLocation.Host)
    ke = max(3, min(K, 20))

    def e2e_run(cloud, steps):
        for _ in range(2):
            pe.ingest(cloud); pe.finalize()
        barrier()
        pe.timer_begin()
        t0 = time.perf_counter()
        for _ in range(steps):
            pe.ingest(cloud)
            pe.finalize()
        ms = pe.timer_end()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        return max_over_ranks(max(ms, wall)) / steps      # host staging is part of e2e: take the larger clock

    e2e_ms = e2e_run(pinned, ke)
    e2e_pageable_ms = e2e_run(pageable, ke)
    clocks = sampler.stop() if sampler else None

    if dist is not None:
        barrier()
        del p, pe
        dist.destroy_process_group()
    if rank != 0:
        return

    ms_per_step = ms_dev / K
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
                       "; N>1: partial grids merged over NVLink peer memory at every finalize, bands assembled on rank 0" if world > 1 else ""), "timer": "CUDA events on the pipeline stream, max over ranks",
                   "wall_ms_per_step": round(t_wall / K, 5), "rank0_affinity": numa,
                   "kernel_timing": f"CUDA events around the kernels of every {PROF_EVERY}th step of the timed region"},
        "clocks": clocks,
        "e2e": {"value": round(total_points / (e2e_ms * 1e-3) / 1e6, 1), "unit": "Mpts/s",
                "h2d_bytes_per_step": N_POINTS * BYTES_PER_POINT, "d2h_bytes_per_step": GRID * GRID * 4 * 3,
                "ms_per_step": round(e2e_ms, 4), "steps": ke, "input": "pinned host PointCloud",
                "pageable_input_mpts": round(total_points / (e2e_pageable_ms * 1e-3) / 1e6, 1)},
        "gpu_launches": int(prof["kernel_launches"]),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                     "kernel": "k_point_direct/k_point_tma (fused route+accumulate)",
                     "algorithmic_bytes_per_launch": N_POINTS * BYTES_PER_POINT,
                     "mean_launch_ms": round(acc_ms, 5),
                     "finalize_mean_ms": round(prof["finalize_ms"] / max(1, int(prof["finalize_launches"])), 5)},
    }
    if world > 1 and int(prof.get("push_launches", 0)):
        # N>1: the slice push runs on the ingest stream; the merge/finalize above runs on its own stream under
        # the next step's ingest kernel
        out["roofline"]["push_mean_ms"] = round(prof["push_ms"] / int(prof["push_launches"]), 5)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(host_sets[0], budget_s=25.0)
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
    if orc.reference_available():
        t0 = time.perf_counter()
        _time_reference(arrays, 500_000, 0, 1, 0)
        probe = time.perf_counter() - t0
        n = int(min(N_POINTS, max(200_000, 500_000 * (150.0 / (K + W)) / max(probe, 1e-3))))
        val = _time_reference(arrays, n, 0, K, W)
        kind = "reference"
        sample = (f"each step = ingest+finalize of the first {n} of the 5M points through the unmodified "
                  f"reference (oracle/_ref), ExecutionMode.CPU, cpu_threads=0 ({cores} OpenMP threads)")
    else:
        n = N_POINTS
        val = _time_oracle_port(arrays, n, K, W)
        kind, cores = "port", 1
        sample = "each step = all 5M points through the C oracle port (oracle/_ref not built on this box)"
    ms = n / (val * 1e6) * 1e3
    out = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": "Mpts/s", "n_gpus": world,
           "steps": K, "warmup": W, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64 routing / f32 accumulate", "data": "synthetic",
           "config": {"workload": WORKLOAD, "points_per_step": n, "grid": [GRID, GRID],
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
