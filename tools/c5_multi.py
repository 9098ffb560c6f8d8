#!/usr/bin/env python
"""BASELINE config 5 at full scale: 1B clustered LiDAR-like points, Average + Max + Count fused, on a
20000 x 20000 grid (25 reference tiles, 6.4 GB of accumulator records per GPU), sharded over N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        tools/c5_multi.py [total_points]

Each rank holds total/N points as device-resident clouds (two distinct 25M-point clouds ingested
alternately), ingests them and calls finalize_device(): partial grids are merged over NVLink peer memory,
bands assembled on rank 0.  Timed with the pipeline's CUDA-event stopwatch, max over ranks; rank 0
checks that the Count band sums to the number of ingested points."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_raster_b200 import pcr                        # noqa: E402
from pointcloud_raster_b200._lib import lib                   # noqa: E402

W = 20000
TOTAL = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
CHUNK = 25_000_000


def clustered(n, seed):
    centres = np.random.default_rng(5)                        # the same 64 clusters on every rank
    K = 64
    cx, cy = centres.uniform(0, W, K), centres.uniform(0, W, K)
    sig = np.exp(centres.uniform(np.log(50), np.log(2000), K))
    rng = np.random.default_rng(seed)
    which = rng.integers(0, K, n)
    x = np.clip(rng.normal(cx[which], sig[which]), 0, W)
    y = np.clip(rng.normal(cy[which], sig[which]), 0, W)
    return x, y, (which / K + rng.normal(0, 0.05, n)).astype(np.float32)


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import faulthandler
        faulthandler.dump_traceback_later(900, exit=True)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per_rank = TOTAL // world
    n_ingests = max(1, per_rank // CHUNK)
    chunk = per_rank // n_ingests

    gc = pcr.GridConfig(); gc.bounds.min_x = gc.bounds.min_y = 0.0; gc.bounds.max_x = gc.bounds.max_y = float(W)
    gc.compute_dimensions()
    specs = []
    for t in (pcr.ReductionType.Average, pcr.ReductionType.Max, pcr.ReductionType.Count):
        s = pcr.ReductionSpec(); s.value_channel = "value"; s.type = t; specs.append(s)
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.cuda_device_id = local; cfg.async_ingest = True; cfg.comm_root_only = True
    p = pcr.Pipeline.create(cfg)
    assert p is not None
    if world > 1:
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pcr.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        p.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)

    clouds, pinned = [], []
    for r in range(min(2, n_ingests)):
        x, y, v = clustered(chunk, 1000 * rank + r)
        c = pcr.PointCloud.create(chunk); c.set_x_array(x); c.set_y_array(y)
        c.add_channel("value", pcr.DataType.Float32); c.set_channel_array_f32("value", v)
        clouds.append(c.to_device(local))
        pinned.append(c.to_pinned())
        del c, x, y, v

    def sync_all():
        p.synchronize()
        if dist is not None:
            import torch
            dist.barrier(); torch.cuda.synchronize()

    # warm-up round (allocations, IPC mapping, first-touch), then reset to an empty grid
    p.ingest(clouds[0]); p.finalize_device(); sync_all()
    p.reset(); sync_all()

    p.profile_enable(True); p.profile_reset()
    t0 = time.perf_counter()
    p.timer_begin()
    for i in range(n_ingests):
        p.ingest(clouds[i % len(clouds)])
    p.finalize_device()
    ms = p.timer_end()
    wall = time.perf_counter() - t0
    prof = p.profile_read()
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sync_all()

    # host-fed leg: the same points from pinned host memory through the ingest ring (PCIe inside the timer)
    p.reset(); sync_all()
    t0 = time.perf_counter()
    for i in range(n_ingests):
        p.ingest(pinned[i % len(pinned)])
    p.finalize_device()
    p.synchronize()
    host_ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        import torch
        t = torch.tensor([host_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        host_ms = float(t.item())
    sync_all()
    if rank == 0:
        ptr, rows, cols = p.result_band_device_ptr(2)
        cnt = np.empty((rows, cols), np.float32)
        lib.pcr_mem_copy(C.c_void_p(cnt.ctypes.data), 0, C.c_void_p(ptr), 2, cnt.nbytes, local)
        counted = float(np.nansum(cnt, dtype=np.float64))
        ingested = chunk * n_ingests * world
        print(json.dumps({"workload": "BASELINE config 5: clustered points, Average+Max+Count, 20000x20000 grid",
                          "n_gpus": world, "points": ingested, "ms": round(ms, 3), "wall_ms": round(wall * 1e3, 3),
                          "mpts_per_s": round(ingested / (ms * 1e-3) / 1e6, 1),
                          "host_fed_ms": round(host_ms, 3), "host_fed_mpts_per_s": round(ingested / (host_ms * 1e-3) / 1e6, 1),
                          "rank0_accumulate_ms": round(prof["accumulate_ms"], 3),
                          "rank0_push_ms": round(prof.get("push_ms", 0.0), 3),
                          "rank0_merge_finalize_ms": round(prof["finalize_ms"], 3),
                          "count_band_sum": counted, "count_ok": counted == float(ingested),
                          "cells_with_data": int(np.count_nonzero(~np.isnan(cnt)))}), flush=True)
    sync_all()
    del p
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
