"""The N>1 path: points sharded by rank, partial grid states combined at finalize.

CPU (gloo, world_size 2): the combine ALGORITHM — row-slice ownership from the product's own
pcr_comm_slice_rows, all-to-all of partial record slices, Op::merge in rank order, touched-tile OR,
all-gather of finalized slices — restated with the oracle's state arrays and torch.distributed, and
checked against a single-rank run.  GPU (2+ devices): the real thing, NCCL + k_finalize, vs the oracle.
"""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import ctypes as C
    import torch
    import torch.distributed as dist
    import oracle as orc
    import make_golden as mg
    from pointcloud_raster_b200 import pcr
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    o = orc.Oracle()
    w, h = 40, 37                                    # 37 rows over 2 ranks: uneven slices
    gd = orc.GridDesc(0, 0, w, h, tile_width=16, tile_height=16)
    rng = np.random.default_rng(123)
    n = 20000
    x, y = rng.uniform(-1, w + 1, n), rng.uniform(-1, h / 2, n)     # north tiles stay untouched
    v = rng.normal(0, 5, n).astype(np.float32)
    specs = [mg.Spec("v", t) for t in (orc.SUM, orc.MAX, orc.MIN, orc.AVERAGE, orc.COUNT)]
    lo, hi = rank * n // world, (rank + 1) * n // world
    _, states, touched = o.run(gd, [(x[lo:hi], y[lo:hi], {"v": v[lo:hi]})], specs, return_state=True)

    t_touched = torch.from_numpy(touched.astype(np.int32))
    dist.all_reduce(t_touched, op=dist.ReduceOp.MAX)
    touched_all = t_touched.numpy().astype(np.uint8)
    cells = w * h
    bands = []
    for s, st in zip(specs, states):
        k = o.lib.orc_state_floats(int(s.type))
        st2 = st.reshape(k, cells)
        r0, r1 = pcr.comm_slice_rows(h, world, rank)
        # every rank sends slice j of its partial state to rank j
        parts = [None] * world
        for peer in range(world):
            p0, p1 = pcr.comm_slice_rows(h, world, peer)
            send = torch.from_numpy(np.ascontiguousarray(st2[:, p0 * w:p1 * w]))
            recv = torch.empty(k, (r1 - r0) * w)
            if peer == rank:
                parts[rank] = send.numpy().copy()
                continue
            ops = [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            parts[peer] = recv.numpy()
        # owner merges the parts in RANK ORDER (Op::merge), then finalizes its slice
        acc = np.ascontiguousarray(parts[0]).reshape(-1).copy()
        sl_cells = (r1 - r0) * w
        for peer in range(1, world):
            src = np.ascontiguousarray(parts[peer]).reshape(-1)
            o.lib.orc_state_merge(int(s.type), acc.ctypes.data, src.ctypes.data, sl_cells)
        full_state = np.zeros(k * cells, np.float32).reshape(k, cells)
        full_state[:, r0 * w:r1 * w] = acc.reshape(k, sl_cells)
        out = np.empty((h, w), np.float32)
        g = o.grid(gd)
        o.lib.orc_finalize(C.byref(g), int(s.type), full_state.ctypes.data, touched_all.ctypes.data, out.ctypes.data)
        mine = torch.from_numpy(out[r0:r1].copy())
        gathered = [torch.empty((pcr.comm_slice_rows(h, world, p)[1] - pcr.comm_slice_rows(h, world, p)[0], w))
                    for p in range(world)]
        # all-gather with uneven slices: pad to the largest
        m = max(g_.shape[0] for g_ in gathered)
        padded = torch.zeros(m, w); padded[:mine.shape[0]] = mine
        outl = [torch.zeros(m, w) for _ in range(world)]
        dist.all_gather(outl, padded)
        bands.append(np.concatenate([outl[p][:gathered[p].shape[0]].numpy() for p in range(world)], axis=0))
    if rank == 0:
        np.savez(os.path.join(out_dir, "bands.npz"), *bands)
    dist.barrier()
    dist.destroy_process_group()


def test_combine_algorithm_gloo_world2(oracle):
    import torch.multiprocessing as mp
    import oracle as orc
    import make_golden as mg
    from util import compare_bands
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_gloo_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        z = np.load(os.path.join(d, "bands.npz"))
        got = [z[f"arr_{i}"] for i in range(5)]
    w, h = 40, 37
    gd = orc.GridDesc(0, 0, w, h, tile_width=16, tile_height=16)
    rng = np.random.default_rng(123)
    n = 20000
    x, y = rng.uniform(-1, w + 1, n), rng.uniform(-1, h / 2, n)
    v = rng.normal(0, 5, n).astype(np.float32)
    specs = [mg.Spec("v", t) for t in (orc.SUM, orc.MAX, orc.MIN, orc.AVERAGE, orc.COUNT)]
    ref = oracle.run(gd, [(x, y, {"v": v})], specs)
    compare_bands(oracle, gd, [(x, y, {"v": v})], specs, ref, got, "gloo world 2")
    assert np.isnan(got[0][:16]).all()               # untouched north tiles: NaN on every rank's slice


def _gloo_exchange_worker(rank, world, port, out_dir):
    """The tile-partitioned layout restated on CPU: every rank routes its own points (oracle), sends each entry
    {cell, value} to the rank that owns the entry's bin (ownership = the product's pcr_comm_partition_cells), the
    owners fold what they received; no reduce."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import torch
    import torch.distributed as dist
    import oracle as orc
    from pointcloud_raster_b200 import pcr
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    o = orc.Oracle()
    w, h = 61, 47
    gd = orc.GridDesc(0, 0, w, h, tile_width=16, tile_height=16)
    rng = np.random.default_rng(321)
    n = 30000
    x, y = rng.uniform(-1, w + 1, n), rng.uniform(-1, h + 1, n)
    v = rng.normal(0, 5, n).astype(np.float32)
    lo, hi = rank * n // world, (rank + 1) * n // world
    cell, _tile, valid = o.assign(gd, x[lo:hi], y[lo:hi])
    valid = valid.astype(bool)
    cell, vals = cell[valid].astype(np.int64), v[lo:hi][valid]
    shift, nbins, c0, c1 = pcr.comm_partition_cells(w * h, 4, world, rank, bin_cells_log2=6)
    owners = [pcr.comm_partition_cells(w * h, 4, world, r, bin_cells_log2=6)[2:] for r in range(world)]
    assert owners[0][0] == 0 and owners[-1][1] == w * h and all(a[1] == b[0] for a, b in zip(owners, owners[1:]))
    dest = np.searchsorted([a for a, _ in owners], cell, side="right") - 1            # owner of each entry's bin
    assert all(((cell[dest == r] >> shift) * (1 << shift) >= owners[r][0]).all() for r in range(world))
    # all-to-all of the entries (counts first, then payloads)
    send = [torch.from_numpy(np.stack([cell[dest == r].astype(np.float64), vals[dest == r].astype(np.float64)], 1).copy())
            for r in range(world)]
    counts = torch.tensor([len(t) for t in send])
    got_counts = torch.zeros(world, dtype=torch.long)
    dist.all_to_all_single(got_counts, counts)
    recv = [torch.empty((int(k), 2), dtype=torch.float64) for k in got_counts]
    recv[rank] = send[rank]
    ops = []
    for peer in range(world):                       # gloo has no list all_to_all: point-to-point pairs
        if peer == rank:
            continue
        ops += [dist.P2POp(dist.isend, send[peer], peer), dist.P2POp(dist.irecv, recv[peer], peer)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    ent = torch.cat(recv).numpy()
    ec, ev = ent[:, 0].astype(np.int64), ent[:, 1].astype(np.float32)
    assert ((ec >= c0) & (ec < c1)).all()                                        # only cells I own arrive here
    cnt = np.bincount(ec - c0, minlength=c1 - c0).astype(np.float32)
    mx = np.full(c1 - c0, -np.inf, np.float32)
    np.maximum.at(mx, ec - c0, ev)
    sm = np.bincount(ec - c0, weights=ev.astype(np.float64), minlength=c1 - c0)
    np.savez(os.path.join(out_dir, f"own{rank}.npz"), c0=c0, c1=c1, cnt=cnt, mx=mx, sm=sm)
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_exchange_algorithm_gloo_world2(oracle):
    import torch.multiprocessing as mp
    import oracle as orc
    import make_golden as mg
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_gloo_exchange_worker, args=(world, _free_port(), d), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, f"own{r}.npz")) for r in range(world)]
        w, h = 61, 47
        cnt = np.concatenate([p["cnt"] for p in parts]).reshape(h, w)
        mx = np.concatenate([p["mx"] for p in parts]).reshape(h, w)
        sm = np.concatenate([p["sm"] for p in parts]).reshape(h, w)
    gd = orc.GridDesc(0, 0, w, h, tile_width=16, tile_height=16)
    rng = np.random.default_rng(321)
    n = 30000
    x, y = rng.uniform(-1, w + 1, n), rng.uniform(-1, h + 1, n)
    v = rng.normal(0, 5, n).astype(np.float32)
    specs = [mg.Spec("v", t) for t in (orc.COUNT, orc.MAX, orc.SUM)]
    ref = oracle.run(gd, [(x, y, {"v": v})], specs)
    has = cnt > 0
    assert np.array_equal(np.isnan(ref[0]), ~has) and np.array_equal(ref[0][has], cnt[has])    # Count: exact
    assert np.array_equal(ref[1][has], mx[has])                                                 # Max: exact
    np.testing.assert_allclose(ref[2][has], sm[has], rtol=1e-5, atol=1e-4)


def test_partition_cells_tile_the_grid(pcr):
    for cells in (1, 17, 85581, 1_000_000, 400_000_000):
        for world in (1, 2, 3, 8):
            for log2 in (0, 6, 12):
                edges = [pcr.comm_partition_cells(cells, 4, world, r, log2) for r in range(world)]
                assert edges[0][2] == 0 and edges[-1][3] == cells
                assert all(a[3] == b[2] for a, b in zip(edges, edges[1:]))
                shift, nbins = edges[0][0], edges[0][1]
                assert nbins <= 1024 and (nbins - 1) << shift < cells <= nbins << shift
                assert all(e[2] % (1 << shift) == 0 or e[2] == cells for e in edges)   # slices start on bin boundaries


def test_slice_rows_cover_grid(pcr):
    for h in (1, 2, 7, 37, 1000, 20000):
        for world in (1, 2, 3, 4, 8):
            edges = [pcr.comm_slice_rows(h, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert all(0 <= a <= b <= h for a, b in edges)


# ---- the real N>1 path on GPUs ------------------------------------------------------------------
def _case(pcr, deterministic, big):
    """Grid, cloud and reductions shared by the workers and the checker."""
    from util import make_grid, spec
    w, h = (350, 260) if big else (300, 211)
    gc = make_grid(pcr, w, h, tile=64)
    rng = np.random.default_rng(77)
    n = 400_000
    x, y = rng.uniform(-2, w + 2, n), rng.uniform(-2, h * 0.57, n)
    ch = {"value": rng.normal(0, 3, n).astype(np.float32), "hl": rng.uniform(0, 8, n).astype(np.float32)}
    R = pcr.ReductionType
    specs = [spec(pcr, "value", t) for t in (R.Sum, R.Max, R.Min, R.Average, R.Count)]
    if not deterministic:
        specs.append(pcr.line_splat_spec("value", default_direction=0.3, half_length_channel="hl", max_radius_cells=9.0))
        specs.append(pcr.gaussian_splat_spec("value", default_sigma=1.5, max_radius_cells=5.0))
    return gc, x, y, ch, specs


def _gpu_worker(rank, world, id_path, out_dir, deterministic, comm_mode, overlapped=False, big=False):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import time
    from pointcloud_raster_b200 import pcr
    from util import make_grid, spec, cloud as mk
    if rank == 0:
        open(id_path + ".tmp", "wb").write(pcr.comm_unique_id())
        os.rename(id_path + ".tmp", id_path)
    while not os.path.exists(id_path):
        time.sleep(0.01)
    uid = open(id_path, "rb").read()
    gc, x, y, ch, specs = _case(pcr, deterministic, big)
    n = len(x)
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.cuda_device_id = rank; cfg.deterministic = deterministic; cfg.comm_mode = comm_mode
    cfg.async_ingest = overlapped
    cfg.comm_band_copy = 2 if big else 0      # `big` doubles as: ship the bands with the copy engine
    p = pcr.Pipeline.create(cfg)
    assert p is not None
    p.comm_init(uid, rank, world)
    lo, hi = rank * n // world, (rank + 1) * n // world
    if overlapped:
        # bench.py's loop: device-resident clouds, ingest + finalize_device back to back without any host
        # sync, so the merge of round r (finalize stream) runs under the ingest kernels of round r+1
        edges = np.linspace(lo, hi, 9).astype(int)
        parts = [mk(pcr, x[a:b], y[a:b], {k: v[a:b] for k, v in ch.items()}).to_device(rank)
                 for a, b in zip(edges, edges[1:])]
        for c in parts:
            p.ingest(c)
            p.finalize_device()
        p.synchronize()
        p.finalize()
    else:
        half = (lo + hi) // 2                      # two ingest+finalize rounds: the combine must be repeatable
        p.ingest(mk(pcr, x[lo:half], y[lo:half], {k: v[lo:half] for k, v in ch.items()}))
        p.finalize()
        p.ingest(mk(pcr, x[half:hi], y[half:hi], {k: v[half:hi] for k, v in ch.items()}))
        p.finalize()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), *[np.array(p.result().band_array(i)) for i in range(len(specs))])
    p.comm_barrier()


@pytest.mark.gpu
def test_multi_gpu_bands_shipped_by_copy_engine(gpu_pcr, oracle):
    test_multi_gpu_finalize_matches_oracle(gpu_pcr, oracle, False, 2, overlapped=True, big=True)


@pytest.mark.gpu
def test_multi_gpu_overlapped_steps_match_oracle(gpu_pcr, oracle):
    test_multi_gpu_finalize_matches_oracle(gpu_pcr, oracle, False, 2, overlapped=True)


@pytest.mark.gpu
@pytest.mark.parametrize("comm_mode", [1, 2], ids=["nccl", "peer"])
@pytest.mark.parametrize("deterministic", [False, True])
def test_multi_gpu_finalize_matches_oracle(gpu_pcr, oracle, deterministic, comm_mode, overlapped=False, big=False):
    if gpu_pcr.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import multiprocessing as mp
    import oracle as orc
    from util import make_grid, spec, compare_bands, grid_desc
    world = min(gpu_pcr.device_count(), 4)
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as d:
        procs = [ctx.Process(target=_gpu_worker, args=(r, world, os.path.join(d, "id"), d, deterministic, comm_mode, overlapped, big))
                 for r in range(world)]
        for pr in procs: pr.start()
        for pr in procs: pr.join(300)
        assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
        per_rank = []
        for r in range(world):
            z = np.load(os.path.join(d, f"rank{r}.npz"))
            per_rank.append([z[k] for k in sorted(z.files, key=lambda s: int(s.split("_")[1]))])
    for r in range(1, world):                          # every rank ends with the same complete bands
        for a, b in zip(per_rank[0], per_rank[r]):
            assert np.array_equal(a, b, equal_nan=True)
    pcr = gpu_pcr
    gc, x, y, ch, specs = _case(pcr, deterministic, big)
    gd = grid_desc(gc)
    ref = oracle.run(gd, [(x, y, ch)], specs)
    compare_bands(oracle, gd, [(x, y, ch)], specs, ref, per_rank[0], f"{world} GPUs", device_weights=True)


# ---- tile-partitioned grid + point exchange (comm_layout = 2) -----------------------------------------
def _part_case(pcr):
    from util import make_grid, spec
    w, h = 333, 257
    gc = make_grid(pcr, w, h, tile=64)
    rng = np.random.default_rng(91)
    n = 500_000
    x, y = rng.uniform(-2, w + 2, n), rng.uniform(-2, h * 0.7, n)           # the north tiles stay untouched
    ch = {"value": rng.normal(0, 3, n).astype(np.float32), "other": rng.uniform(0, 1, n).astype(np.float32)}
    R = pcr.ReductionType
    specs = [spec(pcr, "value", t) for t in (R.Sum, R.Max, R.Min, R.Average, R.Count)] + [spec(pcr, "other", R.Sum)]
    return gc, x, y, ch, specs


def _part_worker(rank, world, id_path, out_dir, root_only):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import time
    from pointcloud_raster_b200 import pcr
    from util import cloud as mk
    if rank == 0:
        open(id_path + ".tmp", "wb").write(pcr.comm_unique_id())
        os.rename(id_path + ".tmp", id_path)
    while not os.path.exists(id_path):
        time.sleep(0.01)
    uid = open(id_path, "rb").read()
    gc, x, y, ch, specs = _part_case(pcr)
    n = len(x)
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.cuda_device_id = rank
    cfg.point_kernel = 3; cfg.bin_cells_log2 = 9; cfg.comm_layout = 2; cfg.comm_root_only = root_only
    p = pcr.Pipeline.create(cfg)
    assert p is not None
    p.comm_init(uid, rank, world)
    # uneven shards, three ingest + finalize rounds (the exchange and the pools must be reusable), the last
    # round with two ingests before the finalize; rank 0 ingests nothing in round 1
    cut = np.linspace(0, n, 3 * world + 1).astype(int)
    for r in range(3):
        a, b = cut[r * world + rank], cut[r * world + rank + 1]
        if not (r == 1 and rank == 0):
            m = (a + b) // 2 if r == 2 else b
            p.ingest(mk(pcr, x[a:m], y[a:m], {k: v[a:m] for k, v in ch.items()}))
            if r == 2:
                p.ingest(mk(pcr, x[m:b], y[m:b], {k: v[m:b] for k, v in ch.items()}).to_device(rank))
        p.finalize()
    c0, c1 = p.owned_cells()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), np.array([c0, c1]),
             *[np.array(p.result().band_array(i)) for i in range(len(specs))])
    p.comm_barrier()


@pytest.mark.gpu
@pytest.mark.parametrize("root_only", [0, 2], ids=["bands_everywhere", "bands_distributed"])
def test_multi_gpu_partitioned_grid_point_exchange(gpu_pcr, oracle, root_only):
    if gpu_pcr.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import multiprocessing as mp
    from util import compare_bands, grid_desc
    world = min(gpu_pcr.device_count(), 4)
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as d:
        procs = [ctx.Process(target=_part_worker, args=(r, world, os.path.join(d, "id"), d, root_only)) for r in range(world)]
        for pr in procs: pr.start()
        for pr in procs: pr.join(300)
        assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
        ranks = []
        for r in range(world):
            z = np.load(os.path.join(d, f"rank{r}.npz"))
            arrs = [z[k] for k in sorted(z.files, key=lambda s: int(s.split("_")[1]))]
            ranks.append((arrs[0], arrs[1:]))
    pcr = gpu_pcr
    gc, x, y, ch, specs = _part_case(pcr)
    n = len(x)
    # points rank 0 skipped in round 1 were never ingested
    cut = np.linspace(0, n, 3 * world + 1).astype(int)
    keep = np.ones(n, bool)
    keep[cut[world]:cut[world + 1]] = False
    xs, ys, chs = x[keep], y[keep], {k: v[keep] for k, v in ch.items()}
    gd = grid_desc(gc)
    ref = oracle.run(gd, [(xs, ys, chs)], specs)
    # ownership: contiguous cell ranges that tile the grid
    edges = sorted((int(o[0]), int(o[1])) for o, _ in ranks)
    assert edges[0][0] == 0 and edges[-1][1] == gc.width * gc.height
    assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    if root_only == 0:
        for _, bands in ranks:                            # every rank ends with the same complete bands
            compare_bands(oracle, gd, [(xs, ys, chs)], specs, ref, bands, f"partitioned, {world} GPUs")
    else:                                                 # every rank holds exactly the cells it owns
        got = [np.full(gc.width * gc.height, np.nan, np.float32) for _ in specs]
        for (c0, c1), bands in ranks:
            for g, b in zip(got, bands):
                g[int(c0):int(c1)] = b.reshape(-1)[int(c0):int(c1)]
        got = [g.reshape(gc.height, gc.width) for g in got]
        compare_bands(oracle, gd, [(xs, ys, chs)], specs, ref, got, f"partitioned + distributed bands, {world} GPUs")


# ---- deterministic mode 2: the same bits on 1 and N GPUs ---------------------------------------------------
def _exact_worker(rank, world, id_path, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import time
    from pointcloud_raster_b200 import pcr
    from util import make_grid, cloud as mk
    from test_exact_mode_gpu import _glyph_specs, _cloud
    if rank == 0:
        open(id_path + ".tmp", "wb").write(pcr.comm_unique_id())
        os.rename(id_path + ".tmp", id_path)
    while not os.path.exists(id_path):
        time.sleep(0.01)
    uid = open(id_path, "rb").read()
    w, h = 160, 120
    gc = make_grid(pcr, w, h, tile=64)
    x, y, ch = _cloud(60_000, w, h, 5)
    specs = _glyph_specs(pcr)
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.cuda_device_id = rank; cfg.deterministic = 2
    p = pcr.Pipeline.create(cfg)
    assert p is not None
    p.comm_init(uid, rank, world)
    n = len(x)
    lo, hi = rank * n // world, (rank + 1) * n // world
    half = (lo + 2 * hi) // 3                          # two unequal ingest + finalize rounds
    for a, b in ((lo, half), (half, hi)):
        p.ingest(mk(pcr, x[a:b], y[a:b], {k: v[a:b] for k, v in ch.items()}))
        p.finalize()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), *[np.array(p.result().band_array(i)) for i in range(len(specs))])
    p.comm_barrier()


@pytest.mark.gpu
def test_multi_gpu_exact_mode_equals_single_gpu_bit_for_bit(gpu_pcr):
    if gpu_pcr.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import make_grid, run_product
    from test_exact_mode_gpu import _glyph_specs, _cloud
    world = min(gpu_pcr.device_count(), 4)
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as d:
        procs = [ctx.Process(target=_exact_worker, args=(r, world, os.path.join(d, "id"), d)) for r in range(world)]
        for pr in procs: pr.start()
        for pr in procs: pr.join(300)
        assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
        per_rank = []
        for r in range(world):
            z = np.load(os.path.join(d, f"rank{r}.npz"))
            per_rank.append([z[k] for k in sorted(z.files, key=lambda s: int(s.split("_")[1]))])
    pcr = gpu_pcr
    w, h = 160, 120
    gc = make_grid(pcr, w, h, tile=64)
    x, y, ch = _cloud(60_000, w, h, 5)
    single, _ = run_product(pcr, gc, [(x, y, ch)], _glyph_specs(pcr), deterministic=2)
    for r in range(world):
        for i, (a, b) in enumerate(zip(single, per_rank[r])):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"rank {r} band {i}: {world}-GPU bits differ from 1 GPU"
