"""GPU parity tests of the tile-binned Point path (point_kernel=3; csrc/bin_kernels.cu): the entries that
ingest appends to per-bin page chains, folded bin by bin at finalize, must give exactly what the direct
kernel and the oracle give — Count/Max/Min bit-exact, sums within the stated fp32 bound — for any bin size,
pool size (early folds), ingest pattern, filter mask and tile layout."""
import numpy as np
import pytest

import oracle as orc
from util import (boundary_cloud, clustered_cloud, compare_bands, grid_desc, make_grid, run_product, spec,
                  uniform_cloud, cloud as mk)

pytestmark = pytest.mark.gpu


def _specs(pcr, chans=("value",)):
    R = pcr.ReductionType
    out = []
    for c in chans:
        out += [spec(pcr, c, t) for t in (R.Sum, R.Max, R.Min, R.Average, R.Count)]
    return out


@pytest.mark.parametrize("log2,pool", [(4, 0), (8, 0), (12, 0), (0, 0), (6, 4096), (10, 20000)],
                         ids=["bins16", "bins256", "bins4096", "auto", "tiny_pool", "small_pool"])
def test_binned_matches_oracle_uniform(gpu_pcr, oracle, log2, pool):
    pcr = gpu_pcr
    w, h = 300, 211
    gc = make_grid(pcr, w, h, tile=64)
    x, y, ch = uniform_cloud(150_000, w, h, seed=5, margin=-3.0)          # some points fall outside
    ch = {"value": (ch["value"] * 10 - 5).astype(np.float32)}
    specs = _specs(pcr)
    got, _ = run_product(pcr, gc, [(x, y, ch)], specs, point_kernel=3, bin_cells_log2=log2, bin_pool_points=pool)
    gd = grid_desc(gc)
    ref = oracle.run(gd, [(x, y, ch)], specs)
    compare_bands(oracle, gd, [(x, y, ch)], specs, ref, got, f"binned log2={log2} pool={pool}")


def test_binned_equals_direct_kernel_exact_bands(gpu_pcr):
    """Same inputs through both Point kernels: Count/Max/Min identical bit for bit, NaN masks identical."""
    pcr = gpu_pcr
    gc = make_grid(pcr, 512, 384, tile=128)
    x, y, ch = clustered_cloud(400_000, 512, 384, seed=9)
    specs = _specs(pcr)
    a, _ = run_product(pcr, gc, [(x, y, ch)], specs, point_kernel=1)
    b, _ = run_product(pcr, gc, [(x, y, ch)], specs, point_kernel=3, bin_cells_log2=11)
    for i, (u, v) in enumerate(zip(a, b)):
        assert np.array_equal(np.isnan(u), np.isnan(v)), i
        if i in (1, 2, 4):
            assert np.array_equal(u, v, equal_nan=True), i
        else:
            np.testing.assert_allclose(u, v, rtol=1e-5, atol=1e-5, equal_nan=True)


def test_binned_boundary_probes_and_two_channels(gpu_pcr, oracle):
    pcr = gpu_pcr
    w, h = 64, 48
    gc = make_grid(pcr, w, h, tile=16)
    x, y, ch = boundary_cloud(w, h)
    rng = np.random.default_rng(3)
    ch["other"] = rng.normal(0, 100, len(x)).astype(np.float32)
    ch["value"][::7] = np.nan                                   # NaN values: skipped by Max/Min, poison Sum
    specs = _specs(pcr, ("value", "other"))
    got, _ = run_product(pcr, gc, [(x, y, ch)], specs, point_kernel=3, bin_cells_log2=5)
    gd = grid_desc(gc)
    ref = oracle.run(gd, [(x, y, ch)], specs)
    compare_bands(oracle, gd, [(x, y, ch)], specs, ref, got, "binned boundary + 2 channels")


def test_binned_count_only_and_too_many_channels(gpu_pcr, oracle):
    """Entries without a value (4-byte entries, runs padded to 4) and a pass with more channels than an entry
    carries (falls back to the direct kernel even when the binned path is requested)."""
    pcr = gpu_pcr
    w, h = 190, 133
    gc = make_grid(pcr, w, h, tile=64)
    x, y, ch = uniform_cloud(120_000, w, h, seed=8, margin=-2.0)
    rng = np.random.default_rng(4)
    chans = {"value": ch["value"], "b": rng.normal(0, 2, len(x)).astype(np.float32),
             "c": rng.uniform(0, 9, len(x)).astype(np.float32)}
    gd = grid_desc(gc)
    specs = [spec(pcr, "value", pcr.ReductionType.Count)]
    got, _ = run_product(pcr, gc, [(x, y, chans)], specs, point_kernel=3, bin_cells_log2=7)
    compare_bands(oracle, gd, [(x, y, chans)], specs, oracle.run(gd, [(x, y, chans)], specs), got, "binned Count only")
    R = pcr.ReductionType
    specs = [spec(pcr, "value", R.Sum), spec(pcr, "b", R.Max), spec(pcr, "c", R.Average)]
    got, _ = run_product(pcr, gc, [(x, y, chans)], specs, point_kernel=3, bin_cells_log2=7)
    compare_bands(oracle, gd, [(x, y, chans)], specs, oracle.run(gd, [(x, y, chans)], specs), got, "3 channels")


def test_binned_many_ingests_refinalize_and_reset(gpu_pcr, oracle):
    """Entries pile up over several ingests; finalize folds them; later ingests keep accumulating; the pool
    is small enough that some ingests fold early; reset drops pending entries."""
    pcr = gpu_pcr
    w, h = 200, 150
    gc = make_grid(pcr, w, h, tile=64)
    specs = _specs(pcr)
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.point_kernel = 3; cfg.bin_cells_log2 = 9; cfg.bin_pool_points = 50_000
    p = pcr.Pipeline.create(cfg)
    assert p is not None
    gd = grid_desc(gc)
    clouds = []
    for r in range(6):
        x, y, ch = uniform_cloud(30_000 + 7000 * r, w, h, seed=100 + r, margin=-1.0)
        ch = {"value": ch["value"]}
        clouds.append((x, y, ch))
        p.ingest(mk(pcr, x, y, ch) if r % 2 == 0 else mk(pcr, x, y, ch).to_device())
        if r in (2, 5):
            p.finalize()
            got = [np.array(p.result().band_array(i)) for i in range(len(specs))]
            ref = oracle.run(gd, clouds, specs)
            compare_bands(oracle, gd, clouds, specs, ref, got, f"binned after ingest {r}")
    assert p.stats().points_processed == sum(len(c[0]) for c in clouds)
    p.ingest(mk(pcr, *clouds[0]))
    p.reset()                                                   # pending entries must not survive
    p.ingest(mk(pcr, *clouds[1]))
    p.finalize()
    got = [np.array(p.result().band_array(i)) for i in range(len(specs))]
    ref = oracle.run(gd, [clouds[1]], specs)
    compare_bands(oracle, gd, [clouds[1]], specs, ref, got, "binned after reset")


def test_binned_with_filter_and_untouched_tiles(gpu_pcr, oracle):
    pcr = gpu_pcr
    w, h = 256, 256
    gc = make_grid(pcr, w, h, tile=64)
    rng = np.random.default_rng(11)
    n = 100_000
    x, y = rng.uniform(0, w, n), rng.uniform(0, h * 0.4, n)            # the north tiles stay untouched -> NaN
    v = rng.uniform(0, 1, n).astype(np.float32)
    cls = rng.integers(0, 5, n).astype(np.float32)
    specs = [spec(pcr, "value", pcr.ReductionType.Sum), spec(pcr, "value", pcr.ReductionType.Count)]
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.point_kernel = 3; cfg.bin_cells_log2 = 10
    cfg.filter.add("cls", pcr.CompareOp.Equal, 2.0)
    p = pcr.Pipeline.create(cfg)
    p.ingest(mk(pcr, x, y, {"value": v, "cls": cls}))
    p.finalize()
    got = [np.array(p.result().band_array(i)) for i in range(2)]
    keep = cls == 2.0
    gd = grid_desc(gc)
    clouds = [(x[keep], y[keep], {"value": v[keep]})]
    ref = oracle.run(gd, clouds, specs)
    compare_bands(oracle, gd, clouds, specs, ref, got, "binned + filter")
    assert np.isnan(got[0][:64]).all() and p.stats().points_processed == int(keep.sum())


def test_binned_large_grid_auto(gpu_pcr):
    """A grid whose records exceed 256 MB picks the binned path by itself: 6000 x 6000 x 16 B = 576 MB.
    Count and Max are checked cell by cell against a numpy replay, Average within the fp32 bound."""
    pcr = gpu_pcr
    W = 6000
    gc = make_grid(pcr, W, W)
    x, y, ch = clustered_cloud(6_000_000, W, W, seed=21)
    R = pcr.ReductionType
    specs = [spec(pcr, "value", R.Average), spec(pcr, "value", R.Max), spec(pcr, "value", R.Count)]
    got, p = run_product(pcr, gc, [(x, y, ch)], specs, loc=pcr.MemoryLocation.Device)
    col = np.clip(np.floor(x).astype(np.int64), 0, W - 1)
    row = np.clip(np.floor((y - W) / -1.0).astype(np.int64), 0, W - 1)
    cell = row * W + col
    cnt = np.bincount(cell, minlength=W * W).astype(np.float32).reshape(W, W)
    mx = np.full(W * W, -np.inf, np.float32)
    np.maximum.at(mx, cell, ch["value"])
    sums = np.bincount(cell, weights=ch["value"].astype(np.float64), minlength=W * W).reshape(W, W)
    has = cnt > 0
    assert np.array_equal(np.isnan(got[2]), ~has)
    assert np.array_equal(got[2][has], cnt[has])
    assert np.array_equal(got[1][has], mx.reshape(W, W)[has])
    avg = sums[has] / cnt[has]
    np.testing.assert_allclose(got[0][has], avg, rtol=1e-5, atol=1e-6)
