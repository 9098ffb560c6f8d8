"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the public pcr API,
i.e. through the C-ABI of libpcr_b200.so; the oracle (oracle/pcr_oracle.c) and the committed
reference fixtures (tests/golden/) are only the checkers.  /root/reference is never read.

Bars (BASELINE.json north_star):
  * cell indices (via Count), Count, Max, Min: bit-exact;
  * Sum / Average / WeightedAverage: within util.compare_bands' stated fp32 bound;
  * NaN masks identical (touched-tile rule included);
  * deterministic mode: byte-identical across runs.
"""
import glob
import os

import numpy as np
import pytest

import oracle as orc
import make_golden as mg
from known_answers import ACCUMULATOR, pipeline_cases
from util import (boundary_cloud, clustered_cloud, compare_bands, grid_desc, make_grid, run_product,
                  spec, uniform_cloud)

pytestmark = pytest.mark.gpu

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("pcrt_", "pcrp_")))
# Line endpoints use f64 cos/sin rounded to f32 where the reference uses glibc cosf/sinf; the two
# differ in the last bit for a tiny fraction of angles, which can move an endpoint across a .5
# rounding boundary.  Measured flip rate is reported by test_line_flip_rate; fixtures allow 0.
ALL_POINT = ("Sum", "Max", "Min", "Average", "WeightedAverage", "Count")


def to_product_spec(pcr, s):
    p = pcr.ReductionSpec()
    p.value_channel = s.value_channel
    p.type = pcr.ReductionType(int(s.type))
    p.output_band_name = s.output_band_name
    for k, v in vars(s.glyph).items():
        setattr(p.glyph, k, pcr.GlyphType(int(v)) if k == "type" else v)
    return p


def product_grid(pcr, gd):
    gc = pcr.GridConfig()
    gc.bounds.min_x, gc.bounds.min_y, gc.bounds.max_x, gc.bounds.max_y = gd.min_x, gd.min_y, gd.max_x, gd.max_y
    gc.cell_size_x, gc.cell_size_y = gd.cell_size_x, gd.cell_size_y
    gc.tile_width, gc.tile_height = gd.tile_width, gd.tile_height
    gc.compute_dimensions()
    return gc


# ---- reference fixtures ---------------------------------------------------------------
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_reference_fixture(gpu_pcr, oracle, path):
    gd, clouds, specs, ref_bands = mg.load(path)
    got, _ = run_product(gpu_pcr, product_grid(gpu_pcr, gd), clouds, [to_product_spec(gpu_pcr, s) for s in specs])
    compare_bands(oracle, gd, clouds, specs, ref_bands, got, os.path.basename(path), device_weights=True)


@pytest.mark.parametrize("name", sorted(ACCUMULATOR))
def test_reference_gtest_accumulator_vectors(gpu_pcr, name):
    k = ACCUMULATOR[name]
    gc = make_grid(gpu_pcr, k["w"], k["h"], tile=k["tile"])
    got, _ = run_product(gpu_pcr, gc, k["clouds"], [spec(gpu_pcr, "v", gpu_pcr.ReductionType(t)) for t in k["types"]])
    for band, exp in zip(got, k["expect_head"]):
        np.testing.assert_array_equal(band[0], np.array(exp, np.float32))


@pytest.mark.parametrize("name", sorted(pipeline_cases()))
def test_reference_gtest_pipeline_vectors(gpu_pcr, name):
    k = pipeline_cases()[name]
    gc = make_grid(gpu_pcr, k["w"], k["h"], tile=k["tile"])
    got, p = run_product(gpu_pcr, gc, k["clouds"],
                         [spec(gpu_pcr, k["channel"], gpu_pcr.ReductionType(t)) for t in k["types"]])
    for band, exp in zip(got, k["expect"]):
        np.testing.assert_array_equal(band, exp)
    st = p.stats()                                            # test_pipeline.cpp:398-441
    assert st.collections_processed == len(k["clouds"])
    assert st.points_processed == sum(len(c[0]) for c in k["clouds"])


# ---- differential vs the oracle, Point glyph ---------------------------------------------
def point_specs(pcr, channel="value"):
    return [spec(pcr, channel, getattr(pcr.ReductionType, n)) for n in ALL_POINT]


def check_vs_oracle(pcr, oracle, gc, clouds, specs, what, device_weights=False, **knobs):
    gd = grid_desc(gc)
    ref = oracle.run(gd, clouds, specs)
    got, p = run_product(pcr, gc, clouds, specs, **knobs)
    compare_bands(oracle, gd, clouds, specs, ref, got, what, device_weights=device_weights)
    return got, p


@pytest.mark.parametrize("loc", ["Host", "HostPinned", "Device"])
def test_point_uniform_all_reducers(gpu_pcr, oracle, loc):
    gc = make_grid(gpu_pcr, 256, 192, tile=64)
    x, y, ch = uniform_cloud(300_000, 256, 192, seed=42, margin=-3.0)   # some points outside
    check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch)], point_specs(gpu_pcr), f"uniform/{loc}",
                    loc=getattr(gpu_pcr.MemoryLocation, loc))


def test_point_boundary_probes(gpu_pcr, oracle):
    for (w, h, cell, tile) in [(64, 48, 1.0, 16), (30, 20, 0.3, 7), (50, 50, 2.0, 4096), (17, 33, 1.7, 5)]:
        gc = make_grid(gpu_pcr, w * cell, h * cell, cell=cell, tile=tile)
        x, y, ch = boundary_cloud(w * cell, h * cell)
        ch["value"][::53] = np.nan
        ch["value"][3::59] = -np.inf
        ch["value"][5::61] = np.inf
        ch["value"][7::67] = np.float32(-3.4028235e38)
        check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch)], point_specs(gpu_pcr), f"boundary {w}x{h} cs={cell}")


def test_point_offset_bounds_negative_coords(gpu_pcr, oracle):
    gc = make_grid(gpu_pcr, 37.3, 21.9, cell=0.7, tile=8, min_x=-1234.5678, min_y=987654.321, cell_y=-0.35)
    rng = np.random.default_rng(5)
    n = 100_000
    x = rng.uniform(gc.bounds.min_x - 1, gc.bounds.max_x + 1, n)
    y = rng.uniform(gc.bounds.min_y - 1, gc.bounds.max_y + 1, n)
    ch = {"value": rng.normal(0, 100, n).astype(np.float32)}
    check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch)], point_specs(gpu_pcr), "offset bounds")


def test_point_clustered_clipped_to_bbox(gpu_pcr, oracle):
    gc = make_grid(gpu_pcr, 512, 512, tile=128)
    x, y, ch = clustered_cloud(400_000, 512, 512, seed=42)
    check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch)], point_specs(gpu_pcr), "clustered")


def test_point_sorted_input_exercises_run_aggregation(gpu_pcr, oracle):
    """Scan-ordered cloud: long runs of consecutive points in one cell (warp run aggregation)."""
    gc = make_grid(gpu_pcr, 64, 64, tile=32)
    rng = np.random.default_rng(9)
    n = 200_000
    x = np.sort(rng.uniform(0, 64, n)); y = np.repeat(rng.uniform(0, 64, n // 100), 100)
    ch = {"value": rng.uniform(-1, 1, n).astype(np.float32)}
    ch["value"][::41] = np.nan
    specs = [spec(gpu_pcr, "value", gpu_pcr.ReductionType.Max), spec(gpu_pcr, "value", gpu_pcr.ReductionType.Min),
             spec(gpu_pcr, "value", gpu_pcr.ReductionType.Count)]
    for knobs in ({}, {"warp_aggregate": 2}, {"point_kernel": 2}):
        check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch)], specs, f"sorted {knobs}", **knobs)
    ch2 = {"value": rng.uniform(0, 1, n).astype(np.float32)}
    check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch2)], point_specs(gpu_pcr), "sorted sums")


@pytest.mark.parametrize("knobs", [{"point_kernel": 1}, {"point_kernel": 2}, {"point_kernel": 2, "warp_aggregate": 2},
                                   {"ring_slot_points": 4096, "ring_depth": 2}])
def test_point_kernel_variants_and_ring(gpu_pcr, oracle, knobs):
    gc = make_grid(gpu_pcr, 300, 200, tile=128)
    x, y, ch = uniform_cloud(123_457, 300, 200, seed=3, margin=-1.0)   # odd count: tails
    for loc in ("Host", "Device"):
        check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch)], point_specs(gpu_pcr), f"{knobs}/{loc}",
                        loc=getattr(gpu_pcr.MemoryLocation, loc), **knobs)


def test_point_multiple_ingests_and_refinalize(gpu_pcr, oracle):
    gc = make_grid(gpu_pcr, 128, 128, tile=32)
    clouds = [uniform_cloud(50_000, 128, 40, seed=s) for s in (1, 2, 3)]     # only the south rows
    specs = point_specs(gpu_pcr)
    gd = grid_desc(gc)
    cfg = gpu_pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = gpu_pcr.ExecutionMode.GPU
    p = gpu_pcr.Pipeline.create(cfg)
    from util import cloud as mk
    for i, c in enumerate(clouds):
        p.ingest(mk(gpu_pcr, *c))
        p.finalize()                                # finalize after every ingest, like the benchmarks do
        got = [np.array(p.result().band_array(b)) for b in range(len(specs))]
        ref = oracle.run(gd, clouds[:i + 1], specs)
        compare_bands(oracle, gd, clouds[:i + 1], specs, ref, got, f"after ingest {i}")
    assert np.isnan(got[0][:64]).all()              # untouched north tiles stay NaN, even for Sum


def test_fused_multi_channel_and_many_reductions(gpu_pcr, oracle):
    """More reductions than one record can hold: the planner must split passes."""
    gc = make_grid(gpu_pcr, 96, 96, tile=32)
    rng = np.random.default_rng(17)
    n = 80_000
    x, y = rng.uniform(0, 96, n), rng.uniform(0, 96, n)
    ch = {k: rng.normal(i, 2, n).astype(np.float32) for i, k in enumerate("abcdef")}
    R = gpu_pcr.ReductionType
    specs = [spec(gpu_pcr, k, t) for k in "abcdef" for t in (R.Sum, R.Max, R.Min, R.Average)] + \
            [spec(gpu_pcr, "a", R.Count)]
    check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch)], specs, "25 reductions / 6 channels")


def test_empty_and_all_outside_clouds(gpu_pcr):
    gc = make_grid(gpu_pcr, 16, 16)
    specs = point_specs(gpu_pcr)
    got, p = run_product(gpu_pcr, gc, [(np.array([]), np.array([]), {"value": np.array([], np.float32)})], specs)
    assert all(np.isnan(b).all() for b in got)
    assert p.stats().collections_processed == 0       # empty cloud is a no-op (pipeline.cpp:284-287)
    got, p = run_product(gpu_pcr, gc, [(np.array([-5.0, 99.0]), np.array([3.0, 3.0]),
                                        {"value": np.array([1, 2], np.float32)})], specs)
    assert all(np.isnan(b).all() for b in got)         # no tile touched -> NaN everywhere
    assert p.stats().points_processed == 2             # but the points count (pipeline.cpp:749)


# ---- glyphs vs the oracle ---------------------------------------------------------------------
def test_line_vs_oracle(gpu_pcr, oracle):
    gc = make_grid(gpu_pcr, 200, 160, tile=64)
    x, y, ch = uniform_cloud(60_000, 200, 160, seed=8, margin=-1.0)
    rng = np.random.default_rng(8)
    ch["hl"] = rng.uniform(0, 20, len(x)).astype(np.float32)
    specs = []
    for t in ("WeightedAverage", "Sum", "Count", "Average"):
        s = gpu_pcr.line_splat_spec("value", "direction", "hl", max_radius_cells=18.0)
        s.type = getattr(gpu_pcr.ReductionType, t)
        specs.append(s)
    check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch)], specs, "line", device_weights=True)


def test_line_flip_rate(gpu_pcr, oracle):
    """Cell-set parity of the Line glyph at scale: count the cells whose Count differs from the
    oracle's (device f64 cos/sin vs glibc cosf/sinf).  Stated bound: <= 1e-6 of painted cells."""
    gc = make_grid(gpu_pcr, 1000, 1000)
    x, y, ch = uniform_cloud(1_000_000, 1000, 1000, seed=42)
    ch["hl"] = np.full(len(x), 16.0, np.float32)
    s = gpu_pcr.line_splat_spec("value", "direction", "hl", max_radius_cells=18.0)
    s.type = gpu_pcr.ReductionType.Count
    gd = grid_desc(gc)
    ref = oracle.run(gd, [(x, y, ch)], [s])[0]
    got, _ = run_product(gpu_pcr, gc, [(x, y, ch)], [s])
    painted = float(np.nansum(ref))
    flips = float(np.nansum(np.abs(np.nan_to_num(got[0]) - np.nan_to_num(ref))))
    print(f"line flip rate: {flips:.0f} cell-visits differ of {painted:.0f} painted ({flips / painted:.2e})")
    assert flips <= max(2.0, 1e-6 * painted)


@pytest.mark.parametrize("kernel", [1, 2, 3], ids=["scatter", "gather", "bin_gemm"])
def test_gaussian_vs_oracle(gpu_pcr, oracle, kernel):
    gc = make_grid(gpu_pcr, 160, 120, tile=64)
    rng = np.random.default_rng(21)
    n = 6000
    x, y = rng.uniform(-1, 161, n), rng.uniform(-1, 121, n)
    ch = {"value": rng.uniform(0, 1, n).astype(np.float32), "sigma": rng.uniform(-0.5, 5, n).astype(np.float32),
          "s2": rng.uniform(0.3, 3, n).astype(np.float32), "rot": rng.uniform(-3.2, 3.2, n).astype(np.float32)}
    specs = []
    for t in ("WeightedAverage", "Sum", "Count"):
        s = gpu_pcr.gaussian_splat_spec("value", "sigma", "sigma", default_sigma=1.5, max_radius_cells=12.0)
        s.type = getattr(gpu_pcr.ReductionType, t)
        specs.append(s)
    specs.append(gpu_pcr.gaussian_splat_spec("value", "sigma", "s2", "rot", default_sigma=2.0, max_radius_cells=10.0))
    specs.append(gpu_pcr.gaussian_splat_spec("value", default_sigma_x=16.0, default_sigma_y=16.0, max_radius_cells=32.0))
    check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch)], specs, "gaussian", device_weights=True,
                    gaussian_kernel=kernel)


def test_gaussian_gather_edge_cases(gpu_pcr, oracle):
    """Gather-specific seams: grid not a multiple of the 32-cell gather tile, reference tiles smaller
    than / misaligned with gather tiles, points on the max edges (centre cell == width), radius cap
    < 1, huge cap with small sigma, non-unit cells, several chunks."""
    rng = np.random.default_rng(33)
    for (w, h, cell, tile, cap, sig) in [(45, 37, 1.0, 10, 6.0, 1.3), (64, 64, 1.0, 4096, 0.4, 2.0),
                                         (70, 33, 0.5, 24, 300.0, 0.6), (33, 95, 2.0, 7, 12.0, 9.0)]:
        gc = make_grid(gpu_pcr, w * cell, h * cell, cell=cell, tile=tile)
        n = 3000
        x = rng.uniform(-1, w * cell + 1, n); y = rng.uniform(-1, h * cell + 1, n)
        x[:40] = w * cell; y[40:80] = 0.0; x[80:90] = 0.0; y[90:100] = h * cell
        ch = {"value": rng.uniform(-1, 1, n).astype(np.float32)}
        s1 = gpu_pcr.gaussian_splat_spec("value", default_sigma=sig, max_radius_cells=cap)
        s2 = gpu_pcr.gaussian_splat_spec("value", default_sigma=sig, max_radius_cells=cap)
        s2.type = gpu_pcr.ReductionType.Count
        for knobs in ({"gaussian_kernel": 2}, {"gaussian_kernel": 2, "ring_slot_points": 1024},
                      {"gaussian_kernel": 3}, {"gaussian_kernel": 3, "ring_slot_points": 1024}):
            check_vs_oracle(gpu_pcr, oracle, gc, [(x, y, ch)], [s1, s2], f"gather {w}x{h} cap={cap} {knobs}",
                            device_weights=True, **knobs)


def test_gaussian_gather_is_bit_reproducible(gpu_pcr):
    gc = make_grid(gpu_pcr, 200, 150, tile=64)
    rng = np.random.default_rng(5)
    n = 50_000
    x, y = rng.uniform(0, 200, n), rng.uniform(0, 150, n)
    ch = {"value": rng.uniform(0, 1, n).astype(np.float32), "sigma": rng.uniform(0.5, 4, n).astype(np.float32)}
    s = gpu_pcr.gaussian_splat_spec("value", "sigma", "sigma", max_radius_cells=12.0)
    runs = [run_product(gpu_pcr, gc, [(x, y, ch)], [s], deterministic=True)[0][0].tobytes() for _ in range(3)]
    assert runs[0] == runs[1] == runs[2]


def test_glyph_with_max_is_not_implemented(gpu_pcr):
    gc = make_grid(gpu_pcr, 16, 16)
    s = gpu_pcr.line_splat_spec("value")
    s.type = gpu_pcr.ReductionType.Max
    cfg = gpu_pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = [s]; cfg.exec_mode = gpu_pcr.ExecutionMode.GPU
    p = gpu_pcr.Pipeline.create(cfg)
    assert p is not None
    from util import cloud as mk
    with pytest.raises(RuntimeError, match="glyph splatting only supports"):
        p.ingest(mk(gpu_pcr, [1.0], [1.0], {"value": [1.0]}))


def test_pipeline_memory_released_without_gc(gpu_pcr):
    """A finalized pipeline must give its device memory back when the last reference goes — by reference
    counting, not by the cyclic garbage collector (result Grid and band views keep the pipeline alive, the
    pipeline does not keep them)."""
    import gc
    pcr = gpu_pcr
    gc.collect()
    gc.disable()
    try:
        free0, _ = pcr.device_mem_info()
        g = make_grid(pcr, 3000, 3000)
        x, y, ch = uniform_cloud(200_000, 3000, 3000, seed=1)
        s = spec(pcr, "value", pcr.ReductionType.Sum)
        for _ in range(3):
            got, p = run_product(pcr, g, [(x, y, ch)], [s], point_kernel=3, bin_pool_points=60_000_000)
            grid = p.result()
            assert grid is p.result()                        # one wrapper while somebody holds it
            view = grid.band_array(0)
            del p, grid                                      # the view alone keeps the pinned result alive
            assert np.array_equal(view, got[0], equal_nan=True)
            held, _ = pcr.device_mem_info()
            assert free0 - held > 300 << 20                  # ... and with it the pipeline (pool + records + bands)
            del view, got
        free1, _ = pcr.device_mem_info()
        assert free0 - free1 < 64 << 20, f"{(free0 - free1) >> 20} MB still held after the pipelines went out of scope"
    finally:
        gc.enable()


# ---- API behaviour ---------------------------------------------------------------------------
def test_error_behaviour(gpu_pcr):
    from util import cloud as mk
    gc = make_grid(gpu_pcr, 8, 8)
    cfg = gpu_pcr.PipelineConfig(); cfg.grid = gc; cfg.exec_mode = gpu_pcr.ExecutionMode.GPU
    p = gpu_pcr.Pipeline.create(cfg)                        # no reductions: create ok, validate fails
    assert p is not None
    with pytest.raises(RuntimeError, match="at least one reduction"):
        p.validate()
    cfg.reductions = [spec(gpu_pcr, "intensity", gpu_pcr.ReductionType.Sum)]
    p = gpu_pcr.Pipeline.create(cfg)
    p.validate()
    with pytest.raises(RuntimeError, match="value channel not found: intensity"):
        p.ingest(mk(gpu_pcr, [1.0], [1.0], {"other": [1.0]}))
    c = gpu_pcr.PointCloud.create(4)
    c.set_x_array(np.ones(4)); c.set_y_array(np.ones(4)); c.add_channel("intensity", gpu_pcr.DataType.Int32)
    with pytest.raises(RuntimeError, match="must be Float32"):
        p.ingest(c)
    cfg.reductions = [spec(gpu_pcr, "intensity", gpu_pcr.ReductionType.Median)]
    assert gpu_pcr.Pipeline.create(cfg) is None             # unregistered op: create fails (nullptr upstream)
    cfg.reductions = [spec(gpu_pcr, "intensity", gpu_pcr.ReductionType.Sum)]
    cfg.exec_mode = gpu_pcr.ExecutionMode.CPU
    assert gpu_pcr.Pipeline.create(cfg) is None             # no CPU path behind this API
    for mode in ("Auto", "Hybrid"):                         # both are the GPU path
        cfg.exec_mode = getattr(gpu_pcr.ExecutionMode, mode)
        assert gpu_pcr.Pipeline.create(cfg) is not None


def test_progress_callback_and_cancel(gpu_pcr):
    from util import cloud as mk
    gc = make_grid(gpu_pcr, 8, 8)
    cfg = gpu_pcr.PipelineConfig(); cfg.grid = gc; cfg.exec_mode = gpu_pcr.ExecutionMode.GPU
    cfg.reductions = [spec(gpu_pcr, "v", gpu_pcr.ReductionType.Count)]
    p = gpu_pcr.Pipeline.create(cfg)
    seen = []
    p.set_progress_callback(lambda info: (seen.append((info.collections_processed, info.points_processed)), True)[1])
    c = mk(gpu_pcr, [1.0, 2.0], [1.0, 2.0], {"v": [1.0, 1.0]})
    p.ingest(c); p.ingest(c)
    assert seen == [(1, 2), (2, 4)]
    p.set_progress_callback(lambda info: False)
    with pytest.raises(RuntimeError, match="cancelled by user"):
        p.ingest(c)


def test_band_names_and_result_grid(gpu_pcr):
    gc = make_grid(gpu_pcr, 8, 6)
    specs = [spec(gpu_pcr, "v", gpu_pcr.ReductionType.Average), spec(gpu_pcr, "v", gpu_pcr.ReductionType.Max, "peak")]
    got, p = run_product(gpu_pcr, gc, [([1.5], [1.5], {"v": [2.0]})], specs)
    g = p.result()
    assert (g.cols(), g.rows(), g.num_bands()) == (8, 6, 2)
    assert g.band_desc(0).name == "v_3" and g.band_desc(1).name == "peak"   # pipeline.cpp:1178-1180
    assert g.band_array(0).shape == (6, 8) and g.band_array(0).dtype == np.float32
    assert got[0][4, 1] == 2.0 and got[1][4, 1] == 2.0


# ---- deterministic mode -----------------------------------------------------------------------
def test_deterministic_mode_bit_reproducible(gpu_pcr, oracle):
    gc = make_grid(gpu_pcr, 128, 128, tile=64)
    x, y, ch = clustered_cloud(300_000, 128, 128, seed=4)
    specs = point_specs(gpu_pcr)
    runs = []
    for _ in range(3):
        got, _ = run_product(gpu_pcr, gc, [(x, y, ch), (x[::3], y[::3], {"value": ch["value"][::3]})], specs,
                             deterministic=True)
        runs.append([b.tobytes() for b in got])
    assert runs[0] == runs[1] == runs[2]
    clouds = [(x, y, ch), (x[::3], y[::3], {"value": ch["value"][::3]})]
    gd = grid_desc(gc)
    compare_bands(oracle, gd, clouds, specs, oracle.run(gd, clouds, specs),
                  [np.frombuffer(b, np.float32).reshape(128, 128) for b in runs[0]], "deterministic")
    # in deterministic mode the fold is the oracle's: original point order per cell -> Sum bit-exact
    ref = oracle.run(gd, [clouds[0]], [specs[0]])[0]
    got, _ = run_product(gpu_pcr, gc, [clouds[0]], [specs[0]], deterministic=True)
    assert np.array_equal(got[0], ref, equal_nan=True)


def test_deterministic_mode_hot_cells_in_order(gpu_pcr, oracle):
    """Runs far longer than what one thread folds on its own (cells holding 150 000 and 40 000 points, one run
    crossing many CTAs): the warp-assisted fold must still add in original point order — Sum equal to the
    oracle's in-order fold bit for bit, over two ingests."""
    pcr = gpu_pcr
    gc = make_grid(pcr, 64, 64, tile=32)
    rng = np.random.default_rng(11)
    n_hot, n_warm, n_rest = 150_000, 40_000, 110_000
    x = np.concatenate([np.full(n_hot, 10.25), rng.uniform(40.0, 41.0, n_warm), rng.uniform(0, 64, n_rest)])
    y = np.concatenate([np.full(n_hot, 20.75), rng.uniform(7.0, 8.0, n_warm), rng.uniform(0, 64, n_rest)])
    v = rng.normal(0, 100, len(x)).astype(np.float32)            # cancellation: the order of the adds shows
    perm = rng.permutation(len(x))
    x, y, v = x[perm], y[perm], v[perm]
    R = pcr.ReductionType
    specs = [spec(pcr, "value", t) for t in (R.Sum, R.Max, R.Min, R.Count, R.Average)]
    clouds = [(x, y, {"value": v}), (x[::2], y[::2], {"value": v[::2]})]
    gd = grid_desc(gc)
    ref = oracle.run(gd, clouds, specs)
    got, _ = run_product(pcr, gc, clouds, specs, deterministic=True)
    for i in range(len(specs)):
        assert np.array_equal(got[i], ref[i], equal_nan=True), f"band {i} differs from the in-order fold"
    again, _ = run_product(pcr, gc, clouds, specs, deterministic=True)
    assert all(a.tobytes() == b.tobytes() for a, b in zip(got, again))


# ---- full-size properties (BASELINE config 2: 5M points, 1000x1000) ------------------------------
def test_full_size_point_config(gpu_pcr, oracle):
    gc = make_grid(gpu_pcr, 1000, 1000)
    x, y, ch = uniform_cloud(5_000_000, 1000, 1000, seed=42)
    R = gpu_pcr.ReductionType
    specs = [spec(gpu_pcr, "value", R.Sum), spec(gpu_pcr, "value", R.Count), spec(gpu_pcr, "value", R.Max)]
    got, p = run_product(gpu_pcr, gc, [(x, y, ch)], specs)
    assert np.nansum(got[1].astype(np.float64)) == 5_000_000            # every point counted exactly once
    assert abs(np.nansum(got[0].astype(np.float64)) - ch["value"].astype(np.float64).sum()) < 1.0
    assert np.nanmax(got[2]) == ch["value"].max()
    gd = grid_desc(gc)
    ref = oracle.run(gd, [(x, y, ch)], specs)                             # the C oracle does 5M in < 1 s
    compare_bands(oracle, gd, [(x, y, ch)], specs, ref, got, "config 2 full size")
    # idempotence of finalize, exact doubling of Count on a second ingest
    from util import cloud as mk
    p.ingest(mk(gpu_pcr, x, y, ch)); p.finalize()
    assert np.array_equal(np.array(p.result().band_array(1)), 2 * ref[1], equal_nan=True)


# ---- BASELINE config 3 at full size: 5M lines, half length 16, per-point direction ------------------
def test_full_size_line_config(gpu_pcr, oracle):
    gc = make_grid(gpu_pcr, 1000, 1000)
    x, y, ch = uniform_cloud(5_000_000, 1000, 1000, seed=42)
    ch["hl"] = np.full(len(x), 16.0, np.float32)
    specs = []
    for t in ("WeightedAverage", "Count"):
        s = gpu_pcr.line_splat_spec("value", "direction", "hl", max_radius_cells=18.0)
        s.type = getattr(gpu_pcr.ReductionType, t)
        specs.append(s)
    gd = grid_desc(gc)
    ref = oracle.run(gd, [(x, y, ch)], specs)               # 165M cell visits: a few seconds in the C oracle
    got, _ = run_product(gpu_pcr, gc, [(x, y, ch)], specs)
    painted = float(np.nansum(ref[1].astype(np.float64)))
    flips = float(np.nansum(np.abs(np.nan_to_num(got[1]).astype(np.float64) - np.nan_to_num(ref[1]))))
    assert painted > 1.4e8                                  # ~29.6 cells per line
    assert flips <= max(2.0, 1e-6 * painted), (flips, painted)
    if flips == 0:                                          # same cell sets: the averages must agree cell by cell
        compare_bands(oracle, gd, [(x, y, ch)], specs, ref, got, "config 3 full size", device_weights=True)


# ---- BASELINE config 4 at full size: 5M Gaussians, sigma 4 and 16, radius cap 32 ---------------------
# The oracle needs minutes for 21G cell visits, so at full size the two independent device
# implementations (scatter: per-point REDs; gather: tile-binned tensor-core products) check each other,
# cell by cell against the weight mass, next to exact invariants; both are checked against the oracle at
# smaller sizes above.
@pytest.mark.parametrize("sigma", [4.0, 16.0])
def test_full_size_gaussian_config(gpu_pcr, sigma):
    from util import GLYPH_WEIGHT_RTOL
    gc = make_grid(gpu_pcr, 1000, 1000)
    x, y, ch = uniform_cloud(5_000_000, 1000, 1000, seed=42)
    ch["sigma"] = np.full(len(x), sigma, np.float32)
    specs = []
    for t in ("Sum", "Count", "WeightedAverage"):
        s = gpu_pcr.gaussian_splat_spec("value", "sigma", "sigma", max_radius_cells=32.0)
        s.type = getattr(gpu_pcr.ReductionType, t)
        specs.append(s)
    scatter, _ = run_product(gpu_pcr, gc, [(x, y, ch)], specs, gaussian_kernel=1)
    gather, _ = run_product(gpu_pcr, gc, [(x, y, ch)], specs, gaussian_kernel=2)
    mass = gather[1].astype(np.float64)                     # sum of weights per cell; values are in [0,1]
    assert not np.isnan(mass).any() and mass.min() > 0
    n_terms = (2 * min(3 * sigma, 32.0) + 1) ** 2 * 5.0     # ~ footprint cells x points per cell
    tol = (GLYPH_WEIGHT_RTOL + 2 * n_terms * 2.0 ** -24) * mass
    for k in (0, 1):
        assert (np.abs(scatter[k].astype(np.float64) - gather[k]) <= tol).all(), (sigma, k)
    # WeightedAverage = Sum / Count of the same run, bit for bit (same state words, one IEEE division)
    assert np.array_equal(gather[2], (gather[0] / gather[1]).astype(np.float32))
    assert np.abs(scatter[2].astype(np.float64) - gather[2]).max() <= 1e-4
    # a weighted average of values in [0,1) stays in [0,1); total weight mass is the sum of the kernels
    assert gather[2].min() >= 0.0 and gather[2].max() < 1.0
    # interior points carry the full kernel; its mass is the product of two 1-D sums (no clipping away
    # from the border), identical for every point up to the sub-cell offset: compare totals
    total = mass.sum()
    assert abs(total - scatter[1].astype(np.float64).sum()) <= 1e-6 * total
    # gather is deterministic: a second run reproduces every bit
    again, _ = run_product(gpu_pcr, gc, [(x, y, ch)], specs, gaussian_kernel=2)
    for a, b in zip(gather, again):
        assert np.array_equal(a, b, equal_nan=True)
    # the per-bin GEMM (the default kernel for this configuration): same bounds against the scatter kernel
    gemm, _ = run_product(gpu_pcr, gc, [(x, y, ch)], specs, gaussian_kernel=3)
    auto, _ = run_product(gpu_pcr, gc, [(x, y, ch)], specs)
    for got in (gemm, auto):
        for k in (0, 1):
            assert (np.abs(scatter[k].astype(np.float64) - got[k]) <= tol).all(), (sigma, k, "bin gemm")
        assert np.array_equal(got[2], (got[0] / got[1]).astype(np.float32))
        assert got[2].min() >= 0.0 and got[2].max() < 1.0
        assert abs(got[1].astype(np.float64).sum() - total) <= 1e-6 * total


# BASELINE config 4 against the ORACLE (not kernel against kernel): the first 200k points of the config-4 cloud,
# per-point sigma channel 4 / 16, max_radius_cells 32 (625 / 4225 cells per point), on the full 1000 x 1000
# grid, through both Gaussian kernels.  The C oracle needs ~5 ns per painted cell, which bounds the subsample.
@pytest.mark.parametrize("sigma", [4.0, 16.0])
def test_config4_subsample_vs_oracle(gpu_pcr, oracle, sigma):
    gc = make_grid(gpu_pcr, 1000, 1000)
    x, y, ch = uniform_cloud(5_000_000, 1000, 1000, seed=42)
    n = 200_000
    x, y = x[:n], y[:n]
    ch = {"value": ch["value"][:n], "sigma": np.full(n, sigma, np.float32)}
    specs = []
    for t in ("WeightedAverage", "Count"):
        s = gpu_pcr.gaussian_splat_spec("value", "sigma", "sigma", max_radius_cells=32.0)
        s.type = getattr(gpu_pcr.ReductionType, t)
        specs.append(s)
    gd = grid_desc(gc)
    ref = oracle.run(gd, [(x, y, ch)], specs)

    class Memo:                                  # the f64 replays behind the tolerances: once per spec, not per kernel
        def __init__(self, o): self.o, self.c = o, {}
        def bounds(self, gd_, clouds_, s_, want_weight=False):
            k = (id(s_), want_weight)
            if k not in self.c:
                self.c[k] = self.o.bounds(gd_, clouds_, s_, want_weight=want_weight)
            return self.c[k]
    memo = Memo(oracle)
    for kernel, name in ((1, "scatter"), (2, "gather"), (3, "bin_gemm")):
        got, _ = run_product(gpu_pcr, gc, [(x, y, ch)], specs, gaussian_kernel=kernel)
        compare_bands(memo, gd, [(x, y, ch)], specs, ref, got, f"config 4 sigma={sigma} {name}", device_weights=True)


# ---- BASELINE config 5 at single-GPU scale: 20000 x 20000 grid (25 reference tiles), clustered ----
def test_config5_like_large_grid(gpu_pcr):
    W = 20000
    free, _total = gpu_pcr.device_mem_info()
    if free < 16 << 30:
        pytest.skip("needs 16 GB of free HBM")
    gc = make_grid(gpu_pcr, W, W)
    assert (gc.tiles_x, gc.tiles_y) == (5, 5)
    n = 20_000_000
    x, y, ch = clustered_cloud(n, W, W, seed=5, k=12)
    R = gpu_pcr.ReductionType
    specs = [spec(gpu_pcr, "value", R.Average), spec(gpu_pcr, "value", R.Max), spec(gpu_pcr, "value", R.Count)]
    got, p = run_product(gpu_pcr, gc, [(x, y, ch)], specs)
    ok = (x >= 0) & (x <= W) & (y >= 0) & (y <= W)
    col = np.minimum(np.floor(x[ok]).astype(np.int64), W - 1)
    row = np.minimum(np.floor(W - y[ok]).astype(np.int64), W - 1)      # (y - max_y) / -1
    cells, inverse, counts = np.unique(row * W + col, return_inverse=True, return_counts=True)
    cnt = got[2].reshape(-1)
    assert np.count_nonzero(~np.isnan(cnt)) == len(cells)                 # Count finalizes empty cells to NaN
    assert np.array_equal(cnt[cells], counts.astype(np.float32))          # exact, every cell
    mx = np.full(len(cells), -np.inf, np.float32)
    np.maximum.at(mx, inverse, ch["value"][ok])
    assert np.array_equal(got[1].reshape(-1)[cells], mx)                  # Max bit-exact
    sums = np.bincount(inverse, weights=ch["value"][ok].astype(np.float64))
    avg = got[0].reshape(-1)[cells].astype(np.float64)
    assert (np.abs(avg - sums / counts) <= 2 * counts * 2.0 ** -24 * 1.5 + 1e-6).all()   # any-order fp32 sum bound
    # touched-tile rule (R11): Max/Average are NaN wherever Count is; untouched 4096-cell tiles are all NaN
    assert np.array_equal(np.isnan(got[0]), np.isnan(got[2])) and np.array_equal(np.isnan(got[1]), np.isnan(got[2]))
    assert p.stats().points_processed == n
    touched = np.zeros((5, 5), bool)
    touched[np.minimum(row // 4096, 4), np.minimum(col // 4096, 4)] = True
    assert p.stats().tiles_active == int(touched.sum())


# ---- independent pipelines from several host threads (the reference's ThreadSafety_ConcurrentAccess idea) ----
def test_independent_pipelines_in_parallel_threads(gpu_pcr, oracle):
    import threading
    gc = make_grid(gpu_pcr, 220, 170, tile=64)
    gd = grid_desc(gc)
    R = gpu_pcr.ReductionType
    results, errors = {}, []

    def work(tid):
        try:
            x, y, ch = uniform_cloud(120_000, 220, 170, seed=100 + tid, margin=-1.0)
            ch["sigma"] = np.full(len(x), 1.5 + tid, np.float32)
            specs = [spec(gpu_pcr, "value", R.Sum), spec(gpu_pcr, "value", R.Max), spec(gpu_pcr, "value", R.Count),
                     gpu_pcr.gaussian_splat_spec("value", "sigma", "sigma", max_radius_cells=12.0)]
            for rep in range(3):                                  # pipelines are created and torn down concurrently
                got, _ = run_product(gpu_pcr, gc, [(x, y, ch)], specs,
                                     loc=[None, gpu_pcr.MemoryLocation.HostPinned, gpu_pcr.MemoryLocation.Device][rep])
            results[tid] = ((x, y, ch), specs, got)
        except Exception as e:          # noqa
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in threads: t.start()
    for t in threads: t.join()
    assert not errors, errors
    for tid, (cl, specs, got) in results.items():
        ref = oracle.run(gd, [cl], specs)
        compare_bands(oracle, gd, [cl], specs, ref, got, f"thread {tid}", device_weights=True)
