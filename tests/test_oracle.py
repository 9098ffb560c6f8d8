"""CPU tests: pin oracle/pcr_oracle.c (the C restatement) against
  (a) the known-answer vectors of the reference's own gtests,
  (b) the committed golden fixtures produced by the unmodified reference (oracle/_ref),
  (c) when oracle/_ref is present (build container), live differential runs of the reference.
"""
import glob
import os

import numpy as np
import pytest

import oracle as orc
import make_golden as mg
from known_answers import ACCUMULATOR, pipeline_cases
from util import compare_bands

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith(("pcrt_", "pcrp_")))


# ---- (a) reference gtest vectors -------------------------------------------------
def test_world_to_cell_known_answers(oracle):
    gd = orc.GridDesc(0, 0, 100, 100)                       # test_grid_config.cpp:81-110
    assert oracle.world_to_cell(gd, 50.0, 50.0) == (50, 50, True)
    assert oracle.world_to_cell(gd, 0.0, 100.0) == (0, 0, True)
    assert oracle.world_to_cell(gd, -10.0, 50.0)[2] is False
    assert oracle.compute_dimensions(orc.GridDesc(0, 0, 100.5, 100.5)) == (101, 101)   # :31-44
    assert oracle.compute_dimensions(orc.GridDesc(0, 0, 100, 100)) == (100, 100)


def test_assign_known_answers(oracle):
    gd = orc.GridDesc(0, 0, 10, 10, tile_width=5, tile_height=5)     # test_tile_router.cpp:48-84
    i = np.arange(100)
    x = 0.5 + (i % 10); y = 9.5 - (i // 10)
    cell, tile, valid = oracle.assign(gd, x, y)
    assert valid.all()
    assert np.array_equal(cell, i)
    assert np.array_equal(tile, ((i // 10) // 5) * 2 + (i % 10) // 5)
    _, _, valid = oracle.assign(gd, [-1, 5, 15, 5, 5], [5, -1, 5, 15, 5])   # :86-120
    assert valid.tolist() == [0, 0, 0, 0, 1]


class _S:   # minimal spec-like
    def __init__(self, ch, t):
        self.value_channel, self.type, self.output_band_name = ch, t, ""
        self.glyph = mg.Glyph()


@pytest.mark.parametrize("name", sorted(ACCUMULATOR))
def test_accumulator_known_answers(oracle, name):
    k = ACCUMULATOR[name]
    gd = orc.GridDesc(0, 0, k["w"], k["h"], tile_width=k["tile"], tile_height=k["tile"])
    bands = oracle.run(gd, k["clouds"], [_S("v", t) for t in k["types"]])
    for band, exp in zip(bands, k["expect_head"]):
        np.testing.assert_array_equal(band[0], np.array(exp, np.float32))


@pytest.mark.parametrize("name", sorted(pipeline_cases()))
def test_pipeline_known_answers(oracle, name):
    k = pipeline_cases()[name]
    gd = orc.GridDesc(0, 0, k["w"], k["h"], tile_width=k["tile"], tile_height=k["tile"])
    bands = oracle.run(gd, k["clouds"], [_S(k["channel"], t) for t in k["types"]])
    for band, exp in zip(bands, k["expect"]):
        np.testing.assert_array_equal(band, exp)


def test_merge_is_op_merge(oracle):
    """Op::merge, builtin_ops.h:15,28,41,54,67,95 — the multi-GPU combine rule."""
    import ctypes as C
    for t, f in ((orc.SUM, np.add), (orc.MAX, np.fmax), (orc.MIN, np.fmin), (orc.COUNT, np.add)):
        a = np.array([1, 5, -3, 0], np.float32); b = np.array([2, -7, 9, 0], np.float32)
        d = a.copy()
        oracle.lib.orc_state_merge(t, d.ctypes.data, b.ctypes.data, 4)
        np.testing.assert_array_equal(d, f(a, b))


# ---- (b) golden fixtures from the unmodified reference -----------------------------
def _compare(oracle, gd, clouds, specs, ref_bands, got_bands, what):
    compare_bands(oracle, gd, clouds, specs, ref_bands, got_bands, what)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_golden(oracle, path):
    gd, clouds, specs, ref_bands = mg.load(path)
    got = oracle.run(gd, clouds, specs)
    _compare(oracle, gd, clouds, specs, ref_bands, got, os.path.basename(path))


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 16


def test_glyph_rejects_max_min(oracle):
    s = mg.Spec("v", orc.MAX, type=orc.GLYPH_LINE)
    with pytest.raises(RuntimeError, match="glyph splatting only supports"):
        oracle.run(orc.GridDesc(0, 0, 8, 8), [([1.0], [1.0], {"v": [1.0]})], [s])


# ---- (c) live differential against the reference (build container only) ------------
needs_ref = pytest.mark.skipif(not orc.reference_available(),
                               reason="oracle/_ref not built (no /root/reference on this box)")


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_live_differential_point(oracle, seed):
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(20, 90)), int(rng.integers(20, 90))
    cs = float(rng.choice([1.0, 0.5, 0.3, 2.0, 1.7]))
    gd = orc.GridDesc(-3.5, 10.25, -3.5 + w * cs, 10.25 + h * cs, cs, -cs * float(rng.choice([1.0, 0.7])),
                      int(rng.integers(4, 40)), int(rng.integers(4, 40)))
    n = 5000
    x = rng.uniform(gd.min_x - 2, gd.max_x + 2, n); y = rng.uniform(gd.min_y - 2, gd.max_y + 2, n)
    x[:50] = np.clip(x[:50], gd.min_x, gd.max_x); y[:50] = gd.min_y
    x[50:100] = gd.max_x
    ch = {"v": rng.normal(0, 10, n).astype(np.float32)}
    specs = [mg.Spec("v", t) for t in (orc.SUM, orc.MAX, orc.MIN, orc.AVERAGE, orc.WEIGHTED_AVERAGE, orc.COUNT)]
    ref = orc.reference_run(gd, [(x, y, ch)], specs)
    got = oracle.run(gd, [(x, y, ch)], specs)
    _compare(oracle, gd, [(x, y, ch)], specs, ref, got, f"live point seed {seed}")


@needs_ref
@pytest.mark.parametrize("seed", [11, 12])
def test_live_differential_glyphs(oracle, seed):
    rng = np.random.default_rng(seed)
    w, h = 70, 55
    gd = orc.GridDesc(0, 0, w, h, 1.0, -1.0, int(rng.integers(16, 40)), int(rng.integers(16, 40)))
    n = 400
    x = rng.uniform(-1, w + 1, n); y = rng.uniform(-1, h + 1, n)
    ch = {"v": rng.uniform(0, 1, n).astype(np.float32), "d": rng.uniform(-4, 4, n).astype(np.float32),
          "hl": rng.uniform(0, 12, n).astype(np.float32), "s": rng.uniform(0.2, 3, n).astype(np.float32),
          "r": rng.uniform(-3, 3, n).astype(np.float32)}
    specs = [mg.Spec("v", orc.WEIGHTED_AVERAGE, type=orc.GLYPH_LINE, direction_channel="d",
                     half_length_channel="hl", max_radius_cells=9.0),
             mg.Spec("v", orc.SUM, type=orc.GLYPH_GAUSSIAN, sigma_x_channel="s", sigma_y_channel="s",
                     rotation_channel="r", max_radius_cells=7.0),
             mg.Spec("v", orc.COUNT, type=orc.GLYPH_GAUSSIAN, default_sigma_x=2.0, default_sigma_y=0.7,
                     default_rotation=0.4, max_radius_cells=6.0)]
    ref = orc.reference_run(gd, [(x, y, ch)], specs)
    got = oracle.run(gd, [(x, y, ch)], specs)
    _compare(oracle, gd, [(x, y, ch)], specs, ref, got, f"live glyph seed {seed}")


@needs_ref
def test_live_differential_random_configs(oracle, pcr):
    """The 96 seeded random pipelines of tests/test_random_configs_gpu.py (awkward cell sizes, far origins,
    tiny reference tiles, points on every edge, NaN / inf coordinates and values, Line and Gaussian glyphs
    with per-point channels) through the UNMODIFIED reference and through the C oracle: the same cases the
    CUDA path is held to on the GPU box are first used to pin the oracle itself."""
    import test_random_configs_gpu as cases
    from util import compare_bands, grid_desc
    for seed in range(96):
        gc, clouds, specs, _knobs, _loc = cases._random_case(pcr, seed)
        gd = grid_desc(gc)
        ref = orc.reference_run(gd, clouds, specs)
        got = oracle.run(gd, clouds, specs)
        compare_bands(oracle, gd, clouds, specs, ref, got, f"reference vs oracle, seed {seed}")


@needs_ref
def test_live_differential_world_to_cell(pcr):
    """GridConfig.compute_dimensions / world_to_cell of the product library against the reference's own, on
    awkward cell sizes, far origins, exact cell boundaries, both inclusive edges and their neighbours, NaN / inf."""
    ref = orc.load_reference(False)
    rng = np.random.default_rng(0)
    checked = 0
    for _ in range(40):
        cs = float(rng.choice([1.0, 0.5, 0.1, 0.3, 2.7, 1e-3, 1e3, 7.0]))
        csy = -cs * float(rng.choice([1.0, 0.5, 3.0]))
        ox = float(rng.choice([0.0, -1e6 + 0.3, 4.5e6 + 0.1, 12.5]))
        oy = float(rng.choice([0.0, 7e5 + 0.7, -33.25]))
        w, h = int(rng.integers(1, 500)), int(rng.integers(1, 500))
        g, r = pcr.GridConfig(), ref.GridConfig()
        for c in (g, r):
            c.bounds.min_x, c.bounds.min_y = ox, oy
            c.bounds.max_x, c.bounds.max_y = ox + w * cs, oy + h * abs(csy)
            c.cell_size_x, c.cell_size_y = cs, csy
            c.compute_dimensions()
        assert (g.width, g.height, g.tiles_x, g.tiles_y) == (r.width, r.height, r.tiles_x, r.tiles_y)
        xs = np.concatenate([rng.uniform(ox - cs, ox + (w + 1) * cs, 200), ox + cs * rng.integers(0, w + 1, 60),
                             [ox, ox + w * cs, np.nextafter(ox, -np.inf), np.nextafter(ox + w * cs, np.inf),
                              np.nan, np.inf, -np.inf, -0.0]])
        ys = np.concatenate([rng.uniform(oy - cs, oy + (h + 1) * abs(csy), 200), oy + abs(csy) * rng.integers(0, h + 1, 60),
                             [oy, g.bounds.max_y, np.nextafter(oy, -np.inf), np.nextafter(g.bounds.max_y, np.inf),
                              np.nan, np.inf, -np.inf, 0.0]])
        for x, y in zip(xs, rng.permutation(ys)):
            a, b = g.world_to_cell(float(x), float(y)), r.world_to_cell(float(x), float(y))
            assert bool(a[2]) == bool(b[2]) and (not a[2] or (a[0], a[1]) == (b[0], b[1])), (cs, csy, ox, oy, x, y, a, b)
            checked += 1
    assert checked > 10000
