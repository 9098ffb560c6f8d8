"""GPU tests of deterministic mode 2 (exact fixed-point accumulation, csrc/exact_acc.cuh): every additive band
is the correctly rounded EXACT sum of its float32 contributions, hence bit-identical whatever the point order,
the ingest chunking and (tests/test_multi_rank.py) the number of GPUs — for the Point, Line and Gaussian glyphs."""
from fractions import Fraction

import numpy as np
import pytest

from util import compare_bands, grid_desc, make_grid, run_product, spec, uniform_cloud, cloud as mk

pytestmark = pytest.mark.gpu


def round_to_f32_bits(n):
    """Exact integer n (value = n * 2^-149) -> float32 bit pattern, round to nearest even."""
    sign = 0x80000000 if n < 0 else 0
    n = abs(n)
    if n == 0:
        return 0
    p = n.bit_length() - 1
    if p <= 23:
        return sign | n
    shift = p - 23
    top = n >> shift
    rem = n & ((1 << shift) - 1)
    half = 1 << (shift - 1)
    if rem > half or (rem == half and (top & 1)):
        top += 1
    bits = (shift << 23) + top
    return sign | (0x7f800000 if bits >= 0x7f800000 else bits)


def test_sum_is_the_correctly_rounded_exact_sum(gpu_pcr):
    pcr = gpu_pcr
    gc = make_grid(pcr, 8, 8)
    rng = np.random.default_rng(1)
    n = 40_000
    # magnitudes from subnormal to 1e30, both signs, heavy cancellation; 64 cells
    expo = rng.uniform(-44, 30, n)
    v = (rng.choice([-1.0, 1.0], n) * 10.0 ** expo).astype(np.float32)
    v[:100] = np.float32(1e-45)                      # subnormals
    v[100:200] = np.float32(-3.0e38)
    v[200:300] = np.float32(3.0e38)                  # cancels the line above exactly
    x, y = rng.uniform(0, 8, n), rng.uniform(0, 8, n)
    R = pcr.ReductionType
    specs = [spec(pcr, "value", R.Sum), spec(pcr, "value", R.Count), spec(pcr, "value", R.Average)]
    got, _ = run_product(pcr, gc, [(x, y, {"value": v})], specs, deterministic=2)
    col = np.clip(np.floor(x).astype(int), 0, 7)
    row = np.clip(np.floor((y - 8.0) / -1.0).astype(int), 0, 7)
    for r in range(8):
        for c in range(8):
            m = (row == r) & (col == c)
            total = sum(int(Fraction(float(t)) * 2 ** 149) for t in v[m])
            want = np.array([round_to_f32_bits(total)], np.uint32).view(np.float32)[0]
            assert got[0][r, c].view(np.uint32) == want.view(np.uint32), (r, c, got[0][r, c], want)
            assert got[1][r, c] == np.float32(m.sum())
            assert got[2][r, c].view(np.uint32) == np.float32(want / np.float32(m.sum())).view(np.uint32)


def _glyph_specs(pcr):
    R = pcr.ReductionType
    specs = [spec(pcr, "value", t) for t in (R.Sum, R.Max, R.Min, R.Average, R.Count)]
    specs.append(pcr.line_splat_spec("value", "direction", default_half_length=6.0, max_radius_cells=9.0))
    specs.append(pcr.gaussian_splat_spec("value", "sigma", "sigma", default_sigma=1.5, max_radius_cells=6.0))
    rot = pcr.gaussian_splat_spec("value", "sigma", default_sigma_y=2.5, rotation_channel="direction", max_radius_cells=7.0)
    rot.type = R.Sum
    specs.append(rot)
    return specs


def _cloud(n, w, h, seed):
    x, y, ch = uniform_cloud(n, w, h, seed=seed, margin=-2.0)
    rng = np.random.default_rng(seed + 1)
    ch["value"] = rng.normal(0, 50, n).astype(np.float32)
    ch["sigma"] = rng.uniform(0.4, 2.0, n).astype(np.float32)
    return x, y, ch


def test_bits_do_not_depend_on_order_or_chunking_and_match_the_oracle(gpu_pcr, oracle):
    pcr = gpu_pcr
    w, h = 160, 120
    gc = make_grid(pcr, w, h, tile=64)
    x, y, ch = _cloud(60_000, w, h, 5)
    specs = _glyph_specs(pcr)
    one, _ = run_product(pcr, gc, [(x, y, ch)], specs, deterministic=2)
    perm = np.random.default_rng(9).permutation(len(x))
    cuts = [0, 1, 777, 20_000, 20_001, 45_000, len(x)]
    clouds = [(x[perm[a:b]], y[perm[a:b]], {k: v[perm[a:b]] for k, v in ch.items()}) for a, b in zip(cuts, cuts[1:])]
    many, _ = run_product(pcr, gc, clouds, specs, deterministic=2, loc=pcr.MemoryLocation.Device)
    for i, (a, b) in enumerate(zip(one, many)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"band {i} depends on the order of the points"
    gd = grid_desc(gc)
    ref = oracle.run(gd, [(x, y, ch)], specs)
    compare_bands(oracle, gd, [(x, y, ch)], specs, ref, one, "deterministic mode 2", device_weights=True)


def test_nan_and_inf_follow_float_semantics(gpu_pcr):
    pcr = gpu_pcr
    gc = make_grid(pcr, 4, 1)
    x = np.array([0.5, 0.5, 1.5, 1.5, 2.5, 2.5, 3.5], np.float64)
    y = np.full(7, 0.5)
    v = np.array([1.0, np.nan, np.inf, -np.inf, np.inf, 5.0, 7.0], np.float32)
    got, _ = run_product(pcr, gc, [(x, y, {"value": v})], [spec(pcr, "value", pcr.ReductionType.Sum)], deterministic=2)
    assert np.isnan(got[0][0, 0]) and np.isnan(got[0][0, 1]) and got[0][0, 2] == np.inf and got[0][0, 3] == 7.0
