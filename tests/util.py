"""Shared helpers for the parity tests: grid/spec builders and seeded cloud generators
(SURVEY §8(d) M1).  Works with the product `pcr` API; the oracle takes the same objects."""
import numpy as np

import oracle as orc


def make_grid(pcr, w, h, cell=1.0, tile=None, min_x=0.0, min_y=0.0, cell_y=None):
    gc = pcr.GridConfig()
    gc.bounds.min_x, gc.bounds.min_y = float(min_x), float(min_y)
    gc.bounds.max_x, gc.bounds.max_y = float(min_x) + w, float(min_y) + h
    gc.cell_size_x = cell
    gc.cell_size_y = -cell if cell_y is None else cell_y
    if tile:
        gc.tile_width = gc.tile_height = tile
    gc.compute_dimensions()
    return gc


def spec(pcr, channel, rtype, name=""):
    s = pcr.ReductionSpec()
    s.value_channel = channel
    s.type = rtype
    s.output_band_name = name
    return s


def cloud(pcr, x, y, chans, loc=None):
    x = np.ascontiguousarray(x, np.float64)
    c = pcr.PointCloud.create(max(len(x), 1), pcr.MemoryLocation.Host if loc is None else loc)
    c.set_x_array(x)
    c.set_y_array(np.ascontiguousarray(y, np.float64))
    for k, v in chans.items():
        c.add_channel(k, pcr.DataType.Float32)
        c.set_channel_array_f32(k, np.ascontiguousarray(v, np.float32))
    return c


def run_product(pcr, gc, clouds, specs, loc=None, **knobs):
    """clouds = [(x, y, {name: f32 array})]; returns list of band arrays (copies)."""
    cfg = pcr.PipelineConfig()
    cfg.grid = gc
    cfg.reductions = list(specs)
    cfg.exec_mode = pcr.ExecutionMode.GPU
    for k, v in knobs.items():
        setattr(cfg, k, v)
    p = pcr.Pipeline.create(cfg)
    assert p is not None, "Pipeline.create failed"
    for (x, y, ch) in clouds:
        c = cloud(pcr, x, y, ch)
        if loc is not None and loc != pcr.MemoryLocation.Host:
            c = c.to_device() if loc == pcr.MemoryLocation.Device else c.to_pinned()
        p.ingest(c)
    p.finalize()
    return [np.array(p.result().band_array(i)) for i in range(len(specs))], p


def uniform_cloud(n, w, h, seed=42, margin=2.0):
    """benchmark_glyph_full.py:62-78 of the reference: U(margin, W-margin), value U(0,1),
    direction U(0,pi)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(margin, w - margin, n)
    y = rng.uniform(margin, h - margin, n)
    ch = {"value": rng.uniform(0, 1, n).astype(np.float32),
          "direction": rng.uniform(0, np.pi, n).astype(np.float32)}
    return x, y, ch


def clustered_cloud(n, w, h, seed=42, k=64):
    """LiDAR-like: K Gaussian clusters, sigma log-uniform, clipped to the bbox so that
    mass sits exactly on max_x / min_y (generate_gaussian_clusters,
    python/pcr/test_generators.py:560-632 of the reference)."""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, w, k)
    cy = rng.uniform(0, h, k)
    sig = np.exp(rng.uniform(np.log(w / 400.0), np.log(w / 10.0), k))
    which = rng.integers(0, k, n)
    x = np.clip(rng.normal(cx[which], sig[which]), 0, w)
    y = np.clip(rng.normal(cy[which], sig[which]), 0, h)
    ch = {"value": (which / k + rng.normal(0, 0.05, n)).astype(np.float32)}
    return x, y, ch


def boundary_cloud(w, h, seed=7, n=4000):
    """Adversarial routing probes: exact edges, corners, one ulp either side of every
    edge, cell boundaries, NaN/inf coordinates, far outside."""
    rng = np.random.default_rng(seed)
    edges_x = [0.0, w, np.nextafter(0.0, -1), np.nextafter(0.0, 1), np.nextafter(w, 0), np.nextafter(w, 2 * w + 1)]
    edges_y = [0.0, h, np.nextafter(0.0, -1), np.nextafter(0.0, 1), np.nextafter(h, 0), np.nextafter(h, 2 * h + 1)]
    xs, ys = [], []
    for ex in edges_x:
        for ey in edges_y:
            xs.append(ex); ys.append(ey)
    for ex in edges_x:
        xs += [ex] * 8; ys += list(rng.uniform(0, h, 8))
    for ey in edges_y:
        ys += [ey] * 8; xs += list(rng.uniform(0, w, 8))
    # integer cell boundaries and their neighbours
    for _ in range(n // 4):
        c = float(rng.integers(0, int(w) + 1)); r = float(rng.integers(0, int(h) + 1))
        xs += [c, np.nextafter(c, -1), np.nextafter(c, c + 1)]
        ys += [r, np.nextafter(r, r + 1), np.nextafter(r, -1)]
    xs += [np.nan, 1.0, np.inf, -np.inf, 1e300, -1e300, w / 2]
    ys += [1.0, np.nan, 1.0, 1.0, 1.0, 1.0, np.inf]
    m = n - len(xs)
    if m > 0:
        xs += list(rng.uniform(-0.1 * w, 1.1 * w, m)); ys += list(rng.uniform(-0.1 * h, 1.1 * h, m))
    x = np.array(xs, np.float64); y = np.array(ys, np.float64)
    v = rng.uniform(-5, 5, len(x)).astype(np.float32)
    return x, y, {"value": v}


def grid_desc(gc):
    return orc.GridDesc.from_config(gc)


def assert_same_nan_mask(a, b, what=""):
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"NaN mask differs {what}"


def assert_exact(a, b, what=""):
    """Bit-level parity for Count/Max/Min: identical NaN mask, identical values
    (+0.0 == -0.0: the reference's own fmaxf picks either, order-dependently)."""
    assert a.shape == b.shape
    assert_same_nan_mask(a, b, what)
    m = ~np.isnan(a)
    assert np.array_equal(a[m], b[m]), f"{what}: {np.sum(a[m] != b[m])} cells differ"


def sum_tolerance(abs64, cnt):
    """Rigorous bound for an fp32 sum of n terms in ANY order vs the exact sum:
    |err| <= (n-1) * u * sum|x_i| (+1 ulp for the final rounding), u = 2^-24.
    Two such sums (ours and the reference's) differ by at most twice that."""
    return 2.0 * (np.maximum(cnt, 1)) * 2.0 ** -24 * abs64 + 1e-30


EXACT_TYPES = (orc.MAX, orc.MIN, orc.COUNT)

# Stated float32 tolerances (BASELINE north_star: "Sum, Average and WeightedAverage must agree
# within a stated float32 relative tolerance"):
#   * any fp32 sum of n terms, in any order, is within (n-1)*2^-24*sum|x_i| of the exact sum, so
#     two implementations differ by at most twice that: sum_tolerance();
#   * device glyph weights (CUDA expf <= 2 ulp; f64->f32 cos/sin vs glibc cosf/sinf <= 1 ulp,
#     amplified by |exponent| <= 13.8 before the 1e-6 cut) are within 2^-18 relative of the
#     reference's: GLYPH_WEIGHT_RTOL * sum|contribution| is added for Line/Gaussian bands.
GLYPH_WEIGHT_RTOL = 2.0 ** -18


def compare_bands(oracle, gd, clouds, specs, ref_bands, got_bands, what, device_weights=False,
                  mismatch_budget=0):
    """ref_bands: reference/oracle output; got_bands: implementation under test.
    Point Count/Max/Min: bit-exact.  Sums/ratios: within the stated bound.  Up to
    `mismatch_budget` cells per band may violate (used only for the Line flip-rate test)."""
    report = []
    for i, (s, ref, got) in enumerate(zip(specs, ref_bands, got_bands)):
        t = int(s.type)
        glyph = int(s.glyph.type)
        tag = f"{what} band {i} (type {t}, glyph {glyph})"
        assert got.shape == ref.shape, tag
        if glyph == orc.GLYPH_POINT and t in EXACT_TYPES:
            assert_exact(got, ref, tag)
            continue
        extra = GLYPH_WEIGHT_RTOL if (device_weights and glyph != orc.GLYPH_POINT) else 0.0
        if t in (orc.AVERAGE, orc.WEIGHTED_AVERAGE):
            s64, a64, cnt = oracle.bounds(gd, clouds, s, want_weight=False)
            w64, wa64, _ = oracle.bounds(gd, clouds, s, want_weight=True)
            ts = sum_tolerance(a64, cnt) + extra * a64
            tw = sum_tolerance(wa64, cnt) + extra * wa64
            with np.errstate(all="ignore"):
                tol = 1.5 * (ts / np.maximum(np.abs(w64), 1e-30)
                             + np.abs(s64) * tw / np.maximum(w64 * w64, 1e-30)) \
                    + np.abs(ref.astype(np.float64)) * 2.0 ** -22
        else:
            _, a64, cnt = oracle.bounds(gd, clouds, s, want_weight=(t == orc.COUNT))
            tol = sum_tolerance(a64, cnt) + extra * a64 + np.abs(ref.astype(np.float64)) * 2.0 ** -23
        if t not in (orc.AVERAGE, orc.WEIGHTED_AVERAGE):
            # Rounding is unbiased, a different accumulation rule is not: the mean SIGNED error over many
            # cells, relative to the summed magnitudes, must vanish (this is what caught the tensor core's
            # truncating accumulate in the Gaussian gather: -6e-5 there).
            m = np.isfinite(ref) & np.isfinite(got) & (a64 > 0)
            if int(m.sum()) >= 1000:
                with np.errstate(all="ignore"):
                    bias = float(np.mean((got[m].astype(np.float64) - ref[m].astype(np.float64)) / a64[m]))
                assert abs(bias) <= 2.0 ** -21, f"{tag}: systematic bias {bias:.3e} relative to sum|x|"
        bad = np.isnan(ref) != np.isnan(got)
        fin = np.isfinite(ref) & np.isfinite(got)
        with np.errstate(all="ignore"):
            err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
        bad |= fin & (err > tol)
        nf = ~np.isfinite(ref) & ~np.isnan(ref)           # +-inf must agree exactly
        bad |= nf & (got != ref)
        nbad = int(bad.sum())
        report.append(nbad)
        assert nbad <= mismatch_budget, \
            f"{tag}: {nbad} cells outside tolerance (budget {mismatch_budget}); " \
            f"first at {np.argwhere(bad)[:3].tolist()}"
    return report
