"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED
reference (oracle/_ref, built by oracle/Makefile from /root/reference) in CPU mode
with cpu_threads=1.  Run in the build container:

    make -C oracle ref && python oracle/make_golden.py

Each fixture stores the grid, the input clouds, the reduction specs (JSON) and the
reference's output bands.  The fixtures pin (a) the C oracle (tests/test_oracle.py,
CPU) and (b) the CUDA path (tests/test_golden_gpu.py, GPU box, where /root/reference
does not exist).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle as orc  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

SUM, MAX, MIN, AVG, WAVG, COUNT = 0, 1, 2, 3, 4, 5
POINT, LINE, GAUSS = 0, 1, 2


class Glyph:
    def __init__(self, **kw):
        self.type = POINT
        self.direction_channel = ""
        self.default_direction = 0.0
        self.half_length_channel = ""
        self.default_half_length = 1.0
        self.sigma_x_channel = ""
        self.default_sigma_x = 1.0
        self.sigma_y_channel = ""
        self.default_sigma_y = 1.0
        self.rotation_channel = ""
        self.default_rotation = 0.0
        self.max_radius_cells = 32.0
        self.normalize_weights = False
        for k, v in kw.items():
            setattr(self, k, v)


class Spec:
    def __init__(self, value_channel, rtype, name="", **glyph):
        self.value_channel = value_channel
        self.type = rtype
        self.output_band_name = name
        self.glyph = Glyph(**glyph)

    def to_json(self):
        return {"value_channel": self.value_channel, "type": int(self.type),
                "output_band_name": self.output_band_name, "glyph": dict(vars(self.glyph))}

    @staticmethod
    def from_json(d):
        return Spec(d["value_channel"], d["type"], d["output_band_name"], **d["glyph"])


def save(name, gd, clouds, specs, note=""):
    bands = orc.reference_run(gd, clouds, specs, cpu_threads=1)
    arrays = {"grid": np.array([gd.min_x, gd.min_y, gd.max_x, gd.max_y, gd.cell_size_x, gd.cell_size_y,
                                gd.tile_width, gd.tile_height], np.float64),
              "specs": np.frombuffer(json.dumps([s.to_json() for s in specs]).encode(), np.uint8),
              "note": np.frombuffer(note.encode(), np.uint8),
              "n_clouds": np.array([len(clouds)])}
    for i, (x, y, ch) in enumerate(clouds):
        arrays[f"c{i}_x"] = np.asarray(x, np.float64)
        arrays[f"c{i}_y"] = np.asarray(y, np.float64)
        for k, v in ch.items():
            arrays[f"c{i}_ch_{k}"] = np.asarray(v, np.float32)
    for i, b in enumerate(bands):
        arrays[f"band{i}"] = b
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print(f"{name}: {len(specs)} bands, {sum(len(c[0]) for c in clouds)} points")


def load(path):
    """-> (GridDesc, clouds, specs, bands)"""
    z = np.load(path)
    g = z["grid"]
    gd = orc.GridDesc(g[0], g[1], g[2], g[3], g[4], g[5], int(g[6]), int(g[7]))
    specs = [Spec.from_json(d) for d in json.loads(bytes(z["specs"]).decode())]
    clouds = []
    for i in range(int(z["n_clouds"][0])):
        ch = {k[len(f"c{i}_ch_"):]: z[k] for k in z.files if k.startswith(f"c{i}_ch_")}
        clouds.append((z[f"c{i}_x"], z[f"c{i}_y"], ch))
    bands = [z[f"band{i}"] for i in range(len(specs))]
    return gd, clouds, specs, bands


def boundary_points(w, h, rng, n):
    ex = [0.0, w, np.nextafter(0.0, -1), np.nextafter(0.0, 1), np.nextafter(w, 0), np.nextafter(w, 3 * w)]
    ey = [0.0, h, np.nextafter(0.0, -1), np.nextafter(0.0, 1), np.nextafter(h, 0), np.nextafter(h, 3 * h)]
    xs = [a for a in ex for _ in ey] + [np.nan, 1.0, np.inf, -np.inf]
    ys = [b for _ in ex for b in ey] + [1.0, np.nan, 1.0, 1.0]
    for _ in range(n):
        c = float(rng.integers(0, int(w) + 1)); r = float(rng.integers(0, int(h) + 1))
        xs += [c, np.nextafter(c, -1), rng.uniform(0, w)]
        ys += [rng.uniform(0, h), r, np.nextafter(r, r + 1)]
    return np.array(xs), np.array(ys)


def main():
    rng = np.random.default_rng(2024)
    all_point = [Spec("v", t) for t in (SUM, MAX, MIN, AVG, WAVG, COUNT)]

    # --- SURVEY Appendix B probes -------------------------------------------------
    save("probe_inclusive_bounds", orc.GridDesc(0, 0, 4, 4),
         [([4, 0, 2, 4, -0.5, 2], [2, 0, 4, 4, 2, 4.5], {"v": [1] * 6})], [Spec("v", COUNT)],
         "R1: inclusive max edges are clamped into the last row/col; outside points dropped")
    save("probe_touched_tile", orc.GridDesc(0, 0, 4, 4, tile_width=2, tile_height=2),
         [([0.5], [3.5], {"v": [7.0]})], all_point,
         "R11: untouched tiles are NaN for every op, touched tiles follow Op::finalize")
    save("probe_wavg_is_avg", orc.GridDesc(0, 0, 4, 4),
         [([0.5, 0.6], [3.5, 3.4], {"v": [1.0, 3.0]})], [Spec("v", WAVG), Spec("v", MAX), Spec("v", AVG)],
         "WeightedAverage(Point) == Average; Max of an empty cell in a touched tile is NaN")

    # --- Point glyph ----------------------------------------------------------------
    w, h = 64, 48
    bx, by = boundary_points(w, h, rng, 300)
    ux, uy = rng.uniform(-3, w + 3, 3000), rng.uniform(-3, h + 3, 3000)
    x = np.concatenate([bx, ux]); y = np.concatenate([by, uy])
    v = rng.uniform(-10, 10, len(x)).astype(np.float32)
    v[::97] = np.nan                      # NaN values: ignored by Max/Min, poison Sum/Average
    v[5::131] = -np.inf
    v[7::137] = np.float32(-3.4028235e38)   # genuine -FLT_MAX finalizes to NaN for Max
    save("point_all_ops_tiles16", orc.GridDesc(0, 0, w, h, tile_width=16, tile_height=16),
         [(x, y, {"v": v})], all_point, "every Point reducer, 4x3 reference tiles, boundary + NaN/inf values")

    x2, y2 = rng.uniform(0, w, 2000), rng.uniform(0, h / 3, 2000)
    v2 = rng.uniform(0, 1, 2000).astype(np.float32)
    save("point_two_clouds", orc.GridDesc(0, 0, w, h, tile_width=16, tile_height=16),
         [(x2, y2, {"v": v2}), (x2[::-1] * 0.5, y2[::-1], {"v": v2 + 1})], all_point,
         "two ingests accumulate; tiles never hit stay NaN (test_pipeline.cpp:235-303)")

    gd = orc.GridDesc(1000.25, -500.5, 1000.25 + 19.1, -500.5 + 33.3, 0.3, -0.7, 16, 16)
    x3 = rng.uniform(gd.min_x - 1, gd.max_x + 1, 4000); y3 = rng.uniform(gd.min_y - 1, gd.max_y + 1, 4000)
    x3[:4] = [gd.min_x, gd.max_x, gd.min_x, gd.max_x]; y3[:4] = [gd.min_y, gd.min_y, gd.max_y, gd.max_y]
    k = np.arange(1, 40); x3[10:49] = gd.min_x + k * 0.3; y3[60:99] = gd.max_y - k * 0.7
    save("point_nonpow2_cells", gd, [(x3, y3, {"v": rng.uniform(0, 100, 4000).astype(np.float32)})],
         [Spec("v", COUNT), Spec("v", MAX), Spec("v", MIN), Spec("v", SUM)],
         "cell size 0.3 x -0.7 with offset bounds: true f64 division path, ceil() dimensions")

    two_ch = [Spec("a", SUM), Spec("b", SUM), Spec("a", MAX), Spec("b", MIN), Spec("a", AVG), Spec("b", COUNT)]
    save("point_two_channels", orc.GridDesc(0, 0, 32, 32, tile_width=8, tile_height=8),
         [(rng.uniform(0, 32, 1500), rng.uniform(0, 32, 1500),
           {"a": rng.normal(0, 3, 1500).astype(np.float32), "b": rng.normal(5, 1, 1500).astype(np.float32)})],
         two_ch, "reductions over two different value channels in one pipeline")

    # --- Line glyph -------------------------------------------------------------------
    lw, lh = 96, 80
    n = 1500
    lx, ly = rng.uniform(-1, lw + 1, n), rng.uniform(-1, lh + 1, n)
    lch = {"v": rng.uniform(0, 1, n).astype(np.float32),
           "dir": rng.uniform(-np.pi, 2 * np.pi, n).astype(np.float32),
           "hl": rng.uniform(0, 20, n).astype(np.float32)}
    line = dict(type=LINE, direction_channel="dir", half_length_channel="hl", max_radius_cells=18.0)
    save("line_channels_tiles32", orc.GridDesc(0, 0, lw, lh, tile_width=32, tile_height=32),
         [(lx, ly, lch)], [Spec("v", t, **line) for t in (WAVG, SUM, COUNT, AVG)],
         "Line with per-point direction/half_length, clipped at reference tile seams (R8, R9)")
    save("line_defaults", orc.GridDesc(0, 0, lw, lh),
         [(lx, ly, {"v": lch["v"]})],
         [Spec("v", WAVG, type=LINE, default_direction=0.7, default_half_length=6.0, max_radius_cells=8.0),
          Spec("v", COUNT, type=LINE, default_direction=0.0, default_half_length=2.5, max_radius_cells=4.0)],
         "Line with defaults only; two different line footprints in one pipeline")
    save("probe_line_cap", orc.GridDesc(0, 0, 41, 41),
         [([20.5], [20.5], {"v": [1.0]})],
         [Spec("v", COUNT, type=LINE, default_direction=float(np.float32(np.pi / 2)), default_half_length=10.0, max_radius_cells=3.0),
          Spec("v", COUNT, type=LINE, default_direction=0.0, default_half_length=10.0, max_radius_cells=3.0)],
         "R8: with cell_size_y<0 the Y half-length is negative and never capped (21 rows vs 7 cols)")
    gdl = orc.GridDesc(10.0, 20.0, 10.0 + 30.0, 20.0 + 25.0, 0.5, -0.25, 24, 40)
    save("line_nonunit_cells", gdl,
         [(rng.uniform(10, 40, 800), rng.uniform(20, 45, 800),
           {"v": rng.uniform(0, 1, 800).astype(np.float32), "dir": rng.uniform(0, np.pi, 800).astype(np.float32)})],
         [Spec("v", WAVG, type=LINE, direction_channel="dir", default_half_length=1.5, max_radius_cells=12.0)],
         "Line on 0.5 x -0.25 cells (world->cell scaling of half_length differs per axis)")

    # --- Gaussian glyph ---------------------------------------------------------------
    gw, gh = 72, 64
    n = 600
    gx, gy = rng.uniform(-1, gw + 1, n), rng.uniform(-1, gh + 1, n)
    sig = rng.uniform(-0.5, 4.0, n).astype(np.float32)      # <= 0 falls back to the default
    gch = {"v": rng.uniform(0, 1, n).astype(np.float32), "s": sig,
           "s2": rng.uniform(0.3, 2.0, n).astype(np.float32),
           "rot": rng.uniform(-np.pi, np.pi, n).astype(np.float32)}
    save("gauss_sigma_channel_tiles32", orc.GridDesc(0, 0, gw, gh, tile_width=32, tile_height=32),
         [(gx, gy, gch)],
         [Spec("v", t, type=GAUSS, sigma_x_channel="s", sigma_y_channel="s", default_sigma_x=1.5,
               default_sigma_y=1.5, max_radius_cells=9.0) for t in (WAVG, SUM, COUNT, AVG)],
         "Gaussian with one per-point sigma channel for both axes, tile-seam clipping (R9, R10)")
    save("gauss_aniso_rotation", orc.GridDesc(0, 0, gw, gh),
         [(gx, gy, gch)],
         [Spec("v", WAVG, type=GAUSS, sigma_x_channel="s", sigma_y_channel="s2", rotation_channel="rot",
               default_sigma_x=2.0, default_sigma_y=1.0, max_radius_cells=10.0),
          Spec("v", COUNT, type=GAUSS, default_sigma_x=3.0, default_sigma_y=0.5, default_rotation=0.6,
               max_radius_cells=32.0)],
         "anisotropic sigma + per-point rotation; a defaults-only rotated footprint")
    save("probe_gauss_radius", orc.GridDesc(0, 0, 41, 41),
         [([20.5], [20.5], {"v": [1.0]})],
         [Spec("v", COUNT, type=GAUSS, default_sigma_x=0.5, default_sigma_y=4.0),
          Spec("v", COUNT, type=GAUSS, default_sigma_x=4.0, default_sigma_y=0.5)],
         "R10: with cell_size_y<0 the footprint radius depends on sigma_x only; corner-sampled weight")
    gdg = orc.GridDesc(-8.0, -8.0, 8.0, 8.0, 0.25, -0.5, 20, 12)
    save("gauss_nonunit_cells", gdg,
         [(rng.uniform(-8, 8, 300), rng.uniform(-8, 8, 300), {"v": rng.uniform(0, 1, 300).astype(np.float32)})],
         [Spec("v", WAVG, type=GAUSS, default_sigma_x=0.6, default_sigma_y=0.9, max_radius_cells=16.0)],
         "Gaussian on 0.25 x -0.5 cells with small reference tiles")

    # --- Mixed glyphs in one pipeline ---------------------------------------------------
    save("mixed_glyphs", orc.GridDesc(0, 0, gw, gh, tile_width=32, tile_height=32),
         [(gx, gy, gch)],
         [Spec("v", AVG), Spec("v", MAX),
          Spec("v", SUM, type=LINE, default_direction=1.0, default_half_length=4.0, max_radius_cells=6.0),
          Spec("v", WAVG, type=GAUSS, default_sigma_x=1.2, default_sigma_y=1.2, max_radius_cells=5.0),
          Spec("s2", MIN)],
         "Point + Line + Gaussian reductions side by side in one pipeline")


def wide_gaussian_fixture():
    """Footprint radii beyond 16 cells (sigma up to 12, cap 32): 65 x 65-cell footprints across several
    reference tiles — the range the per-bin GEMM kernel covers with its 96-cell neighbourhood.  Own rng: added
    after the other fixtures were generated, which keep their random streams."""
    rng = np.random.default_rng(2026)
    w, h, n = 160, 128, 500
    x, y = rng.uniform(-1, w + 1, n), rng.uniform(-1, h + 1, n)
    x[:6] = [0.0, w, w / 2, 63.999, 64.0, 64.001]
    y[:6] = [0.0, h, h, 64.0, 63.999, 0.25]
    ch = {"v": rng.uniform(-1, 1, n).astype(np.float32), "s": rng.uniform(3.0, 12.0, n).astype(np.float32)}
    save("gauss_wide_sigma_tiles64", orc.GridDesc(0, 0, w, h, tile_width=64, tile_height=64), [(x, y, ch)],
         [Spec("v", t, type=GAUSS, sigma_x_channel="s", sigma_y_channel="s", default_sigma_x=4.0,
               default_sigma_y=4.0, max_radius_cells=32.0) for t in (WAVG, SUM, COUNT)],
         "wide Gaussians (radius 9..32 cells) clipped at 64-cell reference tile seams")


def pcrt_fixtures():
    """Reference-written .pcrt tile-state files (the reference flushes every dirty tile to state_dir at
    finalize, tile_manager.cpp:416-426): one single-reduction pipeline per op, so that one file name per
    tile is unambiguous.  Stored: inputs, the reference bands, and the raw bytes of every tile file."""
    import glob
    import shutil
    import tempfile
    ref = orc.load_reference()
    rng = np.random.default_rng(77)
    gd = orc.GridDesc(0, 0, 40, 24, tile_width=16, tile_height=16)      # 3 x 2 tiles, clipped edge tiles
    n = 700
    x = rng.uniform(0, 30, n); y = rng.uniform(0, 24, n)                 # the east tile column stays untouched
    v = rng.normal(1, 4, n).astype(np.float32)
    for name, t in (("sum", SUM), ("max", MAX), ("min", MIN), ("count", COUNT), ("average", AVG), ("wavg", WAVG)):
        cfg = ref.PipelineConfig()
        cfg.grid = orc.reference_grid(ref, gd)
        cfg.reductions = [orc.to_reference_spec(ref, Spec("v", t))]
        cfg.exec_mode = ref.ExecutionMode.CPU
        cfg.cpu_threads = 1
        tmp = tempfile.mkdtemp(prefix="pcr_pcrt_")
        cfg.state_dir = tmp
        p = ref.Pipeline.create(cfg)
        p.ingest(orc.reference_cloud(ref, x, y, {"v": v}))
        p.finalize()
        band = np.array(p.result().band_array(0))
        files = {os.path.basename(f): np.frombuffer(open(f, "rb").read(), np.uint8) for f in sorted(glob.glob(tmp + "/*.pcrt"))}
        shutil.rmtree(tmp)
        np.savez_compressed(os.path.join(OUT, f"pcrt_{name}.npz"), grid=np.array([0, 0, 40, 24, 1, -1, 16, 16], np.float64),
                            x=x, y=y, v=v, rtype=np.array([t]), band=band,
                            **{"file_" + k[:-5]: b for k, b in files.items()})
        print(f"pcrt_{name}: {len(files)} tile files")


def pcrp_fixture():
    """A PCRP point-cloud file written by the reference's write_point_cloud (point_cloud_io.cpp:74-148)."""
    import tempfile
    ref = orc.load_reference()
    rng = np.random.default_rng(5)
    n = 37
    x, y = rng.uniform(0, 100, n), rng.uniform(-50, 50, n)
    ch = {"intensity": rng.uniform(0, 255, n).astype(np.float32), "z": rng.normal(10, 2, n).astype(np.float32)}
    c = orc.reference_cloud(ref, x, y, ch)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "cloud.pcr")
        ref.write_point_cloud(path, c, ref.PointCloudFormat.PCR_Binary)
        raw = np.frombuffer(open(path, "rb").read(), np.uint8)
    np.savez_compressed(os.path.join(OUT, "pcrp_reference_file.npz"), raw_file=raw, x=x, y=y, **{"ch_" + k: v for k, v in ch.items()})
    print(f"pcrp_reference_file: {raw.size} bytes")


if __name__ == "__main__":
    if "--wide-gauss-only" in sys.argv:
        wide_gaussian_fixture()
        sys.exit(0)
    if "--pcrt-only" not in sys.argv and "--io-only" not in sys.argv:
        main()
        wide_gaussian_fixture()
    if "--io-only" not in sys.argv:
        pcrt_fixtures()
    pcrp_fixture()
