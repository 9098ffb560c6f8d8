"""Point filter: FilterSpec predicates evaluated on the device in front of routing.
Known answers from the reference's tests/cpp/test_filter.cpp (the 100-point fixture: intensity = i,
classification = i % 5) and tests/cpp/test_pipeline.cpp:305-352 (DISABLED_WithFilter upstream because
the reference's pipeline integration is broken — its expected result, 50 of 100 points, is the
contract here); differential runs against the oracle fed the numpy-filtered cloud."""
import numpy as np
import pytest

from util import compare_bands, grid_desc, make_grid, spec, cloud

pytestmark = pytest.mark.gpu


def run(pcr, gc, arrays, specs, flt, **knobs):
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.filter = flt
    for k, v in knobs.items():
        setattr(cfg, k, v)
    p = pcr.Pipeline.create(cfg)
    assert p is not None
    p.ingest(cloud(pcr, *arrays))
    p.finalize()
    return [np.array(p.result().band_array(i)) for i in range(len(specs))], p


def fixture100():
    i = np.arange(100)
    return (0.5 + (i % 10), 9.5 - (i // 10),
            {"intensity": i.astype(np.float32), "classification": (i % 5).astype(np.float32)})


@pytest.mark.parametrize("build,expected", [
    (lambda f, C: f.add("classification", C.Equal, 2.0), 20),                      # test_filter.cpp:50-69
    (lambda f, C: f.add("intensity", C.Less, 50.0), 50),                           # :71-89
    (lambda f, C: f.add("intensity", C.GreaterEqual, 75.0), 25),                   # :91-109
    (lambda f, C: f.add_in_set("classification", [1.0, 3.0]), 40),                 # :111-130
    (lambda f, C: f.add("intensity", C.GreaterEqual, 50.0).add("intensity", C.Less, 60.0)
                   .add("classification", C.Equal, 0.0), 2),                       # :157-184
    (lambda f, C: f.add("intensity", C.Greater, 1000.0), 0),                       # :186-199
])
def test_reference_filter_vectors(gpu_pcr, build, expected):
    pcr = gpu_pcr
    f = pcr.FilterSpec()
    build(f, pcr.CompareOp)
    got, p = run(pcr, make_grid(pcr, 10, 10, tile=5), fixture100(), [spec(pcr, "intensity", pcr.ReductionType.Count)], f)
    assert int(np.nansum(got[0])) == expected
    assert p.stats().points_processed == expected           # points_processed += filtered_count


def test_not_in_set(gpu_pcr):
    pcr = gpu_pcr
    f = pcr.FilterSpec()
    p = pcr.FilterPredicate(); p.channel_name = "classification"; p.op = pcr.CompareOp.NotInSet; p.value_set = [0.0, 4.0]
    f.predicates.append(p)                                                          # test_filter.cpp:132-155
    got, _ = run(pcr, make_grid(pcr, 10, 10, tile=5), fixture100(), [spec(pcr, "intensity", pcr.ReductionType.Count)], f)
    assert int(np.nansum(got[0])) == 60


def test_pipeline_with_filter_counts_half(gpu_pcr):
    """test_pipeline.cpp:305-352: classification = idx % 2, keep class 1 -> 50 points counted."""
    pcr = gpu_pcr
    i = np.arange(100)
    arrays = (0.5 + (i % 10), 9.5 - (i // 10), {"intensity": np.ones(100, np.float32),
                                                "classification": (i % 2).astype(np.float32)})
    f = pcr.FilterSpec().add("classification", pcr.CompareOp.Equal, 1.0)
    got, _ = run(pcr, make_grid(pcr, 10, 10, tile=5), arrays, [spec(pcr, "intensity", pcr.ReductionType.Count)], f)
    assert int(np.nansum(got[0])) == 50
    assert np.isnan(got[0]).sum() == 50 and np.nanmax(got[0]) == 1.0


@pytest.mark.parametrize("knobs", [{}, {"deterministic": True}, {"point_kernel": 2}, {"ring_slot_points": 2048}])
def test_filter_differential_vs_oracle(gpu_pcr, oracle, knobs):
    pcr = gpu_pcr
    rng = np.random.default_rng(12)
    n = 60_000
    gc = make_grid(pcr, 120, 90, tile=32)
    x, y = rng.uniform(-2, 122, n), rng.uniform(-2, 92, n)
    ch = {"value": rng.normal(0, 3, n).astype(np.float32), "cls": rng.integers(0, 6, n).astype(np.float32),
          "q": rng.uniform(0, 1, n).astype(np.float32)}
    ch["q"][::17] = np.nan                                   # NaN fails every ordered comparison
    f = pcr.FilterSpec().add_in_set("cls", [1.0, 2.0, 5.0]).add("q", pcr.CompareOp.LessEqual, 0.8) \
        .add("value", pcr.CompareOp.NotEqual, 0.0)
    keep = np.isin(ch["cls"], [1.0, 2.0, 5.0]) & (ch["q"] <= 0.8) & (ch["value"] != 0.0)
    R = pcr.ReductionType
    specs = [spec(pcr, "value", t) for t in (R.Sum, R.Max, R.Min, R.Average, R.Count)]
    if "deterministic" not in knobs:
        specs.append(pcr.line_splat_spec("value", default_direction=0.4, default_half_length=3.0, max_radius_cells=5.0))
        specs.append(pcr.gaussian_splat_spec("value", default_sigma=1.2, max_radius_cells=4.0))
    specs.append(pcr.gaussian_splat_spec("value", default_sigma=2.5, max_radius_cells=9.0))     # gather kernel
    got, p = run(pcr, gc, (x, y, ch), specs, f, **knobs)
    kept = [(x[keep], y[keep], {k: v[keep] for k, v in ch.items()})]
    gd = grid_desc(gc)
    compare_bands(oracle, gd, kept, specs, oracle.run(gd, kept, specs), got, f"filter {knobs}", device_weights=True)
    assert p.stats().points_processed == int(keep.sum())


def test_filter_errors(gpu_pcr):
    pcr = gpu_pcr
    gc = make_grid(pcr, 8, 8)
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.reductions = [spec(pcr, "v", pcr.ReductionType.Sum)]
    cfg.filter = pcr.FilterSpec().add("nonexistent", pcr.CompareOp.Equal, 1.0)
    p = pcr.Pipeline.create(cfg)
    with pytest.raises(RuntimeError, match="filter_points: channel not found: nonexistent"):   # test_filter.cpp:201-212
        p.ingest(cloud(pcr, [1.0], [1.0], {"v": [1.0]}))
    c = pcr.PointCloud.create(2); c.set_x_array(np.ones(2)); c.set_y_array(np.ones(2))
    c.add_channel("v", pcr.DataType.Float32); c.add_channel("nonexistent", pcr.DataType.Int32)
    with pytest.raises(RuntimeError, match="only Float32 channels supported for filtering"):
        p.ingest(c)
