"""Point-cloud file IO: PCRP binary + CSV + streaming reader, CPU-only tests.
Interop: a PCRP file WRITTEN BY THE REFERENCE is committed under tests/golden/; where oracle/_ref exists
(build container) the reference also reads a file written here."""
import os

import numpy as np
import pytest

import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pcrp_reference_file.npz")


def make(pcr, n=1000, seed=0):
    rng = np.random.default_rng(seed)
    c = pcr.PointCloud.create(n)
    c.set_x_array(rng.uniform(0, 1e6, n)); c.set_y_array(rng.uniform(-1e6, 1e6, n))
    c.add_channel("intensity", pcr.DataType.Float32); c.set_channel_array_f32("intensity", rng.uniform(0, 1, n).astype(np.float32))
    c.add_channel("z", pcr.DataType.Float32); c.set_channel_array_f32("z", rng.normal(0, 5, n).astype(np.float32))
    c.set_crs(pcr.CRS.from_wkt('PROJCS["demo"]'))
    return c


def test_read_reference_written_pcrp(pcr, tmp_path):
    z = np.load(GOLD)
    path = tmp_path / "ref.pcr"
    path.write_bytes(bytes(z["raw_file"]))
    info = pcr.read_point_cloud_info(str(path))
    assert info.num_points == 37 and sorted(c.name for c in info.channels) == ["intensity", "z"]
    assert all(c.dtype == pcr.DataType.Float32 for c in info.channels)
    c = pcr.read_point_cloud(str(path))
    assert c.count() == 37
    assert np.array_equal(c.x_array(), z["x"]) and np.array_equal(c.y_array(), z["y"])
    for k in ("intensity", "z"):
        assert np.array_equal(c.channel_array_f32(k), z["ch_" + k])


def test_pcrp_roundtrip_and_streaming(pcr, tmp_path):
    c = make(pcr, 1000)
    path = str(tmp_path / "a.pcr")
    pcr.write_point_cloud(path, c)
    d = pcr.read_point_cloud(path)
    assert d.count() == 1000 and d.crs().wkt == 'PROJCS["demo"]'
    assert np.array_equal(d.x_array(), c.x_array()) and np.array_equal(d.channel_array_f32("z"), c.channel_array_f32("z"))
    r = pcr.PointCloudReader.open(path)
    assert r.info().num_points == 1000 and not r.eof()
    chunk = pcr.PointCloud.create(300)
    xs, zs, sizes = [], [], []
    while not r.eof():
        n = r.read_chunk(chunk, 300)
        sizes.append(n); xs.append(chunk.x_array().copy()); zs.append(chunk.channel_array_f32("z").copy())
    assert sizes == [300, 300, 300, 100] and r.read_chunk(chunk, 300) == 0
    assert np.array_equal(np.concatenate(xs), c.x_array()) and np.array_equal(np.concatenate(zs), c.channel_array_f32("z"))
    r.rewind()
    assert r.read_chunk(chunk, 5) == 5 and np.array_equal(chunk.y_array(), c.y_array()[:5])


def test_csv_roundtrip(pcr, tmp_path):
    c = make(pcr, 50)
    path = str(tmp_path / "a.csv")
    pcr.write_point_cloud(path, c, pcr.PointCloudFormat.CSV)
    assert open(path).readline().strip() == "x,y,intensity,z"
    d = pcr.read_point_cloud(path)
    assert d.count() == 50 and d.channel("z").dtype == pcr.DataType.Float64       # CSV channels come back as f64
    assert np.allclose(d.x_array(), c.x_array(), rtol=1e-14)


def test_io_errors(pcr, tmp_path):
    bad = tmp_path / "bad.pcr"; bad.write_bytes(b"NOPE" + b"\0" * 40)
    with pytest.raises(RuntimeError, match="invalid magic"):
        pcr.read_point_cloud_info(str(bad), pcr.PointCloudFormat.PCR_Binary)
    with pytest.raises(RuntimeError, match="Failed to open"):
        pcr.PointCloudReader.open(str(tmp_path / "missing.pcr"))
    with pytest.raises(RuntimeError, match="LAS"):
        pcr.read_point_cloud_info(str(tmp_path / "x.las"))
    csv = tmp_path / "c.csv"; csv.write_text("a,b\n1,2\n")
    with pytest.raises(RuntimeError, match="x,y columns"):
        pcr.read_point_cloud_info(str(csv))


@pytest.mark.skipif(not orc.reference_available(), reason="oracle/_ref not built on this box")
def test_reference_reads_our_pcrp(pcr, tmp_path):
    ref = orc.load_reference()
    c = make(pcr, 200, seed=4)
    path = str(tmp_path / "ours.pcr")
    pcr.write_point_cloud(path, c)
    rc = ref.read_point_cloud(path, ref.PointCloudFormat.PCR_Binary)
    assert rc.count() == 200
    assert np.array_equal(np.array(rc.x_array()), c.x_array())
    assert np.array_equal(np.array(rc.channel_array_f32("intensity")), c.channel_array_f32("intensity"))


@pytest.mark.gpu
def test_stream_file_into_pipeline(gpu_pcr, tmp_path):
    """The N4 use case: a file-backed cloud streamed chunk by chunk into Pipeline.ingest."""
    pcr = gpu_pcr
    from util import make_grid, spec
    rng = np.random.default_rng(8)
    n = 100_000
    c = pcr.PointCloud.create(n); c.set_x_array(rng.uniform(0, 64, n)); c.set_y_array(rng.uniform(0, 64, n))
    c.add_channel("v", pcr.DataType.Float32); c.set_channel_array_f32("v", rng.uniform(0, 1, n).astype(np.float32))
    path = str(tmp_path / "big.pcr")
    pcr.write_point_cloud(path, c)
    cfg = pcr.PipelineConfig(); cfg.grid = make_grid(pcr, 64, 64); cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.reductions = [spec(pcr, "v", pcr.ReductionType.Count), spec(pcr, "v", pcr.ReductionType.Max)]
    whole = pcr.Pipeline.create(cfg); whole.ingest(c); whole.finalize()
    streamed = pcr.Pipeline.create(cfg)
    r = pcr.PointCloudReader.open(path)
    chunk = pcr.PointCloud.create(16384, pcr.MemoryLocation.HostPinned)
    while not r.eof():
        r.read_chunk(chunk, 16384)
        streamed.ingest(chunk)
    streamed.finalize()
    for b in range(2):
        assert np.array_equal(np.array(whole.result().band_array(b)), np.array(streamed.result().band_array(b)), equal_nan=True)
    assert streamed.stats().points_processed == n
