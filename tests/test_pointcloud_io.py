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
        pcr.read_point_cloud_info(str(tmp_path / "x.laz"))            # compressed LAS: as upstream, not implemented
    notlas = tmp_path / "x.las"; notlas.write_bytes(b"\0" * 300)
    with pytest.raises(RuntimeError, match="LASF"):
        pcr.read_point_cloud_info(str(notlas))
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


# ---- LAS (uncompressed): NotImplemented upstream, read here --------------------------------------
def test_las_roundtrip_and_streaming(pcr, tmp_path):
    rng = np.random.default_rng(3)
    n = 5000
    x = np.round(rng.uniform(500000, 500100, n), 3)          # mm grid: the LAS quantisation is then exact
    y = np.round(rng.uniform(4100000, 4100080, n), 3)
    c = pcr.PointCloud.create(n)
    c.set_x_array(x); c.set_y_array(y)
    vals = {"z": np.round(rng.uniform(10, 90, n), 3), "intensity": rng.integers(0, 65535, n),
            "classification": rng.integers(0, 19, n), "return_number": rng.integers(1, 6, n),
            "number_of_returns": rng.integers(1, 6, n)}
    for k, v in vals.items():
        c.add_channel(k, pcr.DataType.Float32); c.set_channel_array_f32(k, v.astype(np.float32))
    c.set_crs(pcr.CRS.from_epsg(32610))
    path = str(tmp_path / "cloud.las")
    pcr.write_point_cloud(path, c, pcr.PointCloudFormat.LAS)
    info = pcr.read_point_cloud_info(path)
    assert info.num_points == n and info.crs.epsg == 32610
    assert [ch.name for ch in info.channels] == ["z", "intensity", "classification", "return_number", "number_of_returns"]
    assert abs(info.bounds.min_x - x.min()) < 1e-6 and abs(info.bounds.max_y - y.max()) < 1e-6
    back = pcr.read_point_cloud(path)                        # format auto-detected from the extension
    assert back.count() == n
    assert np.abs(back.x_array() - x).max() < 1e-6 and np.abs(back.y_array() - y).max() < 1e-6
    for k, v in vals.items():
        assert np.allclose(back.channel_array_f32(k), v.astype(np.float32), atol=2e-3 if k == "z" else 0), k
    r = pcr.PointCloudReader.open(path)
    chunk = pcr.PointCloud.create(1024)
    got, xs = 0, []
    while not r.eof():
        m = r.read_chunk(chunk, 1024)
        xs.append(np.array(chunk.x_array()[:m])); got += m
    assert got == n and np.array_equal(np.concatenate(xs), back.x_array())
    r.rewind()
    assert r.read_chunk(chunk, 10) == 10 and np.array_equal(chunk.x_array()[:10], back.x_array()[:10])


def test_las14_format6_records_built_by_hand(pcr, tmp_path):
    """A LAS 1.4 file with point data record format 6 (30-byte records + 4 bytes of extra data per point),
    assembled field by field from the ASPRS layout — independent of the writer above."""
    import struct
    n = 7
    X = np.arange(n) * 1000 + 5; Y = np.arange(n) * -250 + 17; Z = np.arange(n) * 10
    scale, off = (0.01, 0.01, 0.001), (1000.0, -2000.0, 50.0)
    recs = b""
    for i in range(n):
        recs += struct.pack("<iiiHBBBBhHd", int(X[i]), int(Y[i]), int(Z[i]), 100 + i, (i % 15 + 1) | ((15 - i) << 4),
                            0, 2 + i, 0, -300 + i, 9, 1e5 + i) + b"\xAB\xCD\xEF\x01"
    head = bytearray(375)
    head[0:4] = b"LASF"; head[24], head[25] = 1, 4
    struct.pack_into("<HII", head, 94, 375, 375, 0)
    head[104] = 6
    struct.pack_into("<H", head, 105, 34)
    struct.pack_into("<I", head, 107, 0)                       # legacy count unused in 1.4
    struct.pack_into("<3d", head, 131, *scale)
    struct.pack_into("<3d", head, 155, *off)
    xs, ys = X * scale[0] + off[0], Y * scale[1] + off[1]
    struct.pack_into("<6d", head, 179, xs.max(), xs.min(), ys.max(), ys.min(), 1.0, 0.0)
    struct.pack_into("<Q", head, 247, n)
    path = str(tmp_path / "v14.las")
    open(path, "wb").write(bytes(head) + recs)
    info = pcr.read_point_cloud_info(path)
    assert info.num_points == n and "gps_time" in [c.name for c in info.channels]
    c = pcr.read_point_cloud(path)
    assert np.allclose(c.x_array(), xs, rtol=0, atol=1e-9) and np.allclose(c.y_array(), ys, rtol=0, atol=1e-9)
    assert np.array_equal(c.channel_array_f32("intensity"), 100 + np.arange(n, dtype=np.float32))
    assert np.array_equal(c.channel_array_f32("classification"), 2 + np.arange(n, dtype=np.float32))
    assert np.array_equal(c.channel_array_f32("return_number"), (np.arange(n) % 15 + 1).astype(np.float32))
    assert np.array_equal(c.channel_array_f32("number_of_returns"), (15 - np.arange(n)).astype(np.float32))
    assert np.allclose(c.channel_array_f32("z"), Z * scale[2] + off[2])
    # compressed files are refused, not misread
    bad = bytearray(bytes(head)); bad[104] = 6 | 0x80
    open(str(tmp_path / "c.las"), "wb").write(bytes(bad) + recs)
    with pytest.raises(RuntimeError, match="LAZ"):
        pcr.read_point_cloud_info(str(tmp_path / "c.las"))
    with pytest.raises(RuntimeError):
        pcr.read_point_cloud_info(str(tmp_path / "missing.laz"), pcr.PointCloudFormat.LAZ)


@pytest.mark.gpu
def test_las_ground_points_to_dem(gpu_pcr, tmp_path):
    """The LiDAR use case of the reference's scripts/data/test_dc_lidar.py: a LAS tile streamed into the
    pipeline, ground returns only (classification == 2, FilterSpec on the device), z -> Max/Average/Count."""
    pcr = gpu_pcr
    from util import make_grid, spec
    rng = np.random.default_rng(12)
    n = 60_000
    x = np.round(rng.uniform(323000.0, 323100.0, n), 3)
    y = np.round(rng.uniform(4306000.0, 4306080.0, n), 3)
    z = np.round(rng.uniform(5, 60, n), 3)
    cls = rng.choice([1, 2, 5, 6], n).astype(np.float32)
    c = pcr.PointCloud.create(n); c.set_x_array(x); c.set_y_array(y)
    for k, v in (("z", z), ("classification", cls), ("intensity", rng.integers(0, 4000, n))):
        c.add_channel(k, pcr.DataType.Float32); c.set_channel_array_f32(k, np.asarray(v, np.float32))
    path = str(tmp_path / "tile.las")
    pcr.write_point_cloud(path, c, pcr.PointCloudFormat.LAS)

    cfg = pcr.PipelineConfig()
    cfg.grid = make_grid(pcr, 100, 80, cell=2.0, min_x=323000.0, min_y=4306000.0)
    cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.filter.add("classification", pcr.CompareOp.Equal, 2.0)
    cfg.reductions = [spec(pcr, "z", pcr.ReductionType.Max), spec(pcr, "z", pcr.ReductionType.Average),
                      spec(pcr, "z", pcr.ReductionType.Count)]
    p = pcr.Pipeline.create(cfg)
    r = pcr.PointCloudReader.open(path)
    chunk = pcr.PointCloud.create(8192, pcr.MemoryLocation.HostPinned)
    while not r.eof():
        r.read_chunk(chunk, 8192)
        p.ingest(chunk)
    p.finalize()
    back = pcr.read_point_cloud(path)
    bx, by = np.array(back.x_array()), np.array(back.y_array())
    bz, bc = np.array(back.channel_array_f32("z")), np.array(back.channel_array_f32("classification"))
    g = bc == 2
    col = np.minimum(np.floor((bx[g] - 323000.0) / 2.0).astype(int), 49)
    row = np.minimum(np.floor((by[g] - 4306080.0) / -2.0).astype(int), 39)
    cnt = np.zeros((40, 50)); np.add.at(cnt, (row, col), 1)
    mx = np.full((40, 50), -np.inf, np.float32); np.maximum.at(mx, (row, col), bz[g])
    sm = np.zeros((40, 50)); np.add.at(sm, (row, col), bz[g].astype(np.float64))
    got = [np.array(p.result().band_array(i)) for i in range(3)]
    assert p.stats().points_processed == int(g.sum())
    has = cnt > 0
    assert np.array_equal(np.isnan(got[2]), ~has) and np.array_equal(got[2][has], cnt[has].astype(np.float32))
    assert np.array_equal(got[0][has], mx[has])
    assert np.allclose(got[1][has], (sm[has] / cnt[has]), rtol=1e-5, atol=0)
