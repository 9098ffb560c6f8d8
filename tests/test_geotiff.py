"""GDAL-free GeoTIFF writer: the file is re-read by an independent pure-Python TIFF
parser written here (tags, tiles, zlib), so the test does not trust the writer's own reader."""
import struct
import zlib

import numpy as np
import pytest


def parse_tiff(path):
    b = open(path, "rb").read()
    assert b[:2] == b"II"
    magic = struct.unpack_from("<H", b, 2)[0]
    big = magic == 43
    assert big or magic == 42
    ifd = struct.unpack_from("<Q", b, 8)[0] if big else struct.unpack_from("<I", b, 4)[0]
    n = struct.unpack_from("<Q", b, ifd)[0] if big else struct.unpack_from("<H", b, ifd)[0]
    esz, base, inl = (20, ifd + 8, 8) if big else (12, ifd + 2, 4)
    fmt = {1: "B", 2: "c", 3: "H", 4: "I", 12: "d", 16: "Q"}
    tags = {}
    prev = 0
    for i in range(n):
        e = base + i * esz
        tag, typ = struct.unpack_from("<HH", b, e)
        assert tag > prev, "IFD entries must be sorted"
        prev = tag
        cnt = struct.unpack_from("<Q", b, e + 4)[0] if big else struct.unpack_from("<I", b, e + 4)[0]
        size = struct.calcsize(fmt[typ]) * cnt
        cell = e + (12 if big else 8)
        off = cell if size <= inl else (struct.unpack_from("<Q", b, cell)[0] if big else struct.unpack_from("<I", b, cell)[0])
        vals = struct.unpack_from("<" + fmt[typ] * cnt, b, off)
        tags[tag] = b"".join(vals).rstrip(b"\0").decode() if typ == 2 else list(vals)
    w, h, nb = tags[256][0], tags[257][0], tags[277][0]
    tw, th = tags[322][0], tags[323][0]
    tx, ty = -(-w // tw), -(-h // th)
    assert len(tags[324]) == tx * ty * nb == len(tags[325])
    assert tags[258] == [32] * nb and tags[339] == [3] * nb
    bands = np.empty((nb, h, w), np.float32)
    k = 0
    for band in range(nb):
        for j in range(ty):
            for i in range(tx):
                raw = b[tags[324][k]:tags[324][k] + tags[325][k]]
                if tags[259][0] == 8:
                    raw = zlib.decompress(raw)
                t = np.frombuffer(raw, "<f4").reshape(th, tw)
                y0, x0 = j * th, i * tw
                bands[band, y0:y0 + th, x0:x0 + tw] = t[:min(th, h - y0), :min(tw, w - x0)]
                k += 1
    return tags, bands


def make(pcr, w, h, nb, seed=0):
    gc = pcr.GridConfig()
    gc.bounds.min_x, gc.bounds.min_y = 500000.0, 4100000.0
    gc.bounds.max_x, gc.bounds.max_y = 500000.0 + w * 0.5, 4100000.0 + h * 0.5
    gc.cell_size_x, gc.cell_size_y = 0.5, -0.5
    gc.crs = pcr.CRS.from_epsg(32610)
    gc.compute_dimensions()
    rng = np.random.default_rng(seed)
    bands = [pcr.BandDesc(f"band <{i}> & co", pcr.DataType.Float32) for i in range(nb)]
    g = pcr.Grid.create(gc.width, gc.height, bands)
    for i in range(nb):
        a = rng.normal(0, 100, (gc.height, gc.width)).astype(np.float32)
        a[rng.random(a.shape) < 0.1] = np.nan
        g.set_band_array(i, a)
    return gc, g


@pytest.mark.parametrize("compress,bigtiff,nb,tile", [("NONE", True, 1, 256), ("NONE", False, 3, 64),
                                                     ("DEFLATE", True, 2, 128), ("DEFLATE", False, 1, 256)])
def test_geotiff_roundtrip(pcr, tmp_path, compress, bigtiff, nb, tile):
    gc, g = make(pcr, 300, 173, nb)
    o = pcr.GeoTiffOptions()
    o.compress, o.bigtiff, o.tile_width, o.tile_height = compress, bigtiff, tile, tile
    path = str(tmp_path / "out.tif")
    pcr.write_geotiff(path, g, gc, o)
    tags, bands = parse_tiff(path)
    assert bands.shape == (nb, 173, 300)
    for i in range(nb):
        assert np.array_equal(bands[i], g.band_array(i), equal_nan=True)
    assert tags[33550] == [0.5, 0.5, 0.0]                                      # ModelPixelScale
    assert tags[33922] == [0, 0, 0, 500000.0, 4100000.0 + 173 * 0.5, 0]        # tiepoint = top-left corner
    keys = tags[34735]
    assert keys[:4] == [1, 1, 0, 3] and [3072, 0, 1, 32610] == keys[12:16] and [1025, 0, 1, 1] == keys[8:12]
    assert tags[42113] == "nan" and "band &lt;0&gt; &amp; co" in tags[42112]
    assert tags[284] == [2 if nb > 1 else 1]
    w, h, n, crs, bb = pcr.read_geotiff_info(path)
    assert (w, h, n, crs.epsg) == (300, 173, nb, 32610)
    assert (bb.min_x, bb.max_x, bb.min_y, bb.max_y) == (gc.bounds.min_x, gc.bounds.max_x, gc.bounds.min_y, gc.bounds.max_y)


def test_geotiff_errors(pcr, tmp_path):
    gc, g = make(pcr, 32, 32, 1)
    o = pcr.GeoTiffOptions()                       # default compress = "LZW" upstream; not supported here
    with pytest.raises(RuntimeError, match="LZW"):
        pcr.write_geotiff(str(tmp_path / "a.tif"), g, gc, o)
    o.compress = "NONE"
    with pytest.raises(RuntimeError, match="failed to create"):
        pcr.write_geotiff(str(tmp_path / "nodir" / "a.tif"), g, gc, o)
    gc.width += 1
    with pytest.raises(RuntimeError, match="dimensions mismatch"):
        pcr.write_geotiff(str(tmp_path / "a.tif"), g, gc, o)
    with pytest.raises(RuntimeError, match="failed to open"):
        pcr.read_geotiff_info(str(tmp_path / "missing.tif"))


@pytest.mark.gpu
def test_pipeline_output_path_writes_geotiff(gpu_pcr, tmp_path):
    from util import make_grid, spec, cloud
    pcr = gpu_pcr
    gc = make_grid(pcr, 100, 80)
    gc.crs = pcr.CRS.from_epsg(4326)
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.reductions = [spec(pcr, "v", pcr.ReductionType.Average), spec(pcr, "v", pcr.ReductionType.Count, "n")]
    for cog in (False, True):
        cfg.output_path = str(tmp_path / f"out{int(cog)}.tif"); cfg.write_cog = cog
        p = pcr.Pipeline.create(cfg)
        rng = np.random.default_rng(1)
        p.run([cloud(pcr, rng.uniform(0, 100, 5000), rng.uniform(0, 80, 5000), {"v": rng.uniform(0, 1, 5000)})])
        tags, bands = parse_tiff(cfg.output_path)
        assert tags[259] == [8 if cog else 1]                         # pipeline.cpp:1352-1354
        for i in range(2):
            assert np.array_equal(bands[i], p.result().band_array(i), equal_nan=True)
        assert tags[34735][4:8] == [1024, 0, 1, 2] and tags[34735][12:16] == [2048, 0, 1, 4326]
        assert "v_3" in tags[42112] and ">n<" in tags[42112]
