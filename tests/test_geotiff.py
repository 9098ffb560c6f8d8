"""GDAL-free GeoTIFF writer: the file is re-read by an independent pure-Python TIFF
parser written here (tags, tiles, zlib), so the test does not trust the writer's own reader."""
import struct
import zlib

import numpy as np
import pytest


def lzw_decode(data):
    """TIFF 6.0 LZW, written independently of the writer: MSB-first codes, 9..12 bits, Clear 256, EOI 257,
    width bumped one code early."""
    out = bytearray()
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    width, prev, acc, nbits = 9, None, 0, 0
    for byte in data:
        acc = (acc << 8) | byte
        nbits += 8
        while nbits >= width:
            code = (acc >> (nbits - width)) & ((1 << width) - 1)
            nbits -= width
            if code == 257:
                return bytes(out)
            if code == 256:
                table = table[:258]
                width, prev = 9, None
                continue
            if prev is None:
                entry = table[code]
            elif code < len(table):
                entry = table[code]
                table.append(prev + entry[:1])
            else:
                assert code == len(table)
                entry = prev + prev[:1]
                table.append(entry)
            out += entry
            prev = entry
            if len(table) >= (1 << width) - 1 and width < 12:
                width += 1
    return bytes(out)


def parse_tiff(path, image=0):
    b = open(path, "rb").read()
    assert b[:2] == b"II"
    magic = struct.unpack_from("<H", b, 2)[0]
    big = magic == 43
    assert big or magic == 42
    ifd = struct.unpack_from("<Q", b, 8)[0] if big else struct.unpack_from("<I", b, 4)[0]
    for _ in range(image):                                   # follow the IFD chain to overview `image`
        n = struct.unpack_from("<Q", b, ifd)[0] if big else struct.unpack_from("<H", b, ifd)[0]
        nxt = ifd + (8 + 20 * n if big else 2 + 12 * n)
        ifd = struct.unpack_from("<Q", b, nxt)[0] if big else struct.unpack_from("<I", b, nxt)[0]
        assert ifd != 0, "no such overview"
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
    nxt = ifd + (8 + 20 * n if big else 2 + 12 * n)
    tags["next_ifd"] = struct.unpack_from("<Q", b, nxt)[0] if big else struct.unpack_from("<I", b, nxt)[0]
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
                elif tags[259][0] == 5:
                    raw = lzw_decode(raw)
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
                                                     ("DEFLATE", True, 2, 128), ("DEFLATE", False, 1, 256),
                                                     ("LZW", True, 2, 64), ("LZW", False, 1, 256)])
def test_geotiff_roundtrip(pcr, tmp_path, compress, bigtiff, nb, tile):
    if compress == "LZW":                 # the independent LZW decoder above is pure Python: keep its input small
        return _roundtrip_small_lzw(pcr, tmp_path, bigtiff, nb, tile)
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
    for i in range(nb):                                    # the library's own band reader (read_geotiff_band)
        assert np.array_equal(pcr.read_geotiff_band(path, i, 300, 173), g.band_array(i), equal_nan=True)
    w, h, n, crs, bb = pcr.read_geotiff_info(path)
    assert (w, h, n, crs.epsg) == (300, 173, nb, 32610)
    assert (bb.min_x, bb.max_x, bb.min_y, bb.max_y) == (gc.bounds.min_x, gc.bounds.max_x, gc.bounds.min_y, gc.bounds.max_y)


def _roundtrip_small_lzw(pcr, tmp_path, bigtiff, nb, tile):
    gc, g = make(pcr, 150, 97, nb)
    o = pcr.GeoTiffOptions()
    o.compress, o.bigtiff, o.tile_width, o.tile_height = "LZW", bigtiff, tile, tile
    path = str(tmp_path / "out.tif")
    pcr.write_geotiff(path, g, gc, o)
    tags, bands = parse_tiff(path)
    assert tags[259] == [5]
    for i in range(nb):
        assert np.array_equal(bands[i], g.band_array(i), equal_nan=True)
        assert np.array_equal(pcr.read_geotiff_band(path, i, 150, 97), g.band_array(i), equal_nan=True)


def test_geotiff_errors(pcr, tmp_path):
    gc, g = make(pcr, 32, 32, 1)
    o = pcr.GeoTiffOptions()
    assert o.compress == "LZW"                     # the reference's default (grid_io.h:18) is written as such
    pcr.write_geotiff(str(tmp_path / "lzw.tif"), g, gc, o)
    o.compress = "ZSTD"
    with pytest.raises(RuntimeError, match="ZSTD"):
        pcr.write_geotiff(str(tmp_path / "a.tif"), g, gc, o)
    with pytest.raises(RuntimeError, match="dimension mismatch"):
        pcr.read_geotiff_band(str(tmp_path / "lzw.tif"), 0, 31, 32)
    with pytest.raises(RuntimeError, match="out of range"):
        pcr.read_geotiff_band(str(tmp_path / "lzw.tif"), 1, 32, 32)
    o.compress = "NONE"
    with pytest.raises(RuntimeError, match="failed to create"):
        pcr.write_geotiff(str(tmp_path / "nodir" / "a.tif"), g, gc, o)
    gc.width += 1
    with pytest.raises(RuntimeError, match="dimensions mismatch"):
        pcr.write_geotiff(str(tmp_path / "a.tif"), g, gc, o)
    with pytest.raises(RuntimeError, match="failed to open"):
        pcr.read_geotiff_info(str(tmp_path / "missing.tif"))


def test_lzw_handles_repetitive_and_random_tiles(pcr, tmp_path):
    """Dictionary resets (4094 entries) and the KwKwK case: constant, ramp and random rasters."""
    gc, g = make(pcr, 160, 160, 3)
    a = np.zeros((160, 160), np.float32)
    g.set_band_array(0, a)                                                       # one long run
    g.set_band_array(1, np.tile(np.arange(160, dtype=np.float32), (160, 1)))    # ramp
    o = pcr.GeoTiffOptions(); o.compress = "LZW"; o.tile_width = o.tile_height = 160
    path = str(tmp_path / "lzw.tif")
    pcr.write_geotiff(path, g, gc, o)
    _, bands = parse_tiff(path)
    for i in range(3):
        assert np.array_equal(bands[i], g.band_array(i), equal_nan=True)
        assert np.array_equal(pcr.read_geotiff_band(path, i, 160, 160), g.band_array(i), equal_nan=True)
    # a large raster through the library's own decoder only (many dictionary resets per tile)
    gc, g = make(pcr, 700, 520, 2)
    o.tile_width = o.tile_height = 256
    pcr.write_geotiff(path, g, gc, o)
    for i in range(2):
        assert np.array_equal(pcr.read_geotiff_band(path, i, 700, 520), g.band_array(i), equal_nan=True)


def test_cloud_optimized_writes_the_reference_overview_pyramid(pcr, tmp_path):
    """cloud_optimized: levels 2, 4, ... while min(w, h) / level >= 256 (grid_io.cpp:155-176), NaN-aware average."""
    gc, g = make(pcr, 1100, 1030, 2)
    o = pcr.GeoTiffOptions(); o.compress = "DEFLATE"; o.cloud_optimized = True
    path = str(tmp_path / "cog.tif")
    pcr.write_geotiff(path, g, gc, o)
    tags0, full = parse_tiff(path)
    assert tags0["next_ifd"] != 0 and 254 not in tags0
    tags1, ov2 = parse_tiff(path, 1)
    tags2, ov4 = parse_tiff(path, 2)
    assert tags1[254] == [1] and tags2[254] == [1] and tags2["next_ifd"] == 0     # 1030 / 8 < 256: two levels
    assert ov2.shape == (2, 515, 550) and ov4.shape == (2, 258, 275)
    a = full[0]
    blk = np.stack([a[0:1030:2, 0:1100:2], a[1:1030:2, 0:1100:2], a[0:1030:2, 1:1100:2], a[1:1030:2, 1:1100:2]])
    with np.errstate(all="ignore"):
        want = np.nanmean(blk.astype(np.float64), axis=0).astype(np.float32)
    assert np.array_equal(np.isnan(want), np.isnan(ov2[0]))
    np.testing.assert_allclose(ov2[0][~np.isnan(want)], want[~np.isnan(want)], rtol=1e-6)
    # without the flag there is exactly one image
    o.cloud_optimized = False
    pcr.write_geotiff(path, g, gc, o)
    assert parse_tiff(path)[0]["next_ifd"] == 0


def test_tiled_geotiff_writer(pcr, tmp_path):
    """TiledGeoTiffWriter (grid_io.h:44-70): reference tiles written one at a time; unwritten tiles stay NaN."""
    gc, g = make(pcr, 300, 173, 2)
    gc.tile_width, gc.tile_height = 128, 64
    gc.compute_dimensions()
    o = pcr.GeoTiffOptions(); o.compress = "DEFLATE"
    path = str(tmp_path / "tiled.tif")
    w = pcr.TiledGeoTiffWriter.open(path, gc, ["a", "b"], o)
    assert w is not None
    skipped = (1, 2)
    for tr in range(gc.tiles_y):
        for tc in range(gc.tiles_x):
            if (tr, tc) == skipped:
                continue
            c0, r0, cols, rows = gc.tile_cell_range(pcr.TileIndex(tr, tc))
            data = np.stack([g.band_array(b)[r0:r0 + rows, c0:c0 + cols] for b in range(2)])
            w.write_tile(pcr.TileIndex(tr, tc), data, 2)
    with pytest.raises(RuntimeError, match="band count"):
        w.write_tile(pcr.TileIndex(0, 0), np.zeros((1, 64, 128), np.float32), 1)
    w.close()
    tags, bands = parse_tiff(path)
    c0, r0, cols, rows = gc.tile_cell_range(pcr.TileIndex(*skipped))
    for b in range(2):
        want = np.array(g.band_array(b))
        want[r0:r0 + rows, c0:c0 + cols] = np.nan
        assert np.array_equal(bands[b], want, equal_nan=True)
    assert ">a<" in tags[42112] and ">b<" in tags[42112]
    assert pcr.TiledGeoTiffWriter.open(str(tmp_path / "nodir" / "x.tif"), gc, ["a"], o) is not None   # fails at close
    bad = pcr.TiledGeoTiffWriter.open(str(tmp_path / "nodir" / "x.tif"), gc, ["a"], o)
    with pytest.raises(RuntimeError, match="failed to create"):
        bad.close()


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
