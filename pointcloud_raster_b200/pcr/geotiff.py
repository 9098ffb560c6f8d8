"""write_geotiff / read_geotiff_info of the reference API (python/bindings.cpp:507-527) over the
GDAL-free writer in csrc/geotiff.cu."""
import ctypes as C

import numpy as np

from .._lib import lib, GridDesc


def _raise():
    raise RuntimeError((lib.pcr_geotiff_last_error() or b"GeoTIFF error").decode("utf-8", "replace"))


def write_geotiff(path, grid, config, options):
    if grid.cols() != config.width or grid.rows() != config.height:
        raise RuntimeError("grid dimensions mismatch config")
    n = grid.num_bands()
    arrays = [np.ascontiguousarray(grid.band_array(i), np.float32) for i in range(n)]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
    names = (C.c_char_p * n)(*[grid.band_desc(i).name.encode() for i in range(n)])
    d = config._desc()
    rc = lib.pcr_geotiff_write(str(path).encode(), ptrs, n, C.byref(d), names, int(config.crs.epsg),
                               str(options.compress).upper().encode(), int(options.compress_level),
                               int(options.tile_width), int(options.tile_height), int(bool(options.bigtiff)),
                               int(bool(options.cloud_optimized)))
    if rc != 0:
        _raise()


def read_geotiff_band(path, band_index, width, height):
    """read_geotiff_band (src/io/grid_io.cpp:445-497): (height, width) float32 array of one band."""
    out = np.empty((int(height), int(width)), np.float32)
    if lib.pcr_geotiff_read_band(str(path).encode(), int(band_index), out.ctypes.data_as(C.POINTER(C.c_float)),
                                 int(width), int(height)) != 0:
        _raise()
    return out


class TiledGeoTiffWriter:
    """TiledGeoTiffWriter (include/pcr/io/grid_io.h:44-70): open, write one reference tile at a time, close."""

    def __init__(self, handle, config, num_bands):
        self._h, self._config, self._nb = handle, config, num_bands

    @staticmethod
    def open(path, config, band_names, options=None):
        from . import GeoTiffOptions
        o = options or GeoTiffOptions()
        n = len(band_names)
        names = (C.c_char_p * max(n, 1))(*[str(b).encode() for b in band_names])
        d = config._desc()
        h = C.c_void_p()
        rc = lib.pcr_geotiff_tiled_open(str(path).encode(), C.byref(d), names, n, int(config.crs.epsg),
                                        str(o.compress).upper().encode(), int(o.compress_level), int(o.tile_width),
                                        int(o.tile_height), int(bool(o.bigtiff)), int(bool(o.cloud_optimized)), C.byref(h))
        if rc != 0 or not h.value:
            return None                      # the reference returns nullptr
        return TiledGeoTiffWriter(h, config, n)

    def write_tile(self, tile, data, num_bands=None):
        """tile: TileIndex; data: band-sequential float32, tile_cols * tile_rows values per band."""
        if not self._h:
            raise RuntimeError("writer not open")
        a = np.ascontiguousarray(data, np.float32)
        nb = self._nb if num_bands is None else int(num_bands)
        _, _, cols, rows = self._config.tile_cell_range(tile)
        if a.size != nb * cols * rows:
            raise RuntimeError("tile data size mismatch")
        if lib.pcr_geotiff_tiled_write_tile(self._h, int(tile.row), int(tile.col),
                                            a.ctypes.data_as(C.POINTER(C.c_float)), nb) != 0:
            _raise()

    def close(self):
        if self._h:
            h, self._h = self._h, None
            if lib.pcr_geotiff_tiled_close(h) != 0:
                _raise()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_geotiff_info(path):
    from . import CRS, BBox
    w, h, nb, epsg = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int32(0)
    b = (C.c_double * 4)()
    if lib.pcr_geotiff_read_info(str(path).encode(), C.byref(w), C.byref(h), C.byref(nb), C.byref(epsg), b) != 0:
        _raise()
    crs = CRS.from_epsg(epsg.value) if epsg.value else CRS()
    bb = BBox()
    bb.min_x, bb.min_y, bb.max_x, bb.max_y = b[0], b[1], b[2], b[3]
    return (w.value, h.value, nb.value, crs, bb)
