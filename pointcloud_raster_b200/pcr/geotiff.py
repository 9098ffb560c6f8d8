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
                               int(options.tile_width), int(options.tile_height), int(bool(options.bigtiff)))
    if rc != 0:
        _raise()


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
