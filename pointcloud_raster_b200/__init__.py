"""pointcloud_raster_b200 — B200-native (sm_100a) ingest/finalize path of
BigHippo123/pointcloud-raster, behind the reference's own Python API.

    from pointcloud_raster_b200 import pcr        # drop-in for `import pcr`

The package holds only the hot path: ``csrc/`` (hand-written CUDA kernels, the
C++ engine and the C-ABI of ``include/pcr_b200.h``), the built ``libpcr_b200.so``
and ``pcr/`` (the ctypes mirror of the reference's pybind11 module).  There is no
CPU fallback: importing works anywhere, but every compute call needs the built
library and a B200.
"""
from . import _lib  # noqa: F401  (fails loudly if libpcr_b200.so is missing)

__all__ = ["pcr"]
__version__ = "0.1.0"


def __getattr__(name):
    if name == "pcr":
        import importlib
        return importlib.import_module(".pcr", __name__)
    raise AttributeError(name)
