import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # The product library and the C oracle are build artefacts (git-ignored): build them on a fresh
    # checkout so that the suite does not depend on someone having run __graft_entry__.build() first.
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "pointcloud_raster_b200", "libpcr_b200.so")):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "pointcloud_raster_b200", "csrc")])
    if not os.path.exists(os.path.join(ROOT, "oracle", "libpcr_oracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/pcr_oracle.c), built on demand with gcc."""
    import oracle as orc
    return orc.Oracle()


@pytest.fixture(scope="session")
def pcr():
    from pointcloud_raster_b200 import pcr as _pcr
    return _pcr


@pytest.fixture(scope="session")
def gpu_pcr(pcr):
    """The product API, with a loud failure (not a skip) when the box has no usable GPU:
    -m gpu tests must never pass on a fallback."""
    assert pcr.device_count() > 0, "GPU test selected but no CUDA device is visible"
    return pcr


@pytest.fixture(autouse=True)
def _mem_trace(request):
    """PCR_MEM_TRACE=<file>: free HBM before every GPU test (finds pipelines that outlive their test)."""
    path = os.environ.get("PCR_MEM_TRACE")
    if path and request.node.get_closest_marker("gpu"):
        try:
            from pointcloud_raster_b200 import pcr as _pcr
            free, total = _pcr.device_mem_info()
            with open(path, "a") as f:
                f.write(f"{free / 2**30:8.1f} GB free  {request.node.nodeid}\n")
        except Exception:
            pass
    yield
