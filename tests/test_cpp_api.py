"""The C++ API header (include/pcr_b200.hpp) — builds tests/cpp/test_cpp_api.cpp with
plain g++ against libpcr_b200.so and runs it.  The cases mirror the reference's own
tests/cpp/test_pipeline.cpp / test_gpu_pipeline.cpp; see that file's header."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pointcloud_raster_b200")
SRC = os.path.join(ROOT, "tests", "cpp", "test_cpp_api.cpp")


def _gxx():
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    pytest.skip("no g++")


def _build(tmp_path):
    exe = str(tmp_path / "test_cpp_api")
    cmd = [_gxx(), "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
           "-o", exe, "-L", PKG, "-lpcr_b200", f"-Wl,-rpath,{PKG}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_compiles_and_links(pcr, tmp_path):
    exe = _build(tmp_path)
    names = subprocess.run([exe, "--list"], capture_output=True, text=True, check=True).stdout.split()
    assert "single_cloud_sum" in names and "device_and_pinned_clouds" in names and len(names) >= 18


def test_header_is_self_contained_in_two_translation_units(pcr, tmp_path):
    # header-only: every function must be inline, or two TUs collide at link time
    a, b = tmp_path / "a.cpp", tmp_path / "b.cpp"
    a.write_text('#include "pcr_b200.hpp"\nint other();\nint main() { return other() + pcr::cuda_device_count() * 0; }\n')
    b.write_text('#include "pcr_b200.hpp"\nint other() { pcr::GridConfig g; g.compute_dimensions(); return g.width; }\n')
    r = subprocess.run([_gxx(), "-std=c++17", "-I", os.path.join(ROOT, "include"), str(a), str(b), "-o",
                        str(tmp_path / "two"), "-L", PKG, "-lpcr_b200", f"-Wl,-rpath,{PKG}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_cpp_api_cases(gpu_pcr, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed expectations" in r.stdout
