"""Tile-state checkpoints in the reference's .pcrt format.  The fixtures under
tests/golden/pcrt_*.npz hold tile files WRITTEN BY THE UNMODIFIED REFERENCE (oracle/make_golden.py):
  * load them into the B200 pipeline -> finalize must reproduce the reference's band;
  * save from the B200 pipeline      -> same file set, identical headers, identical payload for the
                                        order-free ops (Count/Max/Min), payload within tolerance for sums;
  * save -> load into a fresh pipeline -> keep ingesting == one uninterrupted run."""
import glob
import os
import struct

import numpy as np
import pytest

from util import make_grid, spec, cloud, run_product

pytestmark = pytest.mark.gpu
FIX = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "pcrt_*.npz")))


def load_fix(path):
    z = np.load(path)
    files = {k[5:] + ".pcrt": bytes(z[k]) for k in z.files if k.startswith("file_")}
    return z, files


def parse(b):
    magic, ver, tr, tc, cols, rows, sf, red = struct.unpack_from("<IIiiiiiB", b, 0)
    return dict(magic=magic, ver=ver, tile=(tr, tc), cols=cols, rows=rows, sf=sf, red=red), \
        np.frombuffer(b, "<f4", offset=36).reshape(sf, rows, cols)


def pipeline(pcr, rtype, **kw):
    gc = make_grid(pcr, 40, 24, tile=16)
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.exec_mode = pcr.ExecutionMode.GPU
    cfg.reductions = [spec(pcr, "v", pcr.ReductionType(int(rtype)))]
    for k, v in kw.items():
        setattr(cfg, k, v)
    p = pcr.Pipeline.create(cfg)
    assert p is not None
    return p


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[5:-4] for p in FIX])
def test_load_reference_written_tiles(gpu_pcr, tmp_path, path):
    z, files = load_fix(path)
    for name, b in files.items():
        (tmp_path / name).write_bytes(b)
    p = pipeline(gpu_pcr, z["rtype"][0])
    p.load_state(tmp_path)
    p.finalize()
    assert np.array_equal(np.array(p.result().band_array(0)), z["band"], equal_nan=True)
    # resume flag: same thing through PipelineConfig
    p2 = pipeline(gpu_pcr, z["rtype"][0], state_dir=str(tmp_path), resume=True)
    p2.finalize()
    assert np.array_equal(np.array(p2.result().band_array(0)), z["band"], equal_nan=True)


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[5:-4] for p in FIX])
def test_save_matches_reference_files(gpu_pcr, tmp_path, path):
    z, files = load_fix(path)
    p = pipeline(gpu_pcr, z["rtype"][0])
    p.ingest(cloud(gpu_pcr, z["x"], z["y"], {"v": z["v"]}))
    p.save_state(tmp_path)
    mine = {os.path.basename(f): open(f, "rb").read() for f in glob.glob(str(tmp_path / "*.pcrt"))}
    assert sorted(mine) == sorted(files)                     # one file per TOUCHED tile, same names
    for name in files:
        hr, dr = parse(files[name]); hm, dm = parse(mine[name])
        assert hr == hm, name
        if int(z["rtype"][0]) in (1, 2, 5):                  # Max, Min, Count: payload bit-identical
            assert mine[name] == files[name], name
        else:
            assert np.allclose(dm, dr, rtol=1e-5, atol=1e-5), name


def test_checkpoint_resume_equals_uninterrupted_run(gpu_pcr, tmp_path):
    pcr = gpu_pcr
    gc = make_grid(pcr, 100, 70, tile=32)
    rng = np.random.default_rng(3)
    R = pcr.ReductionType
    specs = [spec(pcr, "v", t) for t in (R.Count, R.Max, R.Min, R.Average, R.Sum)]
    c1 = (rng.uniform(0, 60, 20000), rng.uniform(0, 70, 20000), {"v": rng.normal(0, 2, 20000).astype(np.float32)})
    c2 = (rng.uniform(30, 100, 20000), rng.uniform(0, 40, 20000), {"v": rng.normal(3, 2, 20000).astype(np.float32)})
    ref, _ = run_product(pcr, gc, [c1, c2], specs, deterministic=True)
    cfg = pcr.PipelineConfig(); cfg.grid = gc; cfg.reductions = specs; cfg.exec_mode = R and pcr.ExecutionMode.GPU
    cfg.deterministic = True
    a = pcr.Pipeline.create(cfg)
    a.ingest(cloud(pcr, *c1))
    a.save_state(tmp_path)
    assert sorted(os.listdir(tmp_path)) == [f"band_{i}" for i in range(5)]     # multi-reduction layout
    del a
    cfg.state_dir, cfg.resume = str(tmp_path), True
    b = pcr.Pipeline.create(cfg)
    b.ingest(cloud(pcr, *c2))
    b.finalize()
    for i in range(5):
        assert np.array_equal(np.array(b.result().band_array(i)), ref[i], equal_nan=True), i


def test_mismatched_or_missing_files_are_ignored(gpu_pcr, tmp_path):
    z, files = load_fix(FIX[0])
    name, b = next(iter(files.items()))
    (tmp_path / name).write_bytes(b)
    gc = make_grid(gpu_pcr, 40, 24, tile=8)                   # other tile size: header dims do not match
    cfg = gpu_pcr.PipelineConfig(); cfg.grid = gc; cfg.exec_mode = gpu_pcr.ExecutionMode.GPU
    cfg.reductions = [spec(gpu_pcr, "v", gpu_pcr.ReductionType(int(z["rtype"][0])))]
    p = gpu_pcr.Pipeline.create(cfg)
    p.load_state(tmp_path)
    p.finalize()
    assert np.isnan(np.array(p.result().band_array(0))).all()
    (tmp_path / name).write_bytes(b[:100])                    # truncated payload with a matching header
    p2 = pipeline(gpu_pcr, z["rtype"][0])
    with pytest.raises(RuntimeError, match="truncated"):
        p2.load_state(tmp_path)
