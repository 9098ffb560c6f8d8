"""Known-answer vectors lifted from the reference's own gtests, expressed
as pipeline runs so that the same table checks the C oracle (CPU) and the CUDA path (GPU).
Accumulator-level vectors (cell indices + values) are mapped to a 10x1 grid with the
point for cell i at x = i + 0.5.  Citations: /root/reference/tests/cpp/<file>:<line>."""
import numpy as np

SUM, MAX, MIN, AVG, WAVG, COUNT = 0, 1, 2, 3, 4, 5
NAN = float("nan")


def _acc(cells, values, rtype, expect):
    """(grid w,h,tile), cloud list, reduction type, expected first cells of row 0"""
    x = np.array(cells, np.float64) + 0.5
    y = np.full(len(cells), 0.5)
    return dict(w=10, h=1, tile=4096, clouds=[(x, y, {"v": np.array(values, np.float32)})],
                types=[rtype], expect_head=[expect])


ACCUMULATOR = {
    # test_accumulator.cpp:19-51
    "Sum_SingleBatch": _acc([0, 1, 2, 1, 0], [10, 20, 30, 40, 50], SUM, [60, 60, 30, 0, 0, 0, 0, 0, 0, 0]),
    # test_accumulator.cpp:53-80   (untouched cells of a touched tile: Max finalizes -FLT_MAX -> NaN)
    "Max_SingleBatch": _acc([0, 1, 0, 1, 2], [10, 20, 50, 15, 100], MAX, [50, 20, 100] + [NAN] * 7),
    # test_accumulator.cpp:82-108
    "Min_SingleBatch": _acc([0, 1, 0, 1, 2], [10, 20, 5, 15, 100], MIN, [5, 15, 100] + [NAN] * 7),
    # test_accumulator.cpp:110-137
    "Count_SingleBatch": _acc([0, 0, 1, 1, 1, 2], [1, 2, 3, 4, 5, 6], COUNT, [2, 3, 1] + [NAN] * 7),
    # test_accumulator.cpp:139-181  (state 30,2 / 30,1 -> 15, 30)
    "Average_SingleBatch": _acc([0, 1, 0], [10, 30, 20], AVG, [15, 30] + [NAN] * 8),
}


def pipeline_grid_cloud(value_fn, per_cell=1):
    """test_pipeline.cpp fixture: 10x10 grid, 5x5 tiles, points at cell centres,
    x = 0.5 + j, y = 9.5 - i for cell (row i, col j)."""
    xs, ys, vs = [], [], []
    for i in range(10):
        for j in range(10):
            for k in range(per_cell):
                xs.append(0.5 + j); ys.append(9.5 - i); vs.append(value_fn(i * 10 + j, k))
    return np.array(xs), np.array(ys), {"intensity": np.array(vs, np.float32)}


def pipeline_cases():
    cases = {}
    # test_pipeline.cpp:66-120  SingleCloud_Sum: every cell 1.0
    cases["SingleCloud_Sum"] = dict(w=10, h=10, tile=5, clouds=[pipeline_grid_cloud(lambda c, k: 1.0)],
                                    types=[SUM], channel="intensity",
                                    expect=[np.full((10, 10), 1.0, np.float32)])
    # test_pipeline.cpp:122-171  SingleCloud_Average: two points per cell, 10 and 20 -> 15
    cases["SingleCloud_Average"] = dict(w=10, h=10, tile=5,
                                        clouds=[pipeline_grid_cloud(lambda c, k: 10.0 + 10.0 * k, per_cell=2)],
                                        types=[AVG], channel="intensity",
                                        expect=[np.full((10, 10), 15.0, np.float32)])
    # test_pipeline.cpp:173-233  MultipleReductions: intensity = cell index -> Sum i, Max i, Count 1
    idx = np.arange(100, dtype=np.float32).reshape(10, 10)
    cases["MultipleReductions"] = dict(w=10, h=10, tile=5, clouds=[pipeline_grid_cloud(lambda c, k: float(c))],
                                       types=[SUM, MAX, COUNT], channel="intensity",
                                       expect=[idx, idx, np.ones((10, 10), np.float32)])
    # test_pipeline.cpp:235-303  MultipleClouds: two clouds of 50 points over the first 50 cells
    # (rows 0-4), values 10 and 20 -> Sum 30 there; rows 5-9 lie in untouched tiles -> NaN
    def half(v):
        xs, ys, vs = [], [], []
        for i in range(50):
            xs.append(0.5 + (i % 10)); ys.append(9.5 - (i // 10)); vs.append(v)
        return np.array(xs), np.array(ys), {"intensity": np.array(vs, np.float32)}
    exp = np.full((10, 10), np.nan, np.float32)
    exp[:5, :] = 30.0
    cases["MultipleClouds"] = dict(w=10, h=10, tile=5, clouds=[half(10.0), half(20.0)], types=[SUM],
                                   channel="intensity", expect=[exp])
    return cases
