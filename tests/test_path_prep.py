"""Host path preparation of the DWA controller (Path::interpolate LINEAR + Path::segment,
ref src/datatypes/path.cpp:167-330) against the oracle's restatement. CPU only: no device needed."""
import math

import numpy as np
import pytest

import orc
import workloads as wl


def straight():  # ref: tests/controller_test_helpers.h:35-41
    pts, x = [], 0.0
    while x <= 10.0:
        pts.append((x, 0.0))
        x += 0.5
    return pts


PATHS = {
    "global_path_json": wl.GLOBAL_PATH_XY,
    "two_points": [(0.0, 0.0), (20.0, 0.0)],
    "straight": straight(),
    "uturn": wl.uturn_points(),
    "circle": wl.circle_test_points(),
    "short": [(0.0, 0.0), (0.004, 0.003)],
}


@pytest.mark.parametrize("name", sorted(PATHS))
@pytest.mark.parametrize("interp,seg_len", [(0.01, 1.0), (0.05, 0.7), (0.003, 2.5)])
def test_interpolate_and_segment_match_oracle(pkg, name, interp, seg_len):
    pts = PATHS[name]
    max_pts = int(seg_len / interp + 1)
    ref = orc.Path(pts, interp, seg_len, max_pts)
    got = pkg.path_prepare(pts, True, interp, seg_len, max_pts)
    assert len(got["X"]) == ref.n
    assert np.array_equal(got["X"], ref.X) and np.array_equal(got["Y"], ref.Y)
    assert np.array_equal(got["acc"], ref.acc)
    assert np.array_equal(got["curvature"], ref.curv)
    assert np.array_equal(got["seg_starts"], ref.seg_starts)
    # Path::totalPathLength() is 0 for a path that collapsed to one point (path.cpp:152-154)
    expect = np.float32(ref.total_length) if ref.n >= 2 else np.float32(0.0)
    assert np.float32(got["total_length"]) == expect


def test_non_interpolated_path_keeps_points_and_edge_lengths(pkg):
    pts = PATHS["uturn"]
    got = pkg.path_prepare(pts, False, 0.01, 1.0)
    assert len(got["X"]) == len(pts)
    x = np.asarray(pts, np.float32)
    d = np.sqrt((x[1:, 0] - x[:-1, 0]) ** 2 + (x[1:, 1] - x[:-1, 1]) ** 2, dtype=np.float32)
    # q13: without interpolation the "accumulated" lengths are per-edge lengths, the last entry reads 0
    assert np.allclose(got["acc"][:-1], d, rtol=1e-6) and got["acc"][-1] == 0.0
    assert got["seg_starts"][0] == 0


def test_bad_arguments(pkg):
    with pytest.raises(ValueError):
        pkg.path_prepare([(0.0, 0.0)])
