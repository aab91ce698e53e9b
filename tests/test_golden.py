"""Committed golden vectors (tests/golden/oracle_vectors.json, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them (drift guard). GPU: the CUDA path reproduces them through the
C-ABI without calling the oracle at all."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

import workloads as wl

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def f32hex(a):
    return [format(int(v), "08x") for v in np.asarray(a, np.float32).view(np.uint32).ravel()]


def test_oracle_reproduces_golden_vectors():
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    import orc
    orc.build()
    for name, fn in make_golden.cases().items():
        assert fn() == GOLD[name], name


def _path_arrays(pkg, pts):
    return pkg.path_prepare(pts, True, 0.01, 1.0)


def _planner(pkg, kw, pts):
    p = pkg.Planner(pkg.planner_config(**kw))
    pa = _path_arrays(pkg, pts)
    p.set_path(pa["X"], pa["Y"], pa["acc"], pa["total_length"])
    return p, len(pa["X"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,weights", [("dwa_c1_default_weights", (3.0, 3.0, 1.0, 0.0, 0.0)),
                                          ("dwa_c1_all_weights", (1.0, 1.0, 1.0, 1.0, 1.0))])
def test_gpu_dwa_c1_matches_golden(pkg, name, weights):
    g = GOLD[name]
    planner, n = _planner(pkg, wl.cfg_c1(weights=weights), wl.GLOBAL_PATH_XY)
    planner.set_tuning(7, 0)  # the fixture hashes the exact total of every slot (no branch and bound)
    assert n == g["path_points"]
    ranges, angles = wl.scan_360()
    r = planner.cycle_scan((0.0, 0.0, 0.0), (-0.51731912, 0.0, 0.0), ranges, angles, g["seg"][0], g["seg"][1])
    assert (r.slot, r.n_admissible, r.n_points) == (g["slot"], g["n_admissible"], g["P"])
    assert f32hex([r.cost])[0] == g["cost"]
    assert f32hex(r.x) == g["x"] and f32hex(r.y) == g["y"]
    costs, adm = planner.fetch_costs(r.n_slots)
    assert sha(np.where(adm == 1, costs, np.float32(np.finfo(np.float32).max)).astype(np.float32)) == g["costs_sha"]
    planner.close()


@pytest.mark.gpu
def test_gpu_dwa_c2_reduced_matches_golden(pkg):
    g = GOLD["dwa_c2_reduced"]
    planner, n = _planner(pkg, wl.cfg_c2(n_lin=20, n_ang=20), wl.straight_points(20.0))
    cloud = wl.cloud_c2(3, n=4000)
    assert sha(cloud) == g["cloud_sha"]
    start, count = 0, min(max(101, int(math.ceil(2.0 / 0.01)) + 1), n - 1) + 1
    r = planner.cycle_cloud((1.0, 0.0, 0.0), (0.0, 0.0, 0.0), cloud, start, count)
    assert (r.slot, r.n_admissible, r.n_points) == (g["slot"], g["n_admissible"], g["P"])
    assert f32hex([r.cost])[0] == g["cost"]
    planner.close()


@pytest.mark.gpu
def test_gpu_mapper_binning_critical_zone_match_golden(pkg):
    g = GOLD["mapper_c4_sine_scan"]
    angles, ranges = wl.mapping_scan(1080)
    mp = pkg.LocalMapperGPU(400, 400, 0.05, (0.0, 0.0, 0.0), 0.0, False, 1080, 0.01, 2.0, 0.1, 20.0, 256)
    grid = mp.scan_to_grid(angles, ranges)
    assert sha(grid.astype(np.int32)) == g["grid_sha"]
    assert (int((grid == 100).sum()), int((grid == 0).sum()), int((grid == -1).sum())) == \
        (g["occupied"], g["empty"], g["unexplored"])
    gb, pb = mp.scan_to_grid_baysian(angles, ranges)
    assert sha(gb.astype(np.int32)) == g["bayes_grid_sha"] and sha(pb.astype(np.float32)) == g["bayes_prob_sha"]
    mp.set_previous_grid(pb)
    mp.get_previous_grid_in_current_pose((0.35, -0.2), 0.3)
    assert sha(mp.get_previous_grid().astype(np.float32)) == g["warp_sha"]
    mp.close()

    g = GOLD["cloud_binning_c4"]
    pts = wl.cloud_lattice(0)
    data = wl.cloud_bytes_xyz16(pts)
    r = pkg.pointcloud_to_laserscan(data, 16, 16 * len(pts), 1, len(pts), 0, 4, 8, 20.0, 0.1, 2.0, 1080)
    assert sha(np.asarray(r, np.float64)) == g["ranges_sha"] and int((np.asarray(r) < 20.0).sum()) == g["hit_bins"]

    g = GOLD["critical_zone_c4"]
    ang = np.array([2 * math.pi * i / 360 for i in range(360)], np.float64)
    cz = pkg.CriticalZoneCheckerGPU(pkg.SensorInputType.POINTCLOUD, pkg.RobotGeometry.CYLINDER, (0.51, 2.0),
                                    (0.22, 0.0, 0.4), (0.0, 0.0, 0.99, 0.0), 160.0, 0.3, 0.6, ang, 0.1, 2.0, 20.0)
    assert f32hex([cz.check(data, 16, 16 * len(pts), 1, len(pts), 0, 4, 8, True)])[0] == g["cloud_forward"]
    cz.close()
    a2, r2 = wl.dense_slowdown_scan()
    cz = pkg.CriticalZoneCheckerGPU(pkg.SensorInputType.LASERSCAN, pkg.RobotGeometry.CYLINDER, (0.51, 2.0),
                                    (0.22, 0.0, 0.4), (0.0, 0.0, 0.99, 0.0), 160.0, 0.3, 0.6, a2, 0.1, 2.0, 20.0)
    assert f32hex([cz.check(r2, True)])[0] == g["scan_forward"]
    assert f32hex([cz.check(r2, False)])[0] == g["scan_backward"]
    cz.close()
