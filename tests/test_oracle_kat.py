"""Pins the CPU oracle against every known-answer test the reference holds for the hot path.

Each test restates one case of the reference's own Boost tests (same inputs, same expected value,
same tolerance):
  - src/kompass_cpp/tests/cost_evaluator_test.cpp:217-461 (12 cost cases, tol 1e-4 / 0.02)
  - src/kompass_cpp/tests/critical_zone_test.cpp:39-333   (14 emergency-stop cases)
  - src/kompass_cpp/tests/collisions_test.cpp:11-77       (3 FCL booleans)
"""
import ctypes as C
import math
import struct

import numpy as np
import pytest

import orc


# ------------------------------------------------------------------ helpers (cost_evaluator_test.cpp:29-142)
def straight_path(length, interp, seg):
    return orc.Path([(0.0, 0.0), (length, 0.0)], interp, seg)


def circle34_path(R, input_pts, interp, seg):
    max_theta = 3.0 * math.pi / 2.0
    pts = []
    for i in range(input_pts):
        th = (i / (input_pts - 1)) * max_theta
        pts.append((R * math.cos(th), R * math.sin(th)))
    return orc.Path(pts, interp, seg)


def sample_at_endpoint(n, pt):
    return dict(x=np.full((1, n), pt[0], np.float32), y=np.full((1, n), pt[1], np.float32),
                vx=np.zeros((1, n - 1), np.float32), vy=np.zeros((1, n - 1), np.float32),
                omega=np.zeros((1, n - 1), np.float32))


def sample_with_path(pts, vels=None):
    pts = np.asarray(pts, np.float32)
    n = len(pts)
    s = dict(x=pts[None, :, 0].copy(), y=pts[None, :, 1].copy(),
             vx=np.zeros((1, n - 1), np.float32), vy=np.zeros((1, n - 1), np.float32),
             omega=np.zeros((1, n - 1), np.float32))
    if vels is not None:
        v = np.asarray(vels, np.float64)
        s["vx"][0], s["vy"][0], s["omega"][0] = v[:, 0], v[:, 1], v[:, 2]
    return s


def solo(name, value=1.0):
    w = dict(w_path=0.0, w_goal=0.0, w_obstacles=0.0, w_smooth=0.0, w_jerk=0.0)
    w[name] = value
    return orc.cost_cfg(acc_limits=(1.0, 1.0, 1.0), **w)


def eval_cost(ccfg, ref, seg_idx, samples, obstacles=None):
    obs, D = None, 0.0
    if obstacles is not None:
        cloud = np.asarray(obstacles, np.float32).reshape(-1, 3)
        obs = orc.cost_points(ccfg, (0.0, 0.0, 0.0), cloud=cloud)
        D = float(np.float32(30.0) / np.float32(3.0))
    found, idx, cost, _ = orc.cost_evaluate(ccfg, samples, ref, ref.segment(seg_idx), obs, D)
    assert found
    return cost


def close(a, b, tol):
    # boost tt::tolerance(tol): relative difference w.r.t. both operands
    if a == b:
        return True
    d = abs(a - b)
    return d <= tol * abs(a) and d <= tol * abs(b) if (a != 0 and b != 0) else d <= tol


# ------------------------------------------------------------------ cost KATs
def test_goal_cost_on_straight_path():
    ref = straight_path(10.0, 1.0, 5.0)
    assert ref.n == 11
    c = eval_cost(solo("w_goal"), ref, 0, sample_at_endpoint(5, (4.0, 0.0)))
    assert close(c, 0.6, 1e-4)


def test_goal_cost_arc_remaining_on_curved_path():
    R = 2.0
    ref = circle34_path(R, 60, 0.05, 20.0)
    total = ref.total_length
    follow_pt = (R * math.cos(0.5), R * math.sin(0.5))
    follow = eval_cost(solo("w_goal"), ref, 0, sample_at_endpoint(5, follow_pt))
    chord = eval_cost(solo("w_goal"), ref, 0, sample_at_endpoint(5, (1.5, -0.5)))
    assert close(follow, (total - R * 0.5) / total, 0.02)
    assert close(chord, 1.0 + math.sqrt(0.5) / total, 0.02)
    assert follow < chord


def test_goal_cost_tie_breaker():
    ref = straight_path(10.0, 1.0, 5.0)
    a = eval_cost(solo("w_goal"), ref, 0, sample_at_endpoint(5, (4.0, 0.1)))
    b = eval_cost(solo("w_goal"), ref, 0, sample_at_endpoint(5, (4.0, 0.5)))
    assert close(a, 0.61, 1e-4) and close(b, 0.65, 1e-4) and a < b


def test_path_cost_centered_sample():
    ref = straight_path(10.0, 1.0, 5.0)
    pts = [(float(i), 0.0) for i in range(5)]
    c = eval_cost(solo("w_path"), ref, 0, sample_with_path(pts))
    assert abs(c) <= 1e-4


def test_path_cost_constant_lateral_offset():
    ref = straight_path(10.0, 1.0, 5.0)
    d, seg_len = 0.5, 4.0
    pts = [(float(i), d) for i in range(5)]
    c = eval_cost(solo("w_path"), ref, 0, sample_with_path(pts))
    assert close(c, (d + d / seg_len) / 2.0, 1e-4)


def test_smoothness_cost_constant_velocity():
    ref = straight_path(10.0, 1.0, 5.0)
    c = eval_cost(solo("w_smooth"), ref, 0, sample_with_path([(0, 0)] * 5, [(1.0, 0, 0)] * 4))
    assert abs(c) <= 1e-4


def test_smoothness_cost_single_step_change():
    ref = straight_path(10.0, 1.0, 5.0)
    vels = [(0.0, 0, 0), (1.0, 0, 0), (1.0, 0, 0), (1.0, 0, 0)]
    c = eval_cost(solo("w_smooth"), ref, 0, sample_with_path([(0, 0)] * 5, vels))
    assert close(c, 1.0 / 12.0, 1e-4)


def test_jerk_cost_constant_acceleration():
    ref = straight_path(10.0, 1.0, 5.0)
    vels = [(0.1, 0, 0), (0.2, 0, 0), (0.3, 0, 0), (0.4, 0, 0)]
    c = eval_cost(solo("w_jerk"), ref, 0, sample_with_path([(0, 0)] * 5, vels))
    assert abs(c) <= 1e-4


def test_jerk_cost_known_second_diff():
    ref = straight_path(10.0, 1.0, 5.0)
    vels = [(0.0, 0, 0), (1.0, 0, 0), (3.0, 0, 0), (6.0, 0, 0)]
    c = eval_cost(solo("w_jerk"), ref, 0, sample_with_path([(0, 0)] * 5, vels))
    assert close(c, 2.0 / 12.0, 1e-4)


@pytest.mark.parametrize("obst,expected", [((20.0, 0.0, 0.0), 0.0), ((0.0, 0.0, 0.0), 1.0),
                                           ((5.0, 0.0, 0.0), 0.5)])
def test_obstacles_cost(obst, expected):
    ref = straight_path(10.0, 1.0, 5.0)
    c = eval_cost(solo("w_obstacles"), ref, 0, sample_at_endpoint(5, (0.0, 0.0)), [obst])
    assert abs(c - expected) <= 1e-4


# ------------------------------------------------------------------ critical zone KATs (test.h:55-116)
def init_laserscan(n, r):
    angles = np.array([2.0 * math.pi * i / n for i in range(n)], np.float64)
    return np.full(n, r, np.float64), angles


def set_at_angle(angle, value, ranges, angles):
    a = math.fmod(angle, 2 * math.pi)
    if a < 0:
        a += 2 * math.pi
    ranges[int(np.argmin(np.abs(angles - a)))] = value


def test_critical_zone_laserscan():
    cfg = orc.cz_cfg()
    ranges, angles = init_laserscan(360, 10.0)
    chk = lambda fwd: orc.cz_check_scan(cfg, angles, ranges, fwd)
    # 1: behind & forward
    for a in (0.0, 0.1, -0.1):
        set_at_angle(a, 0.2, ranges, angles)
    assert chk(True) == 1.0
    # 2: far & forward
    ranges, angles = init_laserscan(360, 10.0)
    assert chk(True) == 1.0
    # 3: front close & forward
    for a in (math.pi, math.pi + 0.1, math.pi - 0.1):
        set_at_angle(a, 0.2, ranges, angles)
    assert chk(True) == 0.0
    # 4: front close & backward
    assert chk(False) == 1.0
    # 5: back close & backward
    for a in (0.0, 0.1, -0.1):
        set_at_angle(a, 0.2, ranges, angles)
    assert chk(False) == 0.0
    # 6: back slowdown & backward
    ranges, angles = init_laserscan(360, 10.0)
    set_at_angle(0.0, 1.3, ranges, angles)
    r = chk(False)
    assert 0.0 < r < 1.0
    # 7: back slowdown & forward
    assert chk(True) == 1.0
    # 8: front slowdown & forward
    set_at_angle(math.pi, 0.7, ranges, angles)
    r = chk(True)
    assert 0.0 < r < 1.0


def cloud_bytes(pts):
    b = b"".join(struct.pack("<ffff", x, y, z, 0.0) for (x, y, z) in pts)
    return np.frombuffer(b, dtype=np.int8) if b else np.zeros(0, np.int8)


def test_critical_zone_pointcloud():
    cfg = orc.cz_cfg(sensor_position=(0.0, 0.0, 0.0), sensor_rotation=(0.0, 0.0, 0.0, 1.0))
    _, angles = init_laserscan(360, 10.0)

    def run(pts, fwd):
        d = cloud_bytes(pts)
        n = len(pts)
        return orc.cz_check_cloud(cfg, angles, d, 16, n * 16, 1, n, 0, 4, 8, fwd)

    assert run([], True) == 1.0                                    # 9
    assert run([(0.7, 0.0, 0.5)], True) == 0.0                     # 10
    assert run([(0.7, 0.0, 3.0)], True) == 1.0                     # 11
    assert 0.4 < run([(0.95, 0.0, 0.5)], True) < 0.6               # 12
    pts13 = [(0.95, 0, 0.5), (1, 1, 0.5), (-1, -1, 0.5), (-0.1, -0.1, 3.0), (-0.1, -0.1, -3.0),
             (0.1, 0.2, 4.0), (0.1, 0.2, -4.0), (0.75, 0.0, 0.5)]
    assert run(pts13, True) == 0.0                                 # 13
    pts14 = [(0.95, 0, 0.5), (-0.95, 0, 0.5), (1, 1, 0.5), (-1, -1, 0.5), (-0.1, -0.1, 3.0),
             (-0.1, -0.1, -3.0), (0.1, 0.2, 4.0), (0.1, 0.2, -4.0)]
    assert 0.4 < run(pts14, False) < 0.6                           # 14


# ------------------------------------------------------------------ FCL booleans (collisions_test.cpp)
def test_collision_booleans():
    # Eigen::Quaternionf{0,0,0,1} is the (w,x,y,z) ctor: w=0, z=1 -> coeffs (x,y,z,w) = (0,0,1,0)
    cfg = orc.sampler_cfg(shape=orc.BOX, dims=(0.4, 0.4, 1.0), sensor_position=(0.0, 0.0, 1.0),
                          sensor_rotation=(0.0, 0.0, 1.0, 0.0), octree_resolution=0.1)
    angles = [0.0, 0.1, 0.2]
    assert orc.check_collision(cfg, (0, 0, 0), (0, 0, 0), scan=([1.0, 1.0, 1.0], angles)) == 0
    assert orc.check_collision(cfg, (3, 5, 0), (3, 5, 0), scan=([0.25, 0.5, 0.5], angles)) == 1
    assert orc.check_collision(cfg, (3, 5, 0), (3, 5, 0), cloud=[(3.1, 5.1, -0.5)]) == 1


def test_batched_collision_states_agree_with_single_checks():
    """orc_check_collision_states (row f4 oracle) == orc_check_collision per state, for the three
    sensor-frame modes of CollisionChecker::updateSensorData"""
    rng = np.random.default_rng(5)
    cfg = orc.sampler_cfg(shape=orc.BOX, dims=(0.5, 0.3, 0.6), sensor_position=(0.1, 0.0, 0.2),
                          sensor_rotation=(0.0, 0.0, 0.2, 0.98), octree_resolution=0.08)
    body = (0.4, -0.2, 0.5)
    states = np.column_stack([rng.uniform(-3, 3, 300), rng.uniform(-3, 3, 300), rng.uniform(-3, 3, 300)])
    ang = np.linspace(0, 2 * np.pi, 90, endpoint=False)
    scan = (rng.uniform(0.5, 3.0, 90), ang)
    any_s, per_s = orc.check_collision_states(cfg, body, states, scan=scan)
    assert [orc.check_collision(cfg, body, s, scan=scan) for s in states] == per_s.tolist()
    cloud = np.column_stack([rng.uniform(-3, 3, 400), rng.uniform(-3, 3, 400), rng.uniform(-0.2, 0.4, 400)])
    any_c, per_c = orc.check_collision_states(cfg, body, states, cloud=cloud, global_frame=True)
    assert [orc.check_collision(cfg, body, s, cloud=cloud) for s in states] == per_c.tolist()
    assert any_s == bool(per_s.any()) and any_c == bool(per_c.any()) and 0 < per_c.sum() < 300
    # a body-frame cloud moves with the body pose: shifting body and states together keeps the answers
    _, local0 = orc.check_collision_states(cfg, (0, 0, 0), states, cloud=cloud, global_frame=False)
    shifted = states + np.array([2.0, -1.0, 0.0])
    _, local1 = orc.check_collision_states(cfg, (2.0, -1.0, 0.0), shifted, cloud=cloud, global_frame=False)
    assert (local0 != local1).mean() < 0.02  # only float-rounding of the shifted poses may flip a tangent case


def test_upside_down_mount_mirrors_the_octree():
    """sensor_tf_world_ = body * sensor with a 180 deg flip about x: the octree's cubes are carried to
    (x, -y, -z) + t (collision_check.cpp:118-123 hands the full transform to FCL)"""
    cfg = orc.sampler_cfg(shape=orc.CYLINDER, dims=(0.2, 1.0, 0.0), sensor_position=(0.0, 0.0, 0.1),
                          sensor_rotation=(1.0, 0.0, 0.0, 0.0), octree_resolution=0.05)
    ang, r = np.arctan2(0.6, 1.0), np.hypot(1.0, 0.6)
    assert orc.check_collision(cfg, (0, 0, 0), (1.0, -0.6, 0.0), scan=([r], [ang])) == 1
    assert orc.check_collision(cfg, (0, 0, 0), (1.0, 0.6, 0.0), scan=([r], [ang])) == 0
    # a sensor_rotation that is not a rotation (|q| = 0.996) stays unsupported (negative return code)
    tilted = orc.sampler_cfg(sensor_rotation=(0.3, 0.0, 0.0, 0.95))
    assert orc.lib().orc_check_collision(C.byref(tilted), orc.dp(orc.f64((0, 0, 0))), orc.dp(orc.f64((0, 0, 0))),
                                         0, orc.dp(orc.f64([1.0])), orc.dp(orc.f64([0.0])), 1) < 0


# ------------------------------------------------------------------ sizes (trajectory.h:19-51)
def test_sizes():
    L = orc.lib()
    assert L.orc_num_trajectories(orc.DIFFERENTIAL_DRIVE, 20, 20) == 441
    assert L.orc_num_trajectories(orc.DIFFERENTIAL_DRIVE, 100, 100) == 10201
    assert L.orc_num_trajectories(orc.ACKERMANN, 224, 224) == 50625
    assert L.orc_num_trajectories(orc.OMNI, 224, 224) == 169 * 57 + 169 * 225
    assert L.orc_num_points(0.1, 1.0) == 10
    assert L.orc_num_points(0.02, 1.0) == 50


# ------------------------------------------------------------------ general (tilted) voxel test
_GENERAL_SCRIPT = '''
import sys, math, numpy as np
sys.path.insert(0, sys.argv[2])
import orc, workloads as wl
rng = np.random.default_rng(5)
out = []
for shape, dims in ((0, (0.25, 0.6, 0)), (1, (0.5, 0.3, 0.4)), (2, (0.3, 0, 0))):
    for rot in ((0, 0, 0, 1), (0, 0, math.sin(0.3), math.cos(0.3)), (1, 0, 0, 0)):
        cfg = orc.sampler_cfg(shape=shape, dims=dims, sensor_position=(0.1, -0.05, 0.2), sensor_rotation=rot,
                              octree_resolution=0.1)
        ranges, angles = wl.scan_360(3, n=720, lo=0.3, hi=3.0)
        st = np.stack([rng.uniform(-3, 3, 1500), rng.uniform(-3, 3, 1500), rng.uniform(-3.2, 3.2, 1500)], 1)
        out.append(orc.check_collision_states(cfg, (0.2, 0.1, 0.4), st, scan=(ranges, angles))[1])
np.save(sys.argv[1], np.concatenate(out))
'''


def test_general_voxel_test_equals_the_planar_one(tmp_path):
    """oracle/voxel_model.h carries two exact evaluations: the planar one (sensor z axis vertical: z test
    folded into the insertion, disc / rectangle against a square) and the general one for tilted sensors
    (oriented cubes: sphere / box / cylinder vs OBB). On planar frames they must agree; the environment
    switch ORC_FORCE_GENERAL_VOXEL routes planar frames through the general code."""
    import os
    import subprocess
    import sys
    script = tmp_path / "gen.py"
    script.write_text(_GENERAL_SCRIPT)
    here = os.path.dirname(os.path.abspath(__file__))
    env = {k: v for k, v in os.environ.items() if k != "ORC_FORCE_GENERAL_VOXEL"}
    subprocess.check_call([sys.executable, str(script), str(tmp_path / "a.npy"), here], env=env)
    subprocess.check_call([sys.executable, str(script), str(tmp_path / "b.npy"), here],
                          env=dict(env, ORC_FORCE_GENERAL_VOXEL="1"))
    a, b = np.load(tmp_path / "a.npy"), np.load(tmp_path / "b.npy")
    assert 0.2 < a.mean() < 0.9 and np.array_equal(a, b), (a != b).sum()


def test_tilted_mount_known_answer():
    """Sensor rolled 90 deg about x: a sensor-frame point (x, y, z) sits at body (x, -z, y). A laser
    return at range 1 m, bearing 90 deg (sensor +y, z = -sensor_z/2 = -0.2) therefore becomes an obstacle
    ABOVE the sensor at body (0, 0.2, 0.4 + 1.0): far above a 0.5 m tall cylinder, no collision anywhere
    near; bearing 0 (sensor +x) stays at body (1, 0.2, 0.4): blocks a robot standing at x = 1."""
    s = math.sin(math.pi / 4)
    cfg = orc.sampler_cfg(shape=orc.CYLINDER, dims=(0.2, 1.0, 0.0), sensor_position=(0.0, 0.0, 0.4),
                          sensor_rotation=(s, 0.0, 0.0, s), octree_resolution=0.05)
    up = ([1.0], [math.pi / 2])
    for x, y in ((0.0, 0.0), (0.0, 0.2), (0.0, 1.0), (0.0, -1.0)):
        assert orc.check_collision(cfg, (0, 0, 0), (x, y, 0.0), scan=up) == 0
    ahead = ([1.0], [0.0])
    assert orc.check_collision(cfg, (0, 0, 0), (1.0, 0.2, 0.0), scan=ahead) == 1
    assert orc.check_collision(cfg, (0, 0, 0), (1.0, 0.6, 0.0), scan=ahead) == 0
    assert orc.check_collision(cfg, (0, 0, 0), (0.3, 0.2, 0.0), scan=ahead) == 0
