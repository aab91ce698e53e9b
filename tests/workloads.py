"""Synthetic workloads C1..C5 of SURVEY.md §8(d), shared by the parity tests and bench.py.

Every generator is seeded (numpy default_rng(20261018 + offset)) so the oracle and the CUDA path
see identical inputs on any box. The reference path of C1 is the 5-pose path of the reference's own
fixture tests/resources/control/global_path.json (positions only).
"""
import math
import struct

import numpy as np

SEED = 20261018

# x, y of /root/reference/tests/resources/control/global_path.json (5 poses)
GLOBAL_PATH_XY = [
    (-0.51731912, 0.0),
    (-0.4864472592186867, 0.5561681993192574),
    (-0.39707157220754075, 0.7827921730680379),
    (0.4401074774387863, 2.375158424203812),
    (1.641465663909912, 3.210279703140259),
]


def circle34_points(R=10.0, n=200):
    """3/4 CCW circle, ref: src/kompass_cpp/tests/controller_test_helpers.h:63-72 shape."""
    mt = 3.0 * math.pi / 2.0
    return [(R * math.cos(i / (n - 1) * mt), R * math.sin(i / (n - 1) * mt)) for i in range(n)]


def straight_test_points():
    """ref: src/kompass_cpp/tests/controller_test_helpers.h:35-41 createStraightPath"""
    pts, x = [], 0.0
    while x <= 10.0:
        pts.append((x, 0.0))
        x += 0.5
    return pts


def uturn_points():
    """ref: controller_test_helpers.h:43-61 createUTurnPath"""
    pts, x = [], 0.0
    while x <= 5.0:
        pts.append((x, 0.0))
        x += 0.5
    a = -math.pi / 2
    while a <= math.pi / 2:
        pts.append((5.0 + 5.5 * math.cos(a), 2.5 + 5.5 * math.sin(a)))
        a += 0.2
    x = 5.0
    while x >= 0.0:
        pts.append((x, 5.0))
        x -= 0.5
    return pts


def circle_test_points():
    """ref: controller_test_helpers.h:63-72 createCirclePath"""
    pts, a = [], 0.0
    while a <= 3.0 * math.pi / 2.0:
        pts.append((10.0 * math.cos(a), 10.0 * math.sin(a)))
        a += 0.1
    return pts


def round_obstacle(x, y, radius, resolution=0.1):
    """ref: controller_test_helpers.h:75-92 createRoundObstacle"""
    cloud, r = [], 0.0
    while r <= radius:
        theta = 0.0
        while theta < 2 * math.pi:
            cloud.append((x + r * math.cos(theta), y + r * math.sin(theta), 0.0))
            theta = theta + (resolution / r if r != 0 else math.inf)
        if r == 0:
            cloud.append((x, y, 0.0))
        r += resolution
    return np.asarray(cloud, np.float32)


def straight_points(length=20.0):
    return [(0.0, 0.0), (length, 0.0)]


def scan_360(seed_off=0, n=360, lo=0.5, hi=10.0):
    rng = np.random.default_rng(SEED + seed_off)
    angles = np.array([2.0 * math.pi * i / n for i in range(n)], np.float64)
    ranges = rng.uniform(lo, hi, n)
    return ranges, angles


def cloud_c2(seed_off=0, n=100_000, center=(0.0, 0.0), r_min=1.5, intruder_r=(0.3, 1.5)):
    """98% ring points r~U[1.5,10], 2% intruders r~U[0.3,1.5] in a 30 deg wedge at bearing 40 deg,
    z~U[0,0.3]. r_min > reach + robot radius keeps every non-intruder sample
    admissible (the heavy case for cost evaluation; used by bench.py)."""
    rng = np.random.default_rng(SEED + seed_off)
    n_in = int(n * 0.02)
    n_out = n - n_in
    r = np.concatenate([rng.uniform(r_min, 10.0, n_out), rng.uniform(intruder_r[0], intruder_r[1], n_in)])
    a = np.concatenate([rng.uniform(0.0, 2 * math.pi, n_out),
                        math.radians(40.0) + rng.uniform(-math.radians(15), math.radians(15), n_in)])
    z = rng.uniform(0.0, 0.3, n)
    pts = np.stack([center[0] + r * np.cos(a), center[1] + r * np.sin(a), z], axis=1).astype(np.float32)
    perm = rng.permutation(n)
    return np.ascontiguousarray(pts[perm])


def cloud_lattice(seed_off=0, n=100_000):
    """ref: benchmarks/benchmark_runner.cpp:93-109 generate_heavy_pointcloud_bytes (seeded)."""
    rng = np.random.default_rng(SEED + 1000 + seed_off)
    x = rng.integers(0, 2000, n) / np.float32(100.0) - np.float32(10.0)
    y = rng.integers(0, 2000, n) / np.float32(100.0) - np.float32(10.0)
    z = rng.integers(0, 300, n) / np.float32(100.0)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0], pts[:, 1], pts[:, 2] = x, y, z
    return pts


def cloud_bytes_xyz16(pts4):
    return np.frombuffer(np.ascontiguousarray(pts4, dtype=np.float32).tobytes(), dtype=np.int8)


def mapping_scan(n=1080):
    """ref: benchmarks/benchmark_runner.cpp:112-121 generate_mapping_scan"""
    step = (2.0 * math.pi) / n
    angles = np.array([-math.pi + i * step for i in range(n)], np.float64)
    ranges = 5.0 + 2.0 * np.sin(angles * 20.0)
    return angles, ranges


def dense_slowdown_scan(n=3600, sensor_x=0.22, target=0.96):
    """ref: benchmarks/benchmark_runner.cpp:317-352 (every ray lands in the slowdown band)"""
    step = (2.0 * math.pi) / n
    angles = np.array([-math.pi + i * step for i in range(n)], np.float64)
    b = 2.0 * sensor_x * np.cos(angles)
    c = sensor_x * sensor_x - target * target
    disc = b * b - 4.0 * c
    ranges = np.where(disc >= 0, (-b + np.sqrt(np.maximum(disc, 0))) / 2.0, 10.0)
    return angles, ranges


# ------------------------------------------------------------------------------------------------
# planner configurations (keyword dicts usable for both orc.sampler_cfg/cost_cfg and planner_config)
# ------------------------------------------------------------------------------------------------
def cfg_c1(weights=(3.0, 3.0, 1.0, 0.0, 0.0)):
    return dict(control_type=1, time_step=0.1, prediction_horizon=1.0, control_horizon=0.2,
                max_linear_samples=20, max_angular_samples=20, vx=(1.0, 5.0, 10.0), vy=(0.0, 0.0, 0.0),
                omega=(4.0, 3.0, 3.0), shape=0, dims=(0.1, 0.4, 0.0), sensor_position=(0, 0, 0),
                sensor_rotation=(0, 0, 0, 1), octree_resolution=0.1, drop_samples=True,
                weights=weights, max_local_range=10.0)


def cfg_c2(weights=(1.0, 1.0, 1.0, 1.0, 1.0), n_lin=100, n_ang=100, dims=(0.2, 0.4, 0.0)):
    return dict(control_type=1, time_step=0.02, prediction_horizon=1.0, control_horizon=0.1,
                max_linear_samples=n_lin, max_angular_samples=n_ang, vx=(2.0, 50.0, 50.0),
                vy=(0.0, 0.0, 0.0), omega=(4.0, 100.0, 100.0), shape=0, dims=dims,
                sensor_position=(0, 0, 0), sensor_rotation=(0, 0, 0, 1), octree_resolution=0.1,
                drop_samples=True, weights=weights, max_local_range=10.0)


def cfg_c3(control_type=0, n=224, drop_samples=True, shape=1, dims=(0.5, 0.3, 0.4)):
    return dict(control_type=control_type, time_step=0.02, prediction_horizon=2.0, control_horizon=0.2,
                max_linear_samples=n, max_angular_samples=n, vx=(2.0, 50.0, 50.0),
                vy=(2.0, 50.0, 50.0), omega=(4.0, 100.0, 100.0), shape=shape, dims=dims,
                sensor_position=(0, 0, 0), sensor_rotation=(0, 0, 0, 1), octree_resolution=0.1,
                drop_samples=drop_samples, weights=(1.0, 1.0, 1.0, 1.0, 1.0), max_local_range=10.0)


def tracked_segment(path, closest_index, max_forward_distance, interp=0.01, segment_length=1.0):
    """ref: DWA::findTrackedPathSegment (src/controllers/dwa.cpp:208-233) with
    max_segment_size_ = path_segment_length/max_point_interpolation_distance + 1
    (src/controllers/follower.cpp:54-59)."""
    max_segment_size = int(segment_length / interp + 1)
    lookahead = max(max_segment_size, int(math.ceil(max_forward_distance / interp)) + 1)
    start = min(closest_index, path.n - 1)
    end = min(start + lookahead, path.n - 1)
    return start, end - start + 1


def cloud_bench(seed_off=0, n=100_000, center=(0.0, 0.0)):
    """bench.py's config-2 cloud: ring beyond the robot's reach (r >= 2.3 m) and the intruder wedge
    at 0.9-1.5 m, so roughly a quarter of the 10k samples collide and the rest go through all
    five cost terms (the expensive case for the evaluator)."""
    return cloud_c2(seed_off, n, center, r_min=2.3, intruder_r=(0.9, 1.5))


# ------------------------------------------------------------------------------------------------
# Config-2 cloud FAMILY. The reference's brute force is data independent; the pruned pipeline is
# not, so latency and parity are reported over distributions that stress its different stages and
# the bench headline is the worst of them.
# ------------------------------------------------------------------------------------------------
def _polar_cloud(rng, r, a, n, center=(0.0, 0.0), z=(0.0, 0.3)):
    zz = rng.uniform(z[0], z[1], n)
    pts = np.stack([center[0] + r * np.cos(a), center[1] + r * np.sin(a), zz], axis=1).astype(np.float32)
    return np.ascontiguousarray(pts[rng.permutation(n)])


def cloud_clutter(seed_off=0, n=100_000, center=(0.0, 0.0)):
    """Uniform clutter INSIDE the robot's reach: half of the points uniform (by area) over the annulus
    0.45-2.5 m around the start pose, half over 2.5-10 m. Most rollouts end in a collision; the
    survivors hug obstacles (every query cell has near neighbours, long candidate lists)."""
    rng = np.random.default_rng(SEED + 31 + seed_off)
    n_in = n // 2
    r = np.concatenate([np.sqrt(rng.uniform(0.45 ** 2, 2.5 ** 2, n_in)), rng.uniform(2.5, 10.0, n - n_in)])
    a = rng.uniform(0.0, 2 * math.pi, n)
    return _polar_cloud(rng, r, a, n, center)


def cloud_pillars(seed_off=0, n=100_000, center=(0.0, 0.0), n_pillars=40):
    """Sparse clutter inside reach: `n_pillars` thin pillars (5 cm blobs, 20 % of the points) dropped
    between 0.6 and 2.4 m, the rest of the points beyond 2.5 m. Many admissible rollouts thread
    between obstacles: the collision stage and the exact obstacle search both work."""
    rng = np.random.default_rng(SEED + 57 + seed_off)
    n_in = n // 5
    pr = rng.uniform(0.6, 2.4, n_pillars)
    pa = rng.uniform(0.0, 2 * math.pi, n_pillars)
    k = rng.integers(0, n_pillars, n_in)
    px = pr[k] * np.cos(pa[k]) + rng.normal(0.0, 0.025, n_in)
    py = pr[k] * np.sin(pa[k]) + rng.normal(0.0, 0.025, n_in)
    r_out = rng.uniform(2.5, 10.0, n - n_in)
    a_out = rng.uniform(0.0, 2 * math.pi, n - n_in)
    x = np.concatenate([px, r_out * np.cos(a_out)]) + center[0]
    y = np.concatenate([py, r_out * np.sin(a_out)]) + center[1]
    z = rng.uniform(0.0, 0.3, n)
    pts = np.stack([x, y, z], axis=1).astype(np.float32)
    return np.ascontiguousarray(pts[rng.permutation(n)])


def cloud_cluster(seed_off=0, n=100_000, center=(0.0, 0.0)):
    """One dense cluster ON the tracked path: 30 % of the points inside a 0.25 m blob at (1.3, 0.05)
    ahead of the robot (30 000 points in a handful of grid cells: per-cell candidate lists overflow
    their pool share), the rest on the far ring."""
    rng = np.random.default_rng(SEED + 83 + seed_off)
    n_in = (3 * n) // 10
    cx = 1.3 + rng.normal(0.0, 0.08, n_in)
    cy = 0.05 + rng.normal(0.0, 0.08, n_in)
    r_out = rng.uniform(2.3, 10.0, n - n_in)
    a_out = rng.uniform(0.0, 2 * math.pi, n - n_in)
    x = np.concatenate([cx, r_out * np.cos(a_out)]) + center[0]
    y = np.concatenate([cy, r_out * np.sin(a_out)]) + center[1]
    z = rng.uniform(0.0, 0.3, n)
    pts = np.stack([x, y, z], axis=1).astype(np.float32)
    return np.ascontiguousarray(pts[rng.permutation(n)])


def cloud_far(seed_off=0, n=100_000, center=(0.0, 0.0)):
    """Every point farther than reach + D (2 + 3.34 m) from the start pose: the obstacle term is 0 for
    every slot. With obstacle-only weights all admissible slots TIE at 0 and the branch and bound
    prunes nothing (the all-ties case)."""
    rng = np.random.default_rng(SEED + 101 + seed_off)
    r = rng.uniform(5.8, 10.0, n)
    a = rng.uniform(0.0, 2 * math.pi, n)
    return _polar_cloud(rng, r, a, n, center)


def cloud_empty(seed_off=0, n=0, center=(0.0, 0.0)):
    return np.zeros((0, 3), np.float32)


# name -> (generator, weights override or None). "survey_c2" is SURVEY 8(d)'s own C2 cloud.
CLOUD_FAMILY = {
    "survey_c2": (cloud_c2, None),
    "friendly_ring": (cloud_bench, None),
    "clutter_in_reach": (cloud_clutter, None),
    "pillars_in_reach": (cloud_pillars, None),
    "dense_cluster_on_path": (cloud_cluster, None),
    "all_ties_far_obstacles": (cloud_far, (0.0, 0.0, 1.0, 0.0, 0.0)),
    "empty_cloud": (cloud_empty, None),
}


def family_cloud(name, seed_off=0, n=100_000, center=(0.0, 0.0)):
    gen, w = CLOUD_FAMILY[name]
    return gen(seed_off, n=n, center=center) if name != "empty_cloud" else cloud_empty(), w


def heavy_trajectory_samples(prediction_horizon=10.0, time_step=0.01, n_samples=5001):
    """ref: benchmarks/benchmark_runner.cpp:36-90 generate_heavy_trajectory_samples: one straight row,
    then pairs with a sinusoidal lateral-velocity / heading fluctuation of growing amplitude
    (CostEvaluator_5k_Trajs: 5001 rows x 1000 points)."""
    P = int(prediction_horizon / time_step)
    v1, max_fl = 1.0, 0.5
    i = np.arange(P, dtype=np.float64)
    pairs = (n_samples - 1) // 2
    n = 1 + 2 * pairs
    x = np.zeros((n, P), np.float32)
    y = np.zeros((n, P), np.float32)
    vx = np.full((n, P - 1), v1, np.float32)
    vy = np.zeros((n, P - 1), np.float32)
    om = np.zeros((n, P - 1), np.float32)
    x[0] = time_step * v1 * i
    amp = (np.arange(1, pairs + 1, dtype=np.float64) * (max_fl / max(pairs, 1)))[:, None]
    fl_v = amp * np.sin(2 * math.pi * i / P)[None, :]
    x[1::2] = (time_step * v1 * i)[None, :]
    y[1::2] = time_step * fl_v * i[None, :]
    vy[1::2] = fl_v[:, :P - 1]
    fl_a = amp * np.cos(2 * math.pi * i / P)[None, :]
    x[2::2] = time_step * v1 * i[None, :] * np.cos(fl_a)
    y[2::2] = time_step * v1 * i[None, :] * np.sin(fl_a)
    om[2::2] = fl_a[:, :P - 1]
    return dict(x=x, y=y, vx=vx, vy=vy, omega=om)
