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
