"""Randomised parity sweeps (DWA cycle first; mapper, cloud binning and critical zone at the end): configurations drawn at random (kinematics, limits,
horizon, sample counts, robot solid, sensor mount incl. yawed, upside-down and tilted ones, octree resolution,
dropping mode, weights incl. zeros, velocity, pose, scan or cloud) run through the C-ABI and the CPU
oracle. Every case is checked twice by check_cycle: every slot evaluated exactly (all per-slot costs
bit-identical) and with the branch and bound forced on (winner identical, pruned bounds valid)."""
import math

import numpy as np
import pytest

import orc
import workloads as wl
from test_gpu_planner import check_cycle

pytestmark = pytest.mark.gpu


def _draw(seed, big=False):
    rng = np.random.default_rng(90_000 + seed)
    ctrl = int(rng.integers(0, 3))
    dt = float(rng.choice([0.05, 0.1, 0.2]))
    steps = int(rng.integers(4, 40))
    n_lin, n_ang = int(rng.integers(3, 26)), int(rng.integers(3, 26))
    if big:  # thousands of slots: the default branch and bound and the reach mask are in play
        n_lin, n_ang = int(rng.integers(50, 90)), int(rng.integers(50, 90))
    vmax = float(rng.uniform(0.4, 2.5))
    shape = int(rng.integers(0, 3))
    dims = {0: (float(rng.uniform(0.08, 0.4)), float(rng.uniform(0.3, 1.5)), 0.0),
            1: (float(rng.uniform(0.2, 0.9)), float(rng.uniform(0.15, 0.6)), float(rng.uniform(0.3, 1.5))),
            2: (float(rng.uniform(0.1, 0.5)), 0.0, 0.0)}[shape]
    yaw_m = float(rng.uniform(-math.pi, math.pi)) if rng.random() < 0.6 else 0.0
    qz, qw = math.sin(yaw_m / 2), math.cos(yaw_m / 2)
    flip = rng.random() < 0.25  # upside-down mount: q = q_z(yaw) * q_x(pi) = (cos, sin, 0, 0) of yaw / 2
    rot = (qw, qz, 0.0, 0.0) if flip else (0.0, 0.0, qz, qw)
    trng = np.random.default_rng(190_000 + seed)  # (own stream: the other draws keep their values)
    if trng.random() < 0.3:  # tilted mount (pitched lidar, depth camera): any unit quaternion
        q = trng.normal(size=4)
        q /= np.linalg.norm(q)
        rot = tuple(float(v) for v in q)
    weights = tuple(float(w) for w in np.where(rng.random(5) < 0.2, 0.0, rng.uniform(0.1, 4.0, 5)))
    kw = dict(control_type=ctrl, time_step=dt, prediction_horizon=steps * dt,
              control_horizon=float(rng.integers(1, 4)) * dt, max_linear_samples=n_lin,
              max_angular_samples=n_ang, vx=(vmax, float(rng.uniform(0.5, 20)), float(rng.uniform(0.5, 20))),
              vy=(float(rng.uniform(0.3, 1.5)), float(rng.uniform(0.5, 10)), float(rng.uniform(0.5, 10))),
              omega=(float(rng.uniform(0.5, 4.0)), float(rng.uniform(0.5, 20)), float(rng.uniform(0.5, 20))),
              shape=shape, dims=dims,
              sensor_position=(float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.3, 0.3)), float(rng.uniform(0.0, 0.6))),
              sensor_rotation=rot, octree_resolution=float(rng.choice([0.04, 0.07, 0.1, 0.2])),
              drop_samples=bool(rng.random() < 0.6), weights=weights,
              max_local_range=float(rng.uniform(4.0, 15.0)))
    vel = (float(rng.uniform(-vmax, vmax)), float(rng.uniform(-0.5, 0.5)) if ctrl == 2 else 0.0,
           float(rng.uniform(-2.0, 2.0)))
    pts = [wl.straight_points(12.0), wl.uturn_points(), wl.circle_test_points()][int(rng.integers(0, 3))]
    path = orc.Path(pts, 0.01, 1.0)
    closest = int(rng.integers(0, max(1, path.n // 2)))
    pose = (float(path.X[closest] + rng.uniform(-0.4, 0.4)), float(path.Y[closest] + rng.uniform(-0.4, 0.4)),
            float(rng.uniform(-math.pi, math.pi)))
    seg = wl.tracked_segment(path, closest, vmax * steps * dt)
    reach = vmax * steps * dt
    clear = max(dims[0], dims[1]) + float(rng.uniform(0.15, 0.8))  # nothing closer than this at the start
    if rng.random() < 0.5:
        n = int(rng.integers(1, 120))
        ang = np.sort(rng.uniform(-math.pi, math.pi, n))
        rr = rng.uniform(clear + 0.4, max(clear + 0.8, 1.5 * reach + 3.0), n)
        rr[rng.random(n) < 0.05] = np.inf
        rr[rng.random(n) < 0.03] = np.nan
        data = dict(scan=(rr, ang))
    else:
        n = int(rng.integers(1, 400)) * (4 if big else 1)
        r = rng.uniform(clear, max(clear + 0.5, 1.5 * reach + 4.0), n)
        th = rng.uniform(0, 2 * math.pi, n)
        cloud = np.stack([pose[0] + r * np.cos(th), pose[1] + r * np.sin(th), rng.uniform(-0.8, 1.2, n)], 1)
        data = dict(cloud=cloud.astype(np.float32))
    return kw, path, seg, vel, pose, data


@pytest.mark.parametrize("seed", range(48))
def test_random_configuration_matches_oracle(pkg, seed):
    kw, path, seg, vel, pose, data = _draw(seed)
    got, ref = check_cycle(pkg, kw, path, seg, vel, pose, **data)
    assert got.n_slots >= 0 and got.n_admissible == ref["n_admissible"]


@pytest.mark.parametrize("seed", range(100, 108))
def test_random_large_configuration_matches_oracle(pkg, seed):
    kw, path, seg, vel, pose, data = _draw(seed, big=True)
    got, ref = check_cycle(pkg, kw, path, seg, vel, pose, **data)
    assert got.n_slots > 2048 and got.n_admissible == ref["n_admissible"]
    # the default settings (branch and bound on automatically at this size) give the same winner
    from parity_util import make_planner
    pl = make_planner(pkg, kw, path)
    d = pl.cycle_scan(vel, pose, data["scan"][0], data["scan"][1], seg[0], seg[1]) if "scan" in data \
        else pl.cycle_cloud(vel, pose, data["cloud"], seg[0], seg[1])
    assert (d.is_found, d.slot, d.n_admissible) == (got.is_found, got.slot, got.n_admissible)
    assert np.float32(d.cost) == np.float32(got.cost)
    if d.is_found:
        assert pl.fetch_pruned(d.n_slots).sum() >= 0
    pl.close()


# ------------------------------------------------------------------------------------------------
# mapper / cloud binning / critical zone: random configurations, every output bit-exact
# ------------------------------------------------------------------------------------------------
def _cloud_layout(rng, pts):
    """pack [n, 3] float32 points into a PointCloud2-style byte buffer with a random layout"""
    n = len(pts)
    h = int(rng.choice([1, 1, 2, 3]))
    w = n // h
    offs = sorted(rng.choice(np.arange(0, 20), 3, replace=False) * 1)  # may be unaligned
    offs = [int(offs[0]), int(offs[0] + 4 + rng.integers(0, 3)), 0]
    offs[2] = int(offs[1] + 4 + rng.integers(0, 3))
    perm = rng.permutation(3)
    xo, yo, zo = (offs[perm[0]], offs[perm[1]], offs[perm[2]])
    ps = max(offs) + 4 + int(rng.integers(0, 6))
    rs = w * ps + int(rng.integers(0, 9))
    buf = np.zeros(h * rs, np.uint8)
    raw = np.ascontiguousarray(pts[:h * w, :3], np.float32).view(np.uint8).reshape(h * w, 12)
    for r in range(h):
        rows = raw[r * w:(r + 1) * w]
        base = r * rs + np.arange(w) * ps
        for k, off in enumerate((xo, yo, zo)):
            for b in range(4):
                buf[base + off + b] = rows[:, 4 * k + b]
    return buf.view(np.int8), ps, rs, h, w, xo, yo, zo


def _random_cloud(rng, n, rmax):
    r = rng.uniform(0.0, rmax, n) ** rng.choice([1.0, 2.0]) / rmax ** rng.choice([0.0, 0.0])
    r = np.minimum(r, rmax)
    a = rng.uniform(-math.pi, math.pi, n)
    pts = np.zeros((n, 3), np.float32)
    pts[:, 0], pts[:, 1] = r * np.cos(a), r * np.sin(a)
    pts[:, 2] = rng.uniform(-0.5, 3.0, n)
    pts[rng.random(n) < 0.01] = 0.0  # points at the origin are dropped (pointcloud.h r^2 < 1e-6)
    return pts


@pytest.mark.parametrize("seed", range(24))
def test_random_mapper_configuration_matches_oracle(pkg, seed):
    rng = np.random.default_rng(70_000 + seed)
    H, W = int(rng.integers(8, 520)), int(rng.integers(8, 520))
    res = float(rng.choice([0.02, 0.05, 0.1, 0.25]))
    span = 0.5 * min(H, W) * res
    pos = (float(rng.uniform(-0.6, 0.6) * span), float(rng.uniform(-0.6, 0.6) * span), float(rng.uniform(0, 0.5)))
    orient = float(rng.uniform(-math.pi, math.pi))
    range_max = float(rng.choice([5.0, 20.0, 40.0]))
    n = int(rng.integers(1, 2400))
    if rng.random() < 0.5:
        angles = np.array([-math.pi + 2 * math.pi * i / n for i in range(n)])
    else:
        angles = np.sort(rng.uniform(-2 * math.pi, 2 * math.pi, n))
    ranges = rng.uniform(0.0, rng.choice([0.5, 1.5, 3.0]) * span + 0.1, n)
    ranges[rng.random(n) < 0.05] = 0.0
    ranges[rng.random(n) < 0.05] = 3.0 * range_max  # beyond the grid and the sensor range
    bay = dict(p_prior=float(rng.choice([0.5, 0.6, 0.3])), p_occupied=float(rng.uniform(0.55, 0.95)),
               p_empty=float(rng.uniform(0.05, 0.45)), range_sure=float(rng.uniform(0.05, 2.0)),
               range_max=range_max, wall_size=float(rng.uniform(0.05, 0.5)))
    mp = pkg.LocalMapperGPU(H, W, res, pos, orient, False, n, 0.01, 2.0, 0.1, range_max, 256)
    mp.set_bayesian_params(bay["p_prior"], bay["p_occupied"], bay["p_empty"], bay["range_sure"], bay["wall_size"])
    ref = orc.mapper_scan_to_grid(H, W, res, pos, orient, angles, ranges)
    got = mp.scan_to_grid(angles, ranges)
    assert np.array_equal(got, ref), f"{(got != ref).sum()} cells differ"
    prev = None
    for k in range(3):  # Bayesian update, feedback, robot motion + warp, update again
        g_ref, p_ref = orc.mapper_scan_to_grid_bayes(H, W, res, pos, orient, angles, ranges, prev=prev, **bay)
        g, p = mp.scan_to_grid_baysian(angles, ranges)
        assert np.array_equal(g, g_ref), (k, (g != g_ref).sum())
        assert np.array_equal(p.view(np.uint32), p_ref.view(np.uint32)), (k, np.abs(p - p_ref).max())
        mp.set_previous_grid(p)
        move = (float(rng.uniform(-3, 3) * res), float(rng.uniform(-3, 3) * res))
        yaw = float(rng.uniform(-0.5, 0.5))
        mp.get_previous_grid_in_current_pose(move, yaw)
        prev = orc.mapper_warp_previous(H, W, res, bay["p_prior"], move, yaw, p_ref)
        w = mp.get_previous_grid()
        assert np.array_equal(w.view(np.uint32), prev.view(np.uint32)), (k, np.abs(w - prev).max())
        ranges = np.clip(ranges + rng.normal(0, 0.05, n), 0.0, None)
    mp.close()


@pytest.mark.parametrize("seed", range(16))
def test_random_cloud_binning_and_cloud_mapper_match_oracle(pkg, seed):
    rng = np.random.default_rng(71_000 + seed)
    n = int(rng.integers(1, 30_000))
    range_max = float(rng.choice([5.0, 20.0]))
    pts = _random_cloud(rng, n, 1.3 * range_max)
    data, ps, rs, h, w, xo, yo, zo = _cloud_layout(rng, pts)
    min_z = float(rng.uniform(-0.5, 0.5))
    max_z = float(rng.choice([-1.0, rng.uniform(0.6, 3.0)]))
    bins = int(rng.integers(1, 4000))
    got = pkg.pointcloud_to_laserscan(data, ps, rs, h, w, xo, yo, zo, range_max, min_z, max_z, bins)
    ref = orc.pointcloud_to_laserscan(data, ps, rs, h, w, xo, yo, zo, range_max, min_z, max_z, bins)
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (got != ref).sum()
    step = float(rng.choice([np.float32(0.01), 0.05, 2 * math.pi / 720, rng.uniform(0.002, 0.5)]))
    got_r, got_a = pkg.pointcloud_to_laserscan_step(data, ps, rs, h, w, xo, yo, zo, range_max, min_z, max_z, step)
    ref_r, ref_a = orc.pointcloud_to_laserscan_step(data, ps, rs, h, w, xo, yo, zo, range_max, min_z, max_z, step)
    assert np.array_equal(got_a.view(np.uint64), ref_a.view(np.uint64))
    assert np.array_equal(got_r.view(np.uint64), ref_r.view(np.uint64)), (got_r != ref_r).sum()
    # the mapper's raw-cloud overloads: binning (num_bins = scan_size / angle_step of the ctor) + rays
    H, W, res = int(rng.integers(40, 420)), int(rng.integers(40, 420)), float(rng.choice([0.05, 0.1]))
    orient = float(rng.uniform(-1, 1))
    pos = (float(rng.uniform(-0.5, 0.5)), float(rng.uniform(-0.5, 0.5)), 0.0)
    scan_size = int(rng.integers(16, 1500))
    astep = float(np.float32(rng.choice([0.01, 0.02, 0.005])))
    mp = pkg.LocalMapperGPU(H, W, res, pos, orient, True, scan_size, astep, max_z, min_z, range_max, 256)
    g = mp.scan_to_grid(data, ps, rs, h, w, xo, yo, zo)
    rr = orc.pointcloud_to_laserscan(data, ps, rs, h, w, xo, yo, zo, range_max, min_z, max_z, scan_size)
    aa = np.array([i * (2.0 * math.pi) / scan_size for i in range(scan_size)])
    g_ref = orc.mapper_scan_to_grid(H, W, res, pos, orient, aa, rr)
    assert np.array_equal(g, g_ref), f"{(g != g_ref).sum()} cells differ"
    rb, ab = orc.pointcloud_to_laserscan_step(data, ps, rs, h, w, xo, yo, zo, range_max, min_z, max_z, astep)
    gb_ref, pb_ref = orc.mapper_scan_to_grid_bayes(H, W, res, pos, orient, ab, rb, range_max=range_max)
    gb, pb = mp.scan_to_grid_baysian(data, ps, rs, h, w, xo, yo, zo)
    assert np.array_equal(gb, gb_ref), f"{(gb != gb_ref).sum()} cells differ"
    assert np.array_equal(pb.view(np.uint32), pb_ref.view(np.uint32))
    mp.close()


@pytest.mark.parametrize("seed", range(32))
def test_random_critical_zone_matches_oracle(pkg, seed):
    rng = np.random.default_rng(72_000 + seed)
    shape = int(rng.integers(0, 3))  # 0 cylinder, 1 box, 2 sphere (RobotGeometry order of the package)
    dims = {0: (float(rng.uniform(0.1, 0.8)), float(rng.uniform(0.2, 2.0))),
            1: (float(rng.uniform(0.2, 1.2)), float(rng.uniform(0.2, 1.2)), float(rng.uniform(0.2, 2.0))),
            2: (float(rng.uniform(0.1, 0.8)),)}[shape]
    pos = (float(rng.uniform(-0.4, 0.4)), float(rng.uniform(-0.3, 0.3)), float(rng.uniform(0.0, 0.6)))
    yaw = float(rng.uniform(-math.pi, math.pi))
    scale = float(rng.choice([1.0, 0.99, 1.0]))  # the reference test passes an un-normalised quaternion
    rot = (0.0, 0.0, scale * math.sin(yaw / 2), scale * math.cos(yaw / 2))
    if rng.random() < 0.25:
        rot = (0.0, 0.0, 0.0, 1.0)
    crit_angle = float(rng.choice([20.0, 90.0, 160.0, 180.0, 270.0, 360.0, rng.uniform(1.0, 360.0)]))
    crit_d = float(rng.uniform(0.05, 0.5))
    slow_d = crit_d + float(rng.uniform(0.05, 1.0))
    n = int(rng.integers(1, 4000))
    angles = np.array([2.0 * math.pi * i / n for i in range(n)]) if rng.random() < 0.7 else \
        np.sort(rng.uniform(-math.pi, 2 * math.pi, n))
    min_h, max_h, range_max = float(rng.uniform(-0.2, 0.3)), float(rng.uniform(0.8, 2.5)), 20.0
    kw = dict(shape=shape, dims=dims, sensor_position=pos, sensor_rotation=rot, critical_angle=crit_angle,
              critical_distance=crit_d, slowdown_distance=slow_d, min_height=min_h, max_height=max_h,
              range_max=range_max)
    cfg = orc.cz_cfg(**kw)
    radius = math.hypot(dims[0], dims[1]) / 2 if shape == 1 else dims[0]
    z = pkg.CriticalZoneCheckerGPU(0, shape, dims, pos, rot, crit_angle, crit_d, slow_d, angles, min_h, max_h,
                                   range_max)
    for lo in (slow_d + 0.4, crit_d + 0.02, 0.0):  # clear / slowdown band / inside the critical zone
        ranges = radius + lo + rng.uniform(0.0, 1.5, n) ** 2
        for fwd in (True, False):
            g = z.check(ranges, fwd)
            assert g == orc.cz_check_scan(cfg, angles, ranges, fwd), (lo, fwd)
            assert 0.0 <= g <= 1.0
    # mostly clear with a handful of rays inside the slowdown band: intermediate factors
    for _ in range(4):
        ranges = np.full(n, 10.0)
        k = rng.integers(0, n, min(n, 6))
        ranges[k] = radius + math.hypot(pos[0], pos[1]) + crit_d + rng.uniform(0.0, slow_d - crit_d, len(k))
        for fwd in (True, False):
            assert z.check(ranges, fwd) == orc.cz_check_scan(cfg, angles, ranges, fwd)
    z.close()
    zc = pkg.CriticalZoneCheckerGPU(1, shape, dims, pos, rot, crit_angle, crit_d, slow_d, angles, min_h, max_h,
                                    range_max)
    for lo in (slow_d + 0.4, crit_d + 0.02, 0.0):
        m = int(rng.integers(1, 20_000))
        pts = _random_cloud(rng, m, 6.0)
        rr = np.hypot(pts[:, 0], pts[:, 1])
        keep = rr >= radius + lo + 0.3  # (sensor offset up to 0.5 m is not compensated: bands overlap)
        pts = pts[keep] if keep.any() else pts[:1]
        data, ps, rs, h, w, xo, yo, zo = _cloud_layout(rng, pts)
        for fwd in (True, False):
            g = zc.check(data, ps, rs, h, w, xo, yo, zo, fwd)
            assert g == orc.cz_check_cloud(cfg, angles, data, ps, rs, h, w, xo, yo, zo, fwd), (lo, fwd, m)
    zc.close()
