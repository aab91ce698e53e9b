"""Randomised parity sweep of the DWA cycle: configurations drawn at random (kinematics, limits,
horizon, sample counts, robot solid, sensor mount incl. yawed and upside-down ones, octree resolution,
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
