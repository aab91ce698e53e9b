"""Stand-alone CollisionChecker through the C-ABI against the CPU oracle: the
reference's own three booleans (src/kompass_cpp/tests/collisions_test.cpp:11-77), then batched state
checks the way PurePursuit (pure_pursuit.cpp:154-155), the OMPL validity checker (ompl.cpp:95-97) and
TrajectorySampler::checkStatesFeasibility (trajectory_sampler.cpp:378-408) use the class. Booleans
are bit-exact against the oracle for every state."""
import math

import numpy as np
import pytest

import orc
import workloads as wl

pytestmark = pytest.mark.gpu

SHAPES = {"cylinder": (0, (0.25, 0.6, 0.0)), "box": (1, (0.6, 0.35, 0.8)), "sphere": (2, (0.3, 0.0, 0.0))}


def _orc_cfg(shape, dims, pos, rot, res):
    return orc.sampler_cfg(control_type=1, time_step=0.1, prediction_horizon=1.0, control_horizon=0.2,
                           max_linear_samples=3, max_angular_samples=3, vx=(1, 1, 1), vy=(0, 0, 0),
                           omega=(1, 1, 1), shape=shape, dims=dims, sensor_position=pos, sensor_rotation=rot,
                           octree_resolution=res, drop_samples=True, max_num_threads=1)


def test_reference_fcl_cases(pkg):
    """collisions_test.cpp: BOX 0.4 x 0.4 x 1.0, sensor at z = 1, octree 0.1"""
    # Eigen::Quaternionf{0,0,0,1} is the (w,x,y,z) ctor: coefficients (x,y,z,w) = (0,0,1,0)
    cc = pkg.CollisionChecker(1, (0.4, 0.4, 1.0), (0.0, 0.0, 1.0), (0, 0, 1, 0), 0.1)
    assert abs(cc.get_radius() - math.sqrt(0.32) / 2) < 1e-6
    cc.update_state(0.0, 0.0, 0.0)
    assert cc.check_collisions([1.0, 1.0, 1.0], [0.0, 0.1, 0.2]) is False
    cc.update_state(3.0, 5.0, 0.0)
    assert cc.check_collisions([0.25, 0.5, 0.5], [0.0, 0.1, 0.2]) is True
    cc.update_sensor_data(cloud=[(3.1, 5.1, -0.5)], global_frame=True)
    assert cc.check_collisions() is True
    assert cc.check_collisions((0.0, 0.0, 0.0)) is False
    cc.close()


@pytest.mark.parametrize("sensor", ["scan", "cloud_global", "cloud_local"])
@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_batched_states_match_oracle(pkg, shape, sensor):
    sh, dims = SHAPES[shape]
    yaw_mount = 0.6
    pos, rot = (0.15, -0.05, 0.3), (0.0, 0.0, math.sin(yaw_mount / 2), math.cos(yaw_mount / 2))
    res = 0.07
    rng = np.random.default_rng(wl.SEED + 31 + sh)
    body = (1.2, -0.7, 0.9)  # body pose when the sensor data arrives
    if sensor == "scan":
        n = 720
        ang = np.linspace(-math.pi, math.pi, n, endpoint=False)
        rngs = rng.uniform(0.4, 6.0, n)
        rngs[::37] = np.inf
        rngs[5::41] = np.nan
        data = dict(scan=(rngs, ang))
        centre = body[:2]
    else:
        pts = wl.cloud_c2(7, n=1_500)
        pts[:, 2] = rng.uniform(-0.6, 0.9, len(pts)).astype(np.float32)
        data = dict(cloud=pts)
        centre = (0.0, 0.0) if sensor == "cloud_global" else body[:2]
    gf = sensor != "cloud_local"
    n_states = 4000
    states = np.zeros((n_states, 3))
    states[:, 0] = centre[0] + rng.uniform(-7, 7, n_states)
    states[:, 1] = centre[1] + rng.uniform(-7, 7, n_states)
    states[:, 2] = rng.uniform(-math.pi, math.pi, n_states)
    states[3] = (np.nan, 0.0, 0.0)
    states[4] = (np.inf, 1.0, 0.0)
    cfg = _orc_cfg(sh, dims, pos, rot, res)
    ref_any, ref = orc.check_collision_states(cfg, body, states, global_frame=gf, **data)
    cc = pkg.CollisionChecker(sh, dims, pos, rot, res)
    cc.update_state(*body)
    cc.update_sensor_data(global_frame=gf, **data)
    got_any, got = cc.check_states(states)
    assert got_any == ref_any
    assert np.array_equal(got, ref), f"{(got != ref).sum()} of {n_states} booleans differ"
    assert 0.01 < ref.mean() < 0.99  # the scene exercises both answers
    # a later, tighter batch is answered from the cached bitmap; a far-away batch rebuilds it
    for lo, hi in [(-1.0, 1.0), (20.0, 25.0), (-3.0, 3.0)]:
        s2 = states[:500].copy()
        s2[:, 0] = centre[0] + rng.uniform(lo, hi, 500)
        s2[:, 1] = centre[1] + rng.uniform(lo, hi, 500)
        r_any, r = orc.check_collision_states(cfg, body, s2, global_frame=gf, **data)
        g_any, g = cc.check_states(s2)
        assert g_any == r_any and np.array_equal(g, r), (lo, hi, (g != r).sum())
    # single-state overloads: updateState + checkCollisions(), and checkCollisions(state)
    for i in range(0, 60, 7):
        cc.update_state(*states[i + 5])
        assert cc.check_collisions() == bool(ref[i + 5])
        assert cc.check_collisions(tuple(states[i + 5])) == bool(ref[i + 5])
    cc.close()


@pytest.mark.parametrize("rot", [(1.0, 0.0, 0.0, 0.0), (0.0, 1.0, 0.0, 0.0)])
def test_upside_down_mount_mirrors_the_scan(pkg, rot):
    """180 deg about x maps a sensor-frame hit at (x, y) to body (x, -y); about y to (-x, y)"""
    cc = pkg.CollisionChecker(0, (0.2, 1.0), (0.0, 0.0, 0.1), rot, 0.05)
    cc.update_state(0.0, 0.0, 0.0)
    hit = (1.0, 0.6)
    ang, rng_ = math.atan2(hit[1], hit[0]), math.hypot(*hit)
    cc.update_sensor_data(scan=([rng_], [ang]))
    mirrored = (hit[0], -hit[1]) if rot[0] == 1.0 else (-hit[0], hit[1])
    assert cc.check_collisions((mirrored[0], mirrored[1], 0.0)) is True
    assert cc.check_collisions((hit[0], hit[1], 0.0)) is False
    cfg = _orc_cfg(0, (0.2, 1.0, 0.0), (0.0, 0.0, 0.1), rot, 0.05)
    rng = np.random.default_rng(3)
    states = np.column_stack([rng.uniform(-1.5, 1.5, 500), rng.uniform(-1.5, 1.5, 500), rng.uniform(-3, 3, 500)])
    ref_any, ref = orc.check_collision_states(cfg, (0, 0, 0), states, scan=([rng_], [ang]))
    got_any, got = cc.check_states(states)
    assert np.array_equal(got, ref) and ref_any
    cc.close()


def test_rollout_feasibility_like_pure_pursuit(pkg):
    """pure_pursuit.cpp:140-160: simulate states along an arc until the first collision"""
    cc = pkg.CollisionChecker(0, (0.2, 0.5), (0.0, 0.0, 0.0), (0, 0, 0, 1), 0.05)
    cloud = wl.round_obstacle(1.5, 0.4, 0.3)
    cc.update_state(0.0, 0.0, 0.0)
    cc.update_sensor_data(cloud=cloud)
    cfg = _orc_cfg(0, (0.2, 0.5, 0.0), (0, 0, 0), (0, 0, 0, 1), 0.05)
    x = y = yaw = 0.0
    states = []
    for _ in range(60):
        x += 0.8 * math.cos(yaw) * 0.05
        y += 0.8 * math.sin(yaw) * 0.05
        yaw += 0.5 * 0.05
        states.append((x, y, yaw))
    ref_any, ref = orc.check_collision_states(cfg, (0, 0, 0), states, cloud=cloud)
    got_any, got = cc.check_states(states)
    assert ref_any and got_any and np.array_equal(got, ref)
    first = int(np.argmax(ref))
    assert 0 < first < 59 and not ref[:first].any()
    cc.close()


def test_edge_cases(pkg):
    cc = pkg.CollisionChecker(0, (0.2, 0.5), (0, 0, 0), (0, 0, 0, 1), 0.1)
    assert cc.check_collisions() is False  # no sensor data yet
    any_, out = cc.check_states(np.zeros((0, 3)))
    assert not any_ and len(out) == 0
    cc.update_sensor_data(cloud=np.zeros((0, 3), np.float32))
    assert cc.check_states([(0, 0, 0)])[0] is False
    cc.update_sensor_data(cloud=[(0.1, 0.0, 0.0)])
    assert cc.check_states([(0, 0, 0)])[0] is True
    # resolution change applies from the next sensor update
    cc.reset_octree_resolution(0.5)
    assert cc.check_states([(0.75, 0.0, 0.0)])[0] is False
    cc.update_sensor_data(cloud=[(0.1, 0.0, 0.0)])  # voxel [0, 0.5]^3 now
    cfg = _orc_cfg(0, (0.2, 0.5, 0.0), (0, 0, 0), (0, 0, 0, 1), 0.5)
    st = [(0.65, 0.0, 0.0), (0.75, 0.0, 0.0), (0.69999, 0.0, 0.0)]
    ref_any, ref = orc.check_collision_states(cfg, (0, 0, 0), st, cloud=[(0.1, 0.0, 0.0)])
    got_any, got = cc.check_states(st)
    assert np.array_equal(got, ref) and got[0] == 1 and got[1] == 0
    with pytest.raises(pkg.KompassB200Error):  # not a rotation (|q| = 0.996): rejected loudly
        bad = pkg.CollisionChecker(0, (0.2, 0.5), (0, 0, 0), (0.3, 0.0, 0.0, 0.95), 0.1)
        bad.update_sensor_data(scan=([1.0], [0.0]))
    with pytest.raises(ValueError):
        pkg.CollisionChecker(7, (0.2, 0.5))
    cc.close()



def _unit_quat(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return tuple(float(np.float32(v)) for v in q)


TILTS = {
    "pitch_20deg": (0.0, math.sin(math.radians(10)), 0.0, math.cos(math.radians(10))),
    "roll_35deg_yaw": None,  # filled below: roll about x composed with a yaw
    "random_a": _unit_quat(np.random.default_rng(wl.SEED + 71)),
    "random_b": _unit_quat(np.random.default_rng(wl.SEED + 72)),
}
_r, _y = math.radians(35) / 2, 0.8 / 2
TILTS["roll_35deg_yaw"] = (math.sin(_r) * math.cos(_y), math.sin(_r) * math.sin(_y), math.cos(_r) * math.sin(_y),
                           math.cos(_r) * math.cos(_y))


@pytest.mark.parametrize("sensor", ["scan", "cloud_local"])
@pytest.mark.parametrize("tilt", sorted(TILTS))
@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_tilted_sensor_mounts_match_oracle(pkg, shape, tilt, sensor):
    """Pitched / rolled / arbitrary sensor mounts (a pitched lidar, a depth camera): the octree's voxel
    cubes are oriented boxes in the robot's frame (collision_check.cpp:118-123 hands FCL the full
    sensor_tf_world_). Sphere vs OBB, box vs OBB (15 axes) and cylinder vs OBB (slab clip + hull
    distance) against the oracle's general voxel model: every boolean identical."""
    sh, dims = SHAPES[shape]
    pos, rot = (0.12, -0.04, 0.35), TILTS[tilt]
    res = 0.09
    rng = np.random.default_rng(wl.SEED + 131 + sh)
    body = (0.8, -0.5, 0.7)
    if sensor == "scan":
        n = 540
        ang = np.linspace(-math.pi, math.pi, n, endpoint=False)
        rngs = rng.uniform(0.3, 4.0, n)
        rngs[::29] = np.inf
        data = dict(scan=(rngs, ang))
    else:
        pts = np.stack([rng.uniform(-3, 3, 2500), rng.uniform(-3, 3, 2500), rng.uniform(-1.0, 1.0, 2500)], 1).astype(np.float32)
        data = dict(cloud=pts)
    n_states = 3000
    states = np.zeros((n_states, 3))
    states[:, 0] = body[0] + rng.uniform(-4, 4, n_states)
    states[:, 1] = body[1] + rng.uniform(-4, 4, n_states)
    states[:, 2] = rng.uniform(-math.pi, math.pi, n_states)
    states[7] = (np.nan, 0.0, 0.0)
    cfg = _orc_cfg(sh, dims, pos, rot, res)
    ref_any, ref = orc.check_collision_states(cfg, body, states, global_frame=False, **data)
    cc = pkg.CollisionChecker(sh, dims, pos, rot, res)
    cc.update_state(*body)
    cc.update_sensor_data(global_frame=False, **data)
    got_any, got = cc.check_states(states)
    assert got_any == ref_any
    assert np.array_equal(got, ref), f"{(got != ref).sum()} of {n_states} booleans differ"
    assert 0.005 < ref.mean() < 0.995
    s2 = states[:400].copy()  # a far-away batch rebuilds the 3-D window
    s2[:, :2] += 30.0
    r_any, r = orc.check_collision_states(cfg, body, s2, global_frame=False, **data)
    g_any, g = cc.check_states(s2)
    assert g_any == r_any and np.array_equal(g, r)
    cc.close()
