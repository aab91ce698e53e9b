"""Closed-loop DWA controller through the C-ABI: the 18 scenarios of the
reference's own closed-loop test (src/kompass_cpp/tests/dwa_test.cpp:161-362: 3 robot types x 3 paths
x obstacle avoidance on/off) driven by the CUDA path, with the reference's assertions (a trajectory is
found at every step, the goal is reached, clearance >= robot radius) and, in lock-step, the CPU oracle:
the follower state (closest index, segment, tracked view, adaptive horizon) is compared at EVERY
step, the planner cycle (winner slot bit-exact, cost, winner rows) at a subset of steps."""
import math

import numpy as np
import pytest

import workloads as wl
from orc_follower import FollowerOracle
from parity_util import assert_cycle_parity, run_oracle_cycle

pytestmark = pytest.mark.gpu

PATHS = {"Straight": wl.straight_test_points, "UTurn": wl.uturn_points, "Circle": wl.circle_test_points}
OBSTACLES = {"Straight": (4.0, 0.0), "UTurn": (10.0, 0.0), "Circle": (5.0, 8.5)}
ROBOTS = {"Ackermann": 0, "DiffDrive": 1, "Omni": 2}
RADIUS = 0.1


def scenario_cfg(control_type):  # ref: dwa_test.cpp:162-205
    return dict(control_type=control_type, time_step=0.1, prediction_horizon=4.0, control_horizon=0.5,
                max_linear_samples=20, max_angular_samples=20, vx=(1.0, 2.0, 2.0), vy=(1.0, 2.0, 2.0),
                omega=(2.0, 3.0, 3.0), shape=0, dims=(RADIUS, 0.4, 0.0), sensor_position=(0, 0, 0),
                sensor_rotation=(0, 0, 0, 1), octree_resolution=0.1, drop_samples=True,
                weights=(1.0, 1.0, 0.0, 0.0, 0.0), max_local_range=10.0)


def apply_control(state, cmd, dt):  # ref: controller_test_helpers.h:12-31
    x, y, yaw = state
    dx = (cmd[0] * math.cos(yaw) - cmd[1] * math.sin(yaw)) * dt
    dy = (cmd[0] * math.sin(yaw) + cmd[1] * math.cos(yaw)) * dt
    x, y, yaw = x + dx, y + dy, yaw + cmd[2] * dt
    while yaw > math.pi:
        yaw -= 2.0 * math.pi
    while yaw < -math.pi:
        yaw += 2.0 * math.pi
    return (x, y, yaw)


@pytest.mark.parametrize("avoid", [False, True])
@pytest.mark.parametrize("robot", sorted(ROBOTS))
@pytest.mark.parametrize("path_name", sorted(PATHS))
def test_closed_loop_scenario(pkg, avoid, robot, path_name):
    kw = scenario_cfg(ROBOTS[robot])
    pts = PATHS[path_name]()
    dwa = pkg.DWA(pkg.planner_config(**kw), pkg.follower_params(goal_dist_tolerance=0.3))
    if avoid:  # half of the scenarios with the branch and bound forced on (off by default at 441 slots)
        dwa.planner.set_tuning(7, 2)
    ref = FollowerOracle(kw, goal_dist_tolerance=0.3)
    dwa.set_current_path(pts)
    ref.set_current_path(pts)
    got_path = dwa.get_current_path()
    assert np.array_equal(got_path["X"], ref.path.X) and np.array_equal(got_path["Y"], ref.path.Y)
    assert got_path["n_segments"] == len(ref.path.seg_starts)

    p0, p1 = np.float32(pts[0]), np.float32(pts[1])
    start_yaw = float(np.arctan2(np.float32(p1[1] - p0[1]), np.float32(p1[0] - p0[0])))
    state = (float(p0[0]), float(p0[1]), start_yaw)
    if path_name == "Circle":
        state = (state[0] + 0.2, state[1], state[2])
    cloud = wl.round_obstacle(*OBSTACLES[path_name], 0.3) if avoid else np.zeros((0, 3), np.float32)

    vel = (0.0, 0.0, 0.0)
    min_clear, goal, step, checked = math.inf, False, 0, 0
    while not goal and step < 1000:
        dwa.set_current_state(*state)
        ref.set_current_state(*state)
        res = dwa.compute_velocity_commands(vel, cloud=cloud)
        seg = ref.prepare()
        info = dwa.info
        tag = f"{robot}/{path_name}/{avoid} step {step}"
        assert (info.closest_index, info.segment_index) == (ref.c_index, ref.current_segment_index), tag
        assert (info.seg_start, info.seg_count) == seg, tag
        assert info.n_points == ref.n_points and info.horizon == ref.horizon, tag
        assert info.segment_position == ref.c_seglen, tag
        assert abs(info.crosstrack_error - ref.c_parallel) <= 1e-12 + 1e-9 * abs(ref.c_parallel), tag
        assert abs(info.heading_error - ref.heading_error) <= 1e-6, tag
        assert res.is_found, "DWA failed to find trajectory: " + tag
        if step < 4 or step % 3 == 0:
            oracle = ref.run_cycle(vel, seg, cloud=cloud)
            assert_cycle_parity(res, oracle)
            assert np.float32(res.cost) == np.float32(oracle["cost"]), tag
            checked += 1
        cmd = (dwa.get_vx_cmd(), dwa.get_vy_cmd(), dwa.get_omega_cmd())
        # the getters clamp to the BASE-CLASS limits (1 m/s, 1 rad/s: DWA never sets them, follower.h:147-165)
        clamp = lambda v: max(min(float(v), 1.0), -1.0)
        assert cmd == (clamp(res.vx[0]), clamp(res.vy[0]), clamp(res.omega[0]))
        vel = cmd
        state = apply_control(state, cmd, kw["time_step"])
        if avoid:
            d = np.hypot(cloud[:, 0].astype(np.float64) - state[0], cloud[:, 1].astype(np.float64) - state[1])
            min_clear = min(min_clear, float(d.min()))
        # as in the reference loop (dwa_test.cpp:259-284) the goal check runs before the next
        # setCurrentState: it sees the state this step's command was computed from
        g1, g2 = dwa.is_goal_reached(), ref.is_goal_reached()
        assert g1 == g2, tag
        goal = g1
        step += 1
    assert goal, f"DWA did not reach goal: {robot}/{path_name}/{avoid} after {step} steps"
    assert checked >= 4
    if avoid:
        assert min_clear >= RADIUS, f"DWA collided with obstacle: clearance {min_clear}"
    dwa.close()


# ---------------------------------------------------------------------------------------------
# DWA::addCustomCost / debugVelocitySearch / getDebuggingSamples (bindings_control.cpp:256-271)
# ---------------------------------------------------------------------------------------------
def _heading_cost(t, path):  # prefers trajectories that end far along +y; float like CustomCostFunction
    return np.float32(3.0) - np.float32(t["y"][-1]) + np.float32(0.01) * np.float32(path["total_length"])


def _turn_cost(t, path):
    return np.float32(abs(float(t["omega"][0]))) * np.float32(0.37) + np.float32(len(path["X"])) * np.float32(1e-4)


@pytest.mark.parametrize("sensor", ["scan", "cloud"])
@pytest.mark.parametrize("robot", ["DiffDrive", "Omni"])
def test_custom_costs_match_oracle(pkg, robot, sensor):
    """A cycle with registered custom costs = the reference's three steps; winner slot, cost bits,
    rows and admissible count equal the oracle's with the same callbacks (two callbacks: the
    float += double accumulation order matters)."""
    kw = scenario_cfg(ROBOTS[robot])
    kw.update(prediction_horizon=1.5, weights=(1.0, 1.0, 1.0, 0.5, 0.25), drop_samples=False)
    pts = wl.uturn_points()
    customs = [(0.8, _heading_cost), (2.5, _turn_cost)]
    dwa = pkg.DWA(pkg.planner_config(**kw), pkg.follower_params())
    ref = FollowerOracle(kw)
    dwa.set_current_path(pts)
    ref.set_current_path(pts)
    for w, fn in customs:
        dwa.add_custom_cost(w, fn)
    rng = np.random.default_rng(wl.SEED + 5)
    if sensor == "cloud":
        data = dict(cloud=np.concatenate([wl.round_obstacle(1.2, 0.4, 0.3), wl.round_obstacle(0.4, -0.9, 0.2)]))
    else:
        ang = np.linspace(0, 2 * np.pi, 180, endpoint=False)
        data = dict(scan=(rng.uniform(0.6, 6.0, 180), ang))
    state, vel = (0.3, 0.05, 0.2), (0.4, 0.0, 0.1)
    for step in range(3):
        dwa.set_current_state(*state)
        ref.set_current_state(*state)
        res = dwa.compute_velocity_commands(vel, **data)
        seg = ref.prepare()
        assert (dwa.info.seg_start, dwa.info.seg_count) == seg
        kwc = dict(ref.kw)
        kwc["prediction_horizon"] = ref.sampler_horizon
        kwc["num_ctrl_points"] = int(kw["control_horizon"] / kw["time_step"])
        oracle = run_oracle_cycle(kwc, ref.path, seg, vel, ref.state, customs=customs, **data)
        assert oracle["found"] and oracle["n_admissible"] > 50
        assert_cycle_parity(res, oracle)
        assert np.float32(res.cost) == np.float32(oracle["cost"])
        # the custom terms really decide: without them the winner differs
        plain = run_oracle_cycle(kwc, ref.path, seg, vel, ref.state, **data)
        assert plain["cost"] != oracle["cost"]
        cmd = (dwa.get_vx_cmd(), dwa.get_vy_cmd(), dwa.get_omega_cmd())
        state = apply_control(state, cmd, kw["time_step"])
        vel = cmd
    # clearing the callbacks returns to the fused cycle and to the plain winner
    dwa.clear_custom_costs()
    dwa.set_current_state(*state)
    ref.set_current_state(*state)
    res = dwa.compute_velocity_commands(vel, **data)
    seg = ref.prepare()
    oracle = ref.run_cycle(vel, seg, **data)
    assert_cycle_parity(res, oracle)
    dwa.close()


def test_custom_cost_exception_is_reraised(pkg):
    kw = scenario_cfg(1)
    dwa = pkg.DWA(pkg.planner_config(**kw))
    dwa.set_current_path(wl.straight_test_points())
    dwa.set_current_state(0.0, 0.0, 0.0)

    def boom(t, p):
        raise ZeroDivisionError("boom")
    dwa.add_custom_cost(1.0, boom)
    with pytest.raises(ZeroDivisionError):
        dwa.compute_velocity_commands((0, 0, 0), cloud=np.zeros((0, 3), np.float32))
    dwa.close()


@pytest.mark.parametrize("drop", [True, False])
def test_debug_velocity_search_matches_oracle_sampler(pkg, drop):
    """debugVelocitySearch = determineTarget + the sampler alone (dwa.h:147-165); the kept samples
    equal the oracle sampler's rows bit for bit, in order, for both dropping modes."""
    kw = scenario_cfg(1)
    kw.update(prediction_horizon=2.0)
    dwa = pkg.DWA(pkg.planner_config(**kw))
    with pytest.raises(ValueError, match="No debugging samples"):
        dwa.get_debugging_samples()
    with pytest.raises(ValueError, match="global path"):
        dwa.debug_velocity_search((0, 0, 0), cloud=np.zeros((0, 3), np.float32))
    dwa.set_current_path(wl.straight_test_points())
    state, vel = (0.1, -0.05, 0.1), (0.3, 0.0, 0.0)
    dwa.set_current_state(*state)
    cloud = wl.round_obstacle(1.0, 0.1, 0.3)
    dwa.debug_velocity_search(vel, cloud=cloud, drop_samples=drop)
    px, py = dwa.get_debugging_samples()
    full = dwa.get_debugging_samples(full=True)
    from parity_util import _split
    kws = dict(kw)
    kws["drop_samples"] = drop
    kws["num_ctrl_points"] = int(kw["control_horizon"] / kw["time_step"])
    common, _ = _split(kws)
    import orc
    samples = orc.sampler_generate(orc.sampler_cfg(max_num_threads=1, **common), vel, state, cloud=cloud)
    assert len(samples["slots"]) > 0 and px.shape == samples["x"].shape
    assert np.array_equal(px, samples["x"]) and np.array_equal(py, samples["y"])
    assert np.array_equal(full["slots"], samples["slots"])
    assert np.array_equal(full["vx"], samples["vx"]) and np.array_equal(full["omega"], samples["omega"])
    if not drop:  # keeping collided samples yields more rows than dropping them
        dwa.debug_velocity_search(vel, cloud=cloud, drop_samples=True)
        assert dwa.get_debugging_samples()[0].shape[0] < px.shape[0]
    dwa.close()
