"""The `kompass_cpp` extension module (pybind11 in place of the absent nanobind)
with the reference's module layout, class names and keyword names for the hot-path classes. The GPU
tests restate the reference's own Python tests against it (tests/test_controllers.py::test_dwa,
tests/test_local_mapper_bindings.py, tests/test_laserscan_emergency_stop.py::test_emergency_stop) and
check that the module and the ctypes front-end (same library underneath) agree bit for bit."""
import importlib.util
import math
import os
import sys

import numpy as np
import pytest

import workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kcpp():
    spec = importlib.util.spec_from_file_location("kc_build", os.path.join(ROOT, "kompass-core_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    ext = b.build_bindings()
    d = os.path.dirname(ext)
    if d not in sys.path:
        sys.path.insert(0, d)
    import kompass_cpp
    return kompass_cpp


def test_module_layout_matches_the_reference_bindings(kcpp):
    # names a `from kompass_cpp.X import Y` in src/kompass_core/ resolves for the hot path
    for sub, names in {
        "types": ["State", "Path", "Velocity2D", "LaserScan", "Trajectory", "TrajectoryPath",
                  "TrajectoryVelocities2D", "RobotGeometry", "SensorInputType", "PointFieldType"],
        "control": ["ControlType", "LinearVelocityControlParams", "AngularVelocityControlParams",
                    "ControlLimitsParams", "TrajectoryCostWeights", "SamplingControlResult", "DWA"],
        "mapping": ["LocalMapperGPU", "OCCUPANCY_TYPE"],
        "utils": ["CriticalZoneCheckerGPU", "CollisionChecker", "pointcloud_to_laserscan_from_raw"],
    }.items():
        mod = getattr(kcpp, sub)
        for n in names:
            assert hasattr(mod, n), f"kompass_cpp.{sub}.{n}"
    for meth in ["compute_velocity_commands", "add_custom_cost", "get_debugging_samples", "debug_velocity_search",
                 "set_resolution", "set_current_path", "set_current_state", "is_goal_reached", "get_vx_cmd",
                 "get_omega_cmd", "has_path", "set_linear_ctr_limits"]:
        assert hasattr(kcpp.control.DWA, meth), meth
    assert kcpp.types.RobotGeometry.get("BOX") == kcpp.types.RobotGeometry.BOX
    assert kcpp.types.PointFieldType.from_int(7) == kcpp.types.PointFieldType.FLOAT32
    assert isinstance(kcpp.get_available_accelerators(), str)
    w = kcpp.control.TrajectoryCostWeights()
    w.from_dict({"goal_distance_weight": 3.0, "not_a_weight": 1.0})
    assert w.get_parameter("goal_distance_weight") == 3.0
    with pytest.raises(RuntimeError):  # out-of-range -> RuntimeError (bindings_config.cpp:28-32)
        w.from_dict({"goal_distance_weight": 5000.0})
    p = kcpp.types.Path(points=[(0, 0, 0), (1, 0, 0), (2, 0, 0)])
    assert p.size() == 3 and p.get_total_length() == 2.0


def test_no_cpu_fallback_behind_the_module(kcpp):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        kcpp.mapping.LocalMapperGPU(grid_height=10, grid_width=10, resolution=0.1, laserscan_position=[0, 0, 0],
                                    laserscan_orientation=0.0, is_pointcloud=False, scan_size=10,
                                    angle_step=0.1, max_height=1.0, min_height=0.0, range_max=10.0)


def _dwa(kcpp, weights=None, **over):
    c = kcpp.control
    lim = c.ControlLimitsParams(vel_x_ctr_params=c.LinearVelocityControlParams(max_vel=1.0, max_acc=5.0, max_decel=10.0),
                                vel_y_ctr_params=c.LinearVelocityControlParams(),
                                omega_ctr_params=c.AngularVelocityControlParams(max_ang=math.pi, max_omega=4.0,
                                                                                max_acc=3.0, max_decel=3.0))
    w = c.TrajectoryCostWeights()
    w.from_dict(weights or {"reference_path_distance_weight": 3.0, "goal_distance_weight": 3.0,
                            "obstacles_distance_weight": 1.0, "smoothness_weight": 0.0, "jerk_weight": 0.0})
    kw = dict(control_limits=lim, control_type=c.ControlType.DIFFERENTIAL_DRIVE, time_step=0.1,
              prediction_horizon=1.0, control_horizon=0.2, max_linear_samples=20, max_angular_samples=20,
              robot_shape_type=kcpp.types.RobotGeometry.CYLINDER, robot_dimensions=[0.1, 0.4],
              sensor_position_robot=[0.0, 0.0, 0.0], sensor_rotation_robot=[0.0, 0.0, 0.0, 1.0],
              octree_resolution=0.1, cost_weights=w, max_num_threads=1)
    kw.update(over)
    return c.DWA(**kw)


@pytest.mark.gpu
def test_dwa_closed_loop_like_the_reference_python_test(kcpp, pkg):
    """tests/test_controllers.py::test_dwa: follow the global path with the default Python weights;
    every step must find a command, the goal must be reached; the first cycle equals the ctypes
    front-end's on the same inputs."""
    planner = _dwa(kcpp)
    pts = [(float(x), float(y), 0.0) for x, y in wl.GLOBAL_PATH_XY]
    planner.set_current_path(kcpp.types.Path(points=pts))
    ranges, angles = wl.scan_360()
    scan = kcpp.types.LaserScan(ranges=list(ranges), angles=list(angles))
    state = [-0.51731912, 0.0, 0.0]
    planner.set_current_state(kcpp.types.State(x=state[0], y=state[1], yaw=state[2], speed=0.0))
    res = planner.compute_velocity_commands(kcpp.types.Velocity2D(), scan)
    assert res.is_found and len(res.trajectory.path.x) == 10 and len(res.trajectory.velocities.vx) == 9

    ref = pkg.DWA(pkg.planner_config(**wl.cfg_c1()))
    ref.set_current_path(np.asarray(wl.GLOBAL_PATH_XY, np.float32))
    ref.set_current_state(*state)
    r2 = ref.compute_velocity_commands((0, 0, 0), scan=(ranges, angles))
    assert np.float32(res.cost) == np.float32(r2.cost)
    assert np.array_equal(np.asarray(res.trajectory.path.x, np.float32), r2.x)
    assert np.array_equal(np.asarray(res.trajectory.velocities.omega, np.float32), r2.omega)
    ref.close()

    free = kcpp.types.LaserScan(ranges=[10.0, 10.1], angles=[0.4, 0.3])
    vel, steps = kcpp.types.Velocity2D(), 0
    while not planner.is_goal_reached() and steps < 600:
        planner.set_current_state(state[0], state[1], state[2], 0.0)
        res = planner.compute_velocity_commands(vel, free)
        assert res.is_found, f"no command at step {steps}"
        vx, vy, om = planner.get_vx_cmd(), planner.get_vy_cmd(), planner.get_omega_cmd()
        state[0] += (vx * math.cos(state[2]) - vy * math.sin(state[2])) * 0.1
        state[1] += (vx * math.sin(state[2]) + vy * math.cos(state[2])) * 0.1
        state[2] += om * 0.1
        vel = kcpp.types.Velocity2D(vx=vx, vy=vy, omega=om)
        steps += 1
    assert planner.is_goal_reached() and steps < 600


@pytest.mark.gpu
def test_dwa_custom_cost_and_debug_samples_through_the_module(kcpp):
    planner = _dwa(kcpp)
    planner.set_current_path(kcpp.types.Path(points=[(0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (2.0, 0.0, 0.0)]))
    planner.set_current_state(-0.5, 0.0, 0.0, 0.0)
    scan = kcpp.types.LaserScan(ranges=[5.0] * 90, angles=list(np.linspace(0, 2 * math.pi, 90, endpoint=False)))
    vel = kcpp.types.Velocity2D(vx=0.2)
    with pytest.raises(ValueError):
        planner.get_debugging_samples()
    planner.debug_velocity_search(vel, scan, True)
    px, py_ = planner.get_debugging_samples()
    assert px.shape == py_.shape and px.shape[1] == 10 and px.shape[0] > 100 and np.all(px[:, 0] == np.float32(-0.5))
    plain = planner.compute_velocity_commands(vel, scan)
    seen = []

    def turn_left(traj, path):
        seen.append(path.size())
        return 10.0 - traj.velocities.omega[0]
    planner.add_custom_cost(100.0, turn_left)
    steered = planner.compute_velocity_commands(vel, scan)
    assert steered.is_found and len(seen) == px.shape[0] and seen[0] == 201
    assert steered.trajectory.velocities.omega[0] > plain.trajectory.velocities.omega[0]


@pytest.mark.gpu
def test_local_mapper_like_the_reference_binding_test(kcpp, pkg):
    """tests/test_local_mapper_bindings.py: ring scans -> values in {-1, 0, 100}, counts partition the
    grid, occupied and empty cells exist; equal to the ctypes front-end cell for cell."""
    args = dict(grid_height=100, grid_width=120, resolution=0.1, laserscan_position=[0.0, 0.0, 0.0],
                laserscan_orientation=0.0, is_pointcloud=False, scan_size=360, angle_step=0.01, max_height=2.0,
                min_height=0.0, range_max=20.0, max_points_per_line=256)
    mapper = kcpp.mapping.LocalMapperGPU(**args)
    angles = np.linspace(0, 2 * math.pi, 360, endpoint=False)
    ranges = np.full(360, 3.0)
    grid = mapper.scan_to_grid(angles=list(angles), ranges=list(ranges))
    assert grid.shape == (100, 120) and grid.dtype == np.int32 and grid.flags.f_contiguous
    vals, counts = np.unique(grid, return_counts=True)
    assert set(vals.tolist()) <= {-1, 0, 100} and counts.sum() == 100 * 120
    assert (grid == 100).sum() > 0 and (grid == 0).sum() > 0
    ref = pkg.LocalMapperGPU(100, 120, 0.1, (0, 0, 0), 0.0, False, 360, 0.01, 2.0, 0.0, 20.0, 256)
    assert np.array_equal(grid, ref.scan_to_grid(angles, ranges))
    g2, p2 = mapper.scan_to_grid_baysian(angles=list(angles), ranges=list(ranges))
    gr, pr = ref.scan_to_grid_baysian(angles, ranges)
    assert np.array_equal(g2, gr) and np.array_equal(p2.view(np.uint32), pr.view(np.uint32))
    ref.close()
    # raw point cloud overload
    pts = wl.cloud_lattice(4, 20_000)
    data = wl.cloud_bytes_xyz16(pts)
    cargs = dict(args, is_pointcloud=True, scan_size=720)
    cm = kcpp.mapping.LocalMapperGPU(**cargs)
    g = cm.scan_to_grid(data=data, point_step=16, row_step=len(pts) * 16, height=1, width=len(pts), x_offset=0,
                        y_offset=4, z_offset=8)
    rc = pkg.LocalMapperGPU(100, 120, 0.1, (0, 0, 0), 0.0, True, 720, 0.01, 2.0, 0.0, 20.0, 256)
    assert np.array_equal(g, rc.scan_to_grid(data, 16, len(pts) * 16, 1, len(pts), 0, 4, 8))
    rc.close()


@pytest.mark.gpu
def test_emergency_stop_like_the_reference_python_test(kcpp):
    """tests/test_laserscan_emergency_stop.py::test_emergency_stop (use_gpu=True branch): the wrapper's
    constructor keywords (src/kompass_core/utils/emergency_stop.py:70-106), Python's default sensor
    rotation [1, 0, 0, 0], factor 1.0 / 0.0 / 1.0."""
    robot_radius, emergency_distance, slowdown_distance = 0.1, 0.5, 1.0
    angles = np.arange(0.0, 2 * math.pi, 0.1)
    checker = kcpp.utils.CriticalZoneCheckerGPU(
        input_type=kcpp.types.SensorInputType.LASERSCAN, scan_angles=list(angles),
        robot_shape=kcpp.types.RobotGeometry.CYLINDER, robot_dimensions=[robot_radius, 0.4],
        sensor_position_body=[0.0, 0.0, 0.173], sensor_rotation_body=[1.0, 0.0, 0.0, 0.0], critical_angle=90.0,
        critical_distance=emergency_distance, slowdown_distance=slowdown_distance, min_height=-0.4,
        max_height=0.4, range_max=20.0, cloud_field_type=kcpp.types.PointFieldType.FLOAT32)
    ranges = np.full(len(angles), 10.0)
    assert checker.check(ranges=list(ranges), forward=True) == 1.0
    ranges[0] = robot_radius + emergency_distance / 2
    assert checker.check(ranges=list(ranges), forward=True) == 0.0
    assert checker.check(ranges=list(ranges), forward=False) == 1.0


@pytest.mark.gpu
def test_collision_checker_through_the_module(kcpp):
    cc = kcpp.utils.CollisionChecker(robot_shape=kcpp.types.RobotGeometry.BOX, robot_dimensions=[0.4, 0.4, 1.0],
                                     sensor_position_body=[0.0, 0.0, 1.0], sensor_rotation_body=[0.0, 0.0, 1.0, 0.0],
                                     octree_resolution=0.1)
    cc.update_state(3.0, 5.0, 0.0)
    cc.update_sensor_data([(3.1, 5.1, -0.5)], True)
    assert cc.check_collisions() is True
    any_hit, per = cc.check_states([kcpp.types.State(0, 0, 0), kcpp.types.State(3.0, 5.0, 0.3)])
    assert any_hit and list(per) == [0, 1]
