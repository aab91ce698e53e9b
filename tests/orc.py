"""ctypes bindings of the CPU parity oracle (oracle/_build/libkompass_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libkompass_oracle.so")

ACKERMANN, DIFFERENTIAL_DRIVE, OMNI = 0, 1, 2
CYLINDER, BOX, SPHERE = 0, 1, 2


class SamplerCfg(C.Structure):
    _fields_ = [
        ("control_type", C.c_int32),
        ("time_step", C.c_double),
        ("prediction_horizon", C.c_double),
        ("control_horizon", C.c_double),
        ("max_linear_samples", C.c_int32),
        ("max_angular_samples", C.c_int32),
        ("vx_max", C.c_double), ("vx_acc", C.c_double), ("vx_dec", C.c_double),
        ("vy_max", C.c_double), ("vy_acc", C.c_double), ("vy_dec", C.c_double),
        ("omega_max", C.c_double), ("omega_acc", C.c_double), ("omega_dec", C.c_double),
        ("robot_shape", C.c_int32),
        ("robot_dims", C.c_float * 3),
        ("sensor_position", C.c_float * 3),
        ("sensor_rotation", C.c_float * 4),
        ("octree_resolution", C.c_double),
        ("drop_samples", C.c_int32),
        ("num_ctrl_points", C.c_int64),
        ("max_num_threads", C.c_int32),
    ]


class CostCfg(C.Structure):
    _fields_ = [
        ("w_path", C.c_double), ("w_goal", C.c_double), ("w_obstacles", C.c_double),
        ("w_smooth", C.c_double), ("w_jerk", C.c_double),
        ("acc_limits", C.c_float * 3),
        ("sensor_position", C.c_float * 3),
        ("sensor_rotation", C.c_float * 4),
    ]


class CzCfg(C.Structure):
    _fields_ = [
        ("robot_shape", C.c_int32),
        ("robot_dims", C.c_float * 3),
        ("sensor_position", C.c_float * 3),
        ("sensor_rotation", C.c_float * 4),
        ("critical_angle", C.c_float),
        ("critical_distance", C.c_float),
        ("slowdown_distance", C.c_float),
        ("min_height", C.c_float), ("max_height", C.c_float), ("range_max", C.c_float),
    ]


REF_LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libkompass_ref.so")

_lib = None
_libs = {}
_backend = "port"


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def ref_available():
    """oracle/_ref/libkompass_ref.so: the reference's OWN sources compiled against the stand-in headers
    of oracle/shim (built in the authoring container by `make -C oracle ref`; travels prebuilt)."""
    return os.path.exists(REF_LIB_PATH)


def set_backend(name):
    """'port' (oracle/kompass_oracle.cpp, the default) or 'ref' (the reference's own sources): every
    wrapper below that exists in both libraries then calls the chosen one."""
    global _backend, _lib
    assert name in ("port", "ref")
    _backend = name
    _lib = None


def _load(path):
    L = C.CDLL(path)
    L.orc_num_trajectories.restype = C.c_int64
    L.orc_num_trajectories.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.orc_num_points.restype = C.c_int64
    L.orc_num_points.argtypes = [C.c_double, C.c_double]
    L.orc_cz_check_scan.restype = C.c_float
    L.orc_cz_check_cloud.restype = C.c_float
    return L


def lib():
    global _lib
    if _lib is None:
        if _backend not in _libs:
            if _backend == "port":
                if not os.path.exists(LIB_PATH):
                    build()
                _libs["port"] = _load(LIB_PATH)
                _libs["port"].orc_segment_length.restype = C.c_float
            else:
                _libs["ref"] = _load(REF_LIB_PATH)
                _libs["ref"].orc_ref_segment_length.restype = C.c_float
        _lib = _libs[_backend]
    return _lib


# ---- entry points that exist only in the _ref library (they need the Path OBJECT, so they take the
# original way points instead of the interpolated arrays) -------------------------------------------
def ref_path_segment(pts, interp, seg_len, max_pts_per_seg=10000):
    assert _backend == "ref"
    pts = np.asarray(pts, dtype=np.float32)
    x, y = f32(pts[:, 0]), f32(pts[:, 1])
    starts = np.zeros(1 << 16, np.int32)
    ns = lib().orc_ref_path_segment(fp(x), fp(y), len(x), C.c_double(interp), C.c_double(seg_len),
                                    C.c_int64(max_pts_per_seg), ip(starts), len(starts))
    return starts[:ns].copy()


def ref_segment_length(pts, interp, start, count):
    assert _backend == "ref"
    pts = np.asarray(pts, dtype=np.float32)
    x, y = f32(pts[:, 0]), f32(pts[:, 1])
    return float(np.float32(lib().orc_ref_segment_length(fp(x), fp(y), len(x), C.c_double(interp), start, count)))


def ref_cost_evaluate(ccfg, samples, way_points, interp, seg, pose, max_sensor_range, scan=None, cloud=None,
                      want_costs=True):
    """CostEvaluator::setPointScan + getMinTrajectoryCost of the reference class -> (found, best_idx,
    best_cost, per-trajectory costs)."""
    assert _backend == "ref"
    x, y = f32(samples["x"]), f32(samples["y"])
    vx, vy, om = f32(samples["vx"]), f32(samples["vy"]), f32(samples["omega"])
    n, P = x.shape
    wp = np.asarray(way_points, dtype=np.float32)
    wx, wy = f32(wp[:, 0]), f32(wp[:, 1])
    costs = np.zeros(n, np.float32)
    bi, bc = C.c_int32(-1), C.c_float(0)
    p = f64(pose)
    if scan is not None:
        a, b = f64(scan[0]), f64(scan[1])
        is_cloud, pa, pb, n_obs = 0, dp(a), dp(b), len(a)
    elif cloud is not None:
        a = f32(cloud).reshape(-1, 3)
        is_cloud, pa, pb, n_obs = 1, fp(a), None, len(a)
    else:
        is_cloud, pa, pb, n_obs = 1, None, None, 0
    found = lib().orc_ref_cost_evaluate(C.byref(ccfg), n, P, fp(vx), fp(vy), fp(om), fp(x), fp(y), fp(wx), fp(wy),
                                        len(wx), C.c_double(interp), seg[0], seg[1], is_cloud, pa, pb, n_obs,
                                        dp(p), C.c_float(max_sensor_range), fp(costs) if want_costs else None,
                                        C.byref(bi), C.byref(bc))
    return bool(found), bi.value, float(np.float32(bc.value)), costs


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def fp(a):
    return _p(a, C.c_float)


def dp(a):
    return _p(a, C.c_double)


def ip(a):
    return _p(a, C.c_int32)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------------------------------------
class Path:
    """Interpolated + segmented reference path (ref: Path::interpolate / Path::segment)."""

    def __init__(self, pts, interp, seg_len, max_pts_per_seg=10000):
        pts = np.asarray(pts, dtype=np.float32)
        x, y = f32(pts[:, 0]), f32(pts[:, 1])
        cap = 1 << 20
        X, Y = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        acc, curv = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        tot = C.c_float(0)
        n = lib().orc_path_interpolate_linear(fp(x), fp(y), len(x), C.c_double(interp), fp(X), fp(Y),
                                              fp(acc), fp(curv), cap, C.byref(tot))
        assert n > 0, n
        self.X, self.Y, self.acc, self.curv = X[:n].copy(), Y[:n].copy(), acc[:n].copy(), curv[:n].copy()
        self.n = n
        self.total_length = float(np.float32(tot.value))
        self.way_points, self.interp = pts.copy(), interp
        if _backend == "ref":
            self.seg_starts = ref_path_segment(pts, interp, seg_len, max_pts_per_seg)
            return
        starts = np.zeros(n + 1, np.int32)
        ns = lib().orc_path_segment(fp(self.acc), n, C.c_double(seg_len), C.c_int64(max_pts_per_seg),
                                    ip(starts), n + 1)
        self.seg_starts = starts[:ns].copy()

    def segment(self, i):
        s = int(self.seg_starts[i])
        e = int(self.seg_starts[i + 1]) - 1 if i + 1 < len(self.seg_starts) else self.n - 1
        return s, e - s + 1

    def part(self, start, end):
        return start, end - start + 1


def sampler_cfg(control_type=DIFFERENTIAL_DRIVE, time_step=0.1, prediction_horizon=1.0,
                control_horizon=0.2, max_linear_samples=20, max_angular_samples=20,
                vx=(1.0, 5.0, 10.0), vy=(0.0, 0.0, 0.0), omega=(4.0, 3.0, 3.0),
                shape=CYLINDER, dims=(0.1, 0.4, 0.0), sensor_position=(0, 0, 0),
                sensor_rotation=(0, 0, 0, 1), octree_resolution=0.1, drop_samples=True,
                num_ctrl_points=None, max_num_threads=1):
    c = SamplerCfg()
    c.control_type = control_type
    c.time_step, c.prediction_horizon, c.control_horizon = time_step, prediction_horizon, control_horizon
    c.max_linear_samples, c.max_angular_samples = max_linear_samples, max_angular_samples
    c.vx_max, c.vx_acc, c.vx_dec = vx
    c.vy_max, c.vy_acc, c.vy_dec = vy
    c.omega_max, c.omega_acc, c.omega_dec = omega
    c.robot_shape = shape
    d = list(dims) + [0.0] * (3 - len(dims))
    c.robot_dims = (C.c_float * 3)(*d)
    c.sensor_position = (C.c_float * 3)(*sensor_position)
    c.sensor_rotation = (C.c_float * 4)(*sensor_rotation)
    c.octree_resolution = octree_resolution
    c.drop_samples = 1 if drop_samples else 0
    c.num_ctrl_points = int(control_horizon / time_step) if num_ctrl_points is None else num_ctrl_points
    c.max_num_threads = max_num_threads
    return c


def cost_cfg(w_path=1.0, w_goal=1.0, w_obstacles=1.0, w_smooth=1.0, w_jerk=1.0,
             acc_limits=(1.0, 1.0, 1.0), sensor_position=(0, 0, 0), sensor_rotation=(0, 0, 0, 1)):
    c = CostCfg()
    c.w_path, c.w_goal, c.w_obstacles, c.w_smooth, c.w_jerk = w_path, w_goal, w_obstacles, w_smooth, w_jerk
    c.acc_limits = (C.c_float * 3)(*acc_limits)
    c.sensor_position = (C.c_float * 3)(*sensor_position)
    c.sensor_rotation = (C.c_float * 4)(*sensor_rotation)
    return c


def num_points(cfg):
    return int(lib().orc_num_points(C.c_double(cfg.time_step), C.c_double(cfg.prediction_horizon)))


def num_trajectories(cfg):
    return int(lib().orc_num_trajectories(cfg.control_type, cfg.max_linear_samples, cfg.max_angular_samples))


def velocity_samples(cfg, vel):
    cap = num_trajectories(cfg) + 8
    vx, vy, om = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    v = f64(vel)
    n = lib().orc_velocity_samples(C.byref(cfg), dp(v), dp(vx), dp(vy), dp(om), cap)
    assert n >= 0, n
    return vx[:n], vy[:n], om[:n]


def sampler_generate(cfg, vel, pose, scan=None, cloud=None):
    """Returns dict(vx, vy, omega [n x P-1], x, y [n x P], slots [n])."""
    P = num_points(cfg)
    cap = num_trajectories(cfg) + 8
    vx = np.zeros((cap, P - 1), np.float32)
    vy = np.zeros_like(vx)
    om = np.zeros_like(vx)
    x = np.zeros((cap, P), np.float32)
    y = np.zeros_like(x)
    slots = np.zeros(cap, np.int32)
    v, p = f64(vel), f64(pose)
    if scan is not None:
        r, a = f64(scan[0]), f64(scan[1])
        n = lib().orc_sampler_generate_scan(C.byref(cfg), dp(v), dp(p), dp(r), dp(a), len(r), fp(vx), fp(vy),
                                            fp(om), fp(x), fp(y), ip(slots), cap)
    else:
        pts = f32(cloud).reshape(-1, 3)
        n = lib().orc_sampler_generate_cloud(C.byref(cfg), dp(v), dp(p), fp(pts), len(pts), fp(vx), fp(vy),
                                             fp(om), fp(x), fp(y), ip(slots), cap)
    assert n >= 0, n
    return dict(vx=vx[:n], vy=vy[:n], omega=om[:n], x=x[:n], y=y[:n], slots=slots[:n], P=P)


def check_collision(cfg, sensor_pose, query_pose, scan=None, cloud=None):
    sp, qp = f64(sensor_pose), f64(query_pose)
    if scan is not None:
        r, a = f64(scan[0]), f64(scan[1])
        return lib().orc_check_collision(C.byref(cfg), dp(sp), dp(qp), 0, dp(r), dp(a), len(r))
    pts = f32(cloud).reshape(-1, 3)
    return lib().orc_check_collision(C.byref(cfg), dp(sp), dp(qp), 1, fp(pts), None, len(pts))


def check_collision_states(cfg, sensor_pose, states, scan=None, cloud=None, global_frame=True):
    """-> (any, per-state uint8 array)"""
    sp = f64(sensor_pose)
    st = f64(states).reshape(-1, 3)
    out = np.zeros(len(st), np.uint8)
    op = out.ctypes.data_as(C.POINTER(C.c_uint8))
    if scan is not None:
        r, a = f64(scan[0]), f64(scan[1])
        rc = lib().orc_check_collision_states(C.byref(cfg), dp(sp), 0, 0, dp(r), dp(a), len(r), dp(st),
                                              len(st), op)
    else:
        pts = f32(cloud).reshape(-1, 3)
        rc = lib().orc_check_collision_states(C.byref(cfg), dp(sp), 1, 1 if global_frame else 0, fp(pts),
                                              None, len(pts), dp(st), len(st), op)
    assert rc >= 0, rc
    return bool(rc), out


def cost_points(ccfg, pose, scan=None, cloud=None):
    p = f64(pose)
    if scan is not None:
        r, a = f64(scan[0]), f64(scan[1])
        ox, oy = np.zeros(len(r), np.float32), np.zeros(len(r), np.float32)
        lib().orc_cost_points_scan(C.byref(ccfg), dp(r), dp(a), len(r), dp(p), fp(ox), fp(oy))
    else:
        pts = f32(cloud).reshape(-1, 3)
        ox, oy = np.zeros(len(pts), np.float32), np.zeros(len(pts), np.float32)
        lib().orc_cost_points_cloud(C.byref(ccfg), fp(pts), len(pts), dp(p), fp(ox), fp(oy))
    return ox, oy


def cost_evaluate(ccfg, samples, path, seg, obstacles=None, max_obstacles_dist=0.0, custom=None,
                  n_threads=1):
    """samples: dict with vx,vy,omega,x,y row-major float32; seg = (start,count).
    Returns (found, best_idx, best_cost, costs)."""
    x, y = f32(samples["x"]), f32(samples["y"])
    vx, vy, om = f32(samples["vx"]), f32(samples["vy"]), f32(samples["omega"])
    n, P = x.shape
    costs = np.zeros(n, np.float32)
    bi, bc = C.c_int32(-1), C.c_float(0)
    if obstacles is None:
        ox = oy = np.zeros(0, np.float32)
    else:
        ox, oy = f32(obstacles[0]), f32(obstacles[1])
    cu, ncu = None, 0
    if custom is not None:  # [n, n_custom] weighted callback terms (doubles)
        cua = f64(custom).reshape(n, -1)
        cu, ncu = dp(cua), cua.shape[1]
    found = lib().orc_cost_evaluate(C.byref(ccfg), n, P, fp(vx), fp(vy), fp(om), fp(x), fp(y), fp(path.X),
                                    fp(path.Y), fp(path.acc), path.n, C.c_float(path.total_length),
                                    seg[0], seg[1], fp(ox), fp(oy), len(ox),
                                    C.c_float(max_obstacles_dist), cu, ncu, fp(costs), C.byref(bi),
                                    C.byref(bc), n_threads)
    return bool(found), bi.value, float(np.float32(bc.value)), costs


def mapper_scan_to_grid(H, W, res, laser_pos, laser_orient, angles, ranges):
    a, r = f64(angles), f64(ranges)
    grid = np.zeros((W, H), np.int32)  # column-major [H x W] == C-order [W][H]
    lp = f32(laser_pos)
    lib().orc_mapper_scan_to_grid(H, W, C.c_float(res), fp(lp), C.c_float(laser_orient), dp(a), dp(r),
                                  len(a), ip(grid))
    return grid.T  # grid[i, j]


def mapper_scan_to_grid_bayes(H, W, res, laser_pos, laser_orient, angles, ranges, prev=None, p_prior=0.5,
                              p_occupied=0.6, p_empty=0.4, range_sure=1.0, range_max=20.0, wall_size=0.2):
    """ref: LocalMapper::scanToGridBaysian. prev: [H, W] previous probabilities (default: prior).
    Returns (grid[i, j], prob[i, j])."""
    a, r = f64(angles), f64(ranges)
    grid = np.zeros((W, H), np.int32)
    prob = np.zeros((W, H), np.float32)
    pv = np.full((W, H), p_prior, np.float32) if prev is None else np.ascontiguousarray(np.asarray(prev, np.float32).T)
    lp = f32(laser_pos)
    lib().orc_mapper_scan_to_grid_bayes(H, W, C.c_float(res), fp(lp), C.c_float(laser_orient),
                                        C.c_float(p_prior), C.c_float(p_occupied), C.c_float(p_empty),
                                        C.c_float(range_sure), C.c_float(range_max), C.c_float(wall_size),
                                        dp(a), dp(r), len(a), fp(pv), ip(grid), fp(prob))
    return grid.T, prob.T


def mapper_warp_previous(H, W, res, p_prior, pos, orientation, prev):
    """ref: LocalMapper::getPreviousGridInCurrentPose. prev: [H, W]; returns the warped [H, W]."""
    pv = np.ascontiguousarray(np.asarray(prev, np.float32).T)
    out = np.zeros((W, H), np.float32)
    lib().orc_mapper_warp_previous(H, W, C.c_float(res), C.c_float(p_prior), C.c_float(pos[0]),
                                   C.c_float(pos[1]), C.c_double(orientation), fp(pv), fp(out))
    return out.T


def pointcloud_to_laserscan(data, point_step, row_step, height, width, xo, yo, zo, max_range, min_z,
                            max_z, num_bins):
    d = np.ascontiguousarray(data, dtype=np.int8)
    out = np.zeros(num_bins, np.float64)
    lib().orc_pointcloud_to_laserscan(_p(d, C.c_int8), C.c_int64(d.size), point_step, row_step, height,
                                      width, xo, yo, zo, C.c_double(max_range), C.c_double(min_z),
                                      C.c_double(max_z), num_bins, dp(out))
    return out


def pointcloud_to_laserscan_step(data, point_step, row_step, height, width, xo, yo, zo, max_range, min_z,
                                 max_z, angle_step):
    """ref: pointcloud.h:116-177 -> (ranges, angles)"""
    d = np.ascontiguousarray(data, dtype=np.int8)
    cap = int(np.ceil(2.0 * np.pi / angle_step)) + 2
    ranges, angles = np.zeros(cap, np.float64), np.zeros(cap, np.float64)
    n = lib().orc_pointcloud_to_laserscan_step(_p(d, C.c_int8), C.c_int64(d.size), point_step, row_step,
                                               height, width, xo, yo, zo, C.c_double(max_range),
                                               C.c_double(min_z), C.c_double(max_z),
                                               C.c_double(angle_step), dp(ranges), dp(angles))
    return ranges[:n].copy(), angles[:n].copy()


def cz_cfg(shape=CYLINDER, dims=(0.51, 2.0, 0.0), sensor_position=(0.22, 0.0, 0.4),
           sensor_rotation=(0, 0, 0.99, 0.0), critical_angle=160.0, critical_distance=0.3,
           slowdown_distance=0.6, min_height=0.1, max_height=2.0, range_max=20.0):
    c = CzCfg()
    c.robot_shape = shape
    d = list(dims) + [0.0] * (3 - len(dims))
    c.robot_dims = (C.c_float * 3)(*d)
    c.sensor_position = (C.c_float * 3)(*sensor_position)
    c.sensor_rotation = (C.c_float * 4)(*sensor_rotation)
    c.critical_angle, c.critical_distance, c.slowdown_distance = critical_angle, critical_distance, slowdown_distance
    c.min_height, c.max_height, c.range_max = min_height, max_height, range_max
    return c


def cz_check_scan(cfg, angles, ranges, forward):
    a, r = f64(angles), f64(ranges)
    return float(lib().orc_cz_check_scan(C.byref(cfg), dp(a), len(a), dp(r), 1 if forward else 0))


def cz_check_cloud(cfg, angles, data, point_step, row_step, height, width, xo, yo, zo, forward):
    a = f64(angles)
    d = np.ascontiguousarray(data, dtype=np.int8)
    return float(lib().orc_cz_check_cloud(C.byref(cfg), dp(a), len(a), _p(d, C.c_int8), C.c_int64(d.size),
                                          point_step, row_step, height, width, xo, yo, zo,
                                          1 if forward else 0))


def cz_indices(cfg, angles, forward):
    a = f64(angles)
    out = np.zeros(len(a), np.int32)
    n = lib().orc_cz_indices(C.byref(cfg), dp(a), len(a), 1 if forward else 0, ip(out))
    return out[:n]


# ---- the reference's own DWA controller object (oracle/_ref only) ---------------------------------
class RefDwaInfo(C.Structure):
    _fields_ = [("closest_index", C.c_int32), ("segment_index", C.c_int32), ("seg_start", C.c_int32),
                ("seg_count", C.c_int32), ("n_points", C.c_int32), ("found", C.c_int32), ("cost", C.c_float),
                ("_pad", C.c_float), ("segment_position", C.c_double), ("crosstrack_error", C.c_double),
                ("heading_error", C.c_double), ("horizon", C.c_double), ("cmd", C.c_double * 3)]


class RefDWA:
    """Kompass::Control::DWA compiled from the reference's sources (oracle/_ref): setCurrentPath,
    setCurrentState, computeVelocityCommandsSet(vel, cloud), isGoalReached."""

    def __init__(self, scfg, ccfg, interp=0.01, seg_len=1.0, goal_tol=0.1, loosing=0.5, kappa_tol=1.5,
                 max_local_range=10.0):
        assert ref_available()
        L = _libs.get("ref") or _load(REF_LIB_PATH)
        _libs["ref"] = L
        L.orc_ref_segment_length.restype = C.c_float
        L.orc_ref_dwa_create.restype = C.c_void_p
        self.L = L
        self.h = C.c_void_p(L.orc_ref_dwa_create(C.byref(scfg), C.byref(ccfg), C.c_double(interp), C.c_double(seg_len),
                                                 C.c_double(goal_tol), C.c_double(loosing), C.c_double(kappa_tol),
                                                 C.c_float(max_local_range)))

    def close(self):
        if self.h:
            self.L.orc_ref_dwa_destroy(self.h)
            self.h = None

    def set_current_path(self, pts):
        pts = np.asarray(pts, dtype=np.float32)
        x, y = f32(pts[:, 0]), f32(pts[:, 1])
        return self.L.orc_ref_dwa_set_path(self.h, fp(x), fp(y), len(x))

    def set_current_state(self, x, y, yaw):
        self.L.orc_ref_dwa_set_state(self.h, C.c_double(x), C.c_double(y), C.c_double(yaw))

    def is_goal_reached(self):
        return bool(self.L.orc_ref_dwa_goal_reached(self.h))

    def compute(self, vel, cloud):
        v = f64(vel)
        pts = f32(cloud).reshape(-1, 3)
        info = RefDwaInfo()
        rows = np.zeros(5 * 4096, np.float32)
        P = self.L.orc_ref_dwa_compute_cloud(self.h, dp(v), fp(pts), len(pts), C.byref(info), fp(rows), len(rows))
        assert P >= 0
        out = dict(info=info, P=P)
        if P > 0:
            out.update(vx=rows[:P - 1].copy(), vy=rows[P - 1:2 * (P - 1)].copy(), omega=rows[2 * (P - 1):3 * (P - 1)].copy(),
                       x=rows[3 * (P - 1):3 * (P - 1) + P].copy(), y=rows[3 * (P - 1) + P:3 * (P - 1) + 2 * P].copy())
        return out


def ref_cycle_cloud(scfg, ccfg, way_points, interp, seg, vel, pose, cloud, max_sensor_range, max_traj=0):
    """One cycle of the reference's own classes (oracle/_ref), timed inside: -> dict n_admissible, P,
    evaluated, found, cost, t_sampler, t_points, t_cost (seconds)."""
    assert ref_available()
    L = _libs.get("ref") or _load(REF_LIB_PATH)
    _libs["ref"] = L
    wp = np.asarray(way_points, dtype=np.float32)
    wx, wy = f32(wp[:, 0]), f32(wp[:, 1])
    v, p = f64(vel), f64(pose)
    pts = f32(cloud).reshape(-1, 3)
    times = np.zeros(3, np.float64)
    P, ev, fnd, bc = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_float(0)
    n = L.orc_ref_cycle_cloud(C.byref(scfg), C.byref(ccfg), fp(wx), fp(wy), len(wx), C.c_double(interp), seg[0], seg[1],
                              dp(v), dp(p), fp(pts), len(pts), C.c_float(max_sensor_range), int(max_traj), dp(times),
                              C.byref(P), C.byref(ev), C.byref(fnd), C.byref(bc))
    return dict(n_admissible=n, P=P.value, evaluated=ev.value, found=bool(fnd.value), cost=float(np.float32(bc.value)),
                t_sampler=float(times[0]), t_points=float(times[1]), t_cost=float(times[2]))
