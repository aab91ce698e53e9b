"""CPU oracle (test infrastructure) of the reference's Follower / DWA host logic around the planner
cycle: closest-point tracking, curvature-adaptive horizon, tracked segment, goal check. Pure Python
with numpy float32 scalars where the reference computes in float; path interpolation/segmentation
and the planner cycle itself come from the C++ oracle (orc.py).

Follows, line by line:
  src/controllers/follower.cpp:81-107 (setCurrentPath), :111-145 (isGoalReached),
  :149-176 (findClosestSegmentIndex), :194-260 (findClosestPointOnSegment), :262-304 (determineTarget)
  src/controllers/dwa.cpp:157-206 (adaptPredictionHorizonToCurvature), :208-233 (findTrackedPathSegment)
  include/controllers/dwa.h:183-230 (findBestPath), src/utils/trajectory_sampler.cpp:316-326
"""
import math

import numpy as np

import orc
from parity_util import run_oracle_cycle

F = np.float32
FLT_MAX = float(np.finfo(np.float32).max)


def dist2_state(sx, sy, px, py):
    """Path::distanceSquared(State, Point): (Point(state.x, state.y, 0) - point).squaredNorm() in float"""
    dx = F(F(sx) - F(px))
    dy = F(F(sy) - F(py))
    return F(F(dx * dx) + F(dy * dy))


class FollowerOracle:
    def __init__(self, kw, max_point_interpolation_distance=0.01, path_segment_length=1.0,
                 goal_dist_tolerance=0.1, loosing_goal_distance=0.5, curvature_horizon_tolerance=1.5,
                 ctrl_vx_max=1.0):
        self.kw = dict(kw)
        self.interp = max_point_interpolation_distance
        self.seg_len = path_segment_length
        self.goal_tol = goal_dist_tolerance
        self.loosing = loosing_goal_distance
        self.kappa_tol = curvature_horizon_tolerance
        self.v_max = ctrl_vx_max  # Controller::ctrlimitsParams default (control.h:190-193); DWA never sets it
        self.max_segment_size = int(path_segment_length / max_point_interpolation_distance + 1)
        self.base_horizon = kw["prediction_horizon"]
        self.path = None
        self.path_processing = False
        self.reached_goal = False
        self.state = (0.0, 0.0, 0.0)
        # Path::PathPosition defaults (path.h:300-308)
        self.c_index, self.c_segment, self.c_seglen = 0, 0, -1.0
        self.c_parallel, self.c_normal = 0.0, 0.0
        self.c_state = (0.0, 0.0, 0.0)
        self.current_segment_index = 0
        self.goal_distance = float("inf")

    # -- follower.cpp:81-107
    def set_current_path(self, pts):
        self.path = orc.Path(pts, self.interp, self.seg_len, self.max_segment_size)
        self.max_segment_index = len(self.path.seg_starts) - 1
        self.path_processing = True
        self.current_segment_index = 0
        self.goal_distance = float("inf")
        self.reached_goal = False

    def set_current_state(self, x, y, yaw):
        self.state = (x, y, yaw)

    def _seg_bounds(self, k):
        s, n = self.path.segment(k)
        return s, s + n - 1

    # -- follower.cpp:149-176
    def _closest_segment(self, left, right):
        if left == right:
            return left
        mid = (left + right) // 2
        p = self.path
        sl, sr = int(p.seg_starts[left]), int(p.seg_starts[right])
        dl = dist2_state(self.state[0], self.state[1], p.X[sl], p.Y[sl])
        dr = dist2_state(self.state[0], self.state[1], p.X[sr], p.Y[sr])
        if mid == right or mid == left:
            return left if dl <= dr else right
        return self._closest_segment(left, mid) if dl <= dr else self._closest_segment(mid, right)

    # -- follower.cpp:194-260
    def _closest_on_segment(self, k):
        p = self.path
        s, e = self._seg_bounds(k)
        n = e - s + 1
        min_d2 = FLT_MAX
        best, seg_pos, cx, cy = 0, 0.0, 0.0, 0.0
        heading = float(np.arctan2(F(p.Y[e] - p.Y[s]), F(p.X[e] - p.X[s])))  # atan2f
        for i in range(n):
            d2 = float(dist2_state(self.state[0], self.state[1], p.X[s + i], p.Y[s + i]))
            if d2 <= min_d2:
                min_d2 = d2
                cx, cy = float(p.X[s + i]), float(p.Y[s + i])
                best = i
                seg_pos = i / (n - 1) if n > 1 else 1.0
        self.c_index = best + s
        self.c_segment = k
        self.c_seglen = seg_pos
        self.c_state = (cx, cy, heading)
        self.c_normal = math.sqrt(min_d2)
        vx, vy = self.state[0] - cx, self.state[1] - cy
        cross = math.cos(heading) * vy - math.sin(heading) * vx
        self.c_parallel = self.c_normal if cross > 0 else -self.c_normal

    # -- follower.cpp:262-304
    def determine_target(self):
        _, seg_end = self._seg_bounds(self.current_segment_index)
        if self.c_seglen <= 0.0 or self.c_index >= seg_end or self.c_seglen >= 0.9:
            self.current_segment_index = self._closest_segment(0, self.max_segment_index)
            self._closest_on_segment(self.current_segment_index)
        else:
            self._closest_on_segment(self.c_segment)
        a = math.fmod(self.c_state[2] - self.state[2] + math.pi, 2 * math.pi)
        if a < 0:
            a += 2 * math.pi
        self.heading_error = a - math.pi

    # -- dwa.cpp:157-206 + trajectory_sampler.cpp:316-326
    def adapt_horizon(self):
        base, v_max = self.base_horizon, self.v_max
        p = self.path
        horizon = base
        if not (v_max < 1e-3) and not (self.interp <= 0.0):
            start = min(self.c_index, p.n - 1)
            peek = int(math.ceil(base * v_max / self.interp))
            end = min(start + peek, p.n - 1)
            kappa = F(0.0)
            for i in range(start, end + 1):
                kappa = max(kappa, F(abs(F(p.curv[i]))))
            if float(kappa) > self.kappa_tol:
                horizon = min(base, math.sqrt(8.0 * self.kappa_tol / float(kappa)) / v_max)
        self.max_forward_distance = horizon * v_max
        dt = self.kw["time_step"]
        clamped = min(max(horizon, 2.0 * dt), base)
        self.horizon, self.sampler_horizon = horizon, clamped
        self.n_points = int(clamped / dt)

    # -- dwa.cpp:208-233
    def tracked_segment(self):
        p = self.path
        s = min(self.c_index, p.n - 1)
        look = self.max_segment_size
        if self.interp > 0.0:
            look = max(look, int(math.ceil(self.max_forward_distance / self.interp)) + 1)
        e = min(s + look, p.n - 1)
        return s, e - s + 1

    # -- follower.cpp:111-145
    def is_goal_reached(self):
        if not self.path_processing:
            return True
        p = self.path
        dist = math.hypot(self.state[0] - float(p.X[-1]), self.state[1] - float(p.Y[-1]))
        end_reached = dist <= self.goal_tol
        loosing = False
        if self.current_segment_index + 1 >= self.max_segment_index:
            if dist < self.goal_distance:
                self.goal_distance = dist
            elif abs(dist - self.goal_distance) > self.loosing:
                loosing = True
        if end_reached or loosing:
            self.path_processing = False
            self.reached_goal = True
        return self.reached_goal

    # -- dwa.h:183-230 (findBestPath), host part only; run_cycle adds the planner
    def prepare(self):
        self.determine_target()
        self.adapt_horizon()
        return self.tracked_segment()

    def run_cycle(self, vel, seg, scan=None, cloud=None):
        kw = dict(self.kw)
        kw["prediction_horizon"] = self.sampler_horizon
        if "num_ctrl_points" not in kw:  # numCtrlPoints_ is fixed at construction (trajectory_sampler.cpp:88)
            kw["num_ctrl_points"] = int(self.kw["control_horizon"] / self.kw["time_step"])
        return run_oracle_cycle(kw, self.path, seg, vel, self.state, scan=scan, cloud=cloud)
