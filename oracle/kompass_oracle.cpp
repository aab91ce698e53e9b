/*
 * kompass_oracle.cpp — CPU parity ORACLE (test infrastructure; never linked into the product).
 *
 * Line-faithful restatement of the reference CPU path (kompass-core 0.8.1). Each function cites
 * the reference file:line it follows ("ref:" paths are relative to
 * /root/reference/src/kompass_cpp/kompass_cpp/). Build with -ffp-contract=off: the reference
 * release build targets baseline x86-64 (no FMA ISA), so every float/double op rounds separately.
 *
 * Mixed-precision rules that were checked against libstdc++ overload resolution
 * (see DESIGN.md §Numerics): unqualified cos/sin/sqrt/pow on a float resolve to the double
 * ::cos/::sin/::sqrt/::pow; std::sqrt/std::atan2 on floats stay float; std::pow(float,int)
 * promotes to double.
 *
 * Third-party arithmetic that is NOT in /root/reference and is restated from its published
 * algorithm: Eigen 3.4 (eigen_order.h), tk::spline linear mode (vendored, restated here),
 * FCL 0.7.0 + octomap (analytic voxel-cube model below; "parity unpinned" beyond the three
 * booleans of tests/collisions_test.cpp).
 */
#include "kompass_oracle.h"
#include "eigen_order.h"
#include "voxel_model.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {
using vox::CollisionWorld;
using vox::initWorld;
using vox::insertPoint;
using vox::poseCollides;

constexpr double MIN_VEL = 0.01; /* ref: include/utils/trajectory_sampler.h:13-15 */
constexpr float DEFAULT_MIN_DIST = std::numeric_limits<float>::max(); /* ref: trajectory.h:12 */

/* ------------------------------------------------------------------------------------------
 * sizes — ref: include/datatypes/trajectory.h:19-51
 * ---------------------------------------------------------------------------------------- */
void computeLinearSampleSplit(int ctrType, int maxLinearSamples, int &vx_n, int &vy_n) {
  auto makeOdd = [](int n) { return (n % 2 == 0) ? n + 1 : n; };
  if (ctrType == ORC_OMNI) {
    vx_n = makeOdd(std::max(3, maxLinearSamples * 3 / 4));
    vy_n = makeOdd(std::max(3, maxLinearSamples * 1 / 4));
  } else {
    vx_n = makeOdd(std::max(3, maxLinearSamples));
    vy_n = 1;
  }
}

/* ------------------------------------------------------------------------------------------
 * velocity window — ref: src/utils/trajectory_sampler.cpp:328-372
 * ---------------------------------------------------------------------------------------- */
struct Window {
  double min_vx, max_vx, min_vy, max_vy, min_om, max_om;
  double res_x, res_y, res_om;
};

Window reachableWindow(const orc_sampler_cfg &c, const double vel[3]) {
  Window w;
  int lin_x, lin_y;
  computeLinearSampleSplit(c.control_type, c.max_linear_samples, lin_x, lin_y);
  /* ref: trajectory_sampler.cpp:48 */
  const int ang_n = c.max_angular_samples + 1 - (c.max_angular_samples % 2);
  /* ref: trajectory_sampler.cpp:51-54: vy limits discarded for non-omni */
  double vy_max = c.vy_max, vy_acc = c.vy_acc, vy_dec = c.vy_dec;
  if (c.control_type != ORC_OMNI) vy_max = vy_acc = vy_dec = 0.0;

  w.max_vx = std::min(c.vx_max, vel[0] + c.vx_acc * c.time_step);
  w.min_vx = std::max(-c.vx_max, vel[0] - c.vx_dec * c.time_step);
  if (c.control_type == ORC_OMNI) {
    w.max_vy = std::min(vy_max, vel[1] + vy_acc * c.time_step);
    w.min_vy = std::max(-vy_max, vel[1] - vy_dec * c.time_step);
  } else {
    w.max_vy = 0.0;
    w.min_vy = 0.0;
  }
  w.res_x = std::max((w.max_vx - w.min_vx) / (lin_x - 1), 0.001);
  w.res_y = (lin_y > 1) ? std::max((w.max_vy - w.min_vy) / (lin_y - 1), 0.001) : 0.001;
  w.max_om = std::min(c.omega_max, vel[2] + c.omega_acc * c.time_step);
  w.min_om = std::max(-c.omega_max, vel[2] - c.omega_dec * c.time_step);
  w.res_om = std::max((w.max_om - w.min_om) / (ang_n - 1), 0.001);
  return w;
}

struct Vel {
  double vx, vy, om;
};

/* serial enumeration order — ref: trajectory_sampler.cpp:207-217 (non-holonomic),
 * :256-272 (holonomic, single-thread branch; the threaded branch has a different order, quirk q11) */
std::vector<Vel> enumerateVelocities(const orc_sampler_cfg &c, const double vel[3]) {
  const Window w = reachableWindow(c, vel);
  std::vector<Vel> out;
  if (c.control_type == ORC_OMNI) {
    for (double vx = w.min_vx; vx <= w.max_vx; vx += w.res_x) {
      for (double vy = w.min_vy; vy <= w.max_vy; vy += w.res_y) out.push_back({vx, vy, 0.0});
      if (std::abs(vx) >= MIN_VEL) {
        for (double om = w.min_om; om <= w.max_om; om += w.res_om) out.push_back({vx, 0.0, om});
      }
    }
  } else {
    for (double vx = w.min_vx; vx <= w.max_vx; vx += w.res_x) {
      if (std::abs(vx) >= MIN_VEL) {
        for (double om = w.min_om; om <= w.max_om; om += w.res_om) out.push_back({vx, 0.0, om});
      }
    }
  }
  return out;
}

/* ------------------------------------------------------------------------------------------
 * Collision model. ref: include/utils/collision_check.h:91-136 (octomap rebuild per call),
 * src/utils/collision_check.cpp:38-58 (robot solid), :118-135 (transforms), :149-162 (collide).
 *
 * Third-party restatement (FCL 0.7.0 / octomap, pinned by build_dependencies/install_linux.sh:41,54):
 *  - octomap::OcTree::insertPointCloud(cloud, origin): after a clear(), the occupied leaves are
 *    exactly the unique keys of the end points, key = floor(coord / resolution) per axis
 *    (coordToKeyChecked, |key| < 2^15); free-space ray cells never collide.
 *  - fcl::collide(shape, OcTree): true iff the shape intersects an occupied leaf cube
 *    [k*res,(k+1)*res]^3 placed by the octree object's transform (sensor_tf_world_).
 *  - The boolean is evaluated here as an exact closed-set test in double (FCL uses float GJK/MPR
 *    with 1e-6 tolerance, so results can differ only within that tolerance of tangency).
 * The robot is upright (pose = x,y,yaw) and the octree transform is required to be planar
 * (rotation about z only); non-planar sensor mounts are reported as unsupported (-2).
 * -------------------------------------------------------------------------------------- */
/* (the voxel model itself lives in voxel_model.h, shared with the FCL/octomap stand-in of the _ref build) */
orc::Quat quatOf(const float q[4]) { return orc::Quat{q[0], q[1], q[2], q[3]}; }

/* ref: collision_check.h:99-116 (laserscan branch) */
int buildWorldScan(CollisionWorld &W, const orc_sampler_cfg &c, const double pose[3],
                   const double *ranges, const double *angles, int32_t n) {
  const orc::Iso3 sensor_tf_body = orc::makeTransform(quatOf(c.sensor_rotation), c.sensor_position);
  /* ref: collision_check.cpp:125-135 updateState(current pose) then :101 */
  const orc::Iso3 body_tf = orc::stateToTransform(pose[0], pose[1], pose[2]);
  const orc::Iso3 sensor_tf_world = orc::mul(body_tf, sensor_tf_body);
  initWorld(W, c.robot_shape, c.robot_dims, c.octree_resolution, sensor_tf_world);
  if (!vox::worldSupported(W)) return -2;
  const float height_in_sensor = (float)(-(double)sensor_tf_body.t[2] / 2.0);
  for (int32_t i = 0; i < n; ++i) {
    const double angle = angles[i], r = ranges[i];
    if (std::isfinite(r)) {
      const float x = (float)(r * std::cos(angle));
      const float y = (float)(r * std::sin(angle));
      insertPoint(W, x, y, height_in_sensor);
    }
  }
  return 0;
}

/* ref: collision_check.h:119-131 (cloud branch, global_frame = true => identity transform) */
int buildWorldCloud(CollisionWorld &W, const orc_sampler_cfg &c, const float *xyz, int32_t n) {
  orc::Iso3 ident;
  ident.L = orc::identity3();
  ident.t[0] = ident.t[1] = ident.t[2] = 0.0f;
  initWorld(W, c.robot_shape, c.robot_dims, c.octree_resolution, ident);
  for (int32_t i = 0; i < n; ++i) insertPoint(W, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * rollout — ref: src/utils/trajectory_sampler.cpp:118-179, include/datatypes/path.h:24-30
 * returns true if admissible; fills the row buffers (P-1 velocities, P points)
 * ---------------------------------------------------------------------------------------- */
bool rolloutSample(const orc_sampler_cfg &c, const CollisionWorld &W, const Vel &v,
                   const double pose[3], size_t P, float *rvx, float *rvy, float *rom, float *rx,
                   float *ry) {
  if (std::abs(v.vx) < MIN_VEL && std::abs(v.vy) < MIN_VEL && std::abs(v.om) < MIN_VEL)
    return false;
  double sx = pose[0], sy = pose[1], syaw = pose[2];
  const float timeStep = (float)c.time_step; /* ref: path.h:24 `const float timeStep` */
  rx[0] = (float)pose[0];
  ry[0] = (float)pose[1];
  bool is_collision = false;
  size_t last_free_index = P - 1;
  for (size_t i = 0; i < P - 1; ++i) {
    /* ref: path.h:25-29 (all three use the pre-update yaw) */
    const double cy = std::cos(syaw), sn = std::sin(syaw);
    sx += (v.vx * cy - v.vy * sn) * timeStep;
    sy += (v.vx * sn + v.vy * cy) * timeStep;
    syaw += v.om * timeStep;
    is_collision = poseCollides(W, sx, sy, syaw);
    if (is_collision) {
      if (i > 0) last_free_index = i - 1;
      break;
    }
    rvx[i] = (float)v.vx;
    rvy[i] = (float)v.vy;
    rom[i] = (float)v.om;
    rx[i + 1] = (float)sx;
    ry[i + 1] = (float)sy;
  }
  if (!c.drop_samples && is_collision && (int64_t)last_free_index > c.num_ctrl_points &&
      last_free_index < P - 1) {
    const float lx = rx[last_free_index], ly = ry[last_free_index];
    for (size_t j = last_free_index + 1; j < P - 1; ++j) {
      rvx[j] = rvy[j] = rom[j] = 0.0f;
      rx[j + 1] = lx;
      ry[j + 1] = ly;
    }
    is_collision = false;
  }
  return !is_collision;
}

int32_t generate(const orc_sampler_cfg &c, const CollisionWorld &W, const double vel[3],
                 const double pose[3], float *vx, float *vy, float *om, float *x, float *y,
                 int32_t *slot_of_row, int32_t cap) {
  const size_t P = (size_t)orc_num_points(c.time_step, c.prediction_horizon);
  if (P < 2) return -1;
  const std::vector<Vel> vels = enumerateVelocities(c, vel);
  const size_t n = vels.size();
  std::vector<uint8_t> ok(n, 0);
  std::vector<float> tvx(n * (P - 1)), tvy(n * (P - 1)), tom(n * (P - 1)), tx(n * P), ty(n * P);
  auto work = [&](size_t lo, size_t hi) {
    for (size_t s = lo; s < hi; ++s)
      ok[s] = rolloutSample(c, W, vels[s], pose, P, &tvx[s * (P - 1)], &tvy[s * (P - 1)],
                            &tom[s * (P - 1)], &tx[s * P], &ty[s * P]);
  };
  const int nt = std::max(1, c.max_num_threads);
  if (nt == 1) {
    work(0, n);
  } else {
    /* one task per sample in the reference (ThreadPool, trajectory_sampler.cpp:192-205);
     * here contiguous chunks, results compacted in enumeration order (deterministic). */
    std::vector<std::thread> th;
    std::atomic<size_t> next{0};
    const size_t chunk = 16;
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&] {
        for (;;) {
          const size_t lo = next.fetch_add(chunk);
          if (lo >= n) break;
          work(lo, std::min(n, lo + chunk));
        }
      });
    for (auto &t : th) t.join();
  }
  int32_t rows = 0;
  for (size_t s = 0; s < n; ++s) {
    if (!ok[s]) continue;
    if (rows >= cap) return -3;
    std::memcpy(vx + (size_t)rows * (P - 1), &tvx[s * (P - 1)], (P - 1) * sizeof(float));
    std::memcpy(vy + (size_t)rows * (P - 1), &tvy[s * (P - 1)], (P - 1) * sizeof(float));
    std::memcpy(om + (size_t)rows * (P - 1), &tom[s * (P - 1)], (P - 1) * sizeof(float));
    std::memcpy(x + (size_t)rows * P, &tx[s * P], P * sizeof(float));
    std::memcpy(y + (size_t)rows * P, &ty[s * P], P * sizeof(float));
    if (slot_of_row) slot_of_row[rows] = (int32_t)s;
    ++rows;
  }
  return rows;
}

/* ------------------------------------------------------------------------------------------
 * cost functions — ref: src/utils/cost_evaluator.cpp:111-233, include/datatypes/trajectory.h:218-235
 * ---------------------------------------------------------------------------------------- */
struct SegView {
  const float *X, *Y;
  const float *acc; /* parent's accumulated lengths (absolute), size acc_n */
  int32_t acc_n;
  int32_t start, count;
};

/* Eigen: (p1 - p2).squaredNorm() on Vector3f with z == 0: dx*dx + (dy*dy + 0) */
inline float sqDist(float ax, float ay, float bx, float by) {
  const float dx = ax - bx, dy = ay - by;
  return dx * dx + (dy * dy + 0.0f);
}

float pathCostFunc(const float *px, const float *py, int32_t P, const SegView &seg,
                   float tracked_segment_length) {
  float total_cost = 0.0f;
  for (int32_t i = 0; i < P; ++i) {
    float min_dist = DEFAULT_MIN_DIST;
    for (int32_t j = 0; j < seg.count; ++j) {
      const float d =
          std::sqrt(sqDist(seg.X[seg.start + j], seg.Y[seg.start + j], px[i], py[i]));
      if (d < min_dist) min_dist = d;
    }
    total_cost += min_dist;
  }
  const int32_t last = seg.start + seg.count - 1;
  const float end_dist_error =
      std::sqrt(sqDist(px[P - 1], py[P - 1], seg.X[last], seg.Y[last])) / tracked_segment_length;
  return (total_cost / (float)(int64_t)P + end_dist_error) / 2;
}

float goalCostFunc(const float *px, const float *py, int32_t P, const SegView &seg,
                   float ref_path_length) {
  const float ex = px[P - 1], ey = py[P - 1];
  float min_dist_sq = DEFAULT_MIN_DIST;
  size_t closest_local_idx = 0;
  for (int32_t i = 0; i < seg.count; ++i) {
    const float d_sq = sqDist(ex, ey, seg.X[seg.start + i], seg.Y[seg.start + i]);
    if (d_sq < min_dist_sq) {
      min_dist_sq = d_sq;
      closest_local_idx = (size_t)i;
    }
  }
  const size_t closest_abs_idx = closest_local_idx + (size_t)seg.start;
  /* ref: path.h:190-194 getDistanceAtIndex: out of range -> 0 */
  const float at = (closest_abs_idx >= (size_t)seg.acc_n) ? 0.0f : seg.acc[closest_abs_idx];
  const float arc_remaining_normalized = (ref_path_length - at) / ref_path_length;
  return arc_remaining_normalized + (std::sqrt(min_dist_sq) / ref_path_length);
}

/* ref: trajectory.h:218-235 minDist2D: obstacles outer, points inner; pow(float,2) -> double */
float minDist2D(const float *px, const float *py, int32_t P, const float *ox, const float *oy,
                int32_t n_obs) {
  if (n_obs <= 0) return 0.0f;
  float minDist = DEFAULT_MIN_DIST;
  for (int32_t i = 0; i < n_obs; ++i) {
    for (int32_t j = 0; j < P; ++j) {
      const float dx = ox[i] - px[j], dy = oy[i] - py[j];
      const float dist = (float)((double)dx * (double)dx + (double)dy * (double)dy);
      if (dist < minDist) minDist = dist;
    }
  }
  return (float)std::sqrt((double)minDist);
}

float obstaclesDistCostFunc(const float *px, const float *py, int32_t P, const float *ox,
                            const float *oy, int32_t n_obs, float maxObstaclesDist) {
  const float dist = minDist2D(px, py, P, ox, oy, n_obs);
  return std::max(maxObstaclesDist - dist, 0.0f) / maxObstaclesDist;
}

float smoothnessCostFunc(const float *vx, const float *vy, const float *om, int32_t nv,
                         const float acc[3]) {
  float cost = 0.0f;
  for (int32_t i = 1; i < nv; ++i) {
    if (acc[0] > 0) {
      const float d = vx[i] - vx[i - 1];
      cost += (double)d * (double)d / (double)acc[0];
    }
    if (acc[1] > 0) {
      const float d = vy[i] - vy[i - 1];
      cost += (double)d * (double)d / (double)acc[1];
    }
    if (acc[2] > 0) {
      const float d = om[i] - om[i - 1];
      cost += (double)d * (double)d / (double)acc[2];
    }
  }
  return cost / (float)(int64_t)(3 * (int64_t)nv);
}

float jerkCostFunc(const float *vx, const float *vy, const float *om, int32_t nv,
                   const float acc[3]) {
  float cost = 0.0f;
  for (int32_t i = 2; i < nv; ++i) {
    if (acc[0] > 0) {
      const float j = vx[i] - 2 * vx[i - 1] + vx[i - 2];
      cost += (double)j * (double)j / (double)acc[0];
    }
    if (acc[1] > 0) {
      const float j = vy[i] - 2 * vy[i - 1] + vy[i - 2];
      cost += (double)j * (double)j / (double)acc[1];
    }
    if (acc[2] > 0) {
      const float j = om[i] - 2 * om[i - 1] + om[i - 2];
      cost += (double)j * (double)j / (double)acc[2];
    }
  }
  return cost / (float)(int64_t)(3 * (int64_t)nv);
}

} // namespace

/* ============================================================================================
 * exported API
 * ========================================================================================== */
extern "C" {

int64_t orc_num_trajectories(int32_t control_type, int32_t max_linear, int32_t max_angular) {
  /* ref: trajectory.h:32-45 */
  const int angSlots = max_angular + 1 - (max_angular % 2);
  int vx_n, vy_n;
  computeLinearSampleSplit(control_type, max_linear, vx_n, vy_n);
  if (control_type == ORC_OMNI) return (int64_t)vx_n * angSlots + (int64_t)vx_n * vy_n;
  return (int64_t)vx_n * angSlots;
}

int64_t orc_num_points(double time_step, double prediction_horizon) {
  /* ref: trajectory.h:48-51 (size_t truncation of the double quotient) */
  return (int64_t)(size_t)(prediction_horizon / time_step);
}

/* ref: src/datatypes/path.cpp:167-288 with tk::spline::linear (include/utils/spline.h:211-224,
 * 390-419): b[i] = (y[i+1]-y[i])/(x[i+1]-x[i]), b[n-1] = b[n-2];
 * value(s) = b[idx]*h + y[idx] with idx = upper_bound(x, s) - 1 (clamped at 0), h = s - x[idx];
 * (the c/d Horner terms are exact zeros in linear mode: ((0*h + 0)*h + b)*h + y). */
int32_t orc_path_interpolate_linear(const float *x, const float *y, int32_t n, double max_dist,
                                    float *X, float *Y, float *acc, float *curv, int32_t cap,
                                    float *total_length) {
  if (n < 2) return -1;
  std::vector<double> s_vals, x_vals, y_vals;
  s_vals.push_back(0.0);
  x_vals.push_back(x[0]);
  y_vals.push_back(y[0]);
  float current_total_length = 0.0f; /* ref: path.h:296 float member, += double */
  for (int32_t i = 1; i < n; ++i) {
    const double seg_dist = std::hypot(x[i] - x[i - 1], y[i] - y[i - 1]); /* float args -> hypotf? */
    current_total_length += seg_dist;
    s_vals.push_back(current_total_length);
    x_vals.push_back(x[i]);
    y_vals.push_back(y[i]);
  }
  const int m = (int)s_vals.size();
  std::vector<double> bx(m), by(m);
  for (int i = 0; i < m - 1; ++i) {
    bx[i] = (x_vals[i + 1] - x_vals[i]) / (s_vals[i + 1] - s_vals[i]);
    by[i] = (y_vals[i + 1] - y_vals[i]) / (s_vals[i + 1] - s_vals[i]);
  }
  bx[m - 1] = bx[m - 2];
  by[m - 1] = by[m - 2];
  auto eval = [&](const std::vector<double> &b, const std::vector<double> &yv, double s) {
    auto it = std::upper_bound(s_vals.begin(), s_vals.end(), s);
    const size_t idx = (size_t)std::max((int)(it - s_vals.begin()) - 1, 0);
    const double h = s - s_vals[idx];
    if (s < s_vals[0]) return (0.0 * h + b[0]) * h + yv[0];
    if (s > s_vals[m - 1]) return (0.0 * h + b[m - 1]) * h + yv[m - 1];
    return ((0.0 * h + 0.0) * h + b[idx]) * h + yv[idx];
  };
  const size_t new_size = (size_t)(current_total_length / max_dist) + 1;
  if ((int64_t)new_size > cap) return -3;
  for (size_t i = 0; i < new_size; ++i) {
    X[i] = Y[i] = 0.0f; /* Eigen resize leaves garbage; only [0,idx) is ever read */
    acc[i] = 0.0f;      /* std::vector::resize value-initialises */
    curv[i] = 0.0f;
  }
  size_t idx = 0;
  for (double s = 0.0; s <= current_total_length && idx < new_size; s += max_dist) {
    acc[idx] = (float)s;
    X[idx] = (float)eval(bx, x_vals, s);
    Y[idx] = (float)eval(by, y_vals, s);
    idx++;
  }
  if (idx < new_size && idx > 0) { /* appended exact end point; its acc entry stays 0 (quirk) */
    X[idx] = (float)eval(bx, x_vals, current_total_length);
    Y[idx] = (float)eval(by, y_vals, current_total_length);
    idx++;
  }
  /* curvature: ref path.cpp:260-287 */
  if (idx >= 2) {
    float dx_old = X[1] - X[0], dy_old = Y[1] - Y[0];
    for (size_t i = 1; i + 1 < idx; ++i) {
      const float dx = X[i + 1] - X[i], dy = Y[i + 1] - Y[i];
      const float ddx = dx - dx_old, ddy = dy - dy_old;
      const float val = dx * dx + dy * dy;
      const float den = val * std::sqrt(val);
      curv[i] = (den > 1e-6f) ? (dx_old * ddy - ddx * dy_old) / den : 0.0f;
      dx_old = dx;
      dy_old = dy;
    }
  }
  /* ref: path.cpp:146-155 totalPathLength(): a path that collapsed to one point reports 0 */
  *total_length = (idx < 2) ? 0.0f : current_total_length;
  return (int32_t)idx;
}

int32_t orc_path_segment(const float *acc, int32_t n, double segment_length,
                         int64_t max_points_per_segment, int32_t *seg_starts, int32_t cap) {
  if (n < 2) return 0;
  int32_t ns = 0;
  seg_starts[ns++] = 0;
  size_t segmentStartIdx = 0;
  float segmentStartLength = acc[0];
  for (int32_t i = 1; i < n; ++i) {
    const size_t pointsInSegment = (size_t)i - segmentStartIdx + 1;
    const float segmentLength = acc[i] - segmentStartLength;
    const bool lengthExceeded = (segment_length > 0.0 && segmentLength >= segment_length);
    const bool pointsExceeded =
        (max_points_per_segment > 0 && pointsInSegment > (size_t)max_points_per_segment);
    if (lengthExceeded || pointsExceeded) {
      if (ns >= cap) return -3;
      seg_starts[ns++] = i;
      segmentStartIdx = (size_t)i;
      segmentStartLength = acc[i];
    }
  }
  return ns;
}

float orc_segment_length(const float *X, const float *Y, int32_t start, int32_t count) {
  float length = 0.0f;
  for (int32_t i = 0; i + 1 < count; ++i)
    length += std::sqrt(sqDist(X[start + i], Y[start + i], X[start + i + 1], Y[start + i + 1]));
  return length;
}

int32_t orc_velocity_samples(const orc_sampler_cfg *cfg, const double vel[3], double *vx,
                             double *vy, double *omega, int32_t cap) {
  const std::vector<Vel> v = enumerateVelocities(*cfg, vel);
  if ((int64_t)v.size() > cap) return -3;
  for (size_t i = 0; i < v.size(); ++i) {
    vx[i] = v[i].vx;
    vy[i] = v[i].vy;
    omega[i] = v[i].om;
  }
  return (int32_t)v.size();
}

int32_t orc_sampler_generate_scan(const orc_sampler_cfg *cfg, const double vel[3],
                                  const double pose[3], const double *ranges, const double *angles,
                                  int32_t n, float *vx, float *vy, float *omega, float *x, float *y,
                                  int32_t *slot_of_row, int32_t cap) {
  CollisionWorld W;
  const int rc = buildWorldScan(W, *cfg, pose, ranges, angles, n);
  if (rc) return rc;
  return generate(*cfg, W, vel, pose, vx, vy, omega, x, y, slot_of_row, cap);
}

int32_t orc_sampler_generate_cloud(const orc_sampler_cfg *cfg, const double vel[3],
                                   const double pose[3], const float *xyz, int32_t n, float *vx,
                                   float *vy, float *omega, float *x, float *y,
                                   int32_t *slot_of_row, int32_t cap) {
  CollisionWorld W;
  const int rc = buildWorldCloud(W, *cfg, xyz, n);
  if (rc) return rc;
  return generate(*cfg, W, vel, pose, vx, vy, omega, x, y, slot_of_row, cap);
}

int32_t orc_check_collision(const orc_sampler_cfg *cfg, const double sensor_pose[3],
                            const double query_pose[3], int32_t is_cloud, const void *a,
                            const void *b, int32_t n) {
  CollisionWorld W;
  int rc;
  if (is_cloud)
    rc = buildWorldCloud(W, *cfg, (const float *)a, n);
  else
    rc = buildWorldScan(W, *cfg, sensor_pose, (const double *)a, (const double *)b, n);
  if (rc) return rc;
  return poseCollides(W, query_pose[0], query_pose[1], query_pose[2]) ? 1 : 0;
}

/* CollisionChecker as its other users drive it (pure_pursuit.cpp:154-155, ompl.cpp:95-97,
 * trajectory_sampler.cpp:378-408): sensor data inserted with the body at sensor_pose
 * (collision_check.h:91-136; a cloud with global_frame != 0 uses the identity transform), then
 * checkCollisions for each of n_states states (x, y, yaw). Returns 1 if any state collides, a
 * negative value for an unsupported (tilted) mount. */
int32_t orc_check_collision_states(const orc_sampler_cfg *cfg, const double sensor_pose[3],
                                   int32_t is_cloud, int32_t global_frame, const void *a,
                                   const void *b, int32_t n, const double *states,
                                   int32_t n_states, uint8_t *out) {
  CollisionWorld W;
  int rc = 0;
  if (!is_cloud) {
    rc = buildWorldScan(W, *cfg, sensor_pose, (const double *)a, (const double *)b, n);
  } else if (global_frame) {
    rc = buildWorldCloud(W, *cfg, (const float *)a, n);
  } else { /* ref: collision_check.h:124-125: sensor_tf_world_ = body->tf * sensor_tf_body_ */
    const orc::Iso3 sensor_tf_body =
        orc::makeTransform(quatOf(cfg->sensor_rotation), cfg->sensor_position);
    const orc::Iso3 body_tf = orc::stateToTransform(sensor_pose[0], sensor_pose[1], sensor_pose[2]);
    initWorld(W, cfg->robot_shape, cfg->robot_dims, cfg->octree_resolution, orc::mul(body_tf, sensor_tf_body));
    if (!vox::worldSupported(W)) return -2;
    const float *xyz = (const float *)a;
    for (int32_t i = 0; i < n; ++i) insertPoint(W, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  }
  if (rc) return rc;
  int32_t any = 0;
  for (int32_t i = 0; i < n_states; ++i) {
    const bool hit = poseCollides(W, states[3 * i], states[3 * i + 1], states[3 * i + 2]);
    if (out) out[i] = hit ? 1 : 0;
    any |= hit ? 1 : 0;
  }
  return any;
}

/* ref: include/utils/cost_evaluator.h:174-193. NB operand order sensor_tf_body_ * body_tf_world_
 * (quirk q7) and no isfinite filter (quirk q8). */
void orc_cost_points_scan(const orc_cost_cfg *cfg, const double *ranges, const double *angles,
                          int32_t n, const double pose[3], float *ox, float *oy) {
  const orc::Iso3 sensor_tf_body =
      orc::makeTransform(quatOf(cfg->sensor_rotation), cfg->sensor_position);
  const orc::Iso3 body_tf_world = orc::stateToTransform(pose[0], pose[1], pose[2]);
  const orc::Iso3 T = orc::mul(sensor_tf_body, body_tf_world);
  for (int32_t i = 0; i < n; ++i) {
    const double point_x = ranges[i] * std::cos(angles[i]);
    const double point_y = ranges[i] * std::sin(angles[i]);
    const float p[3] = {(float)point_x, (float)point_y, 0.0f};
    float o[3];
    orc::apply(T, p, o);
    ox[i] = o[0];
    oy[i] = o[1];
  }
}

void orc_cost_points_cloud(const orc_cost_cfg *cfg, const float *xyz, int32_t n,
                           const double pose[3], float *ox, float *oy) {
  const orc::Iso3 sensor_tf_body =
      orc::makeTransform(quatOf(cfg->sensor_rotation), cfg->sensor_position);
  const orc::Iso3 body_tf_world = orc::stateToTransform(pose[0], pose[1], pose[2]);
  const orc::Iso3 T = orc::mul(sensor_tf_body, body_tf_world);
  for (int32_t i = 0; i < n; ++i) {
    float o[3];
    orc::apply(T, &xyz[3 * i], o);
    ox[i] = o[0];
    oy[i] = o[1];
  }
}

int32_t orc_cost_evaluate(const orc_cost_cfg *cfg, int32_t n_traj, int32_t P, const float *vx,
                          const float *vy, const float *omega, const float *x, const float *y,
                          const float *pathX, const float *pathY, const float *pathAcc,
                          int32_t path_n, float path_total_length, int32_t seg_start,
                          int32_t seg_count, const float *ox, const float *oy, int32_t n_obs,
                          float max_obstacles_dist, const double *custom, int32_t n_custom,
                          float *costs_out,
                          int32_t *best_idx, float *best_cost, int32_t n_threads) {
  SegView seg{pathX, pathY, pathAcc, path_n, seg_start, seg_count};
  /* ref: cost_evaluator.cpp:71: totalSegmentLength() re-evaluated per trajectory; same value */
  const float seg_len = (seg_count > 0) ? orc_segment_length(pathX, pathY, seg_start, seg_count) : 0.f;
  std::vector<float> totals(n_traj);
  auto work = [&](int32_t lo, int32_t hi) {
    for (int32_t t = lo; t < hi; ++t) {
      const float *tx = x + (size_t)t * P, *ty = y + (size_t)t * P;
      const float *tvx = vx + (size_t)t * (P - 1), *tvy = vy + (size_t)t * (P - 1),
                  *tom = omega + (size_t)t * (P - 1);
      double weight;
      float total_cost = 0.0f;
      const float ref_path_length = path_total_length;
      if (ref_path_length > 0.0) {
        if ((weight = cfg->w_goal) > 0.0) {
          const float goalCost = goalCostFunc(tx, ty, P, seg, ref_path_length);
          total_cost += weight * goalCost;
        }
        if ((weight = cfg->w_path) > 0.0) {
          const float refPathCost = pathCostFunc(tx, ty, P, seg, seg_len);
          total_cost += weight * refPathCost;
        }
      }
      if (n_obs > 0 && (weight = cfg->w_obstacles) > 0.0) {
        const float objCost = obstaclesDistCostFunc(tx, ty, P, ox, oy, n_obs, max_obstacles_dist);
        total_cost += weight * objCost;
      }
      if ((weight = cfg->w_smooth) > 0.0) {
        const float c = smoothnessCostFunc(tvx, tvy, tom, P - 1, cfg->acc_limits);
        total_cost += weight * c;
      }
      if ((weight = cfg->w_jerk) > 0.0) {
        const float c = jerkCostFunc(tvx, tvy, tom, P - 1, cfg->acc_limits);
        total_cost += weight * c;
      }
      /* ref: cost_evaluator.cpp:96-100: one `float += double` per registered callback */
      if (custom)
        for (int32_t k = 0; k < n_custom; ++k) total_cost += custom[(size_t)t * n_custom + k];
      totals[t] = total_cost;
    }
  };
  if (n_threads <= 1) {
    work(0, n_traj);
  } else {
    std::vector<std::thread> th;
    std::atomic<int32_t> next{0};
    for (int t = 0; t < n_threads; ++t)
      th.emplace_back([&] {
        for (;;) {
          const int32_t lo = next.fetch_add(8);
          if (lo >= n_traj) break;
          work(lo, std::min(n_traj, lo + 8));
        }
      });
    for (auto &t : th) t.join();
  }
  /* running argmin with strict '<' from FLT_MAX: lowest index wins ties (cost_evaluator.cpp:102) */
  float minCost = DEFAULT_MIN_DIST;
  int32_t idx = -1;
  for (int32_t t = 0; t < n_traj; ++t) {
    if (costs_out) costs_out[t] = totals[t];
    if (totals[t] < minCost) {
      minCost = totals[t];
      idx = t;
    }
  }
  *best_idx = idx;
  *best_cost = minCost;
  return idx >= 0 ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * mapper — ref: src/mapping/local_mapper.cpp:80-104,127-159,204-220,
 *          include/mapping/local_mapper.h:26-27,210-222, include/mapping/line_drawing.h:55-124
 * ---------------------------------------------------------------------------------------- */
void orc_mapper_scan_to_grid(int32_t H, int32_t W, float resolution, const float laser_pos[3],
                             float laser_orientation, const double *angles, const double *ranges,
                             int32_t n, int32_t *grid) {
  const int UNEXPLORED = -1, EMPTY = 0, OCCUPIED = 100;
  const int c0 = (int)std::round(H / 2) - 1, c1 = (int)std::round(W / 2) - 1;
  auto localToGrid = [&](float px, float py, int &gi, int &gj) {
    gi = c0 + static_cast<int>(px / resolution);
    gj = c1 + static_cast<int>(py / resolution);
  };
  int s0, s1;
  localToGrid(laser_pos[0], laser_pos[1], s0, s1);
  for (int64_t i = 0; i < (int64_t)H * W; ++i) grid[i] = UNEXPLORED;
  for (int32_t r = 0; r < n; ++r) {
    const float angle = (float)angles[r], range = (float)ranges[r];
    /* float + (float * double cos(float)) -> double, narrowed on assignment */
    const float x = (float)((double)laser_pos[0] + ((double)range * std::cos((double)(laser_orientation + angle))));
    const float y = (float)((double)laser_pos[1] + ((double)range * std::sin((double)(laser_orientation + angle))));
    int t0, t1;
    localToGrid(x, y, t0, t1);
    auto visit = [&](int px, int py) {
      if (px >= 0 && px < H && py >= 0 && py < W) {
        int32_t &cell = grid[(int64_t)px + (int64_t)py * H];
        if (px == t0 && py == t1)
          cell = OCCUPIED; /* fillGridAroundPoint(pad 0) */
        else
          cell = std::max(cell, EMPTY);
      }
    };
    /* bresenhamEnhanced(start, to) */
    int px = s0, py = s1;
    int dx = t0 - s0, dy = t1 - s1;
    visit(px, py);
    const int xstep = (dx >= 0) ? 1 : -1, ystep = (dy >= 0) ? 1 : -1;
    dx = std::abs(dx);
    dy = std::abs(dy);
    const int ddy = 2 * dy, ddx = 2 * dx;
    if (ddx >= ddy) {
      int errorprev = dx, error = dx;
      for (int i = 0; i < dx; i++) {
        px += xstep;
        error += ddy;
        if (error > ddx) {
          py += ystep;
          error -= ddx;
          if (error + errorprev < ddx) {
            visit(px, py - ystep);
          } else if (error + errorprev > ddx) {
            visit(px - xstep, py);
          } else {
            visit(px - xstep, py);
            visit(px, py - ystep);
          }
        }
        visit(px, py);
        errorprev = error;
      }
    } else {
      int errorprev = dy, error = dy;
      for (int i = 0; i < dy; i++) {
        py += ystep;
        error += ddx;
        if (error > ddy) {
          px += xstep;
          error -= ddy;
          if (error + errorprev < ddy) {
            visit(px - xstep, py);
          } else if (error + errorprev > ddy) {
            visit(px, py - ystep);
          } else {
            visit(px - xstep, py);
            visit(px, py - ystep);
          }
        }
        visit(px, py);
        errorprev = error;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Bayesian mapper — ref: src/mapping/local_mapper.cpp:106-125
 * (updateGridCellProbability), :161-202 (updateGridBaysian_), :222-238 (scanToGridBaysian),
 * :17-78 (getPreviousGridInCurrentPose). Serial ray order: the LAST ray that crosses a cell
 * decides its probability (each write overwrites). PARITY UNPINNED against a real Eigen build for:
 *  - (pt - m_startPoint).norm() on Vector2i: Eigen's integer norm is sqrt_impl<int> =
 *    (int)std::sqrt(int) (Eigen/src/Core/MathFunctions.h), i.e. the TRUNCATED cell distance;
 *  - Matrix3f::inverse(): compute_inverse_size3 (Eigen/src/LU/InverseImpl.h): cofactor method,
 *    det = c00*m00 + (c10*m10 + c20*m20), inverse(i,j) = cofactor<j,i> * (1/det);
 *  - Matrix3f * Vector3f: row dot products as a0 + (a1 + a2).
 * The reference's tests only log these grids (tests/mapper_test.cpp:137-215), nothing pins them.
 * ---------------------------------------------------------------------------------------- */
static float bayesCellProbability(float distance, float currentRange, float previousProb,
                                  float resolution, float pPrior, float pEmpty, float pOccupied,
                                  float rangeSure, float rangeMax, float wallSize) {
  distance = distance * resolution;
  currentRange = currentRange - wallSize;
  float pF = (distance < currentRange) ? pEmpty : pOccupied;
  float delta = (distance < rangeSure) ? 0.0 : 1.0;
  float pSensor = pF + (delta * ((distance - rangeSure) / rangeMax) * (pPrior - pF));
  float pCurr = 1 - (1 / (1 + ((previousProb / (1 - previousProb)) * (pSensor / (1.0 - pSensor)) *
                               ((1 - pPrior) / pPrior))));
  return pCurr;
}

void orc_mapper_scan_to_grid_bayes(int32_t H, int32_t W, float resolution, const float laser_pos[3],
                                   float laser_orientation, float pPrior, float pOccupied,
                                   float pEmpty, float rangeSure, float rangeMax, float wallSize,
                                   const double *angles, const double *ranges, int32_t n,
                                   const float *prev, int32_t *grid, float *prob) {
  const int UNEXPLORED = -1, EMPTY = 0, OCCUPIED = 100;
  const int c0 = (int)std::round(H / 2) - 1, c1 = (int)std::round(W / 2) - 1;
  const int s0 = c0 + static_cast<int>(laser_pos[0] / resolution);
  const int s1 = c1 + static_cast<int>(laser_pos[1] / resolution);
  for (int64_t i = 0; i < (int64_t)H * W; ++i) {
    grid[i] = UNEXPLORED;
    prob[i] = pPrior;
  }
  for (int32_t r = 0; r < n; ++r) {
    const float angle = (float)angles[r], range = (float)ranges[r];
    const float x = (float)((double)laser_pos[0] + ((double)range * std::cos((double)(laser_orientation + angle))));
    const float y = (float)((double)laser_pos[1] + ((double)range * std::sin((double)(laser_orientation + angle))));
    const int t0 = c0 + static_cast<int>(x / resolution), t1 = c1 + static_cast<int>(y / resolution);
    std::vector<std::pair<int, int>> pts;
    auto push = [&](int a, int b) { pts.emplace_back(a, b); };
    /* bresenhamEnhanced(start, to): include/mapping/line_drawing.h:55-124 */
    int px = s0, py = s1;
    int dx = t0 - s0, dy = t1 - s1;
    push(px, py);
    const int xstep = (dx >= 0) ? 1 : -1, ystep = (dy >= 0) ? 1 : -1;
    dx = std::abs(dx);
    dy = std::abs(dy);
    const int ddy = 2 * dy, ddx = 2 * dx;
    if (ddx >= ddy) {
      int errorprev = dx, error = dx;
      for (int i = 0; i < dx; i++) {
        px += xstep;
        error += ddy;
        if (error > ddx) {
          py += ystep;
          error -= ddx;
          if (error + errorprev < ddx) {
            push(px, py - ystep);
          } else if (error + errorprev > ddx) {
            push(px - xstep, py);
          } else {
            push(px - xstep, py);
            push(px, py - ystep);
          }
        }
        push(px, py);
        errorprev = error;
      }
    } else {
      int errorprev = dy, error = dy;
      for (int i = 0; i < dy; i++) {
        py += ystep;
        error += ddx;
        if (error > ddy) {
          px += xstep;
          error -= ddy;
          if (error + errorprev < ddy) {
            push(px - xstep, py);
          } else if (error + errorprev > ddy) {
            push(px, py - ystep);
          } else {
            push(px - xstep, py);
            push(px, py - ystep);
          }
        }
        push(px, py);
        errorprev = error;
      }
    }
    for (auto &pt : pts) {
      if (pt.first >= 0 && pt.first < H && pt.second >= 0 && pt.second < W) {
        const int ddx2 = pt.first - s0, ddy2 = pt.second - s1;
        const int normi = (int)std::sqrt(ddx2 * ddx2 + ddy2 * ddy2); /* Vector2i::norm() */
        const float distance = normi;
        const int64_t idx = (int64_t)pt.first + (int64_t)pt.second * H;
        const float newValue = bayesCellProbability(distance, range, prev[idx], resolution, pPrior, pEmpty,
                                                    pOccupied, rangeSure, rangeMax, wallSize);
        if (pt.first == t0 && pt.second == t1)
          grid[idx] = OCCUPIED;
        else
          grid[idx] = std::max(grid[idx], EMPTY);
        prob[idx] = newValue;
      }
    }
  }
}

/* ref: src/mapping/local_mapper.cpp:17-78; prev / out are column-major [H x W] */
void orc_mapper_warp_previous(int32_t H, int32_t W, float resolution, float pPrior, float pos_x,
                              float pos_y, double orientation, const float *prev, float *out) {
  const int c0 = (int)std::round(H / 2) - 1, c1 = (int)std::round(W / 2) - 1;
  const int cc0 = c0 + static_cast<int>(pos_x / resolution), cc1 = c1 + static_cast<int>(pos_y / resolution);
  const double a = -1 * orientation;
  const double cosT = std::cos(a), sinT = std::sin(a);
  float m[3][3];
  m[0][0] = (float)cosT;
  m[0][1] = (float)-sinT;
  m[0][2] = (float)(0.5 * H - cc1 + (cc0 * sinT - cc1 * cosT));
  m[1][0] = (float)sinT;
  m[1][1] = (float)cosT;
  m[1][2] = (float)(0.5 * W - cc0 - (cc0 * cosT + cc1 * sinT));
  m[2][0] = 0.0f;
  m[2][1] = 0.0f;
  m[2][2] = 1.0f;
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
  };
  const float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const float det = orc::sum3(c00 * m[0][0], c10 * m[1][0], c20 * m[2][0]);
  const float invdet = 1.0f / det;
  float inv[3][3];
  inv[0][0] = c00 * invdet;
  inv[0][1] = c10 * invdet;
  inv[0][2] = c20 * invdet;
  inv[1][0] = cof(0, 1) * invdet;
  inv[1][1] = cof(1, 1) * invdet;
  inv[1][2] = cof(2, 1) * invdet;
  inv[2][0] = cof(0, 2) * invdet;
  inv[2][1] = cof(1, 2) * invdet;
  inv[2][2] = cof(2, 2) * invdet;
  for (int64_t i = 0; i < (int64_t)H * W; ++i) out[i] = pPrior;
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      const float sx = (float)x, sy = (float)y, sw = 1.0f;
      const float d0 = orc::sum3(inv[0][0] * sx, inv[0][1] * sy, inv[0][2] * sw);
      const float d1 = orc::sum3(inv[1][0] * sx, inv[1][1] * sy, inv[1][2] * sw);
      const double srcX = d0, srcY = d1;
      if (srcX >= 0 && srcX < W - 1 && srcY >= 0 && srcY < H - 1) {
        const int x0 = static_cast<int>(std::floor(srcX)), y0 = static_cast<int>(std::floor(srcY));
        const int x1 = x0 + 1, y1 = y0 + 1;
        const float w0 = srcX - x0, w1 = 1.0f - w0, h0 = srcY - y0, h1 = 1.0f - h0;
        auto P = [&](int r, int c) { return prev[(int64_t)r + (int64_t)c * H]; };
        const float value = h1 * (w1 * P(y0, x0) + w0 * P(y0, x1)) + h0 * (w1 * P(y1, x0) + w0 * P(y1, x1));
        out[(int64_t)y + (int64_t)x * H] = value;
      }
    }
  }
}

/* ref: include/utils/pointcloud.h:205-259 */
void orc_pointcloud_to_laserscan(const int8_t *data, int64_t nbytes, int32_t point_step,
                                 int32_t row_step, int32_t height, int32_t width, int32_t x_off,
                                 int32_t y_off, int32_t z_off, double max_range, double min_z,
                                 double max_z, int32_t num_bins, double *ranges_out) {
  (void)width;
  const double two_pi = 2.0 * M_PI;
  for (int32_t i = 0; i < num_bins; ++i) ranges_out[i] = max_range;
  if (point_step <= 0) return;
  for (int row = 0; row < height; ++row) {
    for (int col = 0; col < row_step; col += point_step) {
      const std::size_t point_start = (std::size_t)(row * row_step + col);
      const std::size_t max_offset =
          point_start + (std::size_t)std::max({x_off, y_off, z_off}) + sizeof(float);
      if (max_offset > (std::size_t)nbytes) continue;
      float x, y, z;
      std::memcpy(&x, &data[point_start + x_off], sizeof(float));
      std::memcpy(&y, &data[point_start + y_off], sizeof(float));
      std::memcpy(&z, &data[point_start + z_off], sizeof(float));
      const float range_sq = x * x + y * y;
      if (range_sq < 1e-6) continue;
      if (z < min_z || (max_z >= 0.0 && z > max_z)) continue;
      double angle = std::atan2(y, x); /* float overload (atan2f), widened */
      if (angle < 0.0) angle += two_pi;
      int bin = static_cast<int>((angle / two_pi) * num_bins);
      bin = std::min(bin, num_bins - 1);
      const double distance = std::sqrt(range_sq); /* float sqrt, widened */
      if (distance < ranges_out[bin]) ranges_out[bin] = distance;
    }
  }
}

/* ref: include/utils/pointcloud.h:116-177 (angle_step overload). Returns the number of bins
 * (ceil(2 pi / angle_step)); ranges_out / angles_out must hold that many. */
int32_t orc_pointcloud_to_laserscan_step(const int8_t *data, int64_t nbytes, int32_t point_step,
                                         int32_t row_step, int32_t height, int32_t width,
                                         int32_t x_off, int32_t y_off, int32_t z_off,
                                         double max_range, double min_z, double max_z,
                                         double angle_step, double *ranges_out, double *angles_out) {
  (void)width;
  const double two_pi = 2.0 * M_PI;
  const int num_bins = static_cast<int>(std::ceil(two_pi / angle_step));
  for (int i = 0; i < num_bins; ++i) {
    angles_out[i] = i * angle_step;
    ranges_out[i] = max_range;
  }
  if (point_step <= 0) return num_bins;
  for (int row = 0; row < height; ++row) {
    for (int col = 0; col < row_step; col += point_step) {
      const std::size_t point_start = (std::size_t)(row * row_step + col);
      const std::size_t max_offset =
          point_start + (std::size_t)std::max({x_off, y_off, z_off}) + sizeof(float);
      if (max_offset > (std::size_t)nbytes) continue;
      float x, y, z;
      std::memcpy(&x, &data[point_start + x_off], sizeof(float));
      std::memcpy(&y, &data[point_start + y_off], sizeof(float));
      std::memcpy(&z, &data[point_start + z_off], sizeof(float));
      const float range_sq = x * x + y * y;
      if (range_sq < 1e-6) continue;
      if (z < min_z || (max_z >= 0.0 && z > max_z)) continue;
      double angle = std::atan2(y, x); /* float overload (atan2f), widened */
      if (angle < 0.0) angle += two_pi;
      int bin = static_cast<int>(angle / angle_step);
      bin = std::min(bin, num_bins - 1);
      const double distance = std::sqrt(range_sq); /* float sqrt, widened */
      if (distance < ranges_out[bin]) ranges_out[bin] = distance;
    }
  }
  return num_bins;
}

/* ------------------------------------------------------------------------------------------
 * critical zone — ref: src/utils/critical_zone_check.cpp:13-131, include/utils/angles.h:21-29
 * ---------------------------------------------------------------------------------------- */
namespace {
struct CZ {
  double robotRadius;
  float critical_angle;
  orc::Iso3 sensor_tf_body;
  std::vector<float> cos_a, sin_a;
  std::vector<size_t> fwd, bwd;
};
CZ czInit(const orc_cz_cfg &c, const double *angles, int32_t n) {
  CZ z;
  if (c.robot_shape == ORC_CYLINDER)
    z.robotRadius = c.robot_dims[0];
  else if (c.robot_shape == ORC_BOX)
    z.robotRadius = std::sqrt(std::pow(c.robot_dims[0], 2) + std::pow(c.robot_dims[1], 2)) / 2;
  else
    z.robotRadius = c.robot_dims[0];
  z.sensor_tf_body = orc::makeTransform(quatOf(c.sensor_rotation), c.sensor_position);
  const float angle_rad = (float)(c.critical_angle * M_PI / 180.0);
  double a = std::fmod((double)(angle_rad / 2) + M_PI, 2 * M_PI);
  if (a < 0) a += 2 * M_PI;
  a -= M_PI;
  z.critical_angle = (float)a;
  z.cos_a.resize(n);
  z.sin_a.resize(n);
  for (int32_t i = 0; i < n; ++i) {
    z.cos_a[i] = (float)std::cos(angles[i]);
    z.sin_a[i] = (float)std::sin(angles[i]);
    const float p[3] = {z.cos_a[i], z.sin_a[i], 0.0f};
    float q[3];
    orc::apply(z.sensor_tf_body, p, q);
    const float abs_theta = std::abs(std::atan2(q[1], q[0]));
    if (abs_theta <= z.critical_angle) z.fwd.push_back((size_t)i);
    if (abs_theta >= M_PI - z.critical_angle) z.bwd.push_back((size_t)i);
  }
  return z;
}
float czCheck(const CZ &z, const orc_cz_cfg &c, const double *ranges, bool forward) {
  const std::vector<size_t> &ind = forward ? z.fwd : z.bwd;
  float slowdown_factor = 1.0f;
  for (size_t index : ind) {
    const float x = (float)(ranges[index] * z.cos_a[index]);
    const float y = (float)(ranges[index] * z.sin_a[index]);
    const float p[3] = {x, y, 0.0f};
    float q[3];
    orc::apply(z.sensor_tf_body, p, q);
    const float converted_range =
        (float)std::sqrt((double)q[1] * (double)q[1] + (double)q[0] * (double)q[0]);
    const float distance = (float)((double)converted_range - z.robotRadius);
    if (distance <= c.critical_distance) {
      return 0.0f;
    } else if (distance <= c.slowdown_distance) {
      slowdown_factor =
          std::min(slowdown_factor, (distance - c.critical_distance) /
                                        (c.slowdown_distance - c.critical_distance));
    }
  }
  return slowdown_factor;
}
} // namespace

float orc_cz_check_scan(const orc_cz_cfg *cfg, const double *angles, int32_t n_angles,
                        const double *ranges, int32_t forward) {
  const CZ z = czInit(*cfg, angles, n_angles);
  return czCheck(z, *cfg, ranges, forward != 0);
}

float orc_cz_check_cloud(const orc_cz_cfg *cfg, const double *angles, int32_t n_angles,
                         const int8_t *data, int64_t nbytes, int32_t point_step, int32_t row_step,
                         int32_t height, int32_t width, int32_t x_off, int32_t y_off,
                         int32_t z_off, int32_t forward) {
  const CZ z = czInit(*cfg, angles, n_angles);
  std::vector<double> ranges((size_t)n_angles);
  orc_pointcloud_to_laserscan(data, nbytes, point_step, row_step, height, width, x_off, y_off,
                              z_off, cfg->range_max, cfg->min_height, cfg->max_height, n_angles,
                              ranges.data());
  return czCheck(z, *cfg, ranges.data(), forward != 0);
}

int32_t orc_cz_indices(const orc_cz_cfg *cfg, const double *angles, int32_t n_angles,
                       int32_t forward, int32_t *idx_out) {
  const CZ z = czInit(*cfg, angles, n_angles);
  const std::vector<size_t> &ind = forward ? z.fwd : z.bwd;
  for (size_t i = 0; i < ind.size(); ++i) idx_out[i] = (int32_t)ind[i];
  return (int32_t)ind.size();
}

} /* extern "C" */
