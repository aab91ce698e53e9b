/*
 * ref_bridge.cpp — TEST INFRASTRUCTURE ONLY: the C entry points of kompass_oracle.h implemented by
 * calling the REFERENCE'S OWN classes, compiled from the reference's own sources where they lie under
 * /root/reference (oracle/Makefile target `_ref` -> oracle/_ref/libkompass_ref.so). No reference
 * source is copied into this repository: the Makefile hands g++ the paths of
 *   src/datatypes/path.cpp, src/utils/cost_evaluator.cpp, src/utils/trajectory_sampler.cpp,
 *   src/utils/collision_check.cpp, src/mapping/local_mapper.cpp, src/utils/critical_zone_check.cpp,
 *   src/controllers/{controller,follower,dwa}.cpp
 * together with -I oracle/shim, which holds stand-ins for the three third-party dependencies that
 * are absent from this image: Eigen (the API subset the reference uses, arithmetic of the published
 * Eigen 3.4 algorithms), FCL and octomap (the occupied-voxel model of voxel_model.h).
 *
 * What a port-vs-_ref comparison pins (tests/test_oracle_vs_ref.py): the reference's control flow,
 * operand widths and operation order for path prep, the sampler's enumeration / rollout / drop-pad
 * logic, all five cost terms and the argmin, the mapper (plain, Bayesian, warp), cloud binning, the
 * critical zone and the follower/DWA glue. What it does not pin: the Eigen kernels and the
 * collision query itself, which both arms share by construction.
 */
#include "kompass_oracle.h"

#include <chrono>
#include <cstring>
#include <memory>
#include <vector>

#include "controllers/dwa.h"
#include "datatypes/control.h"
#include "datatypes/path.h"
#include "datatypes/trajectory.h"
#include "mapping/local_mapper.h"
#include "utils/collision_check.h"
#include "utils/cost_evaluator.h"
#include "utils/critical_zone_check.h"
#include "utils/pointcloud.h"
#include "utils/trajectory_sampler.h"

using namespace Kompass;

namespace {

Control::ControlLimitsParams limitsOf(const orc_sampler_cfg &c) {
  Control::LinearVelocityControlParams x(c.vx_max, c.vx_acc, c.vx_dec), y(c.vy_max, c.vy_acc, c.vy_dec);
  Control::AngularVelocityControlParams a(M_PI, c.omega_max, c.omega_acc, c.omega_dec);
  return Control::ControlLimitsParams(x, y, a);
}

CollisionChecker::ShapeType shapeOf(int32_t s) {
  return s == ORC_CYLINDER ? CollisionChecker::ShapeType::CYLINDER
                           : (s == ORC_BOX ? CollisionChecker::ShapeType::BOX : CollisionChecker::ShapeType::SPHERE);
}

std::vector<float> dimsOf(int32_t shape, const float d[3]) {
  if (shape == ORC_CYLINDER) return {d[0], d[1]};
  if (shape == ORC_BOX) return {d[0], d[1], d[2]};
  return {d[0]};
}

Control::ControlType ctrlOf(int32_t t) {
  return t == ORC_ACKERMANN ? Control::ControlType::ACKERMANN
                            : (t == ORC_OMNI ? Control::ControlType::OMNI : Control::ControlType::DIFFERENTIAL_DRIVE);
}

std::unique_ptr<Control::TrajectorySampler> makeSampler(const orc_sampler_cfg &c) {
  Control::TrajectorySampler::TrajectorySamplerParameters p;
  p.setParameter("time_step", c.time_step);
  p.setParameter("prediction_horizon", c.prediction_horizon);
  p.setParameter("control_horizon", c.control_horizon);
  p.setParameter("max_linear_samples", (int)c.max_linear_samples);
  p.setParameter("max_angular_samples", (int)c.max_angular_samples);
  p.setParameter("octree_map_resolution", c.octree_resolution);
  p.setParameter("drop_samples", c.drop_samples != 0);
  const Eigen::Vector3f pos(c.sensor_position[0], c.sensor_position[1], c.sensor_position[2]);
  // Eigen::Quaternionf(w, x, y, z) from coefficients stored x, y, z, w
  const Eigen::Quaternionf rot(c.sensor_rotation[3], c.sensor_rotation[0], c.sensor_rotation[1], c.sensor_rotation[2]);
  return std::make_unique<Control::TrajectorySampler>(p, limitsOf(c), ctrlOf(c.control_type), shapeOf(c.robot_shape),
                                                      dimsOf(c.robot_shape, c.robot_dims), pos, rot,
                                                      std::max(1, (int)c.max_num_threads));
}

int32_t copySamples(const Control::TrajectorySamples2D &s, float *vx, float *vy, float *omega, float *x, float *y,
                    int32_t *slot_of_row, int32_t cap) {
  const int32_t n = (int32_t)s.size();
  const size_t P = s.numPointsPerTrajectory_;
  if (n > cap) return -1;
  for (int32_t i = 0; i < n; ++i) {
    for (size_t j = 0; j + 1 < P; ++j) {
      vx[(size_t)i * (P - 1) + j] = s.velocities.vx(i, j);
      vy[(size_t)i * (P - 1) + j] = s.velocities.vy(i, j);
      omega[(size_t)i * (P - 1) + j] = s.velocities.omega(i, j);
    }
    for (size_t j = 0; j < P; ++j) {
      x[(size_t)i * P + j] = s.paths.x(i, j);
      y[(size_t)i * P + j] = s.paths.y(i, j);
    }
    if (slot_of_row) slot_of_row[i] = -1;  // the reference does not keep the enumeration index of a row
  }
  return n;
}

Path::Path makeInterpolatedPath(const float *x, const float *y, int32_t n, double max_dist) {
  std::vector<Path::Point> pts;
  for (int32_t i = 0; i < n; ++i) pts.emplace_back(x[i], y[i], 0.0f);
  Path::Path p(pts);
  p.interpolate(max_dist, Path::InterpolationType::LINEAR);
  return p;
}

// a Path whose arrays are exactly the caller's (already interpolated) arrays: X/Y through the
// constructor, prefix lengths through interpolate()'s own bookkeeping would re-sample, so the cost
// evaluation below rebuilds the path from its ORIGINAL way points instead (see orc_ref_cost_evaluate_path)
struct CzAccess : public CriticalZoneChecker {
  using CriticalZoneChecker::CriticalZoneChecker;
  const std::vector<size_t> &fwd() const { return indicies_forward_; }
  const std::vector<size_t> &bwd() const { return indicies_backward_; }
};

std::unique_ptr<CzAccess> newCz(const orc_cz_cfg &c, bool cloud, const double *angles, int32_t n);

// checker objects are kept between calls with the same configuration and angles (a caller of the
// reference constructs the checker once), so that a timed call measures check() alone
CzAccess *makeCz(const orc_cz_cfg &c, bool cloud, const double *angles, int32_t n) {
  struct Slot {
    orc_cz_cfg cfg;
    bool cloud;
    std::vector<double> angles;
    std::unique_ptr<CzAccess> cz;
  };
  static thread_local Slot slots[2];
  Slot &s = slots[cloud ? 1 : 0];
  if (!s.cz || std::memcmp(&s.cfg, &c, sizeof(c)) != 0 || (int32_t)s.angles.size() != n ||
      std::memcmp(s.angles.data(), angles, sizeof(double) * (size_t)n) != 0) {
    s.cfg = c;
    s.cloud = cloud;
    s.angles.assign(angles, angles + n);
    s.cz = newCz(c, cloud, angles, n);
  }
  return s.cz.get();
}

std::unique_ptr<CzAccess> newCz(const orc_cz_cfg &c, bool cloud, const double *angles, int32_t n) {
  const Eigen::Vector3f pos(c.sensor_position[0], c.sensor_position[1], c.sensor_position[2]);
  const Eigen::Vector4f rot(c.sensor_rotation[0], c.sensor_rotation[1], c.sensor_rotation[2], c.sensor_rotation[3]);
  std::vector<double> a(angles, angles + n);
  return std::make_unique<CzAccess>(cloud ? CriticalZoneChecker::InputType::POINTCLOUD
                                          : CriticalZoneChecker::InputType::LASERSCAN,
                                    shapeOf(c.robot_shape), dimsOf(c.robot_shape, c.robot_dims), pos, rot,
                                    c.critical_angle, c.critical_distance, c.slowdown_distance, a, c.min_height,
                                    c.max_height, c.range_max);
}

struct MapperAccess : public Mapping::LocalMapper {
  using Mapping::LocalMapper::LocalMapper;
  Eigen::MatrixXf &prev() { return previousGridDataProb; }
};

}  // namespace

extern "C" {

const char *orc_backend(void) { return "reference sources (oracle/_ref)"; }

int64_t orc_num_trajectories(int32_t control_type, int32_t max_linear, int32_t max_angular) {
  return (int64_t)Control::getNumTrajectories(ctrlOf(control_type), max_linear, max_angular);
}
int64_t orc_num_points(double time_step, double prediction_horizon) {
  return (int64_t)Control::getNumPointsPerTrajectory(time_step, prediction_horizon);
}

int32_t orc_path_interpolate_linear(const float *x, const float *y, int32_t n, double max_dist, float *X, float *Y,
                                    float *acc, float *curv, int32_t cap, float *total_length) {
  Path::Path p = makeInterpolatedPath(x, y, n, max_dist);
  const int32_t m = (int32_t)p.getSize();
  if (m > cap) return -1;
  for (int32_t i = 0; i < m; ++i) {
    const Path::Point q = p.getIndex(i);
    X[i] = q.x();
    Y[i] = q.y();
    acc[i] = p.getDistanceAtIndex(i);
    curv[i] = (float)p.getCurvature(i);
  }
  if (total_length) *total_length = p.totalPathLength();
  return m;
}

/* Path::segment needs a Path object: rebuilt here from the ORIGINAL way points (x, y, n, max_dist) */
int32_t orc_ref_path_segment(const float *x, const float *y, int32_t n, double max_dist, double segment_length,
                             int64_t max_points_per_segment, int32_t *seg_starts, int32_t cap) {
  Path::Path p = makeInterpolatedPath(x, y, n, max_dist);
  p.segment(segment_length, (size_t)max_points_per_segment);
  const int32_t ns = (int32_t)p.getNumSegments();
  if (ns > cap) return -1;
  for (int32_t i = 0; i < ns; ++i) seg_starts[i] = (int32_t)p.getSegmentStartIndex(i);
  return ns;
}

float orc_ref_segment_length(const float *x, const float *y, int32_t n, double max_dist, int32_t start,
                             int32_t count) {
  Path::Path p = makeInterpolatedPath(x, y, n, max_dist);
  return p.getPart(start, start + count - 1).totalSegmentLength();
}

int32_t orc_sampler_generate_scan(const orc_sampler_cfg *cfg, const double vel[3], const double pose[3],
                                  const double *ranges, const double *angles, int32_t n, float *vx, float *vy,
                                  float *omega, float *x, float *y, int32_t *slot_of_row, int32_t cap) {
  auto s = makeSampler(*cfg);
  Control::LaserScan scan(std::vector<double>(ranges, ranges + n), std::vector<double>(angles, angles + n));
  auto out = s->generateTrajectories(Control::Velocity2D(vel[0], vel[1], vel[2]),
                                     Path::State(pose[0], pose[1], pose[2]), scan);
  return copySamples(*out, vx, vy, omega, x, y, slot_of_row, cap);
}

int32_t orc_sampler_generate_cloud(const orc_sampler_cfg *cfg, const double vel[3], const double pose[3],
                                   const float *xyz, int32_t n, float *vx, float *vy, float *omega, float *x,
                                   float *y, int32_t *slot_of_row, int32_t cap) {
  auto s = makeSampler(*cfg);
  std::vector<Path::Point> cloud;
  cloud.reserve(n);
  for (int32_t i = 0; i < n; ++i) cloud.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  auto out = s->generateTrajectories(Control::Velocity2D(vel[0], vel[1], vel[2]),
                                     Path::State(pose[0], pose[1], pose[2]), cloud);
  return copySamples(*out, vx, vy, omega, x, y, slot_of_row, cap);
}

int32_t orc_check_collision_states(const orc_sampler_cfg *cfg, const double sensor_pose[3], int32_t is_cloud,
                                   int32_t global_frame, const void *a, const void *b, int32_t n,
                                   const double *states, int32_t n_states, uint8_t *out) {
  const Eigen::Vector3f pos(cfg->sensor_position[0], cfg->sensor_position[1], cfg->sensor_position[2]);
  const Eigen::Quaternionf rot(cfg->sensor_rotation[3], cfg->sensor_rotation[0], cfg->sensor_rotation[1],
                               cfg->sensor_rotation[2]);
  CollisionChecker cc(shapeOf(cfg->robot_shape), dimsOf(cfg->robot_shape, cfg->robot_dims), pos, rot,
                      cfg->octree_resolution);
  cc.updateState(sensor_pose[0], sensor_pose[1], sensor_pose[2]);
  if (is_cloud) {
    const float *xyz = static_cast<const float *>(a);
    std::vector<Path::Point> cloud;
    for (int32_t i = 0; i < n; ++i) cloud.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    cc.updateSensorData(cloud, global_frame != 0);
  } else {
    const double *r = static_cast<const double *>(a), *g = static_cast<const double *>(b);
    cc.updateSensorData(Control::LaserScan(std::vector<double>(r, r + n), std::vector<double>(g, g + n)));
  }
  int32_t any = 0;
  for (int32_t i = 0; i < n_states; ++i) {
    const bool hit = cc.checkCollisions(Path::State(states[3 * i], states[3 * i + 1], states[3 * i + 2]));
    if (out) out[i] = hit ? 1 : 0;
    any |= hit ? 1 : 0;
  }
  return any;
}

/* CostEvaluator::setPointScan + getMinTrajectoryCost through the reference class. The reference only
 * returns the argmin, so per-trajectory totals come from evaluating every trajectory as a batch of
 * one (same code path, same arithmetic); the winner comes from the whole batch. The reference path is
 * rebuilt from its ORIGINAL way points (wx, wy, n_way) with Path::interpolate(max_dist). sensor data:
 * is_cloud = 1: a = xyz floats; 0: a = ranges, b = angles (doubles); n_obs = 0: setPointScan is not
 * called at all (the obstacle term is skipped, cost_evaluator.cpp:76). */
int32_t orc_ref_cost_evaluate(const orc_cost_cfg *cfg, int32_t n_traj, int32_t P, const float *vx, const float *vy,
                              const float *omega, const float *x, const float *y, const float *wx, const float *wy,
                              int32_t n_way, double max_dist, int32_t seg_start, int32_t seg_count,
                              int32_t is_cloud, const void *a, const void *b, int32_t n_obs, const double pose[3],
                              float max_sensor_range, float *costs_out, int32_t *best_idx, float *best_cost) {
  Path::Path path = makeInterpolatedPath(wx, wy, n_way, max_dist);
  Control::CostEvaluator::TrajectoryCostsWeights w;
  w.setParameter("reference_path_distance_weight", cfg->w_path);
  w.setParameter("goal_distance_weight", cfg->w_goal);
  w.setParameter("obstacles_distance_weight", cfg->w_obstacles);
  w.setParameter("smoothness_weight", cfg->w_smooth);
  w.setParameter("jerk_weight", cfg->w_jerk);
  Control::LinearVelocityControlParams lx(1.0, cfg->acc_limits[0], cfg->acc_limits[0]),
      ly(1.0, cfg->acc_limits[1], cfg->acc_limits[1]);
  Control::AngularVelocityControlParams la(M_PI, 1.0, cfg->acc_limits[2], cfg->acc_limits[2]);
  Control::ControlLimitsParams lim(lx, ly, la);
  const Eigen::Vector3f pos(cfg->sensor_position[0], cfg->sensor_position[1], cfg->sensor_position[2]);
  const Eigen::Quaternionf rot(cfg->sensor_rotation[3], cfg->sensor_rotation[0], cfg->sensor_rotation[1],
                               cfg->sensor_rotation[2]);
  Control::CostEvaluator ev(w, pos, rot, lim, (size_t)n_traj, (size_t)P, (size_t)seg_count);
  if (n_obs > 0) {
    const Path::State st(pose[0], pose[1], pose[2]);
    if (is_cloud) {
      const float *xyz = static_cast<const float *>(a);
      std::vector<Path::Point> cloud;
      for (int32_t i = 0; i < n_obs; ++i) cloud.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      ev.setPointScan(cloud, st, max_sensor_range);
    } else {
      const double *r = static_cast<const double *>(a), *g = static_cast<const double *>(b);
      ev.setPointScan(Control::LaserScan(std::vector<double>(r, r + n_obs), std::vector<double>(g, g + n_obs)), st,
                      max_sensor_range);
    }
  }
  const Path::Path::View view = path.getPart(seg_start, seg_start + seg_count - 1);
  auto rowOf = [&](int32_t i, Control::TrajectoryVelocities2D &v, Control::TrajectoryPath &p) {
    for (int32_t j = 0; j + 1 < P; ++j)
      v.add(j, vx[(size_t)i * (P - 1) + j], vy[(size_t)i * (P - 1) + j], omega[(size_t)i * (P - 1) + j]);
    for (int32_t j = 0; j < P; ++j) p.add(j, x[(size_t)i * P + j], y[(size_t)i * P + j], 0.0f);
  };
  auto all = std::make_unique<Control::TrajectorySamples2D>((size_t)n_traj, (size_t)P);
  for (int32_t i = 0; i < n_traj; ++i) {
    Control::TrajectoryVelocities2D v((size_t)P);
    Control::TrajectoryPath p((size_t)P);
    rowOf(i, v, p);
    all->push_back(v, p);
    if (costs_out) {
      auto one = std::make_unique<Control::TrajectorySamples2D>((size_t)1, (size_t)P);
      one->push_back(v, p);
      costs_out[i] = ev.getMinTrajectoryCost(one, &path, view).trajCost;
    }
  }
  const Control::TrajSearchResult res = ev.getMinTrajectoryCost(all, &path, view);
  int32_t idx = -1;
  if (res.isTrajFound)  // the reference returns the winning trajectory, not its index: match the rows
    for (int32_t i = 0; i < n_traj && idx < 0; ++i) {
      bool same = true;
      for (int32_t j = 0; j < P && same; ++j)
        same = res.trajectory.path.x(j) == x[(size_t)i * P + j] && res.trajectory.path.y(j) == y[(size_t)i * P + j];
      for (int32_t j = 0; j + 1 < P && same; ++j)
        same = res.trajectory.velocities.vx(j) == vx[(size_t)i * (P - 1) + j] &&
               res.trajectory.velocities.vy(j) == vy[(size_t)i * (P - 1) + j] &&
               res.trajectory.velocities.omega(j) == omega[(size_t)i * (P - 1) + j];
      if (same) idx = i;
    }
  if (best_idx) *best_idx = idx;
  if (best_cost) *best_cost = res.trajCost;
  return res.isTrajFound ? 1 : 0;
}

/* One DWA cycle of the reference as DWA::findBestPath runs it (dwa.h:183-230) on its own classes, timed
 * inside: generateTrajectories (the sampler's ThreadPool when cfg->max_num_threads > 1) -> setPointScan ->
 * getMinTrajectoryCost (single-threaded in the reference, cost_evaluator.cpp:49-109) over the first
 * `max_traj` admissible samples (<= 0: all). times_s = {sampler, setPointScan, cost evaluation}.
 * Returns the admissible count. bench.py --impl reference. */
int32_t orc_ref_cycle_cloud(const orc_sampler_cfg *cfg, const orc_cost_cfg *ccfg, const float *wx, const float *wy,
                            int32_t n_way, double max_dist, int32_t seg_start, int32_t seg_count, const double vel[3],
                            const double pose[3], const float *xyz, int32_t n, float max_sensor_range,
                            int32_t max_traj, double times_s[3], int32_t *n_points, int32_t *evaluated,
                            int32_t *found, float *best_cost) {
  Path::Path path = makeInterpolatedPath(wx, wy, n_way, max_dist);
  auto sampler = makeSampler(*cfg);
  Control::CostEvaluator::TrajectoryCostsWeights w;
  w.setParameter("reference_path_distance_weight", ccfg->w_path);
  w.setParameter("goal_distance_weight", ccfg->w_goal);
  w.setParameter("obstacles_distance_weight", ccfg->w_obstacles);
  w.setParameter("smoothness_weight", ccfg->w_smooth);
  w.setParameter("jerk_weight", ccfg->w_jerk);
  const Eigen::Vector3f pos(ccfg->sensor_position[0], ccfg->sensor_position[1], ccfg->sensor_position[2]);
  const Eigen::Quaternionf rot(ccfg->sensor_rotation[3], ccfg->sensor_rotation[0], ccfg->sensor_rotation[1],
                               ccfg->sensor_rotation[2]);
  Control::CostEvaluator ev(w, pos, rot, limitsOf(*cfg), sampler->numTrajectories, sampler->numPointsPerTrajectory,
                            (size_t)seg_count);
  std::vector<Path::Point> cloud;
  cloud.reserve(n);
  for (int32_t i = 0; i < n; ++i) cloud.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  const Path::State st(pose[0], pose[1], pose[2]);
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double>(b - a).count();
  };
  auto t0 = now();
  std::unique_ptr<Control::TrajectorySamples2D> samples =
      sampler->generateTrajectories(Control::Velocity2D(vel[0], vel[1], vel[2]), st, cloud);
  auto t1 = now();
  const int32_t n_adm = (int32_t)samples->size();
  const size_t P = samples->numPointsPerTrajectory_;
  if (n_points) *n_points = (int32_t)P;
  times_s[0] = secs(t0, t1);
  times_s[1] = times_s[2] = 0.0;
  if (evaluated) *evaluated = 0;
  if (found) *found = 0;
  if (n_adm == 0) return 0;
  t0 = now();
  if (n > 0) ev.setPointScan(cloud, st, max_sensor_range);
  t1 = now();
  times_s[1] = secs(t0, t1);
  const Path::Path::View view = path.getPart(seg_start, seg_start + seg_count - 1);
  int32_t m = n_adm;
  if (max_traj > 0 && max_traj < n_adm) {  // bounded sample: the first m admissible rows (copied outside the clock)
    m = max_traj;
    auto sub = std::make_unique<Control::TrajectorySamples2D>((size_t)m, P);
    for (int32_t i = 0; i < m; ++i) {
      Control::Trajectory2D t = samples->getIndex(i);
      sub->push_back(t.velocities, t.path);
    }
    samples = std::move(sub);
  }
  t0 = now();
  const Control::TrajSearchResult res = ev.getMinTrajectoryCost(samples, &path, view);
  t1 = now();
  times_s[2] = secs(t0, t1);
  if (evaluated) *evaluated = m;
  if (found) *found = res.isTrajFound ? 1 : 0;
  if (best_cost) *best_cost = res.trajCost;
  return n_adm;
}

void orc_mapper_scan_to_grid(int32_t H, int32_t W, float resolution, const float laser_pos[3],
                             float laser_orientation, const double *angles, const double *ranges, int32_t n,
                             int32_t *grid) {
  // the mapper object is kept between calls with the same geometry (as a caller of the reference
  // would), so that a timed call measures scanToGrid and not the constructor's allocations
  struct Key {
    int32_t H, W;
    float res, px, py, pz, orient;
    bool operator==(const Key &o) const {
      return H == o.H && W == o.W && res == o.res && px == o.px && py == o.py && pz == o.pz && orient == o.orient;
    }
  };
  static thread_local Key key{};
  static thread_local std::unique_ptr<Mapping::LocalMapper> m;
  const Key k{H, W, resolution, laser_pos[0], laser_pos[1], laser_pos[2], laser_orientation};
  if (!m || !(k == key)) {
    m = std::make_unique<Mapping::LocalMapper>(H, W, resolution,
                                               Eigen::Vector3f(laser_pos[0], laser_pos[1], laser_pos[2]),
                                               laser_orientation, false, n, 0.01f, 2.0f, 0.0f, 20.0f, 10000);
    key = k;
  }
  Eigen::MatrixXi &g = m->scanToGrid(std::vector<double>(angles, angles + n), std::vector<double>(ranges, ranges + n));
  std::memcpy(grid, g.data(), sizeof(int32_t) * (size_t)H * W);
}

void orc_mapper_scan_to_grid_bayes(int32_t H, int32_t W, float resolution, const float laser_pos[3],
                                   float laser_orientation, float pPrior, float pOccupied, float pEmpty,
                                   float rangeSure, float rangeMax, float wallSize, const double *angles,
                                   const double *ranges, int32_t n, const float *prev, int32_t *grid, float *prob) {
  MapperAccess m(H, W, resolution, Eigen::Vector3f(laser_pos[0], laser_pos[1], laser_pos[2]), laser_orientation,
                 false, n, pPrior, pOccupied, pEmpty, rangeSure, rangeMax, wallSize, 0.01f, 2.0f, 0.0f, 10000);
  if (prev) std::memcpy(m.prev().data(), prev, sizeof(float) * (size_t)H * W);
  auto out = m.scanToGridBaysian(std::vector<double>(angles, angles + n), std::vector<double>(ranges, ranges + n));
  std::memcpy(grid, std::get<0>(out).data(), sizeof(int32_t) * (size_t)H * W);
  std::memcpy(prob, std::get<1>(out).data(), sizeof(float) * (size_t)H * W);
}

void orc_mapper_warp_previous(int32_t H, int32_t W, float resolution, float pPrior, float pos_x, float pos_y,
                              double orientation, const float *prev, float *out) {
  MapperAccess m(H, W, resolution, Eigen::Vector3f(0.0f, 0.0f, 0.0f), 0.0f, false, 1, pPrior, 0.6f, 0.4f, 1.0f, 20.0f,
                 0.2f, 0.01f, 2.0f, 0.0f, 10000);
  std::memcpy(m.prev().data(), prev, sizeof(float) * (size_t)H * W);
  m.getPreviousGridInCurrentPose(Eigen::Vector2f(pos_x, pos_y), orientation);
  std::memcpy(out, m.prev().data(), sizeof(float) * (size_t)H * W);
}

void orc_pointcloud_to_laserscan(const int8_t *data, int64_t nbytes, int32_t point_step, int32_t row_step,
                                 int32_t height, int32_t width, int32_t x_off, int32_t y_off, int32_t z_off,
                                 double max_range, double min_z, double max_z, int32_t num_bins, double *ranges_out) {
  std::vector<int8_t> d(data, data + nbytes);
  std::vector<double> r;
  pointCloudToLaserScanFromRaw(d, point_step, row_step, height, width, x_off, y_off, z_off, max_range, min_z, max_z,
                               (int)num_bins, r);
  std::memcpy(ranges_out, r.data(), sizeof(double) * (size_t)num_bins);
}

int32_t orc_pointcloud_to_laserscan_step(const int8_t *data, int64_t nbytes, int32_t point_step, int32_t row_step,
                                         int32_t height, int32_t width, int32_t x_off, int32_t y_off, int32_t z_off,
                                         double max_range, double min_z, double max_z, double angle_step,
                                         double *ranges_out, double *angles_out) {
  std::vector<int8_t> d(data, data + nbytes);
  std::vector<double> r, a;
  pointCloudToLaserScanFromRaw(d, point_step, row_step, height, width, x_off, y_off, z_off, max_range, min_z, max_z,
                               angle_step, r, a);
  std::memcpy(ranges_out, r.data(), sizeof(double) * r.size());
  std::memcpy(angles_out, a.data(), sizeof(double) * a.size());
  return (int32_t)r.size();
}

float orc_cz_check_scan(const orc_cz_cfg *cfg, const double *angles, int32_t n_angles, const double *ranges,
                        int32_t forward) {
  auto cz = makeCz(*cfg, false, angles, n_angles);
  return cz->check(std::vector<double>(ranges, ranges + n_angles), forward != 0);
}

float orc_cz_check_cloud(const orc_cz_cfg *cfg, const double *angles, int32_t n_angles, const int8_t *data,
                         int64_t nbytes, int32_t point_step, int32_t row_step, int32_t height, int32_t width,
                         int32_t x_off, int32_t y_off, int32_t z_off, int32_t forward) {
  auto cz = makeCz(*cfg, true, angles, n_angles);
  return cz->check(std::vector<int8_t>(data, data + nbytes), point_step, row_step, height, width, x_off, y_off,
                   z_off, forward != 0);
}

int32_t orc_cz_indices(const orc_cz_cfg *cfg, const double *angles, int32_t n_angles, int32_t forward,
                       int32_t *idx_out) {
  auto cz = makeCz(*cfg, false, angles, n_angles);
  const std::vector<size_t> &v = forward ? cz->fwd() : cz->bwd();
  for (size_t i = 0; i < v.size(); ++i) idx_out[i] = (int32_t)v[i];
  return (int32_t)v.size();
}

}  // extern "C"
