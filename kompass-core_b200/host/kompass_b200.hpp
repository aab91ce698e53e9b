// kompass_b200.hpp — header-only C++17 host classes over the C-ABI (include/kompass_b200.h).
//
// They mirror the reference's operator interface for the hot path — same class names, method
// names, argument order/meaning and exception behaviour — so the reference's own call sites
// (DWA::findBestPath, the Python wrappers, the Boost tests, benchmark_runner.cpp) compile against
// them with only the include swapped:
//   Kompass::Control::TrajectorySampler   ref: include/utils/trajectory_sampler.h:20-239
//   Kompass::Control::CostEvaluator       ref: include/utils/cost_evaluator.h:20-433
//   Kompass::Control::DWA                 ref: include/controllers/dwa.h:22-261 (hot-path subset)
//   Kompass::Mapping::LocalMapperGPU      ref: include/mapping/local_mapper_gpu.h:12-147
//   Kompass::CriticalZoneCheckerGPU       ref: include/utils/critical_zone_check_gpu.h:17-192
//
// Eigen is not available in this image, so fixed-size Eigen arguments are std::array and the
// dynamic matrices are thin row-/column-major views (`MatrixXfR`, `MatrixXi`) with operator()(i,j),
// rows(), cols(), data(). With Eigen present a maintainer maps them 1:1.
// Errors: KC_ERR_INVALID_ARG -> std::invalid_argument, KC_ERR_OUT_OF_RANGE -> std::out_of_range,
// everything else -> std::runtime_error (ref error conventions: dwa.h:187-191, path.cpp:61-66,
// parameter.h:134-146). There is no CPU fallback behind these classes.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <exception>
#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/kompass_b200.h"

namespace Kompass {

using Vector3f = std::array<float, 3>;
using Vector4f = std::array<float, 4>;  // quaternion coefficients (x, y, z, w) as Eigen::Vector4f

inline void kcThrow(int32_t rc) {
  if (rc == KC_OK) return;
  const std::string msg = kc_last_error();
  if (rc == KC_ERR_INVALID_ARG) throw std::invalid_argument(msg);
  if (rc == KC_ERR_OUT_OF_RANGE) throw std::out_of_range(msg);
  throw std::runtime_error(msg);
}

// ref: src/utils/gpu_check.cpp:7-22
inline std::string getAvailableAccelerators() {
  char buf[4096];
  kc_available_accelerators(buf, sizeof(buf));
  return buf;
}

struct LaserScanView {  // Control::LaserScan is declared further down; the checker only needs the arrays
  const std::vector<double> &ranges, &angles;
};

// ref: include/utils/collision_check.h:23-180
class CollisionChecker {
public:
  enum class ShapeType { CYLINDER = 0, BOX = 1, SPHERE = 2 };  // collision_check.h:25

  CollisionChecker(const ShapeType robot_shape_type, const std::vector<float> &robot_dimensions,
                   const Vector3f &sensor_position_body, const Vector4f &sensor_rotation_body,
                   const double octree_resolution = 0.01) {
    kc_collision_config c{};
    c.robot_shape = static_cast<int32_t>(robot_shape_type);
    for (size_t i = 0; i < 3; ++i) c.robot_dims[i] = i < robot_dimensions.size() ? robot_dimensions[i] : 0.0f;
    for (int i = 0; i < 3; ++i) c.sensor_position[i] = sensor_position_body[i];
    for (int i = 0; i < 4; ++i) c.sensor_rotation[i] = sensor_rotation_body[i];
    c.octree_resolution = octree_resolution;
    kcThrow(kc_collision_create(&c, &h_));
  }
  ~CollisionChecker() { kc_collision_destroy(h_); }
  CollisionChecker(const CollisionChecker &) = delete;
  CollisionChecker &operator=(const CollisionChecker &) = delete;

  void resetOctreeResolution(const double resolution) { kcThrow(kc_collision_reset_octree_resolution(h_, resolution)); }
  float getRadius() const { return kc_collision_get_radius(h_); }
  void updateState(const double x, const double y, const double yaw) { kcThrow(kc_collision_update_state(h_, x, y, yaw)); }
  template <class State>
  void updateState(const State &s) { updateState(s.x, s.y, s.yaw); }

  // updateSensorData<T>(data, global_frame): LaserScan-like {ranges, angles} or a vector of points
  template <class Scan>
  auto updateSensorData(const Scan &scan, const bool /*global_frame*/ = true) -> decltype(scan.ranges, void()) {
    if (scan.ranges.size() != scan.angles.size())
      throw std::invalid_argument("LaserScan ranges and angles must have the same size");
    kcThrow(kc_collision_update_scan(h_, scan.ranges.data(), scan.angles.data(),
                                     static_cast<int32_t>(scan.ranges.size())));
  }
  void updateSensorData(const std::vector<std::array<float, 3>> &cloud, const bool global_frame = true) {
    static_assert(sizeof(std::array<float, 3>) == 12, "packed points");
    kcThrow(kc_collision_update_cloud(h_, cloud.empty() ? nullptr : cloud.front().data(),
                                      static_cast<int32_t>(cloud.size()), global_frame ? 1 : 0));
  }
  bool checkCollisions() {
    int32_t r = 0;
    kcThrow(kc_collision_check(h_, &r));
    return r != 0;
  }
  template <class State>
  auto checkCollisions(const State &s) -> decltype(s.yaw, bool()) {
    const double st[3] = {s.x, s.y, s.yaw};
    int32_t any = 0;
    kcThrow(kc_collision_check_states(h_, st, 1, nullptr, &any));
    return any != 0;
  }
  bool checkCollisions(const std::vector<double> &ranges, const std::vector<double> &angles,
                       double /*height*/ = 0.1) {
    updateSensorData(LaserScanView{ranges, angles});
    return checkCollisions();
  }
  // batched checkCollisions(state) (TrajectorySampler::checkStatesFeasibility): true if any collides
  template <class State>
  bool checkStates(const std::vector<State> &states, std::vector<uint8_t> *per_state = nullptr) {
    std::vector<double> st(states.size() * 3);
    for (size_t i = 0; i < states.size(); ++i) {
      st[3 * i] = states[i].x;
      st[3 * i + 1] = states[i].y;
      st[3 * i + 2] = states[i].yaw;
    }
    if (per_state) per_state->assign(states.size(), 0);
    int32_t any = 0;
    kcThrow(kc_collision_check_states(h_, st.data(), static_cast<int32_t>(states.size()),
                                      per_state ? per_state->data() : nullptr, &any));
    return any != 0;
  }

private:
  kc_collision *h_ = nullptr;
};

// row-major float matrix (ref: trajectory.h:53-54 MatrixXfR)
struct MatrixXfR {
  std::vector<float> v;
  size_t r = 0, c = 0;
  MatrixXfR() = default;
  MatrixXfR(size_t rows, size_t cols) : v(rows * cols), r(rows), c(cols) {}
  float &operator()(size_t i, size_t j) { return v[i * c + j]; }
  float operator()(size_t i, size_t j) const { return v[i * c + j]; }
  size_t rows() const { return r; }
  size_t cols() const { return c; }
  float *data() { return v.data(); }
  const float *data() const { return v.data(); }
};

// column-major int matrix (Eigen::MatrixXi layout: (i,j) at i + j*rows)
struct MatrixXi {
  std::vector<int32_t> v;
  size_t r = 0, c = 0;
  MatrixXi() = default;
  MatrixXi(size_t rows, size_t cols) : v(rows * cols), r(rows), c(cols) {}
  int32_t &operator()(size_t i, size_t j) { return v[i + j * r]; }
  int32_t operator()(size_t i, size_t j) const { return v[i + j * r]; }
  size_t rows() const { return r; }
  size_t cols() const { return c; }
  int32_t *data() { return v.data(); }
  const int32_t *data() const { return v.data(); }
};

// column-major float matrix (Eigen::MatrixXf layout)
struct MatrixXf {
  std::vector<float> v;
  size_t r = 0, c = 0;
  MatrixXf() = default;
  MatrixXf(size_t rows, size_t cols) : v(rows * cols), r(rows), c(cols) {}
  float &operator()(size_t i, size_t j) { return v[i + j * r]; }
  float operator()(size_t i, size_t j) const { return v[i + j * r]; }
  size_t rows() const { return r; }
  size_t cols() const { return c; }
  float *data() { return v.data(); }
  const float *data() const { return v.data(); }
};

namespace Control {

enum class ControlType { ACKERMANN = 0, DIFFERENTIAL_DRIVE = 1, OMNI = 2 };  // control.h:14

// ref: include/datatypes/control.h:181-232
struct LinearVelocityControlParams {
  double maxVel = 1.0, maxAcceleration = 10.0, maxDeceleration = 10.0;
  LinearVelocityControlParams(double v = 1.0, double a = 10.0, double d = 10.0)
      : maxVel(v), maxAcceleration(a), maxDeceleration(d) {}
};
struct AngularVelocityControlParams {
  double maxAngle = M_PI, maxOmega = 1.0, maxAcceleration = 10.0, maxDeceleration = 10.0;
  AngularVelocityControlParams(double ang = M_PI, double om = 1.0, double a = 10.0, double d = 10.0)
      : maxAngle(ang), maxOmega(om), maxAcceleration(a), maxDeceleration(d) {}
};
struct ControlLimitsParams {
  LinearVelocityControlParams velXParams, velYParams;
  AngularVelocityControlParams omegaParams;
  ControlLimitsParams() = default;
  ControlLimitsParams(const LinearVelocityControlParams &x, const LinearVelocityControlParams &y,
                      const AngularVelocityControlParams &o)
      : velXParams(x), velYParams(y), omegaParams(o) {}
};

// ref: control.h:112-140
class Velocity2D {
public:
  Velocity2D() = default;
  Velocity2D(double vx, double vy, double omega, double steer = 0.0) : v_{vx, vy, omega, steer} {}
  double vx() const { return v_[0]; }
  double vy() const { return v_[1]; }
  double omega() const { return v_[2]; }
  double steer_ang() const { return v_[3]; }
  void setVx(double v) { v_[0] = v; }
  void setVy(double v) { v_[1] = v; }
  void setOmega(double v) { v_[2] = v; }
  void setSteerAng(double v) { v_[3] = v; }

private:
  std::array<double, 4> v_{0, 0, 0, 0};
};

// ref: control.h:237-243
struct LaserScan {
  std::vector<double> ranges, angles;
  LaserScan(std::vector<double> r, std::vector<double> a) : ranges(std::move(r)), angles(std::move(a)) {}
};

}  // namespace Control
}  // namespace Kompass

namespace Path {

// ref: include/datatypes/path.h:14-31
struct State {
  double x, y, yaw, speed;
  State(double px = 0.0, double py = 0.0, double pyaw = 0.0, double s = 0.0)
      : x(px), y(py), yaw(pyaw), speed(s) {}
};
using Point = std::array<float, 3>;  // Eigen::Vector3f

// Reference path. Two ways in, as in the reference (path.h:112-125):
//  * Path(points): raw way-points; DWA::setCurrentPath interpolates and segments them
//    (follower.cpp:81-107) through kc_dwa_set_current_path;
//  * Path(X, Y, accumulated, total): arrays Path::interpolate already produced (X_, Y_,
//    accumulated_path_length_, total length: path.h:287-297), used as they are.
struct Path {
  struct View {  // ref: path.h:39-61
    const Path *parent = nullptr;
    size_t start_idx_ = 0, length = 0;
    size_t getSize() const { return length; }
    size_t getStartIndex() const { return start_idx_; }
    Point getIndex(size_t i) const { return {parent->X[start_idx_ + i], parent->Y[start_idx_ + i], 0.0f}; }
  };
  std::vector<float> X, Y, accumulated;
  float total_length = 0.0f;
  bool prepared = false;  // true: arrays are already interpolated (second constructor)
  Path() = default;
  explicit Path(const std::vector<Point> &points) {  // ref: path.cpp:12-26
    if (points.size() < 2) throw std::invalid_argument("At least two points are required to create a path.");
    for (const Point &p : points) {
      X.push_back(p[0]);
      Y.push_back(p[1]);
    }
  }
  Path(std::vector<float> x, std::vector<float> y, std::vector<float> acc, float total)
      : X(std::move(x)), Y(std::move(y)), accumulated(std::move(acc)), total_length(total), prepared(true) {
    if (X.size() < 2) throw std::invalid_argument("At least two points are required to create a path.");
    if (X.size() != Y.size() || X.size() != accumulated.size())
      throw std::invalid_argument("X, Y and accumulated-length vectors must have the same size.");
  }
  size_t getSize() const { return X.size(); }
  const std::vector<float> &getX() const { return X; }
  const std::vector<float> &getY() const { return Y; }
  Point getIndex(size_t i) const { return {X.at(i), Y.at(i), 0.0f}; }
  float totalPathLength() const {  // ref: path.cpp:150-165 (sum of way-point distances until interpolated)
    if (prepared || X.size() < 2) return total_length;
    float t = 0.0f;
    for (size_t i = 1; i < X.size(); ++i) {
      const float dx = X[i - 1] - X[i], dy = Y[i - 1] - Y[i];
      t += std::sqrt(dx * dx + (dy * dy + 0.0f));
    }
    return t;
  }
  View getPart(size_t start, size_t end) const {  // ref: path.cpp:80-91
    if (start >= X.size() || end >= X.size() || start > end)
      throw std::out_of_range("Invalid range for path part. Maximum path size is " +
                              std::to_string(X.size()) + ", but requested part start= " +
                              std::to_string(start) + ", and requested end= " + std::to_string(end));
    return View{this, start, end - start + 1};
  }
};

}  // namespace Path

namespace Kompass {
namespace Control {

// ref: include/datatypes/trajectory.h:57-323
struct TrajectoryVelocities2D {
  std::vector<float> vx, vy, omega;
  size_t numPointsPerTrajectory_ = 0;
  Velocity2D getFront() const { return Velocity2D(vx.at(0), vy.at(0), omega.at(0)); }
};
struct TrajectoryPath {
  std::vector<float> x, y, z;
  size_t numPointsPerTrajectory_ = 0;
  ::Path::Point getEnd() const { return {x.back(), y.back(), z.back()}; }
};
struct Trajectory2D {
  TrajectoryVelocities2D velocities;
  TrajectoryPath path;
  size_t numPointsPerTrajectory_ = 0;
};
struct TrajSearchResult {  // trajectory.h:611-618
  Trajectory2D trajectory;
  bool isTrajFound = false;
  float trajCost = 0.0f;
};

// ref: trajectory.h:326-603 (SoA batches, row-major float)
struct TrajectoryVelocitySamples2D {
  MatrixXfR vx, vy, omega;
};
struct TrajectoryPathSamples {
  MatrixXfR x, y, z;
};
struct TrajectorySamples2D {
  TrajectoryVelocitySamples2D velocities;
  TrajectoryPathSamples paths;
  size_t maxNumTrajectories_ = 0, numPointsPerTrajectory_ = 0;
  size_t count = 0;
  std::vector<int32_t> slots;  // enumeration index of each row (extra, for diagnostics)
  TrajectorySamples2D() = default;
  TrajectorySamples2D(size_t maxN, size_t P) : maxNumTrajectories_(maxN), numPointsPerTrajectory_(P) {
    velocities.vx = velocities.vy = velocities.omega = MatrixXfR(maxN, P - 1);
    paths.x = paths.y = paths.z = MatrixXfR(maxN, P);
  }
  size_t size() const { return count; }
  void push_back(const TrajectoryVelocities2D &v, const TrajectoryPath &p) {
    const size_t P = numPointsPerTrajectory_;
    for (size_t j = 0; j + 1 < P; ++j) {
      velocities.vx(count, j) = v.vx[j];
      velocities.vy(count, j) = v.vy[j];
      velocities.omega(count, j) = v.omega[j];
    }
    for (size_t j = 0; j < P; ++j) {
      paths.x(count, j) = p.x[j];
      paths.y(count, j) = p.y[j];
      paths.z(count, j) = p.z.empty() ? 0.0f : p.z[j];
    }
    ++count;
  }
  Trajectory2D getIndex(size_t i) const {
    const size_t P = numPointsPerTrajectory_;
    Trajectory2D t;
    t.numPointsPerTrajectory_ = t.velocities.numPointsPerTrajectory_ = t.path.numPointsPerTrajectory_ = P;
    t.velocities.vx.assign(&velocities.vx.v[i * (P - 1)], &velocities.vx.v[i * (P - 1)] + (P - 1));
    t.velocities.vy.assign(&velocities.vy.v[i * (P - 1)], &velocities.vy.v[i * (P - 1)] + (P - 1));
    t.velocities.omega.assign(&velocities.omega.v[i * (P - 1)], &velocities.omega.v[i * (P - 1)] + (P - 1));
    t.path.x.assign(&paths.x.v[i * P], &paths.x.v[i * P] + P);
    t.path.y.assign(&paths.y.v[i * P], &paths.y.v[i * P] + P);
    t.path.z.assign(P, 0.0f);
    return t;
  }
};

namespace detail {
inline kc_planner_config makeConfig(const ControlLimitsParams &lim, ControlType type, double timeStep,
                                    double predictionHorizon, double controlHorizon, int maxLin,
                                    int maxAng, CollisionChecker::ShapeType shape,
                                    const std::vector<float> &dims, const Vector3f &pos,
                                    const Vector4f &rot, double octreeRes, int maxNumThreads) {
  kc_planner_config c{};
  c.control_type = static_cast<int32_t>(type);
  c.time_step = timeStep;
  c.prediction_horizon = predictionHorizon;
  c.control_horizon = controlHorizon;
  c.max_linear_samples = maxLin;
  c.max_angular_samples = maxAng;
  c.vx_max = lim.velXParams.maxVel;
  c.vx_acc = lim.velXParams.maxAcceleration;
  c.vx_dec = lim.velXParams.maxDeceleration;
  c.vy_max = lim.velYParams.maxVel;
  c.vy_acc = lim.velYParams.maxAcceleration;
  c.vy_dec = lim.velYParams.maxDeceleration;
  c.omega_max = lim.omegaParams.maxOmega;
  c.omega_acc = lim.omegaParams.maxAcceleration;
  c.omega_dec = lim.omegaParams.maxDeceleration;
  c.robot_shape = static_cast<int32_t>(shape);
  for (size_t i = 0; i < 3; ++i) c.robot_dims[i] = i < dims.size() ? dims[i] : 0.0f;
  for (int i = 0; i < 3; ++i) c.sensor_position[i] = pos[i];
  for (int i = 0; i < 4; ++i) c.sensor_rotation[i] = rot[i];
  c.octree_resolution = octreeRes;
  c.drop_samples = 1;
  c.num_ctrl_points = -1;
  c.w_path = c.w_goal = c.w_obstacles = c.w_smooth = c.w_jerk = 1.0;
  c.max_local_range = 10.0f;
  c.max_num_threads = maxNumThreads;
  return c;
}
struct PlannerHandle {
  kc_planner *h = nullptr;
  explicit PlannerHandle(const kc_planner_config &c) { kcThrow(kc_planner_create(&c, &h)); }
  ~PlannerHandle() { kc_planner_destroy(h); }
  PlannerHandle(const PlannerHandle &) = delete;
  PlannerHandle &operator=(const PlannerHandle &) = delete;
};
// std::vector<Path::Point> (Eigen::Vector3f in the reference, std::array<float, 3> here) is already
// the packed xyz float array the C-ABI takes: no copy
struct XyzView {
  const float *ptr;
  const float *data() const { return ptr; }
};
inline XyzView flatten(const std::vector<::Path::Point> &cloud) {
  static_assert(sizeof(::Path::Point) == 3 * sizeof(float), "Path::Point must be three packed floats");
  return XyzView{cloud.empty() ? nullptr : cloud.front().data()};
}
inline Trajectory2D toTrajectory(const kc_cycle_result &r) {
  Trajectory2D t;
  const size_t P = static_cast<size_t>(r.n_points);
  t.numPointsPerTrajectory_ = t.velocities.numPointsPerTrajectory_ = t.path.numPointsPerTrajectory_ = P;
  if (r.found && P >= 2) {
    t.velocities.vx.assign(r.vx, r.vx + P - 1);
    t.velocities.vy.assign(r.vy, r.vy + P - 1);
    t.velocities.omega.assign(r.omega, r.omega + P - 1);
    t.path.x.assign(r.x, r.x + P);
    t.path.y.assign(r.y, r.y + P);
    t.path.z.assign(P, 0.0f);
  }
  return t;
}
}  // namespace detail

// ---------------------------------------------------------------------------------------------
class TrajectorySampler {
public:
  // ref: trajectory_sampler.h:22-59 TrajectorySamplerParameters (same names, defaults and ranges;
  // out-of-range -> std::out_of_range, unknown name -> std::invalid_argument as parameter.h:134-146)
  class TrajectorySamplerParameters {
  public:
    TrajectorySamplerParameters() {
      p_ = {{"time_step", {0.1, 0.001, 1000.0}},          {"prediction_horizon", {1.0, 0.001, 1000.0}},
            {"control_horizon", {1.0, 0.001, 1000.0}},    {"max_linear_samples", {10, 1, 1000}},
            {"max_angular_samples", {10, 1, 1000}},       {"octree_map_resolution", {0.1, 0.0, 1000.0}},
            {"drop_samples", {1, 0, 1}}};
    }
    void setParameter(const std::string &name, double value) {
      auto it = p_.find(name);
      if (it == p_.end()) throw std::invalid_argument("Parameter not found: " + name);
      if (value < it->second.lo || value > it->second.hi)
        throw std::out_of_range("Value out of range for parameter " + name);
      it->second.v = value;
    }
    template <typename T = double>
    T getParameter(const std::string &name) const {
      auto it = p_.find(name);
      if (it == p_.end()) throw std::invalid_argument("Parameter not found: " + name);
      return static_cast<T>(it->second.v);
    }

  private:
    struct Entry {
      double v, lo, hi;
    };
    std::map<std::string, Entry> p_;
  };

  // ref: trajectory_sampler.h:84-91 / trajectory_sampler.cpp:62-92 (configuration ctor; this one
  // initialises numCtrlPoints_ = control_horizon / time_step and takes drop_samples from the config)
  TrajectorySampler(TrajectorySamplerParameters config, ControlLimitsParams controlLimits,
                    ControlType controlType, const CollisionChecker::ShapeType robotShapeType,
                    const std::vector<float> robotDimensions, const Vector3f &sensor_position_body,
                    const Vector4f &sensor_rotation_body, const int maxNumThreads = 1)
      : cfg_(configFromParams(config, controlLimits, controlType, robotShapeType, robotDimensions,
                              sensor_position_body, sensor_rotation_body, maxNumThreads)),
        handle_(std::make_shared<detail::PlannerHandle>(cfg_)),
        base_horizon_(config.getParameter<double>("prediction_horizon")) {
    numTrajectories = static_cast<size_t>(kc_planner_num_trajectories(handle_->h));
    numPointsPerTrajectory = static_cast<size_t>(kc_planner_num_points(handle_->h));
  }
  static kc_planner_config configFromParams(const TrajectorySamplerParameters &config,
                                            const ControlLimitsParams &controlLimits, ControlType controlType,
                                            CollisionChecker::ShapeType robotShapeType,
                                            const std::vector<float> &robotDimensions,
                                            const Vector3f &sensor_position_body,
                                            const Vector4f &sensor_rotation_body, int maxNumThreads) {
    kc_planner_config c = detail::makeConfig(
        controlLimits, controlType, config.getParameter<double>("time_step"),
        config.getParameter<double>("prediction_horizon"), config.getParameter<double>("control_horizon"),
        config.getParameter<int>("max_linear_samples"), config.getParameter<int>("max_angular_samples"),
        robotShapeType, robotDimensions, sensor_position_body, sensor_rotation_body,
        config.getParameter<double>("octree_map_resolution"), maxNumThreads);
    c.drop_samples = config.getParameter<int>("drop_samples") != 0;
    return c;
  }

  // ref: trajectory_sampler.h:75-83 (explicit-argument ctor)
  TrajectorySampler(ControlLimitsParams controlLimits, ControlType controlType, double timeStep,
                    double predictionHorizon, double controlHorizon, int maxLinearSamples,
                    int maxAngularSamples, const CollisionChecker::ShapeType robotShapeType,
                    const std::vector<float> robotDimensions, const Vector3f &sensor_position_body,
                    const Vector4f &sensor_rotation_body, const double octreeRes,
                    const int maxNumThreads = 1)
      : cfg_(detail::makeConfig(controlLimits, controlType, timeStep, predictionHorizon, controlHorizon,
                                maxLinearSamples, maxAngularSamples, robotShapeType, robotDimensions,
                                sensor_position_body, sensor_rotation_body, octreeRes, maxNumThreads)),
        handle_(std::make_shared<detail::PlannerHandle>(cfg_)),
        base_horizon_(predictionHorizon) {
    numTrajectories = static_cast<size_t>(kc_planner_num_trajectories(handle_->h));
    numPointsPerTrajectory = static_cast<size_t>(kc_planner_num_points(handle_->h));
  }

  void setSampleDroppingMode(const bool drop) { kcThrow(kc_planner_set_drop_samples(handle_->h, drop)); }
  void resetOctreeResolution(const double res) {
    kcThrow(kc_planner_set_octree_resolution(handle_->h, res));
    cfg_.octree_resolution = res;
    if (checker_) checker_->resetOctreeResolution(res);
  }
  // ref: trajectory_sampler.cpp:374-408: updateState + checkStatesFeasibility<T>(states, sensor data):
  // true when ANY of the states collides
  void updateState(const ::Path::State &s) { checker().updateState(s.x, s.y, s.yaw); }
  template <class Sensor>
  bool checkStatesFeasibility(const std::vector<::Path::State> &states, const Sensor &sensor) {
    checker().updateSensorData(sensor);
    return checker().checkStates(states);
  }
  double getBasePredictionHorizon() const { return base_horizon_; }
  void setPredictionHorizon(double horizon) {
    int32_t n = 0;
    kcThrow(kc_planner_set_prediction_horizon(handle_->h, horizon, &n));
    numPointsPerTrajectory = static_cast<size_t>(n);
  }

  std::unique_ptr<TrajectorySamples2D> generateTrajectories(const Velocity2D &vel, const ::Path::State &pose,
                                                            const LaserScan &scan) {
    const double v[3] = {vel.vx(), vel.vy(), vel.omega()}, p[3] = {pose.x, pose.y, pose.yaw};
    if (scan.ranges.size() != scan.angles.size())
      throw std::invalid_argument("LaserScan ranges and angles must have the same size");
    kc_samples s{};
    kcThrow(kc_sampler_generate_scan(handle_->h, v, p, scan.ranges.data(), scan.angles.data(),
                                     static_cast<int32_t>(scan.ranges.size()), &s));
    return wrap(s);
  }
  std::unique_ptr<TrajectorySamples2D> generateTrajectories(const Velocity2D &vel, const ::Path::State &pose,
                                                            const std::vector<::Path::Point> &cloud) {
    const double v[3] = {vel.vx(), vel.vy(), vel.omega()}, p[3] = {pose.x, pose.y, pose.yaw};
    const detail::XyzView xyz = detail::flatten(cloud);
    kc_samples s{};
    kcThrow(kc_sampler_generate_cloud(handle_->h, v, p, xyz.data(), static_cast<int32_t>(cloud.size()), &s));
    return wrap(s);
  }

  size_t numTrajectories = 0;
  size_t numPointsPerTrajectory = 0;
  kc_planner *native() const { return handle_->h; }

private:
  std::unique_ptr<TrajectorySamples2D> wrap(const kc_samples &s) const {
    const size_t P = static_cast<size_t>(s.n_points), n = static_cast<size_t>(s.count);
    auto out = std::make_unique<TrajectorySamples2D>(std::max(numTrajectories, n), P);
    out->count = n;
    if (n) {
      std::copy(s.vx, s.vx + n * (P - 1), out->velocities.vx.data());
      std::copy(s.vy, s.vy + n * (P - 1), out->velocities.vy.data());
      std::copy(s.omega, s.omega + n * (P - 1), out->velocities.omega.data());
      std::copy(s.x, s.x + n * P, out->paths.x.data());
      std::copy(s.y, s.y + n * P, out->paths.y.data());
      out->slots.assign(s.slots, s.slots + n);
    }
    return out;
  }
  CollisionChecker &checker() {
    if (!checker_)
      checker_ = std::make_unique<CollisionChecker>(
          static_cast<CollisionChecker::ShapeType>(cfg_.robot_shape),
          std::vector<float>{cfg_.robot_dims[0], cfg_.robot_dims[1], cfg_.robot_dims[2]},
          Vector3f{cfg_.sensor_position[0], cfg_.sensor_position[1], cfg_.sensor_position[2]},
          Vector4f{cfg_.sensor_rotation[0], cfg_.sensor_rotation[1], cfg_.sensor_rotation[2],
                   cfg_.sensor_rotation[3]},
          cfg_.octree_resolution);
    return *checker_;
  }
  kc_planner_config cfg_;
  std::shared_ptr<detail::PlannerHandle> handle_;
  double base_horizon_;
  std::unique_ptr<CollisionChecker> checker_;
};

// ---------------------------------------------------------------------------------------------
class CostEvaluator {
public:
  // ref: cost_evaluator.h:22-50. Same parameter names, range [0, 1000] enforced like Parameter.
  class TrajectoryCostsWeights {
  public:
    TrajectoryCostsWeights() {
      for (const char *k : {"reference_path_distance_weight", "goal_distance_weight",
                            "obstacles_distance_weight", "smoothness_weight", "jerk_weight"})
        w_[k] = 1.0;
    }
    void setParameter(const std::string &name, double value) {
      auto it = w_.find(name);
      if (it == w_.end()) throw std::invalid_argument("Parameter not found: " + name);
      if (value < 0.0 || value > 1000.0) throw std::out_of_range("Value out of range for parameter " + name);
      it->second = value;
    }
    template <typename T = double>
    T getParameter(const std::string &name) const {
      auto it = w_.find(name);
      if (it == w_.end()) throw std::invalid_argument("Parameter not found: " + name);
      return static_cast<T>(it->second);
    }

  private:
    std::map<std::string, double> w_;
  };
  using CustomCostFunction = std::function<float(const Trajectory2D &, const ::Path::Path &)>;

  // ref: cost_evaluator.h:69-71 / :89-93
  CostEvaluator(TrajectoryCostsWeights &costsWeights, ControlLimitsParams ctrLimits,
                size_t /*maxNumTrajectories*/, size_t numPointsPerTrajectory,
                size_t /*maxRefPathSegmentSize*/)
      : CostEvaluator(costsWeights, Vector3f{0, 0, 0}, Vector4f{0, 0, 0, 1}, ctrLimits, 0,
                      numPointsPerTrajectory, 0) {}
  CostEvaluator(TrajectoryCostsWeights &costsWeights, const Vector3f &sensor_position_body,
                const Vector4f &sensor_rotation_body, ControlLimitsParams ctrLimits,
                size_t /*maxNumTrajectories*/, size_t numPointsPerTrajectory,
                size_t /*maxRefPathSegmentSize*/) {
    kc_planner_config c = detail::makeConfig(ctrLimits, ControlType::ACKERMANN, 0.1,
                                             0.1 * static_cast<double>(std::max<size_t>(numPointsPerTrajectory, 2)),
                                             0.1, 3, 3, CollisionChecker::ShapeType::CYLINDER, {0.1f, 0.1f},
                                             sensor_position_body, sensor_rotation_body, 0.1, 1);
    handle_ = std::make_shared<detail::PlannerHandle>(c);
    updateCostWeights(costsWeights);
  }

  void updateCostWeights(TrajectoryCostsWeights &w) {
    kcThrow(kc_planner_set_weights(handle_->h, w.getParameter<double>("reference_path_distance_weight"),
                                   w.getParameter<double>("goal_distance_weight"),
                                   w.getParameter<double>("obstacles_distance_weight"),
                                   w.getParameter<double>("smoothness_weight"),
                                   w.getParameter<double>("jerk_weight")));
  }
  void addCustomCost(double weight, CustomCostFunction f) { custom_.emplace_back(weight, std::move(f)); }

  void setPointScan(const LaserScan &scan, const ::Path::State &state, const float max_sensor_range,
                    const float multiple = 3.0f) {
    const double p[3] = {state.x, state.y, state.yaw};
    kcThrow(kc_cost_set_points_scan(handle_->h, scan.ranges.data(), scan.angles.data(),
                                    static_cast<int32_t>(scan.ranges.size()), p, max_sensor_range, multiple));
  }
  void setPointScan(const std::vector<::Path::Point> &cloud, const ::Path::State &state,
                    const float max_sensor_range, const float multiple = 3.0f) {
    const double p[3] = {state.x, state.y, state.yaw};
    const detail::XyzView xyz = detail::flatten(cloud);
    kcThrow(kc_cost_set_points_cloud(handle_->h, xyz.data(), static_cast<int32_t>(cloud.size()), p,
                                     max_sensor_range, multiple));
  }

  // ref: cost_evaluator.h:139-142
  TrajSearchResult getMinTrajectoryCost(const std::unique_ptr<TrajectorySamples2D> &trajs,
                                        const ::Path::Path *reference_path,
                                        const ::Path::Path::View &tracked_segment) {
    if (!reference_path) throw std::invalid_argument("reference path is NULL");
    if (reference_path != uploaded_ || reference_path->getSize() != uploaded_n_) {
      kcThrow(kc_planner_set_path(handle_->h, reference_path->X.data(), reference_path->Y.data(),
                                  reference_path->accumulated.data(),
                                  static_cast<int32_t>(reference_path->getSize()),
                                  reference_path->totalPathLength()));
      uploaded_ = reference_path;
      uploaded_n_ = reference_path->getSize();
    }
    const size_t n = trajs->size(), P = trajs->numPointsPerTrajectory_;
    // host callbacks: weight * value as double per callback, added on the device as float += double
    // in registration order (cost_evaluator.cpp:96-100)
    const size_t nc = custom_.size();
    std::vector<double> terms(n * nc);
    for (size_t i = 0; i < n && nc; ++i) {
      const Trajectory2D t = trajs->getIndex(i);
      for (size_t k = 0; k < nc; ++k)
        terms[i * nc + k] = custom_[k].first * static_cast<double>(custom_[k].second(t, *reference_path));
    }
    kc_cycle_result r{};
    kcThrow(kc_cost_evaluate(handle_->h, static_cast<int32_t>(n), static_cast<int32_t>(P),
                             trajs->velocities.vx.data(), trajs->velocities.vy.data(),
                             trajs->velocities.omega.data(), trajs->paths.x.data(), trajs->paths.y.data(),
                             static_cast<int32_t>(tracked_segment.getStartIndex()),
                             static_cast<int32_t>(tracked_segment.getSize()),
                             terms.empty() ? nullptr : terms.data(), static_cast<int32_t>(nc), nullptr, &r));
    TrajSearchResult out;
    out.isTrajFound = r.found != 0;
    out.trajCost = r.cost;
    out.trajectory = detail::toTrajectory(r);
    return out;
  }
  bool hasCustomCosts() const { return !custom_.empty(); }

private:
  std::shared_ptr<detail::PlannerHandle> handle_;
  std::vector<std::pair<double, CustomCostFunction>> custom_;
  const ::Path::Path *uploaded_ = nullptr;
  size_t uploaded_n_ = 0;
};

// ---------------------------------------------------------------------------------------------
// DWA (ref include/controllers/dwa.h + the Follower / Controller methods it inherits). Path
// following state lives behind kc_dwa_* (closest point, segments, adaptive horizon, tracked view);
// with a Path built from already-interpolated arrays the caller may also pin the tracked segment by
// hand (setTrackedSegment), which is what the cost-evaluator tests do.
// ref: include/controllers/controller.h:18-28 Controller::Result
struct Controller {
  struct Result {
    enum class Status { GOAL_REACHED, LOOSING_GOAL, COMMAND_FOUND, NO_COMMAND_POSSIBLE };
    Status status = Status::NO_COMMAND_POSSIBLE;
    Velocity2D velocity_command;
  };
};

class DWA {
public:
  // ref: dwa.h:24-32 / dwa.cpp:14-41
  DWA(ControlLimitsParams controlLimits, ControlType controlType, double timeStep,
      double predictionHorizon, double controlHorizon, int maxLinearSamples, int maxAngularSamples,
      const CollisionChecker::ShapeType robotShapeType, const std::vector<float> robotDimensions,
      const Vector3f &sensor_position_body, const Vector4f &sensor_rotation_body, const double octreeRes,
      CostEvaluator::TrajectoryCostsWeights costWeights, const int maxNumThreads = 1) {
    kc_follower_params_default(&follower_);
    configure(controlLimits, controlType, timeStep, predictionHorizon, controlHorizon, maxLinearSamples,
              maxAngularSamples, robotShapeType, robotDimensions, sensor_position_body,
              sensor_rotation_body, octreeRes, costWeights, maxNumThreads);
  }
  // ref: dwa.h:34-41 / dwa.cpp:43-69
  DWA(TrajectorySampler::TrajectorySamplerParameters config, ControlLimitsParams controlLimits,
      ControlType controlType, const CollisionChecker::ShapeType robotShapeType,
      const std::vector<float> robotDimensions, const Vector3f &sensor_position_body,
      const Vector4f &sensor_rotation_body, CostEvaluator::TrajectoryCostsWeights costWeights,
      const int maxNumThreads = 1) {
    kc_follower_params_default(&follower_);
    configure(config, controlLimits, controlType, robotShapeType, robotDimensions, sensor_position_body,
              sensor_rotation_body, costWeights, maxNumThreads);
  }
  ~DWA() { kc_dwa_destroy(d_); }
  DWA(const DWA &) = delete;
  DWA &operator=(const DWA &) = delete;

  // ref: dwa.h:53-74 / dwa.cpp:94-135: rebuilds the sampler and the cost evaluator. As in the
  // reference, custom costs registered on the old evaluator are gone afterwards.
  void configure(ControlLimitsParams controlLimits, ControlType controlType, double timeStep,
                 double predictionHorizon, double controlHorizon, int maxLinearSamples,
                 int maxAngularSamples, const CollisionChecker::ShapeType robotShapeType,
                 const std::vector<float> robotDimensions, const Vector3f &sensor_position_body,
                 const Vector4f &sensor_rotation_body, const double octreeRes,
                 CostEvaluator::TrajectoryCostsWeights costWeights, const int maxNumThreads = 1) {
    kc_planner_config c = detail::makeConfig(controlLimits, controlType, timeStep, predictionHorizon,
                                             controlHorizon, maxLinearSamples, maxAngularSamples,
                                             robotShapeType, robotDimensions, sensor_position_body,
                                             sensor_rotation_body, octreeRes, maxNumThreads);
    reconfigure(c, costWeights);
  }
  void configure(TrajectorySampler::TrajectorySamplerParameters config, ControlLimitsParams controlLimits,
                 ControlType controlType, const CollisionChecker::ShapeType robotShapeType,
                 const std::vector<float> robotDimensions, const Vector3f &sensor_position_body,
                 const Vector4f &sensor_rotation_body, CostEvaluator::TrajectoryCostsWeights costWeights,
                 const int maxNumThreads = 1) {
    kc_planner_config c = TrajectorySampler::configFromParams(config, controlLimits, controlType,
                                                              robotShapeType, robotDimensions,
                                                              sensor_position_body, sensor_rotation_body,
                                                              maxNumThreads);
    reconfigure(c, costWeights);
  }

  // ref: follower.cpp:17-48 setParams (the parameters the DWA path reads)
  void setFollowerParams(const kc_follower_params &fp) {
    follower_ = fp;
    create();
  }
  void resetOctreeResolution(const double res) { kcThrow(kc_planner_set_octree_resolution(planner(), res)); }
  void setSensorMaxRange(const float r) { kcThrow(kc_planner_set_max_range(planner(), r)); }
  void setCurrentState(const ::Path::State &s) {
    state_ = s;
    kcThrow(kc_dwa_set_current_state(d_, s.x, s.y, s.yaw, s.speed));
  }
  // ref: dwa.cpp:147-150 -> cost_evaluator.h:150-154. The callback sees each admissible trajectory
  // and the current (interpolated) reference path; it runs on the calling thread inside compute*.
  void addCustomCost(double weight, CostEvaluator::CustomCostFunction custom_cost_function) {
    customs_.push_back(std::make_unique<CustomHolder>(CustomHolder{this, std::move(custom_cost_function), weight}));
    kcThrow(kc_dwa_add_custom_cost(d_, weight, &DWA::customTrampoline, customs_.back().get()));
  }
  // ref: controller.cpp:22-33
  void setLinearControlLimits(const LinearVelocityControlParams &vx, const LinearVelocityControlParams &vy) {
    base_vx_ = vx.maxVel;
    base_vy_ = vy.maxVel;
    kcThrow(kc_dwa_set_control_limits(d_, base_vx_, base_vy_, base_om_));
  }
  void setAngularControlLimits(const AngularVelocityControlParams &p) {
    base_om_ = p.maxOmega;
    kcThrow(kc_dwa_set_control_limits(d_, base_vx_, base_vy_, base_om_));
  }
  // ref: follower.cpp:81-107
  void setCurrentPath(const ::Path::Path &path, const bool interpolate = true) {
    path_ = std::make_unique<::Path::Path>(path);
    callback_path_.reset();
    if (path_->prepared) {
      manual_ = true;
      kcThrow(kc_planner_set_path(planner(), path_->X.data(), path_->Y.data(), path_->accumulated.data(),
                                  static_cast<int32_t>(path_->getSize()), path_->totalPathLength()));
      seg_start_ = 0;
      seg_count_ = path_->getSize();
    } else {
      manual_ = false;
      kcThrow(kc_dwa_set_current_path(d_, path_->X.data(), path_->Y.data(),
                                      static_cast<int32_t>(path_->getSize()), interpolate ? 1 : 0));
    }
  }
  void clearCurrentPath() {
    path_.reset();
    callback_path_.reset();
    kcThrow(kc_dwa_clear_current_path(d_));
  }
  void setTrackedSegment(size_t start, size_t end) {
    if (!path_) throw std::invalid_argument("Pointer to global path is NULL. Cannot use DWA local planner without setting a global path");
    if (!manual_) throw std::invalid_argument("setTrackedSegment needs a Path built from interpolated arrays");
    const ::Path::Path::View v = path_->getPart(start, end);
    seg_start_ = v.getStartIndex();
    seg_count_ = v.getSize();
  }
  void setPredictionHorizon(double h) { kcThrow(kc_planner_set_prediction_horizon(planner(), h, nullptr)); }
  bool isGoalReached() {  // ref: follower.cpp:111-145
    int32_t r = 0;
    kcThrow(kc_dwa_is_goal_reached(d_, &r));
    return r != 0;
  }
  bool hasPath() const { return kc_dwa_has_path(d_) != 0; }
  size_t getCurrentSegmentIndex() const { return static_cast<size_t>(info_.segment_index); }
  const kc_dwa_info &getTrackingInfo() const { return info_; }
  // ref: follower.h:147-165 (clamped to the base-class limits)
  double getLinearVelocityCmdX() const { return cmd(0); }
  double getLinearVelocityCmdY() const { return cmd(1); }
  double getAngularVelocityCmd() const { return cmd(2); }

  // ref: dwa.h:113-128 computeVelocityCommand<T>
  template <typename T>
  Controller::Result computeVelocityCommand(const Velocity2D &global_vel, const T &scan_points) {
    const TrajSearchResult res = computeVelocityCommandsSet(global_vel, scan_points);
    Controller::Result out;
    if (res.isTrajFound) {
      out.status = Controller::Result::Status::COMMAND_FOUND;
      out.velocity_command = res.trajectory.velocities.getFront();
    } else {
      out.status = Controller::Result::Status::NO_COMMAND_POSSIBLE;
    }
    return out;
  }

  // ref: dwa.h:130-139 computeVelocityCommandsSet<T>
  TrajSearchResult computeVelocityCommandsSet(const Velocity2D &vel, const LaserScan &scan) {
    if (scan.ranges.size() != scan.angles.size())
      throw std::invalid_argument("LaserScan ranges and angles must have the same size");
    return cycle(vel, false, scan.ranges.data(), scan.angles.data(), static_cast<int32_t>(scan.ranges.size()));
  }
  TrajSearchResult computeVelocityCommandsSet(const Velocity2D &vel, const std::vector<::Path::Point> &cloud) {
    const detail::XyzView xyz = detail::flatten(cloud);
    return cycle(vel, true, xyz.data(), nullptr, static_cast<int32_t>(cloud.size()));
  }
  Velocity2D latestVelocityCommand() const { return latest_; }

  // ref: dwa.h:147-165 debugVelocitySearch<T> (the dropping mode stays set, as in the reference)
  void debugVelocitySearch(const Velocity2D &vel, const LaserScan &scan, const bool &drop_samples) {
    if (scan.ranges.size() != scan.angles.size())
      throw std::invalid_argument("LaserScan ranges and angles must have the same size");
    debugSearch(vel, false, scan.ranges.data(), scan.angles.data(), static_cast<int32_t>(scan.ranges.size()),
                drop_samples);
  }
  void debugVelocitySearch(const Velocity2D &vel, const std::vector<::Path::Point> &cloud,
                           const bool &drop_samples) {
    const detail::XyzView xyz = detail::flatten(cloud);
    debugSearch(vel, true, xyz.data(), nullptr, static_cast<int32_t>(cloud.size()), drop_samples);
  }
  // ref: dwa.cpp:235-250
  std::tuple<MatrixXfR, MatrixXfR> getDebuggingSamples() const {
    const TrajectorySamples2D s = getDebuggingSamplesPure();
    MatrixXfR px(s.size(), s.numPointsPerTrajectory_), py(s.size(), s.numPointsPerTrajectory_);
    std::copy(s.paths.x.v.begin(), s.paths.x.v.begin() + px.v.size(), px.v.begin());
    std::copy(s.paths.y.v.begin(), s.paths.y.v.begin() + py.v.size(), py.v.begin());
    return std::make_tuple(px, py);
  }
  TrajectorySamples2D getDebuggingSamplesPure() const {
    if (!debug_) throw std::invalid_argument("No debugging samples are available");
    return *debug_;
  }

private:
  struct CustomHolder {
    DWA *self;
    CostEvaluator::CustomCostFunction f;
    double weight;
  };
  // C callback -> std::function. Exceptions cannot cross the C boundary: the first one is kept and
  // rethrown by the compute call that triggered it.
  static double customTrampoline(const kc_trajectory_view *t, const kc_path_view *p, void *user) {
    CustomHolder *h = static_cast<CustomHolder *>(user);
    DWA *self = h->self;
    if (self->callback_error_) return 0.0;
    try {
      if (!self->callback_path_ || self->callback_path_->X.data() == nullptr ||
          self->callback_path_->getSize() != static_cast<size_t>(p->n))
        self->callback_path_ = std::make_unique<::Path::Path>(
            std::vector<float>(p->X, p->X + p->n), std::vector<float>(p->Y, p->Y + p->n),
            std::vector<float>(p->acc, p->acc + p->n), p->total_length);
      const size_t P = static_cast<size_t>(t->n_points);
      Trajectory2D traj;
      traj.numPointsPerTrajectory_ = traj.velocities.numPointsPerTrajectory_ = traj.path.numPointsPerTrajectory_ = P;
      traj.velocities.vx.assign(t->vx, t->vx + P - 1);
      traj.velocities.vy.assign(t->vy, t->vy + P - 1);
      traj.velocities.omega.assign(t->omega, t->omega + P - 1);
      traj.path.x.assign(t->x, t->x + P);
      traj.path.y.assign(t->y, t->y + P);
      traj.path.z.assign(P, 0.0f);
      return static_cast<double>(h->f(traj, *self->callback_path_));
    } catch (...) {
      self->callback_error_ = std::current_exception();
      return 0.0;
    }
  }
  void rethrowCallbackError() {
    if (callback_error_) {
      std::exception_ptr e = callback_error_;
      callback_error_ = nullptr;
      std::rethrow_exception(e);
    }
  }
  void reconfigure(kc_planner_config c, CostEvaluator::TrajectoryCostsWeights &costWeights) {
    c.w_path = costWeights.getParameter<double>("reference_path_distance_weight");
    c.w_goal = costWeights.getParameter<double>("goal_distance_weight");
    c.w_obstacles = costWeights.getParameter<double>("obstacles_distance_weight");
    c.w_smooth = costWeights.getParameter<double>("smoothness_weight");
    c.w_jerk = costWeights.getParameter<double>("jerk_weight");
    cfg_ = c;
    customs_.clear();
    create();
  }
  // (re)creates the native controller; the current path and state are carried over like the
  // Follower members that DWA::configure leaves alone
  void create() {
    kc_dwa *fresh = nullptr;
    kcThrow(kc_dwa_create(&cfg_, &follower_, &fresh));
    kc_dwa_destroy(d_);
    d_ = fresh;
    kcThrow(kc_dwa_set_control_limits(d_, base_vx_, base_vy_, base_om_));
    kcThrow(kc_dwa_set_current_state(d_, state_.x, state_.y, state_.yaw, state_.speed));
    for (const auto &h : customs_) kcThrow(kc_dwa_add_custom_cost(d_, h->weight, &DWA::customTrampoline, h.get()));
    if (path_) {
      const std::unique_ptr<::Path::Path> keep = std::move(path_);
      setCurrentPath(*keep);
    }
  }
  kc_planner *planner() const { return kc_dwa_planner(d_); }
  void requirePath() const {
    if (!path_) throw std::invalid_argument("Pointer to global path is NULL. Cannot use DWA local planner without setting a global path");
  }
  TrajSearchResult cycle(const Velocity2D &vel, bool cloud, const void *a, const void *b, int32_t n) {
    const double v[3] = {vel.vx(), vel.vy(), vel.omega()}, p[3] = {state_.x, state_.y, state_.yaw};
    kc_cycle_result r{};
    requirePath();
    if (manual_ && !customs_.empty()) {
      manualCustomCycle(v, p, cloud, a, b, n, r);
    } else if (manual_) {
      if (cloud)
        kcThrow(kc_planner_cycle_cloud(planner(), v, p, static_cast<const float *>(a), n,
                                       static_cast<int32_t>(seg_start_), static_cast<int32_t>(seg_count_), &r));
      else
        kcThrow(kc_planner_cycle_scan(planner(), v, p, static_cast<const double *>(a),
                                      static_cast<const double *>(b), n, static_cast<int32_t>(seg_start_),
                                      static_cast<int32_t>(seg_count_), &r));
    } else {
      const int32_t rc = cloud ? kc_dwa_compute_cloud(d_, v, static_cast<const float *>(a), n, &r, &info_)
                               : kc_dwa_compute_scan(d_, v, static_cast<const double *>(a),
                                                     static_cast<const double *>(b), n, &r, &info_);
      rethrowCallbackError();
      kcThrow(rc);
    }
    return finish(r);
  }
  // hand-pinned tracked segment + custom costs: the reference's three steps through the C-ABI
  void manualCustomCycle(const double v[3], const double p[3], bool cloud, const void *a, const void *b,
                         int32_t n, kc_cycle_result &r) {
    kc_samples s{};
    if (cloud)
      kcThrow(kc_sampler_generate_cloud(planner(), v, p, static_cast<const float *>(a), n, &s));
    else
      kcThrow(kc_sampler_generate_scan(planner(), v, p, static_cast<const double *>(a),
                                       static_cast<const double *>(b), n, &s));
    r.n_points = s.n_points;
    if (s.count == 0) return;
    const float range = kc_planner_get_max_range(planner());
    if (cloud)
      kcThrow(kc_cost_set_points_cloud(planner(), static_cast<const float *>(a), n, p, range, 3.0f));
    else
      kcThrow(kc_cost_set_points_scan(planner(), static_cast<const double *>(a), static_cast<const double *>(b),
                                      n, p, range, 3.0f));
    const size_t nc = customs_.size(), P = static_cast<size_t>(s.n_points);
    const kc_path_view pv{static_cast<int32_t>(path_->getSize()), path_->X.data(), path_->Y.data(),
                          path_->accumulated.data(), path_->totalPathLength()};
    std::vector<double> terms(static_cast<size_t>(s.count) * nc);
    for (int32_t t = 0; t < s.count; ++t) {
      const kc_trajectory_view tv{s.n_points,      s.vx + t * (P - 1), s.vy + t * (P - 1),
                                  s.omega + t * (P - 1), s.x + t * P,        s.y + t * P};
      for (size_t k = 0; k < nc; ++k)
        terms[t * nc + k] = customs_[k]->weight * customTrampoline(&tv, &pv, customs_[k].get());
    }
    rethrowCallbackError();
    kcThrow(kc_cost_evaluate(planner(), s.count, s.n_points, s.vx, s.vy, s.omega, s.x, s.y,
                             static_cast<int32_t>(seg_start_), static_cast<int32_t>(seg_count_), terms.data(),
                             static_cast<int32_t>(nc), nullptr, &r));
  }
  void debugSearch(const Velocity2D &vel, bool cloud, const void *a, const void *b, int32_t n, bool drop) {
    const double v[3] = {vel.vx(), vel.vy(), vel.omega()}, p[3] = {state_.x, state_.y, state_.yaw};
    requirePath();
    kc_samples s{};
    if (manual_) {
      kcThrow(kc_planner_set_drop_samples(planner(), drop));
      if (cloud)
        kcThrow(kc_sampler_generate_cloud(planner(), v, p, static_cast<const float *>(a), n, &s));
      else
        kcThrow(kc_sampler_generate_scan(planner(), v, p, static_cast<const double *>(a),
                                         static_cast<const double *>(b), n, &s));
    } else if (cloud) {
      kcThrow(kc_dwa_debug_velocity_search_cloud(d_, v, static_cast<const float *>(a), n, drop, &s));
    } else {
      kcThrow(kc_dwa_debug_velocity_search_scan(d_, v, static_cast<const double *>(a),
                                                static_cast<const double *>(b), n, drop, &s));
    }
    const size_t P = static_cast<size_t>(s.n_points), cnt = static_cast<size_t>(s.count);
    debug_ = std::make_unique<TrajectorySamples2D>(
        std::max(cnt, static_cast<size_t>(kc_planner_num_trajectories(planner()))), P);
    debug_->count = cnt;
    if (cnt) {
      std::copy(s.vx, s.vx + cnt * (P - 1), debug_->velocities.vx.data());
      std::copy(s.vy, s.vy + cnt * (P - 1), debug_->velocities.vy.data());
      std::copy(s.omega, s.omega + cnt * (P - 1), debug_->velocities.omega.data());
      std::copy(s.x, s.x + cnt * P, debug_->paths.x.data());
      std::copy(s.y, s.y + cnt * P, debug_->paths.y.data());
      debug_->slots.assign(s.slots, s.slots + cnt);
    }
  }
  double cmd(int i) const {
    if (manual_) {
      const double v[3] = {latest_.vx(), latest_.vy(), latest_.omega()}, lim[3] = {base_vx_, base_vy_, base_om_};
      return std::max(std::min(v[i], lim[i]), -lim[i]);
    }
    double c[3];
    kcThrow(kc_dwa_get_command(d_, c));
    return c[i];
  }
  TrajSearchResult finish(const kc_cycle_result &r) {
    TrajSearchResult out;
    out.isTrajFound = r.found != 0;
    out.trajCost = r.cost;
    out.trajectory = detail::toTrajectory(r);
    if (out.isTrajFound) latest_ = out.trajectory.velocities.getFront();
    return out;
  }
  kc_planner_config cfg_{};
  kc_follower_params follower_{};
  kc_dwa *d_ = nullptr;
  kc_dwa_info info_{};
  std::unique_ptr<::Path::Path> path_;
  std::unique_ptr<::Path::Path> callback_path_;  // the interpolated path as the callbacks see it
  std::vector<std::unique_ptr<CustomHolder>> customs_;
  std::exception_ptr callback_error_;
  std::unique_ptr<TrajectorySamples2D> debug_;
  ::Path::State state_;
  bool manual_ = false;
  size_t seg_start_ = 0, seg_count_ = 0;
  double base_vx_ = 1.0, base_vy_ = 1.0, base_om_ = 1.0;  // Controller::ctrlimitsParams defaults
  Velocity2D latest_;
};

}  // namespace Control

// ---------------------------------------------------------------------------------------------
namespace Mapping {
enum class OccupancyType { UNEXPLORED = -1, EMPTY = 0, OCCUPIED = 100 };  // local_mapper.h:9

class LocalMapperGPU {
public:
  // ref: local_mapper_gpu.h:15-21
  LocalMapperGPU(const int gridHeight, const int gridWidth, const float resolution,
                 const Vector3f &laserscanPosition, const float laserscanOrientation,
                 const bool isPointCloud, const int scanSize, const float angleStep,
                 const float maxHeight, const float minHeight, const float rangeMax,
                 const int maxPointsPerLine = 32)
      : gridData(static_cast<size_t>(gridHeight), static_cast<size_t>(gridWidth)) {
    kc_mapper_config c{};
    c.grid_height = gridHeight;
    c.grid_width = gridWidth;
    c.resolution = resolution;
    for (int i = 0; i < 3; ++i) c.laserscan_position[i] = laserscanPosition[i];
    c.laserscan_orientation = laserscanOrientation;
    c.is_pointcloud = isPointCloud;
    c.scan_size = scanSize;
    c.angle_step = angleStep;
    c.max_height = maxHeight;
    c.min_height = minHeight;
    c.range_max = rangeMax;
    c.max_points_per_line = maxPointsPerLine;
    kcThrow(kc_mapper_create(&c, &h_));
  }
  ~LocalMapperGPU() { kc_mapper_destroy(h_); }
  LocalMapperGPU(const LocalMapperGPU &) = delete;
  LocalMapperGPU &operator=(const LocalMapperGPU &) = delete;

  // returns a reference to the member grid, reused by every call (local_mapper.cpp:219)
  MatrixXi &scanToGrid(const std::vector<double> &angles, const std::vector<double> &ranges) {
    if (angles.size() != ranges.size()) throw std::invalid_argument("angles and ranges must have the same size");
    kcThrow(kc_mapper_scan_to_grid(h_, angles.data(), ranges.data(), static_cast<int32_t>(angles.size()),
                                   gridData.data()));
    return gridData;
  }
  MatrixXi &scanToGrid(const std::vector<int8_t> &data, int point_step, int row_step, int height, int width,
                       float x_offset, float y_offset, float z_offset) {
    kcThrow(kc_mapper_cloud_to_grid(h_, data.data(), static_cast<int64_t>(data.size()), point_step, row_step,
                                    height, width, x_offset, y_offset, z_offset, gridData.data()));
    return gridData;
  }

  // Bayesian update (ref: local_mapper.h:58-75 constructor arguments, local_mapper.cpp:222-238)
  void setBayesianParams(float pPrior, float pOccupied, float pEmpty, float rangeSure, float wallSize) {
    kcThrow(kc_mapper_set_bayesian_params(h_, pPrior, pOccupied, pEmpty, rangeSure, wallSize));
  }
  std::tuple<MatrixXi &, MatrixXf &> scanToGridBaysian(const std::vector<double> &angles,
                                                       const std::vector<double> &ranges) {
    if (angles.size() != ranges.size()) throw std::invalid_argument("angles and ranges must have the same size");
    if (gridDataProb.rows() != gridData.rows()) gridDataProb = MatrixXf(gridData.rows(), gridData.cols());
    kcThrow(kc_mapper_scan_to_grid_bayesian(h_, angles.data(), ranges.data(), static_cast<int32_t>(angles.size()),
                                            gridData.data(), gridDataProb.data()));
    return std::tie(gridData, gridDataProb);
  }
  // ref: local_mapper.cpp:253-264 (raw point-cloud overload: angle-step binning, then the scan path)
  std::tuple<MatrixXi &, MatrixXf &> scanToGridBaysian(const std::vector<int8_t> &data, int point_step,
                                                       int row_step, int height, int width, float x_offset,
                                                       float y_offset, float z_offset) {
    if (gridDataProb.rows() != gridData.rows()) gridDataProb = MatrixXf(gridData.rows(), gridData.cols());
    kcThrow(kc_mapper_cloud_to_grid_bayesian(h_, data.data(), static_cast<int64_t>(data.size()), point_step,
                                             row_step, height, width, x_offset, y_offset, z_offset,
                                             gridData.data(), gridDataProb.data()));
    return std::tie(gridData, gridDataProb);
  }
  // ref: local_mapper.cpp:17-78
  void getPreviousGridInCurrentPose(const std::array<float, 2> &currentPositionInPreviousPose,
                                    double currentOrientationInPreviousPose) {
    kcThrow(kc_mapper_previous_grid_in_current_pose(h_, currentPositionInPreviousPose[0],
                                                    currentPositionInPreviousPose[1],
                                                    currentOrientationInPreviousPose));
  }

private:
  kc_mapper *h_ = nullptr;
  MatrixXi gridData;
  MatrixXf gridDataProb;
};
}  // namespace Mapping

// ---------------------------------------------------------------------------------------------
class CriticalZoneChecker {
public:
  enum class InputType { LASERSCAN = 0, POINTCLOUD = 1 };  // critical_zone_check.h:15-18
};

class CriticalZoneCheckerGPU : public CriticalZoneChecker {
public:
  // ref: critical_zone_check_gpu.h:39-47
  CriticalZoneCheckerGPU(InputType input_type, const CollisionChecker::ShapeType robot_shape_type,
                         const std::vector<float> &robot_dimensions, const Vector3f &sensor_position_body,
                         const Vector4f &sensor_rotation_body, const float critical_angle,
                         const float critical_distance, const float slowdown_distance,
                         const std::vector<double> &angles, const float min_height, const float max_height,
                         const float range_max, const int cloud_field_type = KC_FLOAT32) {
    kc_critical_zone_config c{};
    c.input_type = static_cast<int32_t>(input_type);
    c.robot_shape = static_cast<int32_t>(robot_shape_type);
    for (size_t i = 0; i < 3; ++i) c.robot_dims[i] = i < robot_dimensions.size() ? robot_dimensions[i] : 0.0f;
    for (int i = 0; i < 3; ++i) c.sensor_position[i] = sensor_position_body[i];
    for (int i = 0; i < 4; ++i) c.sensor_rotation[i] = sensor_rotation_body[i];
    c.critical_angle = critical_angle;
    c.critical_distance = critical_distance;
    c.slowdown_distance = slowdown_distance;
    c.min_height = min_height;
    c.max_height = max_height;
    c.range_max = range_max;
    c.cloud_field_type = cloud_field_type;
    kcThrow(kc_critical_zone_create(&c, angles.data(), static_cast<int32_t>(angles.size()), &h_));
  }
  ~CriticalZoneCheckerGPU() { kc_critical_zone_destroy(h_); }
  CriticalZoneCheckerGPU(const CriticalZoneCheckerGPU &) = delete;
  CriticalZoneCheckerGPU &operator=(const CriticalZoneCheckerGPU &) = delete;

  float check(const std::vector<double> &ranges, const bool forward) {
    float f = 0.0f;
    kcThrow(kc_critical_zone_check_scan(h_, ranges.data(), static_cast<int32_t>(ranges.size()), forward, &f));
    return f;
  }
  float check(const std::vector<int8_t> &data, int point_step, int row_step, int height, int width,
              int x_offset, int y_offset, int z_offset, const bool forward) {
    float f = 0.0f;
    kcThrow(kc_critical_zone_check_cloud(h_, data.data(), static_cast<int64_t>(data.size()), point_step,
                                         row_step, height, width, x_offset, y_offset, z_offset, forward, &f));
    return f;
  }

private:
  kc_critical_zone *h_ = nullptr;
};

}  // namespace Kompass
