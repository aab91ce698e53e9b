// C++ host-API tests: the reference's own Boost cases restated against the drop-in classes of
// kompass-core_b200/host/kompass_b200.hpp (same class / method names, same expectations).
//   ref: src/kompass_cpp/tests/critical_zone_test_gpu.cpp:118-288 (14 cases)
//        src/kompass_cpp/tests/cost_evaluator_test.cpp:217-461 (known answers)
//        src/kompass_cpp/tests/dwa_test.cpp (planner never fails on a free scene)
// Runs on the GPU box only (tests/test_gpu_cpp_host.py builds and executes it).
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../kompass-core_b200/host/kompass_b200.hpp"

using namespace Kompass;
static int failures = 0;
#define CHECK(cond)                                                   \
  do {                                                                \
    if (!(cond)) {                                                    \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);     \
      ++failures;                                                     \
    }                                                                 \
  } while (0)
static bool near(float a, float b, float tol) { return std::fabs(a - b) <= tol * std::fmax(1.0f, std::fabs(b)); }

// ---- helpers of tests/test.h ----
static void initLaserscan(size_t N, double r, std::vector<double> &ranges, std::vector<double> &angles) {
  angles.resize(N);
  ranges.resize(N);
  for (size_t i = 0; i < N; ++i) {
    angles[i] = 2.0 * M_PI * static_cast<double>(i) / N;
    ranges[i] = r;
  }
}
static void setLaserscanAtAngle(double angle, double value, std::vector<double> &ranges, std::vector<double> &angles) {
  angle = std::fmod(angle, 2 * M_PI);
  if (angle < 0) angle += 2 * M_PI;
  double best = 2.0 * M_PI;
  size_t idx = 0;
  for (size_t i = 0; i < angles.size(); ++i) {
    const double d = std::abs(angles[i] - angle);
    if (d < best) {
      best = d;
      idx = i;
    }
  }
  ranges[idx] = value;
}
struct PointXYZ {
  float x, y, z, padding;
};
static void addPointToCloud(std::vector<int8_t> &c, float x, float y, float z) {
  PointXYZ p{x, y, z, 0.0f};
  const int8_t *raw = reinterpret_cast<const int8_t *>(&p);
  c.insert(c.end(), raw, raw + sizeof(PointXYZ));
}

static void test_critical_zone() {
  std::vector<float> dims{0.51f, 2.0f};
  std::vector<double> angles, ranges;
  initLaserscan(360, 10.0, ranges, angles);
  CriticalZoneCheckerGPU zone(CriticalZoneChecker::InputType::LASERSCAN, CollisionChecker::ShapeType::CYLINDER,
                              dims, {0.22f, 0.0f, 0.4f}, {0.0f, 0.0f, 0.99f, 0.0f}, 160.0f, 0.3f, 0.6f, angles,
                              0.1f, 2.0f, 20.0f);
  for (double a : {0.0, 0.1, -0.1}) setLaserscanAtAngle(a, 0.2, ranges, angles);
  CHECK(zone.check(ranges, true) == 1.0f);                       // 1
  initLaserscan(360, 10.0, ranges, angles);
  CHECK(zone.check(ranges, true) == 1.0f);                       // 2
  for (double a : {M_PI, M_PI + 0.1, M_PI - 0.1}) setLaserscanAtAngle(a, 0.2, ranges, angles);
  CHECK(zone.check(ranges, true) == 0.0f);                       // 3
  CHECK(zone.check(ranges, false) == 1.0f);                      // 4
  for (double a : {0.0, 0.1, -0.1}) setLaserscanAtAngle(a, 0.2, ranges, angles);
  CHECK(zone.check(ranges, false) == 0.0f);                      // 5
  initLaserscan(360, 10.0, ranges, angles);
  setLaserscanAtAngle(0.0, 1.3, ranges, angles);
  float r = zone.check(ranges, false);
  CHECK(r > 0.0f && r < 1.0f);                                   // 6
  CHECK(zone.check(ranges, true) == 1.0f);                       // 7
  setLaserscanAtAngle(M_PI, 0.7, ranges, angles);
  r = zone.check(ranges, true);
  CHECK(r > 0.0f && r < 1.0f);                                   // 8

  std::vector<double> pc_angles, dummy;
  initLaserscan(360, 10.0, dummy, pc_angles);
  CriticalZoneCheckerGPU pc(CriticalZoneChecker::InputType::POINTCLOUD, CollisionChecker::ShapeType::CYLINDER, dims,
                            {0, 0, 0}, {0, 0, 0, 1}, 160.0f, 0.3f, 0.6f, pc_angles, 0.1f, 2.0f, 20.0f);
  std::vector<int8_t> cloud;
  const int ps = sizeof(PointXYZ), xo = offsetof(PointXYZ, x), yo = offsetof(PointXYZ, y), zo = offsetof(PointXYZ, z);
  auto run = [&](bool fwd) {
    const int n = static_cast<int>(cloud.size() / ps);
    return pc.check(cloud, ps, n * ps, 1, n, xo, yo, zo, fwd);
  };
  CHECK(run(true) == 1.0f);                                      // 9
  addPointToCloud(cloud, 0.7f, 0.0f, 0.5f);
  CHECK(run(true) == 0.0f);                                      // 10
  cloud.clear();
  addPointToCloud(cloud, 0.7f, 0.0f, 3.0f);
  CHECK(run(true) == 1.0f);                                      // 11
  cloud.clear();
  addPointToCloud(cloud, 0.95f, 0.0f, 0.5f);
  r = run(true);
  CHECK(r > 0.4f && r < 0.6f);                                   // 12
  cloud.clear();
  for (auto p : std::vector<std::array<float, 3>>{{0.95f, 0, 0.5f}, {1, 1, 0.5f}, {-1, -1, 0.5f}, {-0.1f, -0.1f, 3.0f},
                                                  {-0.1f, -0.1f, -3.0f}, {0.1f, 0.2f, 4.0f}, {0.1f, 0.2f, -4.0f}, {0.75f, 0, 0.5f}})
    addPointToCloud(cloud, p[0], p[1], p[2]);
  CHECK(run(true) == 0.0f);                                      // 13
  cloud.clear();
  for (auto p : std::vector<std::array<float, 3>>{{0.95f, 0, 0.5f}, {-0.95f, 0, 0.5f}, {1, 1, 0.5f}, {-1, -1, 0.5f},
                                                  {-0.1f, -0.1f, 3.0f}, {-0.1f, -0.1f, -3.0f}, {0.1f, 0.2f, 4.0f}, {0.1f, 0.2f, -4.0f}})
    addPointToCloud(cloud, p[0], p[1], p[2]);
  r = run(false);
  CHECK(r > 0.4f && r < 0.6f);                                   // 14
  bool threw = false;
  try {
    CriticalZoneCheckerGPU bad(CriticalZoneChecker::InputType::LASERSCAN, CollisionChecker::ShapeType::CYLINDER, dims,
                               {0, 0, 0}, {0, 0, 0, 1}, 160.0f, 0.6f, 0.3f, angles, 0.1f, 2.0f, 20.0f);
  } catch (const std::invalid_argument &) {
    threw = true;
  }
  CHECK(threw);  // slowdown <= critical (critical_zone_check.cpp:53-57)
}

// straight path (0,0)->(10,0) interpolated at 1 m: X = 0..10, prefix lengths 0..10
static ::Path::Path straightPath() {
  std::vector<float> X, Y, acc;
  for (int i = 0; i <= 10; ++i) {
    X.push_back((float)i);
    Y.push_back(0.0f);
    acc.push_back((float)i);
  }
  return ::Path::Path(X, Y, acc, 10.0f);
}
static Control::ControlLimitsParams dummyLimits() {
  return Control::ControlLimitsParams(Control::LinearVelocityControlParams(1, 1, 1),
                                      Control::LinearVelocityControlParams(1, 1, 1),
                                      Control::AngularVelocityControlParams(1, 1, 1, 1));
}
static Control::CostEvaluator::TrajectoryCostsWeights solo(const char *name) {
  Control::CostEvaluator::TrajectoryCostsWeights w;
  for (const char *k : {"reference_path_distance_weight", "goal_distance_weight", "obstacles_distance_weight",
                        "smoothness_weight", "jerk_weight"})
    w.setParameter(k, 0.0);
  w.setParameter(name, 1.0);
  return w;
}
static std::unique_ptr<Control::TrajectorySamples2D> oneSample(const std::vector<std::array<float, 2>> &pts,
                                                               const std::vector<std::array<double, 3>> &vels = {}) {
  const size_t n = pts.size();
  auto s = std::make_unique<Control::TrajectorySamples2D>(1, n);
  Control::TrajectoryVelocities2D v;
  Control::TrajectoryPath p;
  for (size_t i = 0; i < n; ++i) {
    p.x.push_back(pts[i][0]);
    p.y.push_back(pts[i][1]);
    p.z.push_back(0.0f);
    if (i + 1 < n) {
      v.vx.push_back(vels.empty() ? 0.0f : (float)vels[i][0]);
      v.vy.push_back(vels.empty() ? 0.0f : (float)vels[i][1]);
      v.omega.push_back(vels.empty() ? 0.0f : (float)vels[i][2]);
    }
  }
  s->push_back(v, p);
  return s;
}
static float evalCost(Control::CostEvaluator::TrajectoryCostsWeights w, const ::Path::Path &ref,
                      std::unique_ptr<Control::TrajectorySamples2D> samples,
                      const std::vector<::Path::Point> &obstacles = {},
                      Control::CostEvaluator::CustomCostFunction custom = nullptr) {
  Control::CostEvaluator ev(w, dummyLimits(), samples->size(), samples->numPointsPerTrajectory_, ref.getSize());
  if (!obstacles.empty()) ev.setPointScan(obstacles, ::Path::State(), 30.0f, 3.0f);
  if (custom) ev.addCustomCost(2.0, custom);
  auto res = ev.getMinTrajectoryCost(samples, &ref, ref.getPart(0, 4));  // segment 0: X = 0..4
  CHECK(res.isTrajFound);
  return res.trajCost;
}

static void test_cost_evaluator() {
  const ::Path::Path ref = straightPath();
  std::vector<std::array<float, 2>> at4(5, {4.0f, 0.0f}), origin(5, {0.0f, 0.0f});
  CHECK(near(evalCost(solo("goal_distance_weight"), ref, oneSample(at4)), 0.6f, 1e-4f));
  CHECK(near(evalCost(solo("goal_distance_weight"), ref, oneSample(std::vector<std::array<float, 2>>(5, {4.0f, 0.1f}))), 0.61f, 1e-4f));
  CHECK(near(evalCost(solo("goal_distance_weight"), ref, oneSample(std::vector<std::array<float, 2>>(5, {4.0f, 0.5f}))), 0.65f, 1e-4f));
  std::vector<std::array<float, 2>> on{{0, 0}, {1, 0}, {2, 0}, {3, 0}, {4, 0}}, off{{0, .5f}, {1, .5f}, {2, .5f}, {3, .5f}, {4, .5f}};
  CHECK(near(evalCost(solo("reference_path_distance_weight"), ref, oneSample(on)), 0.0f, 1e-4f));
  CHECK(near(evalCost(solo("reference_path_distance_weight"), ref, oneSample(off)), (0.5f + 0.5f / 4.0f) / 2.0f, 1e-4f));
  CHECK(near(evalCost(solo("smoothness_weight"), ref, oneSample(origin, {{1, 0, 0}, {1, 0, 0}, {1, 0, 0}, {1, 0, 0}})), 0.0f, 1e-4f));
  CHECK(near(evalCost(solo("smoothness_weight"), ref, oneSample(origin, {{0, 0, 0}, {1, 0, 0}, {1, 0, 0}, {1, 0, 0}})), 1.0f / 12.0f, 1e-4f));
  CHECK(near(evalCost(solo("jerk_weight"), ref, oneSample(origin, {{.1, 0, 0}, {.2, 0, 0}, {.3, 0, 0}, {.4, 0, 0}})), 0.0f, 1e-4f));
  CHECK(near(evalCost(solo("jerk_weight"), ref, oneSample(origin, {{0, 0, 0}, {1, 0, 0}, {3, 0, 0}, {6, 0, 0}})), 2.0f / 12.0f, 1e-4f));
  CHECK(near(evalCost(solo("obstacles_distance_weight"), ref, oneSample(origin), {{20.0f, 0, 0}}), 0.0f, 1e-4f));
  CHECK(near(evalCost(solo("obstacles_distance_weight"), ref, oneSample(origin), {{0.0f, 0, 0}}), 1.0f, 1e-4f));
  CHECK(near(evalCost(solo("obstacles_distance_weight"), ref, oneSample(origin), {{5.0f, 0, 0}}), 0.5f, 1e-4f));
  // custom cost: weight 2 * callback 0.25 added on top of the goal cost
  const float c = evalCost(solo("goal_distance_weight"), ref, oneSample(at4), {},
                           [](const Control::Trajectory2D &t, const ::Path::Path &) { return t.path.getEnd()[0] / 16.0f; });
  CHECK(near(c, 0.6f + 0.5f, 1e-4f));
}

static void test_dwa_and_sampler() {
  Control::ControlLimitsParams lim(Control::LinearVelocityControlParams(1.0, 5.0, 10.0),
                                   Control::LinearVelocityControlParams(0.0, 0.0, 0.0),
                                   Control::AngularVelocityControlParams(M_PI, 4.0, 3.0, 3.0));
  Control::CostEvaluator::TrajectoryCostsWeights w;
  Control::DWA dwa(lim, Control::ControlType::DIFFERENTIAL_DRIVE, 0.1, 1.0, 0.2, 20, 20,
                   CollisionChecker::ShapeType::CYLINDER, {0.1f, 0.4f}, {0, 0, 0}, {0, 0, 0, 1}, 0.1, w);
  std::vector<double> ranges, angles;
  initLaserscan(360, 5.0, ranges, angles);
  bool threw = false;
  try {
    dwa.computeVelocityCommandsSet(Control::Velocity2D(0, 0, 0), Control::LaserScan(ranges, angles));
  } catch (const std::invalid_argument &) {
    threw = true;  // ref dwa.h:187-191: no global path
  }
  CHECK(threw);
  dwa.setCurrentPath(straightPath());
  dwa.setCurrentState(::Path::State(0.0, 0.0, 0.0));
  dwa.setTrackedSegment(0, 4);
  auto res = dwa.computeVelocityCommandsSet(Control::Velocity2D(0.2, 0, 0), Control::LaserScan(ranges, angles));
  CHECK(res.isTrajFound);
  CHECK(res.trajectory.path.x.size() == 10 && res.trajectory.velocities.vx.size() == 9);
  CHECK(res.trajectory.velocities.vx[0] > 0.0f);  // free scene on a straight path: drive forward
  CHECK(res.trajectory.path.x[0] == 0.0f);
  threw = false;
  try {
    dwa.setTrackedSegment(5, 50);
  } catch (const std::out_of_range &) {
    threw = true;  // ref path.cpp:80-86
  }
  CHECK(threw);
  // a wall all around: planner reports "no admissible trajectory" without throwing
  std::vector<double> wall(360, 0.12);
  auto blocked = dwa.computeVelocityCommandsSet(Control::Velocity2D(0, 0, 0), Control::LaserScan(wall, angles));
  CHECK(!blocked.isTrajFound && blocked.trajCost == 0.0f);

  Control::TrajectorySampler sampler(lim, Control::ControlType::DIFFERENTIAL_DRIVE, 0.1, 1.0, 0.2, 20, 20,
                                     CollisionChecker::ShapeType::CYLINDER, {0.1f, 0.4f}, {0, 0, 0}, {0, 0, 0, 1}, 0.1);
  CHECK(sampler.numTrajectories == 441 && sampler.numPointsPerTrajectory == 10);
  auto samples = sampler.generateTrajectories(Control::Velocity2D(0.2, 0, 0), ::Path::State(0, 0, 0),
                                              Control::LaserScan(ranges, angles));
  CHECK(samples->size() > 0 && samples->size() <= 441);
  CHECK(samples->paths.x(0, 0) == 0.0f);
  std::vector<::Path::Point> cloud{{0.3f, 0.0f, 0.0f}};
  auto fewer = sampler.generateTrajectories(Control::Velocity2D(0.2, 0, 0), ::Path::State(0, 0, 0), cloud);
  CHECK(fewer->size() < samples->size());
}

// ref: tests/dwa_test.cpp:21-158 (test_DWA): follow a straight 2 m path to the goal, path given as
// raw way-points (interpolated + segmented by setCurrentPath), closed loop with the getters
static void test_dwa_closed_loop() {
  Control::ControlLimitsParams lim(Control::LinearVelocityControlParams(1.0, 5.0, 10.0),
                                   Control::LinearVelocityControlParams(1, 3, 5),
                                   Control::AngularVelocityControlParams(3.14, 2.0, 3.0, 3.0));
  Control::CostEvaluator::TrajectoryCostsWeights w;
  w.setParameter("reference_path_distance_weight", 1.0);
  w.setParameter("goal_distance_weight", 1.0);
  w.setParameter("obstacles_distance_weight", 0.0);
  w.setParameter("smoothness_weight", 0.0);
  w.setParameter("jerk_weight", 0.0);
  Control::DWA planner(lim, Control::ControlType::ACKERMANN, 0.1, 1.0, 0.2, 20, 20,
                       CollisionChecker::ShapeType::CYLINDER, {0.1f, 0.4f}, {0, 0, 0}, {0, 0, 0, 1}, 0.1, w, 10);
  ::Path::Path path(std::vector<::Path::Point>{{0.0f, 0.0f, 0.0f}, {1.0f, 0.0f, 0.0f}, {2.0f, 0.0f, 0.0f}});
  planner.setCurrentPath(path);
  ::Path::State robot(-0.51731912, 0.0, 0.0, 0.0);
  Control::Velocity2D control;
  Control::LaserScan scan({0.4, 0.3}, {10, 10.1});
  int counter = 0;
  planner.setCurrentState(robot);
  while (!planner.isGoalReached() && counter < 150) {
    counter++;
    planner.setCurrentState(robot);
    Control::TrajSearchResult result = planner.computeVelocityCommandsSet(control, scan);
    if (result.isTrajFound) {
      const double vx = planner.getLinearVelocityCmdX(), vy = planner.getLinearVelocityCmdY(),
                   om = planner.getAngularVelocityCmd();
      // applyControl (tests/controller_test_helpers.h:12-31)
      robot.x += (vx * std::cos(robot.yaw) - vy * std::sin(robot.yaw)) * 0.1;
      robot.y += (vx * std::sin(robot.yaw) + vy * std::cos(robot.yaw)) * 0.1;
      robot.yaw += om * 0.1;
    }
  }
  CHECK(planner.isGoalReached());
  CHECK(counter < 150);
  CHECK(std::abs(robot.y) < 0.2);
}

// DWA::addCustomCost / debugVelocitySearch / getDebuggingSamples / computeVelocityCommand and the
// TrajectorySamplerParameters constructor (ref: dwa.h:34-41,101-165, tests/dwa_test.cpp:88-90)
static void test_dwa_surface() {
  Control::ControlLimitsParams lim(Control::LinearVelocityControlParams(1.0, 5.0, 10.0),
                                   Control::LinearVelocityControlParams(0.0, 0.0, 0.0),
                                   Control::AngularVelocityControlParams(M_PI, 4.0, 3.0, 3.0));
  Control::CostEvaluator::TrajectoryCostsWeights w;
  Control::TrajectorySampler::TrajectorySamplerParameters cfg;
  cfg.setParameter("time_step", 0.1);
  cfg.setParameter("prediction_horizon", 1.0);
  cfg.setParameter("control_horizon", 0.2);
  cfg.setParameter("max_linear_samples", 20);
  cfg.setParameter("max_angular_samples", 20);
  bool threw = false;
  try {
    cfg.setParameter("max_linear_samples", 5000);
  } catch (const std::out_of_range &) {
    threw = true;  // ref parameter.h:134-146
  }
  CHECK(threw);
  Control::DWA planner(cfg, lim, Control::ControlType::DIFFERENTIAL_DRIVE, CollisionChecker::ShapeType::CYLINDER,
                       {0.1f, 0.4f}, {0, 0, 0}, {0, 0, 0, 1}, w);
  threw = false;
  try {
    planner.getDebuggingSamples();
  } catch (const std::invalid_argument &) {
    threw = true;  // ref dwa.cpp:236-238
  }
  CHECK(threw);
  ::Path::Path path(std::vector<::Path::Point>{{0.0f, 0.0f, 0.0f}, {1.0f, 0.0f, 0.0f}, {2.0f, 0.0f, 0.0f}});
  planner.setCurrentPath(path);
  planner.setCurrentState(::Path::State(-0.5, 0.0, 0.0, 0.0));
  std::vector<double> ranges, angles;
  initLaserscan(360, 5.0, ranges, angles);
  const Control::LaserScan scan(ranges, angles);
  const Control::Velocity2D vel(0.2, 0.0, 0.0);
  planner.debugVelocitySearch(vel, scan, true);
  const Control::TrajectorySamples2D dbg = planner.getDebuggingSamplesPure();
  auto [px, py] = planner.getDebuggingSamples();
  CHECK(dbg.size() > 0 && px.rows() == dbg.size() && px.cols() == 10 && py.rows() == dbg.size());
  CHECK(px(0, 0) == -0.5f && py(0, 0) == 0.0f);

  const Control::TrajSearchResult plain = planner.computeVelocityCommandsSet(vel, scan);
  CHECK(plain.isTrajFound);
  const Control::Controller::Result one = planner.computeVelocityCommand(vel, scan);
  CHECK(one.status == Control::Controller::Result::Status::COMMAND_FOUND);
  CHECK(one.velocity_command.vx() == (double)plain.trajectory.velocities.vx[0]);

  // a custom cost that rewards turning left: huge weight on -omega makes the most positive omega win
  size_t calls = 0, path_points = 0;
  planner.addCustomCost(100.0, [&](const Control::Trajectory2D &t, const ::Path::Path &p) {
    ++calls;
    path_points = p.getSize();
    return 10.0f - t.velocities.omega[0];
  });
  const Control::TrajSearchResult steered = planner.computeVelocityCommandsSet(vel, scan);
  CHECK(steered.isTrajFound);
  CHECK(calls == dbg.size());          // one call per admissible trajectory
  CHECK(path_points == 201);           // the interpolated path (2 m at 0.01 m)
  float max_om = -1e9f;
  for (size_t i = 0; i < dbg.size(); ++i) max_om = std::max(max_om, dbg.velocities.omega(i, 0));
  CHECK(steered.trajectory.velocities.omega[0] == max_om);
  CHECK(steered.trajCost > plain.trajCost);
  // an exception thrown by the callback surfaces from the compute call
  planner.addCustomCost(1.0, [](const Control::Trajectory2D &, const ::Path::Path &) -> float {
    throw std::runtime_error("boom");
  });
  threw = false;
  try {
    planner.computeVelocityCommandsSet(vel, scan);
  } catch (const std::runtime_error &e) {
    threw = std::string(e.what()) == "boom";
  }
  CHECK(threw);
}

// ref: tests/collisions_test.cpp:11-77 (test_FCL) against the stand-alone CollisionChecker, plus the
// batched state check and TrajectorySampler::checkStatesFeasibility
static void test_collision_checker() {
  // Eigen::Quaternionf{0,0,0,1} in the reference test is (w,x,y,z): coefficients (x,y,z,w) = (0,0,1,0)
  CollisionChecker checker(CollisionChecker::ShapeType::BOX, {0.4f, 0.4f, 1.0f}, {0.0f, 0.0f, 1.0f},
                           {0, 0, 1, 0}, 0.1);
  CHECK(near(checker.getRadius(), std::sqrt(0.32f) / 2.0f, 1e-6f));
  checker.updateState(::Path::State(0.0, 0.0, 0.0, 0.0));
  CHECK(!checker.checkCollisions({1.0, 1.0, 1.0}, {0.0, 0.1, 0.2}));
  checker.updateState(::Path::State(3.0, 5.0, 0.0, 0.0));
  CHECK(checker.checkCollisions({0.25, 0.5, 0.5}, {0.0, 0.1, 0.2}));
  std::vector<::Path::Point> cloud{{3.1f, 5.1f, -0.5f}};
  checker.updateSensorData(cloud, true);
  CHECK(checker.checkCollisions());
  CHECK(!checker.checkCollisions(::Path::State(0.0, 0.0, 0.0)));
  std::vector<::Path::State> states{::Path::State(0, 0, 0), ::Path::State(3.0, 5.0, 0.3), ::Path::State(9, 9, 0)};
  std::vector<uint8_t> per;
  CHECK(checker.checkStates(states, &per));
  CHECK(per.size() == 3 && per[0] == 0 && per[1] == 1 && per[2] == 0);

  Control::ControlLimitsParams lim(Control::LinearVelocityControlParams(1.0, 5.0, 10.0),
                                   Control::LinearVelocityControlParams(0.0, 0.0, 0.0),
                                   Control::AngularVelocityControlParams(M_PI, 4.0, 3.0, 3.0));
  Control::TrajectorySampler sampler(lim, Control::ControlType::DIFFERENTIAL_DRIVE, 0.1, 1.0, 0.2, 20, 20,
                                     CollisionChecker::ShapeType::CYLINDER, {0.1f, 0.4f}, {0, 0, 0}, {0, 0, 0, 1}, 0.1);
  sampler.updateState(::Path::State(0, 0, 0));
  std::vector<::Path::Point> wall{{1.0f, 0.0f, 0.0f}};
  CHECK(!sampler.checkStatesFeasibility({::Path::State(0, 0, 0), ::Path::State(0.5, 0, 0)}, wall));
  CHECK(sampler.checkStatesFeasibility({::Path::State(0, 0, 0), ::Path::State(0.95, 0, 0)}, wall));
}

static void test_mapper() {
  Mapping::LocalMapperGPU mapper(100, 120, 0.1f, {0.0f, 0.0f, 0.0f}, 0.0f, false, 360, 0.01f, 2.0f, 0.0f, 20.0f, 256);
  std::vector<double> ranges, angles;
  initLaserscan(360, 3.0, ranges, angles);
  MatrixXi &g = mapper.scanToGrid(angles, ranges);
  CHECK(g.rows() == 100 && g.cols() == 120);
  long occ = 0, emp = 0, unk = 0;
  for (size_t i = 0; i < g.rows(); ++i)
    for (size_t j = 0; j < g.cols(); ++j) {
      const int v = g(i, j);
      occ += v == 100;
      emp += v == 0;
      unk += v == -1;
    }
  CHECK(occ > 0 && emp > 0 && occ + emp + unk == 100 * 120);
  CHECK(g(49, 59) == 0);  // central cell = round(H/2)-1, round(W/2)-1 is swept free
  // Bayesian update (tests/mapper_test.cpp:137-215 only logs the grids; assert their invariants)
  auto [gb, pb] = mapper.scanToGridBaysian(angles, ranges);
  long touched = 0;
  bool prior_kept = true;
  for (size_t i = 0; i < gb.rows(); ++i)
    for (size_t j = 0; j < gb.cols(); ++j) {
      if (gb(i, j) == -1)
        prior_kept = prior_kept && pb(i, j) == 0.5f;
      else
        touched += pb(i, j) != 0.5f;
    }
  CHECK(prior_kept && touched > 0);
  mapper.getPreviousGridInCurrentPose({0.2f, 0.1f}, 0.3);
}

int main() {
  std::printf("accelerators: %s", getAvailableAccelerators().c_str());
  test_critical_zone();
  test_cost_evaluator();
  test_dwa_and_sampler();
  test_dwa_closed_loop();
  test_dwa_surface();
  test_collision_checker();
  test_mapper();
  std::printf("%s (%d failures)\n", failures ? "FAILED" : "ALL PASSED", failures);
  return failures ? 1 : 0;
}
