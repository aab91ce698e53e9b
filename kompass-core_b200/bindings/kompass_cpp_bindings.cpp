// kompass_cpp_bindings.cpp — Python extension module `kompass_cpp` for the hot-path classes, over
// the header-only C++ mirror (host/kompass_b200.hpp) and libkompass_b200.so.
//
// The reference binds these classes with nanobind (src/kompass_cpp/bindings/bindings*.cpp); nanobind
// is not in this image, pybind11 is, and the two share the API used here, so this file keeps the
// reference's module layout, class names, method names and keyword names:
//   kompass_cpp.types     State, Path, Velocity2D, LaserScan, TrajectoryVelocities2D, TrajectoryPath,
//                         Trajectory, RobotGeometry, SensorInputType, PointFieldType
//                                                       ref: bindings_types.cpp:30-186
//   kompass_cpp.configure ConfigParameters.from_dict     ref: bindings_config.cpp:9-54
//   kompass_cpp.control   ControlType, Linear/AngularVelocityControlParams, ControlLimitsParams,
//                         TrajectoryCostWeights, TrajectorySamplerParameters, SamplingControlResult,
//                         DWA (+ the Controller / Follower methods it inherits)
//                                                       ref: bindings_control.cpp:20-106,209-273
//   kompass_cpp.mapping   LocalMapperGPU.scan_to_grid x2 (+ Bayesian methods), OCCUPANCY_TYPE
//                                                       ref: bindings_gpu.cpp:13-37, bindings_mapping.cpp:17-80
//   kompass_cpp.utils     CriticalZoneCheckerGPU.check x2, CollisionChecker,
//                         pointcloud_to_laserscan_from_raw
//                                                       ref: bindings_gpu.cpp:42-68, bindings_utils.cpp:50-110
//   kompass_cpp.get_available_accelerators               ref: bindings.cpp:47
// Only the classes of the hot path exist here: this is the GPU backend of those classes, not the
// whole package.
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "../host/kompass_b200.hpp"

namespace py = pybind11;
using namespace Kompass;

// ref: include/utils/pointcloud.h:37-46
enum class PointFieldType : int {
  INT8 = KC_INT8, UINT8 = KC_UINT8, INT16 = KC_INT16, UINT16 = KC_UINT16,
  INT32 = KC_INT32, UINT32 = KC_UINT32, FLOAT32 = KC_FLOAT32, FLOAT64 = KC_FLOAT64
};

namespace {
// Eigen::MatrixXi / MatrixXf as numpy sees them through nanobind: [rows, cols], column-major memory
py::array_t<int32_t> gridToNumpy(const MatrixXi &g) {
  py::array_t<int32_t, py::array::f_style> a({g.rows(), g.cols()});
  std::copy(g.v.begin(), g.v.end(), a.mutable_data());
  return a;
}
py::array_t<float> gridToNumpy(const MatrixXf &g) {
  py::array_t<float, py::array::f_style> a({g.rows(), g.cols()});
  std::copy(g.v.begin(), g.v.end(), a.mutable_data());
  return a;
}
py::array_t<float> rowsToNumpy(const MatrixXfR &m) {
  py::array_t<float> a({m.rows(), m.cols()});
  std::copy(m.v.begin(), m.v.end(), a.mutable_data());
  return a;
}
std::vector<int8_t> rawBytes(const py::array &data) {  // the reference takes std::vector<int8_t>
  py::array_t<int8_t, py::array::c_style | py::array::forcecast> b(data);
  return std::vector<int8_t>(b.data(), b.data() + b.size());
}
// name -> value dictionaries for the two Parameters-derived classes (bindings_config.cpp:9-34)
template <class P>
void fromDict(P &params, const py::dict &d) {
  for (auto item : d) {
    const std::string name = py::cast<std::string>(item.first);
    try {
      double v;
      if (py::isinstance<py::bool_>(item.second))
        v = py::cast<bool>(item.second) ? 1.0 : 0.0;
      else if (py::isinstance<py::float_>(item.second) || py::isinstance<py::int_>(item.second))
        v = py::cast<double>(item.second);
      else
        continue;
      try {
        params.getParameter(name);
      } catch (const std::invalid_argument &) {
        continue;  // unknown names are ignored, as the reference does
      }
      params.setParameter(name, v);
    } catch (const std::exception &e) {
      throw std::runtime_error(e.what());  // bindings_config.cpp:28-32
    }
  }
}
}  // namespace

PYBIND11_MODULE(kompass_cpp, m) {
  m.doc() = "B200 (sm_100a) backend of the kompass_cpp hot-path classes";
  m.def("get_available_accelerators", &getAvailableAccelerators, "Get available accelerators");

  // ------------------------------------------------------------------------------------ types
  auto m_types = m.def_submodule("types", "KOMPASS CPP data types module");
  py::class_<::Path::State>(m_types, "State")
      .def(py::init<double, double, double, double>(), py::arg("x") = 0.0, py::arg("y") = 0.0,
           py::arg("yaw") = 0.0, py::arg("speed") = 0.0)
      .def_readwrite("x", &::Path::State::x)
      .def_readwrite("y", &::Path::State::y)
      .def_readwrite("yaw", &::Path::State::yaw)
      .def_readwrite("speed", &::Path::State::speed);
  py::class_<::Path::Path>(m_types, "Path")
      .def(py::init<const std::vector<::Path::Point> &>(), py::arg("points"))
      .def("get_total_length", &::Path::Path::totalPathLength)
      .def("size", &::Path::Path::getSize)
      .def("getIndex", &::Path::Path::getIndex, py::arg("index"))
      .def("x", &::Path::Path::getX)
      .def("y", &::Path::Path::getY);
  py::class_<Control::Velocity2D>(m_types, "Velocity2D")
      .def(py::init<>())
      .def(py::init<double, double, double, double>(), py::arg("vx") = 0.0, py::arg("vy") = 0.0,
           py::arg("omega") = 0.0, py::arg("steer_ang") = 0.0)
      .def_property("vx", &Control::Velocity2D::vx, &Control::Velocity2D::setVx)
      .def_property("vy", &Control::Velocity2D::vy, &Control::Velocity2D::setVy)
      .def_property("omega", &Control::Velocity2D::omega, &Control::Velocity2D::setOmega)
      .def_property("steer_ang", &Control::Velocity2D::steer_ang, &Control::Velocity2D::setSteerAng);
  py::class_<Control::TrajectoryVelocities2D>(m_types, "TrajectoryVelocities2D")
      .def_readonly("vx", &Control::TrajectoryVelocities2D::vx)
      .def_readonly("vy", &Control::TrajectoryVelocities2D::vy)
      .def_readonly("omega", &Control::TrajectoryVelocities2D::omega);
  py::class_<Control::TrajectoryPath>(m_types, "TrajectoryPath")
      .def_readonly("x", &Control::TrajectoryPath::x)
      .def_readonly("y", &Control::TrajectoryPath::y)
      .def_readonly("z", &Control::TrajectoryPath::z);
  py::class_<Control::Trajectory2D>(m_types, "Trajectory")
      .def(py::init<>())
      .def_readonly("velocities", &Control::Trajectory2D::velocities)
      .def_readonly("path", &Control::Trajectory2D::path);
  py::class_<Control::LaserScan>(m_types, "LaserScan")
      .def(py::init<std::vector<double>, std::vector<double>>(), py::arg("ranges"), py::arg("angles"))
      .def_readonly("ranges", &Control::LaserScan::ranges)
      .def_readonly("angles", &Control::LaserScan::angles);
  py::enum_<CollisionChecker::ShapeType>(m_types, "RobotGeometry")
      .value("CYLINDER", CollisionChecker::ShapeType::CYLINDER)
      .value("BOX", CollisionChecker::ShapeType::BOX)
      .value("SPHERE", CollisionChecker::ShapeType::SPHERE)
      .def_static("get", [](const std::string &key) {
        if (key == "CYLINDER") return CollisionChecker::ShapeType::CYLINDER;
        if (key == "BOX") return CollisionChecker::ShapeType::BOX;
        if (key == "SPHERE") return CollisionChecker::ShapeType::SPHERE;
        throw std::runtime_error("Invalid key");
      });
  py::enum_<CriticalZoneChecker::InputType>(m_types, "SensorInputType")
      .value("LASERSCAN", CriticalZoneChecker::InputType::LASERSCAN)
      .value("POINTCLOUD", CriticalZoneChecker::InputType::POINTCLOUD)
      .def_static("get", [](const std::string &key) {
        if (key == "LASERSCAN") return CriticalZoneChecker::InputType::LASERSCAN;
        if (key == "POINTCLOUD") return CriticalZoneChecker::InputType::POINTCLOUD;
        throw std::runtime_error("Invalid key");
      });
  py::enum_<PointFieldType>(m_types, "PointFieldType")
      .value("INT8", PointFieldType::INT8)
      .value("UINT8", PointFieldType::UINT8)
      .value("INT16", PointFieldType::INT16)
      .value("UINT16", PointFieldType::UINT16)
      .value("INT32", PointFieldType::INT32)
      .value("UINT32", PointFieldType::UINT32)
      .value("FLOAT32", PointFieldType::FLOAT32)
      .value("FLOAT64", PointFieldType::FLOAT64)
      .export_values()
      .def_static(
          "from_int",
          [](int value) {
            if (value < 1 || value > 8)
              throw std::invalid_argument("Invalid integer for PointFieldType. Must be between 1 and 8.");
            return static_cast<PointFieldType>(value);
          },
          py::arg("value"));

  // ------------------------------------------------------------------------------------ control
  auto m_control = m.def_submodule("control", "Control module");
  py::enum_<Control::ControlType>(m_control, "ControlType")
      .value("ACKERMANN", Control::ControlType::ACKERMANN)
      .value("DIFFERENTIAL_DRIVE", Control::ControlType::DIFFERENTIAL_DRIVE)
      .value("OMNI", Control::ControlType::OMNI);
  py::class_<Control::LinearVelocityControlParams>(m_control, "LinearVelocityControlParams")
      .def(py::init<double, double, double>(), py::arg("max_vel") = 0.0, py::arg("max_acc") = 0.0,
           py::arg("max_decel") = 0.0)
      .def_readwrite("max_vel", &Control::LinearVelocityControlParams::maxVel)
      .def_readwrite("max_acc", &Control::LinearVelocityControlParams::maxAcceleration)
      .def_readwrite("max_decel", &Control::LinearVelocityControlParams::maxDeceleration);
  py::class_<Control::AngularVelocityControlParams>(m_control, "AngularVelocityControlParams")
      .def(py::init<double, double, double, double>(), py::arg("max_ang") = M_PI,
           py::arg("max_omega") = 0.0, py::arg("max_acc") = 0.0, py::arg("max_decel") = 0.0)
      .def_readwrite("max_steer_ang", &Control::AngularVelocityControlParams::maxAngle)
      .def_readwrite("max_omega", &Control::AngularVelocityControlParams::maxOmega)
      .def_readwrite("max_acc", &Control::AngularVelocityControlParams::maxAcceleration)
      .def_readwrite("max_decel", &Control::AngularVelocityControlParams::maxDeceleration);
  py::class_<Control::ControlLimitsParams>(m_control, "ControlLimitsParams")
      .def(py::init<>())
      .def(py::init<const Control::LinearVelocityControlParams &, const Control::LinearVelocityControlParams &,
                    const Control::AngularVelocityControlParams &>(),
           py::arg("vel_x_ctr_params") = Control::LinearVelocityControlParams(),
           py::arg("vel_y_ctr_params") = Control::LinearVelocityControlParams(),
           py::arg("omega_ctr_params") = Control::AngularVelocityControlParams())
      .def_readwrite("linear_x_limits", &Control::ControlLimitsParams::velXParams)
      .def_readwrite("linear_y_limits", &Control::ControlLimitsParams::velYParams)
      .def_readwrite("angular_limits", &Control::ControlLimitsParams::omegaParams);
  py::class_<Control::TrajSearchResult>(m_control, "SamplingControlResult")
      .def(py::init<>())
      .def_readwrite("is_found", &Control::TrajSearchResult::isTrajFound)
      .def_readwrite("cost", &Control::TrajSearchResult::trajCost)
      .def_readwrite("trajectory", &Control::TrajSearchResult::trajectory);
  using Weights = Control::CostEvaluator::TrajectoryCostsWeights;
  py::class_<Weights>(m_control, "TrajectoryCostWeights")
      .def(py::init<>())
      .def("from_dict", [](Weights &w, const py::dict &d) { fromDict(w, d); })
      .def("set_parameter", &Weights::setParameter)
      .def("get_parameter", [](const Weights &w, const std::string &n) { return w.getParameter<double>(n); });
  using SamplerParams = Control::TrajectorySampler::TrajectorySamplerParameters;
  py::class_<SamplerParams>(m_control, "TrajectorySamplerParameters")
      .def(py::init<>())
      .def("from_dict", [](SamplerParams &p, const py::dict &d) { fromDict(p, d); })
      .def("set_parameter", &SamplerParams::setParameter)
      .def("get_parameter", [](const SamplerParams &p, const std::string &n) { return p.getParameter<double>(n); });

  py::class_<Control::DWA>(m_control, "DWA")
      .def(py::init<Control::ControlLimitsParams, Control::ControlType, double, double, double, int, int,
                    CollisionChecker::ShapeType, std::vector<float>, const Vector3f &, const Vector4f &, double,
                    Weights, int>(),
           py::arg("control_limits"), py::arg("control_type"), py::arg("time_step"),
           py::arg("prediction_horizon"), py::arg("control_horizon"), py::arg("max_linear_samples"),
           py::arg("max_angular_samples"), py::arg("robot_shape_type"), py::arg("robot_dimensions"),
           py::arg("sensor_position_robot"), py::arg("sensor_rotation_robot"), py::arg("octree_resolution"),
           py::arg("cost_weights"), py::arg("max_num_threads") = 1)
      .def(py::init<SamplerParams, Control::ControlLimitsParams, Control::ControlType,
                    CollisionChecker::ShapeType, std::vector<float>, const Vector3f &, const Vector4f &, Weights,
                    int>(),
           py::arg("config"), py::arg("control_limits"), py::arg("control_type"), py::arg("robot_shape_type"),
           py::arg("robot_dimensions"), py::arg("sensor_position_robot"), py::arg("sensor_rotation_robot"),
           py::arg("cost_weights"), py::arg("max_num_threads") = 1)
      // Controller / Follower surface (bindings_control.cpp:66-106)
      .def("set_linear_ctr_limits", &Control::DWA::setLinearControlLimits)
      .def("set_angular_ctr_limits", &Control::DWA::setAngularControlLimits)
      .def("set_current_state", &Control::DWA::setCurrentState)
      .def("set_current_state", [](Control::DWA &d, double x, double y, double yaw,
                                   double speed) { d.setCurrentState(::Path::State(x, y, yaw, speed)); })
      .def("set_current_path", &Control::DWA::setCurrentPath, py::arg("path"), py::arg("interpolate") = true)
      .def("clear_current_path", &Control::DWA::clearCurrentPath)
      .def("is_goal_reached", &Control::DWA::isGoalReached)
      .def("has_path", &Control::DWA::hasPath)
      .def("get_vx_cmd", &Control::DWA::getLinearVelocityCmdX)
      .def("get_vy_cmd", &Control::DWA::getLinearVelocityCmdY)
      .def("get_omega_cmd", &Control::DWA::getAngularVelocityCmd)
      // DWA surface (bindings_control.cpp:239-273)
      .def("compute_velocity_commands",
           py::overload_cast<const Control::Velocity2D &, const Control::LaserScan &>(
               &Control::DWA::computeVelocityCommandsSet))
      .def("compute_velocity_commands",
           py::overload_cast<const Control::Velocity2D &, const std::vector<::Path::Point> &>(
               &Control::DWA::computeVelocityCommandsSet))
      .def("add_custom_cost", &Control::DWA::addCustomCost)
      .def("get_debugging_samples",
           [](const Control::DWA &d) {
             auto s = d.getDebuggingSamples();
             return py::make_tuple(rowsToNumpy(std::get<0>(s)), rowsToNumpy(std::get<1>(s)));
           })
      .def("debug_velocity_search",
           py::overload_cast<const Control::Velocity2D &, const std::vector<::Path::Point> &, const bool &>(
               &Control::DWA::debugVelocitySearch))
      .def("debug_velocity_search",
           py::overload_cast<const Control::Velocity2D &, const Control::LaserScan &, const bool &>(
               &Control::DWA::debugVelocitySearch))
      .def("set_resolution", &Control::DWA::resetOctreeResolution);

  // ------------------------------------------------------------------------------------ mapping
  auto m_mapping = m.def_submodule("mapping", "Local mapping module");
  py::enum_<Mapping::OccupancyType>(m_mapping, "OCCUPANCY_TYPE")
      .value("UNEXPLORED", Mapping::OccupancyType::UNEXPLORED)
      .value("EMPTY", Mapping::OccupancyType::EMPTY)
      .value("OCCUPIED", Mapping::OccupancyType::OCCUPIED);
  py::class_<Mapping::LocalMapperGPU>(m_mapping, "LocalMapperGPU")
      .def(py::init<const int, const int, float, const Vector3f &, float, bool, int, float, float, float, float,
                    int>(),
           py::arg("grid_height"), py::arg("grid_width"), py::arg("resolution"), py::arg("laserscan_position"),
           py::arg("laserscan_orientation"), py::arg("is_pointcloud"), py::arg("scan_size"),
           py::arg("angle_step"), py::arg("max_height"), py::arg("min_height"), py::arg("range_max"),
           py::arg("max_points_per_line") = 32)
      .def("scan_to_grid",
           [](Mapping::LocalMapperGPU &s, const std::vector<double> &angles, const std::vector<double> &ranges) {
             return gridToNumpy(s.scanToGrid(angles, ranges));
           },
           "Convert laser scan data to occupancy grid", py::arg("angles"), py::arg("ranges"))
      .def("scan_to_grid",
           [](Mapping::LocalMapperGPU &s, const py::array &data, int point_step, int row_step, int height,
              int width, float x_offset, float y_offset, float z_offset) {
             return gridToNumpy(s.scanToGrid(rawBytes(data), point_step, row_step, height, width, x_offset,
                                             y_offset, z_offset));
           },
           "Convert raw point cloud data to occupancy grid", py::arg("data"), py::arg("point_step"),
           py::arg("row_step"), py::arg("height"), py::arg("width"), py::arg("x_offset"), py::arg("y_offset"),
           py::arg("z_offset"))
      .def("set_bayesian_parameters", &Mapping::LocalMapperGPU::setBayesianParams, py::arg("p_prior") = 0.5f,
           py::arg("p_occupied") = 0.6f, py::arg("p_empty") = 0.4f, py::arg("range_sure") = 1.0f,
           py::arg("wall_size") = 0.2f)
      .def("scan_to_grid_baysian",
           [](Mapping::LocalMapperGPU &s, const std::vector<double> &angles, const std::vector<double> &ranges) {
             auto r = s.scanToGridBaysian(angles, ranges);
             return py::make_tuple(gridToNumpy(std::get<0>(r)), gridToNumpy(std::get<1>(r)));
           },
           py::arg("angles"), py::arg("ranges"))
      .def("scan_to_grid_baysian",
           [](Mapping::LocalMapperGPU &s, const py::array &data, int point_step, int row_step, int height,
              int width, float x_offset, float y_offset, float z_offset) {
             auto r = s.scanToGridBaysian(rawBytes(data), point_step, row_step, height, width, x_offset,
                                          y_offset, z_offset);
             return py::make_tuple(gridToNumpy(std::get<0>(r)), gridToNumpy(std::get<1>(r)));
           },
           py::arg("data"), py::arg("point_step"), py::arg("row_step"), py::arg("height"), py::arg("width"),
           py::arg("x_offset"), py::arg("y_offset"), py::arg("z_offset"))
      .def("get_previous_grid_in_current_pose",
           [](Mapping::LocalMapperGPU &s, const std::array<float, 2> &pos, double yaw) {
             s.getPreviousGridInCurrentPose(pos, yaw);
           },
           py::arg("current_position_in_previous_pose"), py::arg("current_orientation_in_previous_pose"));

  // ------------------------------------------------------------------------------------ utils
  auto m_utils = m.def_submodule("utils", "KOMPASS CPP utilities module");
  py::class_<CriticalZoneCheckerGPU>(m_utils, "CriticalZoneCheckerGPU")
      .def(py::init([](CriticalZoneChecker::InputType input_type, CollisionChecker::ShapeType robot_shape,
                       const std::vector<float> &robot_dimensions, const Vector3f &sensor_position_body,
                       const Vector4f &sensor_rotation_body, float critical_angle, float critical_distance,
                       float slowdown_distance, const std::vector<double> &scan_angles, float min_height,
                       float max_height, float range_max, PointFieldType cloud_field_type) {
             return new CriticalZoneCheckerGPU(input_type, robot_shape, robot_dimensions, sensor_position_body,
                                               sensor_rotation_body, critical_angle, critical_distance,
                                               slowdown_distance, scan_angles, min_height, max_height,
                                               range_max, static_cast<int>(cloud_field_type));
           }),
           py::arg("input_type"), py::arg("robot_shape"), py::arg("robot_dimensions"),
           py::arg("sensor_position_body"), py::arg("sensor_rotation_body"), py::arg("critical_angle"),
           py::arg("critical_distance"), py::arg("slowdown_distance"), py::arg("scan_angles"),
           py::arg("min_height"), py::arg("max_height"), py::arg("range_max"),
           py::arg("cloud_field_type") = PointFieldType::FLOAT32)
      .def("check", py::overload_cast<const std::vector<double> &, bool>(&CriticalZoneCheckerGPU::check),
           py::arg("ranges"), py::arg("forward"))
      .def("check",
           [](CriticalZoneCheckerGPU &s, const py::array &data, int point_step, int row_step, int height,
              int width, int x_offset, int y_offset, int z_offset, bool forward) {
             return s.check(rawBytes(data), point_step, row_step, height, width, x_offset, y_offset, z_offset,
                            forward);
           },
           py::arg("data"), py::arg("point_step"), py::arg("row_step"), py::arg("height"), py::arg("width"),
           py::arg("x_offset"), py::arg("y_offset"), py::arg("z_offset"), py::arg("forward"));
  py::class_<CollisionChecker>(m_utils, "CollisionChecker")
      .def(py::init<CollisionChecker::ShapeType, const std::vector<float> &, const Vector3f &, const Vector4f &,
                    double>(),
           py::arg("robot_shape"), py::arg("robot_dimensions"), py::arg("sensor_position_body"),
           py::arg("sensor_rotation_body"), py::arg("octree_resolution") = 0.01)
      .def("reset_octree_resolution", &CollisionChecker::resetOctreeResolution)
      .def("get_radius", &CollisionChecker::getRadius)
      .def("update_state", [](CollisionChecker &c, double x, double y, double yaw) { c.updateState(x, y, yaw); })
      .def("update_state", [](CollisionChecker &c, const ::Path::State &s) { c.updateState(s.x, s.y, s.yaw); })
      .def("update_sensor_data",
           [](CollisionChecker &c, const Control::LaserScan &scan) { c.updateSensorData(scan); })
      .def("update_sensor_data",
           [](CollisionChecker &c, const std::vector<::Path::Point> &cloud, bool global_frame) {
             c.updateSensorData(cloud, global_frame);
           },
           py::arg("cloud"), py::arg("global_frame") = true)
      .def("check_collisions", [](CollisionChecker &c) { return c.checkCollisions(); })
      .def("check_collisions", [](CollisionChecker &c, const ::Path::State &s) { return c.checkCollisions(s); })
      .def("check_states", [](CollisionChecker &c, const std::vector<::Path::State> &states) {
        std::vector<uint8_t> per;
        const bool any = c.checkStates(states, &per);
        return py::make_tuple(any, per);
      });
  m_utils.def(
      "pointcloud_to_laserscan_from_raw",
      [](const py::array &data, int point_step, int row_step, int height, int width, int x_offset, int y_offset,
         int z_offset, double max_range, double min_z, double max_z, double angle_step) {
        const std::vector<int8_t> raw = rawBytes(data);
        const int cap = static_cast<int>(std::ceil(2.0 * M_PI / angle_step)) + 2;
        std::vector<double> ranges(cap), angles(cap);
        int32_t n = 0;
        kcThrow(kc_pointcloud_to_laserscan_step(raw.data(), (int64_t)raw.size(), point_step, row_step, height,
                                                width, x_offset, y_offset, z_offset, max_range, min_z, max_z,
                                                angle_step, cap, ranges.data(), angles.data(), &n));
        ranges.resize(n);
        angles.resize(n);
        return py::make_tuple(ranges, angles);
      },
      py::arg("data"), py::arg("point_step"), py::arg("row_step"), py::arg("height"), py::arg("width"),
      py::arg("x_offset"), py::arg("y_offset"), py::arg("z_offset"), py::arg("max_range"), py::arg("min_z"),
      py::arg("max_z"), py::arg("angle_step"));
}
