/*
 * ref_follower_bridge.cpp — TEST INFRASTRUCTURE ONLY (part of oracle/_ref/libkompass_ref.so): the
 * reference's own DWA controller object (Follower + DWA + TrajectorySampler + CostEvaluator, compiled
 * from /root/reference where it lies) behind a handle, so that tests/test_oracle_vs_ref.py can drive
 * the closed-loop scenarios of the reference's tests/dwa_test.cpp through it in lock-step with the
 * follower oracle (tests/orc_follower.py + the port). Built with -fno-access-control so the tracked
 * state the comparison needs (closest position, tracked segment, adapted horizon) can be read without
 * touching the reference sources.
 */
#include "kompass_oracle.h"

#include <memory>
#include <vector>

#include "controllers/dwa.h"

using namespace Kompass;

namespace {
struct RefDwa {
  std::unique_ptr<Control::DWA> dwa;
};
}  // namespace

extern "C" {

typedef struct orc_ref_dwa_info {
  int32_t closest_index, segment_index, seg_start, seg_count, n_points, found;
  float cost;
  float _pad;
  double segment_position, crosstrack_error, heading_error, horizon;
  double cmd[3];
} orc_ref_dwa_info;

void *orc_ref_dwa_create(const orc_sampler_cfg *c, const orc_cost_cfg *w, double interp, double seg_len,
                         double goal_tol, double loosing, double kappa_tol, float max_local_range) {
  Control::TrajectorySampler::TrajectorySamplerParameters p;
  p.setParameter("time_step", c->time_step);
  p.setParameter("prediction_horizon", c->prediction_horizon);
  p.setParameter("control_horizon", c->control_horizon);
  p.setParameter("max_linear_samples", (int)c->max_linear_samples);
  p.setParameter("max_angular_samples", (int)c->max_angular_samples);
  p.setParameter("octree_map_resolution", c->octree_resolution);
  p.setParameter("drop_samples", c->drop_samples != 0);
  Control::LinearVelocityControlParams x(c->vx_max, c->vx_acc, c->vx_dec), y(c->vy_max, c->vy_acc, c->vy_dec);
  Control::AngularVelocityControlParams a(M_PI, c->omega_max, c->omega_acc, c->omega_dec);
  Control::ControlLimitsParams lim(x, y, a);
  Control::CostEvaluator::TrajectoryCostsWeights cw;
  cw.setParameter("reference_path_distance_weight", w->w_path);
  cw.setParameter("goal_distance_weight", w->w_goal);
  cw.setParameter("obstacles_distance_weight", w->w_obstacles);
  cw.setParameter("smoothness_weight", w->w_smooth);
  cw.setParameter("jerk_weight", w->w_jerk);
  const CollisionChecker::ShapeType shape =
      c->robot_shape == ORC_CYLINDER ? CollisionChecker::ShapeType::CYLINDER
                                     : (c->robot_shape == ORC_BOX ? CollisionChecker::ShapeType::BOX
                                                                  : CollisionChecker::ShapeType::SPHERE);
  std::vector<float> dims;
  if (c->robot_shape == ORC_CYLINDER)
    dims = {c->robot_dims[0], c->robot_dims[1]};
  else if (c->robot_shape == ORC_BOX)
    dims = {c->robot_dims[0], c->robot_dims[1], c->robot_dims[2]};
  else
    dims = {c->robot_dims[0]};
  const Control::ControlType ct = c->control_type == ORC_ACKERMANN
                                      ? Control::ControlType::ACKERMANN
                                      : (c->control_type == ORC_OMNI ? Control::ControlType::OMNI
                                                                     : Control::ControlType::DIFFERENTIAL_DRIVE);
  auto *h = new RefDwa();
  h->dwa = std::make_unique<Control::DWA>(
      p, lim, ct, shape, dims, Eigen::Vector3f(c->sensor_position[0], c->sensor_position[1], c->sensor_position[2]),
      Eigen::Vector4f(c->sensor_rotation[0], c->sensor_rotation[1], c->sensor_rotation[2], c->sensor_rotation[3]), cw, 1);
  Control::Follower::FollowerParameters fp;
  fp.setParameter("max_point_interpolation_distance", interp);
  fp.setParameter("path_segment_length", seg_len);
  fp.setParameter("goal_dist_tolerance", goal_tol);
  fp.setParameter("loosing_goal_distance", loosing);
  fp.setParameter("curvature_horizon_tolerance", kappa_tol);
  h->dwa->setParams(fp);
  h->dwa->setSensorMaxRange(max_local_range);
  return h;
}

void orc_ref_dwa_destroy(void *hp) { delete static_cast<RefDwa *>(hp); }

int32_t orc_ref_dwa_set_path(void *hp, const float *x, const float *y, int32_t n) {
  auto *h = static_cast<RefDwa *>(hp);
  std::vector<Path::Point> pts;
  for (int32_t i = 0; i < n; ++i) pts.emplace_back(x[i], y[i], 0.0f);
  h->dwa->setCurrentPath(Path::Path(pts));
  return (int32_t)h->dwa->getCurrentPath().getSize();
}

void orc_ref_dwa_set_state(void *hp, double x, double y, double yaw) {
  static_cast<RefDwa *>(hp)->dwa->setCurrentState(Path::State(x, y, yaw));
}

int32_t orc_ref_dwa_goal_reached(void *hp) { return static_cast<RefDwa *>(hp)->dwa->isGoalReached() ? 1 : 0; }

/* DWA::computeVelocityCommandsSet(vel, cloud); rows: vx,vy,omega [P-1] each then x,y [P] each */
int32_t orc_ref_dwa_compute_cloud(void *hp, const double vel[3], const float *xyz, int32_t n, orc_ref_dwa_info *info,
                                  float *rows, int32_t rows_cap) {
  auto *h = static_cast<RefDwa *>(hp);
  Control::DWA &d = *h->dwa;
  std::vector<Path::Point> cloud;
  cloud.reserve(n);
  for (int32_t i = 0; i < n; ++i) cloud.emplace_back(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  const Control::TrajSearchResult res = d.computeVelocityCommandsSet(Control::Velocity2D(vel[0], vel[1], vel[2]), cloud);
  const Path::Path::View seg = d.findTrackedPathSegment();
  info->closest_index = (int32_t)d.closestPosition->index;
  info->segment_index = (int32_t)d.current_segment_index_;
  info->seg_start = (int32_t)seg.getStartIndex();
  info->seg_count = (int32_t)seg.getSize();
  info->n_points = (int32_t)d.trajSampler->numPointsPerTrajectory;
  info->found = res.isTrajFound ? 1 : 0;
  info->cost = res.trajCost;
  info->segment_position = d.closestPosition->segment_length;
  info->crosstrack_error = d.closestPosition->parallel_distance;
  info->heading_error = d.currentTrackedTarget_->heading_error;
  info->horizon = d.trajSampler->max_time_;
  info->cmd[0] = d.getLinearVelocityCmdX();
  info->cmd[1] = d.getLinearVelocityCmdY();
  info->cmd[2] = d.getAngularVelocityCmd();
  if (res.isTrajFound && rows) {
    const int32_t P = (int32_t)res.trajectory.path.x.size();
    if (5 * P > rows_cap) return -1;
    for (int32_t j = 0; j + 1 < P; ++j) {
      rows[j] = res.trajectory.velocities.vx(j);
      rows[(P - 1) + j] = res.trajectory.velocities.vy(j);
      rows[2 * (P - 1) + j] = res.trajectory.velocities.omega(j);
    }
    for (int32_t j = 0; j < P; ++j) {
      rows[3 * (P - 1) + j] = res.trajectory.path.x(j);
      rows[3 * (P - 1) + P + j] = res.trajectory.path.y(j);
    }
    return P;
  }
  return 0;
}

}  // extern "C"
