/*
 * kompass_oracle.h — C API of the CPU parity ORACLE (test infrastructure, NOT product code).
 *
 * This library is a from-scratch CPU restatement of the reference CPU path of
 * automatika-robotics/kompass-core (v0.8.1) for the DWA / local-mapper /
 * critical-zone hot path. It exists only so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can check and time the CUDA
 * path against it. Nothing under kompass-core_b200/ may include, link or call it.
 *
 * Parity status:
 *   - costs, path prep, critical zone, cloud binning: PINNED against the reference's
 *     known-answer tests (tests/test_oracle_kat.py restates every case of
 *     src/kompass_cpp/tests/cost_evaluator_test.cpp and critical_zone_test.cpp).
 *   - mapper grid cells: invariants only (the reference has no golden grid).
 *   - collision (FCL 0.7.0 + octomap, third-party, absent from the reference tree):
 *     restated analytically; pinned only by the 3 booleans of collisions_test.cpp.
 *
 * All "ref:" citations are relative to /root/reference/src/kompass_cpp/kompass_cpp/.
 */
#ifndef KOMPASS_ORACLE_H
#define KOMPASS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ref: include/datatypes/control.h:14  enum ControlType */
enum { ORC_ACKERMANN = 0, ORC_DIFFERENTIAL_DRIVE = 1, ORC_OMNI = 2 };
/* ref: include/utils/collision_check.h:25 enum ShapeType */
enum { ORC_CYLINDER = 0, ORC_BOX = 1, ORC_SPHERE = 2 };

typedef struct orc_sampler_cfg {
  int32_t control_type;
  double time_step;          /* ref: trajectory_sampler.h:25-29 */
  double prediction_horizon; /* current (possibly adapted) horizon */
  double control_horizon;
  int32_t max_linear_samples;
  int32_t max_angular_samples;
  /* ref: include/datatypes/control.h:181-232 ControlLimitsParams */
  double vx_max, vx_acc, vx_dec;
  double vy_max, vy_acc, vy_dec;
  double omega_max, omega_acc, omega_dec;
  int32_t robot_shape;
  float robot_dims[3];
  float sensor_position[3];
  float sensor_rotation[4]; /* quaternion coefficients x,y,z,w (Eigen coeffs order) */
  double octree_resolution;
  int32_t drop_samples;
  int64_t num_ctrl_points; /* ref: trajectory_sampler.cpp:88 (control_horizon/time_step) */
  int32_t max_num_threads; /* >1 => std::thread fan-out over samples (timing only) */
} orc_sampler_cfg;

typedef struct orc_cost_cfg {
  /* ref: include/utils/cost_evaluator.h:22-50 TrajectoryCostsWeights */
  double w_path, w_goal, w_obstacles, w_smooth, w_jerk;
  float acc_limits[3]; /* ref: src/utils/cost_evaluator.cpp:18-20 */
  float sensor_position[3];
  float sensor_rotation[4]; /* x,y,z,w */
} orc_cost_cfg;

/* ---- sizes (ref: include/datatypes/trajectory.h:19-51) ---- */
int64_t orc_num_trajectories(int32_t control_type, int32_t max_linear, int32_t max_angular);
int64_t orc_num_points(double time_step, double prediction_horizon);

/* ---- path prep (ref: src/datatypes/path.cpp:167-288, include/utils/spline.h linear) ----
 * Returns the interpolated size; fills X,Y,acc,curv (capacity cap). */
int32_t orc_path_interpolate_linear(const float *x, const float *y, int32_t n, double max_dist,
                                    float *X, float *Y, float *acc, float *curv, int32_t cap,
                                    float *total_length);
/* ref: src/datatypes/path.cpp:290-330 Path::segment (on an interpolated path). Returns #segments */
int32_t orc_path_segment(const float *acc, int32_t n, double segment_length,
                         int64_t max_points_per_segment, int32_t *seg_starts, int32_t cap);
/* ref: include/datatypes/path.h:85-91 View::totalSegmentLength */
float orc_segment_length(const float *X, const float *Y, int32_t start, int32_t count);

/* ---- velocity samples (ref: src/utils/trajectory_sampler.cpp:181-275,328-372) ----
 * Enumerates the (vx,vy,omega) slots in serial reference order. Returns count (<= cap). */
int32_t orc_velocity_samples(const orc_sampler_cfg *cfg, const double vel[3], double *vx, double *vy,
                             double *omega, int32_t cap);

/* ---- sampler (ref: trajectory_sampler.cpp:118-179,295-314; collision_check.{h,cpp}) ----
 * Outputs row-major [cap x (P-1)] velocities and [cap x P] paths (float), admissible rows only, in
 * enumeration order; slot_of_row[i] = enumeration index of row i. Returns admissible count, or <0. */
int32_t orc_sampler_generate_scan(const orc_sampler_cfg *cfg, const double vel[3],
                                  const double pose[3], const double *ranges, const double *angles,
                                  int32_t n, float *vx, float *vy, float *omega, float *x, float *y,
                                  int32_t *slot_of_row, int32_t cap);
int32_t orc_sampler_generate_cloud(const orc_sampler_cfg *cfg, const double vel[3],
                                   const double pose[3], const float *xyz, int32_t n, float *vx,
                                   float *vy, float *omega, float *x, float *y,
                                   int32_t *slot_of_row, int32_t cap);
/* single pose collision boolean (ref: collision_check.cpp:125-162). sensor data given as scan
 * (is_cloud=0: a=ranges,b=angles doubles) or cloud (is_cloud=1: a=xyz floats). */
int32_t orc_check_collision(const orc_sampler_cfg *cfg, const double sensor_pose[3],
                            const double query_pose[3], int32_t is_cloud, const void *a,
                            const void *b, int32_t n);
/* batched CollisionChecker::checkCollisions(state) after updateState(sensor_pose) +
 * updateSensorData(data, global_frame); out[n_states] may be NULL; returns any-collision (0/1) */
int32_t orc_check_collision_states(const orc_sampler_cfg *cfg, const double sensor_pose[3],
                                   int32_t is_cloud, int32_t global_frame, const void *a,
                                   const void *b, int32_t n, const double *states,
                                   int32_t n_states, uint8_t *out);

/* ---- cost evaluator ---- */
/* ref: include/utils/cost_evaluator.h:174-223 setPointScan */
void orc_cost_points_scan(const orc_cost_cfg *cfg, const double *ranges, const double *angles,
                          int32_t n, const double pose[3], float *ox, float *oy);
void orc_cost_points_cloud(const orc_cost_cfg *cfg, const float *xyz, int32_t n,
                           const double pose[3], float *ox, float *oy);
/* ref: src/utils/cost_evaluator.cpp:49-233 getMinTrajectoryCost.
 * n_obs == 0 => obstacle term skipped. custom may be NULL; else row-major [n_traj x n_custom]
 * doubles = weight_k * custom_cost_k, each added as `float += double` (cost_evaluator.cpp:96-100). costs_out may be NULL. Returns found (0/1). */
int32_t orc_cost_evaluate(const orc_cost_cfg *cfg, int32_t n_traj, int32_t P, const float *vx,
                          const float *vy, const float *omega, const float *x, const float *y,
                          const float *pathX, const float *pathY, const float *pathAcc,
                          int32_t path_n, float path_total_length, int32_t seg_start,
                          int32_t seg_count, const float *ox, const float *oy, int32_t n_obs,
                          float max_obstacles_dist, const double *custom, int32_t n_custom,
                          float *costs_out,
                          int32_t *best_idx, float *best_cost, int32_t n_threads);

/* ---- local mapper ---- */
/* ref: src/mapping/local_mapper.cpp:127-159,204-220; include/mapping/line_drawing.h:55-124.
 * grid is column-major int32 [H x W] (cell (i,j) at i + j*H). */
void orc_mapper_scan_to_grid(int32_t H, int32_t W, float resolution, const float laser_pos[3],
                             float laser_orientation, const double *angles, const double *ranges,
                             int32_t n, int32_t *grid);
/* ref: include/utils/pointcloud.h:205-259 (num_bins overload) */
/* Bayesian mapper + previous-grid warp (row f2); grids column-major [H x W] */
void orc_mapper_scan_to_grid_bayes(int32_t H, int32_t W, float resolution, const float laser_pos[3],
                                   float laser_orientation, float pPrior, float pOccupied,
                                   float pEmpty, float rangeSure, float rangeMax, float wallSize,
                                   const double *angles, const double *ranges, int32_t n,
                                   const float *prev, int32_t *grid, float *prob);
void orc_mapper_warp_previous(int32_t H, int32_t W, float resolution, float pPrior, float pos_x,
                              float pos_y, double orientation, const float *prev, float *out);
void orc_pointcloud_to_laserscan(const int8_t *data, int64_t nbytes, int32_t point_step,
                                 int32_t row_step, int32_t height, int32_t width, int32_t x_off,
                                 int32_t y_off, int32_t z_off, double max_range, double min_z,
                                 double max_z, int32_t num_bins, double *ranges_out);
/* ref: include/utils/pointcloud.h:116-177 (angle_step overload); returns the bin count */
int32_t orc_pointcloud_to_laserscan_step(const int8_t *data, int64_t nbytes, int32_t point_step,
                                         int32_t row_step, int32_t height, int32_t width,
                                         int32_t x_off, int32_t y_off, int32_t z_off,
                                         double max_range, double min_z, double max_z,
                                         double angle_step, double *ranges_out, double *angles_out);

/* ---- critical zone (ref: src/utils/critical_zone_check.cpp:13-131) ---- */
typedef struct orc_cz_cfg {
  int32_t robot_shape;
  float robot_dims[3];
  float sensor_position[3];
  float sensor_rotation[4]; /* x,y,z,w, NOT normalised (ref quirk q17) */
  float critical_angle;     /* degrees */
  float critical_distance;
  float slowdown_distance;
  float min_height, max_height, range_max;
} orc_cz_cfg;
float orc_cz_check_scan(const orc_cz_cfg *cfg, const double *angles, int32_t n_angles,
                        const double *ranges, int32_t forward);
float orc_cz_check_cloud(const orc_cz_cfg *cfg, const double *angles, int32_t n_angles,
                         const int8_t *data, int64_t nbytes, int32_t point_step, int32_t row_step,
                         int32_t height, int32_t width, int32_t x_off, int32_t y_off,
                         int32_t z_off, int32_t forward);
/* preset index lists (for tests): returns count written */
int32_t orc_cz_indices(const orc_cz_cfg *cfg, const double *angles, int32_t n_angles,
                       int32_t forward, int32_t *idx_out);

#ifdef __cplusplus
}
#endif
#endif
