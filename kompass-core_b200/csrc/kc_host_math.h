// kc_host_math.h — host-side float rigid-transform algebra evaluated in the same operation order
// as the Eigen 3.4 expressions the reference uses (per-cycle scalars only; never per point).
//
// ref call sites: include/utils/transformation.h:10-42 (eulerToRotationMatrix, getTransformation),
// include/utils/cost_evaluator.h:180-189 (sensor_tf_body_ * body_tf_world_),
// include/utils/collision_check.h:101 (body->tf * sensor_tf_body_),
// src/utils/critical_zone_check.cpp:43-44.
// Eigen semantics reproduced: Quaternionf(Vector4f) takes coeffs (x,y,z,w) un-normalised;
// toRotationMatrix() is the 12-product form; Quaternionf(Matrix3f) is the trace method;
// Isometry*Isometry = {L1*L2, L1*t2 + t1}; 3-term sums associate as a + (b + c).
#pragma once
#include <cmath>

namespace kc {
namespace hm {

struct Rot {
  float r[9];  // row-major 3x3
  float &operator()(int i, int j) { return r[3 * i + j]; }
  float operator()(int i, int j) const { return r[3 * i + j]; }
};

struct Rigid {
  Rot R;
  float t[3];
};

inline float add3(float a, float b, float c) { return a + (b + c); }

inline Rot rot_from_quat(float qx, float qy, float qz, float qw) {
  const float x2 = 2.0f * qx, y2 = 2.0f * qy, z2 = 2.0f * qz;
  const float wx = x2 * qw, wy = y2 * qw, wz = z2 * qw;
  const float xx = x2 * qx, xy = y2 * qx, xz = z2 * qx;
  const float yy = y2 * qy, yz = z2 * qy, zz = z2 * qz;
  Rot m;
  m(0, 0) = 1.0f - (yy + zz);
  m(0, 1) = xy - wz;
  m(0, 2) = xz + wy;
  m(1, 0) = xy + wz;
  m(1, 1) = 1.0f - (xx + zz);
  m(1, 2) = yz - wx;
  m(2, 0) = xz - wy;
  m(2, 1) = yz + wx;
  m(2, 2) = 1.0f - (xx + yy);
  return m;
}

inline void quat_from_rot(const Rot &m, float q[4] /* x,y,z,w */) {
  float tr = add3(m(0, 0), m(1, 1), m(2, 2));
  if (tr > 0.0f) {
    float s = std::sqrt(tr + 1.0f);
    q[3] = 0.5f * s;
    s = 0.5f / s;
    q[0] = (m(2, 1) - m(1, 2)) * s;
    q[1] = (m(0, 2) - m(2, 0)) * s;
    q[2] = (m(1, 0) - m(0, 1)) * s;
    return;
  }
  int i = 0;
  if (m(1, 1) > m(0, 0)) i = 1;
  if (m(2, 2) > m(i, i)) i = 2;
  const int j = (i + 1) % 3, k = (j + 1) % 3;
  float s = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0f);
  q[i] = 0.5f * s;
  s = 0.5f / s;
  q[3] = (m(k, j) - m(j, k)) * s;
  q[j] = (m(j, i) + m(i, j)) * s;
  q[k] = (m(k, i) + m(i, k)) * s;
}

// AngleAxisf(yaw, Z) * AngleAxisf(0, Y) * AngleAxisf(0, X) -> matrix
inline Rot rot_from_yaw(float yaw) {
  const float half = 0.5f * yaw;
  return rot_from_quat(0.0f, 0.0f, std::sin(half), std::cos(half));
}

inline Rigid rigid_from_quat(const float q[4], const float t[3]) {
  Rigid T;
  T.R = rot_from_quat(q[0], q[1], q[2], q[3]);
  T.t[0] = t[0];
  T.t[1] = t[1];
  T.t[2] = t[2];
  return T;
}

// getTransformation(Matrix3f, t): the matrix goes through Quaternionf(Matrix3f) and back
inline Rigid rigid_from_rot(const Rot &R, const float t[3]) {
  float q[4];
  quat_from_rot(R, q);
  return rigid_from_quat(q, t);
}

// getTransformation(Path::State)
inline Rigid rigid_from_state(double x, double y, double yaw) {
  const float t[3] = {static_cast<float>(x), static_cast<float>(y), 0.0f};
  return rigid_from_rot(rot_from_yaw(static_cast<float>(yaw)), t);
}

inline void rot_apply(const Rot &R, const float v[3], float out[3]) {
  for (int i = 0; i < 3; ++i) out[i] = add3(R(i, 0) * v[0], R(i, 1) * v[1], R(i, 2) * v[2]);
}

inline Rigid compose(const Rigid &A, const Rigid &B) {
  Rigid C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.R(i, j) = add3(A.R(i, 0) * B.R(0, j), A.R(i, 1) * B.R(1, j), A.R(i, 2) * B.R(2, j));
  float Rt[3];
  rot_apply(A.R, B.t, Rt);
  for (int i = 0; i < 3; ++i) C.t[i] = Rt[i] + A.t[i];
  return C;
}

}  // namespace hm
}  // namespace kc
