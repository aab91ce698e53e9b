/*
 * eigen_order.h — ORACLE-ONLY restatement of the handful of Eigen 3.4 float operations the
 * reference hot path relies on (Eigen is a third-party dependency that is NOT vendored in
 * /root/reference: `find_package(Eigen3 3.4 REQUIRED)`, src/kompass_cpp/CMakeLists.txt:2).
 *
 * What is restated (published Eigen 3.4 algorithms, scalar/unvectorised evaluation order):
 *   - QuaternionBase::toRotationMatrix()              (Eigen/src/Geometry/Quaternion.h)
 *   - Quaternion(Matrix3) a.k.a. quaternionbase_assign_impl<Other,3,3> (trace method)
 *   - AngleAxis -> Quaternion (half-angle) and Quaternion * Quaternion
 *   - Transform<float,3,Isometry>: translate(), rotate(), T*T, T*Vector3f
 *   - 3-term reductions evaluate as a0 + (a1 + a2) (redux_novec_unroller, Length 3)
 * Call sites in the reference that these stand in for:
 *   ref: include/utils/transformation.h:10-71, include/utils/cost_evaluator.h:180-189,
 *        include/utils/collision_check.h:101-125, src/utils/critical_zone_check.cpp:43-44,72,103
 * Where the 3-term order could matter (non-planar sensor mounts) parity with a real Eigen
 * build is unpinned to 1 ulp; for planar transforms (z = 0 inputs, yaw-only rotations) every
 * candidate order gives identical bits because the extra terms are exact zeros.
 */
#pragma once
#include <cmath>

namespace orc {

struct M3 {
  float m[3][3];
};
struct Quat {
  float x, y, z, w;
};
struct Iso3 {
  M3 L;
  float t[3];
};

inline float sum3(float a0, float a1, float a2) { return a0 + (a1 + a2); }

inline M3 identity3() {
  M3 r{};
  r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0f;
  return r;
}

inline M3 quatToMatrix(const Quat &q) {
  const float tx = 2.0f * q.x, ty = 2.0f * q.y, tz = 2.0f * q.z;
  const float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 r;
  r.m[0][0] = 1.0f - (tyy + tzz);
  r.m[0][1] = txy - twz;
  r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz;
  r.m[1][1] = 1.0f - (txx + tzz);
  r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy;
  r.m[2][1] = tyz + twx;
  r.m[2][2] = 1.0f - (txx + tyy);
  return r;
}

inline Quat matrixToQuat(const M3 &a) {
  Quat q;
  float t = sum3(a.m[0][0], a.m[1][1], a.m[2][2]);
  if (t > 0.0f) {
    t = std::sqrt(t + 1.0f);
    q.w = 0.5f * t;
    t = 0.5f / t;
    q.x = (a.m[2][1] - a.m[1][2]) * t;
    q.y = (a.m[0][2] - a.m[2][0]) * t;
    q.z = (a.m[1][0] - a.m[0][1]) * t;
  } else {
    int i = 0;
    if (a.m[1][1] > a.m[0][0]) i = 1;
    if (a.m[2][2] > a.m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(a.m[i][i] - a.m[j][j] - a.m[k][k] + 1.0f);
    float c[3];
    c[i] = 0.5f * t;
    t = 0.5f / t;
    q.w = (a.m[k][j] - a.m[j][k]) * t;
    c[j] = (a.m[j][i] + a.m[i][j]) * t;
    c[k] = (a.m[k][i] + a.m[i][k]) * t;
    q.x = c[0];
    q.y = c[1];
    q.z = c[2];
  }
  return q;
}

/* ref: include/utils/transformation.h:10-18 eulerToRotationMatrix(0,0,yaw):
 * (AngleAxis(yaw,Z) * AngleAxis(0,Y) * AngleAxis(0,X)).matrix(); the two identity quaternion
 * products are exact, leaving q = (0,0,sin(yaw/2),cos(yaw/2)). */
inline M3 yawToMatrix(float yaw) {
  const float ha = 0.5f * yaw;
  Quat q{0.0f, 0.0f, std::sin(ha), std::cos(ha)};
  return quatToMatrix(q);
}

/* ref: transformation.h:20-34 getTransformation(rotation, translation):
 * Identity.translate(t).rotate(Quaternionf(rotation)) */
inline Iso3 makeTransform(const Quat &q, const float t[3]) {
  Iso3 r;
  r.L = quatToMatrix(q);
  r.t[0] = t[0];
  r.t[1] = t[1];
  r.t[2] = t[2];
  return r;
}
inline Iso3 makeTransform(const M3 &R, const float t[3]) {
  return makeTransform(matrixToQuat(R), t);
}
/* ref: transformation.h:36-42 getTransformation(Path::State) */
inline Iso3 stateToTransform(double x, double y, double yaw) {
  const float t[3] = {float(x), float(y), 0.0f};
  return makeTransform(yawToMatrix(float(yaw)), t);
}

inline void mulVec(const M3 &A, const float p[3], float out[3]) {
  for (int i = 0; i < 3; ++i)
    out[i] = sum3(A.m[i][0] * p[0], A.m[i][1] * p[1], A.m[i][2] * p[2]);
}

/* Transform * Transform (isometry mode): linear = L1*L2 ; translation = L1*t2 + t1 */
inline Iso3 mul(const Iso3 &a, const Iso3 &b) {
  Iso3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      r.L.m[i][j] =
          sum3(a.L.m[i][0] * b.L.m[0][j], a.L.m[i][1] * b.L.m[1][j], a.L.m[i][2] * b.L.m[2][j]);
  float lt[3];
  mulVec(a.L, b.t, lt);
  for (int i = 0; i < 3; ++i) r.t[i] = lt[i] + a.t[i];
  return r;
}

/* Transform * Vector3f : translation + linear * v */
inline void apply(const Iso3 &T, const float p[3], float out[3]) {
  float lp[3];
  mulVec(T.L, p, lp);
  for (int i = 0; i < 3; ++i) out[i] = T.t[i] + lp[i];
}

} // namespace orc
