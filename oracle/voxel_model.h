/*
 * voxel_model.h - ORACLE-ONLY (test infrastructure): the occupied-voxel collision model that stands
 * in for FCL 0.7.0 + octomap, shared by the port (kompass_oracle.cpp) and by the FCL/octomap stand-in
 * headers of the _ref build (oracle/shim/fcl, oracle/shim/octomap), so that both arms of a port-vs-_ref
 * comparison evaluate the SAME third-party model: what such a comparison pins is the reference's own
 * sampler control flow around it, never the model itself ("parity unpinned" beyond the reference's
 * three FCL booleans, tests/collisions_test.cpp).
 *
 * ref: include/utils/collision_check.h:91-136 (octomap rebuild per call),
 *      src/utils/collision_check.cpp:38-58 (robot solid), :118-135 (transforms), :149-162 (collide).
 *
 * Third-party restatement (FCL 0.7.0 / octomap, pinned by build_dependencies/install_linux.sh:41,54):
 *  - octomap::OcTree::insertPointCloud(cloud, origin): after a clear(), the occupied leaves are
 *    exactly the unique keys of the end points, key = floor(coord / resolution) per axis
 *    (coordToKeyChecked, |key| < 2^15); free-space ray cells never collide.
 *  - fcl::collide(shape, OcTree): true iff the shape intersects an occupied leaf cube
 *    [k*res,(k+1)*res]^3 placed by the octree object's transform (sensor_tf_world_).
 *  - The boolean is evaluated as an exact closed-set test in double (FCL uses float GJK/MPR with
 *    1e-6 tolerance, so results can differ only within that tolerance of tangency).
 * Two evaluation paths, both exact:
 *  - PLANAR octree frames (the sensor's z axis stays vertical: upright or upside-down mounts): the z
 *    test is folded into the insertion and the xy test is a disc / rectangle against a square;
 *  - GENERAL frames (pitched or rolled sensors): every voxel cube becomes an oriented box in the
 *    robot's frame: sphere vs OBB by clamped projection, box vs OBB by the 15 separating axes,
 *    cylinder vs OBB by clipping the cube with the slab |z| <= h/2, projecting on the xy plane and
 *    measuring the distance from the axis to the convex hull of the projected vertices.
 */
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <unordered_map>
#include <unordered_set>

#include "eigen_order.h"

namespace vox {

enum { VOX_CYLINDER = 0, VOX_BOX = 1, VOX_SPHERE = 2 };

struct Key3 {
  int32_t x, y, z;
  bool operator==(const Key3 &o) const { return x == o.x && y == o.y && z == o.z; }
};
struct Key3Hash {
  size_t operator()(const Key3 &k) const {
    uint64_t h = (uint64_t)(uint32_t)k.x * 0x9E3779B97F4A7C15ull;
    h ^= (uint64_t)(uint32_t)k.y * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (uint64_t)(uint32_t)k.z * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};

struct CollisionWorld {
  int shape;
  double dims[3];
  double res;
  /* octree frame -> world: p_w = A p_s + t (planar) */
  double a00, a01, a10, a11, tx, ty, tz;
  double psi; /* yaw of the octree frame in world */
  double zsign, sigma; /* +-1: z axis kept / flipped; xy block a rotation / a reflection */
  bool planar;
  /* general (tilted sensor) model: every occupied voxel key of the octree frame, p_w = R p_s + t */
  bool orthonormal = true;
  double R[3][3], t[3];
  std::unordered_set<Key3, Key3Hash> voxels;
  int32_t kzmin = INT32_MAX, kzmax = INT32_MIN;
  /* occupied voxel columns after the z test: (kx,ky) -> min over kz of (float)dz^2 (sphere) */
  std::unordered_map<uint64_t, float> columns;
  int32_t kxmin = INT32_MAX, kxmax = INT32_MIN, kymin = INT32_MAX, kymax = INT32_MIN;
  double circ_radius; /* circumscribed xy radius of the footprint */
};

inline uint64_t colKey(int32_t kx, int32_t ky) {
  return ((uint64_t)(uint32_t)kx << 32) | (uint32_t)ky;
}

inline bool keyOf(double res_factor, float coord, int32_t &k) {
  /* octomap coordToKeyChecked: floor(resolution_factor * coordinate), |key| < 32768 */
  const double s = std::floor(res_factor * (double)coord);
  if (!(s >= -32768.0 && s <= 32767.0)) return false; /* also rejects NaN */
  k = (int32_t)s;
  return true;
}

inline void initWorld(CollisionWorld &W, int shape, const float dims[3], double resolution,
                      const orc::Iso3 &sensor_tf_world) {
  W.shape = shape;
  for (int i = 0; i < 3; ++i) W.dims[i] = (double)dims[i];
  W.res = resolution;
  const orc::M3 &L = sensor_tf_world.L;
  W.a00 = L.m[0][0];
  W.a01 = L.m[0][1];
  W.a10 = L.m[1][0];
  W.a11 = L.m[1][1];
  W.tx = sensor_tf_world.t[0];
  W.ty = sensor_tf_world.t[1];
  W.tz = sensor_tf_world.t[2];
  const double tol = 1e-4;
  /* the octree's z axis must stay vertical (upright or upside down: a sensor mounted flipped about
   * x or y keeps its voxel cubes axis-aligned with the upright robot solid); the xy block is then a
   * rotation (det +1) or a reflection (det -1) */
  const double det = W.a00 * W.a11 - W.a01 * W.a10;
  W.zsign = (L.m[2][2] >= 0.0f) ? 1.0 : -1.0;
  W.sigma = (det >= 0.0) ? 1.0 : -1.0;
  W.planar = std::abs(L.m[0][2]) < tol && std::abs(L.m[1][2]) < tol && std::abs(L.m[2][0]) < tol &&
             std::abs(L.m[2][1]) < tol && std::abs(std::abs((double)L.m[2][2]) - 1.0) < tol &&
             std::abs(W.a00 * W.a00 + W.a10 * W.a10 - 1.0) < 1e-3 &&
             std::abs(std::abs(det) - 1.0) < 1e-3;
  /* self-check switch of the oracle's own tests: evaluate planar frames through the general path too */
  if (std::getenv("ORC_FORCE_GENERAL_VOXEL")) W.planar = false;
  W.psi = std::atan2(W.a10, W.a00);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) W.R[i][j] = (double)L.m[i][j];
    W.t[i] = (double)sensor_tf_world.t[i];
  }
  /* general (tilted) mounts need a rotation: unit quaternion within 1e-3 */
  W.orthonormal = true;
  for (int i = 0; i < 3 && W.orthonormal; ++i)
    for (int j = i; j < 3; ++j) {
      const double d = W.R[0][i] * W.R[0][j] + W.R[1][i] * W.R[1][j] + W.R[2][i] * W.R[2][j];
      if (std::abs(d - (i == j ? 1.0 : 0.0)) > 1e-3) W.orthonormal = false;
    }
  if (W.shape == VOX_CYLINDER)
    W.circ_radius = W.dims[0];
  else if (W.shape == VOX_BOX)
    W.circ_radius = 0.5 * std::sqrt(W.dims[0] * W.dims[0] + W.dims[1] * W.dims[1]);
  else
    W.circ_radius = W.dims[0];
}

/* robot centre z in the octree frame: z_w = zsign * z_s + tz = 0  =>  z_s = -zsign * tz */
inline void insertPointPlanar(CollisionWorld &W, float px, float py, float pz) {
  const double rf = 1.0 / W.res;
  int32_t kx, ky, kz;
  if (!keyOf(rf, px, kx) || !keyOf(rf, py, ky) || !keyOf(rf, pz, kz)) return;
  const double lo = (double)kz * W.res, hi = (double)(kz + 1) * W.res;
  const double cz = -W.zsign * W.tz;
  float dz2 = 0.0f;
  if (W.shape == VOX_SPHERE) {
    const double dz = std::max(std::max(lo - cz, 0.0), cz - hi);
    dz2 = (float)(dz * dz);
  } else {
    const double hh = 0.5 * (W.shape == VOX_CYLINDER ? W.dims[1] : W.dims[2]);
    if (!(lo <= cz + hh && hi >= cz - hh)) return; /* closed z-interval overlap */
  }
  auto it = W.columns.find(colKey(kx, ky));
  if (it == W.columns.end())
    W.columns.emplace(colKey(kx, ky), dz2);
  else if (dz2 < it->second)
    it->second = dz2;
  W.kxmin = std::min(W.kxmin, kx);
  W.kxmax = std::max(W.kxmax, kx);
  W.kymin = std::min(W.kymin, ky);
  W.kymax = std::max(W.kymax, ky);
}

/* exact closed test robot-vs-voxel-column, all in double, fixed operation order */
inline bool columnHit(const CollisionWorld &W, int32_t kx, int32_t ky, float dz2f, double cx,
                      double cy, double cth, double sth) {
  const double lox = (double)kx * W.res, hix = (double)(kx + 1) * W.res;
  const double loy = (double)ky * W.res, hiy = (double)(ky + 1) * W.res;
  if (W.shape == VOX_BOX) {
    const double a = 0.5 * W.dims[0], b = 0.5 * W.dims[1];
    const double ex = 0.5 * (hix - lox), ey = 0.5 * (hiy - loy);
    const double dx = 0.5 * (lox + hix) - cx, dy = 0.5 * (loy + hiy) - cy;
    const double ac = std::abs(cth), as = std::abs(sth);
    if (std::abs(dx) > ex + (a * ac + b * as)) return false;
    if (std::abs(dy) > ey + (a * as + b * ac)) return false;
    if (std::abs(dx * cth + dy * sth) > a + (ex * ac + ey * as)) return false;
    if (std::abs(dy * cth - dx * sth) > b + (ex * as + ey * ac)) return false;
    return true;
  }
  const double dx = std::max(std::max(lox - cx, 0.0), cx - hix);
  const double dy = std::max(std::max(loy - cy, 0.0), cy - hiy);
  const double r = W.dims[0];
  double d2 = dx * dx + dy * dy;
  if (W.shape == VOX_SPHERE) d2 = d2 + (double)dz2f;
  return d2 <= r * r;
}

inline bool poseCollidesPlanar(const CollisionWorld &W, double x, double y, double yaw) {
  if (W.columns.empty()) return false;
  /* ref: collision_check.cpp:128-131: pose narrowed to float */
  const double fx = (double)(float)x, fy = (double)(float)y, fyaw = (double)(float)yaw;
  const double dx = fx - W.tx, dy = fy - W.ty;
  const double cx = W.a00 * dx + W.a10 * dy; /* A^T d */
  const double cy = W.a01 * dx + W.a11 * dy;
  double cth = 1.0, sth = 0.0;
  if (W.shape == VOX_BOX) {
    /* heading in the octree frame: A^T u_w = (cos(th), sigma sin(th)), A = R(psi) diag(1, sigma) */
    const double th = fyaw - W.psi;
    cth = std::cos(th);
    sth = W.sigma * std::sin(th);
  }
  const double R = W.circ_radius;
  int32_t kx0 = (int32_t)std::floor((cx - R) / W.res) - 1;
  int32_t kx1 = (int32_t)std::floor((cx + R) / W.res) + 1;
  int32_t ky0 = (int32_t)std::floor((cy - R) / W.res) - 1;
  int32_t ky1 = (int32_t)std::floor((cy + R) / W.res) + 1;
  kx0 = std::max(kx0, W.kxmin);
  kx1 = std::min(kx1, W.kxmax);
  ky0 = std::max(ky0, W.kymin);
  ky1 = std::min(ky1, W.kymax);
  for (int32_t ky = ky0; ky <= ky1; ++ky)
    for (int32_t kx = kx0; kx <= kx1; ++kx) {
      auto it = W.columns.find(colKey(kx, ky));
      if (it == W.columns.end()) continue;
      if (columnHit(W, kx, ky, it->second, cx, cy, cth, sth)) return true;
    }
  return false;
}


/* ---- general (tilted) frames ------------------------------------------------------------------ */
struct Obb {
  double c[3];      /* cube centre in the body frame */
  double ax[3][3];  /* ax[j] = cube axis j in the body frame */
  double e;         /* half side */
};

inline bool sphereHitsObb(double r, const Obb &b) {
  double d2 = 0.0;
  for (int j = 0; j < 3; ++j) {
    const double u = std::abs(b.c[0] * b.ax[j][0] + b.c[1] * b.ax[j][1] + b.c[2] * b.ax[j][2]);
    const double ex = std::max(u - b.e, 0.0);
    d2 = d2 + ex * ex;
  }
  return d2 <= r * r;
}

/* 15 separating axes, strict '>' (closed sets: touching counts as a hit) */
inline bool boxHitsObb(const double a[3], const Obb &b) {
  double Rm[3][3], A[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Rm[i][j] = b.ax[j][i];
      A[i][j] = std::abs(Rm[i][j]);
    }
  const double e = b.e;
  for (int i = 0; i < 3; ++i)
    if (std::abs(b.c[i]) > a[i] + e * (A[i][0] + A[i][1] + A[i][2])) return false;
  for (int j = 0; j < 3; ++j) {
    const double tj = b.c[0] * Rm[0][j] + b.c[1] * Rm[1][j] + b.c[2] * Rm[2][j];
    if (std::abs(tj) > (a[0] * A[0][j] + a[1] * A[1][j] + a[2] * A[2][j]) + e) return false;
  }
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double lhs = std::abs(b.c[i2] * Rm[i1][j] - b.c[i1] * Rm[i2][j]);
      const double ra = a[i1] * A[i2][j] + a[i2] * A[i1][j];
      const double rb = e * A[i][j1] + e * A[i][j2];
      if (lhs > ra + rb) return false;
    }
  }
  return true;
}

/* squared distance from the origin to the convex hull of n 2-D points (n <= 32) */
inline double hullDist2(double (*p)[2], int n) {
  auto seg2 = [](const double *a, const double *b) {
    const double dx = b[0] - a[0], dy = b[1] - a[1];
    const double len2 = dx * dx + dy * dy;
    double t = 0.0;
    if (len2 > 0.0) t = std::min(1.0, std::max(0.0, -(a[0] * dx + a[1] * dy) / len2));
    const double x = a[0] + t * dx, y = a[1] + t * dy;
    return x * x + y * y;
  };
  if (n == 1) return p[0][0] * p[0][0] + p[0][1] * p[0][1];
  int idx[32];
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx, idx + n, [&](int a, int b) { return p[a][0] < p[b][0] || (p[a][0] == p[b][0] && p[a][1] < p[b][1]); });
  auto cross = [&](int o, int a, int b) {
    return (p[a][0] - p[o][0]) * (p[b][1] - p[o][1]) - (p[a][1] - p[o][1]) * (p[b][0] - p[o][0]);
  };
  int h[66], k = 0;
  for (int i = 0; i < n; ++i) {
    while (k >= 2 && cross(h[k - 2], h[k - 1], idx[i]) <= 0.0) --k;
    h[k++] = idx[i];
  }
  for (int i = n - 2, lo = k + 1; i >= 0; --i) {
    while (k >= lo && cross(h[k - 2], h[k - 1], idx[i]) <= 0.0) --k;
    h[k++] = idx[i];
  }
  const int m = k - 1; /* last equals first */
  if (m < 2) return seg2(p[h[0]], p[h[m > 0 ? 1 : 0]]);
  if (m == 2) return seg2(p[h[0]], p[h[1]]);
  bool inside = true;
  double best = INFINITY;
  for (int i = 0; i < m; ++i) {
    const double *a = p[h[i]], *b = p[h[i + 1]];
    /* counter-clockwise hull: the origin is inside iff it is left of (or on) every edge */
    if ((b[0] - a[0]) * (0.0 - a[1]) - (b[1] - a[1]) * (0.0 - a[0]) < 0.0) inside = false;
    best = std::min(best, seg2(a, b));
  }
  return inside ? 0.0 : best;
}

inline bool cylinderHitsObb(double r, double hh, const Obb &b) {
  double v[8][3];
  for (int s = 0; s < 8; ++s)
    for (int d = 0; d < 3; ++d)
      v[s][d] = b.c[d] + b.e * (((s & 1) ? 1.0 : -1.0) * b.ax[0][d] + (((s & 2) ? 1.0 : -1.0) * b.ax[1][d] +
                                                                     ((s & 4) ? 1.0 : -1.0) * b.ax[2][d]));
  bool below = true, above = true;
  for (int s = 0; s < 8; ++s) {
    if (v[s][2] <= hh) above = false;
    if (v[s][2] >= -hh) below = false;
  }
  if (above || below) return false;
  double pts[32][2];
  int n = 0;
  for (int s = 0; s < 8; ++s)
    if (v[s][2] >= -hh && v[s][2] <= hh) {
      pts[n][0] = v[s][0];
      pts[n][1] = v[s][1];
      ++n;
    }
  for (int s = 0; s < 8; ++s)
    for (int bit = 1; bit < 8; bit <<= 1) {
      if (s & bit) continue;
      const int q = s | bit; /* edge s - q */
      for (int side = 0; side < 2; ++side) {
        const double zp = side ? hh : -hh;
        const double da = v[s][2] - zp, db = v[q][2] - zp;
        if ((da < 0.0 && db > 0.0) || (da > 0.0 && db < 0.0)) {
          const double tt = da / (da - db);
          pts[n][0] = v[s][0] + tt * (v[q][0] - v[s][0]);
          pts[n][1] = v[s][1] + tt * (v[q][1] - v[s][1]);
          ++n;
        }
      }
    }
  if (n == 0) return false;
  return hullDist2(pts, n) <= r * r;
}

inline void insertPoint(CollisionWorld &W, float px, float py, float pz) {
  if (W.planar) {
    insertPointPlanar(W, px, py, pz);
    return;
  }
  const double rf = 1.0 / W.res;
  int32_t kx, ky, kz;
  if (!keyOf(rf, px, kx) || !keyOf(rf, py, ky) || !keyOf(rf, pz, kz)) return;
  W.voxels.insert(Key3{kx, ky, kz});
  W.kxmin = std::min(W.kxmin, kx);
  W.kxmax = std::max(W.kxmax, kx);
  W.kymin = std::min(W.kymin, ky);
  W.kymax = std::max(W.kymax, ky);
  W.kzmin = std::min(W.kzmin, kz);
  W.kzmax = std::max(W.kzmax, kz);
}

/* upright robot body at (x, y, 0, yaw) against every occupied voxel cube near it */
inline bool poseCollidesGeneral(const CollisionWorld &W, double x, double y, double yaw) {
  if (W.voxels.empty()) return false;
  const double fx = (double)(float)x, fy = (double)(float)y, fyaw = (double)(float)yaw;
  if (!(std::abs(fx) < 1e9 && std::abs(fy) < 1e9 && std::abs(fyaw) < 1e18)) return false; /* non-finite pose */
  const double cb = std::cos(fyaw), sb = std::sin(fyaw);
  /* body centre in the octree frame: R^T (c_w - t) */
  const double d[3] = {fx - W.t[0], fy - W.t[1], 0.0 - W.t[2]};
  double cs[3];
  for (int j = 0; j < 3; ++j) cs[j] = W.R[0][j] * d[0] + (W.R[1][j] * d[1] + W.R[2][j] * d[2]);
  double rho, hb[3] = {0, 0, 0};
  if (W.shape == VOX_SPHERE) {
    rho = W.dims[0];
  } else if (W.shape == VOX_CYLINDER) {
    rho = std::sqrt(W.dims[0] * W.dims[0] + 0.25 * W.dims[1] * W.dims[1]);
  } else {
    hb[0] = 0.5 * W.dims[0];
    hb[1] = 0.5 * W.dims[1];
    hb[2] = 0.5 * W.dims[2];
    rho = std::sqrt(hb[0] * hb[0] + hb[1] * hb[1] + hb[2] * hb[2]);
  }
  int32_t lo[3], hi[3];
  const int32_t kmin[3] = {W.kxmin, W.kymin, W.kzmin}, kmax[3] = {W.kxmax, W.kymax, W.kzmax};
  for (int j = 0; j < 3; ++j) {
    lo[j] = std::max((int32_t)std::floor((cs[j] - rho) / W.res) - 1, kmin[j]);
    hi[j] = std::min((int32_t)std::floor((cs[j] + rho) / W.res) + 1, kmax[j]);
  }
  /* cube axes in the body frame: Rb^T R, Rb = Rz(yaw) */
  Obb b;
  b.e = 0.5 * W.res;
  for (int j = 0; j < 3; ++j) {
    b.ax[j][0] = cb * W.R[0][j] + sb * W.R[1][j];
    b.ax[j][1] = cb * W.R[1][j] - sb * W.R[0][j];
    b.ax[j][2] = W.R[2][j];
  }
  for (int32_t kz = lo[2]; kz <= hi[2]; ++kz)
    for (int32_t ky = lo[1]; ky <= hi[1]; ++ky)
      for (int32_t kx = lo[0]; kx <= hi[0]; ++kx) {
        if (W.voxels.find(Key3{kx, ky, kz}) == W.voxels.end()) continue;
        /* cube centre: octree frame -> world -> body */
        const double s[3] = {((double)kx + 0.5) * W.res, ((double)ky + 0.5) * W.res, ((double)kz + 0.5) * W.res};
        double w[3];
        for (int i = 0; i < 3; ++i) w[i] = (W.R[i][0] * s[0] + (W.R[i][1] * s[1] + W.R[i][2] * s[2])) + W.t[i];
        const double rx = w[0] - fx, ry = w[1] - fy;
        b.c[0] = cb * rx + sb * ry;
        b.c[1] = cb * ry - sb * rx;
        b.c[2] = w[2];
        bool hit;
        if (W.shape == VOX_SPHERE)
          hit = sphereHitsObb(W.dims[0], b);
        else if (W.shape == VOX_CYLINDER)
          hit = cylinderHitsObb(W.dims[0], 0.5 * W.dims[1], b);
        else
          hit = boxHitsObb(hb, b);
        if (hit) return true;
      }
  return false;
}

inline bool poseCollides(const CollisionWorld &W, double x, double y, double yaw) {
  return W.planar ? poseCollidesPlanar(W, x, y, yaw) : poseCollidesGeneral(W, x, y, yaw);
}

inline bool worldSupported(const CollisionWorld &W) { return W.planar || W.orthonormal; }

}  // namespace vox
