// kc_mapper.cu — local mapper (scan -> occupancy grid), point-cloud -> laser-scan binning and the
// critical-zone checker. All three are tiny, latency-bound passes: one fill + one or two kernels +
// one D2H per call, on the handle's own stream.
//
// Parity target is the reference CPU path (NOT the reference SYCL kernels, which use a different
// DDA / per-point cone algorithm):
//   ref: src/mapping/local_mapper.cpp:127-159,204-251; include/mapping/local_mapper.h:26-27,210-222;
//        include/mapping/line_drawing.h:55-124 (super-cover Bresenham);
//        include/utils/pointcloud.h:205-259; src/utils/critical_zone_check.cpp:13-131.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <vector>

#include "kc_common.cuh"
#include "kc_host_math.h"
#include "kc_hostcopy.h"
#include "kc_libm_compat.cuh"

using namespace kc;

namespace {

__device__ __forceinline__ double bin_range(unsigned int bits, double max_range) {
  if (bits == 0xffffffffu) return max_range;
  const double d = (double)__uint_as_float(bits);
  return d < max_range ? d : max_range;  // ranges_out starts at max_range; strict '<' update
}

__global__ void k_bins_to_ranges(const unsigned int *__restrict__ bins, int n, double max_range,
                                 double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = bin_range(bins[i], max_range);
}

// ------------------------------------------------------------------------------------------------
// scan -> grid: one WARP per ray; the i-th step of the super-cover line is computed in closed form
// (below), so the 32 lanes stamp 32 steps of the ray at once instead of one thread walking it.
// Final cell value is order independent: 100 if any ray ends in it, else 0 if any ray crosses it,
// else -1 (local_mapper.cpp:143-155) -> atomicMax reproduces the serial result exactly.
// ------------------------------------------------------------------------------------------------
struct MapParams {
  int H, W;
  float res;
  float px, py, orient;
  int c0, c1;  // central cell
  int s0, s1;  // start cell of every ray
};

__device__ __forceinline__ void map_visit(int *grid, const MapParams &mp, int px, int py, int t0,
                                          int t1) {
  if (px >= 0 && px < mp.H && py >= 0 && py < mp.W) {
    const int v = (px == t0 && py == t1) ? KC_OCCUPIED : KC_EMPTY;
    int *cell = &grid[(size_t)px + (size_t)py * mp.H];
    // cells only grow (-1 < 0 < 100): a cell that already holds >= v needs no atomic. The cells next
    // to the sensor are crossed by every ray; this keeps thousands of same-address atomics off L2.
    if (__ldcg(cell) < v) atomicMax(cell, v);
  }
}

// Closed form of the super-cover walk (ref: line_drawing.h:55-124 bresenhamEnhanced). In the serial
// loop of the major-axis case (ddx >= ddy; the other case swaps the roles of x and y)
//     error += ddy; if (error > ddx) { py += ystep; error -= ddx; ... }
// `error` starts at dx and stays in (0, ddx], so after step i (1-based)
//     E_i = dx + i * ddy,   k_i = (E_i - 1) / ddx  minor-axis steps so far,   error_i = E_i - k_i * ddx
// and step i moved the minor axis iff k_i > k_(i-1); the corner cells it adds depend on
// error_i + error_(i-1) against ddx exactly as in the loop. Every step is therefore independent of
// the others: lane l of the ray's warp stamps steps l+1, l+33, ...
struct LineSetup {
  int s0, s1, t0, t1;
  int xstep, ystep, dx, dy;  // dx, dy absolute
  int n_steps;               // steps of the major axis that can still touch the grid
  bool x_major;
  bool small;                // dmaj + n_steps * 2 dmin < 2^30: 32-bit arithmetic in line_step
};

__device__ __forceinline__ LineSetup line_setup(const MapParams &mp, int t0, int t1) {
  LineSetup L;
  L.s0 = mp.s0;
  L.s1 = mp.s1;
  L.t0 = t0;
  L.t1 = t1;
  const int dx = t0 - mp.s0, dy = t1 - mp.s1;
  L.xstep = (dx >= 0) ? 1 : -1;
  L.ystep = (dy >= 0) ? 1 : -1;
  L.dx = abs(dx);
  L.dy = abs(dy);
  L.x_major = (2 * (long long)L.dx >= 2 * (long long)L.dy);
  // once the walk is outside the grid on the side it moves towards it never re-enters (the serial
  // loop breaks there): steps beyond the far border of the major axis cannot stamp anything
  int room;
  if (L.x_major)
    room = (L.xstep > 0) ? (mp.H - mp.s0) : (mp.s0 + 1);
  else
    room = (L.ystep > 0) ? (mp.W - mp.s1) : (mp.s1 + 1);
  const int major = L.x_major ? L.dx : L.dy;
  L.n_steps = max(0, min(major, room + 1));
  const long long minor = L.x_major ? L.dy : L.dx;
  L.small = (long long)major * 2 < (1LL << 30) && (long long)major + (long long)L.n_steps * 2 * minor < (1LL << 30);
  return L;
}

// cells of step i (1 <= i <= n_steps): visit(px, py) up to three times, in the serial loop's order.
// T = unsigned (L.small: every E_i of the ray fits 31 bits, the case of any ray near the grid) or
// long long (rays whose end cell lies millions of cells away)
template <typename T, class Visit>
__device__ __forceinline__ void line_step(const LineSetup &L, int i, Visit visit) {
  const T dmaj = (T)(L.x_major ? L.dx : L.dy), dmin = (T)(L.x_major ? L.dy : L.dx);
  const T ddmaj = 2 * dmaj, ddmin = 2 * dmin;
  const T e1 = dmaj + (T)i * ddmin;  // E_i
  const T e0 = e1 - ddmin;           // E_(i-1)  (>= dmaj >= 1)
  const T k1 = (e1 - 1) / ddmaj, k0 = (e0 - 1) / ddmaj;
  const T err1 = e1 - k1 * ddmaj, err0 = e0 - k0 * ddmaj;
  int px, py;
  if (L.x_major) {
    px = L.s0 + i * L.xstep;
    py = L.s1 + (int)k1 * L.ystep;
    if (k1 > k0) {
      const T sum = err1 + err0;
      if (sum < ddmaj) {
        visit(px, py - L.ystep);
      } else if (sum > ddmaj) {
        visit(px - L.xstep, py);
      } else {
        visit(px - L.xstep, py);
        visit(px, py - L.ystep);
      }
    }
  } else {
    py = L.s1 + i * L.ystep;
    px = L.s0 + (int)k1 * L.xstep;
    if (k1 > k0) {
      const T sum = err1 + err0;
      if (sum < ddmaj) {
        visit(px - L.xstep, py);
      } else if (sum > ddmaj) {
        visit(px, py - L.ystep);
      } else {
        visit(px - L.xstep, py);
        visit(px, py - L.ystep);
      }
    }
  }
  visit(px, py);
}

// ray end cell (ref local_mapper.cpp:129-134): float + (float * double cos(float sum)) narrowed to
// float, then localToGrid (truncation toward zero)
__device__ __forceinline__ void ray_end_cell(const MapParams &mp, float angle, float range, int &t0,
                                             int &t1) {
  double s, c;
  sincos((double)(mp.orient + angle), &s, &c);
  const float x = (float)((double)mp.px + ((double)range * c));
  const float y = (float)((double)mp.py + ((double)range * s));
  t0 = mp.c0 + (int)(x / mp.res);
  t1 = mp.c1 + (int)(y / mp.res);
}

// angles: double[n]; ranges: double[n] (RANGES_FROM_BINS: uint bins converted on the fly)
template <bool RANGES_FROM_BINS>
__global__ void k_scan_to_grid(MapParams mp, const double *__restrict__ angles,
                               const double *__restrict__ ranges,
                               const unsigned int *__restrict__ bins, double max_range, int n,
                               int *__restrict__ grid) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    const float angle = (float)angles[r];
    const float range = (float)(RANGES_FROM_BINS ? bin_range(bins[r], max_range) : ranges[r]);
    int t0, t1;
    ray_end_cell(mp, angle, range, t0, t1);  // every lane computes the same end cell (no shuffle needed)
    const LineSetup L = line_setup(mp, t0, t1);
    if (lane == 0) map_visit(grid, mp, L.s0, L.s1, t0, t1);
    auto visit = [&](int px, int py) { map_visit(grid, mp, px, py, t0, t1); };
    if (L.small)
      for (int i = lane + 1; i <= L.n_steps; i += 32) line_step<unsigned>(L, i, visit);
    else
      for (int i = lane + 1; i <= L.n_steps; i += 32) line_step<long long>(L, i, visit);
  }
}

// ------------------------------------------------------------------------------------------------
// Bayesian update (ref local_mapper.cpp:106-125,161-202,222-238). The reference
// walks the rays serially and every ray OVERWRITES the probability of the cells it crosses, so a
// cell ends with the value computed by the last (highest-index) ray that touched it. In parallel:
// 64-bit atomicMax of (ray index + 1) << 32 | float bits; a second pass unpacks (untouched cells
// keep the prior).
// ------------------------------------------------------------------------------------------------
struct BayesParams {
  float p_prior, p_occupied, p_empty, range_sure, range_max, wall_size;
};

// ref local_mapper.cpp:106-125 (operand widths as written: float products, one double chain)
__device__ __forceinline__ float bayes_cell_probability(const BayesParams &bp, float res, float distance,
                                                        float current_range, float previous) {
  distance = distance * res;
  current_range = current_range - bp.wall_size;
  const float pF = (distance < current_range) ? bp.p_empty : bp.p_occupied;
  const float delta = (distance < bp.range_sure) ? 0.0f : 1.0f;
  const float pSensor = pF + (delta * ((distance - bp.range_sure) / bp.range_max) * (bp.p_prior - pF));
  const float a = previous / (1 - previous);
  const double b = (double)pSensor / (1.0 - (double)pSensor);
  const float c = (1 - bp.p_prior) / bp.p_prior;
  const double pCurr = 1 - (1 / (1 + (((double)a * b) * (double)c)));
  return (float)pCurr;
}

// RANGES_FROM_BINS: the point-cloud overload (local_mapper.cpp:253-264): ray r has angle
// r * angle_step (double product, pointcloud.h:131) and the binned range
template <bool RANGES_FROM_BINS>
__global__ void k_scan_to_grid_bayes(MapParams mp, BayesParams bp, const double *__restrict__ angles,
                                     const double *__restrict__ ranges,
                                     const unsigned int *__restrict__ bins, double max_range,
                                     double angle_step, int n, const float *__restrict__ prev,
                                     int *__restrict__ grid, unsigned long long *__restrict__ keys) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
    const float angle = (float)(RANGES_FROM_BINS ? (double)r * angle_step : angles[r]);
    const float range = (float)(RANGES_FROM_BINS ? bin_range(bins[r], max_range) : ranges[r]);
    int t0, t1;
    ray_end_cell(mp, angle, range, t0, t1);
    const LineSetup L = line_setup(mp, t0, t1);
    auto visit = [&](int px, int py) {
      if (px >= 0 && px < mp.H && py >= 0 && py < mp.W) {
        const size_t idx = (size_t)px + (size_t)py * mp.H;
        const int v = (px == t0 && py == t1) ? KC_OCCUPIED : KC_EMPTY;
        if (__ldcg(&grid[idx]) < v) atomicMax(&grid[idx], v);
        // Vector2i::norm(): Eigen's integer sqrt_impl truncates (int)sqrt(dx^2 + dy^2)
        const int ddx = px - mp.s0, ddy = py - mp.s1;
        const float distance = (float)(int)sqrt((double)(ddx * ddx + ddy * ddy));
        const float pv = bayes_cell_probability(bp, mp.res, distance, range, prev[idx]);
        const unsigned long long key = ((unsigned long long)(unsigned)(r + 1) << 32) | __float_as_uint(pv);
        // the LAST ray crossing a cell decides it: a cell already claimed by a later ray needs no atomic
        if (__ldcg(&keys[idx]) < key) atomicMax(&keys[idx], key);
      }
    };
    if (lane == 0) visit(L.s0, L.s1);
    if (L.small)
      for (int i = lane + 1; i <= L.n_steps; i += 32) line_step<unsigned>(L, i, visit);
    else
      for (int i = lane + 1; i <= L.n_steps; i += 32) line_step<long long>(L, i, visit);
  }
}

__global__ void k_bayes_finalize(const unsigned long long *__restrict__ keys, float prior, size_t cells,
                                 float *__restrict__ prob) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cells;
       i += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    prob[i] = k ? __uint_as_float((unsigned)(k & 0xffffffffull)) : prior;
  }
}

// previous-grid warp (ref local_mapper.cpp:17-78): inverse transform computed once on the host
// (Eigen's cofactor inverse, float), bilinear sample per cell
struct WarpParams {
  int H, W;
  float inv[9];
  float prior;
};
__global__ void k_warp_previous(WarpParams wp, const float *__restrict__ prev, float *__restrict__ out) {
  const size_t cells = (size_t)wp.H * wp.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < cells;
       i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i % wp.H), x = (int)(i / wp.H);  // out(y, x), column-major
    const float sx = (float)x, sy = (float)y;
    const float d0 = wp.inv[0] * sx + (wp.inv[1] * sy + wp.inv[2] * 1.0f);
    const float d1 = wp.inv[3] * sx + (wp.inv[4] * sy + wp.inv[5] * 1.0f);
    const double srcX = d0, srcY = d1;
    float value = wp.prior;
    if (srcX >= 0 && srcX < wp.W - 1 && srcY >= 0 && srcY < wp.H - 1) {
      const int x0 = (int)floor(srcX), y0 = (int)floor(srcY);
      const int x1 = x0 + 1, y1 = y0 + 1;
      const float w0 = (float)(srcX - x0), w1 = 1.0f - w0, h0 = (float)(srcY - y0), h1 = 1.0f - h0;
      const float p00 = prev[(size_t)y0 + (size_t)x0 * wp.H], p01 = prev[(size_t)y0 + (size_t)x1 * wp.H];
      const float p10 = prev[(size_t)y1 + (size_t)x0 * wp.H], p11 = prev[(size_t)y1 + (size_t)x1 * wp.H];
      value = h1 * (w1 * p00 + w0 * p01) + h0 * (w1 * p10 + w0 * p11);
    }
    out[i] = value;
  }
}

// ------------------------------------------------------------------------------------------------
// critical zone: indexed rays -> slowdown factor. The serial reference returns 0 at the first
// critical ray and otherwise the min factor; both are order independent (factor > 0 off-critical),
// so a min-reduction with "critical -> 0" is exact. Factors are in [0,1]: uint-ordered bits.
// ------------------------------------------------------------------------------------------------
struct CzParams {
  float T[12];
  double robot_radius;
  float critical_distance, slowdown_distance;
};

// ------------------------------------------------------------------------------------------------
// One-launch forms of the critical-zone check (latency path). The result is a 16-byte record in
// page-locked host memory that the kernel writes itself - factor first, sequence number last behind a
// system-scope fence - and the host watches, instead of a D2H copy plus a stream synchronisation.
//   scan:  one CTA reads the ranges straight from the handle's page-locked staging buffer over PCIe
//          (28.8 KB at 3600 rays), reduces and publishes: memcpy + ONE launch per call.
//   cloud: the binning kernel's last CTA (ticket counter) runs the zone reduction over the finished
//          bins and publishes: memset + ONE launch per call.
// ------------------------------------------------------------------------------------------------
struct CzRecord {
  float factor;
  uint32_t pad0;
  volatile uint32_t seq;
  uint32_t pad1;
};

struct CzTail {  // zone reduction appended to the binning kernel (enabled != 0)
  int enabled;
  CzParams cp;
  const int *idx;
  int n_idx;
  const float *cos_a, *sin_a;
  double max_range;
  unsigned int *ticket;  // self-resetting CTA counter
  CzRecord *record;      // mapped page-locked host memory
  uint32_t seq;
};

__device__ __forceinline__ float cz_ray_factor(const CzParams &cp, double r, float ca, float sa, float f) {
  const float x = (float)(r * (double)ca);
  const float y = (float)(r * (double)sa);
  const float *T = cp.T;
  const float qx = T[9] + (T[0] * x + (T[1] * y + T[2] * 0.0f));
  const float qy = T[10] + (T[3] * x + (T[4] * y + T[5] * 0.0f));
  const float conv = (float)sqrt((double)qy * (double)qy + (double)qx * (double)qx);
  const float dist = (float)((double)conv - cp.robot_radius);
  if (dist <= cp.critical_distance) return 0.0f;
  if (dist <= cp.slowdown_distance)
    return fminf(f, (dist - cp.critical_distance) / (cp.slowdown_distance - cp.critical_distance));
  return f;
}

// block-wide min of f (blockDim.x a multiple of 32, <= 1024); valid in thread 0
__device__ __forceinline__ float block_min(float f, float *s_red) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) f = fminf(f, __shfl_xor_sync(FULL, f, m));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = f;
  __syncthreads();
  if (threadIdx.x < 32) {
    f = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 1.0f;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) f = fminf(f, __shfl_xor_sync(FULL, f, m));
  }
  return f;
}

__device__ __forceinline__ void cz_publish(CzRecord *rec, float f, uint32_t seq) {
  rec->factor = f;
  __threadfence_system();
  rec->seq = seq;
}

__global__ void __launch_bounds__(1024) k_critical_zone_scan1(CzParams cp, const int *__restrict__ idx, int n_idx,
                                                              const float *__restrict__ cos_a,
                                                              const float *__restrict__ sin_a,
                                                              const double *ranges_host, int n_stage,
                                                              CzRecord *rec, uint32_t seq) {
  __shared__ float s_red[32];
  extern __shared__ double s_ranges[];  // n_stage doubles (0: the ranges are read in place)
  // every PCIe round trip costs more than the whole reduction: pull the scan into shared memory with
  // 16-byte loads, all in flight at once (28.8 KB at 3600 rays = two rounds of the CTA)
  const double *src = ranges_host;
  if (n_stage > 0) {
    const double2 *v = reinterpret_cast<const double2 *>(ranges_host);
    double2 *d = reinterpret_cast<double2 *>(s_ranges);
    for (int k = threadIdx.x; k < n_stage / 2; k += blockDim.x) d[k] = v[k];
    if (threadIdx.x == 0 && (n_stage & 1)) s_ranges[n_stage - 1] = ranges_host[n_stage - 1];
    __syncthreads();
    src = s_ranges;
  }
  float f = 1.0f;
  for (int k = threadIdx.x; k < n_idx; k += blockDim.x) {
    const int i = idx[k];
    f = cz_ray_factor(cp, src[i], cos_a[i], sin_a[i], f);
  }
  f = block_min(f, s_red);
  if (threadIdx.x == 0) cz_publish(rec, f, seq);
}

// called by every thread of every CTA at the end of a binning kernel
__device__ __forceinline__ void cz_tail(const CzTail &t, const unsigned int *bins) {
  if (!t.enabled) return;
  __shared__ float s_red[32];
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(t.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float f = 1.0f;
  for (int k = threadIdx.x; k < t.n_idx; k += blockDim.x) {
    const int i = t.idx[k];
    f = cz_ray_factor(t.cp, bin_range(__ldcg(&bins[i]), t.max_range), t.cos_a[i], t.sin_a[i], f);
  }
  f = block_min(f, s_red);
  if (threadIdx.x == 0) {
    *t.ticket = 0u;  // ready for the next call
    cz_publish(t.record, f, t.seq);
  }
}

// ------------------------------------------------------------------------------------------------
// point cloud -> laser scan: per point filter, float atan2 (glibc-compatible), bin, atomic min.
// Range candidates are non-negative floats, so their bit patterns order like unsigned ints.
// ------------------------------------------------------------------------------------------------
__global__ void k_cloud_to_bins(const int8_t *__restrict__ data, long long nbytes, int point_step,
                                int row_step, int height, int x_off, int y_off, int z_off,
                                double min_z, double max_z, int num_bins, double angle_step,
                                unsigned int *__restrict__ bins, CzTail tail) {
  const int per_row = (row_step + point_step - 1) / point_step;
  const long long total = (long long)height * per_row;
  const double two_pi = 2.0 * M_PI;
  const int max_off = max(x_off, max(y_off, z_off));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / per_row), col = (int)(i % per_row) * point_step;
    const size_t point_start = (size_t)(row * row_step + col);
    if (point_start + (size_t)max_off + sizeof(float) > (size_t)nbytes) continue;
    float x, y, z;  // byte-wise loads: offsets need not be 4-aligned
    {
      const unsigned char *b = reinterpret_cast<const unsigned char *>(data) + point_start;
      unsigned int ux = b[x_off] | (b[x_off + 1] << 8) | (b[x_off + 2] << 16) | ((unsigned)b[x_off + 3] << 24);
      unsigned int uy = b[y_off] | (b[y_off + 1] << 8) | (b[y_off + 2] << 16) | ((unsigned)b[y_off + 3] << 24);
      unsigned int uz = b[z_off] | (b[z_off + 1] << 8) | (b[z_off + 2] << 16) | ((unsigned)b[z_off + 3] << 24);
      x = __uint_as_float(ux);
      y = __uint_as_float(uy);
      z = __uint_as_float(uz);
    }
    const float range_sq = x * x + y * y;
    if ((double)range_sq < 1e-6) continue;
    if ((double)z < min_z || (max_z >= 0.0 && (double)z > max_z)) continue;
    double angle = (double)compat_atan2f(y, x);
    if (angle < 0.0) angle += two_pi;
    if (!(angle == angle)) continue;  // NaN coordinates: int(NaN) is undefined in the reference
    // num_bins overload (pointcloud.h:249) or angle_step overload (pointcloud.h:167)
    int bin = angle_step > 0.0 ? (int)(angle / angle_step) : (int)((angle / two_pi) * num_bins);
    bin = min(bin, num_bins - 1);
    const float dist = sqrtf(range_sq);
    if (!(dist == dist)) continue;
    atomicMin(&bins[bin], __float_as_uint(dist));
  }
  cz_tail(tail, bins);
}

// 4-aligned fast path (the common PointCloud2 layout): one 16-byte vector load per point
__global__ void k_cloud_to_bins_xyz16(const float4 *__restrict__ pts, int n, double min_z,
                                      double max_z, int num_bins, double angle_step,
                                      unsigned int *__restrict__ bins, CzTail tail) {
  const double two_pi = 2.0 * M_PI;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(&pts[i]);
    const float range_sq = p.x * p.x + p.y * p.y;
    if ((double)range_sq < 1e-6) continue;
    if ((double)p.z < min_z || (max_z >= 0.0 && (double)p.z > max_z)) continue;
    double angle = (double)compat_atan2f(p.y, p.x);
    if (angle < 0.0) angle += two_pi;
    if (!(angle == angle)) continue;
    // num_bins overload (pointcloud.h:249) or angle_step overload (pointcloud.h:167)
    int bin = angle_step > 0.0 ? (int)(angle / angle_step) : (int)((angle / two_pi) * num_bins);
    bin = min(bin, num_bins - 1);
    const float dist = sqrtf(range_sq);
    if (!(dist == dist)) continue;
    atomicMin(&bins[bin], __float_as_uint(dist));
  }
  cz_tail(tail, bins);
}

int32_t launch_binning(cudaStream_t st, const int8_t *d_data, int64_t nbytes, int point_step,
                       int row_step, int height, int x_off, int y_off, int z_off, double min_z,
                       double max_z, int num_bins, unsigned int *d_bins, double angle_step = 0.0,
                       const CzTail *tail_in = nullptr) {
  CzTail tail{};
  if (tail_in) tail = *tail_in;
  KC_CUDA(cudaMemsetAsync(d_bins, 0xFF, (size_t)num_bins * 4, st));
  if (point_step <= 0 || height <= 0 || row_step <= 0 || nbytes <= 0) {
    if (tail.enabled) {  // no point at all: the zone reduction still has to run (every bin = max_range)
      k_cloud_to_bins<<<1, 256, 0, st>>>(d_data, 0, 1, 1, 0, 0, 0, 0, min_z, max_z, num_bins, angle_step, d_bins, tail);
      KC_CUDA(cudaGetLastError());
    }
    return KC_OK;
  }
  const int per_row = (row_step + point_step - 1) / point_step;
  const long long total = (long long)height * per_row;
  const bool fast = point_step == 16 && x_off == 0 && y_off == 4 && z_off == 8 &&
                    (row_step % 16 == 0) && ((int64_t)height * row_step <= nbytes) &&
                    ((uintptr_t)d_data % 16 == 0);
  const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, 8LL * sm_count()));
  if (fast) {
    const int n = (int)((int64_t)height * row_step / 16);
    k_cloud_to_bins_xyz16<<<grid, 256, 0, st>>>(reinterpret_cast<const float4 *>(d_data), n, min_z,
                                                max_z, num_bins, angle_step, d_bins, tail);
  } else {
    k_cloud_to_bins<<<grid, 256, 0, st>>>(d_data, nbytes, point_step, row_step, height, x_off, y_off,
                                          z_off, min_z, max_z, num_bins, angle_step, d_bins, tail);
  }
  KC_CUDA(cudaGetLastError());
  return KC_OK;
}

// Make a raw PointCloud2 buffer readable by the binning kernel, its only consumer. A page-locked
// caller buffer (kc_pinned_alloc / cudaHostAlloc / cudaHostRegister) is read in place over PCIe: no
// copy at all. Pageable memory goes through the handle's pinned staging buffer in pieces copied by the
// process's copy pool, the DMA of finished pieces overlapping the host copy of the others. *dev receives the pointer the kernel reads;
// *resident tells whether that is the handle's own device copy (replay needs one).
// KOMPASS_B200_STAGE=dma: pageable clouds are DMA-ed from the staging buffer piece by piece;
// default: the binning kernel reads the staging buffer in place, like a page-locked caller buffer
inline bool stage_in_place() {
  static const bool v = [] {
    const char *e = getenv("KOMPASS_B200_STAGE");
    return !(e && std::strcmp(e, "dma") == 0);
  }();
  return v;
}
int32_t stage_cloud(cudaStream_t st, const int8_t *data, int64_t nbytes, PinnedBuf<uint8_t> &h_stage,
                    DevBuf<int8_t> &d_raw, const int8_t **dev, bool *resident, bool *in_stage) {
  KC_TRY(d_raw.reserve((size_t)std::max<int64_t>(nbytes, 16)));
  *dev = d_raw.ptr;
  *resident = true;
  *in_stage = false;
  if (nbytes <= 0) return KC_OK;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, data) == cudaSuccess && a.type == cudaMemoryTypeHost) {
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, const_cast<int8_t *>(data), 0) == cudaSuccess && dp) {
      *dev = static_cast<const int8_t *>(dp);
      *resident = false;
      return KC_OK;
    }
  }
  cudaGetLastError();
  KC_TRY(h_stage.reserve((size_t)nbytes));
  void *sp = nullptr;
  if (stage_in_place() && cudaHostGetDevicePointer(&sp, h_stage.ptr, 0) == cudaSuccess && sp) {
    // the copy pool's threads and this one copy pieces side by side (kc_hostcopy.h); the kernel then
    // pulls the staging buffer over PCIe itself, as it does with a page-locked caller buffer
    CopyPool::instance().copy(h_stage.ptr, reinterpret_cast<const uint8_t *>(data), (size_t)nbytes,
                              CopyPool::piece_for((size_t)nbytes), [](size_t, size_t) {}, (size_t)nbytes);
    *dev = static_cast<const int8_t *>(sp);
    *resident = false;
    *in_stage = true;
    return KC_OK;
  }
  cudaGetLastError();
  // finished runs of pieces go to the DMA engine in order, so the PCIe transfer overlaps the rest of
  // the host copy
  cudaError_t err = cudaSuccess;
  CopyPool::instance().copy(h_stage.ptr, reinterpret_cast<const uint8_t *>(data), (size_t)nbytes,
                            CopyPool::piece_for((size_t)nbytes), [&](size_t off, size_t len) {
                              if (err == cudaSuccess)
                                err = cudaMemcpyAsync(d_raw.ptr + off, h_stage.ptr + off, len,
                                                      cudaMemcpyHostToDevice, st);
                            }, (size_t)nbytes / 3);
  KC_CUDA(err);
  return KC_OK;
}

// replay needs a resident copy: a cloud that was read in place from the staging buffer is uploaded now
int32_t make_resident(cudaStream_t st, PinnedBuf<uint8_t> &h_stage, DevBuf<int8_t> &d_raw, int64_t nbytes,
                      const int8_t **dev, bool *resident, bool *in_stage) {
  if (*resident || !*in_stage) return KC_OK;
  KC_CUDA(cudaMemcpyAsync(d_raw.ptr, h_stage.ptr, (size_t)nbytes, cudaMemcpyHostToDevice, st));
  *dev = d_raw.ptr;
  *resident = true;
  *in_stage = false;
  return KC_OK;
}

}  // namespace

// =================================================================================================
// mapper handle
// =================================================================================================
struct kc_mapper {
  kc_mapper_config cfg;
  MapParams mp;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  DevBuf<int> d_grid;
  // Bayesian mapper state (allocated on first use): posterior, previous grid (+ warp scratch), keys
  BayesParams bp{0.5f, 0.6f, 0.4f, 1.0f, 20.0f, 0.2f};  // local_mapper.h:22-25 defaults
  DevBuf<float> d_prob, d_prev, d_prev_tmp;
  DevBuf<unsigned long long> d_keys;
  PinnedBuf<float> h_prob;
  bool bayes_ready = false;
  DevBuf<double> d_scan;  // angles | ranges
  DevBuf<double> d_init_angles;
  DevBuf<int8_t> d_raw;
  DevBuf<unsigned int> d_bins;
  PinnedBuf<uint8_t> h_stage;
  PinnedBuf<int> h_grid;
  int last_n = 0;
  bool last_cloud = false;
  // cached launch graph of scanToGrid(angles, ranges)
  cudaGraphExec_t scan_graph = nullptr;
  int scan_graph_n = -1;
  const void *scan_graph_dst = nullptr, *scan_graph_stage = nullptr, *scan_graph_dev = nullptr;
  const int8_t *cloud_dev = nullptr;  // where the binning kernel reads the last cloud
  bool cloud_resident = true;         // false: a page-locked caller buffer read in place
  bool cloud_in_stage = false;        // ... or the handle's own staging buffer (replay uploads it)
  // last cloud call geometry (replay)
  int64_t last_nbytes = 0;
  int last_ps = 0, last_rs = 0, last_h = 0, last_xo = 0, last_yo = 0, last_zo = 0;
};

namespace {
// one warp per ray, 8 rays per CTA; more rays than resident warps are strided over
constexpr int kRayThreads = 256;
inline int ray_blocks(int n_rays) {
  return std::max(1, std::min((n_rays + 7) / 8, 16 * sm_count()));
}
int32_t mapper_run_scan(kc_mapper *m, int n) {
  const size_t cells = (size_t)m->cfg.grid_height * m->cfg.grid_width;
  KC_CUDA(cudaMemsetAsync(m->d_grid.ptr, 0xFF, cells * 4, m->stream));  // UNEXPLORED = -1
  if (n > 0) {
    k_scan_to_grid<false><<<ray_blocks(n), kRayThreads, 0, m->stream>>>(
        m->mp, m->d_scan.ptr, m->d_scan.ptr + n, nullptr, 0.0, n, m->d_grid.ptr);
    KC_CUDA(cudaGetLastError());
  }
  return KC_OK;
}
int32_t mapper_run_cloud(kc_mapper *m) {
  const size_t cells = (size_t)m->cfg.grid_height * m->cfg.grid_width;
  const int bins = m->cfg.scan_size;
  KC_TRY(launch_binning(m->stream, m->cloud_dev, m->last_nbytes, m->last_ps, m->last_rs, m->last_h,
                        m->last_xo, m->last_yo, m->last_zo, (double)m->cfg.min_height,
                        (double)m->cfg.max_height, bins, m->d_bins.ptr));
  KC_CUDA(cudaMemsetAsync(m->d_grid.ptr, 0xFF, cells * 4, m->stream));
  k_scan_to_grid<true><<<ray_blocks(bins), kRayThreads, 0, m->stream>>>(
      m->mp, m->d_init_angles.ptr, nullptr, m->d_bins.ptr, (double)m->cfg.range_max, bins,
      m->d_grid.ptr);
  KC_CUDA(cudaGetLastError());
  return KC_OK;
}
// page-locked caller memory (kc_pinned_alloc / cudaHostAlloc / cudaHostRegister) is the target of the
// DMA itself; anything else goes through the handle's pinned buffer and a host copy
bool is_page_locked(const void *q) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, q) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}
int32_t mapper_fetch(kc_mapper *m, int32_t *grid_out) {
  const size_t cells = (size_t)m->cfg.grid_height * m->cfg.grid_width;
  if (is_page_locked(grid_out)) {
    KC_CUDA(cudaMemcpyAsync(grid_out, m->d_grid.ptr, cells * 4, cudaMemcpyDeviceToHost, m->stream));
    KC_CUDA(cudaStreamSynchronize(m->stream));
    return KC_OK;
  }
  KC_CUDA(cudaMemcpyAsync(m->h_grid.ptr, m->d_grid.ptr, cells * 4, cudaMemcpyDeviceToHost, m->stream));
  KC_CUDA(cudaStreamSynchronize(m->stream));
  memcpy(grid_out, m->h_grid.ptr, cells * 4);
  return KC_OK;
}
}  // namespace

extern "C" {

int32_t kc_mapper_create(const kc_mapper_config *cfg, kc_mapper **out) {
  KC_REQUIRE(cfg && out, KC_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  KC_REQUIRE(cfg->grid_height > 0 && cfg->grid_width > 0 && cfg->resolution > 0.0f,
             KC_ERR_INVALID_ARG, "grid dimensions and resolution must be positive");
  KC_REQUIRE(!cfg->is_pointcloud || cfg->scan_size > 0, KC_ERR_INVALID_ARG,
             "scan_size must be positive for point-cloud input");
  KC_TRY(ensure_device());
  kc_mapper *m = new kc_mapper();
  m->cfg = *cfg;
  MapParams &mp = m->mp;
  mp.H = cfg->grid_height;
  mp.W = cfg->grid_width;
  mp.res = cfg->resolution;
  mp.px = cfg->laserscan_position[0];
  mp.py = cfg->laserscan_position[1];
  mp.orient = cfg->laserscan_orientation;
  // ref local_mapper.h:26-27: round(H / 2) - 1 with integer division
  mp.c0 = (int)std::round(cfg->grid_height / 2) - 1;
  mp.c1 = (int)std::round(cfg->grid_width / 2) - 1;
  mp.s0 = mp.c0 + static_cast<int>(mp.px / mp.res);
  mp.s1 = mp.c1 + static_cast<int>(mp.py / mp.res);
  m->bp.range_max = cfg->range_max;
  cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&m->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&m->ev1);
  if (e != cudaSuccess) {
    delete m;
    return cuda_fail(e, "stream/event creation", __FILE__, __LINE__);
  }
  const size_t cells = (size_t)mp.H * mp.W;
  int32_t rc = m->d_grid.reserve(cells);
  if (rc == KC_OK) rc = m->h_grid.reserve(cells);
  if (rc == KC_OK && cfg->is_pointcloud) {
    // ref local_mapper.h:39-55: initializedAngles[i] = i * (2*pi / scanSize)
    const int n = cfg->scan_size;
    std::vector<double> a(n);
    const double step = (2.0 * M_PI) / static_cast<double>(n);
    for (int i = 0; i < n; ++i) a[i] = i * step;
    rc = m->d_init_angles.reserve(n);
    if (rc == KC_OK) rc = m->d_bins.reserve(n);
    if (rc == KC_OK &&
        cudaMemcpy(m->d_init_angles.ptr, a.data(), (size_t)n * 8, cudaMemcpyHostToDevice) != cudaSuccess)
      rc = cuda_fail(cudaGetLastError(), "angle upload", __FILE__, __LINE__);
  }
  if (rc != KC_OK) {
    kc_mapper_destroy(m);
    return rc;
  }
  *out = m;
  return KC_OK;
}

void kc_mapper_destroy(kc_mapper *m) {
  kc::ensure_device();  // the handle's device on this thread (frees below)
  if (!m) return;
  if (m->stream) cudaStreamSynchronize(m->stream);
  if (m->scan_graph) cudaGraphExecDestroy(m->scan_graph);
  m->d_grid.release();
  m->d_prob.release();
  m->d_prev.release();
  m->d_prev_tmp.release();
  m->d_keys.release();
  m->h_prob.release();
  m->d_scan.release();
  m->d_init_angles.release();
  m->d_raw.release();
  m->d_bins.release();
  m->h_stage.release();
  m->h_grid.release();
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int32_t kc_mapper_scan_to_grid(kc_mapper *m, const double *angles, const double *ranges, int32_t n,
                               int32_t *grid_out) {
  KC_REQUIRE(m && grid_out, KC_ERR_INVALID_ARG, "null argument");
  KC_REQUIRE(n >= 0 && (n == 0 || (angles && ranges)), KC_ERR_INVALID_ARG, "bad scan arrays");
  KC_TRY(kc::ensure_device());
  const size_t cells = (size_t)m->cfg.grid_height * m->cfg.grid_width;
  if (n > 0) {
    KC_TRY(m->d_scan.reserve(2 * (size_t)n));
    KC_TRY(m->h_stage.reserve(16 * (size_t)n));
    memcpy(m->h_stage.ptr, angles, (size_t)n * 8);
    memcpy(m->h_stage.ptr + (size_t)n * 8, ranges, (size_t)n * 8);
  }
  m->last_n = n;
  m->last_cloud = false;
  // upload + fill + ray kernel + read-back are the same four stream operations every call: captured
  // once per (beam count, destination) into a CUDA graph and replayed with a single launch
  const bool direct = is_page_locked(grid_out);
  int32_t *dst = direct ? grid_out : m->h_grid.ptr;
  if (!m->scan_graph || m->scan_graph_n != n || m->scan_graph_dst != dst || m->scan_graph_stage != m->h_stage.ptr ||
      m->scan_graph_dev != m->d_scan.ptr) {
    if (m->scan_graph) cudaGraphExecDestroy(m->scan_graph);
    m->scan_graph = nullptr;
    cudaGraph_t g = nullptr;
    KC_CUDA(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = cudaSuccess;
    if (n > 0)
      e = cudaMemcpyAsync(m->d_scan.ptr, m->h_stage.ptr, 16 * (size_t)n, cudaMemcpyHostToDevice, m->stream);
    int32_t rc = (e == cudaSuccess) ? mapper_run_scan(m, n) : KC_ERR_CUDA;
    if (rc == KC_OK) e = cudaMemcpyAsync(dst, m->d_grid.ptr, cells * 4, cudaMemcpyDeviceToHost, m->stream);
    const cudaError_t ec = cudaStreamEndCapture(m->stream, &g);
    if (rc != KC_OK || e != cudaSuccess || ec != cudaSuccess) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      KC_REQUIRE(false, KC_ERR_CUDA, "capturing the scan-to-grid launch set failed");
    }
    const cudaError_t ei = cudaGraphInstantiate(&m->scan_graph, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) m->scan_graph = nullptr;
    KC_CUDA(ei);
    m->scan_graph_n = n;
    m->scan_graph_dst = dst;
    m->scan_graph_stage = m->h_stage.ptr;
    m->scan_graph_dev = m->d_scan.ptr;
  }
  KC_CUDA(cudaGraphLaunch(m->scan_graph, m->stream));
  KC_CUDA(cudaStreamSynchronize(m->stream));
  if (!direct) memcpy(grid_out, m->h_grid.ptr, cells * 4);
  return KC_OK;
}

int32_t kc_mapper_cloud_to_grid(kc_mapper *m, const int8_t *data, int64_t nbytes, int32_t point_step,
                                int32_t row_step, int32_t height, int32_t width, float x_offset,
                                float y_offset, float z_offset, int32_t *grid_out) {
  (void)width;
  KC_REQUIRE(m && grid_out, KC_ERR_INVALID_ARG, "null argument");
  KC_TRY(kc::ensure_device());
  KC_REQUIRE(m->cfg.is_pointcloud, KC_ERR_INVALID_ARG,
             "mapper was not constructed for point-cloud input");
  KC_REQUIRE(nbytes >= 0 && (nbytes == 0 || data), KC_ERR_INVALID_ARG, "bad cloud buffer");
  KC_REQUIRE(x_offset >= 0 && y_offset >= 0 && z_offset >= 0, KC_ERR_INVALID_ARG,
             "negative field offset");
  KC_TRY(stage_cloud(m->stream, data, nbytes, m->h_stage, m->d_raw, &m->cloud_dev, &m->cloud_resident,
                     &m->cloud_in_stage));
  m->last_nbytes = nbytes;
  m->last_ps = point_step;
  m->last_rs = row_step;
  m->last_h = height;
  m->last_xo = (int)x_offset;
  m->last_yo = (int)y_offset;
  m->last_zo = (int)z_offset;
  m->last_cloud = true;
  KC_TRY(mapper_run_cloud(m));
  return mapper_fetch(m, grid_out);
}

// ---- Bayesian mapper (row f2) ------------------------------------------------------------------
static int32_t bayes_prepare(kc_mapper *m) {
  if (m->bayes_ready) return KC_OK;
  const size_t cells = (size_t)m->cfg.grid_height * m->cfg.grid_width;
  KC_TRY(m->d_prob.reserve(cells));
  KC_TRY(m->d_prev.reserve(cells));
  KC_TRY(m->d_prev_tmp.reserve(cells));
  KC_TRY(m->d_keys.reserve(cells));
  KC_TRY(m->h_prob.reserve(cells));
  // previousGridDataProb.fill(m_pPrior) (local_mapper.h:33-34)
  for (size_t i = 0; i < cells; ++i) m->h_prob.ptr[i] = m->bp.p_prior;
  KC_CUDA(cudaMemcpyAsync(m->d_prev.ptr, m->h_prob.ptr, cells * 4, cudaMemcpyHostToDevice, m->stream));
  KC_CUDA(cudaStreamSynchronize(m->stream));
  m->bayes_ready = true;
  return KC_OK;
}

static int32_t bayes_finish(kc_mapper *m, int32_t *grid_out, float *prob_out) {
  const size_t cells = (size_t)m->cfg.grid_height * m->cfg.grid_width;
  const int gb = std::max(1, std::min((int)((cells + 255) / 256), 4 * sm_count()));
  k_bayes_finalize<<<gb, 256, 0, m->stream>>>(m->d_keys.ptr, m->bp.p_prior, cells, m->d_prob.ptr);
  KC_CUDA(cudaGetLastError());
  const bool direct = is_page_locked(grid_out) && is_page_locked(prob_out);
  KC_CUDA(cudaMemcpyAsync(direct ? grid_out : m->h_grid.ptr, m->d_grid.ptr, cells * 4, cudaMemcpyDeviceToHost,
                          m->stream));
  KC_CUDA(cudaMemcpyAsync(direct ? prob_out : m->h_prob.ptr, m->d_prob.ptr, cells * 4, cudaMemcpyDeviceToHost,
                          m->stream));
  KC_CUDA(cudaStreamSynchronize(m->stream));
  if (!direct) {
    memcpy(grid_out, m->h_grid.ptr, cells * 4);
    memcpy(prob_out, m->h_prob.ptr, cells * 4);
  }
  return KC_OK;
}

// ref: local_mapper.h:58-75 (the Bayesian constructor's extra arguments)
int32_t kc_mapper_set_bayesian_params(kc_mapper *m, float p_prior, float p_occupied, float p_empty,
                                      float range_sure, float wall_size) {
  KC_REQUIRE(m, KC_ERR_INVALID_ARG, "null handle");
  KC_REQUIRE(p_prior > 0.0f && p_prior < 1.0f && p_occupied > 0.0f && p_occupied < 1.0f &&
                 p_empty > 0.0f && p_empty < 1.0f,
             KC_ERR_OUT_OF_RANGE, "probabilities must lie in (0, 1)");
  KC_TRY(kc::ensure_device());
  m->bp.p_prior = p_prior;
  m->bp.p_occupied = p_occupied;
  m->bp.p_empty = p_empty;
  m->bp.range_sure = range_sure;
  m->bp.wall_size = wall_size;
  m->bayes_ready = false;  // the previous grid restarts from the new prior, as a fresh LocalMapper would
  return KC_OK;
}

// ref: local_mapper.cpp:222-238 scanToGridBaysian(angles, ranges) -> (gridData, gridDataProb)
int32_t kc_mapper_scan_to_grid_bayesian(kc_mapper *m, const double *angles, const double *ranges,
                                        int32_t n, int32_t *grid_out, float *prob_out) {
  KC_REQUIRE(m && grid_out && prob_out, KC_ERR_INVALID_ARG, "null argument");
  KC_REQUIRE(n >= 0 && (n == 0 || (angles && ranges)), KC_ERR_INVALID_ARG, "bad scan arrays");
  KC_TRY(kc::ensure_device());
  KC_TRY(bayes_prepare(m));
  const size_t cells = (size_t)m->cfg.grid_height * m->cfg.grid_width;
  if (n > 0) {
    KC_TRY(m->d_scan.reserve(2 * (size_t)n));
    KC_TRY(m->h_stage.reserve(16 * (size_t)n));
    memcpy(m->h_stage.ptr, angles, (size_t)n * 8);
    memcpy(m->h_stage.ptr + (size_t)n * 8, ranges, (size_t)n * 8);
    KC_CUDA(cudaMemcpyAsync(m->d_scan.ptr, m->h_stage.ptr, 16 * (size_t)n, cudaMemcpyHostToDevice,
                            m->stream));
  }
  KC_CUDA(cudaMemsetAsync(m->d_grid.ptr, 0xFF, cells * 4, m->stream));
  KC_CUDA(cudaMemsetAsync(m->d_keys.ptr, 0, cells * 8, m->stream));
  if (n > 0)
    k_scan_to_grid_bayes<false><<<ray_blocks(n), kRayThreads, 0, m->stream>>>(
        m->mp, m->bp, m->d_scan.ptr, m->d_scan.ptr + n, nullptr, 0.0, 0.0, n, m->d_prev.ptr,
        m->d_grid.ptr, m->d_keys.ptr);
  return bayes_finish(m, grid_out, prob_out);
}

// ref: local_mapper.cpp:253-264 scanToGridBaysian(raw cloud): pointCloudToLaserScanFromRaw with the
// ANGLE-STEP overload (pointcloud.h:116-177: ceil(2 pi / angle_step) bins, bin = int(angle /
// angle_step), angles_out[i] = i * angle_step; the step is the constructor's angleStep as given —
// unlike the scanToGrid overload it is not replaced by 2 pi / scan_size), then the scan overload.
int32_t kc_mapper_cloud_to_grid_bayesian(kc_mapper *m, const int8_t *data, int64_t nbytes,
                                         int32_t point_step, int32_t row_step, int32_t height,
                                         int32_t width, float x_offset, float y_offset, float z_offset,
                                         int32_t *grid_out, float *prob_out) {
  (void)width;
  KC_REQUIRE(m && grid_out && prob_out, KC_ERR_INVALID_ARG, "null argument");
  KC_TRY(kc::ensure_device());
  KC_REQUIRE(nbytes >= 0 && (nbytes == 0 || data), KC_ERR_INVALID_ARG, "bad cloud buffer");
  KC_REQUIRE(x_offset >= 0 && y_offset >= 0 && z_offset >= 0, KC_ERR_INVALID_ARG,
             "negative field offset");
  const double step = (double)m->cfg.angle_step;  // m_angleStep is a float member (local_mapper.h:253)
  KC_REQUIRE(step > 0.0, KC_ERR_OUT_OF_RANGE, "angle_step must be positive");
  const double nb = std::ceil(2.0 * M_PI / step);
  KC_REQUIRE(nb >= 1.0 && nb <= 16777216.0, KC_ERR_OUT_OF_RANGE, "angle_step gives %g bins", nb);
  const int bins = (int)nb;
  KC_TRY(bayes_prepare(m));
  KC_TRY(m->d_bins.reserve((size_t)bins));
  const int8_t *cloud_dev = nullptr;
  bool cloud_resident = true, cloud_in_stage = false;
  KC_TRY(stage_cloud(m->stream, data, nbytes, m->h_stage, m->d_raw, &cloud_dev, &cloud_resident, &cloud_in_stage));
  const size_t cells = (size_t)m->cfg.grid_height * m->cfg.grid_width;
  KC_TRY(launch_binning(m->stream, cloud_dev, nbytes, point_step, row_step, height, (int)x_offset,
                        (int)y_offset, (int)z_offset, (double)m->cfg.min_height,
                        (double)m->cfg.max_height, bins, m->d_bins.ptr, step));
  KC_CUDA(cudaMemsetAsync(m->d_grid.ptr, 0xFF, cells * 4, m->stream));
  KC_CUDA(cudaMemsetAsync(m->d_keys.ptr, 0, cells * 8, m->stream));
  k_scan_to_grid_bayes<true><<<ray_blocks(bins), kRayThreads, 0, m->stream>>>(
      m->mp, m->bp, nullptr, nullptr, m->d_bins.ptr, (double)m->cfg.range_max, step, bins,
      m->d_prev.ptr, m->d_grid.ptr, m->d_keys.ptr);
  return bayes_finish(m, grid_out, prob_out);
}

// ref: local_mapper.cpp:17-78 getPreviousGridInCurrentPose
int32_t kc_mapper_previous_grid_in_current_pose(kc_mapper *m, float pos_x, float pos_y,
                                                double orientation) {
  KC_REQUIRE(m, KC_ERR_INVALID_ARG, "null handle");
  KC_TRY(kc::ensure_device());
  KC_TRY(bayes_prepare(m));
  const MapParams &mp = m->mp;
  const int cc0 = mp.c0 + static_cast<int>(pos_x / mp.res), cc1 = mp.c1 + static_cast<int>(pos_y / mp.res);
  const double a = -1 * orientation;
  const double cosT = std::cos(a), sinT = std::sin(a);
  float t[3][3];
  t[0][0] = (float)cosT;
  t[0][1] = (float)-sinT;
  t[0][2] = (float)(0.5 * mp.H - cc1 + (cc0 * sinT - cc1 * cosT));
  t[1][0] = (float)sinT;
  t[1][1] = (float)cosT;
  t[1][2] = (float)(0.5 * mp.W - cc0 - (cc0 * cosT + cc1 * sinT));
  t[2][0] = 0.0f;
  t[2][1] = 0.0f;
  t[2][2] = 1.0f;
  // Eigen compute_inverse_size3: cofactors of column 0 give the determinant (a0 + (a1 + a2)),
  // inverse(i, j) = cofactor<j, i> / det
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return t[i1][j1] * t[i2][j2] - t[i1][j2] * t[i2][j1];
  };
  const float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const float det = c00 * t[0][0] + (c10 * t[1][0] + c20 * t[2][0]);
  const float invdet = 1.0f / det;
  WarpParams wp;
  wp.H = mp.H;
  wp.W = mp.W;
  wp.prior = m->bp.p_prior;
  wp.inv[0] = c00 * invdet;
  wp.inv[1] = c10 * invdet;
  wp.inv[2] = c20 * invdet;
  wp.inv[3] = cof(0, 1) * invdet;
  wp.inv[4] = cof(1, 1) * invdet;
  wp.inv[5] = cof(2, 1) * invdet;
  wp.inv[6] = cof(0, 2) * invdet;
  wp.inv[7] = cof(1, 2) * invdet;
  wp.inv[8] = cof(2, 2) * invdet;
  const size_t cells = (size_t)mp.H * mp.W;
  const int gb = std::max(1, std::min((int)((cells + 255) / 256), 4 * sm_count()));
  k_warp_previous<<<gb, 256, 0, m->stream>>>(wp, m->d_prev.ptr, m->d_prev_tmp.ptr);
  KC_CUDA(cudaGetLastError());
  KC_CUDA(cudaStreamSynchronize(m->stream));
  std::swap(m->d_prev.ptr, m->d_prev_tmp.ptr);
  std::swap(m->d_prev.cap, m->d_prev_tmp.cap);
  return KC_OK;
}

// previousGridDataProb is a protected member in the reference and nothing ever assigns the posterior
// back to it; these two hooks let a caller (and the tests) read it and feed a grid back.
int32_t kc_mapper_get_previous_grid(kc_mapper *m, float *prob_out) {
  KC_REQUIRE(m && prob_out, KC_ERR_INVALID_ARG, "null argument");
  KC_TRY(kc::ensure_device());
  KC_TRY(bayes_prepare(m));
  const size_t cells = (size_t)m->cfg.grid_height * m->cfg.grid_width;
  KC_CUDA(cudaMemcpyAsync(m->h_prob.ptr, m->d_prev.ptr, cells * 4, cudaMemcpyDeviceToHost, m->stream));
  KC_CUDA(cudaStreamSynchronize(m->stream));
  memcpy(prob_out, m->h_prob.ptr, cells * 4);
  return KC_OK;
}

int32_t kc_mapper_set_previous_grid(kc_mapper *m, const float *prob) {
  KC_REQUIRE(m && prob, KC_ERR_INVALID_ARG, "null argument");
  KC_TRY(kc::ensure_device());
  KC_TRY(bayes_prepare(m));
  const size_t cells = (size_t)m->cfg.grid_height * m->cfg.grid_width;
  memcpy(m->h_prob.ptr, prob, cells * 4);
  KC_CUDA(cudaMemcpyAsync(m->d_prev.ptr, m->h_prob.ptr, cells * 4, cudaMemcpyHostToDevice, m->stream));
  KC_CUDA(cudaStreamSynchronize(m->stream));
  return KC_OK;
}

int32_t kc_mapper_replay(kc_mapper *m, int32_t n_iters, float *total_ms) {
  KC_REQUIRE(m && n_iters > 0, KC_ERR_INVALID_ARG, "bad replay arguments");
  if (m->last_cloud)
    KC_TRY(make_resident(m->stream, m->h_stage, m->d_raw, m->last_nbytes, &m->cloud_dev, &m->cloud_resident,
                         &m->cloud_in_stage));
  KC_REQUIRE(!m->last_cloud || m->cloud_resident, KC_ERR_INVALID_ARG,
             "the last cloud was read in place from page-locked caller memory: nothing resident to replay");
  KC_TRY(kc::ensure_device());
  KC_CUDA(cudaEventRecord(m->ev0, m->stream));
  for (int i = 0; i < n_iters; ++i) {
    if (m->last_cloud)
      KC_TRY(mapper_run_cloud(m));
    else
      KC_TRY(mapper_run_scan(m, m->last_n));
  }
  KC_CUDA(cudaEventRecord(m->ev1, m->stream));
  KC_CUDA(cudaStreamSynchronize(m->stream));
  float ms = 0.0f;
  KC_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
  if (total_ms) *total_ms = ms;
  return KC_OK;
}

static int32_t cloud_to_scan(const int8_t *data, int64_t nbytes, int32_t point_step, int32_t row_step,
                             int32_t height, int32_t x_offset, int32_t y_offset, int32_t z_offset,
                             double max_range, double min_z, double max_z, int32_t num_bins,
                             double angle_step, double *ranges_out);

int32_t kc_pointcloud_to_laserscan(const int8_t *data, int64_t nbytes, int32_t point_step,
                                   int32_t row_step, int32_t height, int32_t width, int32_t x_offset,
                                   int32_t y_offset, int32_t z_offset, double max_range,
                                   double min_z, double max_z, int32_t num_bins,
                                   double *ranges_out) {
  (void)width;
  return cloud_to_scan(data, nbytes, point_step, row_step, height, x_offset, y_offset, z_offset,
                       max_range, min_z, max_z, num_bins, 0.0, ranges_out);
}

// ref: pointcloud.h:116-177 (angle_step overload): ceil(2 pi / angle_step) bins, angles i * step
int32_t kc_pointcloud_to_laserscan_step(const int8_t *data, int64_t nbytes, int32_t point_step,
                                        int32_t row_step, int32_t height, int32_t width,
                                        int32_t x_offset, int32_t y_offset, int32_t z_offset,
                                        double max_range, double min_z, double max_z,
                                        double angle_step, int32_t cap, double *ranges_out,
                                        double *angles_out, int32_t *n_bins_out) {
  (void)width;
  KC_REQUIRE(ranges_out && angles_out && n_bins_out, KC_ERR_INVALID_ARG, "null output");
  KC_REQUIRE(angle_step > 0.0, KC_ERR_OUT_OF_RANGE, "angle_step must be positive");
  const double nb = std::ceil(2.0 * M_PI / angle_step);
  KC_REQUIRE(nb >= 1.0 && nb <= (double)cap, KC_ERR_OUT_OF_RANGE,
             "output capacity %d too small for %g bins", cap, nb);
  const int32_t num_bins = (int32_t)nb;
  for (int32_t i = 0; i < num_bins; ++i) angles_out[i] = i * angle_step;
  *n_bins_out = num_bins;
  return cloud_to_scan(data, nbytes, point_step, row_step, height, x_offset, y_offset, z_offset,
                       max_range, min_z, max_z, num_bins, angle_step, ranges_out);
}

static int32_t cloud_to_scan(const int8_t *data, int64_t nbytes, int32_t point_step, int32_t row_step,
                             int32_t height, int32_t x_offset, int32_t y_offset, int32_t z_offset,
                             double max_range, double min_z, double max_z, int32_t num_bins,
                             double angle_step, double *ranges_out) {
  KC_REQUIRE(ranges_out && num_bins > 0, KC_ERR_INVALID_ARG, "bad output");
  KC_REQUIRE(nbytes >= 0 && (nbytes == 0 || data), KC_ERR_INVALID_ARG, "bad cloud buffer");
  KC_REQUIRE(x_offset >= 0 && y_offset >= 0 && z_offset >= 0, KC_ERR_INVALID_ARG,
             "negative field offset");
  KC_TRY(ensure_device());
  DevBuf<int8_t> d_raw;
  DevBuf<unsigned int> d_bins;
  DevBuf<double> d_out;
  int32_t rc = d_raw.reserve((size_t)std::max<int64_t>(nbytes, 16));
  if (rc == KC_OK) rc = d_bins.reserve(num_bins);
  if (rc == KC_OK) rc = d_out.reserve(num_bins);
  auto done = [&](int32_t r) {
    d_raw.release();
    d_bins.release();
    d_out.release();
    return r;
  };
  if (rc != KC_OK) return done(rc);
  if (nbytes > 0 && cudaMemcpy(d_raw.ptr, data, (size_t)nbytes, cudaMemcpyHostToDevice) != cudaSuccess)
    return done(cuda_fail(cudaGetLastError(), "cloud upload", __FILE__, __LINE__));
  rc = launch_binning(0, d_raw.ptr, nbytes, point_step, row_step, height, x_offset, y_offset,
                      z_offset, min_z, max_z, num_bins, d_bins.ptr, angle_step);
  if (rc != KC_OK) return done(rc);
  k_bins_to_ranges<<<(num_bins + 255) / 256, 256>>>(d_bins.ptr, num_bins, max_range, d_out.ptr);
  cudaError_t e = cudaMemcpy(ranges_out, d_out.ptr, (size_t)num_bins * 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return done(cuda_fail(e, "ranges download", __FILE__, __LINE__));
  return done(KC_OK);
}

}  // extern "C"

// =================================================================================================
// critical zone handle
// =================================================================================================
struct kc_critical_zone {
  kc_critical_zone_config cfg;
  CzParams cp;
  int n_angles = 0;
  std::vector<int> fwd, bwd;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  DevBuf<float> d_trig;  // cos | sin
  DevBuf<int> d_idx;     // forward | backward
  DevBuf<int8_t> d_raw;
  DevBuf<unsigned int> d_bins;
  DevBuf<unsigned int> d_ticket;  // CTA counter of the fused binning + zone kernel (self-resetting)
  PinnedBuf<uint8_t> h_stage;
  PinnedBuf<CzRecord> h_rec;      // result record the kernels write and the host watches
  CzRecord *rec_dev = nullptr;    // its device-side address
  const double *stage_dev = nullptr;
  size_t stage_dev_cap = 0;
  uint32_t seq = 0;
  // replay state
  bool last_cloud = false;
  const int8_t *cloud_dev = nullptr;  // where the binning kernel reads the last cloud
  bool cloud_resident = true, cloud_in_stage = false;
  int last_forward = 1;
  int64_t last_nbytes = 0;
  int last_ps = 0, last_rs = 0, last_h = 0, last_xo = 0, last_yo = 0, last_zo = 0;
};

namespace {
// enqueue one check on the handle's stream; the result record carries sequence number z->seq
int32_t cz_launch(kc_critical_zone *z, bool cloud, bool forward) {
  const std::vector<int> &ind = forward ? z->fwd : z->bwd;
  const int n_idx = (int)ind.size();
  const int *d_idx = z->d_idx.ptr + (forward ? 0 : (int)z->fwd.size());
  z->seq += 1;
  if (cloud) {
    CzTail tail{};
    tail.enabled = 1;
    tail.cp = z->cp;
    tail.idx = d_idx;
    tail.n_idx = n_idx;
    tail.cos_a = z->d_trig.ptr;
    tail.sin_a = z->d_trig.ptr + z->n_angles;
    tail.max_range = (double)z->cfg.range_max;
    tail.ticket = z->d_ticket.ptr;
    tail.record = z->rec_dev;
    tail.seq = z->seq;
    return launch_binning(z->stream, z->cloud_dev, z->last_nbytes, z->last_ps, z->last_rs, z->last_h, z->last_xo,
                          z->last_yo, z->last_zo, (double)z->cfg.min_height, (double)z->cfg.max_height,
                          std::max(z->n_angles, 1), z->d_bins.ptr, 0.0, &tail);
  }
  const int n_stage = (z->n_angles <= 5632) ? z->n_angles : 0;  // 44 KB of shared memory at most
  k_critical_zone_scan1<<<1, 1024, (size_t)n_stage * 8, z->stream>>>(z->cp, d_idx, n_idx, z->d_trig.ptr,
                                                                    z->d_trig.ptr + z->n_angles, z->stage_dev,
                                                                    n_stage, z->rec_dev, z->seq);
  KC_CUDA(cudaGetLastError());
  return KC_OK;
}
// watch the record for this call's sequence number (the stream is checked now and then so that a
// failed launch still surfaces as an error)
int32_t cz_wait(kc_critical_zone *z, float *factor_out) {
  volatile CzRecord *rec = z->h_rec.ptr;
  for (unsigned spins = 0; rec->seq != z->seq; ++spins) {
    if ((spins & 0x3fff) == 0x3fff) {
      const cudaError_t q = cudaStreamQuery(z->stream);
      if (q == cudaSuccess) break;
      if (q != cudaErrorNotReady) KC_CUDA(q);
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  if (rec->seq != z->seq) KC_CUDA(cudaStreamSynchronize(z->stream));
  std::atomic_thread_fence(std::memory_order_acquire);
  *factor_out = z->h_rec.ptr->factor;
  return KC_OK;
}
// (re)bind the device-side address of the page-locked staging buffer the scan kernel reads in place
int32_t cz_bind_stage(kc_critical_zone *z, size_t bytes) {
  KC_TRY(z->h_stage.reserve(bytes));
  if (z->stage_dev_cap != z->h_stage.cap) {
    void *dp = nullptr;
    KC_CUDA(cudaHostGetDevicePointer(&dp, z->h_stage.ptr, 0));
    z->stage_dev = static_cast<const double *>(dp);
    z->stage_dev_cap = z->h_stage.cap;
  }
  return KC_OK;
}
}  // namespace

extern "C" {

int32_t kc_critical_zone_create(const kc_critical_zone_config *cfg, const double *angles,
                                int32_t n_angles, kc_critical_zone **out) {
  KC_REQUIRE(cfg && out, KC_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  KC_REQUIRE(n_angles >= 0 && (n_angles == 0 || angles), KC_ERR_INVALID_ARG, "bad angles");
  KC_REQUIRE(cfg->robot_shape >= 0 && cfg->robot_shape <= 2, KC_ERR_INVALID_ARG,
             "Invalid robot geometry type");
  // ref critical_zone_check.cpp:53-57
  KC_REQUIRE(cfg->slowdown_distance > cfg->critical_distance, KC_ERR_INVALID_ARG,
             "SlowDown distance must be greater than the Critical distance!");
  KC_REQUIRE(cfg->cloud_field_type == 0 || cfg->cloud_field_type == KC_FLOAT32, KC_ERR_UNSUPPORTED,
             "only FLOAT32 point fields are supported (the reference CPU path memcpy's floats)");
  KC_TRY(ensure_device());
  kc_critical_zone *z = new kc_critical_zone();
  z->cfg = *cfg;
  z->n_angles = n_angles;
  // ref critical_zone_check.cpp:26-40
  if (cfg->robot_shape == KC_BOX)
    z->cp.robot_radius =
        std::sqrt(std::pow(cfg->robot_dims[0], 2) + std::pow(cfg->robot_dims[1], 2)) / 2;
  else
    z->cp.robot_radius = cfg->robot_dims[0];
  z->cp.critical_distance = cfg->critical_distance;
  z->cp.slowdown_distance = cfg->slowdown_distance;
  const hm::Rigid T = hm::rigid_from_quat(cfg->sensor_rotation, cfg->sensor_position);
  for (int i = 0; i < 9; ++i) z->cp.T[i] = T.R.r[i];
  z->cp.T[9] = T.t[0];
  z->cp.T[10] = T.t[1];
  z->cp.T[11] = T.t[2];
  // ref :47-48 + angles.h:21-29
  const float angle_rad = (float)(cfg->critical_angle * M_PI / 180.0);
  double a = std::fmod((double)(angle_rad / 2) + M_PI, 2 * M_PI);
  if (a < 0) a += 2 * M_PI;
  a -= M_PI;
  const float critical_angle = (float)a;
  // ref :62-85 preset(): float trig tables and the forward / backward index lists (host, once)
  std::vector<float> trig(2 * (size_t)n_angles);
  for (int i = 0; i < n_angles; ++i) {
    const float c = (float)std::cos(angles[i]), s = (float)std::sin(angles[i]);
    trig[i] = c;
    trig[(size_t)n_angles + i] = s;
    const float v[3] = {c, s, 0.0f};
    float Rv[3];
    hm::rot_apply(T.R, v, Rv);
    const float qx = T.t[0] + Rv[0], qy = T.t[1] + Rv[1];
    const float abs_theta = std::abs(std::atan2(qy, qx));
    if (abs_theta <= critical_angle) z->fwd.push_back(i);
    if (abs_theta >= M_PI - critical_angle) z->bwd.push_back(i);
  }
  cudaError_t e = cudaStreamCreateWithFlags(&z->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&z->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&z->ev1);
  if (e != cudaSuccess) {
    delete z;
    return cuda_fail(e, "stream/event creation", __FILE__, __LINE__);
  }
  int32_t rc = z->d_trig.reserve(2 * (size_t)n_angles + 1);
  if (rc == KC_OK) rc = z->d_idx.reserve(z->fwd.size() + z->bwd.size() + 1);
  if (rc == KC_OK) rc = z->d_bins.reserve((size_t)n_angles + 1);
  if (rc == KC_OK) rc = z->d_ticket.reserve(1);
  if (rc == KC_OK) rc = z->h_rec.reserve(1);
  if (rc == KC_OK) {
    memset(z->h_rec.ptr, 0, sizeof(CzRecord));
    void *dp = nullptr;
    e = cudaHostGetDevicePointer(&dp, z->h_rec.ptr, 0);
    if (e == cudaSuccess) e = cudaMemset(z->d_ticket.ptr, 0, 4);
    if (e != cudaSuccess) rc = cuda_fail(e, "result record mapping", __FILE__, __LINE__);
    z->rec_dev = static_cast<CzRecord *>(dp);
  }
  if (rc == KC_OK && n_angles > 0) {
    std::vector<int> idx(z->fwd);
    idx.insert(idx.end(), z->bwd.begin(), z->bwd.end());
    e = cudaMemcpy(z->d_trig.ptr, trig.data(), trig.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !idx.empty())
      e = cudaMemcpy(z->d_idx.ptr, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = cuda_fail(e, "table upload", __FILE__, __LINE__);
  }
  if (rc != KC_OK) {
    kc_critical_zone_destroy(z);
    return rc;
  }
  *out = z;
  return KC_OK;
}

void kc_critical_zone_destroy(kc_critical_zone *z) {
  kc::ensure_device();  // the handle's device on this thread (frees below)
  if (!z) return;
  if (z->stream) cudaStreamSynchronize(z->stream);
  z->d_trig.release();
  z->d_idx.release();
  z->d_raw.release();
  z->d_bins.release();
  z->d_ticket.release();
  z->h_stage.release();
  z->h_rec.release();
  if (z->ev0) cudaEventDestroy(z->ev0);
  if (z->ev1) cudaEventDestroy(z->ev1);
  if (z->stream) cudaStreamDestroy(z->stream);
  delete z;
}

int32_t kc_critical_zone_check_scan(kc_critical_zone *z, const double *ranges, int32_t n,
                                    int32_t forward, float *factor_out) {
  KC_REQUIRE(z && factor_out, KC_ERR_INVALID_ARG, "null argument");
  KC_REQUIRE(n >= z->n_angles && (n == 0 || ranges), KC_ERR_INVALID_ARG,
             "ranges must cover the %d angles given at construction (got %d)", z->n_angles, n);
  KC_TRY(kc::ensure_device());
  // the previous call returned on its result record: its kernel has read the staging buffer already
  KC_TRY(cz_bind_stage(z, (size_t)std::max(z->n_angles, 1) * 8));
  if (z->n_angles > 0) memcpy(z->h_stage.ptr, ranges, (size_t)z->n_angles * 8);
  z->last_cloud = false;
  z->last_forward = forward ? 1 : 0;
  KC_TRY(cz_launch(z, false, forward != 0));
  return cz_wait(z, factor_out);
}

int32_t kc_critical_zone_check_cloud(kc_critical_zone *z, const int8_t *data, int64_t nbytes,
                                     int32_t point_step, int32_t row_step, int32_t height,
                                     int32_t width, int32_t x_offset, int32_t y_offset,
                                     int32_t z_offset, int32_t forward, float *factor_out) {
  (void)width;
  KC_REQUIRE(z && factor_out, KC_ERR_INVALID_ARG, "null argument");
  KC_TRY(kc::ensure_device());
  KC_REQUIRE(nbytes >= 0 && (nbytes == 0 || data), KC_ERR_INVALID_ARG, "bad cloud buffer");
  KC_REQUIRE(x_offset >= 0 && y_offset >= 0 && z_offset >= 0, KC_ERR_INVALID_ARG,
             "negative field offset");
  KC_TRY(stage_cloud(z->stream, data, nbytes, z->h_stage, z->d_raw, &z->cloud_dev, &z->cloud_resident,
                     &z->cloud_in_stage));
  z->last_cloud = true;
  z->last_forward = forward ? 1 : 0;
  z->last_nbytes = nbytes;
  z->last_ps = point_step;
  z->last_rs = row_step;
  z->last_h = height;
  z->last_xo = x_offset;
  z->last_yo = y_offset;
  z->last_zo = z_offset;
  KC_TRY(cz_launch(z, true, forward != 0));
  return cz_wait(z, factor_out);
}

int32_t kc_critical_zone_replay(kc_critical_zone *z, int32_t n_iters, float *total_ms) {
  KC_REQUIRE(z && n_iters > 0, KC_ERR_INVALID_ARG, "bad replay arguments");
  if (z->last_cloud)
    KC_TRY(make_resident(z->stream, z->h_stage, z->d_raw, z->last_nbytes, &z->cloud_dev, &z->cloud_resident,
                         &z->cloud_in_stage));
  KC_REQUIRE(!z->last_cloud || z->cloud_resident, KC_ERR_INVALID_ARG,
             "the last cloud was read in place from page-locked caller memory: nothing resident to replay");
  KC_REQUIRE(z->last_cloud || z->stage_dev, KC_ERR_INVALID_ARG, "no check has run on this handle");
  KC_TRY(kc::ensure_device());
  KC_CUDA(cudaEventRecord(z->ev0, z->stream));
  for (int i = 0; i < n_iters; ++i) KC_TRY(cz_launch(z, z->last_cloud, z->last_forward != 0));
  KC_CUDA(cudaEventRecord(z->ev1, z->stream));
  KC_CUDA(cudaStreamSynchronize(z->stream));
  float ms = 0.0f;
  KC_CUDA(cudaEventElapsedTime(&ms, z->ev0, z->ev1));
  if (total_ms) *total_ms = ms;
  return KC_OK;
}

// host build of the libm-compatible atan2f (same source the kernels compile), for CPU tests
float kc_debug_atan2f(float y, float x) { return kc::compat_atan2f(y, x); }
void kc_debug_atan2f_array(const float *y, const float *x, float *out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = kc::compat_atan2f(y[i], x[i]);
}

}  // extern "C"
