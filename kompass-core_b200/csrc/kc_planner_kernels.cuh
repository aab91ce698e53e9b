// kc_planner_kernels.cuh — sm_100a kernels of the DWA hot path.
//
// One RobotCtx per robot lives in device memory; every kernel takes the ctx array and uses
// blockIdx.y as the robot index, so the single-robot control cycle and the batched multi-robot
// sweep run the same code. Pipeline per cycle (three streams inside one captured launch graph, no
// host round trip; k_path_cand and k_rollout_collide run beside the grid preparation):
//
//   k_prep_points      sensor points -> (a) collision voxel-column bitmap of the octree frame,
//                                        (b) cost-frame obstacle points culled to the reachable
//                                            window and counted per cell of a uniform grid
//   k_scan_dist        chained multi-CTA exclusive scan of the per-cell counts + per grid row the
//                      column distance to the nearest occupied cell
//   k_scatter          counting-sort scatter of the kept obstacle points by cell; its first CTAs classify
//                      the query-window cells (one thread per cell) and file the ones that need a list
//   k_cell_cand        one warp per filed cell: distance to the nearest obstacle point and the list of
//                      points that can be the nearest one of any query inside the cell
//   k_cell_cand_heavy  cells next to dense clusters: exact centre distance only, one CTA per cell
//   k_path_class       files the query-window cells inside the reach set for
//   k_path_cand        the same lists over the tracked reference-path segment (path cost)
//   k_dilate           footprint-dilated and sure-hit maps of the bitmap for
//   k_rollout_collide  one warp per tile of four velocity slots: FP64 Euler rollout (bit-identical
//                      floats to the reference) from the heading table k_prep_points fills, per-pose
//                      collision against the bitmap; stores admissible rows
//   k_cost_bounds      (cycles with >= 2048 slots) goal + path cost and a lower / upper bound of the
//   k_cost_split       total of every admissible slot; slots that provably cannot win are filed away
//   k_cost_eval        the remaining cost terms of the slots that are left (all of them without the
//                      bound stage): exact nearest-obstacle distance from the candidate lists - by
//                      (slot, point) batch while few slots survive, by slot otherwise - smoothness /
//                      jerk; the last CTA resolves the argmin (lowest cost, lowest index on ties) and
//                      publishes the winner record
//
// ref: src/utils/trajectory_sampler.cpp:118-275, include/datatypes/path.h:24-30,
//      src/utils/collision_check.cpp:125-162, src/utils/cost_evaluator.cpp:49-233,
//      include/datatypes/trajectory.h:218-235,621-644.
#pragma once
#include "kc_common.cuh"

namespace kc {

constexpr double kMinVel = 0.01;  // ref: include/utils/trajectory_sampler.h:13-15 MIN_VEL
constexpr int kGridN = 256;       // obstacle grid cells per side
constexpr int kGridWords = kGridN / 32;
constexpr int kEvalWarps = 8;     // warps (= velocity slots) per CTA of the trajectory kernels
constexpr int kScanBlocks = kGridN * kGridN / 1024;

#ifdef KC_DBG_STAMPS
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define KC_STAMP_MIN(i) if (threadIdx.x == 0) atomicMin(&cx.dbg[i], gtime())
#define KC_STAMP_MAX(i) if (threadIdx.x == 0) atomicMax(&cx.dbg[i], gtime())
#define KC_STAMP_SET(i, v) if (threadIdx.x == 0 && blockIdx.x == 0) cx.dbg[i] = (unsigned long long)(v)
// whole-cycle timeline (developer): first CTA in / last CTA out of kernel k -> dbg[32 + 2k], dbg[33 + 2k]
struct TlGuard {
  unsigned long long *d;
  __device__ __forceinline__ TlGuard(unsigned long long *dbg, int k) : d(dbg + 32 + 2 * k) {
    if (threadIdx.x == 0) atomicMin(d, gtime());
  }
  __device__ __forceinline__ ~TlGuard() {
    if (threadIdx.x == 0) atomicMax(d + 1, gtime());
  }
};
#define KC_TL(k) TlGuard tl_guard_(cx.dbg, k)
#define KC_PH_DECL const long long p_t0 = clock64()
#define KC_PH(i) if (threadIdx.x == 0) atomicMax(&cx.dbg[80 + (i)], (unsigned long long)(clock64() - p_t0))
#define KC_CAT_DECL const long long c_t0 = clock64(); const unsigned long long c_g0 = gtime(); int c_cat = 0
#define KC_CAT_SET(c) c_cat = (c)
#define KC_CAT_END                                                                   \
  if ((threadIdx.x & 31) == 0) {                                                     \
    const unsigned long long d_ = (unsigned long long)(clock64() - c_t0);            \
    atomicAdd(&cx.dbg[64 + 4 * c_cat], d_);                                          \
    atomicMax(&cx.dbg[65 + 4 * c_cat], d_);                                          \
    atomicAdd(&cx.dbg[66 + 4 * c_cat], 1ull);                                        \
    atomicMax(&cx.dbg[67 + 4 * c_cat], c_g0);                                        \
  }
// rollout phases (developer): cycles since the CTA / warp began, summed and maxed over the grid
#define KC_RSTAMP_DECL const long long r_t0 = clock64(); const unsigned long long r_g0 = gtime()
#define KC_RSTAMP_SUM(i) if (threadIdx.x == 0) atomicAdd(&cx.dbg[i], (unsigned long long)(clock64() - r_t0))
#define KC_RSTAMP_LANE(i) atomicAdd(&cx.dbg[i], (unsigned long long)(clock64() - r_t0))
#define KC_RSTAMP_WARP(isum, imax, icnt)                                              \
  if ((threadIdx.x & 31) == 0) {                                                     \
    const unsigned long long d_ = (unsigned long long)(clock64() - r_t0);            \
    atomicAdd(&cx.dbg[isum], d_);                                                    \
    atomicMax(&cx.dbg[imax], d_);                                                    \
    atomicAdd(&cx.dbg[icnt], 1ull);                                                  \
    atomicMin(&cx.dbg[14], r_g0);                                                    \
    atomicMax(&cx.dbg[15], gtime());                                                 \
  }
#else
#define KC_RSTAMP_DECL
#define KC_RSTAMP_SUM(i)
#define KC_RSTAMP_LANE(i)
#define KC_RSTAMP_WARP(isum, imax, icnt)
#define KC_STAMP_MIN(i)
#define KC_STAMP_MAX(i)
#define KC_STAMP_SET(i, v)
#define KC_PH_DECL
#define KC_PH(i)
#define KC_CAT_DECL
#define KC_CAT_SET(c)
#define KC_CAT_END
#define KC_TL(k)
#endif

struct ResultHeader {
  int32_t found;
  float cost;
  int32_t slot;
  int32_t n_admissible;
  // written last, after a system-scope fence: a host that polls the mapped record sees a complete
  // record as soon as seq equals the cycle's sequence number
  uint32_t seq;
  uint32_t heavy_cells;  // cells whose search disc exceeded the heavy threshold this cycle (host feedback)
  uint32_t pad[2];
};

struct RobotCtx {
  // ---- velocity slots (ref: trajectory_sampler.cpp:181-275,328-372; enumerated on the host) ----
  double pose_x, pose_y, pose_yaw;
  double dt;  // (double)(float)time_step, ref path.h:24 `const float timeStep`
  int32_t P, n_slots, n_rows, nvy, nom;
  int32_t drop_samples;
  long long num_ctrl_points;
  const double *ax_vx, *ax_vy, *ax_om;  // axis values; slot -> (row, local) via row_off
  const int32_t *row_off;               // [n_rows + 1]
  // ---- sensor input ----
  int32_t sensor_is_cloud, n_sensor;
  const void *sensor;  // scan: ranges[n] then angles[n] (double); cloud: xyz floats
  float scan_z;        // laser points' z in the sensor frame (-sensor_z/2, collision_check.h:104)
  // ---- collision world: occupied voxel columns of the octree frame ----
  int32_t coll_enabled, shape;
  double dim0, dim1, dim2, res, res_factor;
  double a00, a01, a10, a11, tx, ty, tz, psi, circ_r;
  double cz, sigma;  // robot centre z in the octree frame; +1 rotation / -1 reflection of the xy block
  int32_t bm_kx0, bm_ky0, bm_cols, bm_rows, bm_wpr;
  uint32_t *bitmap;   // [bm_rows x bm_wpr] one bit per voxel column
  uint32_t *sph_col;  // sphere only: float bits of min dz^2 per column
  // general (tilted sensor) frames: the octree's cubes are oriented boxes in the robot's frame. The
  // bitmap then holds one bit per VOXEL of a 3-D window [g_nz][g_ny][g_wpr words] and the pose test is
  // sphere / box / cylinder against every occupied cube near the body (pose_collides_general)
  int32_t coll_general;
  double gR[9], gt[3];  // octree frame -> world: p_w = R p_s + t (row-major R)
  int32_t g_kx0, g_ky0, g_kz0, g_nx, g_ny, g_nz, g_wpr;
  int32_t dil_W;      // > 0: CTAs keep a copy of the bitmap dilated by +-dil_W columns/rows in smem
  int32_t hit_W;      // voxel columns farther than this from the pose's own column cannot touch the robot
  // hit_W <= 15: rowmask[|dy|] has bit (dx + hit_W) set when the voxel column at offset (dx, dy)
  // from the pose's own column can touch the bounding circle at all (0 = whole row irrelevant)
  int32_t use_rowmask;
  uint32_t rowmask[16];
  float rho;          // bounding-circle radius in voxels (circ_r / res)
  // k_dilate's two derived maps of the bitmap, dil_stride words each (read by k_rollout_collide):
  //   [0] dilated by the row-mask footprint   clear bit: no occupied column can touch a pose of this voxel
  //   [1] dilated by the sure-mask footprint  set bit: some occupied column touches EVERY pose of this voxel
  uint32_t *dil_maps;
  int32_t dil_stride;
  // suremask[|dy|] bit (dx + hit_W): a column at offset (dx, dy) lies within the robot's inscribed
  // circle wherever the pose sits inside its own voxel (0 everywhere: no such shortcut, e.g. spheres)
  uint32_t suremask[16];
  // ---- cost evaluator ----
  float T[12];  // cost-frame transform: R row-major then t (ref cost_evaluator.h:187-189)
  float D;      // maxObstaclesDist
  int32_t obs_enabled, path_enabled;
  double w_path, w_goal, w_obs, w_smooth, w_jerk, dcap2;
  float acc0, acc1, acc2;
  int32_t seg_start, seg_count, path_n;
  float path_len, seg_len;
  float seg_step;  // upper bound of the spacing of consecutive tracked-segment points
  const float *pathX, *pathY, *pathAcc;
  // uniform grid over the cost-frame obstacle points that can matter
  float gx0, gy0, h, inv_h;
  float win_lo_x, win_hi_x, win_lo_y, win_hi_y;
  int32_t *cell_count;   // [N*N + 1]
  int32_t *cell_start;   // [N*N + 1]
  int32_t *cell_cursor;  // [N*N] (the heavy-cell queue)
  uint32_t *occ;         // [N x N/32]
  uint16_t *cell_nn;     // [N*N] squared cell distance to the nearest occupied cell (0xFFFF: none)
  uint16_t *row_dx;      // [N*N], row-major: per grid row the column distance to the nearest
                         // occupied cell of that row (query-window columns only; 0xFFFF: empty row)
  int4 *cell_info;       // [N*N] query-window cells: {dmin float bits, cand start, cand count (-1: overflow), 0}
  float2 *cand_pool;     // nearest-obstacle candidates of the query-window cells (bump allocated)
  int32_t cand_cap;      // capacity of cand_pool
  int32_t *cand_ctr;     // bump counter (zeroed per cycle)
  // cells whose search disc holds more than kHeavyPoints points are queued by k_cell_cand and built by
  // whole CTAs in k_cell_cand_heavy (the queue reuses cell_cursor, free once k_scatter is done)
  int32_t *heavy_ctr;    // queue length (zeroed per cycle)
  // cells whose lists k_cell_cand builds (relevant and reachable), filed by k_scatter's classifying CTAs
  int32_t *work_cells;   // [N*N]
  int32_t *work_ctr;     // queue length (zeroed per cycle)
  int32_t heavy_points;  // disc size that makes a cell "heavy" (<= 0: no cell is)
  int32_t heavy_queue;   // 1: heavy cells are queued for k_cell_cand_heavy; 0: counted, built in place
  int32_t cand_lists;    // 0: no cell builds a candidate list (only its exact centre distance); the few
                         // exact queries of a branch-and-bound cycle then search their own disc
  // the same structure over the TRACKED SEGMENT points (path cost): per query-window cell the
  // segment points that can be the nearest one of any query inside the cell (k_path_cand)
  int32_t pcand_enabled;
  int2 *pcell_info;      // [N*N] {candidate start, candidate count (-1: overflow)}
  float2 *pcand_pool;
  int32_t pcand_cap;
  int32_t *pcand_ctr;    // bump counter (zeroed per cycle)
  int32_t *pwork_cells;  // [N*N] reachable query-window cells, filed by k_path_class for k_path_cand
  int32_t *pwork_ctr;    // queue length (zeroed per cycle)
  uint32_t *blk_tot;     // [kScanBlocks] per-block count totals | ready flag (zeroed per cycle)
  int32_t q_x0, q_x1, q_y0, q_y1;  // cells that can contain trajectory points (cell_nn is valid there)
  // heading table shared by all slots of the cycle (yaw_table_rows): the yaw chain of a slot depends
  // only on its omega, so sincos(yaw before step k) is computed once per (omega, k):
  // tab_sc[row * (P-1) + k] = {sin, cos}, tab_yaw[row * (P-1) + k] = (float)(yaw after step k);
  // rows 0..nom-1 = the omega axis, row nom = omega 0 (omni vy block)
  int32_t tab_rows, tab_ctas;  // the first tab_ctas CTAs of k_prep_points fill the table (0: none)
  double2 *tab_sc;
  float *tab_yaw;
  // analytic reach set of the velocity window (non-holonomic cycles): cells outside it build no
  // candidate lists; a query that lands in one anyway takes the generic exact search, so the mask
  // only has to be a good guess, never a proof
  int32_t reach_mask, rm_n;
  float rm_px, rm_py, rm_yaw, rm_vf, rm_vr, rm_blo, rm_bhi;
  uint32_t *done_ctr;    // blocks of k_cost_eval that finished (zeroed per cycle)
  unsigned long long *best_key;  // packed (ordered cost, slot) argmin (set to ~0 per cycle)
  int32_t *adm_count;    // admissible samples (zeroed per cycle)
  int32_t *n_list;       // admissible slots appended by k_rollout_collide (zeroed per cycle)
  // branch and bound over the slots (k_cost_bounds -> k_cost_eval): every slot gets a lower and an
  // upper bound of its total from the cheap terms + the per-cell distance brackets; slots whose
  // lower bound exceeds the smallest upper bound cannot win and skip the exact obstacle search
  uint32_t seq;          // sequence number of this cycle (ResultHeader::seq)
  int32_t prune;
  int32_t by_point_max;  // survivors up to which k_cost_eval works by (slot, point) pair
  uint32_t *ub_inv;      // ~ordered(min upper bound), zeroed per cycle (0: no bound)
  float *lbv;            // [n_slots] lower bound of the slot's total
  float *ubd;            // [n_slots] upper bracket of the slot's nearest-obstacle distance
  float2 *sjv;           // [n_slots] padded rows: {smoothness, jerk} cost (unweighted), from k_cost_bounds
  uint8_t *prn;          // [n_slots] 1: costs[slot] is only that lower bound (the slot cannot win)
  uint32_t *bounds_done; // CTAs of k_cost_bounds that finished (zeroed per cycle)
  int32_t *n_surv;       // slots that survive the bound test (zeroed per cycle)
  int32_t *surv;         // [n_slots] their ids, unordered
  unsigned long long *dmin_bits;  // [n_slots] survivors: running min d^2 (double bits order as u64)
  unsigned long long *dbg;        // developer time stamps (KC_DBG_STAMPS builds only)
  int32_t *list;         // [n_slots] admissible slot ids, unordered
  int32_t *cutv;         // [n_slots] velocity cut of every slot (P-1 unless padded)
  int2 *tmp_cell;        // [n_sensor] {grid cell (-1: culled), rank of the point inside its cell}
  float2 *tmp_xy;        // [n_sensor]
  float2 *sorted_xy;     // [n_sensor]
  // ---- outputs ----
  float *costs;    // [n_slots]
  uint8_t *adm;    // [n_slots]
  ResultHeader *result;
  float *res_rows;  // vx,vy,om [P-1] each then x,y [P] each
  // sampler mode: all rows stored per slot
  float *rows_vx, *rows_vy, *rows_om, *rows_x, *rows_y;
  // evaluate mode: caller-provided samples
  const float *in_vx, *in_vy, *in_om, *in_x, *in_y;
  const double *custom;  // [n_traj x n_custom] weighted host-callback terms
  int n_custom;
  int32_t n_traj;
};

// ================================================================================================
// k_prep_points
// ================================================================================================
__device__ __forceinline__ bool voxel_key(double res_factor, float c, int &k) {
  // octomap coordToKeyChecked: floor(resolution_factor * coordinate), |key| < 2^15
  const double s = floor(res_factor * (double)c);
  if (!(s >= -32768.0 && s <= 32767.0)) return false;
  k = (int)s;
  return true;
}

__device__ __forceinline__ void yaw_table_rows(const RobotCtx &cx, int cta);

// Bitmap update of a warp's 32 points. One bitmap word covers 32 voxel columns (3.2 m of a 0.1 m
// octree), so the points of a dense cluster - thousands of them - all land in two or three words, and
// their atomics queue up on those addresses in L2 (the kernel's time on such clouds). Lanes that hit
// the same word merge their bits first and one of them sends the update. (The same merge for the
// per-cell counters was measured without gain: a cell is 32 times smaller than a word's span.)
// Every lane of the warp must call this (w < 0: no voxel).
__device__ __forceinline__ void warp_bitmap_or(uint32_t *bitmap, int w, uint32_t bit) {
  const unsigned peers = __match_any_sync(FULL, w);
  const uint32_t bits = __reduce_or_sync(peers, bit);
  if (w >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicOr(&bitmap[w], bits);
}

// When the cycle has rollouts the first tab_ctas CTAs fill the heading table the rollout kernel
// reads (the table and the bitmap are the two inputs of that kernel); the others walk the points.
__global__ void k_prep_points(const RobotCtx *__restrict__ ctxs) {
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(0);
  const int n = cx.n_sensor;
  const int tc = cx.tab_ctas;
  if ((int)blockIdx.x < tc) {
    yaw_table_rows(cx, blockIdx.x);
    return;
  }
  const int nblk = gridDim.x - tc, bx = blockIdx.x - tc;
  const int lane = threadIdx.x & 31;
  for (int base = bx * blockDim.x + (threadIdx.x & ~31); base < n; base += nblk * blockDim.x) {
    const int i = base + lane;
    const bool has = i < n;
    float px = 0.0f, py = 0.0f, pz = 0.0f;  // collision point (octree/sensor frame)
    float qx = 0.0f, qy = 0.0f, qz = 0.0f;  // cost point before the transform
    bool coll_valid = has;
    if (has) {
      if (cx.sensor_is_cloud) {
        const float *xyz = reinterpret_cast<const float *>(cx.sensor);
        px = xyz[3 * i];
        py = xyz[3 * i + 1];
        pz = xyz[3 * i + 2];
        qx = px;
        qy = py;
        qz = pz;
      } else {
        const double *ranges = reinterpret_cast<const double *>(cx.sensor);
        const double *angles = ranges + n;
        const double r = ranges[i], a = angles[i];
        double sn, cs;
        sincos(a, &sn, &cs);
        px = (float)(r * cs);
        py = (float)(r * sn);
        pz = cx.scan_z;
        coll_valid = isfinite(r);  // ref: collision_check.h:111
        qx = px;                   // ref: cost_evaluator.h:184-188 (no finite filter)
        qy = py;
        qz = 0.0f;
      }
    }
    // ---- (a) collision voxel column ----
    int cbw = -1;  // bitmap word / bit of this lane's voxel, set once every ctx read is done
    uint32_t cbbit = 0u;
    if (cx.coll_enabled) {  // uniform
      int bw = -1;          // bitmap word of this lane's voxel (-1: none)
      uint32_t bbit = 0u;
      int kx, ky, kz;
      if (coll_valid && voxel_key(cx.res_factor, px, kx) && voxel_key(cx.res_factor, py, ky) &&
          voxel_key(cx.res_factor, pz, kz)) {
        if (cx.coll_general) {
          const int col = kx - cx.g_kx0, row = ky - cx.g_ky0, lay = kz - cx.g_kz0;
          if (col >= 0 && col < cx.g_nx && row >= 0 && row < cx.g_ny && lay >= 0 && lay < cx.g_nz) {
            bw = (lay * cx.g_ny + row) * cx.g_wpr + (col >> 5);
            bbit = 1u << (col & 31);
          }
        } else {
          const int col = kx - cx.bm_kx0, row = ky - cx.bm_ky0;
          if (col >= 0 && col < cx.bm_cols && row >= 0 && row < cx.bm_rows) {
            const double lo = (double)kz * cx.res, hi = (double)(kz + 1) * cx.res;
            const double cz = cx.cz;
            bool keep = true;
            if (cx.shape == KC_SPHERE) {
              const double dz = fmax(fmax(lo - cz, 0.0), cz - hi);
              const float dz2 = (float)(dz * dz);
              atomicMin(&cx.sph_col[(size_t)row * cx.bm_cols + col], __float_as_uint(dz2));
            } else {
              const double hh = 0.5 * (cx.shape == KC_CYLINDER ? cx.dim1 : cx.dim2);
              keep = (lo <= cz + hh) && (hi >= cz - hh);
            }
            if (keep) {
              bw = row * cx.bm_wpr + (col >> 5);
              bbit = 1u << (col & 31);
            }
          }
        }
      }
      // (reading the word first to skip the atomic when the bit is already there was measured: it puts
      // an L2 round trip in front of every fire-and-forget update, 3 us on ordinary clouds for 1 us on
      // the dense-cluster one)
      cbw = bw;
      cbbit = bbit;
    }
    // Every read of the ctx comes before the first atomic of the iteration: a load issued behind the
    // atomics waits in the memory pipe for them (dense clusters: thousands of updates of a few hot
    // words), which showed as the kernel's largest stall.
    const bool obs_on = cx.obs_enabled != 0;
    uint32_t *const bitmap = cx.bitmap;
    // ---- (b) cost-frame obstacle point ----
    if (obs_on) {  // uniform
      const float *T = cx.T;
      const float ox = T[9] + (T[0] * qx + (T[1] * qy + T[2] * qz));
      const float oy = T[10] + (T[3] * qx + (T[4] * qy + T[5] * qz));
      int cell = -1, ix = 0, iy = 0;
      if (has && ox >= cx.win_lo_x && ox <= cx.win_hi_x && oy >= cx.win_lo_y && oy <= cx.win_hi_y) {
        ix = (int)((ox - cx.gx0) * cx.inv_h);
        iy = (int)((oy - cx.gy0) * cx.inv_h);
        ix = min(max(ix, 0), kGridN - 1);
        iy = min(max(iy, 0), kGridN - 1);
        cell = iy * kGridN + ix;
      }
      // The count the atomicAdd returns is the rank of the point inside its cell, so k_scatter places
      // it without a second round of atomics. (Merging the updates of a warp's points per cell with
      // __match_any_sync was measured: no gain even on the dense-cluster cloud, 1-2 us lost elsewhere.)
      int32_t *const cell_count = cx.cell_count;
      uint32_t *const occ = cx.occ;
      float2 *const tmp_xy = cx.tmp_xy;
      int2 *const tmp_cell = cx.tmp_cell;
      int rank = 0;
      if (cell >= 0) rank = atomicAdd(&cell_count[cell], 1);
      warp_bitmap_or(bitmap, cbw, cbbit);
      if (cell >= 0) {
        // (the same merge for the occupancy words was measured: 0.2 us gained on the dense cloud, 1.3 us
        // lost on the ring - a word spans 32 cells of ONE row, few of a warp's points share it)
        atomicOr(&occ[iy * kGridWords + (ix >> 5)], 1u << (ix & 31));
        tmp_xy[i] = make_float2(ox, oy);
      }
      if (has) tmp_cell[i] = make_int2(cell, rank);
    } else {
      warp_bitmap_or(bitmap, cbw, cbbit);
    }
  }
}

__device__ __forceinline__ bool cell_reachable(const RobotCtx &cx, float cxm, float cym);

// One THREAD per query-window cell decides what k_cell_cand has to do for it (the decision costs a few
// hundred instructions of scalar work - a column walk, an atan2f - which a warp per cell would repeat
// in 32 lanes for all ~10 000 cells, and which kept eight warps' worth of launch slots busy for cells
// that need nothing):
//   * squared cell distance nn to the nearest occupied cell = min over grid rows of dx(row)^2 + dy^2,
//     rows farther than the cost cut-off skipped (such a distance makes the cell irrelevant anyway)
//   * irrelevant (nearest obstacle beyond the cut-off: the cost term is an exact zero) -> record
//     {inf, no list}; relevant but outside the reach set -> {NaN, generic search}
//   * the others are filed in work_cells: k_cell_cand runs one warp per FILED cell, so its launch
//     holds only warps with real work, all resident at once.
// A CTA takes a tile of 32 columns x 8 rows of the query window (thread = cell, warp = 32 neighbours
// of one row): the dx columns of the tile - every grid row within the cut-off of the tile's rows -
// are staged in shared memory with all loads in flight at once, then every thread walks its column.
constexpr int kClassCols = 32, kClassRows = 8;  // tile (kClassCols * kClassRows = k_scatter's 256 threads)
__host__ __device__ inline int class_ctas_for(int qcells) {
  // >= ceil(qw / 32) * ceil(qh / 8) for every window of qcells cells inside the 256 x 256 grid
  return qcells / (kClassCols * kClassRows) + kGridN / kClassCols + kGridN / kClassRows + 1;
}
__device__ __forceinline__ void classify_cells(const RobotCtx &cx, int b) {
  __shared__ uint16_t s_dx[kGridN][kClassCols];
  const int qw = cx.q_x1 - cx.q_x0 + 1, qh = cx.q_y1 - cx.q_y0 + 1;
  const int tiles_x = (qw + kClassCols - 1) / kClassCols, tiles_y = (qh + kClassRows - 1) / kClassRows;
  if (qw <= 0 || qh <= 0 || b >= tiles_x * tiles_y) return;  // CTA-uniform
  KC_PH_DECL;
  const int tx = b % tiles_x, ty = b / tiles_x;
  const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
  const float h = cx.h;
  // relevant <=> sqrt(nn) < T := D * 1.001 / (h * 0.999) + 1.4144; rows within ceil(T) + 1 decide
  const float T = cx.D * 1.001f / (h * 0.999f) + 1.4144f;
  const int cap = (int)fminf(T + 2.0f, (float)kGridN);
  const int ty0 = cx.q_y0 + ty * kClassRows;
  const int rlo = max(0, ty0 - cap), rhi = min(kGridN - 1, ty0 + kClassRows - 1 + cap);
  const int nrows = rhi - rlo + 1;
  const int col0 = cx.q_x0 + tx * kClassCols;
  // (sixteen loads per thread in flight, then the stores: the staging is two round trips to L2)
  for (int base = 0; base < nrows * kClassCols; base += 16 * 256) {
    uint16_t v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = base + u * 256 + threadIdx.x;
      const int r = idx / kClassCols, c = idx % kClassCols;
      v[u] = 0xFFFF;
      if (idx < nrows * kClassCols && col0 + c <= cx.q_x1) v[u] = __ldg(&cx.row_dx[(rlo + r) * kGridN + col0 + c]);
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = base + u * 256 + threadIdx.x;
      if (idx < nrows * kClassCols) s_dx[idx / kClassCols][idx % kClassCols] = v[u];
    }
  }
  __syncthreads();
  KC_PH(0);
  const int ccx = col0 + lane, ccy = ty0 + wrow;
  const bool valid = ccx <= cx.q_x1 && ccy <= cx.q_y1;
  bool file = false;
  const int cell = ccy * kGridN + ccx;
  if (valid) {
    // outwards from the cell's own row, until the row offset alone exceeds the best distance found
    int best = 1 << 30;
    {
      const int dxv = s_dx[ccy - rlo][lane];
      if (dxv != 0xFFFF) best = dxv * dxv;
    }
#pragma unroll 4
    for (int k = 1; k <= cap && k * k < best; ++k) {
      const int ra = ccy - k, rb = ccy + k;
      const int da = (ra >= rlo) ? s_dx[ra - rlo][lane] : 0xFFFF;
      const int db = (rb <= rhi) ? s_dx[rb - rlo][lane] : 0xFFFF;
      if (da != 0xFFFF) best = min(best, da * da + k * k);
      if (db != 0xFFFF) best = min(best, db * db + k * k);
    }
    KC_PH(1);
    // (a minimum beyond the walked rows' reach is not the true one - and irrelevant either way)
    const unsigned nn = (best <= cap * cap) ? (unsigned)min(best, 0xFFFF) : 0xFFFFu;
    cx.cell_nn[cell] = (uint16_t)nn;  // kept for the generic search
    const float rn = sqrtf((float)nn);
    // the centre's nearest point lies within [(rn - 0.7072) h, (rn + 0.7072) h]; a query of this cell
    // is at most another 0.7072 h closer: beyond D the cost term is an exact zero
    const bool relevant = nn != 0xFFFFu && (fmaxf(0.0f, rn - 1.4144f) * h * 0.999f < cx.D * 1.001f);
    const float cxm = cx.gx0 + ((float)ccx + 0.5f) * h, cym = cx.gy0 + ((float)ccy + 0.5f) * h;
    if (!relevant) {
      cx.cell_info[cell] = make_int4(__float_as_int(INFINITY), 0, 0, 0);
    } else if (!cell_reachable(cx, cxm, cym)) {
      // no list: a query that lands here after all takes the generic exact search, without a bracket
      cx.cell_info[cell] = make_int4(__float_as_int(NAN), 0, -1, 0);
    } else {
      file = true;
    }
  }
  const unsigned m = __ballot_sync(FULL, file);
  if (m) {
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(cx.work_ctr, __popc(m));
    base = __shfl_sync(FULL, base, __ffs(m) - 1);
    if (file) cx.work_cells[base + __popc(m & ((1u << lane) - 1u))] = cell;
  }
  KC_PH(2);
#ifdef KC_DBG_STAMPS
  if (threadIdx.x == 0) cx.dbg[85] = (unsigned long long)cap;
#endif
}

// counting-sort scatter: cell_start (k_scan_dist) + the rank k_prep_points drew = the point's place.
// The first n_class CTAs classify the query-window cells instead (classify_cells: both jobs wait for
// k_scan_dist only).
__global__ void k_scatter(const RobotCtx *__restrict__ ctxs, int n_class) {
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(2);
  if (!cx.obs_enabled) return;
  if ((int)blockIdx.x < n_class) {
    classify_cells(cx, blockIdx.x);
    return;
  }
  const int n = cx.n_sensor;
  KC_PH_DECL;
  const int bx = blockIdx.x - n_class, nblk = gridDim.x - n_class;
  for (int i = bx * blockDim.x + threadIdx.x; i < n; i += nblk * blockDim.x) {
    const int2 cr = cx.tmp_cell[i];
    if (cr.x >= 0) cx.sorted_xy[__ldg(&cx.cell_start[cr.x]) + cr.y] = cx.tmp_xy[i];
  }
  KC_PH(4);
}

// ================================================================================================
// k_cell_dist: for every grid cell, the squared distance (in cells, centre to centre) to the nearest
// occupied cell. A query point in cell c then has an obstacle within (sqrt(nn) + sqrt(2)) * h, which
// gives the nearest-obstacle search a tight, guaranteed starting radius instead of the cost cut-off.
// One thread per cell; the 2 KB occupancy bitmask sits in shared memory.
// ================================================================================================
__device__ __forceinline__ int nearest_set_dx(const uint32_t *row, int cx) {
  int best = 1 << 20;
  const int w = cx >> 5, b = cx & 31;
  for (int ww = w; ww >= 0; --ww) {
    uint32_t bits = row[ww];
    if (ww == w) bits &= 0xffffffffu >> (31 - b);
    if (bits) {
      best = cx - (ww * 32 + 31 - __clz(bits));
      break;
    }
  }
  for (int ww = w; ww < kGridWords; ++ww) {
    uint32_t bits = row[ww];
    if (ww == w) bits &= 0xffffffffu << b;
    if (bits) {
      best = min(best, ww * 32 + __ffs(bits) - 1 - cx);
      break;
    }
  }
  return best;
}

// k_scan_dist: grid (kScanBlocks, robots) x 1024 threads, thread <-> cell (coalesced).
//  A. block-local exclusive scan of the per-cell counts; publish the block total (+ ready flag)
//  B. per grid row the column distance to the nearest occupied cell (independent work that hides
//     the wait of C)
//  C. add the totals of the preceding blocks (chained look-back; blocks of one robot are dispatched
//     in index order, so every block a CTA waits for is already running) -> cell_start / cursor
__global__ void __launch_bounds__(1024) k_scan_dist(const RobotCtx *__restrict__ ctxs) {
  __shared__ uint32_t occ[kGridN * kGridWords];
  __shared__ int warp_sums[32];
  __shared__ int s_prefix;
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(1);
  if (!cx.obs_enabled) return;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, b = blockIdx.x;
  const int cell = b * 1024 + t;
  // ---- A
  const int cnt = cx.cell_count[cell];
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(FULL, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) warp_sums[wid] = incl;
  for (int i = t; i < kGridN * kGridWords; i += 1024) occ[i] = cx.occ[i];
  __syncthreads();
  if (wid == 0) {
    int w = warp_sums[lane], wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(FULL, wi, d);
      if (lane >= d) wi += v;
    }
    warp_sums[lane] = wi - w;
    if (lane == 31) {
      __threadfence();
      atomicExch(&cx.blk_tot[b], (uint32_t)wi | 0x80000000u);
    }
  }
  __syncthreads();
  const int local_excl = warp_sums[wid] + incl - cnt;
  // ---- B: per grid row and query-window column the distance (in columns) to the nearest occupied
  //         cell of that row; k_cell_cand combines the rows into the nearest-occupied-cell distance
  {
    const int qw = cx.q_x1 - cx.q_x0 + 1;
    const int total = (qw > 0) ? kGridN * qw : 0;
    for (int idx = b * 1024 + t; idx < total; idx += kScanBlocks * 1024) {
      const int row = idx / qw, col = cx.q_x0 + idx - row * qw;
      const int dx = nearest_set_dx(&occ[row * kGridWords], col);
      cx.row_dx[row * kGridN + col] = (uint16_t)min(dx, 0xFFFF);  // row-major: classify_cells walks a column, lanes = columns
    }
  }
  // ---- C
  if (wid == 0) {
    int sum = 0;
    for (int j = lane; j < b; j += 32) {
      uint32_t v;
      do {
        v = *((volatile uint32_t *)&cx.blk_tot[j]);
      } while (!(v & 0x80000000u));
      sum += (int)(v & 0x7fffffffu);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(FULL, sum, m);
    if (lane == 0) s_prefix = sum;
  }
  __syncthreads();
  const int start = s_prefix + local_excl;
  cx.cell_start[cell] = start;
  if (b == kScanBlocks - 1 && t == 1023) cx.cell_start[kGridN * kGridN] = start + cnt;
}

// ================================================================================================
// k_cell_cand: one warp per query-window cell (cells that can contain trajectory points).
//   pass A  dmin = exact distance from the cell centre to the nearest binned obstacle point; the
//           search disc comes from cell_nn, lanes walk grid rows (the cells [x0, x1] of one row are
//           ONE contiguous range of the cell-sorted points)
//   pass B  every point within dmin + sqrt2*h (+ margin) of the centre is a candidate: for a query q
//           in the cell, |o* - c| <= |o* - q| + |q - c| <= |o_c - q| + |q - c| <= dmin + 2 |q - c|,
//           |q - c| <= h / sqrt2. Lists are staged in shared memory and bump-allocated in the pool;
//           a list that does not fit marks the cell for the generic search (count = -1).
// ================================================================================================
constexpr int kCandWarps = 8;
// (384 -> 512 measured: ring 83 -> 80 us per cycle, the rest of the family unchanged or 1 us better;
// 576 and 704 lose on the dense cloud what they gain on the ring)
#ifndef KC_CAND_BUF
#define KC_CAND_BUF 512
#endif
constexpr int kCandBuf = KC_CAND_BUF;  // staging capacity per warp (points of the search disc, or candidates)
constexpr int kCandSerial = 64;  // lists up to this length are walked by a single lane

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, m));
  return v;
}

// Can a trajectory point of this cycle fall into the cell centred at (cxm, cym)? Euler steps of
// constant (v, omega) are the vertices of a regular polygon: vertex i lies at bearing
// yaw0 + (i-1) a (a = omega dt / 2; + pi when v < 0) and distance |v| dt sin(i a) / sin(a) from the
// start, so for a bearing offset b the farthest vertex over the whole window is the last one of the
// slowest turn that reaches b: |v|max dt sin(b (1 + 1/n)) / sin(b / n), n = P - 2 (the circle's
// diameter once the polygon turns past a quarter). Cell size enters as a radial and angular slack.
__device__ __forceinline__ bool cell_reachable(const RobotCtx &cx, float cxm, float cym) {
  if (!cx.reach_mask) return true;
  const float dx = cxm - cx.rm_px, dy = cym - cx.rm_py;
  const float r = sqrtf(dx * dx + dy * dy);
  const float slack = 1.5f * cx.h;
  if (!(r > 2.0f * slack)) return true;
  const float kPi = 3.14159265f;
  float th = atan2f(dy, dx) - cx.rm_yaw;
  th -= 2.0f * kPi * rintf(th * (0.5f / kPi));
  const float dth = slack / r, n = (float)cx.rm_n;
  auto side = [&](float t, float vdt) {
    if (!(vdt > 0.0f)) return false;
    if (t < cx.rm_blo - dth || t > cx.rm_bhi + dth) {
      // more than a full turn available: every bearing is reachable through the wrapped branch
      if (cx.rm_bhi - cx.rm_blo < 2.0f * kPi) return false;
    }
    const float b = fmaxf(fabsf(t) - dth, 0.0f);
    float rmax = vdt * (n + 1.0f);
    if (b > 1e-3f) {
      const float a1 = b * (1.0f + 1.0f / n);
      rmax = vdt * (a1 < 0.5f * kPi ? __sinf(a1) : 1.0f) / __sinf(b / n);
    }
    return r - slack <= rmax * 1.002f + slack;
  };
  float tr = th - kPi;
  tr -= 2.0f * kPi * rintf(tr * (0.5f / kPi));
  return side(th, cx.rm_vf) || side(tr, cx.rm_vr);
}

__global__ void __launch_bounds__(kCandWarps * 32) k_cell_cand(const RobotCtx *__restrict__ ctxs) {
  __shared__ float2 s_buf[kCandWarps][kCandBuf];
  __shared__ int s_cnt[kCandWarps];
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(3);
  grid_dep_launch();  // k_cell_cand_heavy may be staged behind this grid (it waits for its completion)
  if (!cx.obs_enabled) return;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  KC_CAT_DECL;
  // one warp per FILED cell (k_scatter's classifying CTAs): relevant, inside the reach set
  const int qi = blockIdx.x * kCandWarps + wid;
  if (qi >= __ldcg(cx.work_ctr)) return;  // warp-uniform
  const int cell = __ldcg(&cx.work_cells[qi]);
  const int ccx = cell % kGridN, ccy = cell / kGridN;
  const float h = cx.h;
  const unsigned nn = __ldcg(&cx.cell_nn[cell]);
  float dmin = INFINITY;
  int start = 0, cnt = 0;
  const float rn = sqrtf((float)nn);
  const float cxm = cx.gx0 + ((float)ccx + 0.5f) * h, cym = cx.gy0 + ((float)ccy + 0.5f) * h;
  {
    // nearest point (+ the candidate ring when lists are built)
    const bool lists = cx.cand_lists != 0;
    const float R2 = ((rn + 0.7072f) * 1.003f + (lists ? 1.4143f * 1.006f : 0.01f)) * h;
    const int rc = (int)(R2 * cx.inv_h) + 2;
    const float hh = 0.5f * h * 1.01f;  // half side of the (slightly inflated) cell
    float2 *buf = s_buf[wid];
    // ---- pass A: walk the grid rows of the disc once; every visited point is staged in shared
    // memory (when it fits) and the nearest one to the centre is tracked. Lanes own grid rows; the
    // few rows that cut through the obstacle front hold most of the points, so rows with more than
    // a handful are walked by the whole warp instead.
    float m = INFINITY, mx = 0.0f, my = 0.0f;
    int staged = 0;  // warp-uniform
    int seen = 0;    // points of the disc's rows met so far (warp-uniform)
    bool is_heavy = false;
    auto look = [&](float2 o) {
      const float dx = o.x - cxm, dy = o.y - cym;
      const float d2 = dx * dx + dy * dy;
      if (d2 < m) {
        m = d2;
        mx = dx;
        my = dy;
      }
    };
    for (int iy0 = ccy - rc; iy0 <= ccy + rc; iy0 += 32) {
      const int iy = iy0 + lane;
      int s = 0, e = 0;
      if (iy <= ccy + rc && iy >= 0 && iy < kGridN) {
        const float dyc = fmaxf(fabsf((float)(iy - ccy)) - 0.5f, 0.0f) * h * 0.999f;
        if (dyc < R2) {
          const int half = (int)(sqrtf(R2 * R2 - dyc * dyc) * cx.inv_h) + 2;
          const int x0 = max(0, ccx - half), x1 = min(kGridN - 1, ccx + half);
          s = __ldg(&cx.cell_start[iy * kGridN + x0]);
          e = __ldg(&cx.cell_start[iy * kGridN + x1 + 1]);
        }
      }
      if (cx.heavy_points > 0 && !is_heavy) {
        // A dense cluster (or a wall seen by a depth camera) next to the cell puts thousands of points
        // into the disc, and one warp sifting them twice is a chain of hundreds of dependent round
        // trips. Such a cell is counted and - when the cycle runs k_cell_cand_heavy - queued for it
        // before the expensive part of the walk.
        int chunk = e - s;
#pragma unroll
        for (int mm = 16; mm > 0; mm >>= 1) chunk += __shfl_xor_sync(FULL, chunk, mm);
        seen += chunk;
        if (seen > cx.heavy_points) {
          is_heavy = true;
          int slot = 0;
          if (lane == 0) slot = atomicAdd(cx.heavy_ctr, 1);  // at most once per cell
          if (cx.heavy_queue) {
            if (lane == 0) cx.cell_cursor[slot] = cell;
            return;  // warp-uniform
          }
        }
      }
      unsigned heavy = __ballot_sync(FULL, e - s > 4);
      const bool mine = !((heavy >> lane) & 1u);
      const int lc = mine ? (e - s) : 0;
      int incl = lc;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += u;
      }
      int off = staged + incl - lc;
      if (mine)
        for (int q = s; q < e; ++q, ++off) {
          const float2 o = __ldg(&cx.sorted_xy[q]);
          if (lists && off < kCandBuf) buf[off] = o;
          look(o);
        }
      staged += __shfl_sync(FULL, incl, 31);
      while (heavy) {
        const int src = __ffs(heavy) - 1;
        heavy &= heavy - 1;
        const int sb = __shfl_sync(FULL, s, src), eb = __shfl_sync(FULL, e, src);
        for (int q0 = sb + lane; q0 < eb; q0 += 128) {  // four independent loads in flight per lane
          float2 o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (q0 + 32 * u < eb) o[u] = __ldg(&cx.sorted_xy[q0 + 32 * u]);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (q0 + 32 * u < eb) {
              const int at = staged + (q0 + 32 * u - sb);
              if (lists && at < kCandBuf) buf[at] = o[u];
              look(o[u]);
            }
        }
        staged += eb - sb;
      }
    }
    int src = lane;
    warp_argmin_f(m, src);
    const float ocx = __shfl_sync(FULL, mx, src), ocy = __shfl_sync(FULL, my, src);
    const float dmin2 = m;
    dmin = sqrtf(m);
    const float rad = dmin * 1.002f + 1.4143f * 1.004f * h;
    const float thr2 = rad * rad * 1.0001f;
    const float tol = 2e-6f * thr2;  // > float error of the expression + the reference's own rounding
    // candidate test: inside the ring and not dominated by o_c everywhere in the cell. o can be the
    // nearest point of some query q of the cell only if |o_c - q|^2 - |o - q|^2 >= 0 somewhere in
    // it; the expression is linear in q, so its max sits at a corner:
    //   (|o_c|^2 - |o|^2) + 2 hh (|dx_c - dx| + |dy_c - dy|)      (centre-relative)
    auto keep = [&](float2 o) {
      const float dx = o.x - cxm, dy = o.y - cym;
      const float d2 = dx * dx + dy * dy;
      if (!(d2 <= thr2)) return false;
      const float f = (dmin2 - d2) + 2.0f * hh * (fabsf(ocx - dx) + fabsf(ocy - dy));
      return f >= -tol;
    };
    __syncwarp();
    if (!lists) {
      cnt = -1;  // bracket only; an empty disc cannot happen for consistently binned points (NaN: no bracket)
      if (!(m <= R2 * R2 * 1.01f)) dmin = NAN;
    } else if (!(rad <= R2)) {
      // cannot happen for consistently binned points; stay exact regardless: no bracket (NaN), generic search
      cnt = -1;
      dmin = NAN;
    } else if (staged <= kCandBuf) {
      KC_CAT_SET(staged <= 64 ? 1 : 2);
      // ---- pass B over the staged points: count, allocate, write
      int n = 0;
      for (int i0 = 0; i0 < staged; i0 += 32) {
        const int i = i0 + lane;
        n += __popc(__ballot_sync(FULL, i < staged && keep(buf[i])));
      }
      if (n > 0) {
        int basep = 0;
        if (lane == 0) basep = atomicAdd(cx.cand_ctr, n);
        basep = __shfl_sync(FULL, basep, 0);
        if (basep + n > cx.cand_cap) {
          cnt = -1;
        } else {
          int w = basep;
          for (int i0 = 0; i0 < staged; i0 += 32) {
            const int i = i0 + lane;
            const float2 o = (i < staged) ? buf[i] : make_float2(0.0f, 0.0f);
            const bool k = i < staged && keep(o);
            const unsigned bal = __ballot_sync(FULL, k);
            if (k) cx.cand_pool[w + __popc(bal & ((1u << lane) - 1u))] = o;
            w += __popc(bal);
          }
          start = basep;
          cnt = n;
        }
      }
    } else {
      KC_CAT_SET(3);
      // ---- the disc holds more points than the staging buffer: walk it a second time, collecting
      // the (far fewer) candidates in the buffer
      if (lane == 0) s_cnt[wid] = 0;
      __syncwarp();
      auto visit_o = [&](float2 o) {
        if (keep(o)) {
          const int slot = atomicAdd(&s_cnt[wid], 1);
          if (slot < kCandBuf) buf[slot] = o;
        }
      };
      auto visit = [&](int q) { visit_o(__ldg(&cx.sorted_xy[q])); };
      for (int iy0 = ccy - rc; iy0 <= ccy + rc; iy0 += 32) {
        const int iy = iy0 + lane;
        int s = 0, e = 0;
        if (iy <= ccy + rc && iy >= 0 && iy < kGridN) {
          const float dyc = fmaxf(fabsf((float)(iy - ccy)) - 0.5f, 0.0f) * h * 0.999f;
          if (dyc < R2) {
            const int half = (int)(sqrtf(R2 * R2 - dyc * dyc) * cx.inv_h) + 2;
            const int x0 = max(0, ccx - half), x1 = min(kGridN - 1, ccx + half);
            s = __ldg(&cx.cell_start[iy * kGridN + x0]);
            e = __ldg(&cx.cell_start[iy * kGridN + x1 + 1]);
          }
        }
        unsigned heavy = __ballot_sync(FULL, e - s > 4);
        if (!((heavy >> lane) & 1u))
          for (int q = s; q < e; ++q) visit(q);
        while (heavy) {
          const int src2 = __ffs(heavy) - 1;
          heavy &= heavy - 1;
          const int sb = __shfl_sync(FULL, s, src2), eb = __shfl_sync(FULL, e, src2);
          for (int q0 = sb + lane; q0 < eb; q0 += 128) {  // four independent loads in flight per lane
            float2 o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (q0 + 32 * u < eb) o[u] = __ldg(&cx.sorted_xy[q0 + 32 * u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (q0 + 32 * u < eb) visit_o(o[u]);
          }
        }
      }
      __syncwarp();
      const int n = s_cnt[wid];
      if (n > kCandBuf) {
        cnt = -1;
      } else if (n > 0) {
        int basep = 0;
        if (lane == 0) basep = atomicAdd(cx.cand_ctr, n);
        basep = __shfl_sync(FULL, basep, 0);
        if (basep + n > cx.cand_cap) {
          cnt = -1;
        } else {
          for (int k = lane; k < n; k += 32) cx.cand_pool[basep + k] = buf[k];
          start = basep;
          cnt = n;
        }
      }
    }
  }
  if (lane == 0) cx.cell_info[cell] = make_int4(__float_as_int(dmin), start, cnt, 0);
  KC_CAT_END;
}

// ================================================================================================
// k_cell_cand_heavy: the cells k_cell_cand queued (search disc with more than heavy_points points,
// e.g. next to a dense cluster or a wall seen by a depth camera). Building a candidate list there
// means sifting thousands of points per cell for lists that the branch and bound almost never
// reads. These cells get only what every slot's bound needs - the exact distance dmin from the cell
// centre to its nearest point - and no list (count -1): the few exact queries that land in them run
// warp_nn_search_one, a warp-cooperative exact search over the query's own (tight) disc.
// ONE CTA per cell: the disc's rows are flattened through a prefix sum in shared memory and the 256
// threads stride over the points with kHeavyUnroll independent loads in flight each.
// ================================================================================================
constexpr int kHeavyThreads = 256;
constexpr int kHeavyDefault = 1024;
constexpr int kHeavyUnroll = 8;  // independent loads in flight per thread

__global__ void __launch_bounds__(kHeavyThreads) k_cell_cand_heavy(const RobotCtx *__restrict__ ctxs) {
  __shared__ int s_rs[kGridN], s_off[kGridN + 1];
  __shared__ float s_m[kHeavyThreads / 32];
  __shared__ int s_wsum[kHeavyThreads / 32];
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(4);
  grid_dep_wait();  // the queue and the sorted points come from the kernels before this one
  if (!cx.obs_enabled) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n_heavy = min(*cx.heavy_ctr, kGridN * kGridN);
  const float h = cx.h;
  for (int hi = blockIdx.x; hi < n_heavy; hi += gridDim.x) {
    const int cell = cx.cell_cursor[hi];
    const int ccx = cell % kGridN, ccy = cell / kGridN;
    const unsigned nn = cx.cell_nn[cell];
    const float rn = sqrtf((float)nn);
    const float cxm = cx.gx0 + ((float)ccx + 0.5f) * h, cym = cx.gy0 + ((float)ccy + 0.5f) * h;
    // the nearest occupied cell holds a point within (rn + 0.7072) h of the centre
    const float RA = ((rn + 0.7072f) * 1.003f + 0.01f) * h;
    const int rc = (int)(RA * cx.inv_h) + 2;
    const int iy_lo = max(0, ccy - rc), iy_hi = min(kGridN - 1, ccy + rc);
    const int nrows = iy_hi - iy_lo + 1;  // <= kGridN = blockDim.x
    // row ranges + exclusive prefix of their lengths
    int s = 0, len = 0;
    if (tid < nrows) {
      const int iy = iy_lo + tid;
      const float dyc = fmaxf(fabsf((float)(iy - ccy)) - 0.5f, 0.0f) * h * 0.999f;
      if (dyc < RA) {
        const int half = (int)(sqrtf(RA * RA - dyc * dyc) * cx.inv_h) + 2;
        const int x0 = max(0, ccx - half), x1 = min(kGridN - 1, ccx + half);
        s = __ldg(&cx.cell_start[iy * kGridN + x0]);
        len = __ldg(&cx.cell_start[iy * kGridN + x1 + 1]) - s;
      }
    }
    int incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += u;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < wid; ++w) wbase += s_wsum[w];
    s_rs[tid] = s;
    s_off[tid] = wbase + incl - len;
    if (tid == kHeavyThreads - 1) s_off[kGridN] = wbase + incl;
    __syncthreads();
    const int T = s_off[kGridN];
    // (rows beyond nrows have length 0: their s_off equals T, the row search never reaches them)
    float m = INFINITY;
    int row = 0;
    for (int f0 = tid; f0 < T; f0 += kHeavyUnroll * kHeavyThreads) {
      float2 o[kHeavyUnroll];
      bool v[kHeavyUnroll];
#pragma unroll
      for (int u = 0; u < kHeavyUnroll; ++u) {
        const int f = f0 + u * kHeavyThreads;
        v[u] = f < T;
        if (v[u]) {
          while (f >= s_off[row + 1]) ++row;  // flat index -> row (advances monotonically)
          o[u] = __ldg(&cx.sorted_xy[s_rs[row] + (f - s_off[row])]);
        }
      }
#pragma unroll
      for (int u = 0; u < kHeavyUnroll; ++u)
        if (v[u]) {
          const float dx = o[u].x - cxm, dy = o[u].y - cym;
          m = fminf(m, dx * dx + dy * dy);
        }
    }
    m = warp_min_f(m);
    if (lane == 0) s_m[wid] = m;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < kHeavyThreads / 32; ++w) m = fminf(m, s_m[w]);
      // the same float expression as k_cell_cand's pass A on the same points: the same dmin. A disc
      // that came up empty cannot happen for consistently binned points; stay exact regardless (NaN:
      // no bracket).
      const float dmin = (m <= RA * RA * 1.01f) ? sqrtf(m) : NAN;
      cx.cell_info[cell] = make_int4(__float_as_int(dmin), 0, -1, 1);
    }
    __syncthreads();  // shared state is reused by the next queued cell
  }
}

// ================================================================================================
// k_path_cand: the candidate-list idea of k_cell_cand applied to the tracked path segment (the
// point set of the path cost, cost_evaluator.cpp:111-141). One warp per query-window cell, lanes
// stride over the S segment points (S is a few hundred: no binning needed):
//   pass A  nearest segment point of the cell centre;  pass B  ring + bisector filter, as above.
// A trajectory point then only measures its distance to its own cell's few candidates: the same
// float operations on a superset of the points that can attain the minimum, hence the same min.
// ================================================================================================
constexpr int kPathCandBuf = 256;

// k_path_class: one THREAD per query-window cell tests the reach set (an atan2f and two sines: scalar
// work a warp per cell would repeat in 32 lanes); cells outside it get "no list" (the evaluator's
// direct search covers stray queries), the others are filed for k_path_cand, which then launches
// warps with real work only.
__global__ void __launch_bounds__(256) k_path_class(const RobotCtx *__restrict__ ctxs) {
  const RobotCtx &cx = ctxs[blockIdx.y];
  if (!cx.pcand_enabled) return;
  const int qw = cx.q_x1 - cx.q_x0 + 1, qh = cx.q_y1 - cx.q_y0 + 1;
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool file = false;
  int cell = 0;
  if (qw > 0 && qh > 0 && qi < qw * qh) {
    const int ccx = cx.q_x0 + qi % qw, ccy = cx.q_y0 + qi / qw;
    cell = ccy * kGridN + ccx;
    const float h = cx.h;
    const float cxm = cx.gx0 + ((float)ccx + 0.5f) * h, cym = cx.gy0 + ((float)ccy + 0.5f) * h;
    if (cell_reachable(cx, cxm, cym))
      file = true;
    else
      cx.pcell_info[cell] = make_int2(0, -1);
  }
  const unsigned m = __ballot_sync(FULL, file);
  if (m) {
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(cx.pwork_ctr, __popc(m));
    base = __shfl_sync(FULL, base, __ffs(m) - 1);
    if (file) cx.pwork_cells[base + __popc(m & ((1u << lane) - 1u))] = cell;
  }
}

__global__ void __launch_bounds__(kCandWarps * 32) k_path_cand(const RobotCtx *__restrict__ ctxs) {
  __shared__ float2 s_buf[kCandWarps][kPathCandBuf];
  __shared__ int s_cnt[kCandWarps];
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(5);
  if (!cx.pcand_enabled) return;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * kCandWarps + wid;
  if (qi >= __ldcg(cx.pwork_ctr)) return;  // warp-uniform: one warp per FILED cell
  const int cell = __ldcg(&cx.pwork_cells[qi]);
  const int ccx = cell % kGridN, ccy = cell / kGridN;
  const float h = cx.h;
  const float cxm = cx.gx0 + ((float)ccx + 0.5f) * h, cym = cx.gy0 + ((float)ccy + 0.5f) * h;
  const float *X = cx.pathX + cx.seg_start, *Y = cx.pathY + cx.seg_start;
  const int S = cx.seg_count;
  float m = INFINITY, mx = 0.0f, my = 0.0f;
  for (int j = lane; j < S; j += 32) {
    const float dx = __ldg(&X[j]) - cxm, dy = __ldg(&Y[j]) - cym;
    const float d2 = dx * dx + dy * dy;
    if (d2 < m) {
      m = d2;
      mx = dx;
      my = dy;
    }
  }
  int src = lane;
  warp_argmin_f(m, src);
  const float ocx = __shfl_sync(FULL, mx, src), ocy = __shfl_sync(FULL, my, src);
  const float dmin2 = m;
  const float rad = sqrtf(m) * 1.002f + 1.4143f * 1.004f * h;
  const float thr2 = rad * rad * 1.0001f;
  const float tol = 2e-6f * thr2;
  const float hh = 0.5f * h * 1.01f;
  if (lane == 0) s_cnt[wid] = 0;
  __syncwarp();
  int start = 0, cnt = 0;
  if (!(m < INFINITY)) {
    cnt = -1;  // non-finite segment: let the evaluator scan it as the reference would
  } else {
    for (int j = lane; j < S; j += 32) {
      const float2 o = make_float2(__ldg(&X[j]), __ldg(&Y[j]));
      const float dx = o.x - cxm, dy = o.y - cym;
      const float d2 = dx * dx + dy * dy;
      if (d2 <= thr2) {
        const float f = (dmin2 - d2) + 2.0f * hh * (fabsf(ocx - dx) + fabsf(ocy - dy));
        if (f >= -tol) {
          const int slot = atomicAdd(&s_cnt[wid], 1);
          if (slot < kPathCandBuf) s_buf[wid][slot] = o;
        }
      }
    }
    __syncwarp();
    const int n = s_cnt[wid];
    if (n > kPathCandBuf || n == 0) {
      cnt = -1;
    } else {
      int basep = 0;
      if (lane == 0) basep = atomicAdd(cx.pcand_ctr, n);
      basep = __shfl_sync(FULL, basep, 0);
      if (basep + n > cx.pcand_cap) {
        cnt = -1;
      } else {
        for (int k = lane; k < n; k += 32) cx.pcand_pool[basep + k] = s_buf[wid][k];
        start = basep;
        cnt = n;
      }
    }
  }
  if (lane == 0) cx.pcell_info[ccy * kGridN + ccx] = make_int2(start, cnt);
}

// ================================================================================================
// device building blocks of the trajectory kernels
// ================================================================================================
struct SlotVel {
  double vx, vy, om;
  int row;   // row of the heading table (omega index, or nom for the omega = 0 block)
  int srow;  // sampler row (index of the vx axis value) the slot belongs to
};

// slot -> velocity triple (serial enumeration order of the reference sampler)
__device__ __forceinline__ SlotVel decode_slot(const RobotCtx &cx, int slot) {
  int lo = 0, hi = cx.n_rows;  // largest row with row_off[row] <= slot
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cx.row_off[mid] <= slot)
      lo = mid;
    else
      hi = mid;
  }
  const int local = slot - cx.row_off[lo];
  SlotVel v;
  v.srow = lo;
  v.vx = cx.ax_vx[lo];
  if (local < cx.nvy) {  // omni (vx, vy, 0) block comes first (trajectory_sampler.cpp:258-262)
    v.vy = cx.ax_vy[local];
    v.om = 0.0;
    v.row = cx.nom;
  } else {
    v.vy = 0.0;
    v.om = cx.ax_om[local - cx.nvy];
    v.row = local - cx.nvy;
  }
  return v;
}

// decode_slot by a whole warp: the lanes probe the row offsets in parallel (one round trip to memory
// instead of a dependent binary search); every lane returns the triple
__device__ __forceinline__ SlotVel warp_decode_slot(const RobotCtx &cx, int slot, int lane) {
  // offsets are non-decreasing: (number of rows with row_off[row] <= slot) - 1; the probes are
  // independent loads, so the whole search costs one round trip
  int cnt = 0;
  for (int base = 0; base < cx.n_rows; base += 32) {
    const int r = base + lane;
    cnt += (r < cx.n_rows && cx.row_off[r] <= slot) ? 1 : 0;
  }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
  const int row = max(cnt - 1, 0);
  const int local = slot - cx.row_off[row];
  SlotVel v;
  v.srow = row;
  v.vx = cx.ax_vx[row];
  if (local < cx.nvy) {
    v.vy = cx.ax_vy[local];
    v.om = 0.0;
    v.row = cx.nom;
  } else {
    v.vy = 0.0;
    v.om = cx.ax_om[local - cx.nvy];
    v.row = local - cx.nvy;
  }
  return v;
}

// the triple of slot `slot` >= s0 given the decoded triple v0 of s0 (row index recovered from v0 by
// the caller's lanes walking forward: tiles are short, so at most a row boundary or two away)
__device__ __forceinline__ SlotVel next_slot(const RobotCtx &cx, const SlotVel &v0, int s0, int slot) {
  if (slot == s0) return v0;
  int row = v0.srow;
  while (row + 1 < cx.n_rows && cx.row_off[row + 1] <= slot) ++row;
  const int local = slot - cx.row_off[row];
  SlotVel v;
  v.srow = row;
  v.vx = (row == v0.srow) ? v0.vx : cx.ax_vx[row];
  if (local < cx.nvy) {
    v.vy = cx.ax_vy[local];
    v.om = 0.0;
    v.row = cx.nom;
  } else {
    v.vy = 0.0;
    v.om = cx.ax_om[local - cx.nvy];
    v.row = local - cx.nvy;
  }
  return v;
}

// One warp per table row: the yaw chain `yaw += omega * dt` in the reference's serial rounding
// order (every lane re-adds up to its own step), then sincos per lane (ref: path.h:24-30).
__device__ __forceinline__ void yaw_table_rows(const RobotCtx &cx, int cta) {
  const int P = cx.P, lane = threadIdx.x & 31;
  const int row = cta * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= cx.tab_rows) return;
  const double om = (row < cx.nom) ? cx.ax_om[row] : 0.0;
  const double w = om * cx.dt;
  double YAW = cx.pose_yaw;
  double2 *tab = cx.tab_sc + (size_t)row * (P - 1);
  float *tyaw = cx.tab_yaw + (size_t)row * (P - 1);
  for (int base = 0; base < P - 1; base += 32) {
    double yk = YAW;  // yaw before step (base + lane): sequential adds keep the rounding order
    const int cnt = min(32, P - 1 - base);
    const int chain = (base + 32 < P - 1) ? 31 : cnt - 1;  // the last batch needs no carry-out
    for (int j = 0; j < chain; ++j)
      if (j < lane) yk = yk + w;
    double s, c;
    sincos(yk, &s, &c);
    if (lane < cnt) {
      tab[base + lane] = make_double2(s, c);
      tyaw[base + lane] = (float)(yk + w);
    }
    YAW = shfl_d(yk, 31) + w;
  }
}

// exact closed robot-vs-voxel-column test (same operation order as the oracle's columnHit)
__device__ __forceinline__ bool column_hit(const RobotCtx &cx, int kx, int ky, float dz2f, double pcx,
                                           double pcy, double cth, double sth) {
  const double lox = (double)kx * cx.res, hix = (double)(kx + 1) * cx.res;
  const double loy = (double)ky * cx.res, hiy = (double)(ky + 1) * cx.res;
  if (cx.shape == KC_BOX) {
    const double a = 0.5 * cx.dim0, b = 0.5 * cx.dim1;
    const double ex = 0.5 * (hix - lox), ey = 0.5 * (hiy - loy);
    const double dx = 0.5 * (lox + hix) - pcx, dy = 0.5 * (loy + hiy) - pcy;
    const double ac = fabs(cth), as = fabs(sth);
    if (fabs(dx) > ex + (a * ac + b * as)) return false;
    if (fabs(dy) > ey + (a * as + b * ac)) return false;
    if (fabs(dx * cth + dy * sth) > a + (ex * ac + ey * as)) return false;
    if (fabs(dy * cth - dx * sth) > b + (ex * as + ey * ac)) return false;
    return true;
  }
  const double dx = fmax(fmax(lox - pcx, 0.0), pcx - hix);
  const double dy = fmax(fmax(loy - pcy, 0.0), pcy - hiy);
  const double r = cx.dim0;
  double d2 = dx * dx + dy * dy;
  if (cx.shape == KC_SPHERE) d2 = d2 + (double)dz2f;
  return d2 <= r * r;
}

// ---- general (tilted sensor) frames (general voxel model): every operation in double, in a
// fixed order, so that the parity checker's CPU restatement reproduces each boolean ----
struct Obb {
  double c[3];      // cube centre in the body frame
  double ax[3][3];  // ax[j] = cube axis j in the body frame
  double e;         // half side
};

__device__ __forceinline__ bool sphere_hits_obb(double r, const Obb &b) {
  double d2 = 0.0;
  for (int j = 0; j < 3; ++j) {
    const double u = fabs(b.c[0] * b.ax[j][0] + b.c[1] * b.ax[j][1] + b.c[2] * b.ax[j][2]);
    const double ex = fmax(u - b.e, 0.0);
    d2 = d2 + ex * ex;
  }
  return d2 <= r * r;
}

__device__ __forceinline__ bool box_hits_obb(const double a[3], const Obb &b) {  // 15 separating axes
  double Rm[3][3], A[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Rm[i][j] = b.ax[j][i];
      A[i][j] = fabs(Rm[i][j]);
    }
  const double e = b.e;
  for (int i = 0; i < 3; ++i)
    if (fabs(b.c[i]) > a[i] + e * (A[i][0] + A[i][1] + A[i][2])) return false;
  for (int j = 0; j < 3; ++j) {
    const double tj = b.c[0] * Rm[0][j] + b.c[1] * Rm[1][j] + b.c[2] * Rm[2][j];
    if (fabs(tj) > (a[0] * A[0][j] + a[1] * A[1][j] + a[2] * A[2][j]) + e) return false;
  }
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double lhs = fabs(b.c[i2] * Rm[i1][j] - b.c[i1] * Rm[i2][j]);
      const double ra = a[i1] * A[i2][j] + a[i2] * A[i1][j];
      const double rb = e * A[i][j1] + e * A[i][j2];
      if (lhs > ra + rb) return false;
    }
  }
  return true;
}

__device__ __forceinline__ double seg_dist2_origin(const double *a, const double *b) {
  const double dx = b[0] - a[0], dy = b[1] - a[1];
  const double len2 = dx * dx + dy * dy;
  double t = 0.0;
  if (len2 > 0.0) t = fmin(1.0, fmax(0.0, -(a[0] * dx + a[1] * dy) / len2));
  const double x = a[0] + t * dx, y = a[1] + t * dy;
  return x * x + y * y;
}

// squared distance from the origin to the convex hull of n 2-D points (n <= 32): monotone chain
__device__ __noinline__ double hull_dist2(double (*p)[2], int n) {
  if (n == 1) return p[0][0] * p[0][0] + p[0][1] * p[0][1];
  int idx[32];
  for (int i = 0; i < n; ++i) {  // insertion sort by (x, y)
    int j = i;
    while (j > 0 && (p[idx[j - 1]][0] > p[i][0] || (p[idx[j - 1]][0] == p[i][0] && p[idx[j - 1]][1] > p[i][1]))) {
      idx[j] = idx[j - 1];
      --j;
    }
    idx[j] = i;
  }
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
  const int m = k - 1;  // last equals first
  if (m < 2) return seg_dist2_origin(p[h[0]], p[h[m > 0 ? 1 : 0]]);
  if (m == 2) return seg_dist2_origin(p[h[0]], p[h[1]]);
  bool inside = true;
  double best = INFINITY;
  for (int i = 0; i < m; ++i) {
    const double *a = p[h[i]], *b = p[h[i + 1]];
    if ((b[0] - a[0]) * (0.0 - a[1]) - (b[1] - a[1]) * (0.0 - a[0]) < 0.0) inside = false;
    best = fmin(best, seg_dist2_origin(a, b));
  }
  return inside ? 0.0 : best;
}

// cylinder (radius r, half height hh, axis = body z) against an oriented cube: clip the cube with the
// slab |z| <= hh, project on the xy plane, distance from the axis to the hull of the projection
__device__ __noinline__ bool cylinder_hits_obb(double r, double hh, const Obb &b) {
  double v[8][3];
  for (int s = 0; s < 8; ++s)
    for (int d = 0; d < 3; ++d)
      v[s][d] = b.c[d] + b.e * (((s & 1) ? 1.0 : -1.0) * b.ax[0][d] +
                                (((s & 2) ? 1.0 : -1.0) * b.ax[1][d] + ((s & 4) ? 1.0 : -1.0) * b.ax[2][d]));
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
      const int q = s | bit;  // edge s - q
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
  return hull_dist2(pts, n) <= r * r;
}

// upright robot body at (x, y, 0, yaw) against every occupied voxel cube near it
__device__ __noinline__ bool pose_collides_general(const RobotCtx &cx, float fxf, float fyf, float fyawf) {
  const double fx = (double)fxf, fy = (double)fyf, fyaw = (double)fyawf;
  if (!(fabs(fx) < 1e9 && fabs(fy) < 1e9 && fabs(fyaw) < 1e18)) return false;  // non-finite pose: no contact
  double sb, cb;
  sincos(fyaw, &sb, &cb);
  const double *R = cx.gR;
  const double d[3] = {fx - cx.gt[0], fy - cx.gt[1], 0.0 - cx.gt[2]};
  double cs[3];
  for (int j = 0; j < 3; ++j) cs[j] = R[0 * 3 + j] * d[0] + (R[1 * 3 + j] * d[1] + R[2 * 3 + j] * d[2]);
  double rho, hb[3] = {0.0, 0.0, 0.0};
  if (cx.shape == KC_SPHERE) {
    rho = cx.dim0;
  } else if (cx.shape == KC_CYLINDER) {
    rho = sqrt(cx.dim0 * cx.dim0 + 0.25 * cx.dim1 * cx.dim1);
  } else {
    hb[0] = 0.5 * cx.dim0;
    hb[1] = 0.5 * cx.dim1;
    hb[2] = 0.5 * cx.dim2;
    rho = sqrt(hb[0] * hb[0] + hb[1] * hb[1] + hb[2] * hb[2]);
  }
  const int k0[3] = {cx.g_kx0, cx.g_ky0, cx.g_kz0}, nn[3] = {cx.g_nx, cx.g_ny, cx.g_nz};
  int lo[3], hi[3];
  for (int j = 0; j < 3; ++j) {
    const double a = floor((cs[j] - rho) / cx.res) - 1.0, b = floor((cs[j] + rho) / cx.res) + 1.0;
    lo[j] = (int)fmax(a, (double)k0[j]) - k0[j];
    hi[j] = (int)fmin(b, (double)(k0[j] + nn[j] - 1)) - k0[j];
  }
  if (lo[0] > hi[0] || lo[1] > hi[1] || lo[2] > hi[2]) return false;
  Obb b;
  b.e = 0.5 * cx.res;
  for (int j = 0; j < 3; ++j) {
    b.ax[j][0] = cb * R[0 * 3 + j] + sb * R[1 * 3 + j];
    b.ax[j][1] = cb * R[1 * 3 + j] - sb * R[0 * 3 + j];
    b.ax[j][2] = R[2 * 3 + j];
  }
  for (int lay = lo[2]; lay <= hi[2]; ++lay)
    for (int row = lo[1]; row <= hi[1]; ++row) {
      const uint32_t *wrow = cx.bitmap + ((size_t)lay * cx.g_ny + row) * cx.g_wpr;
      for (int w = lo[0] >> 5; w <= (hi[0] >> 5); ++w) {
        uint32_t bits = __ldg(&wrow[w]);
        if (w == (lo[0] >> 5)) bits &= 0xffffffffu << (lo[0] & 31);
        if (w == (hi[0] >> 5)) bits &= 0xffffffffu >> (31 - (hi[0] & 31));
        while (bits) {
          const int bit = __ffs(bits) - 1;
          bits &= bits - 1;
          const int kx = k0[0] + w * 32 + bit, ky = k0[1] + row, kz = k0[2] + lay;
          const double sc[3] = {((double)kx + 0.5) * cx.res, ((double)ky + 0.5) * cx.res, ((double)kz + 0.5) * cx.res};
          double wv[3];
          for (int i = 0; i < 3; ++i)
            wv[i] = (R[i * 3 + 0] * sc[0] + (R[i * 3 + 1] * sc[1] + R[i * 3 + 2] * sc[2])) + cx.gt[i];
          const double rx = wv[0] - fx, ry = wv[1] - fy;
          b.c[0] = cb * rx + sb * ry;
          b.c[1] = cb * ry - sb * rx;
          b.c[2] = wv[2];
          bool hit;
          if (cx.shape == KC_SPHERE)
            hit = sphere_hits_obb(cx.dim0, b);
          else if (cx.shape == KC_CYLINDER)
            hit = cylinder_hits_obb(cx.dim0, 0.5 * cx.dim1, b);
          else
            hit = box_hits_obb(hb, b);
          if (hit) return true;
        }
      }
    }
  return false;
}

// bmap: the bitmap to read (a CTA's shared-memory copy, or nullptr for the one in global memory).
// dil / sure: k_dilate's maps of it (nullptr: none). A voxel column can only touch the robot's bounding
// circle when it lies within hit_W = floor(R/res) + 2 columns/rows of the pose's own voxel, so a clear
// bit of the footprint-dilated map proves "no collision" without walking the window, and a set bit
// of the sure map proves a collision.
// GENERAL: the kernel was instantiated for tilted sensor frames (its own instantiation, so that the
// planar kernels keep their register budget: the oriented-cube tests need twice the registers)
template <bool GENERAL>
__device__ __forceinline__ bool pose_collides(const RobotCtx &cx, const uint32_t *bmap,
                                              const uint32_t *dil, const uint32_t *sure, float fx, float fy,
                                              float fyaw) {
  if (GENERAL) return pose_collides_general(cx, fx, fy, fyaw);
  if (!bmap) bmap = cx.bitmap;
  const double dx = (double)fx - cx.tx, dy = (double)fy - cx.ty;
  const double pcx = cx.a00 * dx + cx.a10 * dy;  // A^T d : pose in the octree frame
  const double pcy = cx.a01 * dx + cx.a11 * dy;
  // the pose's own voxel only centres the search window, which carries a spare ring and 1e-6-voxel
  // margins for exactly this rounding (fill_collision_frame): a product instead of the quotient
  const double qx = pcx * cx.res_factor, qy = pcy * cx.res_factor;
  const double fkx = floor(qx), fky = floor(qy);
  if (!(fabs(fkx) < 1e9 && fabs(fky) < 1e9)) return false;  // non-finite pose: FCL reports no contact
  const int kcx = (int)fkx, kcy = (int)fky;                 // the pose's own voxel column
  const int ccol = kcx - cx.bm_kx0, crow = kcy - cx.bm_ky0;
  const bool inside = ccol >= 0 && ccol < cx.bm_cols && crow >= 0 && crow < cx.bm_rows;
  if (dil && inside) {
    const int wi = crow * cx.bm_wpr + (ccol >> 5);
    if (!((dil[wi] >> (ccol & 31)) & 1u)) return false;
    // an occupied column inside the inscribed circle for every position within this voxel (margins as
    // in the FP32 filter below): the exact test can only say "hit"
    if ((sure[wi] >> (ccol & 31)) & 1u) return true;
  }
  double cth = 1.0, sth = 0.0;
  if (cx.shape == KC_BOX) {
    sincos((double)fyaw - cx.psi, &sth, &cth);
    sth = cx.sigma * sth;  // heading in the octree frame: A^T u = (cos, sigma sin)
  }
  const int Wh = cx.hit_W;
  if (cx.use_rowmask) {
    // pose inside its own voxel, in voxel units (FP32 is only a filter: +-1e-4 relative margins)
    const float ux = (float)(qx - fkx), uy = (float)(qy - fky);
    const float rho2 = cx.rho * cx.rho;
    const int c0 = ccol - Wh;  // window column of mask bit 0 (may lie outside the bitmap)
    const int w0 = c0 >> 5, sh = c0 & 31, wpr = cx.bm_wpr;
    // window rows four at a time: their bitmap words and row masks are independent loads, issued
    // together (one round trip per group instead of two or three dependent ones per row)
    for (int d0 = -Wh; d0 <= Wh; d0 += 4) {
      uint32_t rb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int dyi = d0 + j, row = crow + dyi;
        uint32_t bb = 0u;
        if (dyi <= Wh && row >= 0 && row < cx.bm_rows) {
          const uint32_t *wrow = bmap + row * wpr;
          const uint32_t lo = (w0 >= 0 && w0 < wpr) ? wrow[w0] : 0u;
          const uint32_t hi = (w0 + 1 >= 0 && w0 + 1 < wpr) ? wrow[w0 + 1] : 0u;
          bb = __funnelshift_r(lo, hi, sh) & cx.rowmask[abs(dyi)];
        }
        rb[j] = bb;
      }
      if (!(rb[0] | rb[1] | rb[2] | rb[3])) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t bits = rb[j];
        if (!bits) continue;
        const int dyi = d0 + j, row = crow + dyi;
        const float gy = fmaxf(fmaxf((float)dyi - uy, 0.0f), uy - (float)(dyi + 1));
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const int dxi = b - Wh;
          const float gx = fmaxf(fmaxf((float)dxi - ux, 0.0f), ux - (float)(dxi + 1));
          const float g2 = gx * gx + gy * gy;
          if (g2 > rho2 * 1.0002f + 1e-6f) continue;  // clear of the bounding circle
          const int col = c0 + b;
          if (cx.shape == KC_CYLINDER && g2 < rho2 * 0.9998f - 1e-6f) return true;  // well inside
          float dz2 = 0.0f;
          if (cx.shape == KC_SPHERE)
            dz2 = __uint_as_float(__ldg(&cx.sph_col[(size_t)row * cx.bm_cols + col]));
          if (column_hit(cx, cx.bm_kx0 + col, cx.bm_ky0 + row, dz2, pcx, pcy, cth, sth)) return true;
        }
      }
    }
    return false;
  }
  const int kx0 = max(kcx - Wh, cx.bm_kx0), kx1 = min(kcx + Wh, cx.bm_kx0 + cx.bm_cols - 1);
  const int ky0 = max(kcy - Wh, cx.bm_ky0), ky1 = min(kcy + Wh, cx.bm_ky0 + cx.bm_rows - 1);
  if (kx0 > kx1) return false;
  const int c0 = kx0 - cx.bm_kx0, c1 = kx1 - cx.bm_kx0;
  for (int ky = ky0; ky <= ky1; ++ky) {
    const int row = ky - cx.bm_ky0;
    const uint32_t *wrow = bmap + (size_t)row * cx.bm_wpr;
    for (int w = c0 >> 5; w <= (c1 >> 5); ++w) {
      uint32_t bits = wrow[w];
      if (w == (c0 >> 5)) bits &= 0xffffffffu << (c0 & 31);
      if (w == (c1 >> 5)) bits &= 0xffffffffu >> (31 - (c1 & 31));
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const int col = w * 32 + b;
        float dz2 = 0.0f;
        if (cx.shape == KC_SPHERE)
          dz2 = __uint_as_float(__ldg(&cx.sph_col[(size_t)row * cx.bm_cols + col]));
        if (column_hit(cx, cx.bm_kx0 + col, ky, dz2, pcx, pcy, cth, sth)) return true;
      }
    }
  }
  return false;
}

// k_dilate: once per cycle and robot, the three derived maps of the voxel-column bitmap that the pose
// test of k_rollout_collide consults before it walks a window (RobotCtx::dil_maps). One thread per
// bitmap word; the maps are a few KB and stay in L1/L2 for the rollout CTAs.
__device__ __forceinline__ uint32_t dilate_word_cols(uint32_t cur, uint32_t prev, uint32_t next, int W) {
  uint32_t a = cur;
  for (int sft = 1; sft <= W; ++sft) a |= (cur << sft) | (prev >> (32 - sft)) | (cur >> sft) | (next << (32 - sft));
  return a;
}
__global__ void __launch_bounds__(256) k_dilate(const RobotCtx *__restrict__ ctxs) {
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(6);
  grid_dep_launch();  // k_rollout_collide's prologue (slot decode, heading rows) may run beside this grid
  const int W = cx.dil_W;
  if (W <= 0 || !cx.coll_enabled) return;
  const int wpr = cx.bm_wpr, rows = cx.bm_rows, words = rows * wpr;
  const uint32_t *bm = cx.bitmap;
  uint32_t *dil = cx.dil_maps, *sure = dil + cx.dil_stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) {
    const int row = i / wpr, w = i - row * wpr;
    uint32_t d = 0u, sr = 0u;
    for (int dy = -W; dy <= W; ++dy) {
      const int r = row + dy;
      if (r < 0 || r >= rows) continue;
      const uint32_t rm = cx.use_rowmask ? cx.rowmask[abs(dy)] : 0xffffffffu;
      const uint32_t sm = cx.use_rowmask ? cx.suremask[abs(dy)] : 0u;
      if (!(rm | sm)) continue;
      const uint32_t cur = bm[r * wpr + w];
      const uint32_t prev = (w > 0) ? bm[r * wpr + w - 1] : 0u;
      const uint32_t next = (w + 1 < wpr) ? bm[r * wpr + w + 1] : 0u;
      if (!(cur | prev | next)) continue;
      if (!cx.use_rowmask) {  // square footprint
        d |= dilate_word_cols(cur, prev, next, W);
        continue;
      }
      // mask bit (dx + W) set <=> the column at offset dx matters to a pose in this column, i.e. this
      // column is marked when the bitmap has a bit at offset dx
      for (int dxo = 1; dxo <= W; ++dxo) {
        const uint32_t right = (cur >> dxo) | (next << (32 - dxo)), left = (cur << dxo) | (prev >> (32 - dxo));
        if ((rm >> (W + dxo)) & 1u) d |= right;
        if ((rm >> (W - dxo)) & 1u) d |= left;
        if ((sm >> (W + dxo)) & 1u) sr |= right;
        if ((sm >> (W - dxo)) & 1u) sr |= left;
      }
      if ((rm >> W) & 1u) d |= cur;
      if ((sm >> W) & 1u) sr |= cur;
    }
    dil[i] = d;
    sure[i] = sr;
  }
}

// index of the first loop iteration i (pose index i+1) that collides, or P-1 if none
template <bool GENERAL>
__device__ __forceinline__ int warp_first_collision(const RobotCtx &cx, const uint32_t *bmap,
                                                    const uint32_t *dil, const uint32_t *sure, const float *sx,
                                                    const float *sy, const float *syaw, int lane) {
  const int P = cx.P;
  if (!cx.coll_enabled) return P - 1;
  for (int base = 0; base < P - 1; base += 32) {
    const int i = base + lane;
    bool hit = false;
    if (i < P - 1) hit = pose_collides<GENERAL>(cx, bmap, dil, sure, sx[i + 1], sy[i + 1], syaw ? syaw[i + 1] : 0.0f);
    const unsigned m = __ballot_sync(FULL, hit);
    if (m) return base + __ffs(m) - 1;
  }
  return P - 1;
}

// Eigen (p1 - p2).squaredNorm() on Vector3f with z = 0: dx*dx + (dy*dy + 0). The "+ 0" is dropped:
// a square is never -0, so x + 0.0f == x bit for bit (NaN stays NaN)
__device__ __forceinline__ float sq_dist(float ax, float ay, float bx, float by) {
  const float dx = ax - bx, dy = ay - by;
  return dx * dx + dy * dy;
}

// ref: cost_evaluator.cpp:150-177 goalCostFunc
__device__ __forceinline__ float warp_goal_cost(const RobotCtx &cx, const float *segX,
                                                const float *segY, float ex, float ey, int lane) {
  float best = FLT_MAX;
  int bidx = 0x7fffffff;
  for (int j = lane; j < cx.seg_count; j += 32) {
    const float d = sq_dist(ex, ey, segX[j], segY[j]);
    if (d < best) {
      best = d;
      bidx = j;
    }
  }
  warp_argmin_f(best, bidx);
  if (bidx == 0x7fffffff) bidx = 0;  // nothing below FLT_MAX: closest_local_idx stays 0
  const int abs_idx = bidx + cx.seg_start;
  const float at = (abs_idx >= cx.path_n) ? 0.0f : __ldg(&cx.pathAcc[abs_idx]);
  const float arc = (cx.path_len - at) / cx.path_len;
  return arc + (sqrtf(best) / cx.path_len);
}

// ref: cost_evaluator.cpp:111-141 pathCostFunc. pmin: [P] warp scratch.
// cell of a query point inside the query window (where k_cell_cand has filled cell_info)
__device__ __forceinline__ bool query_cell(const RobotCtx &cx, float px, float py, int &cell) {
  const float u = (px - cx.gx0) * cx.inv_h, v = (py - cx.gy0) * cx.inv_h;
  if (!(u >= 0.0f && u < (float)kGridN && v >= 0.0f && v < (float)kGridN)) return false;
  const int ix = (int)u, iy = (int)v;
  if (ix < cx.q_x0 || ix > cx.q_x1 || iy < cx.q_y0 || iy > cx.q_y1) return false;
  cell = iy * kGridN + ix;
  return true;
}

// Half-width of a query point's distance bracket around its cell's centre distance dmin: the distance
// to a point set is 1-Lipschitz, so |d(q) - d(c)| <= |q - c| - the point's ACTUAL offset from the
// centre (0.38 h on average), not the worst case h / sqrt2 of the cell's corner. The brackets of the
// bound stage and the filters of the exact stage are tighter by that much: fewer slots survive the
// bounds and fewer points of a survivor need an exact search (a trajectory that runs along a dense
// cluster had half of its points inside the worst-case bracket of its minimum).
__device__ __forceinline__ float bracket_halfwidth(const RobotCtx &cx, float px, float py, int cell) {
  const int ix = cell & (kGridN - 1), iy = cell / kGridN;
  const float cxm = cx.gx0 + ((float)ix + 0.5f) * cx.h, cym = cx.gy0 + ((float)iy + 0.5f) * cx.h;  // as k_cell_cand forms it
  const float dx = px - cxm, dy = py - cym;
  // (rounded up with slack; never wider than the corner bound)
  return fminf(sqrtf(dx * dx + dy * dy) * 1.002f + 1e-6f * cx.h, 0.7072f * cx.h * 1.002f);
}

// Exact two-level search over the tracked segment: consecutive segment points are at most seg_step
// apart, so |p - seg_j| >= |p - seg_c| - (kPathWin/2) seg_step for every j of the kPathWin-point
// window around its centre c. Windows whose centre is farther than (best centre distance +
// (kPathWin/2) seg_step) cannot hold the minimum and are skipped; the others are evaluated
// exhaustively with the same float operations, so the min is the reference's min.
constexpr int kPathWin = 16;
__device__ __forceinline__ float path_point_min(const RobotCtx &cx, const float *segX,
                                                const float *segY, float px, float py) {
  const int S = cx.seg_count;
  const int nwin = (S + kPathWin - 1) / kPathWin;
  float mc = FLT_MAX;
  for (int k = 0; k < nwin; ++k) {
    const int c = min(kPathWin * k + kPathWin / 2, S - 1);
    mc = fminf(mc, sq_dist(segX[c], segY[c], px, py));
  }
  const float rad = sqrtf(mc) * 1.000001f + (float)(kPathWin / 2) * cx.seg_step;
  const float thr = rad * rad * 1.00001f;  // inf when mc == FLT_MAX: every window is scanned
  float m = FLT_MAX;
  for (int k0 = 0; k0 < nwin; k0 += 32) {
    // windows to scan as a bit mask: the scan below then runs the same instructions in every lane
    // (different windows = different addresses, not different control flow)
    unsigned todo = 0u;
    const int k1 = min(32, nwin - k0);
    for (int k = 0; k < k1; ++k) {
      const int c = min(kPathWin * (k0 + k) + kPathWin / 2, S - 1);
      const float dc = sq_dist(segX[c], segY[c], px, py);
      if (!(dc > thr)) todo |= 1u << k;
    }
    while (todo) {
      const int k = k0 + __ffs(todo) - 1;
      todo &= todo - 1;
      const int j0 = kPathWin * k, j1 = min(j0 + kPathWin, S);
      if (j1 - j0 == kPathWin) {
#pragma unroll
        for (int j = 0; j < kPathWin; ++j) m = fminf(m, sq_dist(segX[j0 + j], segY[j0 + j], px, py));
      } else {
        for (int j = j0; j < j1; ++j) m = fminf(m, sq_dist(segX[j], segY[j], px, py));
      }
    }
  }
  return sqrtf(m);  // min of sqrt == sqrt of min (monotone rounding); FLT_MAX stays finite
}

__device__ __forceinline__ float warp_path_cost(const RobotCtx &cx, const float *segX,
                                                const float *segY, const float *sx, const float *sy,
                                                float *pmin, int lane) {
  const int P = cx.P, S = cx.seg_count;
  for (int i = lane; i < P; i += 32) {
    const float px = sx[i], py = sy[i];
    int cell;
    int2 ci = make_int2(0, -1);
    if (cx.pcand_enabled && query_cell(cx, px, py, cell)) ci = __ldg(&cx.pcell_info[cell]);
    if (ci.y > 0) {  // the cell's candidate list holds every point that can attain the minimum
      const float2 *cand = cx.pcand_pool + ci.x;
      float m = FLT_MAX;
      for (int q = 0; q < ci.y; ++q) {
        const float2 o = __ldg(&cand[q]);
        m = fminf(m, sq_dist(o.x, o.y, px, py));
      }
      pmin[i] = sqrtf(m);
    } else {
      pmin[i] = path_point_min(cx, segX, segY, px, py);
    }
  }
  __syncwarp();
  // index-ordered float sum (the reference's accumulation order; every lane forms it). The cost kernels
  // keep pmin 16-byte aligned: four addends per shared-memory load
  float total = 0.0f;
#ifndef KC_SCALAR_SUM
#define KC_SCALAR_SUM 0
#endif
  if (!KC_SCALAR_SUM && (reinterpret_cast<uintptr_t>(pmin) & 15u) == 0) {
    const float4 *p4 = reinterpret_cast<const float4 *>(pmin);
    const int n4 = P >> 2;
    for (int i = 0; i < n4; ++i) {
      const float4 q = p4[i];
      total += q.x;
      total += q.y;
      total += q.z;
      total += q.w;
    }
    for (int i = n4 << 2; i < P; ++i) total += pmin[i];
  } else {
    for (int i = 0; i < P; ++i) total += pmin[i];
  }
  __syncwarp();
  const float end_err = sqrtf(sq_dist(sx[P - 1], sy[P - 1], segX[S - 1], segY[S - 1])) / cx.seg_len;
  return (total / (float)P + end_err) / 2;
}

// exact min over (trajectory point, obstacle) of d^2 (double, single rounding of dx^2+dy^2 as in
// trajectory.h:228); returns >= dcap2 when nothing is closer than the cost cut-off distance.
//
// Work reduction, all of it exact:
//  * the cell table gives every query point a lower bound LB = (sqrt(nn) - sqrt2) h and an upper
//    bound UB = (sqrt(nn) + sqrt2) h on its nearest-obstacle distance; the trajectory's answer is
//    <= min UB, so only points with LB below that (shrinking) bound are searched at all;
//  * a searched point walks grid rows outwards from its own row inside the current best radius,
//    skipping empty cells through the occupancy bitmask;
//  * pairs are filtered in FP32 (conservatively) and only near-minimal ones re-evaluated in FP64.
__device__ __forceinline__ float conservative_f(double best) {
  return __double2float_ru(best) * 1.000001f;
}

// Generic exact search (fallback): one lane = one query point (has == false: idle lane). Walks grid
// rows outwards from the point's own row inside the current best radius, skipping empty cells
// through the occupancy bitmask. `best` must be warp-uniform on entry; returns the warp-wide min.
__device__ __forceinline__ double nn_search_batch(const RobotCtx &cx, float px, float py, bool has,
                                                  double best) {
  const float h = cx.h;
  const float u = (px - cx.gx0) * cx.inv_h, v = (py - cx.gy0) * cx.inv_h;
  const int ccx = min(max((int)u, 0), kGridN - 1), ccy = min(max((int)v, 0), kGridN - 1);
  double lb2 = 0.0;  // squared lower bound of this point's nearest-obstacle distance
  if (has && u >= 0.0f && u < (float)kGridN && v >= 0.0f && v < (float)kGridN &&
      ccx >= cx.q_x0 && ccx <= cx.q_x1 && ccy >= cx.q_y0 && ccy <= cx.q_y1) {
    const unsigned nn = __ldg(&cx.cell_nn[ccy * kGridN + ccx]);
    if (nn == 0xFFFFu) {
      has = false;  // no obstacle point was binned at all
    } else {
      const float lb = fmaxf(0.0f, sqrtf((float)nn) - 1.45f) * h * 0.999f;
      lb2 = (double)lb * (double)lb;
    }
  }
  float bestf = conservative_f(best);
  for (int r = 0; r < kGridN; ++r) {
    const float lby = fmaxf(0.0f, (float)r - 1.02f) * h;
    const double lby2 = (double)lby * (double)lby;
    const bool active = has && (lby2 < best) && (lb2 < best);
    if (!__any_sync(FULL, active)) break;
    if (active) {
      const float rem = sqrtf((float)(best - lby2)) * 1.0001f;
      const int cm = (int)(rem * cx.inv_h) + 3;
      const int x0 = max(0, ccx - cm), x1 = min(kGridN - 1, ccx + cm);
      for (int sgn = 0; sgn < (r == 0 ? 1 : 2); ++sgn) {
        const int iy = sgn ? ccy - r : ccy + r;
        if (iy < 0 || iy >= kGridN) continue;
        for (int w = x0 >> 5; w <= (x1 >> 5); ++w) {
          uint32_t bits = __ldg(&cx.occ[iy * kGridWords + w]);
          if (w == (x0 >> 5)) bits &= 0xffffffffu << (x0 & 31);
          if (w == (x1 >> 5)) bits &= 0xffffffffu >> (31 - (x1 & 31));
          while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int cell = iy * kGridN + w * 32 + b;
            const int s = __ldg(&cx.cell_start[cell]), e = __ldg(&cx.cell_start[cell + 1]);
            for (int q = s; q < e; ++q) {
              const float2 o = __ldg(&cx.sorted_xy[q]);
              const float dx = o.x - px, dy = o.y - py;
              const float d2f = __fmaf_rn(dx, dx, dy * dy);  // filter only (<= 2 ulp off)
              if (d2f <= bestf) {
                const double d2 = (double)dx * (double)dx + (double)dy * (double)dy;
                if (d2 < best) {
                  best = d2;
                  bestf = conservative_f(best);
                }
              }
            }
          }
        }
      }
    }
    best = warp_min_d(best);
    bestf = conservative_f(best);
  }
  return best;
}

// Exact nearest-obstacle search of ONE in-window query point by the whole warp (px, py and `best`
// warp-uniform): the rows of the disc of radius sqrt(best) around the query are contiguous runs of
// the cell-sorted point array, so lanes first fetch the runs' bounds (one row each), short runs are
// walked by their owner lane and long ones by all 32 lanes with four independent loads in flight.
// Same pairs within the radius, same arithmetic, same min as the reference loop. Used for cells that
// carry no candidate list (dense neighbourhoods, cells outside the reach mask, pool overflow).
#ifndef KC_NN_UNROLL
#define KC_NN_UNROLL 4
#endif
constexpr int kNnUnroll = KC_NN_UNROLL;
__device__ __noinline__ double warp_nn_search_one(const RobotCtx &cx, float px, float py, double best,
                                                     int lane) {
  const float h = cx.h;
  const float fv = (py - cx.gy0) * cx.inv_h;
  const int ccy = min(max((int)fv, 0), kGridN - 1);
  const float rad = __fsqrt_ru(__double2float_ru(best)) * 1.0001f;
  if (!(rad < 1e30f)) return best;
  const int rc = (int)fminf(rad * cx.inv_h + 2.0f, (float)kGridN);
  float bestf = conservative_f(best);
  double mine = best;
  auto look = [&](float2 o) {
    const float dx = o.x - px, dy = o.y - py;
    const float d2f = __fmaf_rn(dx, dx, dy * dy);  // filter only (<= 2 ulp off)
    if (d2f <= bestf) {
      const double d2 = (double)dx * (double)dx + (double)dy * (double)dy;
      if (d2 < mine) {
        mine = d2;
        bestf = conservative_f(mine);
      }
    }
  };
  for (int iy0 = ccy - rc; iy0 <= ccy + rc; iy0 += 32) {
    const int iy = iy0 + lane;
    int s = 0, e = 0;
    if (iy <= ccy + rc && iy >= 0 && iy < kGridN) {
      // vertical gap between the query and the row's band, with 2 % of a cell as binning slack
      const float ylo = cx.gy0 + (float)iy * h;
      const float gap = fmaxf(fmaxf(ylo - py, py - (ylo + h)) - 0.02f * h, 0.0f);
      if (gap < rad) {
        const float half = sqrtf(rad * rad - gap * gap) * 1.0001f;
        const float f0 = (px - half - cx.gx0) * cx.inv_h - 0.05f, f1 = (px + half - cx.gx0) * cx.inv_h + 0.05f;
        if (f1 >= 0.0f && f0 < (float)kGridN) {
          const int x0 = max(0, (int)floorf(f0)), x1 = min(kGridN - 1, (int)floorf(f1));
          s = __ldg(&cx.cell_start[iy * kGridN + x0]);
          e = __ldg(&cx.cell_start[iy * kGridN + x1 + 1]);
        }
      }
    }
    unsigned heavy = __ballot_sync(FULL, e - s > 4);
    if (!((heavy >> lane) & 1u))
      for (int q = s; q < e; ++q) look(__ldg(&cx.sorted_xy[q]));
    while (heavy) {
      const int src = __ffs(heavy) - 1;
      heavy &= heavy - 1;
      const int sb = __shfl_sync(FULL, s, src), eb = __shfl_sync(FULL, e, src);
      for (int q0 = sb + lane; q0 < eb; q0 += 32 * kNnUnroll) {  // kNnUnroll independent loads in flight per lane
        float2 o[kNnUnroll];
#pragma unroll
        for (int u = 0; u < kNnUnroll; ++u)
          if (q0 + 32 * u < eb) o[u] = __ldg(&cx.sorted_xy[q0 + 32 * u]);
#pragma unroll
        for (int u = 0; u < kNnUnroll; ++u)
          if (q0 + 32 * u < eb) look(o[u]);
      }
    }
  }
  return warp_min_d(mine);
}

// Trajectory-wide exact min d^2. Every query-window cell carries (k_cell_cand) the distance dmin from
// its centre to the nearest obstacle point and the list of points that can be the nearest one of
// ANY query inside the cell (all points within dmin + sqrt2*h of the centre). A query point
// therefore has its answer bracketed by dmin -/+ h/sqrt2: the trajectory's result is <= the
// smallest upper bracket, only points whose lower bracket is below the (shrinking) best are
// evaluated, most promising first, and those only against their cell's candidate list, which the
// 32 lanes split: same pairs, same arithmetic, same min as the reference's N*P*M loop.
__device__ __forceinline__ double warp_min_obstacle_d2(const RobotCtx &cx, const float *sx,
                                                       const float *sy, int lane) {
  const int P = cx.P;
  double best = cx.dcap2;
  for (int k = lane; k < P; k += 32) {
    int cell;
    if (query_cell(cx, sx[k], sy[k], cell)) {
      const float dm = __int_as_float(__ldg(&cx.cell_info[cell].x));
      const float ub = dm * 1.001f + bracket_halfwidth(cx, sx[k], sy[k], cell);
      if (ub < FLT_MAX) best = fmin(best, (double)ub * (double)ub);
    }
  }
  best = warp_min_d(best);
  for (int base = 0; base < P; base += 32) {
    const int k = base + lane;
    const bool has = k < P;
    const float px = has ? sx[k] : 0.0f, py = has ? sy[k] : 0.0f;
    bool fallback = false, active = false;
    int4 ci = make_int4(0, 0, 0, 0);
    float lb = 0.0f;
    if (has) {
      int cell;
      if (query_cell(cx, px, py, cell)) {
        ci = __ldg(&cx.cell_info[cell]);
        lb = fmaxf(0.0f, __int_as_float(ci.x) * 0.999f - bracket_halfwidth(cx, px, py, cell));
        active = true;
      } else {
        fallback = true;  // outside the prepared window (caller-provided rows, non-finite poses)
      }
    }
    const double lb2 = (double)lb * (double)lb;
    // (1) the most promising point (smallest lower bracket), its list split over the 32 lanes:
    //     yields a near-final radius. (2) remaining points with short lists: one lane each, in
    //     parallel. (3) remaining points with long lists: one at a time, split over the lanes.
    for (int phase = 0; phase < 3; ++phase) {
      if (phase == 1) {
        // (cells without a list, ci.z < 0, wait for phase 2: the whole warp searches them)
        const bool mineActive = active && (lb2 < best) && ci.z >= 0 && ci.z <= kCandSerial;
        if (__any_sync(FULL, mineActive)) {
          double mine = best;
          if (mineActive) {
            active = false;
            float bestf = conservative_f(best);
            const float2 *cand = cx.cand_pool + ci.y;
            for (int q = 0; q < ci.z; ++q) {
              const float2 o = __ldg(&cand[q]);
              const float dx = o.x - px, dy = o.y - py;
              const float d2f = __fmaf_rn(dx, dx, dy * dy);  // filter only (<= 2 ulp off)
              if (d2f <= bestf) {
                const double d2 = (double)dx * (double)dx + (double)dy * (double)dy;
                if (d2 < mine) {
                  mine = d2;
                  bestf = conservative_f(mine);
                }
              }
            }
          }
          best = warp_min_d(mine);
        }
        continue;
      }
      for (;;) {
        active = active && (lb2 < best);
        const unsigned m = __ballot_sync(FULL, active);
        if (!m) break;
        int src;
        if (phase == 0) {
          float key = active ? lb : FLT_MAX;
          src = lane;
          warp_argmin_f(key, src);
        } else {
          src = __ffs(m) - 1;
        }
        const float qx = __shfl_sync(FULL, px, src), qy = __shfl_sync(FULL, py, src);
        const int start = __shfl_sync(FULL, ci.y, src), cnt = __shfl_sync(FULL, ci.z, src);
        if (lane == src) active = false;
        if (cnt < 0) {  // no candidate list for this cell: exact search of the point's own disc
          best = warp_nn_search_one(cx, qx, qy, best, lane);
        } else if (cnt > 0) {
          const float bestf = conservative_f(best);
          const float2 *cand = cx.cand_pool + start;
          double mine = best;
          for (int q = lane; q < cnt; q += 32) {
            const float2 o = __ldg(&cand[q]);
            const float dx = o.x - qx, dy = o.y - qy;
            const float d2f = __fmaf_rn(dx, dx, dy * dy);  // filter only (<= 2 ulp off)
            if (d2f <= bestf) {
              const double d2 = (double)dx * (double)dx + (double)dy * (double)dy;
              mine = fmin(mine, d2);
            }
          }
          best = warp_min_d(mine);
        }
        if (phase == 0) break;
      }
    }
    if (__any_sync(FULL, fallback)) best = nn_search_batch(cx, px, py, fallback, best);
  }
  return best;
}

// ref: cost_evaluator.cpp:187-206 / 209-233. V(c, j) -> float velocity component c at index j.
template <class V>
__device__ __forceinline__ float warp_smoothness(V vel, int nv, float a0, float a1, float a2,
                                                 int lane) {
  float cost = 0.0f;
  for (int base = 1; base < nv; base += 32) {
    const int i = base + lane;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    if (i < nv) {
      if (a0 > 0) {
        const float d = vel(0, i) - vel(0, i - 1);
        t0 = (double)d * (double)d / (double)a0;
      }
      if (a1 > 0) {
        const float d = vel(1, i) - vel(1, i - 1);
        t1 = (double)d * (double)d / (double)a1;
      }
      if (a2 > 0) {
        const float d = vel(2, i) - vel(2, i - 1);
        t2 = (double)d * (double)d / (double)a2;
      }
    }
    unsigned nz = __ballot_sync(FULL, (t0 != 0.0) || (t1 != 0.0) || (t2 != 0.0));
    while (nz) {  // adding an exact zero never changes the float accumulator: skip those
      const int j = __ffs(nz) - 1;
      nz &= nz - 1;
      cost = (float)((double)cost + shfl_d(t0, j));
      cost = (float)((double)cost + shfl_d(t1, j));
      cost = (float)((double)cost + shfl_d(t2, j));
    }
  }
  return cost / (float)(3 * (long long)nv);
}

template <class V>
__device__ __forceinline__ float warp_jerk(V vel, int nv, float a0, float a1, float a2, int lane) {
  float cost = 0.0f;
  for (int base = 2; base < nv; base += 32) {
    const int i = base + lane;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    if (i < nv) {
      if (a0 > 0) {
        const float j = vel(0, i) - 2 * vel(0, i - 1) + vel(0, i - 2);
        t0 = (double)j * (double)j / (double)a0;
      }
      if (a1 > 0) {
        const float j = vel(1, i) - 2 * vel(1, i - 1) + vel(1, i - 2);
        t1 = (double)j * (double)j / (double)a1;
      }
      if (a2 > 0) {
        const float j = vel(2, i) - 2 * vel(2, i - 1) + vel(2, i - 2);
        t2 = (double)j * (double)j / (double)a2;
      }
    }
    unsigned nz = __ballot_sync(FULL, (t0 != 0.0) || (t1 != 0.0) || (t2 != 0.0));
    while (nz) {
      const int j = __ffs(nz) - 1;
      nz &= nz - 1;
      cost = (float)((double)cost + shfl_d(t0, j));
      cost = (float)((double)cost + shfl_d(t1, j));
      cost = (float)((double)cost + shfl_d(t2, j));
    }
  }
  return cost / (float)(3 * (long long)nv);
}

// weighted total in the reference's term order (cost_evaluator.cpp:52-100):
// float += double * float, one term at a time. Part 1: goal + reference path.
__device__ __forceinline__ float warp_partial_cost(const RobotCtx &cx, const float *segX,
                                                   const float *segY, const float *sx,
                                                   const float *sy, float *pmin, int lane) {
  const int P = cx.P;
  float total = 0.0f;
  if (cx.path_enabled) {
    if (cx.w_goal > 0.0) {
      const float g = warp_goal_cost(cx, segX, segY, sx[P - 1], sy[P - 1], lane);
      total = (float)((double)total + cx.w_goal * (double)g);
    }
    if (cx.w_path > 0.0) {
      const float c = warp_path_cost(cx, segX, segY, sx, sy, pmin, lane);
      total = (float)((double)total + cx.w_path * (double)c);
    }
  }
  return total;
}

// Part 2, continuing from `total`: obstacles, smoothness, jerk.
template <class V>
__device__ __forceinline__ float warp_finish_cost(const RobotCtx &cx, float total, const float *sx,
                                                  const float *sy, V vel, int lane,
                                                  bool constant_velocity) {
  const int P = cx.P;
  if (cx.obs_enabled) {
    const double d2 = warp_min_obstacle_d2(cx, sx, sy, lane);
    if (d2 < cx.dcap2) {  // otherwise dist >= D and the term is an exact zero
      const float md = (float)d2;
      const float dist = (float)sqrt((double)md);
      const float c = fmaxf(cx.D - dist, 0.0f) / cx.D;
      total = (float)((double)total + cx.w_obs * (double)c);
    }
  }
  // a constant velocity row has all-zero differences: both terms are exact zeros, and adding
  // w * 0.0 leaves the float accumulator unchanged
  if (constant_velocity) return total;
  if (cx.w_smooth > 0.0) {
    const float c = warp_smoothness(vel, P - 1, cx.acc0, cx.acc1, cx.acc2, lane);
    total = (float)((double)total + cx.w_smooth * (double)c);
  }
  if (cx.w_jerk > 0.0) {
    const float c = warp_jerk(vel, P - 1, cx.acc0, cx.acc1, cx.acc2, lane);
    total = (float)((double)total + cx.w_jerk * (double)c);
  }
  return total;
}

template <class V>
__device__ __forceinline__ float warp_total_cost(const RobotCtx &cx, const float *segX,
                                                 const float *segY, const float *sx,
                                                 const float *sy, float *pmin, V vel, int lane,
                                                 bool constant_velocity = false) {
  const float partial = warp_partial_cost(cx, segX, segY, sx, sy, pmin, lane);
  return warp_finish_cost(cx, partial, sx, sy, vel, lane, constant_velocity);
}

// Bracket of the trajectory's nearest-obstacle distance from the per-cell table alone (no list is
// walked): a query point of a cell whose centre has its nearest point at dmin has its own at
// dmin -/+ h/sqrt2 (distance to a set is 1-Lipschitz). Points without a table entry (outside the
// prepared window, cells outside the reach mask) leave the lower end at 0.
__device__ __forceinline__ void warp_obstacle_bracket(const RobotCtx &cx, const float *sx,
                                                      const float *sy, int lane, float &lo, float &hi) {
  const int P = cx.P;
  float l = INFINITY, u = INFINITY;
  for (int k = lane; k < P; k += 32) {
    float lk = 0.0f, uk = INFINITY;
    int cell;
    if (query_cell(cx, sx[k], sy[k], cell)) {
      const float dm = __int_as_float(__ldg(&cx.cell_info[cell].x));
      if (dm == dm) {  // NaN: no bracket for this cell
        const float hw = bracket_halfwidth(cx, sx[k], sy[k], cell);
        lk = fmaxf(0.0f, dm * 0.999f - hw);
        uk = dm * 1.001f + hw;
      }
    }
    l = fminf(l, lk);
    u = fminf(u, uk);
  }
  lo = warp_min_f(l);
  hi = warp_min_f(u);
}

// shared-memory layout of k_eval_rows:
//   per warp: acc[64] (double) | segX[S] segY[S] | dilation tmp[DW] dil[DW] |
//   per warp: sx[P] sy[P] syaw[P] pmin[P]
__host__ __device__ inline size_t eval_smem_bytes(int P, int S, int warps, int dil_words) {
  return sizeof(double) * 64 * (size_t)warps +
         sizeof(float) * ((size_t)2 * S + (size_t)2 * dil_words + (size_t)warps * 4 * P);
}

// rollout + collision (+ padding) of one slot; returns admissible flag and the velocity cut
// (velocities are `v` for j < cut and 0 for j >= cut; cut == P-1 when not padded)
// ref: trajectory_sampler.cpp:122-125: a sample whose three components are all ~0 is rejected
__device__ __forceinline__ bool slot_moves(const SlotVel &v) {
  return !(fabs(v.vx) < kMinVel && fabs(v.vy) < kMinVel && fabs(v.om) < kMinVel);
}

// collision test + padding of one rolled-out slot (sx / sy hold its P poses); returns the admissible
// flag and the velocity cut (velocities are `v` for j < cut and 0 beyond; cut == P-1: not padded)
template <bool GENERAL>
__device__ __forceinline__ bool warp_collide_slot(const RobotCtx &cx, const uint32_t *bmap,
                                                  const uint32_t *dil, const uint32_t *sure, float *sx,
                                                  float *sy, const float *syaw, int lane, int &cut) {
  const int P = cx.P;
  cut = P - 1;
  const int i = warp_first_collision<GENERAL>(cx, bmap, dil, sure, sx, sy, syaw, lane);
  if (i >= P - 1) return true;  // no collision
  // ref: trajectory_sampler.cpp:147-168
  const long long last_free = (i > 0) ? (i - 1) : (P - 1);
  if (!cx.drop_samples && last_free > cx.num_ctrl_points && last_free < P - 1) {
    const float lx = sx[last_free], ly = sy[last_free];
    __syncwarp();
    for (int j = (int)last_free + 1 + lane; j < P - 1; j += 32) {
      sx[j + 1] = lx;
      sy[j + 1] = ly;
    }
    __syncwarp();
    cut = (int)last_free + 1;
    return true;
  }
  return false;
}

// ================================================================================================
// The cycle's two trajectory kernels, warp per velocity slot:
//   k_rollout_collide   rollout + per-pose collision (+ padding). Needs only the voxel bitmap, so
//                       in the captured graph it runs BESIDE the obstacle-grid preparation
//                       (scan / scatter / candidate lists). Admissible slots store their path row
//                       and velocity cut and append themselves to an (unordered) list.
//                       STORE_VEL: also materialise the velocity rows (generateTrajectories).
//   k_cost_eval         one warp per admissible slot: the five cost terms from the stored row,
//                       packed 64-bit argmin; the last CTA publishes the winner (its row is already
//                       in memory).
// ================================================================================================
// order-preserving float -> uint map (lower float <=> lower uint), for the packed atomic argmin
__device__ __forceinline__ unsigned int float_to_ordered_u(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_u_to_float(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// shared memory: per slot of the CTA and axis the Euler increments of a window of kStageSteps steps
// (double) | per slot (sx[P] sy[P] syaw[P]) | bitmap[DW] maybe-map[DW] sure-map[DW] (k_dilate)
#ifndef KC_TILE_SLOTS
#define KC_TILE_SLOTS 4
#endif
constexpr int kTileSlots = KC_TILE_SLOTS;
constexpr int kStageSteps = 32;
constexpr int kChainSlots = 16;  // slots whose two chains (x, y) fill one warp in phase A
constexpr int kIncStride = kStageSteps + 1;  // doubles per (slot, axis) row of increments: the chain lanes
                                             // read one row each, an odd stride keeps them off each
                                             // other's banks
__host__ __device__ inline size_t rollout_smem_bytes(int P, int warps, int dil_words) {
  return sizeof(double) * (size_t)warps * kTileSlots * 2 * kIncStride +
         sizeof(float) * ((size_t)warps * kTileSlots * 3 * P + (size_t)3 * dil_words);
}
// shared memory of k_cost_bounds / k_cost_eval: segX[S] segY[S] | per warp sx[P] sy[P] pmin[P]
// (every array starts on a 16-byte boundary: S and P are rounded up to multiples of four floats)
__host__ __device__ inline int pad4(int n) { return (n + 3) & ~3; }
__host__ __device__ inline size_t cost_smem_bytes(int P, int S, int warps) {
  return sizeof(float) * ((size_t)2 * pad4(S) + (size_t)warps * 3 * pad4(P));
}


// One CTA per block of warps x kTileSlots consecutive velocity slots.
//  Phase A, the kinematics, CTA-wide: lane 2c + a of the first warps carries axis a (x or y) of the
//  CTA's slot c through the P-1 Euler steps, i.e. the order-sensitive running sum
//  x += (vx cos(yaw) - vy sin(yaw)) dt in the reference's serial order (ref: path.h:24-30,
//  trajectory_sampler.cpp:134-155), the heading terms coming from the table row of the slot's omega.
//  All 32 lanes of a chain warp advance a chain each (32 slots = 64 chains = two full warps); the
//  heading rows travel through shared memory in windows of kStageSteps steps, staged by ALL warps
//  (warp w stages the rows of its own tile, lane <-> step: one round trip to L2 per window, the next
//  window in flight in registers while the chains run over the current one).
//  Phase B, one warp per tile of kTileSlots slots, slot by slot: per-pose collision test with one lane
//  per pose (k_dilate's maps: clear -> sure hit -> row masks -> FP32 filter -> exact FP64 test),
//  padding, bookkeeping, row store.
template <bool STORE_VEL, bool GENERAL>
__global__ void __launch_bounds__(kEvalWarps * 32, GENERAL ? 2 : 4) k_rollout_collide(const RobotCtx *__restrict__ ctxs) {
  extern __shared__ __align__(16) float smem[];
  __shared__ int s_row[kEvalWarps * kTileSlots];      // heading-table row of every slot of the CTA
  __shared__ float s_vel[kEvalWarps * kTileSlots][3];  // its velocity triple (float, as stored)
  __shared__ double s_vd[kEvalWarps * kTileSlots][2];  // vx, vy (double, as the rollout uses them)
  __shared__ unsigned s_moves;                         // bit c: slot c is a moving sample
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(7);
  const int P = cx.P;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int cta_slots = warps * kTileSlots;
  double *inc_all = reinterpret_cast<double *>(smem);
  float *tile_all = reinterpret_cast<float *>(inc_all + (size_t)cta_slots * 2 * kIncStride);
  // the pose test's maps: the CTA keeps the bitmap and k_dilate's maps of it in shared memory when
  // they are small enough (dil_W > 0), so that the lookups of phase B never leave the SM
  const int DW = (cx.dil_W > 0 && cx.coll_enabled) ? cx.bm_rows * cx.bm_wpr : 0;
  uint32_t *s_maps = reinterpret_cast<uint32_t *>(tile_all + (size_t)cta_slots * 3 * P);
  const int S0 = blockIdx.x * cta_slots;
  const int n_cta = max(0, min(cta_slots, cx.n_slots - S0));
  if (n_cta == 0) return;  // CTA-uniform
  KC_RSTAMP_DECL;
  const int s0 = S0 + wid * kTileSlots;
  const int n_here = max(0, min(kTileSlots, cx.n_slots - s0));
  if (threadIdx.x == 0) s_moves = 0u;
  // ---- the CTA's slots are decoded by the chain lanes (slot c <-> lanes 2c, 2c+1 of warp c / 16) ----
  const int chain_warps = (cta_slots + kChainSlots - 1) / kChainSlots;
  const int cs = wid * kChainSlots + (lane >> 1), axis = lane & 1;
  const bool chain_lane = wid < chain_warps && cs < n_cta;
  SlotVel v;
  v.vx = v.vy = v.om = 0.0;
  v.row = v.srow = 0;
  bool moves = false;
  __syncthreads();  // s_moves cleared
  KC_RSTAMP_SUM(20);
  if (wid < chain_warps) {
    const SlotVel v0 = warp_decode_slot(cx, S0, lane);
    if (chain_lane) {
      v = next_slot(cx, v0, S0, S0 + cs);
      moves = slot_moves(v);
      KC_RSTAMP_SUM(21);
      if (axis == 0) {
        s_row[cs] = v.row;
        s_vel[cs][0] = (float)v.vx;
        s_vel[cs][1] = (float)v.vy;
        s_vel[cs][2] = (float)v.om;
        s_vd[cs][0] = v.vx;
        s_vd[cs][1] = v.vy;
        if (moves) atomicOr(&s_moves, 1u << cs);
      }
    }
  }
  __syncthreads();
  KC_RSTAMP_SUM(8);
  // ---- phase A: increments (all warps, own tile, lane <-> step) / chains (chain warps) per window ----
  // The Euler increment of step k, (vx cos(yaw_k) - vy sin(yaw_k)) dt resp. (vx sin + vy cos) dt, does
  // not depend on the running sum: every lane forms the increments of one step of its tile's slots
  // (same operations, same order as the serial loop) and leaves them in shared memory; the chain is
  // then nothing but the order-sensitive additions x += inc_k, eight operands fetched ahead.
  const double2 *trow[kTileSlots];
  double2 pre[kTileSlots];
#pragma unroll
  for (int s = 0; s < kTileSlots; ++s) {
    const int c = wid * kTileSlots + s;
    trow[s] = cx.tab_sc + (size_t)(s < n_here ? s_row[c] : 0) * (P - 1);
    pre[s] = (s < n_here && lane < P - 1) ? __ldg(&trow[s][lane]) : make_double2(0.0, 0.0);
  }
  {
    float *dst = tile_all + (size_t)(chain_lane ? cs : 0) * 3 * P + (size_t)axis * P;
    double a = axis ? cx.pose_y : cx.pose_x;
    if (moves) dst[0] = (float)a;
    // increments of the window: [slot][axis][kIncStride] doubles
    const double *tab = inc_all + ((size_t)(chain_lane ? cs : 0) * 2 + axis) * kIncStride;
    double *mine = inc_all + (size_t)wid * kTileSlots * 2 * kIncStride;
    const double dt = cx.dt;
    for (int k0 = 0; k0 < P - 1; k0 += kStageSteps) {
#pragma unroll
      for (int s = 0; s < kTileSlots; ++s) {
        const double2 sc = pre[s];  // {sin, cos} of the yaw before step k0 + lane
        const double tvx = s_vd[wid * kTileSlots + s][0], tvy = s_vd[wid * kTileSlots + s][1];
        const double x1 = tvx * sc.y, x2 = tvy * sc.x;
        const double y1 = tvx * sc.x, y2 = tvy * sc.y;
        mine[(s * 2 + 0) * kIncStride + lane] = (x1 - x2) * dt;
        mine[(s * 2 + 1) * kIncStride + lane] = (y1 + y2) * dt;
      }
      if (k0 == 0) { KC_RSTAMP_SUM(16); }
      __syncthreads();
      if (k0 == 0) { KC_RSTAMP_SUM(17); }
      const int kn = k0 + kStageSteps + lane;  // this lane's step of the next window
#pragma unroll
      for (int s = 0; s < kTileSlots; ++s)
        if (s < n_here && kn < P - 1) pre[s] = __ldg(&trow[s][kn]);
      if (moves) {
        const int cnt = min(kStageSteps, P - 1 - k0);
        for (int k = 0; k < cnt; k += 8) {
          double v8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v8[j] = tab[k + j];  // (a row is kStageSteps long: in bounds)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (k + j < cnt) {
              a = a + v8[j];
              dst[k0 + k + j + 1] = (float)a;
            }
        }
      }
      if (k0 == 0) { KC_RSTAMP_SUM(18); }
      __syncthreads();
      if (k0 == 0) { KC_RSTAMP_SUM(19); }
    }
  }
  KC_RSTAMP_SUM(9);
  if (DW > 0) {  // CTA-uniform
    grid_dep_wait();  // k_dilate's maps (everything above reads only what k_prep_points left)
    const uint32_t *gd = cx.dil_maps, *gs = cx.dil_maps + cx.dil_stride;
    for (int i = threadIdx.x; i < DW; i += blockDim.x) {
      s_maps[i] = __ldg(&cx.bitmap[i]);
      s_maps[DW + i] = __ldg(&gd[i]);
      s_maps[2 * DW + i] = __ldg(&gs[i]);
    }
    __syncthreads();
  }
  if (n_here == 0) return;  // warp-uniform (no CTA-wide barrier below)
  float *tile = tile_all + (size_t)wid * kTileSlots * 3 * P;
  const bool box = cx.shape == KC_BOX || GENERAL;  // the pose test needs the heading
  if (box) {  // headings of the poses, from the same table rows
    for (int s = 0; s < n_here; ++s) {
      const int row = s_row[wid * kTileSlots + s];
      float *syaw = tile + (size_t)s * 3 * P + 2 * P;
      const float *gy = cx.tab_yaw + (size_t)row * (P - 1);
      for (int j = lane; j < P - 1; j += 32) syaw[j + 1] = gy[j];
      if (lane == 0) syaw[0] = (float)cx.pose_yaw;
    }
  }
  __syncwarp();
  // ---- phase B ----
  if (lane == 0) { KC_RSTAMP_LANE(13); }
  const bool have_dil = DW > 0;
  const uint32_t *bmap = have_dil ? s_maps : cx.bitmap;
  const uint32_t *dil = have_dil ? s_maps + DW : nullptr;
  const uint32_t *sure = have_dil ? s_maps + 2 * DW : nullptr;
  const unsigned mv = s_moves >> (wid * kTileSlots);
  unsigned okmask = 0;
  for (int s = 0; s < n_here; ++s) {
    const int slot = s0 + s;
    float *sx = tile + (size_t)s * 3 * P, *sy = sx + P, *syaw = sy + P;
    bool ok = (mv >> s) & 1u;
    int cut = P - 1;
    if (ok) ok = warp_collide_slot<GENERAL>(cx, bmap, dil, sure, sx, sy, box ? syaw : nullptr, lane, cut);
    if (ok) okmask |= 1u << s;
    if (lane == 0) {
      cx.adm[slot] = ok ? 1 : 0;
      if (!STORE_VEL) {
        cx.cutv[slot] = cut;
        if (!ok) cx.costs[slot] = FLT_MAX;
      }
    }
    if (ok) {
      const size_t rp = (size_t)slot * P;
      for (int j = lane; j < P; j += 32) {
        cx.rows_x[rp + j] = sx[j];
        cx.rows_y[rp + j] = sy[j];
      }
      if (STORE_VEL) {
        const float *sv = s_vel[wid * kTileSlots + s];
        const float fvx = sv[0], fvy = sv[1], fom = sv[2];
        const size_t rv = (size_t)slot * (P - 1);
        for (int j = lane; j < P - 1; j += 32) {
          cx.rows_vx[rv + j] = (j < cut) ? fvx : 0.0f;
          cx.rows_vy[rv + j] = (j < cut) ? fvy : 0.0f;
          cx.rows_om[rv + j] = (j < cut) ? fom : 0.0f;
        }
      }
    }
    __syncwarp();
  }
  if (!STORE_VEL && lane == 0 && okmask) {  // one append per tile (the list is unordered)
    int at = atomicAdd(cx.n_list, __popc(okmask));
    for (int s = 0; s < n_here; ++s)
      if ((okmask >> s) & 1u) cx.list[at++] = s0 + s;
  }
  KC_RSTAMP_WARP(10, 11, 12);
}

// k_cost_bounds (branch and bound, stage 1): goal + path cost of every admissible slot (kept in
// costs[] for stage 2), bounds of its obstacle term from the distance brackets, smoothness / jerk of
// padded rows; lower bound of the total -> lbv[], smallest upper bound over all slots -> ub_inv.
// All terms are >= 0 and float addition is monotone, so partial sums bound the total from below;
// the brackets carry explicit slack for the float evaluation of the bound itself.
// (Measured and dropped in round 2: the goal + path part as its own kernel behind the rollouts, beside
// the obstacle-grid preparation - the critical path gains nothing because the two branches then
// compete for the same issue slots, and the sweep loses 3 us per robot to the second pass over the rows.)
// resident CTAs per SM the bound stage is compiled for: 5 (48 registers, 36 bytes of spills) puts
// 5 920 warps on the GPU instead of 4 736 - measured 1 us per cycle and 0.6 us per robot of a sweep
// against 4; 6 (40 registers) gains on the dense cloud and loses on the ring
#ifndef KC_BOUNDS_CTAS
#define KC_BOUNDS_CTAS 5
#endif
__global__ void __launch_bounds__(kEvalWarps * 32, KC_BOUNDS_CTAS) k_cost_bounds(const RobotCtx *__restrict__ ctxs) {
  extern __shared__ __align__(16) float smem[];
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(8);
  const int P = cx.P, S = cx.seg_count;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int n_list = *cx.n_list;
  grid_dep_launch();  // k_cost_split may be staged behind this grid
  float *segX = smem, *segY = segX + pad4(S);
  float *sx = segY + pad4(S) + (size_t)wid * 3 * pad4(P);
  float *sy = sx + pad4(P), *pmin = sy + pad4(P);
  const int G = gridDim.x * warps;
  if ((int)blockIdx.x * warps < n_list && cx.path_enabled) {
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
      segX[j] = cx.pathX[cx.seg_start + j];
      segY[j] = cx.pathY[cx.seg_start + j];
    }
  }
  __syncthreads();
  float ub_min = INFINITY;
  for (int li = blockIdx.x * warps + wid; li < n_list; li += G) {
    const int slot = cx.list[li];
    const int cut = cx.cutv[slot];
    const size_t rp = (size_t)slot * P;
    for (int j = lane; j < P; j += 32) {
      sx[j] = cx.rows_x[rp + j];
      sy[j] = cx.rows_y[rp + j];
    }
    __syncwarp();
    const float partial = warp_partial_cost(cx, segX, segY, sx, sy, pmin, lane);
    float c_lo = 0.0f, c_hi = 0.0f, dhi = INFINITY;
    if (cx.obs_enabled) {
      float dlo;
      warp_obstacle_bracket(cx, sx, sy, lane, dlo, dhi);
      c_hi = fminf(fmaxf(cx.D - dlo, 0.0f) / cx.D * 1.00001f + 1e-6f, 1.00001f);
      c_lo = fmaxf(fmaxf(cx.D - dhi, 0.0f) / cx.D * 0.99999f - 1e-6f, 0.0f);
    }
    float sj = 0.0f;  // smoothness + jerk: exact zeros for constant-velocity rows
    if (cut != P - 1) {
      const SlotVel v = decode_slot(cx, slot);
      const float fvx = (float)v.vx, fvy = (float)v.vy, fom = (float)v.om;
      auto vel = [&](int c, int j) -> float {
        return (j < cut) ? (c == 0 ? fvx : (c == 1 ? fvy : fom)) : 0.0f;
      };
      // (the two unweighted terms are kept for k_cost_eval's totals: sjv)
      float cs = 0.0f, cj = 0.0f;
      if (cx.w_smooth > 0.0) {
        cs = warp_smoothness(vel, P - 1, cx.acc0, cx.acc1, cx.acc2, lane);
        sj += (float)(cx.w_smooth * (double)cs);
      }
      if (cx.w_jerk > 0.0) {
        cj = warp_jerk(vel, P - 1, cx.acc0, cx.acc1, cx.acc2, lane);
        sj += (float)(cx.w_jerk * (double)cj);
      }
      if (lane == 0) cx.sjv[slot] = make_float2(cs, cj);
    }
    const float wo = (float)cx.w_obs;
    // slack for the float evaluation of the bounds themselves, sign-aware (a goal term can come out
    // a few ulp below zero when the accumulated length slightly exceeds the total length)
    const float lsum = partial + wo * 0.99999f * c_lo + sj * 0.99999f;
    const float usum = partial + wo * 1.00001f * c_hi + sj * 1.00001f;
    const float lb = lsum - fabsf(lsum) * 1e-5f - 1e-6f;
    const float ub = usum + fabsf(usum) * 1e-5f + 1e-6f;
    if (lane == 0) {
      cx.costs[slot] = partial;
      cx.lbv[slot] = lb;
      cx.ubd[slot] = dhi;
    }
    if (ub < FLT_MAX) ub_min = fminf(ub_min, ub);
    __syncwarp();
  }
  if (lane == 0 && ub_min < FLT_MAX) atomicMax(cx.ub_inv, ~float_to_ordered_u(ub_min));
}

// k_cost_split (branch and bound, between the stages): with the final bound known, one thread per
// admissible slot files it as pruned (done: costs[] keeps the lower bound) or as a survivor whose
// exact obstacle search runs in k_cost_eval.
__global__ void k_cost_split(const RobotCtx *__restrict__ ctxs) {
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(9);
  const int n_list = *cx.n_list;  // (written by the rollout kernel, complete before k_cost_bounds began)
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  grid_dep_launch();
  grid_dep_wait();  // bounds of every slot and the final smallest upper bound
  if (li >= n_list) return;
  float ustar = INFINITY;
  {
    const unsigned int inv = *cx.ub_inv;
    if (inv) ustar = ordered_u_to_float(~inv);
  }
  const int slot = cx.list[li];
  const float lb = cx.lbv[slot];
  if (lb > ustar) {  // its total exceeds some other slot's total: it cannot be the argmin
    cx.costs[slot] = lb;
    cx.prn[slot] = 1;
  } else {
    cx.prn[slot] = 0;
    const float u = cx.ubd[slot] * 1.0001f;
    double d0 = cx.dcap2;  // start of the running minimum: the cut-off, or the slot's upper bracket
    if (u < FLT_MAX) d0 = fmin(d0, (double)u * (double)u);
    cx.dmin_bits[slot] = (unsigned long long)__double_as_longlong(d0);
    cx.surv[atomicAdd(cx.n_surv, 1)] = slot;
  }
}

// CTAS: resident CTAs per SM the kernel is compiled for. A control cycle runs the 80-register build (3 per
// SM: the exact stage of one robot is a latency chain, spills only lengthen it); a sweep runs the
// 64-register build (4 per SM, 172 bytes of spills): measured 1.6 us per robot faster there, 0.5 us per
// cycle slower for a single robot.
template <int CTAS>
__global__ void __launch_bounds__(kEvalWarps * 32, CTAS) k_cost_eval(const RobotCtx *__restrict__ ctxs) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned long long s_key[kEvalWarps];
  __shared__ int s_last;
  const RobotCtx &cx = ctxs[blockIdx.y];
  KC_TL(10);
  KC_STAMP_MIN(0);
  const int P = cx.P, S = cx.seg_count;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int n_list = *cx.n_list;
  float *segX = smem, *segY = segX + pad4(S);
  float *sx = segY + pad4(S) + (size_t)wid * 3 * pad4(P);
  float *sy = sx + pad4(P), *pmin = sy + pad4(P);
  // resident CTAs: every warp strides over the list of admissible slots (uniform work per entry)
  const int G = gridDim.x * warps;
  // (after k_cost_bounds the goal + path cost of every slot is already in costs[]: no segment needed)
  if ((int)blockIdx.x * warps < n_list && cx.path_enabled && !cx.prune) {
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
      segX[j] = cx.pathX[cx.seg_start + j];
      segY[j] = cx.pathY[cx.seg_start + j];
    }
  }
  __syncthreads();
  grid_dep_wait();  // (prologue above: nothing k_cost_bounds / k_cost_split write)
  unsigned long long my_key = ~0ull;
  // stage 2 of the branch and bound: only the slots k_cost_bounds could not rule out are left.
  // Few of them (the usual case): their (slot, point) pairs are spread over all warps, one pair per
  // warp and step, each updating the slot's running minimum; the last CTA then forms the totals.
  // Many of them (ties, loose bounds): one warp per slot as in the unpruned evaluation.
  const int n_work = cx.prune ? *cx.n_surv : n_list;
  const int *work = cx.prune ? cx.surv : cx.list;
  // by point: up to by_point_max survivors (more: ties, loose bounds - one warp per slot as in the
  // unpruned evaluation, whose per-slot radius shrinks as its points are searched)
  const bool by_point = cx.prune && cx.obs_enabled && n_work <= cx.by_point_max;
  KC_STAMP_SET(6, n_work);
  KC_STAMP_SET(7, by_point ? 1 : 0);
  if (by_point) {
    // A warp takes `bw` consecutive (slot, point) pairs at a time, one per lane: the lanes fetch their
    // pair's coordinates, the slot's running minimum and the cell record in parallel (one chain of
    // round trips for the whole batch) and drop the pairs whose lower bracket cannot improve the
    // minimum; the pairs that are left are searched one after the other by the whole warp. bw grows
    // with the number of pairs so that every warp of the grid gets a batch: 1 for a handful of
    // survivors (all pairs searched side by side), 32 for thousands.
    const long long items = (long long)n_work * P;
    const int bw = (int)min(32LL, max(1LL, (items + G - 1) / G));
    for (long long b0 = ((long long)blockIdx.x * warps + wid) * bw; b0 < items; b0 += (long long)G * bw) {
      const long long it = b0 + lane;
      bool act = lane < bw && it < items;
      int slot = 0, kind = 0;  // kind 1: candidate list, 2: no list (own disc), 3: outside the window
      float px = 0.0f, py = 0.0f;
      double best = 0.0;
      float lbf = 0.0f;  // lower bracket of this pair's distance (0: none)
      int4 ci = make_int4(0, 0, 0, 0);
      if (act) {
        slot = work[it / P];
        const int kp = (int)(it % P);
        const size_t rp = (size_t)slot * P;
        px = cx.rows_x[rp + kp];
        py = cx.rows_y[rp + kp];
        best = __longlong_as_double((long long)__ldcg(&cx.dmin_bits[slot]));
        int cell;
        if (query_cell(cx, px, py, cell)) {
          ci = __ldg(&cx.cell_info[cell]);
          const float dm = __int_as_float(ci.x);
          lbf = (dm == dm) ? fmaxf(0.0f, dm * 0.999f - bracket_halfwidth(cx, px, py, cell)) : 0.0f;
          if (!((double)lbf * (double)lbf < best)) act = false;  // this point cannot improve the minimum
          kind = (ci.z < 0) ? 2 : 1;
          if (kind == 1 && ci.z == 0) act = false;
        } else {
          kind = 3;
        }
      }
      // wide batches: pairs with a short candidate list are walked by their own lane, side by side;
      // the others then start from the slot's refreshed minimum
      if (bw >= 4) {
        const bool own = act && kind == 1 && ci.z <= kCandSerial;
        if (__any_sync(FULL, own)) {
          if (own) {
            act = false;
            float bestf = conservative_f(best);
            const float2 *cand = cx.cand_pool + ci.y;
            double mine = best;
            for (int q = 0; q < ci.z; ++q) {
              const float2 o = __ldg(&cand[q]);
              const float dx = o.x - px, dy = o.y - py;
              const float d2f = __fmaf_rn(dx, dx, dy * dy);  // filter only (<= 2 ulp off)
              if (d2f <= bestf) {
                const double d2 = (double)dx * (double)dx + (double)dy * (double)dy;
                if (d2 < mine) {
                  mine = d2;
                  bestf = conservative_f(mine);
                }
              }
            }
            if (mine < best) atomicMin(&cx.dmin_bits[slot], (unsigned long long)__double_as_longlong(mine));
          }
          __syncwarp();
          if (act) {  // (any value read here is a valid upper bound of the slot's minimum)
            best = fmin(best, __longlong_as_double((long long)__ldcg(&cx.dmin_bits[slot])));
            if (!((double)lbf * (double)lbf < best)) act = false;
          }
        }
      }
      // most promising pair first (smallest lower bracket); its result shrinks the radius of the
      // batch's other pairs of the same slot, most of which then drop out without a search
      for (;;) {
        const unsigned todo = __ballot_sync(FULL, act);
        if (!todo) break;
        float key = act ? lbf : FLT_MAX;
        int src = lane;
        warp_argmin_f(key, src);
        const int qs = __shfl_sync(FULL, slot, src), qk = __shfl_sync(FULL, kind, src);
        const float qx = __shfl_sync(FULL, px, src), qy = __shfl_sync(FULL, py, src);
        const double qb = shfl_d(best, src);
        double got = qb;
        KC_PH_DECL;
        if (qk == 1) {
          const int start = __shfl_sync(FULL, ci.y, src), cnt = __shfl_sync(FULL, ci.z, src);
          const float bestf = conservative_f(qb);
          const float2 *cand = cx.cand_pool + start;
          double mine = qb;
          for (int q = lane; q < cnt; q += 32) {
            const float2 o = __ldg(&cand[q]);
            const float dx = o.x - qx, dy = o.y - qy;
            const float d2f = __fmaf_rn(dx, dx, dy * dy);  // filter only (<= 2 ulp off)
            if (d2f <= bestf) {
              const double d2 = (double)dx * (double)dx + (double)dy * (double)dy;
              mine = fmin(mine, d2);
            }
          }
          got = warp_min_d(mine);
        } else if (qk == 2) {
          got = warp_nn_search_one(cx, qx, qy, qb, lane);
        } else {
          got = nn_search_batch(cx, qx, qy, lane == 0, qb);
        }
        if (lane == 0 && got < qb)
          atomicMin(&cx.dmin_bits[qs], (unsigned long long)__double_as_longlong(got));
#ifdef KC_DBG_STAMPS
        if (lane == 0) {
          atomicAdd(&cx.dbg[86 + 2 * qk], (unsigned long long)(clock64() - p_t0));  // cycles of kind qk (1..3)
          atomicAdd(&cx.dbg[87 + 2 * qk], 1ull);
        }
#endif
        if (lane == src) act = false;
        if (act && slot == qs) {
          best = fmin(best, got);
          if (!((double)lbf * (double)lbf < best)) act = false;
        }
      }
    }
  }
  KC_STAMP_MAX(1);
  for (int li = blockIdx.x * warps + wid; li < (by_point ? 0 : n_work); li += G) {
    const int slot = work[li];
    const int cut = cx.cutv[slot];
    const size_t rp = (size_t)slot * P;
    for (int j = lane; j < P; j += 32) {
      sx[j] = cx.rows_x[rp + j];
      sy[j] = cx.rows_y[rp + j];
    }
    __syncwarp();
    const SlotVel v = decode_slot(cx, slot);
    const float fvx = (float)v.vx, fvy = (float)v.vy, fom = (float)v.om;
    auto vel = [&](int c, int j) -> float {
      return (j < cut) ? (c == 0 ? fvx : (c == 1 ? fvy : fom)) : 0.0f;
    };
    const float total = cx.prune ? warp_finish_cost(cx, cx.costs[slot], sx, sy, vel, lane, cut == P - 1)
                                 : warp_total_cost(cx, segX, segY, sx, sy, pmin, vel, lane, cut == P - 1);
    if (lane == 0) cx.costs[slot] = total;
    // strict '<' against FLT_MAX: NaN / inf totals never win (cost_evaluator.cpp:102)
    if (total < FLT_MAX)
      my_key = min(my_key, ((unsigned long long)float_to_ordered_u(total) << 32) | (unsigned int)slot);
    __syncwarp();
  }
  if (lane == 0) s_key[wid] = my_key;
  __syncthreads();
  // ---- block argmin -> one 64-bit atomic; the last CTA of this robot publishes the result ----
  if (threadIdx.x == 0) {
    unsigned long long key = ~0ull;
    for (int w = 0; w < warps; ++w) key = min(key, s_key[w]);
    if (key != ~0ull) atomicMax(cx.best_key, ~key);  // zero-initialised => max of inverted keys
    fence_acq_rel();
    const unsigned int ticket = atomicAdd(cx.done_ctr, 1u);
    s_last = (ticket == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  KC_STAMP_MAX(2);
  if (!s_last) return;
  KC_STAMP_MAX(3);
  unsigned long long final_key = ~0ull;  // thread 0: the winner's key (read back from the global argmin)
  if (by_point) {
    // the running minima are final: total = partial (+) obstacles (+) smoothness (+) jerk per survivor,
    // ONE THREAD per survivor (the smoothness / jerk terms of padded rows come from k_cost_bounds: sjv);
    // the loads of four survivors are in flight together
    fence_acq_rel();
    unsigned long long key2 = ~0ull;
    const int T = blockDim.x;
    for (int s0 = threadIdx.x; s0 < n_work; s0 += 4 * T) {
      int slot[4], cut[4];
      float tot[4];
      double d2[4];
      float2 sj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) slot[u] = (s0 + u * T < n_work) ? work[s0 + u * T] : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (slot[u] >= 0) {
          cut[u] = cx.cutv[slot[u]];
          tot[u] = __ldcg(&cx.costs[slot[u]]);
          d2[u] = __longlong_as_double((long long)__ldcg(&cx.dmin_bits[slot[u]]));
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (slot[u] >= 0 && cut[u] != P - 1) sj[u] = __ldcg(&cx.sjv[slot[u]]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (slot[u] >= 0) {
          float total = tot[u];
          if (d2[u] < cx.dcap2) {
            const float md = (float)d2[u];
            const float dist = (float)sqrt((double)md);
            const float c = fmaxf(cx.D - dist, 0.0f) / cx.D;
            total = (float)((double)total + cx.w_obs * (double)c);
          }
          if (cut[u] != P - 1) {
            if (cx.w_smooth > 0.0) total = (float)((double)total + cx.w_smooth * (double)sj[u].x);
            if (cx.w_jerk > 0.0) total = (float)((double)total + cx.w_jerk * (double)sj[u].y);
          }
          cx.costs[slot[u]] = total;
          if (total < FLT_MAX)
            key2 = min(key2, ((unsigned long long)float_to_ordered_u(total) << 32) | (unsigned int)slot[u]);
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) key2 = min(key2, __shfl_xor_sync(FULL, key2, m));
    if (lane == 0) s_key[wid] = key2;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long key = ~0ull;
      for (int w = 0; w < warps; ++w) key = min(key, s_key[w]);
      s_key[0] = key;  // no other CTA contributes in this mode: the block's argmin is the winner
    }
    __syncthreads();
    final_key = s_key[0];
  } else {
    fence_acq_rel();
    final_key = ~*((volatile unsigned long long *)cx.best_key);
  }
  KC_STAMP_MAX(4);
  if (wid == 0) {
    const uint32_t heavy_seen = cx.obs_enabled ? (uint32_t)__ldcg(cx.heavy_ctr) : 0u;  // (in flight beside the row loads)
    const unsigned long long key = final_key;
    const bool found = key != ~0ull;
    const int win = (int)(unsigned int)(key & 0xffffffffull);
    // The record may live in mapped host memory, where every store instruction is a PCIe write and
    // every system-scope fence a round trip. The warp assembles header and rows and writes them with
    // coalesced stores (32 consecutive words per instruction); every lane then orders ITS OWN stores
    // at system scope, the warp meets, and lane 0 alone publishes the sequence number.
    const int n_out = 3 * (P - 1) + 2 * P;
    float *o = cx.res_rows;
    KC_STAMP_MAX(24);
    if (found) {  // the winner's row is already in memory (k_rollout_collide stored it)
      const SlotVel wv = warp_decode_slot(cx, win, lane);
      const int wcut = cx.cutv[win];
      KC_STAMP_MAX(25);
      const float wvx = (float)wv.vx, wvy = (float)wv.vy, wom = (float)wv.om;
      const size_t rp = (size_t)win * P;
      // loads first (eight words per lane in flight), then the stores
      for (int j0 = 0; j0 < n_out; j0 += 256) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + 32 * u + lane;
          v[u] = 0.0f;
          if (j < 3 * (P - 1)) {
            const int c = j / (P - 1), jj = j - c * (P - 1);
            if (jj < wcut) v[u] = (c == 0) ? wvx : (c == 1 ? wvy : wom);
          } else if (j < n_out) {
            const int jj = j - 3 * (P - 1);
            v[u] = (jj < P) ? cx.rows_x[rp + jj] : cx.rows_y[rp + jj - P];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = j0 + 32 * u + lane;
          if (j < n_out) o[j] = v[u];
        }
      }
    }
    if (lane == 0) {
      cx.result->found = found ? 1 : 0;
      cx.result->cost = found ? ordered_u_to_float((unsigned int)(key >> 32)) : FLT_MAX;
      cx.result->slot = found ? win : -1;
      cx.result->n_admissible = n_list;
      cx.result->heavy_cells = heavy_seen;
    }
    KC_STAMP_MAX(26);
    __threadfence_system();
    __syncwarp();
    KC_STAMP_MAX(27);
    // every lane's stores (lane 0's header words included) were performed at system scope before that
    // lane reached the barrier: the sequence number cannot overtake them
    if (lane == 0) *((volatile uint32_t *)&cx.result->seq) = cx.seq;
  }
  KC_STAMP_MAX(5);
}

// ================================================================================================
// Brute-force obstacle term (verification + FP32 roofline kernel, never on the control path).
// The reference's TrajectoryPath::minDist2D as written (trajectory.h:218-235): every admissible
// trajectory point against EVERY sensor point, no cull, no grid: N*P*M pair evaluations of
// (2 FSUB, 1 FMUL, 1 FFMA, 1 FMNMX). Layout as the north_star sketches it: obstacle points staged
// through shared memory with cp.async (double buffered) and reused by all warps of the CTA; each
// warp keeps 256 (trajectory, point) entries in registers (8 per lane, entries of consecutive
// admissible trajectories packed back to back so no lane idles when P is not a multiple of 32);
// per-entry minima fold into per-trajectory minima with one atomicMin each.
//   pass 1 (REFINE = false): FP32 minimum m32 per trajectory (FFMA-contracted d^2: <= 2 ulp off).
//   pass 2 (REFINE = true): every pair with d32 <= m32 * (1 + 1e-6) is re-evaluated with the
//   reference's arithmetic float(double(dx)^2 + double(dy)^2); the minimum of those IS the
//   reference's minimum (the exact minimiser has d32 <= m32 (1 + 2^-22)^2).
// PACKED (default; KC_BF_SCALAR=1 selects the scalar form): two obstacle points per sm_100 packed
// FP32 instruction (FADD2 with a broadcast scalar subtrahend, FMUL2, FFMA2), the points stored as
// (x0, x1, y0, y1) so one LDS.128 delivers both register pairs: 2.5 instead of 4.5 issue slots per
// pair, each lane operation still one IEEE round-to-nearest -> the same bits. The FP32 pipe
// (4 lane-cycles per pair for 6 FLOP: 75 % of the FMA peak) becomes the bound instead of issue.
// ================================================================================================
constexpr int kBfTile = 2048;    // obstacle points per shared-memory tile (16 KB, two buffers)
constexpr int kBfEntries = 8;    // (trajectory, point) entries per lane
constexpr float kBfFar = 3.0e38f;

// paired = 1: two consecutive points share one 16-byte word as (x0, x1, y0, y1), the operand layout
// of the packed FP32 instructions; an odd count is padded with a far point.
__global__ void k_transform_points(const RobotCtx *__restrict__ ctxs, float2 *__restrict__ out, int paired) {
  const RobotCtx &cx = ctxs[0];
  const int n = cx.n_sensor;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float qx, qy, qz;
    if (cx.sensor_is_cloud) {
      const float *xyz = reinterpret_cast<const float *>(cx.sensor);
      qx = xyz[3 * i];
      qy = xyz[3 * i + 1];
      qz = xyz[3 * i + 2];
    } else {
      const double *ranges = reinterpret_cast<const double *>(cx.sensor);
      const double r = ranges[i], a = ranges[n + i];
      double s, c;
      sincos(a, &s, &c);
      qx = (float)(r * c);
      qy = (float)(r * s);
      qz = 0.0f;
    }
    const float *T = cx.T;  // ref: cost_evaluator.h:187-189
    const float2 q = make_float2(T[9] + (T[0] * qx + (T[1] * qy + T[2] * qz)),
                                 T[10] + (T[3] * qx + (T[4] * qy + T[5] * qz)));
    if (!paired) {
      out[i] = q;
    } else {
      float *w = reinterpret_cast<float *>(out) + 4 * (size_t)(i >> 1) + (i & 1);
      w[0] = q.x;
      w[2] = q.y;
      if (i == n - 1 && !(i & 1)) w[1] = w[3] = kBfFar;
    }
  }
}

template <bool PAIRED>
__device__ __forceinline__ void bf_stage_tile(float2 *dst, const float2 *__restrict__ obs, int base, int M) {
  // 16-byte cp.async per thread-iteration (two points); the ragged tail is filled by hand
  for (int k = threadIdx.x * 2; k < kBfTile; k += blockDim.x * 2) {
    const int g = base + k;
    if (g + 1 < M || (PAIRED && g < M)) {  // paired layout: the odd last point is padded in global memory
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + k);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(obs + g));
    } else {
      dst[k] = (g < M) ? obs[g] : make_float2(kBfFar, kBfFar);
      dst[k + 1] = make_float2(kBfFar, kBfFar);
    }
  }
  asm volatile("cp.async.commit_group;");
}

template <bool REFINE, bool PACKED>
__global__ void __launch_bounds__(256) k_obstacle_bruteforce(const RobotCtx *__restrict__ ctxs,
                                                             const float2 *__restrict__ obs, int M,
                                                             int tiles_per_chunk,
                                                             unsigned int *__restrict__ min32,
                                                             unsigned int *__restrict__ min_exact) {
  __shared__ __align__(16) float2 tile[2][kBfTile];
  const RobotCtx &cx = ctxs[0];
  const int P = cx.P, n_list = *cx.n_list;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long total = (long long)n_list * P;               // flat (trajectory, point) entries
  const long long per_cta = 8LL * 32 * kBfEntries;             // entries per CTA
  // blockIdx.x: block of 2048 entries; blockIdx.y: chunk of obstacle tiles (work units much finer
  // than the number of resident CTAs, partial minima merge through atomicMin)
  const int all_tiles = (M + kBfTile - 1) / kBfTile;
  const int tile0 = blockIdx.y * tiles_per_chunk;
  const int n_tiles = min(tiles_per_chunk, all_tiles - tile0);
  const long long base = (long long)blockIdx.x * per_cta;
  if (base < total && n_tiles > 0) {
    float x[kBfEntries], y[kBfEntries], m[kBfEntries], thr[kBfEntries];
    int tr[kBfEntries];
#pragma unroll
    for (int e = 0; e < kBfEntries; ++e) {
      const long long id = base + (long long)wid * 32 * kBfEntries + e * 32 + lane;
      tr[e] = -1;
      x[e] = y[e] = kBfFar;
      m[e] = FLT_MAX;
      thr[e] = 0.0f;
      if (id < total) {
        const int li = (int)(id / P), j = (int)(id - (long long)li * P);
        const int slot = cx.list[li];
        tr[e] = slot;
        x[e] = cx.rows_x[(size_t)slot * P + j];
        y[e] = cx.rows_y[(size_t)slot * P + j];
        if (REFINE) thr[e] = __uint_as_float(min32[slot]) * 1.000001f + 1e-37f;
      }
    }
    bf_stage_tile<PACKED>(tile[0], obs, tile0 * kBfTile, M);
    for (int t = 0; t < n_tiles; ++t) {
      if (t + 1 < n_tiles) {
        bf_stage_tile<PACKED>(tile[(t + 1) & 1], obs, (tile0 + t + 1) * kBfTile, M);
        asm volatile("cp.async.wait_group 1;");
      } else {
        asm volatile("cp.async.wait_group 0;");
      }
      __syncthreads();
      const float4 *t4 = reinterpret_cast<const float4 *>(tile[t & 1]);
#pragma unroll 2
      for (int k = 0; k < kBfTile / 2; ++k) {
        const float4 o = t4[k];  // two obstacle points, broadcast to the warp
#pragma unroll
        for (int e = 0; e < kBfEntries; ++e) {
          float dx0, dy0, dx1, dy1, d0, d1;
          if (PACKED) {
            // o = (x0, x1, y0, y1); the subtrahend is a broadcast scalar operand of FADD2
            const float2 dx = __fadd2_rn(make_float2(o.x, o.y), make_float2(-x[e], -x[e]));
            const float2 dy = __fadd2_rn(make_float2(o.z, o.w), make_float2(-y[e], -y[e]));
            const float2 d = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
            dx0 = dx.x, dx1 = dx.y, dy0 = dy.x, dy1 = dy.y, d0 = d.x, d1 = d.y;
          } else {
            dx0 = o.x - x[e], dy0 = o.y - y[e];
            dx1 = o.z - x[e], dy1 = o.w - y[e];
            d0 = __fmaf_rn(dx0, dx0, dy0 * dy0);
            d1 = __fmaf_rn(dx1, dx1, dy1 * dy1);
          }
          if (!REFINE) {
            m[e] = fminf(m[e], fminf(d0, d1));
          } else {
            if (d0 <= thr[e]) m[e] = fminf(m[e], (float)((double)dx0 * (double)dx0 + (double)dy0 * (double)dy0));
            if (d1 <= thr[e]) m[e] = fminf(m[e], (float)((double)dx1 * (double)dx1 + (double)dy1 * (double)dy1));
          }
        }
      }
      __syncthreads();
    }
    unsigned int *dst = REFINE ? min_exact : min32;
#pragma unroll
    for (int e = 0; e < kBfEntries; ++e)
      if (tr[e] >= 0 && m[e] < FLT_MAX) atomicMin(&dst[tr[e]], __float_as_uint(m[e]));
  }
}

// reference's obstaclesDistCostFunc on the brute-force minima (cost_evaluator.cpp:179-184)
__global__ void k_bruteforce_cost(const RobotCtx *__restrict__ ctxs, const unsigned int *__restrict__ min_exact,
                                  float *__restrict__ cost) {
  const RobotCtx &cx = ctxs[0];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cx.n_slots) return;
  float c = FLT_MAX;  // inadmissible slot
  if (cx.adm[i]) {
    const float md = __uint_as_float(min_exact[i]);  // 0x7f7fffff = FLT_MAX when nothing was finite
    const float dist = (float)sqrt((double)md);
    c = fmaxf(cx.D - dist, 0.0f) / cx.D;
  }
  cost[i] = c;
}

// ================================================================================================
// k_check_states: CollisionChecker::checkCollisions(state) for a batch of states against the voxel
// bitmap of the current sensor data (ref: collision_check.cpp:125-162,225-246). One thread per state
// (x, y, yaw doubles, narrowed to float as getTransformation / eulerToRotationMatrix do).
// ================================================================================================
template <bool GENERAL>
__global__ void k_check_states(const RobotCtx *__restrict__ ctxs, const double *__restrict__ states,
                               int n, uint8_t *__restrict__ out, int *__restrict__ any) {
  const RobotCtx &cx = ctxs[0];
  bool hit_any = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const bool hit = cx.coll_enabled && pose_collides<GENERAL>(cx, nullptr, nullptr, nullptr, (float)states[3 * i],
                                                      (float)states[3 * i + 1], (float)states[3 * i + 2]);
    out[i] = hit ? 1 : 0;
    hit_any |= hit;
  }
  if (__any_sync(__activemask(), hit_any) && hit_any) atomicOr(any, 1);
}

// ================================================================================================
// k_eval_rows: CostEvaluator::getMinTrajectoryCost on caller-provided samples (warp per row)
// ================================================================================================
__global__ void __launch_bounds__(kEvalWarps * 32) k_eval_rows(const RobotCtx *__restrict__ ctxs) {
  extern __shared__ __align__(16) float smem[];
  const RobotCtx &cx = ctxs[blockIdx.y];
  const int P = cx.P, S = cx.seg_count;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  float *segX = smem + 128 * warps, *segY = segX + S;
  float *sx = segY + S + (size_t)wid * 4 * P;
  float *sy = sx + P, *pmin = sy + 2 * P;
  if (cx.path_enabled) {
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
      segX[j] = cx.pathX[cx.seg_start + j];
      segY[j] = cx.pathY[cx.seg_start + j];
    }
  }
  __syncthreads();
  const int t = blockIdx.x * (blockDim.x >> 5) + wid;
  if (t >= cx.n_traj) return;
  const size_t rp = (size_t)t * P, rv = (size_t)t * (P - 1);
  for (int j = lane; j < P; j += 32) {
    sx[j] = cx.in_x[rp + j];
    sy[j] = cx.in_y[rp + j];
  }
  __syncwarp();
  const float *pvx = cx.in_vx + rv, *pvy = cx.in_vy + rv, *pom = cx.in_om + rv;
  auto vel = [&](int c, int j) -> float {
    return c == 0 ? __ldg(&pvx[j]) : (c == 1 ? __ldg(&pvy[j]) : __ldg(&pom[j]));
  };
  float total = warp_total_cost(cx, segX, segY, sx, sy, pmin, vel, lane);
  if (cx.custom) {  // ref: cost_evaluator.cpp:96-100: total_cost += weight * custom(traj, path), one
                    // float += double per registered callback, in registration order
    const double *cu = cx.custom + (size_t)t * cx.n_custom;
    for (int k = 0; k < cx.n_custom; ++k) total = (float)((double)total + cu[k]);
  }
  if (lane == 0) {
    cx.costs[t] = total;
    cx.adm[t] = 1;
  }
}

// ================================================================================================
// k_select: argmin with lowest-index tie-break over caller-provided rows
// (ref: cost_evaluator.cpp:102-106, trajectory.h:630-636). One CTA of 1024 threads per robot.
// ================================================================================================
__global__ void __launch_bounds__(1024) k_select(const RobotCtx *__restrict__ ctxs) {
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  const RobotCtx &cx = ctxs[blockIdx.y];
  const int n = cx.n_traj;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  float best = FLT_MAX;
  int bidx = 0x7fffffff;
  for (int i = t; i < n; i += blockDim.x) {
    const float c = cx.costs[i];
    if (c < best) {  // strict: FLT_MAX / NaN are never selected
      best = c;
      bidx = i;
    }
  }
  warp_argmin_f(best, bidx);
  if (lane == 0) {
    s_val[wid] = best;
    s_idx[wid] = bidx;
  }
  __syncthreads();
  if (wid == 0) {
    best = s_val[lane];
    bidx = s_idx[lane];
    warp_argmin_f(best, bidx);
    const bool found = bidx != 0x7fffffff;
    if (lane == 0) {
      cx.result->found = found ? 1 : 0;
      cx.result->cost = best;
      cx.result->slot = found ? bidx : -1;
      cx.result->n_admissible = n;
    }
  }
}

// ================================================================================================
// sampler API: order-preserving compaction of admissible rows
// ================================================================================================
__global__ void __launch_bounds__(1024) k_compact_index(const uint8_t *__restrict__ adm, int n,
                                                        int32_t *__restrict__ dst,
                                                        int32_t *__restrict__ count) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  if (t == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + t;
    const int f = (i < n) ? adm[i] : 0;
    int incl = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = warp_sums[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int v = __shfl_up_sync(FULL, wi, d);
        if (lane >= d) wi += v;
      }
      warp_sums[lane] = wi - w;
    }
    __syncthreads();
    const int excl = carry + warp_sums[wid] + incl - f;
    if (i < n) dst[i] = f ? excl : -1;
    __syncthreads();
    if (t == 1023) carry = excl + f;
    __syncthreads();
  }
  if (t == 0) *count = carry;
}

__global__ void k_compact_rows(const RobotCtx *__restrict__ ctxs, const int32_t *__restrict__ dst,
                               float *__restrict__ ovx, float *__restrict__ ovy,
                               float *__restrict__ oom, float *__restrict__ ox,
                               float *__restrict__ oy, int32_t *__restrict__ oslots) {
  const RobotCtx &cx = ctxs[0];
  const int P = cx.P;
  const int slot = blockIdx.x;
  const int d = dst[slot];
  if (d < 0) return;
  const size_t sv = (size_t)slot * (P - 1), sp = (size_t)slot * P;
  const size_t dv = (size_t)d * (P - 1), dp = (size_t)d * P;
  for (int j = threadIdx.x; j < P - 1; j += blockDim.x) {
    ovx[dv + j] = cx.rows_vx[sv + j];
    ovy[dv + j] = cx.rows_vy[sv + j];
    oom[dv + j] = cx.rows_om[sv + j];
  }
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    ox[dp + j] = cx.rows_x[sp + j];
    oy[dp + j] = cx.rows_y[sp + j];
  }
  if (threadIdx.x == 0) oslots[d] = slot;
}

// bounding box of caller-provided sample points (evaluate mode): min/max via ordered-int atomics
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return (i >= 0) ? i : i ^ 0x7fffffff;
}
__host__ __device__ inline float ordered_to_float(int i) {
  int j = (i >= 0) ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
  return __int_as_float(j);
#else
  float f;
  memcpy(&f, &j, sizeof(f));
  return f;
#endif
}
__global__ void k_bbox(const float *__restrict__ x, const float *__restrict__ y, size_t n,
                       int *__restrict__ out /* minx,maxx,miny,maxy ordered ints */) {
  float mnx = FLT_MAX, mxx = -FLT_MAX, mny = FLT_MAX, mxy = -FLT_MAX;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float a = x[i], b = y[i];
    if (isfinite(a)) {
      mnx = fminf(mnx, a);
      mxx = fmaxf(mxx, a);
    }
    if (isfinite(b)) {
      mny = fminf(mny, b);
      mxy = fmaxf(mxy, b);
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(FULL, mnx, m));
    mxx = fmaxf(mxx, __shfl_xor_sync(FULL, mxx, m));
    mny = fminf(mny, __shfl_xor_sync(FULL, mny, m));
    mxy = fmaxf(mxy, __shfl_xor_sync(FULL, mxy, m));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&out[0], float_to_ordered(mnx));
    atomicMax(&out[1], float_to_ordered(mxx));
    atomicMin(&out[2], float_to_ordered(mny));
    atomicMax(&out[3], float_to_ordered(mxy));
  }
}

}  // namespace kc
