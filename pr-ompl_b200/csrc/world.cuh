// K2 device-side world model: sphere-approximated FK of a DH arm vs sphere/box obstacles.
// Specification: the double-precision CPU definition the tests check against
// (clearance(q) = min over (link sphere, obstacle) pairs, valid iff > 0).  This file restates it
// for the device, with an fp32 conservative filter in front of the fp64 decision:
//   level 1  per obstacle: min over link spheres of squared centre distance vs (R + r_max + delta)^2
//   level 2  per (sphere, obstacle): fp32 test with the sphere's own radius (+delta)
//   level 3  fp64 clearance with the specification's formula -> the only place a verdict is made
// delta bounds the fp32 rounding of centres and arithmetic, so levels 1-2 never reject a colliding
// pair; every "collision" verdict comes from level 3.
#pragma once
#include <cuda_fp16.h>
#include <math_constants.h>
#include "common.cuh"

namespace prgpu {

constexpr int WORLD_MAX_DOF = PRGPU_MAX_DIM;
constexpr int WORLD_MAX_S = PRGPU_MAX_LINK_SPHERES;

struct ArmTableF32 { // fp32 copy of the arm for the filter path (built once by prgpu_world_create; kernels copy it to shared memory)
  float dh[WORLD_MAX_DOF][5];        // a, d, cos(alpha), sin(alpha), theta0
  float4 sphere[WORLD_MAX_S];        // local x,y,z and radius + delta, ordered by link
  uint8_t sphere_id[WORLD_MAX_S];    // original sphere index of slot j
  uint8_t link_first[WORLD_MAX_DOF + 2]; // spheres of frame i are slots [link_first[i], link_first[i+1])
};
struct DeviceWorld {
  int dof, S, n_os, n_ob;
  double dh[WORLD_MAX_DOF][5];       // a, d, cos(alpha), sin(alpha), theta0
  int sphere_link[WORLD_MAX_S];
  double sphere_local[WORLD_MAX_S][4]; // x,y,z,r  (padding spheres: far away, r = 0)
  float sphere_r[WORLD_MAX_S];       // fp32 radius + delta
  float box_thr2;                    // (r_max + delta)^2
  float delta;
  const float4 *os_f;                // n_os x (x,y,z,(R + r_max + delta)^2)
  const float4 *ob_f;                // n_ob x 2: (cx,cy,cz,_) (hx,hy,hz,_)
  const double *os_d;                // n_os x 4
  const double *ob_d;                // n_ob x 6
  // broad phase: uniform grid over the obstacles.  An obstacle is listed in every cell its bounding box
  // inflated by (r_max + delta + GRID_SLACK) overlaps, so the cell that contains a link-sphere centre
  // lists every obstacle that sphere could touch.  grid_n[0] == 0: no grid (brute-force filter).
  int grid_n[3];
  float grid_org[3], grid_inv;
  uint32_t grid_items;               // total list length
  const uint32_t *cell_start;        // ncell+1
  const uint16_t *cell_item;         // obstacle index, bit 15 set for boxes
  // clearance field: df[cell] <= distance from ANY point of the (fine) cell to the nearest obstacle,
  // minus the fp32 position error and a cell-index slack.  A link sphere whose centre lies in a cell
  // with df > radius is certainly free (one load instead of the cell's obstacle list), and
  // min_j(df - r_j) is a lower bound of the state's clearance: the states of an edge within that
  // distance (in workspace motion, bounded through rho[]) of a tested state are free without being
  // tested.  df == nullptr: no field.  Outside the field's box every obstacle is >= df_out away.
  // Two-sided: a cell's 32-bit word packs two fp16 numbers (df_unpack): .x, the lower bound above (rounded down),
  // and .y >= the distance from ANY point of the cell to the nearest obstacle surface plus the same slacks
  // (rounded up) -- a link sphere whose true radius exceeds .y certainly penetrates an obstacle, so the state
  // is invalid without walking a list or running the fp64 chain (the specification's verdict is clearance <= 0).
  const uint32_t *df;
  int df_n[3];
  float df_org[3], df_inv, df_out;
  // rho[i]: upper bound, over all configurations and link spheres, of the distance between a sphere centre
  // and the axis of joint i -- a unit change of joint i moves no sphere centre by more than rho[i]
  float rho[WORLD_MAX_DOF];
  ArmTableF32 arm_f32;               // the fp32 arm table the filter kernels keep in shared memory
};
constexpr int WORLD_GRID_MAX_CELLS = 8192;
constexpr int WORLD_GRID_MAX_ITEMS = 32768;

// frames F[(i)*12 ..]: 3x4 row-major [R|p] of frame i = 0..dof
__device__ __forceinline__ void fk_frames(const DeviceWorld &w, const double *q, double *F) {
  F[0] = 1; F[1] = 0; F[2] = 0; F[3] = 0;
  F[4] = 0; F[5] = 1; F[6] = 0; F[7] = 0;
  F[8] = 0; F[9] = 0; F[10] = 1; F[11] = 0;
  for (int i = 0; i < w.dof; ++i) {
    double st, ct;
    sincos(q[i] + w.dh[i][4], &st, &ct);
    const double a = w.dh[i][0], d = w.dh[i][1], cA = w.dh[i][2], sA = w.dh[i][3];
    const double L0[4] = {ct, -st * cA, st * sA, a * ct};
    const double L1[4] = {st, ct * cA, -ct * sA, a * st};
    const double L2[4] = {0.0, sA, cA, d};
    const double *P = F + 12 * i;
    double *N = F + 12 * (i + 1);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double p0 = P[4 * r + 0], p1 = P[4 * r + 1], p2 = P[4 * r + 2], p3 = P[4 * r + 3];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double v = p0 * L0[c] + p1 * L1[c] + p2 * L2[c];
        if (c == 3)
          v += p3;
        N[4 * r + c] = v;
      }
    }
  }
}

__device__ __forceinline__ void sphere_centre(const DeviceWorld &w, const double *F, int s, double c[3]) {
  const double *T = F + 12 * w.sphere_link[s];
  const double lx = w.sphere_local[s][0], ly = w.sphere_local[s][1], lz = w.sphere_local[s][2];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    c[r] = T[4 * r + 0] * lx + T[4 * r + 1] * ly + T[4 * r + 2] * lz + T[4 * r + 3];
}

// level 3: specification formula in double; true = this pair collides (clearance <= 0)
static __device__ __noinline__ bool exact_sphere_hit(const DeviceWorld &w, const double *F, int s, int o) {
  double c[3];
  sphere_centre(w, F, s, c);
  const double *ob = w.os_d + 4 * o;
  const double dx = c[0] - ob[0], dy = c[1] - ob[1], dz = c[2] - ob[2];
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  const double cl = __dsub_rn(__dsqrt_rn(d2), __dadd_rn(w.sphere_local[s][3], ob[3]));
  return !(cl > 0.0);
}
static __device__ __noinline__ bool exact_box_hit(const DeviceWorld &w, const double *F, int s, int o) {
  double c[3];
  sphere_centre(w, F, s, c);
  const double *ob = w.ob_d + 6 * o;
  double ex = fabs(c[0] - ob[0]) - ob[3];
  double ey = fabs(c[1] - ob[1]) - ob[4];
  double ez = fabs(c[2] - ob[2]) - ob[5];
  ex = ex > 0.0 ? ex : 0.0;
  ey = ey > 0.0 ? ey : 0.0;
  ez = ez > 0.0 ? ez : 0.0;
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
  const double cl = __dsub_rn(__dsqrt_rn(d2), w.sphere_local[s][3]);
  return !(cl > 0.0);
}

// StateValidityChecker::isValid for one configuration.  s_os / s_ob: fp32 obstacle tables in
// shared memory.  SP = number of link spheres padded to a multiple of 8 (compile time).
template <int SP>
__device__ __forceinline__ bool state_valid(const DeviceWorld &w, const float4 *__restrict__ s_os,
                                            const float4 *__restrict__ s_ob, const double *q) {
  double F[12 * (WORLD_MAX_DOF + 1)];
  fk_frames(w, q, F);
  float cx[SP], cy[SP], cz[SP];
#pragma unroll
  for (int s = 0; s < SP; ++s) {
    double c[3];
    sphere_centre(w, F, s, c);
    cx[s] = (float)c[0];
    cy[s] = (float)c[1];
    cz[s] = (float)c[2];
  }
  // sphere obstacles
  for (int o = 0; o < w.n_os; ++o) {
    const float4 ob = s_os[o];
    float t = CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < SP; ++s) {
      const float dx = cx[s] - ob.x, dy = cy[s] - ob.y, dz = cz[s] - ob.z;
      t = fminf(t, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    }
    if (t < ob.w) {
      const float R = (float)w.os_d[4 * o + 3];
      unsigned hits = 0;
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        const float dx = cx[s] - ob.x, dy = cy[s] - ob.y, dz = cz[s] - ob.z;
        const float thr = w.sphere_r[s] + R;
        if (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) < thr * thr)
          hits |= 1u << s;
      }
      while (hits) {
        const int s = __ffs(hits) - 1;
        hits &= hits - 1;
        if (s < w.S && exact_sphere_hit(w, F, s, o))
          return false;
      }
    }
  }
  // box obstacles
  for (int o = 0; o < w.n_ob; ++o) {
    const float4 bc = s_ob[2 * o], bh = s_ob[2 * o + 1];
    float t = CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < SP; ++s) {
      const float ex = fmaxf(fabsf(cx[s] - bc.x) - bh.x, 0.0f);
      const float ey = fmaxf(fabsf(cy[s] - bc.y) - bh.y, 0.0f);
      const float ez = fmaxf(fabsf(cz[s] - bc.z) - bh.z, 0.0f);
      t = fminf(t, fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
    }
    if (t < w.box_thr2) {
      unsigned hits = 0;
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        const float ex = fmaxf(fabsf(cx[s] - bc.x) - bh.x, 0.0f);
        const float ey = fmaxf(fabsf(cy[s] - bc.y) - bh.y, 0.0f);
        const float ez = fmaxf(fabsf(cz[s] - bc.z) - bh.z, 0.0f);
        const float thr = w.sphere_r[s];
        if (fmaf(ez, ez, fmaf(ey, ey, ex * ex)) < thr * thr)
          hits |= 1u << s;
      }
      while (hits) {
        const int s = __ffs(hits) - 1;
        hits &= hits - 1;
        if (s < w.S && exact_box_hit(w, F, s, o))
          return false;
      }
    }
  }
  return true;
}

// ---- fast path: fp32 FK in front of the same 3-level test ---------------------------------------
// The link-sphere centres only feed the conservative fp32 filter (levels 1-2), so they can come from an
// fp32 forward-kinematics chain; delta (world_create) bounds its error together with the fp32 rounding
// of the obstacle tables.  A level-2 hit recomputes that sphere's centre with the fp64 chain and applies
// the specification's fp64 formula, so every verdict is still made in double.
__device__ __forceinline__ void fk_frames_f32(const DeviceWorld &w, const double *q, float *F) {
  F[0] = 1.f; F[1] = 0.f; F[2] = 0.f; F[3] = 0.f;
  F[4] = 0.f; F[5] = 1.f; F[6] = 0.f; F[7] = 0.f;
  F[8] = 0.f; F[9] = 0.f; F[10] = 1.f; F[11] = 0.f;
  for (int i = 0; i < w.dof; ++i) {
    float st, ct;
    sincosf((float)(q[i] + w.dh[i][4]), &st, &ct);
    const float a = (float)w.dh[i][0], d = (float)w.dh[i][1], cA = (float)w.dh[i][2], sA = (float)w.dh[i][3];
    const float L0[4] = {ct, -st * cA, st * sA, a * ct};
    const float L1[4] = {st, ct * cA, -ct * sA, a * st};
    const float L2[4] = {0.f, sA, cA, d};
    const float *P = F + 12 * i;
    float *N = F + 12 * (i + 1);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float p0 = P[4 * r + 0], p1 = P[4 * r + 1], p2 = P[4 * r + 2], p3 = P[4 * r + 3];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v = p0 * L0[c] + p1 * L1[c] + p2 * L2[c];
        if (c == 3)
          v += p3;
        N[4 * r + c] = v;
      }
    }
  }
}

// fp64 centre of one link sphere, chaining only up to its link (level 3 of the fast path)
static __device__ __noinline__ void sphere_centre_exact(const DeviceWorld &w, const double *q, int s, double c[3]) {
  double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  const int link = w.sphere_link[s];
  for (int i = 0; i < link; ++i) {
    double st, ct;
    sincos(q[i] + w.dh[i][4], &st, &ct);
    const double a = w.dh[i][0], d = w.dh[i][1], cA = w.dh[i][2], sA = w.dh[i][3];
    const double L0[4] = {ct, -st * cA, st * sA, a * ct};
    const double L1[4] = {st, ct * cA, -ct * sA, a * st};
    const double L2[4] = {0.0, sA, cA, d};
    double N[12];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double p0 = T[4 * r + 0], p1 = T[4 * r + 1], p2 = T[4 * r + 2], p3 = T[4 * r + 3];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        double v = p0 * L0[cc] + p1 * L1[cc] + p2 * L2[cc];
        if (cc == 3)
          v += p3;
        N[4 * r + cc] = v;
      }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k)
      T[k] = N[k];
  }
  const double lx = w.sphere_local[s][0], ly = w.sphere_local[s][1], lz = w.sphere_local[s][2];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    c[r] = T[4 * r + 0] * lx + T[4 * r + 1] * ly + T[4 * r + 2] * lz + T[4 * r + 3];
}
static __device__ __noinline__ bool exact_pair_hit(const DeviceWorld &w, const double *q, int s, int o, bool box) {
  double c[3];
  sphere_centre_exact(w, q, s, c);
  double d2;
  double rr;
  if (!box) {
    const double *ob = w.os_d + 4 * o;
    const double dx = c[0] - ob[0], dy = c[1] - ob[1], dz = c[2] - ob[2];
    d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    rr = __dadd_rn(w.sphere_local[s][3], ob[3]);
  } else {
    const double *ob = w.ob_d + 6 * o;
    double ex = fabs(c[0] - ob[0]) - ob[3];
    double ey = fabs(c[1] - ob[1]) - ob[4];
    double ez = fabs(c[2] - ob[2]) - ob[5];
    ex = ex > 0.0 ? ex : 0.0;
    ey = ey > 0.0 ? ey : 0.0;
    ez = ez > 0.0 ? ez : 0.0;
    d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
    rr = w.sphere_local[s][3];
  }
  const double cl = __dsub_rn(__dsqrt_rn(d2), rr);
  return !(cl > 0.0);
}

template <int SP>
__device__ __forceinline__ bool state_valid_fast(const DeviceWorld &w, const float4 *__restrict__ s_os,
                                                 const float4 *__restrict__ s_ob, const double *q) {
  float F[12 * (WORLD_MAX_DOF + 1)];
  fk_frames_f32(w, q, F);
  float cx[SP], cy[SP], cz[SP];
#pragma unroll
  for (int s = 0; s < SP; ++s) {
    const float *T = F + 12 * w.sphere_link[s];
    const float lx = (float)w.sphere_local[s][0], ly = (float)w.sphere_local[s][1], lz = (float)w.sphere_local[s][2];
    cx[s] = T[0] * lx + T[1] * ly + T[2] * lz + T[3];
    cy[s] = T[4] * lx + T[5] * ly + T[6] * lz + T[7];
    cz[s] = T[8] * lx + T[9] * ly + T[10] * lz + T[11];
  }
  for (int o = 0; o < w.n_os; ++o) {
    const float4 ob = s_os[o];
    float t = CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < SP; ++s) {
      const float dx = cx[s] - ob.x, dy = cy[s] - ob.y, dz = cz[s] - ob.z;
      t = fminf(t, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
    }
    if (t < ob.w) {
      const float R = (float)w.os_d[4 * o + 3];
      unsigned hits = 0;
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        const float dx = cx[s] - ob.x, dy = cy[s] - ob.y, dz = cz[s] - ob.z;
        const float thr = w.sphere_r[s] + R;
        if (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) < thr * thr)
          hits |= 1u << s;
      }
      while (hits) {
        const int s = __ffs(hits) - 1;
        hits &= hits - 1;
        if (s < w.S && exact_pair_hit(w, q, s, o, false))
          return false;
      }
    }
  }
  for (int o = 0; o < w.n_ob; ++o) {
    const float4 bc = s_ob[2 * o], bh = s_ob[2 * o + 1];
    float t = CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < SP; ++s) {
      const float ex = fmaxf(fabsf(cx[s] - bc.x) - bh.x, 0.0f);
      const float ey = fmaxf(fabsf(cy[s] - bc.y) - bh.y, 0.0f);
      const float ez = fmaxf(fabsf(cz[s] - bc.z) - bh.z, 0.0f);
      t = fminf(t, fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
    }
    if (t < w.box_thr2) {
      unsigned hits = 0;
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        const float ex = fmaxf(fabsf(cx[s] - bc.x) - bh.x, 0.0f);
        const float ey = fmaxf(fabsf(cy[s] - bc.y) - bh.y, 0.0f);
        const float ez = fmaxf(fabsf(cz[s] - bc.z) - bh.z, 0.0f);
        const float thr = w.sphere_r[s];
        if (fmaf(ez, ez, fmaf(ey, ey, ex * ex)) < thr * thr)
          hits |= 1u << s;
      }
      while (hits) {
        const int s = __ffs(hits) - 1;
        hits &= hits - 1;
        if (s < w.S && exact_pair_hit(w, q, s, o, true))
          return false;
      }
    }
  }
  return true;
}

// cooperative copy of the fp32 obstacle tables into shared memory
__device__ __forceinline__ void load_obstacles(const DeviceWorld &w, float4 *s_os, float4 *s_ob) {
  for (int i = threadIdx.x; i < w.n_os; i += blockDim.x)
    s_os[i] = w.os_f[i];
  for (int i = threadIdx.x; i < 2 * w.n_ob; i += blockDim.x)
    s_ob[i] = w.ob_f[i];
}

// ---- grid path: per link sphere, only the obstacles listed in the sphere centre's cell ------------
// delta bounds |fp32 distance - exact distance| of a (sphere, obstacle) pair.  Three outcomes per
// candidate pair: farther than r+R+delta -> free; nearer than r+R-delta -> colliding (both certain in
// fp32); in between -> the fp64 chain and the specification's formula decide.
struct ObstacleTables { // shared-memory views
  const float4 *os, *ob;
  const uint32_t *cell_start;
  const uint16_t *cell_item;
  const ArmTableF32 *arm;
};
__host__ __device__ __forceinline__ size_t obstacle_tables_bytes(const DeviceWorld &w) {
  size_t b = ((size_t)w.n_os + 2 * (size_t)w.n_ob) * sizeof(float4) + ((sizeof(ArmTableF32) + 15) & ~(size_t)15);
  if (w.grid_n[0] > 0)
    b += ((size_t)w.grid_n[0] * w.grid_n[1] * w.grid_n[2] + 1) * sizeof(uint32_t) + (((size_t)w.grid_items + 7) & ~(size_t)7) * sizeof(uint16_t);
  return b;
}
// cooperative load of all tables; returns the views (call from every thread, then __syncthreads)
// `stage` = false (tiny launches: a planner's single checkMotion / isValid): only the arm table goes to shared
// memory, the obstacle tables and the grid are read where they lie -- copying up to 100 KB per block would cost
// several times the few look-ups such a launch makes.
__device__ __forceinline__ ObstacleTables load_tables(const DeviceWorld &w, float4 *smem, bool stage = true) {
  ObstacleTables t;
  float4 *s_os = smem, *s_ob = smem + w.n_os;
  t.os = w.os_f;
  t.ob = w.ob_f;
  if (stage) {
    load_obstacles(w, s_os, s_ob);
    t.os = s_os;
    t.ob = s_ob;
  }
  t.cell_start = nullptr;
  t.cell_item = nullptr;
  ArmTableF32 *arm = reinterpret_cast<ArmTableF32 *>(smem + w.n_os + 2 * w.n_ob);
  t.arm = arm;
  { // (spheres grouped by the frame they ride on: prepared by prgpu_world_create)
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&w.arm_f32);
    uint32_t *dst = reinterpret_cast<uint32_t *>(arm);
    for (uint32_t i = threadIdx.x; i < sizeof(ArmTableF32) / 4; i += blockDim.x)
      dst[i] = src[i];
  }
  if (w.grid_n[0] > 0) {
    t.cell_start = w.cell_start;
    t.cell_item = w.cell_item;
    if (stage) {
      const uint32_t ncell = (uint32_t)(w.grid_n[0] * w.grid_n[1] * w.grid_n[2]);
      uint32_t *cs = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(arm) +
                                                  ((sizeof(ArmTableF32) + 15) & ~(size_t)15));
      uint16_t *ci = reinterpret_cast<uint16_t *>(cs + ncell + 1);
      for (uint32_t i = threadIdx.x; i <= ncell; i += blockDim.x)
        cs[i] = w.cell_start[i];
      for (uint32_t i = threadIdx.x; i < w.grid_items; i += blockDim.x)
        ci[i] = w.cell_item[i];
      t.cell_start = cs;
      t.cell_item = ci;
    }
  }
  return t;
}

// fp32 frame [R|p] of the FK chain, in registers
struct FrameF32 {
  float r00, r01, r02, p0, r10, r11, r12, p1, r20, r21, r22, p2;
  __device__ __forceinline__ void reset() {
    r00 = 1.f; r01 = 0.f; r02 = 0.f; p0 = 0.f;
    r10 = 0.f; r11 = 1.f; r12 = 0.f; p1 = 0.f;
    r20 = 0.f; r21 = 0.f; r22 = 1.f; p2 = 0.f;
  }
  // frame i-1 -> frame i:  T <- T * [Rz(theta) Tz(d) Tx(a) Rx(alpha)]
  __device__ __forceinline__ void advance(const ArmTableF32 &arm, const double *q, int i) {
    float st, ct;
    sincosf((float)q[i - 1] + arm.dh[i - 1][4], &st, &ct);
    advance(arm, i, st, ct);
  }
  __device__ __forceinline__ void advance(const ArmTableF32 &arm, int i, float st, float ct) {
    const float a = arm.dh[i - 1][0], d = arm.dh[i - 1][1], cA = arm.dh[i - 1][2], sA = arm.dh[i - 1][3];
    const float l01 = -st * cA, l02 = st * sA, l03 = a * ct;
    const float l11 = ct * cA, l12 = -ct * sA, l13 = a * st;
    float n0, n1, n2, n3;
    n0 = r00 * ct + r01 * st; n1 = r00 * l01 + r01 * l11 + r02 * sA; n2 = r00 * l02 + r01 * l12 + r02 * cA;
    n3 = r00 * l03 + r01 * l13 + r02 * d + p0;
    r00 = n0; r01 = n1; r02 = n2; p0 = n3;
    n0 = r10 * ct + r11 * st; n1 = r10 * l01 + r11 * l11 + r12 * sA; n2 = r10 * l02 + r11 * l12 + r12 * cA;
    n3 = r10 * l03 + r11 * l13 + r12 * d + p1;
    r10 = n0; r11 = n1; r12 = n2; p1 = n3;
    n0 = r20 * ct + r21 * st; n1 = r20 * l01 + r21 * l11 + r22 * sA; n2 = r20 * l02 + r21 * l12 + r22 * cA;
    n3 = r20 * l03 + r21 * l13 + r22 * d + p2;
    r20 = n0; r21 = n1; r22 = n2; p2 = n3;
  }
  __device__ __forceinline__ void centre(const float4 &sp, float &cx, float &cy, float &cz) const {
    cx = r00 * sp.x + r01 * sp.y + r02 * sp.z + p0;
    cy = r10 * sp.x + r11 * sp.y + r12 * sp.z + p1;
    cz = r20 * sp.x + r21 * sp.y + r22 * sp.z + p2;
  }
};

__device__ __forceinline__ float2 df_unpack(uint32_t word) {
  return __half22float2(*reinterpret_cast<const __half2 *>(&word));
}

// Pass 1 (clearance field): fp32 FK frame by frame; the field loads of a link's spheres are issued in
// groups of 4 and consumed only after the NEXT frame has been computed, so the L2 latency of the field
// overlaps the sincos / frame product; no per-sphere branch.  Returns the mask of the sphere slots the
// field could not certify (0: the state is free, and *cmin is a lower bound of its clearance).
// Requires w.df != nullptr.
// `centres` (optional, 3 floats per sphere slot): receives every sphere centre, so that the second pass of the
// same thread need not repeat the FK chain.
// `hit_out` (optional): set to true when some sphere CERTAINLY collides (the field's upper bound lies below its
// true radius): the state is invalid, whatever the returned mask says.
// Only the centres of the spheres the field could NOT certify are stored (they are the ones a list walk needs);
// ALL_CENTRES stores every one (measurement aid).
template <bool ALL_CENTRES = false>
__device__ __forceinline__ uint32_t grid_certify(const DeviceWorld &w, const ArmTableF32 &arm, const double *q,
                                                 float *cmin_out, float *centres = nullptr, bool *hit_out = nullptr) {
  const uint32_t *df = w.df;
  const float two_delta = 2.0f * w.delta;
  float hmin = CUDART_INF_F; // min over spheres of (upper bound - true radius): < 0 -> certain collision
  const int dnx = w.df_n[0], dny = w.df_n[1], dnz = w.df_n[2];
  const float dox = w.df_org[0], doy = w.df_org[1], doz = w.df_org[2], dinv = w.df_inv, dout = w.df_out;
  float cmin = CUDART_INF_F;
  uint32_t unc = 0;
  float dv0 = CUDART_INF_F, dv1 = CUDART_INF_F, dv2 = CUDART_INF_F, dv3 = CUDART_INF_F; // field - radius, in flight
  float3 cc[4];                                                                       // their centres
  int jb = 0;                                                                         // slot of dv0
  FrameF32 f;
  auto consume = [&]() {
    cmin = fminf(fminf(cmin, dv0), fminf(dv1, fminf(dv2, dv3)));
    const uint32_t u4 = (dv0 > 0.f ? 0u : 1u) | (dv1 > 0.f ? 0u : 2u) | (dv2 > 0.f ? 0u : 4u) | (dv3 > 0.f ? 0u : 8u);
    unc |= u4 << jb;
    if (!ALL_CENTRES && centres && u4) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (u4 & (1u << k)) {
          centres[3 * (jb + k)] = cc[k].x;
          centres[3 * (jb + k) + 1] = cc[k].y;
          centres[3 * (jb + k) + 2] = cc[k].z;
        }
    }
  };
  auto lookup = [&](int j, int k) -> float {
    const float4 sp = arm.sphere[j];
    float cx, cy, cz;
    f.centre(sp, cx, cy, cz);
    cc[k] = make_float3(cx, cy, cz);
    if (ALL_CENTRES && centres) {
      centres[3 * j] = cx;
      centres[3 * j + 1] = cy;
      centres[3 * j + 2] = cz;
    }
    const float fx = (cx - dox) * dinv, fy = (cy - doy) * dinv, fz = (cz - doz) * dinv;
    float2 dv = make_float2(dout, CUDART_INF_F); // outside the field's box (or NaN)
    if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)dnx && fy < (float)dny && fz < (float)dnz)
      dv = df_unpack(__ldg(df + ((size_t)(int)fz * dny + (int)fy) * dnx + (int)fx));
    hmin = fminf(hmin, dv.y - (sp.w - two_delta)); // sp.w = radius + delta
    return dv.x - sp.w;
  };
  f.reset();
  for (int i = 0; i <= w.dof; ++i) {
    if (i > 0)
      f.advance(arm, q, i);
    int j = arm.link_first[i];
    const int je = arm.link_first[i + 1];
    while (j < je) {
      consume(); // the previous group: its loads had a frame product (or a group) to land
      jb = j;
      dv0 = lookup(j, 0);
      dv1 = j + 1 < je ? lookup(j + 1, 1) : CUDART_INF_F;
      dv2 = j + 2 < je ? lookup(j + 2, 2) : CUDART_INF_F;
      dv3 = j + 3 < je ? lookup(j + 3, 3) : CUDART_INF_F;
      j += 4;
    }
  }
  consume();
  *cmin_out = cmin < CUDART_INF_F ? cmin : 0.f;
  if (hit_out)
    *hit_out = hmin < 0.f;
  return unc;
}

// Pass 2: the sphere slots in `unc` walk the obstacle list of their centre's cell.
// delta bounds |fp32 distance - exact distance| of a (sphere, obstacle) pair; three outcomes per
// candidate pair: farther than r+R+delta -> free; nearer than r+R-delta -> colliding (both certain in
// fp32); in between -> the fp64 chain and the specification's formula decide.
// false: sphere slot j (centre cx,cy,cz) certainly collides or the fp64 verdict says so
// `state()` returns the fp64 joint values of the sphere's state; it is called only for a thin-band pair
template <class StateFn>
__device__ __forceinline__ bool grid_sphere_free(const DeviceWorld &w, const ObstacleTables &t, StateFn &&state, int j,
                                                 float cx, float cy, float cz) {
  const ArmTableF32 &arm = *t.arm;
  const int nx = w.grid_n[0], ny = w.grid_n[1], nz = w.grid_n[2];
  const float gx = (cx - w.grid_org[0]) * w.grid_inv, gy = (cy - w.grid_org[1]) * w.grid_inv,
              gz = (cz - w.grid_org[2]) * w.grid_inv;
  // outside the grid (or NaN): no inflated obstacle box contains the centre
  if (!(gx >= 0.f && gy >= 0.f && gz >= 0.f && gx < (float)nx && gy < (float)ny && gz < (float)nz))
    return true;
  const float rs = arm.sphere[j].w;
  const int cell = ((int)gz * ny + (int)gy) * nx + (int)gx;
  for (uint32_t it = t.cell_start[cell]; it < t.cell_start[cell + 1]; ++it) {
    const uint32_t id = t.cell_item[it];
    if (id & 0x8000u) {
      const uint32_t o = id & 0x7FFFu;
      const float4 bc = t.ob[2 * o], bh = t.ob[2 * o + 1];
      const float ex = fmaxf(fabsf(cx - bc.x) - bh.x, 0.0f);
      const float ey = fmaxf(fabsf(cy - bc.y) - bh.y, 0.0f);
      const float ez = fmaxf(fabsf(cz - bc.z) - bh.z, 0.0f);
      const float d2 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
      if (d2 < rs * rs) {
        // deeper than the fp32 error bound: certainly colliding; only the thin band needs fp64
        const float rin = rs - 2.0f * w.delta;
        if ((rin > 0.f && d2 < rin * rin) || exact_pair_hit(w, state(), arm.sphere_id[j], (int)o, true))
          return false;
      }
    } else {
      const float4 ob = t.os[id];
      const float dx = cx - ob.x, dy = cy - ob.y, dz = cz - ob.z;
      const float thr = rs + ob.w; // grid tables keep the obstacle radius in .w
      const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      if (d2 < thr * thr) {
        const float tin = thr - 2.0f * w.delta;
        if ((tin > 0.f && d2 < tin * tin) || exact_pair_hit(w, state(), arm.sphere_id[j], (int)id, false))
          return false;
      }
    }
  }
  return true;
}
// with the FK chain repeated (the centres were not kept: another kernel certified the state)
__device__ __forceinline__ bool grid_lists(const DeviceWorld &w, const ObstacleTables &t, const double *q, uint32_t unc) {
  const ArmTableF32 &arm = *t.arm;
  FrameF32 f;
  f.reset();
  for (int i = 0; i <= w.dof; ++i) {
    if (i > 0)
      f.advance(arm, q, i);
    if (arm.link_first[i] >= 32 || (unc >> arm.link_first[i]) == 0)
      break; // no uncertified sphere on this or any later frame
    for (int j = arm.link_first[i]; j < arm.link_first[i + 1]; ++j) {
      if (!((unc >> j) & 1u))
        continue;
      float cx, cy, cz;
      f.centre(arm.sphere[j], cx, cy, cz);
      if (!grid_sphere_free(w, t, [&]() { return q; }, j, cx, cy, cz))
        return false;
    }
  }
  return true;
}
// with the centres the first pass left behind
__device__ __forceinline__ bool grid_lists_kept(const DeviceWorld &w, const ObstacleTables &t, const double *q, uint32_t unc,
                                                const float *centres) {
  while (unc) {
    const int j = __ffs(unc) - 1;
    unc &= unc - 1;
    if (!grid_sphere_free(w, t, [&]() { return q; }, j, centres[3 * j], centres[3 * j + 1], centres[3 * j + 2]))
      return false;
  }
  return true;
}

// `clear` (optional): receives a lower bound of the state's clearance (metres) when every link sphere was
// certified by the clearance field, else 0.
static __device__ __noinline__ bool state_valid_grid(const DeviceWorld *wp, const ObstacleTables *tp, const double *q,
                                                     float *clear = nullptr) {
  const DeviceWorld &w = *wp;
  if (clear)
    *clear = 0.f;
  if (!w.df)
    return grid_lists(w, *tp, q, 0xFFFFFFFFu);
  float cmin, centres[3 * WORLD_MAX_S];
  bool hit;
  const uint32_t unc = grid_certify(w, *tp->arm, q, &cmin, centres, &hit);
  if (hit)
    return false;
  if (unc == 0) {
    if (clear)
      *clear = cmin;
    return true;
  }
  return grid_lists_kept(w, *tp, q, unc, centres);
}
// the two passes as separate out-of-line calls (the sample / resolve kernels of check_motion.cu)
static __device__ __noinline__ uint32_t state_certify_grid(const DeviceWorld *wp, const ObstacleTables *tp, const double *q,
                                                           float *cmin) {
  return grid_certify(*wp, *tp->arm, q, cmin);
}
static __device__ __noinline__ bool state_lists_grid(const DeviceWorld *wp, const ObstacleTables *tp, const double *q,
                                                     uint32_t unc) {
  return grid_lists(*wp, *tp, q, unc);
}

// validity of one state through whichever broad phase the world has
// SP == 0 selects the grid path at compile time, so kernels instantiated for worlds with a grid never
// carry the brute-force filter's register footprint (the host dispatches on grid_n[0] > 0).
template <int SP>
__device__ __forceinline__ bool state_valid_any(const DeviceWorld *wp, const ObstacleTables *tp, const double *q,
                                                float *clear = nullptr) {
  if constexpr (SP == 0)
    return state_valid_grid(wp, tp, q, clear);
  else {
    if (clear)
      *clear = 0.f;
    return state_valid_fast<SP>(*wp, tp->os, tp->ob, q);
  }
}

// out-of-line variant for kernels that test states from several places (keeps their register
// footprint at the checker's own)
template <int SP>
__device__ __noinline__ bool state_valid_call(const DeviceWorld *wp, const ObstacleTables *tp, const double *q) {
  return state_valid_any<SP>(wp, tp, q);
}

} // namespace prgpu
