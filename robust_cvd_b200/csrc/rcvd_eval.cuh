// rcvd_eval.cuh -- residual / Jacobian / normal-equation accumulation kernels.
//
// K1 gn_accumulate : per flow constraint, StaticSceneCost residual + analytic Jacobian
//                    + Cauchy correction + scatter of J^T J and J^T r   (reference hot loop D,
//                    lib/PoseOptimizer.cpp:237-308 evaluated by Ceres autodiff at :1198)
// K2 cost_only     : residual + rho only, for LM step acceptance
// K3 regularisers  : scale / deform / spatial / focal / position rows (:488-656, :1341-1549)
#pragma once
#include "rcvd_device.cuh"

namespace rcvd {

constexpr int kTile = 128;          // constraints per CTA tile (all from one directed pair)
constexpr int kMaxEntries = 144;    // 2 * (6 + 1 + 16*2 + 16*2) + slack

struct DevProblem {
  rcvd_config cfg;
  Layout L;
  int N;
  int num_tiles;
  int64_t num_constraints;
  const float* records;        // [C][6]
  const int32_t* tile_pair;    // [T]
  const int64_t* tile_begin;   // [T]
  const int32_t* tile_count;   // [T]
  const int32_t* pair_frames;  // [P][2]
  const int32_t* blk_of;       // [N*N]: (fa,fb) -> H block id*2 + (fa is the row side), -1 if absent
  const uint8_t* in_range;     // [N]
  const double* median;        // [N]
  const double* adaptive;      // [N*G] or null
  const float* scale_locs;     // [M][2]
  int num_scale_locs;
  int rank, nranks;            // regulariser rows are evaluated by rank f % nranks
  // scene-flow smoothness triplets (optional): records [n][10], tiles of <= kTile constraints with one centre frame
  const float* trip_records; const int32_t* trip_tile_center; const int64_t* trip_tile_begin; const int32_t* trip_tile_count; int num_trip_tiles;
};

__device__ __forceinline__ bool is_const_local(const rcvd_config& c, const Layout& L, int l) {
  if (l < 6) return c.fix_poses != 0;
  if (l == 6) return c.intr_opt == RCVD_INTR_FIXED;
  if (l < L.offS) return c.fix_depth_xforms != 0;
  return c.fix_spatial_xforms != 0;
}

// Adds v to H(fa:la, fb:lb).  Diagonal blocks store the lower triangle only.
__device__ __forceinline__ void add_h(const DevProblem& p, double* H, int fa, int la, int fb, int lb, double v) {
  const int np = p.L.npad;
  if (fa == fb) {
    const int i = max(la, lb), j = min(la, lb);
    red_add(H + (size_t)fa * np * np + (size_t)i * np + j, v);
  } else {
    const int enc = p.blk_of[fa * p.N + fb];
    const size_t base = (size_t)(enc >> 1) * np * np;
    if (enc & 1) red_add(H + base + (size_t)la * np + lb, v);
    else red_add(H + base + (size_t)lb * np + la, v);
  }
}

struct StaticEval {
  double r[3];
  double rho0, scale;
};

// Shared front end of K1/K2: loads the record, gathers, evaluates residual (+ local Jacobian).
template <bool JAC>
__device__ __forceinline__ void eval_constraint(const DevProblem& p, const double* __restrict__ x, int f0, int f1,
                                                const float* __restrict__ rec, Gather& dg0, Gather& sg0, Gather& dg1, Gather& sg1,
                                                StaticEval& ev, double* Jl) {
  const rcvd_config& c = p.cfg; const Layout& L = p.L;
  const double* pf0 = x + (size_t)f0 * L.nf; const double* pf1 = x + (size_t)f1 * L.nf;
  ObsIn o0{rec[0], rec[1], rec[2]}, o1{rec[3], rec[4], rec[5]};
  gather_depth(c, o0.ndcx, o0.ndcy, dg0); gather_depth(c, o1.ndcx, o1.ndcy, dg1);
  gather_spatial(c, o0.ndcx, o0.ndcy, sg0); gather_spatial(c, o1.ndcx, o1.ndcy, sg1);
  double phi0, phi1;
  if (c.intr_opt == RCVD_INTR_SHARED) phi0 = phi1 = x[6];          // &poseParams_[0][6], lib/PoseOptimizer.cpp:1226
  else if (c.intr_opt == RCVD_INTR_PER_FRAME) { phi0 = pf0[6]; phi1 = pf1[6]; }
  else phi0 = phi1 = c.fixed_vfocal;
  const double D0 = depth_value(c, L, dg0, o0.depth, pf0), D1 = depth_value(c, L, dg1, o1.depth, pf1);
  double u0[2], u1[2];
  warp_value(L, sg0, pf0, u0); warp_value(L, sg1, pf1, u1);
  static_scene<JAC>(c, pf0, phi0, D0, u0, pf1, phi1, D1, u1, o0, o1, ev.r, Jl);
  const double s = ev.r[0] * ev.r[0] + ev.r[1] * ev.r[1] + ev.r[2] * ev.r[2];
  double rho1;
  robust_loss(c, s, ev.rho0, rho1);
  ev.scale = sqrt(rho1);    // Corrector with rho'' <= 0: residual and Jacobian scaled by sqrt(rho')
}

// Block-level sum -> partial[blockIdx.x]  (deterministic two-stage cost reduction)
__device__ __forceinline__ void block_store_sum(double v, double* partial) {
  __shared__ double red[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) *partial = t;
  }
}

// K2: cost only.
__global__ void __launch_bounds__(kTile) k_cost_static(DevProblem p, const double* __restrict__ x, double* __restrict__ partial) {
  const int t = blockIdx.x;
  const int pr = p.tile_pair[t];
  const int f0 = p.pair_frames[2 * pr], f1 = p.pair_frames[2 * pr + 1];
  double cost = 0.0;
  if ((int)threadIdx.x < p.tile_count[t]) {
    const float* rec = p.records + (size_t)(p.tile_begin[t] + threadIdx.x) * 6;
    Gather dg0, sg0, dg1, sg1; StaticEval ev;
    eval_constraint<false>(p, x, f0, f1, rec, dg0, sg0, dg1, sg1, ev, nullptr);
    cost = 0.5 * ev.rho0;
  }
  block_store_sum(cost, partial + t);
}

// K1 (generic path): every transform / intrinsics mode; scatter with L2 reductions.
// WANT_H = false gives the gradient-only evaluation used by the bounded line search.
template <bool WANT_H>
__global__ void __launch_bounds__(kTile) k_accumulate_generic(DevProblem p, const double* __restrict__ x, double* __restrict__ H,
                                                               double* __restrict__ g, double* __restrict__ partial) {
  const rcvd_config& c = p.cfg; const Layout& L = p.L;
  const int t = blockIdx.x;
  const int pr = p.tile_pair[t];
  const int f0 = p.pair_frames[2 * pr], f1 = p.pair_frames[2 * pr + 1];
  double cost = 0.0;
  if ((int)threadIdx.x < p.tile_count[t]) {
    const float* rec = p.records + (size_t)(p.tile_begin[t] + threadIdx.x) * 6;
    Gather dg0, sg0, dg1, sg1; StaticEval ev; double Jl[60];
    eval_constraint<true>(p, x, f0, f1, rec, dg0, sg0, dg1, sg1, ev, Jl);
    cost = 0.5 * ev.rho0;
    const double sc = ev.scale;
    const double r0 = ev.r[0] * sc, r1 = ev.r[1] * sc, r2 = ev.r[2] * sc;
    // expanded columns: (frame, local column, 3-vector)
    int ef[kMaxEntries]; short el[kMaxEntries]; double ej[kMaxEntries][3];
    int E = 0;
    auto push = [&](int f, int l, double a0, double a1, double a2) {
      if (is_const_local(c, L, l)) return;
      ef[E] = f; el[E] = (short)l; ej[E][0] = a0 * sc; ej[E][1] = a1 * sc; ej[E][2] = a2 * sc; ++E;
    };
    for (int side = 0; side < 2; ++side) {
      const int f = side ? f1 : f0; const int o = side * 10;
      const Gather& dg = side ? dg1 : dg0; const Gather& sg = side ? sg1 : sg0;
      const double src = (double)(side ? rec[5] : rec[2]);
      for (int q = 0; q < 6; ++q) push(f, q, Jl[o + q], Jl[20 + o + q], Jl[40 + o + q]);
      if (c.intr_opt == RCVD_INTR_PER_FRAME) push(f, 6, Jl[o + 6], Jl[20 + o + 6], Jl[40 + o + 6]);
      for (int q = 0; q < dg.n; ++q) {
        const double w = dg.w[q];
        push(f, L.offD + dg.idx[q] * L.k, Jl[o + 7] * w * src, Jl[20 + o + 7] * w * src, Jl[40 + o + 7] * w * src);
        if (L.k == 2) push(f, L.offD + dg.idx[q] * 2 + 1, Jl[o + 7] * w, Jl[20 + o + 7] * w, Jl[40 + o + 7] * w);
      }
      for (int q = 0; q < sg.n; ++q) {
        const double w = sg.w[q];
        push(f, L.offS + sg.idx[q] * 2, Jl[o + 8] * w, Jl[20 + o + 8] * w, Jl[40 + o + 8] * w);
        push(f, L.offS + sg.idx[q] * 2 + 1, Jl[o + 9] * w, Jl[20 + o + 9] * w, Jl[40 + o + 9] * w);
      }
    }
    if (c.intr_opt == RCVD_INTR_SHARED) push(0, 6, Jl[6] + Jl[16], Jl[26] + Jl[36], Jl[46] + Jl[56]);
    const int np = L.npad;
    for (int a = 0; a < E; ++a) {
      const double a0 = ej[a][0], a1 = ej[a][1], a2 = ej[a][2];
      red_add(g + (size_t)ef[a] * np + el[a], a0 * r0 + a1 * r1 + a2 * r2);
      if (WANT_H) {
        for (int b = 0; b <= a; ++b) {
          const double v = a0 * ej[b][0] + a1 * ej[b][1] + a2 * ej[b][2];
          add_h(p, H, ef[a], el[a], ef[b], el[b], v);
        }
      }
    }
  }
  block_store_sum(cost, partial + t);
}

// ---------------------------------------------------------------------------
// K1 (fast path): depth transform Identity / Global / bilinear Grid with Scale value transform,
// identity spatial transform, PerFrame or Fixed intrinsics, nothing held constant -- the
// reference's default configuration (pose_optimization.py:197-207 + coarse-to-fine grids).
// Per constraint the Jacobian is kept in "local" variables (pose 6, focal, depth D per frame:
// 3 x 16); the dense 14 x 14 pose/focal normal block and its gradient are reduced over the
// 128 constraints of the tile on the fp64 tensor cores (J^T [J | r] as an m8n8k4 DMMA GEMM with
// K = 3 x 128 residual rows staged in shared memory), so they cost 119 L2 reductions per tile
// instead of per constraint.  The spline-node columns (<= 4 nodes per frame) are expanded per
// thread and scattered with RED.ADD.F64.
// ---------------------------------------------------------------------------
constexpr int kJsLd = 20;   // [k][col] staging layout, 20-double rows: conflict-free DMMA fragment loads

__device__ __forceinline__ void dmma_acc(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ bool fast_path_ok(const rcvd_config& c, const Layout& L) {
  return L.k == 1 && c.spatial_type == RCVD_SPATIAL_IDENTITY && c.intr_opt != RCVD_INTR_SHARED && !c.fix_poses && !c.fix_depth_xforms &&
         !c.fix_spatial_xforms && (c.depth_type != RCVD_DEPTH_GRID || !c.depth_cubic);
}

__global__ void __launch_bounds__(kTile) k_accumulate_fast(DevProblem p, const double* __restrict__ x, double* __restrict__ H,
                                                            double* __restrict__ g, double* __restrict__ partial) {
  extern __shared__ __align__(16) double sm[];
  double* Js = sm;                                  // [3*kTile][kJsLd]
  double* Ms = sm + 3 * kTile * kJsLd;              // [4 warps][16][16]
  const rcvd_config& c = p.cfg; const Layout& L = p.L;
  const int t = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pr = p.tile_pair[t];
  const int f0 = p.pair_frames[2 * pr], f1 = p.pair_frames[2 * pr + 1];
  const int np = L.npad;
  const size_t bs = (size_t)np * np;
  const int enc = p.blk_of[f0 * p.N + f1];          // cross block id*2 + (f0 is the row side)
  double* Hx = H + (size_t)(enc >> 1) * bs;
  const bool f0rows = enc & 1;
  double* H0 = H + (size_t)f0 * bs; double* H1 = H + (size_t)f1 * bs;
  double cost = 0.0;
  const bool active = tid < p.tile_count[t];
  double Jl[60]; double r0 = 0, r1 = 0, r2 = 0;
  int nn = 0; int idx0[4], idx1[4]; double w0[4], w1[4];
#pragma unroll
  for (int i = 0; i < 60; ++i) Jl[i] = 0.0;
  if (active) {
    const float* rec = p.records + (size_t)(p.tile_begin[t] + tid) * 6;
    const double* pf0 = x + (size_t)f0 * L.nf; const double* pf1 = x + (size_t)f1 * L.nf;
    ObsIn o0{rec[0], rec[1], rec[2]}, o1{rec[3], rec[4], rec[5]};
    double D0 = (double)o0.depth, D1 = (double)o1.depth;
    if (c.depth_type == RCVD_DEPTH_GLOBAL) {
      nn = 1; idx0[0] = 0; idx1[0] = 0; w0[0] = D0; w1[0] = D1;          // dD/ds = src
      D0 *= pf0[L.offD]; D1 *= pf1[L.offD];
    } else if (c.depth_type == RCVD_DEPTH_GRID) {
      nn = 4;
      int ix, iy; double rx, ry;
      cell_coord(o0.ndcx, c.depth_grid_x, ix, rx); cell_coord(o0.ndcy, c.depth_grid_y, iy, ry);
      idx0[0] = ix + iy * c.depth_grid_x; idx0[1] = idx0[0] + 1; idx0[2] = idx0[0] + c.depth_grid_x; idx0[3] = idx0[2] + 1;
      double ox = __dsub_rn(1.0, rx), oy = __dsub_rn(1.0, ry);
      w0[0] = __dmul_rn(ox, oy); w0[1] = __dmul_rn(rx, oy); w0[2] = __dmul_rn(ox, ry); w0[3] = __dmul_rn(rx, ry);
      cell_coord(o1.ndcx, c.depth_grid_x, ix, rx); cell_coord(o1.ndcy, c.depth_grid_y, iy, ry);
      idx1[0] = ix + iy * c.depth_grid_x; idx1[1] = idx1[0] + 1; idx1[2] = idx1[0] + c.depth_grid_x; idx1[3] = idx1[2] + 1;
      ox = __dsub_rn(1.0, rx); oy = __dsub_rn(1.0, ry);
      w1[0] = __dmul_rn(ox, oy); w1[1] = __dmul_rn(rx, oy); w1[2] = __dmul_rn(ox, ry); w1[3] = __dmul_rn(rx, ry);
      // GridDepthFunctor::eval: res += (src * s_i) * w_i, in node order
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) { a0 += (D0 * pf0[L.offD + idx0[q]]) * w0[q]; a1 += (D1 * pf1[L.offD + idx1[q]]) * w1[q]; }
#pragma unroll
      for (int q = 0; q < 4; ++q) { w0[q] *= D0; w1[q] *= D1; }          // dD/ds_i = w_i * src
      D0 = a0; D1 = a1;
    }
    const double phi0 = (c.intr_opt == RCVD_INTR_PER_FRAME) ? pf0[6] : c.fixed_vfocal;
    const double phi1 = (c.intr_opt == RCVD_INTR_PER_FRAME) ? pf1[6] : c.fixed_vfocal;
    const double u[2] = {0.0, 0.0};
    double r[3];
    static_scene<true>(c, pf0, phi0, D0, u, pf1, phi1, D1, u, o0, o1, r, Jl);
    const double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double rho0, rho1;
    robust_loss(c, s, rho0, rho1);
    cost = 0.5 * rho0;
    const double sc = sqrt(rho1);
    r0 = r[0] * sc; r1 = r[1] * sc; r2 = r[2] * sc;
#pragma unroll
    for (int i = 0; i < 60; ++i) Jl[i] *= sc;
    if (c.intr_opt != RCVD_INTR_PER_FRAME) { Jl[6] = Jl[26] = Jl[46] = 0.0; Jl[16] = Jl[36] = Jl[56] = 0.0; }
  }
  // ---- stage [J_P | r] rows: columns 0..6 frame-0 pose+focal, 7..13 frame-1, 14 = r, 15 = 0 ----
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double* row = Js + (size_t)(tid * 3 + i) * kJsLd;
#pragma unroll
    for (int q = 0; q < 7; ++q) { row[q] = Jl[i * 20 + q]; row[7 + q] = Jl[i * 20 + 10 + q]; }
    row[14] = (i == 0) ? r0 : (i == 1 ? r1 : r2);
    row[15] = 0.0;
  }
  // ---- per-thread scatter of the spline-node columns ----
  if (active && nn > 0) {
    const double ca0 = Jl[7], ca1 = Jl[27], ca2 = Jl[47];      // d r / d D0
    const double cb0 = Jl[17], cb1 = Jl[37], cb2 = Jl[57];     // d r / d D1
    const double gDa = ca0 * r0 + ca1 * r1 + ca2 * r2, gDb = cb0 * r0 + cb1 * r1 + cb2 * r2;
    const double mAA = ca0 * ca0 + ca1 * ca1 + ca2 * ca2, mBB = cb0 * cb0 + cb1 * cb1 + cb2 * cb2, mAB = ca0 * cb0 + ca1 * cb1 + ca2 * cb2;
    double mPa[14], mPb[14];   // M[P, Da], M[P, Db]
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      mPa[q] = Jl[q] * ca0 + Jl[20 + q] * ca1 + Jl[40 + q] * ca2;
      mPa[7 + q] = Jl[10 + q] * ca0 + Jl[30 + q] * ca1 + Jl[50 + q] * ca2;
      mPb[q] = Jl[q] * cb0 + Jl[20 + q] * cb1 + Jl[40 + q] * cb2;
      mPb[7 + q] = Jl[10 + q] * cb0 + Jl[30 + q] * cb1 + Jl[50 + q] * cb2;
    }
    const int nphi = (c.intr_opt == RCVD_INTR_PER_FRAME) ? 7 : 6;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < nn) {
        const int la = L.offD + idx0[i], lb = L.offD + idx1[i];
        const double wa = w0[i], wb = w1[i];
        red_add(g + (size_t)f0 * np + la, gDa * wa);
        red_add(g + (size_t)f1 * np + lb, gDb * wb);
        for (int q = 0; q < nphi; ++q) {
          // node of frame 0 against pose/focal of frame 0 (same block, node row > pose col) and of frame 1 (cross block)
          red_add(H0 + (size_t)la * np + q, mPa[q] * wa);
          if (f0rows) red_add(Hx + (size_t)la * np + q, mPa[7 + q] * wa); else red_add(Hx + (size_t)q * np + la, mPa[7 + q] * wa);
          red_add(H1 + (size_t)lb * np + q, mPb[7 + q] * wb);
          if (f0rows) red_add(Hx + (size_t)q * np + lb, mPb[q] * wb); else red_add(Hx + (size_t)lb * np + q, mPb[q] * wb);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < nn) {
            const int la2 = L.offD + idx0[j], lb2 = L.offD + idx1[j];
            if (la >= la2) red_add(H0 + (size_t)la * np + la2, mAA * wa * w0[j]);
            if (lb >= lb2) red_add(H1 + (size_t)lb * np + lb2, mBB * wb * w1[j]);
            if (f0rows) red_add(Hx + (size_t)la * np + lb2, mAB * wa * w1[j]); else red_add(Hx + (size_t)lb2 * np + la, mAB * wa * w1[j]);
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- M = J_P^T [J_P | r] over this warp's 96 residual rows on the fp64 tensor cores ----
  {
    const int gq = lane >> 2, tq = lane & 3;
    double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
    const double* base = Js + (size_t)warp * 96 * kJsLd;
#pragma unroll 4
    for (int k4 = 0; k4 < 24; ++k4) {
      const double* rowp = base + (size_t)(k4 * 4 + tq) * kJsLd;
      const double a0 = rowp[gq], a1 = rowp[8 + gq];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma_acc(acc[i][j][0], acc[i][j][1], i ? a1 : a0, j ? a1 : a0);
    }
    double* mw = Ms + warp * 256;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { mw[(i * 8 + gq) * 16 + j * 8 + 2 * tq] = acc[i][j][0]; mw[(i * 8 + gq) * 16 + j * 8 + 2 * tq + 1] = acc[i][j][1]; }
  }
  __syncthreads();
  for (int e = tid; e < 256; e += kTile) {
    const int i = e >> 4, j = e & 15;
    if (i >= 14 || j >= 15) continue;
    const double v = Ms[e] + Ms[256 + e] + Ms[512 + e] + Ms[768 + e];
    const int fi = i < 7 ? f0 : f1, li = i < 7 ? i : i - 7;
    if (j == 14) { red_add(g + (size_t)fi * np + li, v); continue; }
    const bool jf0 = j < 7; const int lj = jf0 ? j : j - 7;
    if ((i < 7) == jf0) { if (li >= lj) red_add((i < 7 ? H0 : H1) + (size_t)li * np + lj, v); }
    else if (i < 7) {   // row index in frame 0, column in frame 1: each cross entry appears twice in M (i<7,j>=7 and mirrored); take this one
      if (f0rows) red_add(Hx + (size_t)li * np + lj, v); else red_add(Hx + (size_t)lj * np + li, v);
    }
  }
  block_store_sum(cost, partial + t);
}

// ---------------------------------------------------------------------------
// K1 (run path, round 2): bilinear depth grid, the reference's default once coarse-to-fine has left the Global transform.
// The records of a pair are sorted by (source cell, target cell) when the problem is set up (rcvd_api.cu, device segmented sort), so
// consecutive constraints share their eight spline nodes.  A run = the constraints of one warp with the same cell pair.  Per run the
// node rows of the normal equations,
//       M[node, :] = sum_c  J_node(c)^T [ J_pose(c) | r(c) | J_node(c) ]        (8 x 24: 4 + 4 nodes against 14 pose/focal, r, 8 nodes)
// are reduced over the run on the fp64 tensor cores (m8n8k4, K = the run's residual rows staged in shared memory) and leave the SM as ONE
// reduction per entry per run (156 REDs) instead of one per constraint; the 14 x 14 pose/focal block is reduced over the whole
// 128-constraint tile as before.  matchSeparation 10 (about three constraints per cell pair): ~53 REDs per constraint instead of 157;
// dense constraints (hundreds per cell pair): ~6.
// ---------------------------------------------------------------------------
constexpr int kRunLd = 24;          // [row][col]: 0-13 pose/focal (frame 0, frame 1), 14 r, 15 zero, 16-19 nodes of frame 0, 20-23 nodes of frame 1
constexpr int kRunSmem = ((3 * kTile + 4) * kRunLd + 4 * 256) * (int)sizeof(double) + kTile * (int)sizeof(unsigned);

__global__ void __launch_bounds__(kTile) k_accumulate_runs(DevProblem p, const double* __restrict__ x, double* __restrict__ H,
                                                            double* __restrict__ g, double* __restrict__ partial) {
  extern __shared__ __align__(16) double sm[];
  double* Js = sm;                                        // [3*kTile + 4][kRunLd]
  double* Ms = sm + (3 * kTile + 4) * kRunLd;             // [4 warps][16][16]
  unsigned* skey = reinterpret_cast<unsigned*>(Ms + 4 * 256);
  const rcvd_config& c = p.cfg; const Layout& L = p.L;
  const int t = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pr = p.tile_pair[t];
  const int f0 = p.pair_frames[2 * pr], f1 = p.pair_frames[2 * pr + 1];
  const int np = L.npad; const size_t bs = (size_t)np * np;
  const int enc = p.blk_of[f0 * p.N + f1];
  double* Hx = H + (size_t)(enc >> 1) * bs;
  const bool f0rows = enc & 1;
  double* H0 = H + (size_t)f0 * bs; double* H1 = H + (size_t)f1 * bs;
  const int gx = c.depth_grid_x;
  double cost = 0.0;
  const bool active = tid < p.tile_count[t];
  unsigned key = 0xffffffffu;
  {
    double Jl[60]; double r0 = 0, r1 = 0, r2 = 0; double w0[4] = {0, 0, 0, 0}, w1[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 60; ++i) Jl[i] = 0.0;
    if (active) {
      const float* rec = p.records + (size_t)(p.tile_begin[t] + tid) * 6;
      const double* pf0 = x + (size_t)f0 * L.nf; const double* pf1 = x + (size_t)f1 * L.nf;
      ObsIn o0{rec[0], rec[1], rec[2]}, o1{rec[3], rec[4], rec[5]};
      double D0 = (double)o0.depth, D1 = (double)o1.depth;
      int ix, iy; double rx, ry;
      cell_coord(o0.ndcx, gx, ix, rx); cell_coord(o0.ndcy, c.depth_grid_y, iy, ry);
      const int ba = ix + iy * gx;
      double ox = __dsub_rn(1.0, rx), oy = __dsub_rn(1.0, ry);
      w0[0] = __dmul_rn(ox, oy); w0[1] = __dmul_rn(rx, oy); w0[2] = __dmul_rn(ox, ry); w0[3] = __dmul_rn(rx, ry);
      cell_coord(o1.ndcx, gx, ix, rx); cell_coord(o1.ndcy, c.depth_grid_y, iy, ry);
      const int bb = ix + iy * gx;
      ox = __dsub_rn(1.0, rx); oy = __dsub_rn(1.0, ry);
      w1[0] = __dmul_rn(ox, oy); w1[1] = __dmul_rn(rx, oy); w1[2] = __dmul_rn(ox, ry); w1[3] = __dmul_rn(rx, ry);
      key = ((unsigned)ba << 16) | (unsigned)bb;
      // GridDepthFunctor::eval: res += (src * s_i) * w_i, in node order
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int na = ba + (q & 1) + (q >> 1) * gx, nb = bb + (q & 1) + (q >> 1) * gx;
        a0 += (D0 * pf0[L.offD + na]) * w0[q]; a1 += (D1 * pf1[L.offD + nb]) * w1[q];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { w0[q] *= D0; w1[q] *= D1; }          // dD/ds_i = w_i * src
      D0 = a0; D1 = a1;
      const double phi0 = (c.intr_opt == RCVD_INTR_PER_FRAME) ? pf0[6] : c.fixed_vfocal;
      const double phi1 = (c.intr_opt == RCVD_INTR_PER_FRAME) ? pf1[6] : c.fixed_vfocal;
      const double u[2] = {0.0, 0.0};
      double r[3];
      static_scene<true>(c, pf0, phi0, D0, u, pf1, phi1, D1, u, o0, o1, r, Jl);
      const double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      double rho0, rho1;
      robust_loss(c, s, rho0, rho1);
      cost = 0.5 * rho0;
      const double sc = sqrt(rho1);
      r0 = r[0] * sc; r1 = r[1] * sc; r2 = r[2] * sc;
#pragma unroll
      for (int i = 0; i < 60; ++i) Jl[i] *= sc;
      if (c.intr_opt != RCVD_INTR_PER_FRAME) { Jl[6] = Jl[26] = Jl[46] = 0.0; Jl[16] = Jl[36] = Jl[56] = 0.0; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double* row = Js + (size_t)(tid * 3 + i) * kRunLd;
#pragma unroll
      for (int q = 0; q < 7; ++q) { row[q] = Jl[i * 20 + q]; row[7 + q] = Jl[i * 20 + 10 + q]; }
      row[14] = (i == 0) ? r0 : (i == 1 ? r1 : r2);
      row[15] = 0.0;
      const double ca = Jl[i * 20 + 7], cb = Jl[i * 20 + 17];            // d r_i / d D0, d r_i / d D1
#pragma unroll
      for (int q = 0; q < 4; ++q) { row[16 + q] = ca * w0[q]; row[20 + q] = cb * w1[q]; }
    }
    if (tid < 4 * kRunLd) Js[(size_t)3 * kTile * kRunLd + tid] = 0.0;     // four zero rows behind the tile (the K steps of the last run read them)
    skey[tid] = key;
  }
  __syncthreads();
  const int gq = lane >> 2, tq = lane & 3;
  // ---- pose / focal block of the whole tile: M = J_P^T [J_P | r] over this warp's 96 residual rows ----
  {
    double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
    const double* base = Js + (size_t)warp * 96 * kRunLd;
#pragma unroll 4
    for (int k4 = 0; k4 < 24; ++k4) {
      const double* rowp = base + (size_t)(k4 * 4 + tq) * kRunLd;
      const double a0 = rowp[gq], a1 = rowp[8 + gq];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma_acc(acc[i][j][0], acc[i][j][1], i ? a1 : a0, j ? a1 : a0);
    }
    double* mw = Ms + warp * 256;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { mw[(i * 8 + gq) * 16 + j * 8 + 2 * tq] = acc[i][j][0]; mw[(i * 8 + gq) * 16 + j * 8 + 2 * tq + 1] = acc[i][j][1]; }
  }
  // ---- spline-node rows, one run at a time (runs never cross a warp: at most three extra runs per tile) ----
  {
    const int c0 = warp * 32 + lane;
    const unsigned kprev = lane ? skey[c0 - 1] : ~key;
    const unsigned starts = __ballot_sync(0xffffffffu, key != kprev || lane == 0);
    unsigned rem = starts;
    const bool per_frame = c.intr_opt == RCVD_INTR_PER_FRAME;
    while (rem) {
      const int s0 = __ffs(rem) - 1; rem &= rem - 1;
      const int s1 = rem ? __ffs(rem) - 1 : 32;
      const unsigned rk = skey[warp * 32 + s0];
      if (rk == 0xffffffffu) break;                        // the padding behind the last constraint of the tile
      const int row0 = 3 * (warp * 32 + s0), row1 = 3 * (warp * 32 + s1);
      double acc[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
      for (int rr = row0; rr < row1; rr += 4) {
        const int row = rr + tq;
        const double* rowp = Js + (size_t)row * kRunLd;
        const bool in = row < row1;                                       // the last K step of a run reaches into the next run: masked
        const double a = in ? rowp[16 + gq] : 0.0;
#pragma unroll
        for (int nb = 0; nb < 3; ++nb) dmma_acc(acc[nb][0], acc[nb][1], a, in ? rowp[nb * 8 + gq] : 0.0);
      }
      // lane (gq, tq) holds M[node gq][columns nb*8 + 2 tq, + 1]
      const int ba = (int)(rk >> 16), bb = (int)(rk & 0xffffu);
      const bool rowA = gq < 4;
      const int qn = gq & 3;
      const int ln = L.offD + (rowA ? ba : bb) + (qn & 1) + (qn >> 1) * gx;   // local column of this lane's node
      double* Hn = rowA ? H0 : H1;                                         // diagonal block of the node's frame
      const int fn = rowA ? f0 : f1;
#pragma unroll
      for (int nb = 0; nb < 3; ++nb)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = nb * 8 + 2 * tq + e;
          const double v = acc[nb][e];
          if (col < 14) {
            const int pc = col < 7 ? col : col - 7;
            if (pc == 6 && !per_frame) continue;
            const bool colA = col < 7;
            if (colA == rowA) red_add(Hn + (size_t)ln * np + pc, v);                         // node row > pose column: lower triangle
            else if (rowA == f0rows) red_add(Hx + (size_t)ln * np + pc, v);                  // the node's frame is the row side of the cross block
            else red_add(Hx + (size_t)pc * np + ln, v);
          } else if (col == 14) {
            red_add(g + (size_t)fn * np + ln, v);
          } else if (col >= 16) {
            const int m = col - 16, qm = m & 3; const bool colNodeA = m < 4;
            const int lm = L.offD + (colNodeA ? ba : bb) + (qm & 1) + (qm >> 1) * gx;
            if (colNodeA == rowA) { if (ln >= lm) red_add(Hn + (size_t)ln * np + lm, v); }   // same frame: lower triangle once
            else if (rowA) {                                                                 // (frame-0 node, frame-1 node): once, from the frame-0 row
              if (f0rows) red_add(Hx + (size_t)ln * np + lm, v); else red_add(Hx + (size_t)lm * np + ln, v);
            }
          }
        }
    }
  }
  __syncthreads();
  for (int e = tid; e < 256; e += kTile) {
    const int i = e >> 4, j = e & 15;
    if (i >= 14 || j >= 15) continue;
    const double v = Ms[e] + Ms[256 + e] + Ms[512 + e] + Ms[768 + e];
    const int fi = i < 7 ? f0 : f1, li = i < 7 ? i : i - 7;
    if (j == 14) { red_add(g + (size_t)fi * np + li, v); continue; }
    const bool jf0 = j < 7; const int lj = jf0 ? j : j - 7;
    if ((i < 7) == jf0) { if (li >= lj) red_add((i < 7 ? H0 : H1) + (size_t)li * np + lj, v); }
    else if (i < 7) { if (f0rows) red_add(Hx + (size_t)li * np + lj, v); else red_add(Hx + (size_t)lj * np + li, v); }
  }
  block_store_sum(cost, partial + t);
}

// Sort key of a record for the run path: (top-left node of the source cell) << 16 | (top-left node of the target cell)
__global__ void __launch_bounds__(256) k_record_keys(rcvd_config c, const float* __restrict__ records, long long n, unsigned* __restrict__ keys, int* __restrict__ idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* rec = records + (size_t)i * 6;
  int ix, iy; double rx, ry;
  cell_coord(rec[0], c.depth_grid_x, ix, rx); cell_coord(rec[1], c.depth_grid_y, iy, ry);
  const unsigned ba = (unsigned)(ix + iy * c.depth_grid_x);
  cell_coord(rec[3], c.depth_grid_x, ix, rx); cell_coord(rec[4], c.depth_grid_y, iy, ry);
  keys[i] = (ba << 16) | (unsigned)(ix + iy * c.depth_grid_x);
  idx[i] = (int)i;
}
__global__ void __launch_bounds__(256) k_gather_records(const float* __restrict__ src, const int* __restrict__ idx, long long n, float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 6) return;
  const long long rcd = i / 6; const int e = (int)(i % 6);
  dst[i] = src[(size_t)idx[rcd] * 6 + e];
}

// Marks parameters referenced by at least one residual block (the Ceres program's
// parameter set): used for |x| / |step| norms.  mask has npad stride.
__global__ void __launch_bounds__(kTile) k_mark_static(DevProblem p, uint8_t* __restrict__ mask) {
  const rcvd_config& c = p.cfg; const Layout& L = p.L;
  const int t = blockIdx.x;
  if ((int)threadIdx.x >= p.tile_count[t]) return;
  const int pr = p.tile_pair[t];
  const int fr[2] = {p.pair_frames[2 * pr], p.pair_frames[2 * pr + 1]};
  const float* rec = p.records + (size_t)(p.tile_begin[t] + threadIdx.x) * 6;
  const int np = L.npad;
  for (int side = 0; side < 2; ++side) {
    Gather dg, sg;
    gather_depth(c, rec[side * 3], rec[side * 3 + 1], dg); gather_spatial(c, rec[side * 3], rec[side * 3 + 1], sg);
    uint8_t* m = mask + (size_t)fr[side] * np;
    for (int q = 0; q < 6; ++q) m[q] = 1;
    if (c.intr_opt == RCVD_INTR_PER_FRAME) m[6] = 1;
    for (int q = 0; q < dg.n; ++q) for (int j = 0; j < L.k; ++j) m[L.offD + dg.idx[q] * L.k + j] = 1;
    for (int q = 0; q < sg.n; ++q) { m[L.offS + sg.idx[q] * 2] = 1; m[L.offS + sg.idx[q] * 2 + 1] = 1; }
  }
  if (c.intr_opt == RCVD_INTR_SHARED) mask[6] = 1;
}

// --- K3: regulariser rows ---------------------------------------------------
struct RegCounts { int scale, deform, spatial, focal, per_frame; int position_rows; int total; };

__host__ __device__ inline RegCounts reg_counts(const rcvd_config& c, const Layout& L, int N, int nscale) {
  RegCounts r;
  r.scale = (c.scale_reg > 0.0 && !c.fix_depth_xforms && (c.depth_type == RCVD_DEPTH_GLOBAL || c.depth_type == RCVD_DEPTH_GRID)) ? nscale : 0;
  r.deform = (c.depth_deform_reg > 0.0 && c.depth_type == RCVD_DEPTH_GRID)
                 ? ((c.depth_grid_x - 1) * c.depth_grid_y + c.depth_grid_x * (c.depth_grid_y - 1)) * L.k : 0;
  r.spatial = (c.spatial_deform_reg > 0.0) ? L.ns : 0;
  r.focal = (c.focal_reg > 0.0 && c.intr_opt != RCVD_INTR_FIXED) ? 1 : 0;
  r.per_frame = r.scale + r.deform + r.spatial + r.focal;
  r.position_rows = (c.position_reg > 0.0 && N >= 3) ? (N - 2) * 3 : 0;
  r.total = r.per_frame * N + r.position_rows;
  return r;
}

// MODE 0: cost only, 1: cost + gradient + H, 2: mark active, 3: cost + gradient
template <int MODE>
__global__ void __launch_bounds__(128) k_regularisers(DevProblem p, RegCounts rc, const double* __restrict__ x, double* __restrict__ H,
                                                      double* __restrict__ g, double* __restrict__ partial, uint8_t* __restrict__ mask,
                                                      int first_frame, int last_frame) {
  const rcvd_config& c = p.cfg; const Layout& L = p.L;
  const int np = L.npad;
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  double cost = 0.0;
  int nent = 0; int ef[3] = {0, 0, 0}; int el[32]; double ed[32]; int efr[32];   // entries (frame, local, derivative)
  double r = 0.0; bool valid = false;
  if (id < rc.per_frame * p.N) {
    const int f = id / rc.per_frame; int k = id % rc.per_frame;
    const bool mine = (p.nranks <= 1) || (f % p.nranks == p.rank);
    if (p.in_range[f] && mine) {
      const double* pf = x + (size_t)f * L.nf;
      ef[0] = f;
      if (k < rc.scale) {
        // TargetDisparityCost (lib/PoseOptimizer.cpp:488-517) on the lattice of :1382-1385, ScaledLoss(scaleReg)
        const double sw = sqrt(c.scale_reg);
        const float med = (float)p.median[f];
        Gather gth; gather_depth(c, p.scale_locs[2 * k], p.scale_locs[2 * k + 1], gth);
        const double depth = depth_value(c, L, gth, med, pf);
        const bool clamped = depth < 1e-6;
        r = (1.0 / (clamped ? 1e-6 : depth) - 1.0) * sw;
        const double dd = clamped ? 0.0 : -1.0 / (depth * depth);
        for (int q = 0; q < gth.n; ++q) {
          el[nent] = L.offD + gth.idx[q] * L.k; ed[nent] = dd * gth.w[q] * (double)med * sw; efr[nent] = f; ++nent;
          if (L.k == 2) { el[nent] = L.offD + gth.idx[q] * 2 + 1; ed[nent] = dd * gth.w[q] * sw; efr[nent] = f; ++nent; }
        }
        valid = true;
      } else if ((k -= rc.scale) < rc.deform) {
        // computeGridDeformationCost (lib/DepthMapTransform.cpp:631-667) * weight (DeformationCost / Adaptive, :536-656)
        const int gx = c.depth_grid_x, gy = c.depth_grid_y;
        const int comp = k % L.k; const int e = k / L.k;
        int a, b;
        const int nh = (gx - 1) * gy;
        if (e < nh) { const int yy = e / (gx - 1), xx = e % (gx - 1) + 1; a = xx + yy * gx; b = a - 1; }
        else { const int e2 = e - nh; const int yy = e2 / gx + 1, xx = e2 % gx; a = xx + yy * gx; b = a - gx; }
        double w = c.depth_deform_reg;
        if (c.adaptive_deform > 0.0 && p.adaptive) {
          const double* aw = p.adaptive + (size_t)f * gx * gy;
          w = c.depth_deform_reg + fmax(aw[a], aw[b]) * c.adaptive_deform;
        }
        const int la = L.offD + a * L.k + comp, lb = L.offD + b * L.k + comp;
        const double va = pf[la], vb = pf[lb];
        const double aa = va < 0.0 ? -va : va, ab = vb < 0.0 ? -vb : vb;
        const bool useB = ab < aa;               // min(abs(this), abs(that)) = (that < this) ? that : this
        const double m = useB ? ab : aa;
        r = (va - vb) / m * w;
        double da = 1.0 / m, db = -1.0 / m;
        if (useB) db += -(va - vb) / (m * m) * (vb < 0.0 ? -1.0 : 1.0);
        else da += -(va - vb) / (m * m) * (va < 0.0 ? -1.0 : 1.0);
        el[0] = la; ed[0] = da * w; efr[0] = f; el[1] = lb; ed[1] = db * w; efr[1] = f; nent = 2;
        valid = true;
      } else if ((k -= rc.deform) < rc.spatial) {
        // paramsToResiduals (lib/DepthMapTransform.cpp:61-70) * spatialDeformReg
        r = pf[L.offS + k] * c.spatial_deform_reg; el[0] = L.offS + k; ed[0] = c.spatial_deform_reg; efr[0] = f; nent = 1; valid = true;
      } else {
        // TargetFocalCost (lib/PoseOptimizer.cpp:520-533), ScaledLoss(focalReg)
        const double sw = sqrt(c.focal_reg);
        r = (pf[6] - c.focal_target) * sw; el[0] = 6; ed[0] = sw; efr[0] = f; nent = 1; valid = true;
      }
    }
  } else if (id < rc.total) {
    // ParameterRegularizationCost (lib/PoseOptimizer.cpp:464-483) over in-range triplets (:1420-1426)
    const int k = id - rc.per_frame * p.N;
    const int f = k / 3, i = k % 3;
    const bool mine = (p.nranks <= 1) || (f % p.nranks == p.rank);
    if (mine && f >= first_frame && f < last_frame - 1 && p.in_range[f] && p.in_range[f + 1] && p.in_range[f + 2]) {
      const double sw = sqrt(c.position_reg);
      r = (x[(size_t)f * L.nf + i] - 2.0 * x[(size_t)(f + 1) * L.nf + i] + x[(size_t)(f + 2) * L.nf + i]) * sw;
      el[0] = i; ed[0] = sw; efr[0] = f; el[1] = i; ed[1] = -2.0 * sw; efr[1] = f + 1; el[2] = i; ed[2] = sw; efr[2] = f + 2; nent = 3;
      valid = true;
    }
  }
  if (valid) {
    cost = 0.5 * r * r;
    if (MODE == 2) { for (int a = 0; a < nent; ++a) mask[(size_t)efr[a] * np + el[a]] = 1; }
    if (MODE == 1 || MODE == 3) {
      for (int a = 0; a < nent; ++a) if (is_const_local(c, L, el[a])) ed[a] = 0.0;
      for (int a = 0; a < nent; ++a) {
        if (ed[a] == 0.0) continue;
        red_add(g + (size_t)efr[a] * np + el[a], ed[a] * r);
        if (MODE == 1) for (int b = 0; b <= a; ++b) if (ed[b] != 0.0) add_h(p, H, efr[a], el[a], efr[b], el[b], ed[a] * ed[b]);
      }
    }
  }
  if (MODE != 2) block_store_sum(cost, partial + blockIdx.x);
}

// --- scene-flow smoothness residual blocks (reference addSceneFlowSmoothnessLoss, lib/PoseOptimizer.cpp:1242-1339) ---
// MODE 0: cost only, 1: cost + gradient + H, 2: mark active parameters, 3: cost + gradient
template <int MODE>
__global__ void __launch_bounds__(kTile) k_triplets(DevProblem p, const double* __restrict__ x, double* __restrict__ H, double* __restrict__ g,
                                                    double* __restrict__ partial, uint8_t* __restrict__ mask) {
  const rcvd_config& c = p.cfg; const Layout& L = p.L;
  const int t = blockIdx.x;
  const int fc = p.trip_tile_center[t];
  const bool mine = (p.nranks <= 1) || (fc % p.nranks == p.rank) || MODE == 2;
  double cost = 0.0;
  if ((int)threadIdx.x < p.trip_tile_count[t] && mine) {
    const float* rec = p.trip_records + (size_t)(p.trip_tile_begin[t] + threadIdx.x) * 10;
    const int np = L.npad;
    Gather dg[3], sg[3];
    ObsIn o[3]; const double* pose[3]; double phi[3], D[3], u[3][2];
    for (int i = 0; i < 3; ++i) {
      o[i] = ObsIn{rec[3 * i], rec[3 * i + 1], rec[3 * i + 2]};
      gather_depth(c, o[i].ndcx, o[i].ndcy, dg[i]); gather_spatial(c, o[i].ndcx, o[i].ndcy, sg[i]);
      pose[i] = x + (size_t)(fc - 1 + i) * L.nf;
    }
    if (MODE == 2) {
      for (int i = 0; i < 3; ++i) {
        uint8_t* m = mask + (size_t)(fc - 1 + i) * np;
        for (int q = 0; q < 6; ++q) m[q] = 1;
        if (c.intr_opt == RCVD_INTR_PER_FRAME) m[6] = 1;
        for (int q = 0; q < dg[i].n; ++q) for (int j = 0; j < L.k; ++j) m[L.offD + dg[i].idx[q] * L.k + j] = 1;
        for (int q = 0; q < sg[i].n; ++q) { m[L.offS + sg[i].idx[q] * 2] = 1; m[L.offS + sg[i].idx[q] * 2 + 1] = 1; }
      }
      if (c.intr_opt == RCVD_INTR_SHARED) mask[6] = 1;
      return;
    }
    for (int i = 0; i < 3; ++i) {
      phi[i] = (c.intr_opt == RCVD_INTR_SHARED) ? x[6] : (c.intr_opt == RCVD_INTR_PER_FRAME ? pose[i][6] : c.fixed_vfocal);
      D[i] = depth_value(c, L, dg[i], o[i].depth, pose[i]);
      warp_value(L, sg[i], pose[i], u[i]);
    }
    const double w = (double)rec[9], sw = sqrt(w);   // ScaledLoss(nullptr, w): rho' = w
    double r[3];
    if (MODE == 0) {
      smooth_scene<false>(c, pose, phi, D, u, o, r, nullptr);
      cost = 0.5 * w * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    } else {
      double Jl[90];
      smooth_scene<true>(c, pose, phi, D, u, o, r, Jl);
      cost = 0.5 * w * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      const double r0 = r[0] * sw, r1 = r[1] * sw, r2 = r[2] * sw;
      int ef[kMaxEntries + 72]; short el[kMaxEntries + 72]; double ej[kMaxEntries + 72][3];
      int E = 0;
      auto push = [&](int f, int l, double a0, double a1, double a2) {
        if (is_const_local(c, L, l)) return;
        ef[E] = f; el[E] = (short)l; ej[E][0] = a0 * sw; ej[E][1] = a1 * sw; ej[E][2] = a2 * sw; ++E;
      };
      for (int side = 0; side < 3; ++side) {
        const int f = fc - 1 + side, ofs = side * 10;
        const double src = (double)o[side].depth;
        for (int q = 0; q < 6; ++q) push(f, q, Jl[ofs + q], Jl[30 + ofs + q], Jl[60 + ofs + q]);
        if (c.intr_opt == RCVD_INTR_PER_FRAME) push(f, 6, Jl[ofs + 6], Jl[30 + ofs + 6], Jl[60 + ofs + 6]);
        for (int q = 0; q < dg[side].n; ++q) {
          const double wq = dg[side].w[q];
          push(f, L.offD + dg[side].idx[q] * L.k, Jl[ofs + 7] * wq * src, Jl[30 + ofs + 7] * wq * src, Jl[60 + ofs + 7] * wq * src);
          if (L.k == 2) push(f, L.offD + dg[side].idx[q] * 2 + 1, Jl[ofs + 7] * wq, Jl[30 + ofs + 7] * wq, Jl[60 + ofs + 7] * wq);
        }
        for (int q = 0; q < sg[side].n; ++q) {
          const double wq = sg[side].w[q];
          push(f, L.offS + sg[side].idx[q] * 2, Jl[ofs + 8] * wq, Jl[30 + ofs + 8] * wq, Jl[60 + ofs + 8] * wq);
          push(f, L.offS + sg[side].idx[q] * 2 + 1, Jl[ofs + 9] * wq, Jl[30 + ofs + 9] * wq, Jl[60 + ofs + 9] * wq);
        }
      }
      if (c.intr_opt == RCVD_INTR_SHARED) push(0, 6, Jl[6] + Jl[16] + Jl[26], Jl[36] + Jl[46] + Jl[56], Jl[66] + Jl[76] + Jl[86]);
      for (int a = 0; a < E; ++a) {
        const double a0 = ej[a][0], a1 = ej[a][1], a2 = ej[a][2];
        red_add(g + (size_t)ef[a] * np + el[a], a0 * r0 + a1 * r1 + a2 * r2);
        if (MODE == 1) for (int b = 0; b <= a; ++b) add_h(p, H, ef[a], el[a], ef[b], el[b], a0 * ej[b][0] + a1 * ej[b][1] + a2 * ej[b][2]);
      }
    }
  }
  if (MODE != 2) block_store_sum(cost, partial + t);
}

// Final deterministic reduction of the per-block partial costs: out[slot] = sum(partial[0..n))
__global__ void __launch_bounds__(1024) k_reduce_partials(const double* __restrict__ partial, int n, double* __restrict__ out, int slot) {
  __shared__ double red[32];
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v += partial[i];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) out[slot] = t;
  }
}

}  // namespace rcvd
