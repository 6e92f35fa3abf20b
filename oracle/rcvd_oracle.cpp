// rcvd_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A double-precision restatement of the reference's temporal-consistency
// optimizer (facebookresearch/robust_cvd, lib/PoseOptimizer.cpp,
// lib/DepthMapTransform.cpp) and of the Ceres Solver semantics it relies on.
//
// PARITY UNPINNED: the reference ships no tests/golden vectors for this path
// and cannot be built here (needs Ceres, Eigen, OpenCV, glog, gflags, Boost --
// all absent, no network).  Ceres itself is neither vendored nor version-pinned
// by the reference (lib/CMakeLists.txt:31-38).  The trust-region / loss /
// autodiff semantics below are restated from Ceres' published algorithm
// (TrustRegionMinimizer, LevenbergMarquardtStrategy, Corrector, CauchyLoss,
// rotation.h AngleAxisRotatePoint, Jet) and anchored on the reference's call
// sites cited at each function.  Self-checks: the literal Jet-autodiff
// evaluation (mirrors DynamicAutoDiffCostFunction<.,4>, lib/PoseOptimizer.cpp:1198)
// is compared against the independent analytic Jacobians and finite differences
// in tests/.  What IS pinned by reference code that runs here: the camera model / sign / NDC / aspect conventions of
// StaticSceneCost against the reference's own Python model (utils/geometry.py + loaders/video_dataset.py:177-189, golden
// tests/golden/ref_python_geometry.npz), the hierarchical2 pair sampler (utils/frame_sampling.py) and the .raw wire format
// (utils/image_io.py).  The solver semantics remain unpinned.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may load this library.  The product (robust_cvd_b200/) never does.
//
// Build: see oracle/Makefile (g++ -O3 -fopenmp).
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>
#include <omp.h>

#include "../include/rcvd.h"  // interface structs only (rcvd_config, options, summary)

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace orc {

// ---------------------------------------------------------------------------
// Jet<N>: forward-mode dual number, restating ceres::Jet<double, N>
// (reference use: lib/ValueTransform.h:31-34, kStride = 4).
// ---------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  explicit Jet(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double x, int k) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};
#define JET_BIN(op, body_a, body_v)                                                  \
  template <int N> inline Jet<N> operator op(const Jet<N>& f, const Jet<N>& g) {     \
    Jet<N> h; h.a = body_a; for (int i = 0; i < N; ++i) h.v[i] = body_v; return h; }
JET_BIN(+, f.a + g.a, f.v[i] + g.v[i])
JET_BIN(-, f.a - g.a, f.v[i] - g.v[i])
JET_BIN(*, f.a * g.a, f.a * g.v[i] + f.v[i] * g.a)
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  // Ceres: g_inv = 1/g.a; f/g = (f.a*g_inv, (f.v - f.a*g_inv*g.v)*g_inv)
  Jet<N> h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& f) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) { Jet<N> h = -f; h.a += s; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) { Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, double s) { const double si = 1.0 / s; return f * si; }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) { Jet<N> h; const double m = -s / (g.a * g.a); h.a = s / g.a; for (int i = 0; i < N; ++i) h.v[i] = m * g.v[i]; return h; }
template <int N> inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }
template <int N> inline Jet<N>& operator*=(Jet<N>& f, double s) { f = f * s; return f; }
template <int N> inline bool operator<(const Jet<N>& f, const Jet<N>& g) { return f.a < g.a; }
template <int N> inline bool operator>(const Jet<N>& f, const Jet<N>& g) { return f.a > g.a; }
template <int N> inline bool operator>(const Jet<N>& f, double g) { return f.a > g; }
template <int N> inline Jet<N> jsqrt(const Jet<N>& f) { Jet<N> h; h.a = std::sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
template <int N> inline Jet<N> jcos(const Jet<N>& f) { Jet<N> h; h.a = std::cos(f.a); const double s = -std::sin(f.a); for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
template <int N> inline Jet<N> jsin(const Jet<N>& f) { Jet<N> h; h.a = std::sin(f.a); const double c = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> inline Jet<N> jlog(const Jet<N>& f) { Jet<N> h; h.a = std::log(f.a); const double t = 1.0 / f.a; for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
inline double jsqrt(double x) { return std::sqrt(x); }
inline double jcos(double x) { return std::cos(x); }
inline double jsin(double x) { return std::sin(x); }
inline double jlog(double x) { return std::log(x); }
// std::max/std::min/abs semantics on the scalar part (cv pulls std::max/min/abs
// into scope in the reference; Jet comparisons look at .a only):
//   max(a,b) = (a < b) ? b : a ; min(a,b) = (b < a) ? b : a ; abs(x) = x.a < 0 ? -x : x
template <class T> inline T jmax(const T& a, const T& b) { return (a < b) ? b : a; }
template <class T> inline T jmin(const T& a, const T& b) { return (b < a) ? b : a; }
template <int N> inline Jet<N> jabs(const Jet<N>& f) { return f.a < 0.0 ? -f : f; }
inline double jabs(double x) { return x < 0.0 ? -x : x; }
template <class T> inline double scalar(const T& x) { return x.a; }
template <> inline double scalar<double>(const double& x) { return x; }

// ---------------------------------------------------------------------------
// ceres::AngleAxisRotatePoint restated (ceres/rotation.h); reference call
// sites lib/PoseOptimizer.cpp:185, :211.
// ---------------------------------------------------------------------------
template <class T>
inline void angleAxisRotatePoint(const T aa[3], const T pt[3], T out[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const T theta = jsqrt(theta2);
    const T costheta = jcos(theta);
    const T sintheta = jsin(theta);
    const T theta_inverse = 1.0 / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (1.0 - costheta);
    out[0] = pt[0] * costheta + wxp[0] * sintheta + w[0] * tmp;
    out[1] = pt[1] * costheta + wxp[1] * sintheta + w[1] * tmp;
    out[2] = pt[2] * costheta + wxp[2] * sintheta + w[2] * tmp;
  } else {
    const T wxp[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
    out[0] = pt[0] + wxp[0]; out[1] = pt[1] + wxp[1]; out[2] = pt[2] + wxp[2];
  }
}

// ---------------------------------------------------------------------------
// Layout of one frame's parameters.
// ---------------------------------------------------------------------------
struct Layout {
  int k = 0, G = 0, nd = 0, S = 0, ns = 0, nf = 0, offD = 7, offS = 7;
  bool ok = false;
};
inline Layout makeLayout(const rcvd_config& c) {
  Layout L;
  L.k = (c.value_xform == RCVD_VALUE_SCALESHIFT) ? 2 : 1;
  switch (c.depth_type) {
    case RCVD_DEPTH_IDENTITY: L.G = 0; break;
    case RCVD_DEPTH_GLOBAL: L.G = 1; break;
    case RCVD_DEPTH_GRID:
      if (c.depth_grid_x < 2 || c.depth_grid_y < 2) return L;
      // linearGather indexes params_[i] instead of params_[i*k] (lib/DepthMapTransform.cpp:801-808,
      // :829-832): overlapping Ceres blocks for k=2 -> unsupported in the reference too.
      if (L.k == 2 && !c.depth_cubic) return L;
      L.G = c.depth_grid_x * c.depth_grid_y; break;
    default: return L;
  }
  if (c.depth_type != RCVD_DEPTH_IDENTITY && c.value_xform != RCVD_VALUE_SCALE && c.value_xform != RCVD_VALUE_SCALESHIFT) return L;
  L.nd = L.G * L.k;
  switch (c.spatial_type) {
    case RCVD_SPATIAL_IDENTITY: L.S = 0; break;
    case RCVD_SPATIAL_VERTICAL_LINEAR: L.S = 2; break;
    case RCVD_SPATIAL_CORNERS_BILINEAR: L.S = 4; break;
    case RCVD_SPATIAL_BILINEAR_GRID: case RCVD_SPATIAL_BICUBIC_GRID:
      if (c.spatial_grid_x < 2 || c.spatial_grid_y < 2) return L;
      L.S = c.spatial_grid_x * c.spatial_grid_y; break;
    default: return L;
  }
  L.ns = 2 * L.S;
  L.offD = 7; L.offS = 7 + L.nd; L.nf = 7 + L.nd + L.ns; L.ok = true;
  return L;
}

// ---------------------------------------------------------------------------
// Gathers (node index + weight lists).
// ---------------------------------------------------------------------------
struct Gather { int n = 0; int idx[16]; double w[16]; };

// Cell coordinates: lib/DepthMapTransform.cpp:751-764 (same at :868-881, :1257-1271,
// :1293-1308).  loc is float32, all arithmetic in double.
__attribute__((optimize("fp-contract=off")))
static inline void cellCoord(float loc, int g, int& i, double& r) {
  const double maxc = std::nextafter(double(g - 1), 0.0);
  double s = (double(loc) + 1.0) * (g - 1) / 2.0;
  s = std::min(std::max(s, 0.0), maxc);   // std::clamp(v, lo, hi)
  i = static_cast<int>(s);
  r = s - i;
}
// cubicSpline, lib/DepthMapTransform.cpp:671-678
__attribute__((optimize("fp-contract=off")))
static inline void cubicSpline(double w[4], double t) {
  const double t2 = t * t, t3 = t2 * t;
  w[0] = -0.5 * t3 + t2 - 0.5 * t;
  w[1] = 1.5 * t3 - 2.5 * t2 + 1.0;
  w[2] = -1.5 * t3 + 2.0 * t2 + 0.5 * t;
  w[3] = 0.5 * t3 - 0.5 * t2;
}
// bilinear: linearGather spatial-only branch (:822-840), bilinearSpatialGridGather (:1253-1286)
__attribute__((optimize("fp-contract=off")))
static void gatherBilinear(float lx, float ly, int gx, int gy, Gather& g) {
  int ix, iy; double rx, ry;
  cellCoord(lx, gx, ix, rx); cellCoord(ly, gy, iy, ry);
  g.n = 4;
  g.idx[0] = ix + iy * gx; g.idx[1] = (ix + 1) + iy * gx; g.idx[2] = ix + (iy + 1) * gx; g.idx[3] = (ix + 1) + (iy + 1) * gx;
  g.w[0] = (1.0 - rx) * (1.0 - ry); g.w[1] = rx * (1.0 - ry); g.w[2] = (1.0 - rx) * ry; g.w[3] = rx * ry;
}
// bicubic with border folding: cubicGather (:853-948, gz == 1), bicubicSpatialGridGather (:1288-1343)
__attribute__((optimize("fp-contract=off")))
static void gatherBicubic(float lx, float ly, int gx, int gy, Gather& g) {
  int ix, iy; double rx, ry;
  cellCoord(lx, gx, ix, rx); cellCoord(ly, gy, iy, ry);
  double wx[4], wy[4];
  cubicSpline(wx, rx); cubicSpline(wy, ry);
  const int x0 = (ix == 0 ? 1 : 0), x1 = (ix == gx - 2 ? 3 : 4);
  const int y0 = (iy == 0 ? 1 : 0), y1 = (iy == gy - 2 ? 3 : 4);
  const int xs = x1 - x0, ys = y1 - y0;
  g.n = 0;
  for (int y = y0; y < y1; ++y) for (int x = x0; x < x1; ++x) {
    g.idx[g.n] = (ix - 1 + x) + (iy - 1 + y) * gx; g.w[g.n] = 0.0; ++g.n;
  }
  for (int y = 0; y < 4; ++y) for (int x = 0; x < 4; ++x) {
    const int cx = std::min(std::max(x - x0, 0), xs - 1);
    const int cy = std::min(std::max(y - y0, 0), ys - 1);
    g.w[cx + cy * xs] += wx[x] * wy[y];
  }
}
static void gatherDepth(const rcvd_config& c, float lx, float ly, Gather& g) {
  switch (c.depth_type) {
    case RCVD_DEPTH_IDENTITY: g.n = 0; break;
    case RCVD_DEPTH_GLOBAL: g.n = 1; g.idx[0] = 0; g.w[0] = 1.0; break;  // GlobalDepthFunctor (:495-523): no weight
    default:
      if (c.depth_cubic) gatherBicubic(lx, ly, c.depth_grid_x, c.depth_grid_y, g);
      else gatherBilinear(lx, ly, c.depth_grid_x, c.depth_grid_y, g);
  }
}
__attribute__((optimize("fp-contract=off")))
static void gatherSpatial(const rcvd_config& c, float lx, float ly, Gather& g) {
  switch (c.spatial_type) {
    case RCVD_SPATIAL_IDENTITY: g.n = 0; break;
    case RCVD_SPATIAL_VERTICAL_LINEAR: {  // :1107-1114
      const double w0 = 0.5 + 0.5 * double(ly);
      g.n = 2; g.idx[0] = 0; g.idx[1] = 1; g.w[0] = w0; g.w[1] = 1.0 - w0; break; }
    case RCVD_SPATIAL_CORNERS_BILINEAR: {  // :1181-1191
      const double wx = 0.5 + 0.5 * double(lx), wy = 0.5 + 0.5 * double(ly);
      g.n = 4; for (int i = 0; i < 4; ++i) g.idx[i] = i;
      g.w[0] = wx * wy; g.w[1] = (1.0 - wx) * wy; g.w[2] = wx * (1.0 - wy); g.w[3] = (1.0 - wx) * (1.0 - wy); break; }
    case RCVD_SPATIAL_BILINEAR_GRID: gatherBilinear(lx, ly, c.spatial_grid_x, c.spatial_grid_y, g); break;
    default: gatherBicubic(lx, ly, c.spatial_grid_x, c.spatial_grid_y, g);
  }
}

// ---------------------------------------------------------------------------
// Functors, literal restatement on a generic scalar T.
// ---------------------------------------------------------------------------
// ValueXform (lib/ValueTransform.h:57-94) + Grid/Global/Identity depth functors
// (lib/DepthMapTransform.cpp:457-523, :597-606).  `p` points at the functor's
// parameter blocks laid out contiguously (g.n blocks of k).
template <class T>
inline T depthFunctor(const rcvd_config& c, int k, const Gather& g, float srcDepth, const T* p) {
  const T src = T(static_cast<double>(srcDepth));
  if (c.depth_type == RCVD_DEPTH_IDENTITY) return src;
  if (c.depth_type == RCVD_DEPTH_GLOBAL) return (k == 2) ? (src * p[0]) + p[1] : src * p[0];
  T res(0.0);
  for (int i = 0; i < g.n; ++i) {
    const T v = (k == 2) ? (src * p[i * 2]) + p[i * 2 + 1] : src * p[i];
    res += v * T(g.w[i]);
  }
  return res;
}
// Spatial functors (:1036-1045, :1075-1085, :1146-1160, :1225-1233)
template <class T>
inline void spatialFunctor(const Gather& g, const T* p, T out[2]) {
  out[0] = T(0.0); out[1] = T(0.0);
  for (int i = 0; i < g.n; ++i) { out[0] += p[2 * i] * T(g.w[i]); out[1] += p[2 * i + 1] * T(g.w[i]); }
}

struct ObsData { float ndcx, ndcy, depth; Gather dg, sg; };

// cameraToWorld, lib/PoseOptimizer.cpp:175-192
template <class T>
inline void cameraToWorld(const T pc[3], const T focal[2], const T* pose, T out[3]) {
  T dirCam[3] = {pc[0] * focal[0], pc[1] * focal[1], T(-1.0)};
  T dirWorld[3];
  angleAxisRotatePoint(pose + 3, dirCam, dirWorld);
  out[0] = pose[0] + dirWorld[0] * pc[2];
  out[1] = pose[1] + dirWorld[1] * pc[2];
  out[2] = pose[2] + dirWorld[2] * pc[2];
}
// worldToCamera, lib/PoseOptimizer.cpp:196-221
template <class T>
inline void worldToCamera(const T pw[3], const T focal[2], const T* pose, T out[3]) {
  T rel[3], inv[3], pcam[3];
  for (int i = 0; i < 3; ++i) rel[i] = pw[i] - pose[i];
  for (int i = 0; i < 3; ++i) inv[i] = -pose[i + 3];
  angleAxisRotatePoint(inv, rel, pcam);
  const T depth = -pcam[2];
  out[0] = pcam[0] / depth / focal[0];
  out[1] = pcam[1] / depth / focal[1];
  out[2] = depth;
}

// StaticSceneCost::operator(), lib/PoseOptimizer.cpp:237-308.  x is the stacked
// local parameter vector in Ceres block order:
//   [pose0(6), depth0 blocks, spatial0 blocks, pose1(6), depth1 blocks, spatial1 blocks, focal(s)]
template <class T>
inline void staticSceneCost(const rcvd_config& c, int k, const ObsData& o0, const ObsData& o1, const T* x, T r[3]) {
  int off = 0;
  const T* pose0 = x + off; off += 6;
  const T* d0 = x + off; off += o0.dg.n * k;
  const T* s0 = x + off; off += o0.sg.n * 2;
  const T* pose1 = x + off; off += 6;
  const T* d1 = x + off; off += o1.dg.n * k;
  const T* s1 = x + off; off += o1.sg.n * 2;
  T focal0[2], focal1[2];
  if (c.intr_opt == RCVD_INTR_SHARED) { focal0[1] = focal1[1] = x[off++]; }
  else if (c.intr_opt == RCVD_INTR_PER_FRAME) { focal0[1] = x[off++]; focal1[1] = x[off++]; }
  else { focal0[1] = focal1[1] = T(c.fixed_vfocal); }
  focal0[0] = focal0[1] * c.aspect; focal1[0] = focal1[1] * c.aspect;

  // obsToCamera, :163-171
  T warp0[2], warp1[2];
  const T depth0 = depthFunctor(c, k, o0.dg, o0.depth, d0);
  spatialFunctor(o0.sg, s0, warp0);
  T pc0[3] = {T(double(o0.ndcx)) + warp0[0], T(double(o0.ndcy)) + warp0[1], depth0};
  T pw0[3];
  cameraToWorld(pc0, focal0, pose0, pw0);
  const T depth1 = depthFunctor(c, k, o1.dg, o1.depth, d1);
  spatialFunctor(o1.sg, s1, warp1);
  T pc1[3] = {T(double(o1.ndcx)) + warp1[0], T(double(o1.ndcy)) + warp1[1], depth1};

  if (c.static_loss_type == RCVD_LOSS_EUCLIDEAN) {
    T pw1[3];
    cameraToWorld(pc1, focal1, pose1, pw1);
    for (int i = 0; i < 3; ++i) r[i] = pw1[i] - pw0[i];
    return;
  }
  T p01[3];
  worldToCamera(pw0, focal1, pose1, p01);
  r[0] = (p01[0] - pc1[0]) * T(c.static_spatial_weight);
  r[1] = (p01[1] - pc1[1]) * T(c.static_spatial_weight);
  if (c.static_loss_type == RCVD_LOSS_REPRO_DISPARITY) {
    const T eps(1e-6);
    const T reproDisp = 1.0 / jmax(p01[2], eps);
    const T disp1 = 1.0 / jmax(pc1[2], eps);
    r[2] = (reproDisp - disp1) * T(c.static_depth_weight);
  } else {
    const T maxDepth = jmax(p01[2], pc1[2]);
    const T minDepth = jmin(p01[2], pc1[2]);
    if (c.static_loss_type == RCVD_LOSS_REPRO_DEPTH_RATIO) r[2] = (maxDepth / minDepth - 1.0) * c.static_depth_weight;
    else r[2] = jlog(minDepth / maxDepth) * c.static_depth_weight;
  }
}

// SceneFlowSmoothnessLoss::operator(), lib/PoseOptimizer.cpp:332-413.  x stacked in Ceres block order:
//   [pose0, depth0 blocks, spatial0 blocks, pose1, ..., pose2, ..., focal(s)]
template <class T>
inline void sceneFlowSmoothnessLoss(const rcvd_config& c, int k, const ObsData& o0, const ObsData& o1, const ObsData& o2, const T* x, T r[3]) {
  const ObsData* obs[3] = {&o0, &o1, &o2};
  const T* pose[3]; T pc[3][3];
  int off = 0;
  for (int i = 0; i < 3; ++i) {
    pose[i] = x + off; off += 6;
    const T* d = x + off; off += obs[i]->dg.n * k;
    const T* s = x + off; off += obs[i]->sg.n * 2;
    T warp[2];
    const T depth = depthFunctor(c, k, obs[i]->dg, obs[i]->depth, d);
    spatialFunctor(obs[i]->sg, s, warp);
    pc[i][0] = T(double(obs[i]->ndcx)) + warp[0]; pc[i][1] = T(double(obs[i]->ndcy)) + warp[1]; pc[i][2] = depth;
  }
  T focal[3][2];
  if (c.intr_opt == RCVD_INTR_SHARED) { focal[0][1] = focal[1][1] = focal[2][1] = x[off++]; }
  else if (c.intr_opt == RCVD_INTR_PER_FRAME) { focal[0][1] = x[off++]; focal[1][1] = x[off++]; focal[2][1] = x[off++]; }
  else { focal[0][1] = focal[1][1] = focal[2][1] = T(c.fixed_vfocal); }
  for (int i = 0; i < 3; ++i) focal[i][0] = focal[i][1] * c.aspect;
  const int lt = c.smooth_loss_type;   // SmoothLossType
  if (lt == 0) {   // EuclideanLaplacian
    T w0[3], w1[3], w2[3];
    cameraToWorld(pc[0], focal[0], pose[0], w0); cameraToWorld(pc[1], focal[1], pose[1], w1); cameraToWorld(pc[2], focal[2], pose[2], w2);
    for (int i = 0; i < 3; ++i) r[i] = w0[i] + w2[i] - 2.0 * w1[i];
    return;
  }
  T w0[3], w2[3], p01[3], p21[3];
  cameraToWorld(pc[0], focal[0], pose[0], w0); cameraToWorld(pc[2], focal[2], pose[2], w2);
  worldToCamera(w0, focal[1], pose[1], p01); worldToCamera(w2, focal[1], pose[1], p21);
  r[0] = (p01[0] + p21[0] - pc[1][0] * 2.0) / focal[1][1];
  r[1] = (p01[1] + p21[1] - pc[1][1] * 2.0) / focal[1][1];
  if (lt == 1) {   // ReproDisparityLaplacian
    const T eps(1e-6);
    const T a = 1.0 / jmax(p01[2], eps), b = 1.0 / jmax(pc[1][2], eps), cc = 1.0 / jmax(p21[2], eps);
    r[2] = a + cc - b * 2.0;
  } else {
    const T base = pc[1][2];
    const T other = p01[2] + p21[2] - pc[1][2];
    const T mx = jmax(base, other), mn = jmin(base, other);
    if (lt == 2) r[2] = (mx / mn - 1.0); else r[2] = jlog(mn / mx);
  }
}

// ---------------------------------------------------------------------------
// Analytic Jacobian of the same functor (independent derivation, used for the
// fast CPU baseline "B-analytic" and cross-checked against the Jet version).
// Local variables per frame: t(3), w(3), phi, D, u(2)  -> 10 columns.
// ---------------------------------------------------------------------------
// f = R(v) p, R (row-major 3x3), dfdv[i][j] = d f_i / d v_j ; both branches of
// angleAxisRotatePoint differentiated exactly.
static inline void rotJac(const double v[3], const double p[3], double f[3], double R[9], double dfdv[9]) {
  const double th2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  if (th2 > std::numeric_limits<double>::epsilon()) {
    const double th = std::sqrt(th2), c = std::cos(th), s = std::sin(th), ti = 1.0 / th;
    const double k[3] = {v[0] * ti, v[1] * ti, v[2] * ti};
    const double kxp[3] = {k[1] * p[2] - k[2] * p[1], k[2] * p[0] - k[0] * p[2], k[0] * p[1] - k[1] * p[0]};
    const double kp = k[0] * p[0] + k[1] * p[1] + k[2] * p[2];
    const double omc = 1.0 - c;
    for (int i = 0; i < 3; ++i) f[i] = p[i] * c + kxp[i] * s + k[i] * kp * omc;
    // R = c I + s [k]x + (1-c) k k^T
    R[0] = c + omc * k[0] * k[0]; R[1] = -s * k[2] + omc * k[0] * k[1]; R[2] = s * k[1] + omc * k[0] * k[2];
    R[3] = s * k[2] + omc * k[1] * k[0]; R[4] = c + omc * k[1] * k[1]; R[5] = -s * k[0] + omc * k[1] * k[2];
    R[6] = -s * k[1] + omc * k[2] * k[0]; R[7] = s * k[0] + omc * k[2] * k[1]; R[8] = c + omc * k[2] * k[2];
    for (int j = 0; j < 3; ++j) {
      // dk/dv_j = (e_j - k k_j)/th ; dth/dv_j = k_j
      double dk[3] = {-k[0] * k[j] * ti, -k[1] * k[j] * ti, -k[2] * k[j] * ti};
      dk[j] += ti;
      const double dkxp[3] = {dk[1] * p[2] - dk[2] * p[1], dk[2] * p[0] - dk[0] * p[2], dk[0] * p[1] - dk[1] * p[0]};
      const double dkp = dk[0] * p[0] + dk[1] * p[1] + dk[2] * p[2];
      const double dc = -s * k[j], ds = c * k[j];
      for (int i = 0; i < 3; ++i)
        dfdv[i * 3 + j] = p[i] * dc + dkxp[i] * s + kxp[i] * ds + dk[i] * kp * omc + k[i] * dkp * omc - k[i] * kp * dc;
    }
  } else {
    f[0] = p[0] + (v[1] * p[2] - v[2] * p[1]);
    f[1] = p[1] + (v[2] * p[0] - v[0] * p[2]);
    f[2] = p[2] + (v[0] * p[1] - v[1] * p[0]);
    R[0] = 1; R[1] = -v[2]; R[2] = v[1]; R[3] = v[2]; R[4] = 1; R[5] = -v[0]; R[6] = -v[1]; R[7] = v[0]; R[8] = 1;
    // d(v x p)/dv_j = e_j x p
    dfdv[0] = 0; dfdv[1] = p[2]; dfdv[2] = -p[1];
    dfdv[3] = -p[2]; dfdv[4] = 0; dfdv[5] = p[0];
    dfdv[6] = p[1]; dfdv[7] = -p[0]; dfdv[8] = 0;
  }
}

// X = t + R(w) (pcx*phi*a, pcy*phi, -1) D ; dX[3][10] wrt (t, w, phi, D, u)
static inline void camToWorldJac(const double* pose, double phi, double a, double pcx, double pcy, double D,
                                 double X[3], double dX[30]) {
  const double p[3] = {pcx * phi * a, pcy * phi, -1.0};
  double w[3], R[9], dw[9];
  rotJac(pose + 3, p, w, R, dw);
  for (int i = 0; i < 3; ++i) {
    X[i] = pose[i] + w[i] * D;
    double* row = dX + i * 10;
    row[0] = row[1] = row[2] = 0.0; row[i] = 1.0;
    row[3] = D * dw[i * 3 + 0]; row[4] = D * dw[i * 3 + 1]; row[5] = D * dw[i * 3 + 2];
    row[6] = D * (R[i * 3 + 0] * pcx * a + R[i * 3 + 1] * pcy);
    row[7] = w[i];
    row[8] = D * R[i * 3 + 0] * phi * a;
    row[9] = D * R[i * 3 + 1] * phi;
  }
}

// r[3], Jl[3][20] (local columns: frame0 then frame1)
static inline void staticSceneAnalytic(const rcvd_config& c, const double* pose0, double phi0, double D0, const double u0[2],
                                       const double* pose1, double phi1, double D1, const double u1[2],
                                       const ObsData& o0, const ObsData& o1, double r[3], double Jl[60]) {
  const double a = c.aspect;
  double X0[3], dX0[30];
  camToWorldJac(pose0, phi0, a, double(o0.ndcx) + u0[0], double(o0.ndcy) + u0[1], D0, X0, dX0);
  const double pc1x = double(o1.ndcx) + u1[0], pc1y = double(o1.ndcy) + u1[1];
  for (int i = 0; i < 60; ++i) Jl[i] = 0.0;
  if (c.static_loss_type == RCVD_LOSS_EUCLIDEAN) {
    double X1[3], dX1[30];
    camToWorldJac(pose1, phi1, a, pc1x, pc1y, D1, X1, dX1);
    for (int i = 0; i < 3; ++i) {
      r[i] = X1[i] - X0[i];
      for (int j = 0; j < 10; ++j) { Jl[i * 20 + j] = -dX0[i * 10 + j]; Jl[i * 20 + 10 + j] = dX1[i * 10 + j]; }
    }
    return;
  }
  const double rel[3] = {X0[0] - pose1[0], X0[1] - pose1[1], X0[2] - pose1[2]};
  const double v[3] = {-pose1[3], -pose1[4], -pose1[5]};
  double q[3], R1[9], dq[9];
  rotJac(v, rel, q, R1, dq);
  // dq/d(local0) = R1 * dX0 ; dq/dt1 = -R1 ; dq/dw1 = -dq/dv
  double Q[3][20];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 10; ++j)
      Q[i][j] = R1[i * 3 + 0] * dX0[0 * 10 + j] + R1[i * 3 + 1] * dX0[1 * 10 + j] + R1[i * 3 + 2] * dX0[2 * 10 + j];
    for (int j = 0; j < 3; ++j) { Q[i][10 + j] = -R1[i * 3 + j]; Q[i][13 + j] = -dq[i * 3 + j]; }
    Q[i][16] = Q[i][17] = Q[i][18] = Q[i][19] = 0.0;
  }
  const double depth = -q[2];
  const double fx1 = phi1 * a, fy1 = phi1;
  const double projx = q[0] / depth / fx1, projy = q[1] / depth / fy1;
  const double ws = c.static_spatial_weight, wd = c.static_depth_weight;
  r[0] = (projx - pc1x) * ws; r[1] = (projy - pc1y) * ws;
  // d projx = dq0/(depth fx1) + q0/(depth^2 fx1) * dq2   (ddepth = -dq2)
  for (int j = 0; j < 20; ++j) {
    Jl[0 * 20 + j] = ws * (Q[0][j] / (depth * fx1) + q[0] / (depth * depth * fx1) * Q[2][j]);
    Jl[1 * 20 + j] = ws * (Q[1][j] / (depth * fy1) + q[1] / (depth * depth * fy1) * Q[2][j]);
  }
  Jl[0 * 20 + 16] += ws * (-projx / phi1);
  Jl[1 * 20 + 16] += ws * (-projy / phi1);
  Jl[0 * 20 + 18] += -ws; Jl[1 * 20 + 19] += -ws;
  // depth term: A = depth (function of q2), B = D1
  double dA = 0.0, dB = 0.0;  // d r2 / dA, d r2 / dB
  const double A = depth, B = D1;
  if (c.static_loss_type == RCVD_LOSS_REPRO_DISPARITY) {
    const double eps = 1e-6;
    const double Am = (A < eps) ? eps : A, Bm = (B < eps) ? eps : B;
    r[2] = (1.0 / Am - 1.0 / Bm) * wd;
    dA = (A < eps) ? 0.0 : -wd / (A * A);
    dB = (B < eps) ? 0.0 : wd / (B * B);
  } else {
    // max = (A<B)?B:A ; min = (B<A)?B:A
    const bool mxB = (A < B), mnB = (B < A);
    const double mx = mxB ? B : A, mn = mnB ? B : A;
    // partials of mx, mn wrt A,B
    const double mxA = mxB ? 0.0 : 1.0, mxBd = mxB ? 1.0 : 0.0, mnA = mnB ? 0.0 : 1.0, mnBd = mnB ? 1.0 : 0.0;
    if (c.static_loss_type == RCVD_LOSS_REPRO_DEPTH_RATIO) {
      r[2] = (mx / mn - 1.0) * wd;
      dA = wd * (mxA / mn - mx / (mn * mn) * mnA);
      dB = wd * (mxBd / mn - mx / (mn * mn) * mnBd);
    } else {
      r[2] = std::log(mn / mx) * wd;
      dA = wd * (mnA / mn - mxA / mx);
      dB = wd * (mnBd / mn - mxBd / mx);
    }
  }
  for (int j = 0; j < 20; ++j) Jl[2 * 20 + j] = dA * (-Q[2][j]);
  Jl[2 * 20 + 17] += dB;
}

// ---------------------------------------------------------------------------
// Problem
// ---------------------------------------------------------------------------
struct Problem {
  rcvd_config cfg;
  Layout L;
  int N = 0;
  std::vector<uint8_t> inRange;
  std::vector<double> median, adaptive;
  std::vector<int32_t> pairFrames;
  std::vector<int64_t> offsets;
  std::vector<float> records;
  std::vector<int32_t> tripCenters; std::vector<int64_t> tripOffsets; std::vector<float> tripRecords;   // [n][10]: 3 x (ndc.xy, depth), weight
  std::vector<float> scaleLocs;   // lattice (x,y) float32, computed as lib/PoseOptimizer.cpp:1384-1385
  std::vector<double> x;          // N*nf
  // block-sparse normal matrix: lower blocks (r>=c by frame id), full n x n row-major
  std::map<std::pair<int, int>, std::vector<double>> H;
  std::vector<double> g;
  int jacMode = 0;                // 0 analytic, 1 Jet<4> passes (mirrors DynamicAutoDiffCostFunction<.,4>)
  int U() const { return N * L.nf; }
};

static thread_local std::string g_err;

// Robust loss restated from ceres/loss_function.cc; rho[0..2] at s = |r|^2.
static inline void lossEval(const rcvd_config& c, double s, double rho[3]) {
  if (c.robust_type == RCVD_ROBUST_CAUCHY) {
    const double b = c.robustness * c.robustness, ci = 1.0 / b;
    const double sum = 1.0 + s * ci, inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -ci * (inv * inv);
  } else if (c.robust_type == RCVD_ROBUST_HUBER) {
    const double a = c.robustness, b = a * a;
    if (s > b) {
      const double rr = std::sqrt(s);
      rho[0] = 2.0 * a * rr - b;
      rho[1] = std::max(std::numeric_limits<double>::min(), a / rr);
      rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}
// Corrector (ceres/corrector.cc): rho''<=0 or s==0 -> scale r and J by sqrt(rho').
// Cauchy and Huber always have rho'' <= 0, so the alpha branch never triggers.
static inline double correctorScale(double s, const double rho[3]) { (void)s; return std::sqrt(rho[1]); }

// sparse row: residual value + (col, derivative) list; already weighted.
struct Row { double r; int n; int col[40]; double d[40]; };

struct Accum {
  Problem* P;
  bool wantH, wantG;
  std::vector<omp_lock_t>* locks;  // unused in serial sections
  double* blockPtr(int fr, int fc) {
    auto it = P->H.find({std::max(fr, fc), std::min(fr, fc)});
    return it == P->H.end() ? nullptr : it->second.data();
  }
};

static void ensureBlock(Problem& P, int a, int b) {
  const int n = P.L.nf;
  auto key = std::make_pair(std::max(a, b), std::min(a, b));
  if (!P.H.count(key)) P.H[key].assign(size_t(n) * n, 0.0);
}

// Adds v to H(ci, cj) and (if different) H(cj, ci), in the lower-block storage.
static inline void addH(Problem& P, int ci, int cj, double v) {
  const int n = P.L.nf;
  const int fi = ci / n, fj = cj / n, li = ci % n, lj = cj % n;
  if (fi == fj) {
    double* B = P.H[{fi, fi}].data();
    B[size_t(li) * n + lj] += v;
    if (li != lj) B[size_t(lj) * n + li] += v;
  } else if (fi > fj) {
    P.H[{fi, fj}][size_t(li) * n + lj] += v;
  } else {
    P.H[{fj, fi}][size_t(lj) * n + li] += v;
  }
}

// Scale-regulariser lattice locations in float32, lib/PoseOptimizer.cpp:1382-1385.
__attribute__((optimize("fp-contract=off")))
static void buildScaleLocs(Problem& P) {
  P.scaleLocs.clear();
  const int gx = P.cfg.scale_grid_x, gy = P.cfg.scale_grid_y;
  for (int y = 0; y < gy; ++y) for (int x = 0; x < gx; ++x) {
    const float lx = -1.f + 2.f * x / (gx - 1);
    const float ly = -1.f + 2.f * y / (gy - 1);
    P.scaleLocs.push_back(lx); P.scaleLocs.push_back(ly);
  }
}

// Is parameter `local` (index within a frame) held constant?
static inline bool isConstLocal(const Problem& P, int local) {
  if (local < 6) return P.cfg.fix_poses != 0;
  if (local == 6) return P.cfg.intr_opt == RCVD_INTR_FIXED;
  if (local < P.L.offS) return P.cfg.fix_depth_xforms != 0;
  return P.cfg.fix_spatial_xforms != 0;
}

// Builds the stacked local vector + column map of a static block (Ceres order).
static inline int staticCols(const Problem& P, int f0, int f1, const ObsData& o0, const ObsData& o1, int* cols) {
  const Layout& L = P.L; const int n = L.nf; int m = 0;
  auto pushObs = [&](int f, const ObsData& o) {
    for (int i = 0; i < 6; ++i) cols[m++] = f * n + i;
    for (int i = 0; i < o.dg.n; ++i) for (int j = 0; j < L.k; ++j) cols[m++] = f * n + L.offD + o.dg.idx[i] * L.k + j;
    for (int i = 0; i < o.sg.n; ++i) for (int j = 0; j < 2; ++j) cols[m++] = f * n + L.offS + o.sg.idx[i] * 2 + j;
  };
  pushObs(f0, o0); pushObs(f1, o1);
  if (P.cfg.intr_opt == RCVD_INTR_SHARED) cols[m++] = 0 * n + 6;           // &poseParams_[0][6], :1226
  else if (P.cfg.intr_opt == RCVD_INTR_PER_FRAME) { cols[m++] = f0 * n + 6; cols[m++] = f1 * n + 6; }
  return m;
}

static inline void makeObs(const rcvd_config& c, const float* rec, ObsData& o) {
  o.ndcx = rec[0]; o.ndcy = rec[1]; o.depth = rec[2];
  gatherDepth(c, o.ndcx, o.ndcy, o.dg);
  gatherSpatial(c, o.ndcx, o.ndcy, o.sg);
}

constexpr int kMaxP = 160;

// Evaluates one static block: residual r[3], dense J[3][P] wrt cols. mode 0 analytic, 1 Jet passes.
static int evalStatic(const Problem& P, int f0, int f1, const float* rec, int mode, bool wantJ,
                      double r[3], double* J /*3*kMaxP*/, int* cols) {
  const rcvd_config& c = P.cfg; const Layout& L = P.L; const int n = L.nf;
  ObsData o0, o1; makeObs(c, rec, o0); makeObs(c, rec + 3, o1);
  const int Pn = staticCols(P, f0, f1, o0, o1, cols);
  const double* x = P.x.data();
  if (mode == 1 || !wantJ) {
    double xl[kMaxP];
    for (int i = 0; i < Pn; ++i) xl[i] = x[cols[i]];
    if (!wantJ) { staticSceneCost<double>(c, L.k, o0, o1, xl, r); return Pn; }
    // DynamicAutoDiffCostFunction<., 4>: ceil(P/4) passes, each seeding 4 parameters.
    using J4 = Jet<4>;
    J4 xj[kMaxP], rj[3];
    for (int start = 0; start < Pn; start += 4) {
      for (int i = 0; i < Pn; ++i) xj[i] = J4(xl[i]);
      for (int q = 0; q < 4 && start + q < Pn; ++q) xj[start + q].v[q] = 1.0;
      staticSceneCost<J4>(c, L.k, o0, o1, xj, rj);
      for (int q = 0; q < 4 && start + q < Pn; ++q) for (int i = 0; i < 3; ++i) J[i * kMaxP + start + q] = rj[i].v[q];
    }
    for (int i = 0; i < 3; ++i) r[i] = rj[i].a;
    return Pn;
  }
  // analytic
  const double* p0 = x + size_t(f0) * n; const double* p1 = x + size_t(f1) * n;
  double phi0, phi1;
  if (c.intr_opt == RCVD_INTR_SHARED) phi0 = phi1 = x[6];
  else if (c.intr_opt == RCVD_INTR_PER_FRAME) { phi0 = p0[6]; phi1 = p1[6]; }
  else phi0 = phi1 = c.fixed_vfocal;
  auto depthOf = [&](const double* pf, const ObsData& o) {
    const double src = double(o.depth);
    if (c.depth_type == RCVD_DEPTH_IDENTITY) return src;
    double D = 0.0;
    for (int i = 0; i < o.dg.n; ++i) {
      const double* s = pf + L.offD + o.dg.idx[i] * L.k;
      D += (L.k == 2 ? src * s[0] + s[1] : src * s[0]) * o.dg.w[i];
    }
    return D;
  };
  auto warpOf = [&](const double* pf, const ObsData& o, double u[2]) {
    u[0] = u[1] = 0.0;
    for (int i = 0; i < o.sg.n; ++i) { u[0] += pf[L.offS + o.sg.idx[i] * 2] * o.sg.w[i]; u[1] += pf[L.offS + o.sg.idx[i] * 2 + 1] * o.sg.w[i]; }
  };
  double u0[2], u1[2];
  const double D0 = depthOf(p0, o0), D1 = depthOf(p1, o1);
  warpOf(p0, o0, u0); warpOf(p1, o1, u1);
  double Jl[60];
  staticSceneAnalytic(c, p0, phi0, D0, u0, p1, phi1, D1, u1, o0, o1, r, Jl);
  // expand local -> block columns (same order as staticCols)
  for (int i = 0; i < 3; ++i) {
    int m = 0; double* Ji = J + i * kMaxP; const double* l = Jl + i * 20;
    for (int side = 0; side < 2; ++side) {
      const ObsData& o = side ? o1 : o0; const double* ll = l + side * 10; const double src = double(o.depth);
      for (int q = 0; q < 6; ++q) Ji[m++] = ll[q];
      for (int q = 0; q < o.dg.n; ++q) { Ji[m++] = ll[7] * o.dg.w[q] * src; if (L.k == 2) Ji[m++] = ll[7] * o.dg.w[q]; }
      for (int q = 0; q < o.sg.n; ++q) { Ji[m++] = ll[8] * o.sg.w[q]; Ji[m++] = ll[9] * o.sg.w[q]; }
    }
    if (c.intr_opt == RCVD_INTR_SHARED) Ji[m++] = l[6] + l[16];
    else if (c.intr_opt == RCVD_INTR_PER_FRAME) { Ji[m++] = l[6]; Ji[m++] = l[16]; }
  }
  return Pn;
}

// --- regulariser rows -------------------------------------------------------
// Evaluates all regulariser residual rows of frame f, calling sink(Row&) for each.
// mode 1 reproduces the value/derivative through Jet arithmetic of the literal functor.
template <class Sink>
static void regulariserRows(const Problem& P, int f, int mode, Sink&& sink) {
  const rcvd_config& c = P.cfg; const Layout& L = P.L; const int n = L.nf;
  const double* pf = P.x.data() + size_t(f) * n;
  if (!P.inRange[f]) return;
  Row row;
  // addScaleRegularization + TargetDisparityCost, lib/PoseOptimizer.cpp:488-517, :1341-1415.
  // Not added when fixDepthXforms (:924-938).
  if (c.scale_reg > 0.0 && !c.fix_depth_xforms && (c.depth_type == RCVD_DEPTH_GLOBAL || c.depth_type == RCVD_DEPTH_GRID)) {
    const double sw = std::sqrt(c.scale_reg);  // ScaledLoss(nullptr, w): rho' = w -> sqrt(w) scaling
    const float med = float(P.median[f]);
    const int M = int(P.scaleLocs.size() / 2);
    for (int s = 0; s < M; ++s) {
      Gather g; gatherDepth(c, P.scaleLocs[2 * s], P.scaleLocs[2 * s + 1], g);
      const int Pn = g.n * L.k;
      double xl[32];
      for (int i = 0; i < g.n; ++i) for (int j = 0; j < L.k; ++j) { xl[i * L.k + j] = pf[L.offD + g.idx[i] * L.k + j]; row.col[i * L.k + j] = f * n + L.offD + g.idx[i] * L.k + j; }
      row.n = Pn;
      if (mode == 1) {
        using J4 = Jet<4>; J4 xj[32];
        double val = 0.0;
        for (int start = 0; start < std::max(Pn, 1); start += 4) {
          for (int i = 0; i < Pn; ++i) xj[i] = J4(xl[i]);
          for (int q = 0; q < 4 && start + q < Pn; ++q) xj[start + q].v[q] = 1.0;
          const J4 depth = depthFunctor<J4>(c, L.k, g, med, xj);
          const J4 disp = 1.0 / jmax(depth, J4(1e-6));
          const J4 res = disp - 1.0;
          val = res.a;
          for (int q = 0; q < 4 && start + q < Pn; ++q) row.d[start + q] = res.v[q] * sw;
        }
        row.r = val * sw;
      } else {
        const double depth = depthFunctor<double>(c, L.k, g, med, xl);
        const bool clamped = depth < 1e-6;
        const double dm = clamped ? 1e-6 : depth;
        row.r = (1.0 / dm - 1.0) * sw;
        const double dd = clamped ? 0.0 : -1.0 / (depth * depth);
        for (int i = 0; i < g.n; ++i) {
          if (c.depth_type == RCVD_DEPTH_GLOBAL) { row.d[0] = dd * double(med) * sw; if (L.k == 2) row.d[1] = dd * sw; }
          else { row.d[i * L.k] = dd * g.w[i] * double(med) * sw; if (L.k == 2) row.d[i * L.k + 1] = dd * g.w[i] * sw; }
        }
      }
      sink(row);
    }
  }
  // addDepthDeformRegularization + DeformationCost / AdaptiveDeformationCost
  // (lib/PoseOptimizer.cpp:536-656, :1449-1495) over computeGridDeformationCost
  // (lib/DepthMapTransform.cpp:631-667).  No loss function: residual *= weight.
  if (c.depth_deform_reg > 0.0 && c.depth_type == RCVD_DEPTH_GRID) {
    const int gx = c.depth_grid_x, gy = c.depth_grid_y, k = L.k;
    const double* aw = (c.adaptive_deform > 0.0 && !P.adaptive.empty()) ? P.adaptive.data() + size_t(f) * gx * gy : nullptr;
    auto edge = [&](int a, int b, double w) {
      for (int i = 0; i < k; ++i) {
        const int ca = f * n + L.offD + a * k + i, cb = f * n + L.offD + b * k + i;
        const double va = pf[L.offD + a * k + i], vb = pf[L.offD + b * k + i];
        row.n = 2; row.col[0] = ca; row.col[1] = cb;
        if (mode == 1) {
          using J4 = Jet<4>;
          const J4 ta(va, 0), tb(vb, 1);
          const J4 sc = jmin(jabs(ta), jabs(tb));
          const J4 res = (ta - tb) / sc;
          row.r = res.a * w; row.d[0] = res.v[0] * w; row.d[1] = res.v[1] * w;
        } else {
          const double aa = jabs(va), ab = jabs(vb);
          const bool useB = ab < aa;   // min(abs(this), abs(that)) = (that < this) ? that : this
          const double m = useB ? ab : aa;
          const double e = (va - vb) / m;
          row.r = e * w;
          double da = 1.0 / m, db = -1.0 / m;
          if (useB) db += -(va - vb) / (m * m) * (vb < 0.0 ? -1.0 : 1.0);
          else da += -(va - vb) / (m * m) * (va < 0.0 ? -1.0 : 1.0);
          row.d[0] = da * w; row.d[1] = db * w;
        }
        sink(row);
      }
    };
    for (int y = 0; y < gy; ++y) for (int x = 0; x < gx; ++x) {
      const int a = x + y * gx;
      if (x > 0) { const int b = (x - 1) + y * gx; edge(a, b, aw ? c.depth_deform_reg + std::max(aw[a], aw[b]) * c.adaptive_deform : c.depth_deform_reg); }
      if (y > 0) { const int b = x + (y - 1) * gx; edge(a, b, aw ? c.depth_deform_reg + std::max(aw[a], aw[b]) * c.adaptive_deform : c.depth_deform_reg); }
    }
  }
  // addSpatialDeformRegularization (:1497-1522) + paramsToResiduals (lib/DepthMapTransform.cpp:61-70)
  if (c.spatial_deform_reg > 0.0 && L.ns > 0) {
    for (int i = 0; i < L.ns; ++i) {
      row.n = 1; row.col[0] = f * n + L.offS + i; row.r = pf[L.offS + i] * c.spatial_deform_reg; row.d[0] = c.spatial_deform_reg;
      sink(row);
    }
  }
  // addFocalRegularization + TargetFocalCost (:520-533, :1524-1549)
  if (c.focal_reg > 0.0 && c.intr_opt != RCVD_INTR_FIXED) {
    const double sw = std::sqrt(c.focal_reg);
    row.n = 1; row.col[0] = f * n + 6; row.r = (pf[6] - c.focal_target) * sw; row.d[0] = sw;
    sink(row);
  }
}
// addPositionRegularization + ParameterRegularizationCost (:464-483, :1417-1447): rows for triplet starting at f.
template <class Sink>
static void positionRows(const Problem& P, Sink&& sink) {
  const rcvd_config& c = P.cfg; const int n = P.L.nf;
  if (!(c.position_reg > 0.0)) return;
  int first = -1, last = -1;
  for (int f = 0; f < P.N; ++f) if (P.inRange[f]) { if (first < 0) first = f; last = f; }
  if (first < 0) return;
  const double sw = std::sqrt(c.position_reg);
  Row row;
  // for (frame = firstFrame; frame < lastFrame() - 1; ++frame); FrameRange::lastFrame() is the last
  // in-range frame index (inclusive).
  for (int f = first; f < last - 1; ++f) {
    if (!P.inRange[f] || !P.inRange[f + 1] || !P.inRange[f + 2]) continue;
    for (int i = 0; i < 3; ++i) {
      const double a = P.x[size_t(f) * n + i], b = P.x[size_t(f + 1) * n + i], cc = P.x[size_t(f + 2) * n + i];
      row.n = 3; row.col[0] = f * n + i; row.col[1] = (f + 1) * n + i; row.col[2] = (f + 2) * n + i;
      row.r = (a - 2.0 * b + cc) * sw; row.d[0] = sw; row.d[1] = -2.0 * sw; row.d[2] = sw;
      sink(row);
    }
  }
}

// Scene-flow smoothness blocks (addSceneFlowSmoothnessLoss, lib/PoseOptimizer.cpp:1242-1339): evaluated with Jet<4> passes in
// both Jacobian modes (the literal functor is the only CPU implementation; the CUDA kernel's analytic Jacobian is checked
// against it).  ScaledLoss(nullptr, w): residual and Jacobian scaled by sqrt(w), cost 1/2 w |r|^2.
static int tripletCols(const Problem& P, int fc, const ObsData o[3], int* cols) {
  const Layout& L = P.L; const int n = L.nf; int m = 0;
  for (int i = 0; i < 3; ++i) {
    const int f = fc - 1 + i;
    for (int q = 0; q < 6; ++q) cols[m++] = f * n + q;
    for (int q = 0; q < o[i].dg.n; ++q) for (int j = 0; j < L.k; ++j) cols[m++] = f * n + L.offD + o[i].dg.idx[q] * L.k + j;
    for (int q = 0; q < o[i].sg.n; ++q) for (int j = 0; j < 2; ++j) cols[m++] = f * n + L.offS + o[i].sg.idx[q] * 2 + j;
  }
  if (P.cfg.intr_opt == RCVD_INTR_SHARED) cols[m++] = 6;
  else if (P.cfg.intr_opt == RCVD_INTR_PER_FRAME) { cols[m++] = (fc - 1) * n + 6; cols[m++] = fc * n + 6; cols[m++] = (fc + 1) * n + 6; }
  return m;
}
constexpr int kMaxPT = 240;
static int evalTriplet(const Problem& P, int fc, const float* rec, bool wantJ, double r[3], double* J /*3*kMaxPT*/, int* cols) {
  const rcvd_config& c = P.cfg; const Layout& L = P.L;
  ObsData o[3]; for (int i = 0; i < 3; ++i) makeObs(c, rec + 3 * i, o[i]);
  const int Pn = tripletCols(P, fc, o, cols);
  double xl[kMaxPT]; for (int i = 0; i < Pn; ++i) xl[i] = P.x[cols[i]];
  if (!wantJ) { sceneFlowSmoothnessLoss<double>(c, L.k, o[0], o[1], o[2], xl, r); return Pn; }
  using J4 = Jet<4>; J4 xj[kMaxPT], rj[3];
  for (int start = 0; start < Pn; start += 4) {
    for (int i = 0; i < Pn; ++i) xj[i] = J4(xl[i]);
    for (int q = 0; q < 4 && start + q < Pn; ++q) xj[start + q].v[q] = 1.0;
    sceneFlowSmoothnessLoss<J4>(c, L.k, o[0], o[1], o[2], xj, rj);
    for (int q = 0; q < 4 && start + q < Pn; ++q) for (int i = 0; i < 3; ++i) J[i * kMaxPT + start + q] = rj[i].v[q];
  }
  for (int i = 0; i < 3; ++i) r[i] = rj[i].a;
  return Pn;
}

static void buildStructure(Problem& P) {
  P.H.clear();
  for (int f = 0; f < P.N; ++f) ensureBlock(P, f, f);
  const int np = int(P.pairFrames.size() / 2);
  for (int p = 0; p < np; ++p) {
    const int a = P.pairFrames[2 * p], b = P.pairFrames[2 * p + 1];
    if (P.offsets[p + 1] == P.offsets[p]) continue;
    ensureBlock(P, a, b);
    if (P.cfg.intr_opt == RCVD_INTR_SHARED) { ensureBlock(P, a, 0); ensureBlock(P, b, 0); }
  }
  for (size_t t = 0; t < P.tripCenters.size(); ++t) { const int f = P.tripCenters[t]; if (P.tripOffsets[t + 1] > P.tripOffsets[t]) { ensureBlock(P, f - 1, f); ensureBlock(P, f - 1, f + 1); ensureBlock(P, f, f + 1); if (P.cfg.intr_opt == RCVD_INTR_SHARED) { ensureBlock(P, f - 1, 0); ensureBlock(P, f, 0); ensureBlock(P, f + 1, 0); } } }
  if (P.cfg.position_reg > 0.0)
    for (int f = 0; f + 2 < P.N; ++f) { ensureBlock(P, f, f + 1); ensureBlock(P, f, f + 2); ensureBlock(P, f + 1, f + 2); }
}

// Evaluate: cost (ceres: 1/2 sum rho(|r|^2)), optional gradient g = J^T r and H = J^T J
// (both after loss correction, constant columns zeroed -- Ceres removes constant blocks
// from the reduced program).
static double evaluate(Problem& P, bool wantG, bool wantH, double* costStatic = nullptr) {
  const rcvd_config& c = P.cfg; const int n = P.L.nf;
  const int U = P.U();
  if (wantG) P.g.assign(U, 0.0);
  if (wantH) { if (P.H.empty()) buildStructure(P); for (auto& kv : P.H) std::fill(kv.second.begin(), kv.second.end(), 0.0); }
  std::vector<uint8_t> constCol(n);
  for (int i = 0; i < n; ++i) constCol[i] = isConstLocal(P, i);
  const int np = int(P.pairFrames.size() / 2);
  double cost = 0.0;
  // block locks (one per frame-pair block), only needed for H
  std::map<std::pair<int, int>, int> lockId; std::vector<omp_lock_t> locks;
  if (wantH || wantG) {
    int id = 0; for (auto& kv : P.H) lockId[kv.first] = id++;
    locks.resize(std::max(id, 1)); for (auto& l : locks) omp_init_lock(&l);
  }
  std::vector<omp_lock_t> glocks(wantG ? P.N : 0); for (auto& l : glocks) omp_init_lock(&l);
  // pairs are visited in a strided order: neighbours in the list share their source frame, and the threads of a dynamic schedule would
  // queue on that frame's diagonal-block lock (config 2 on 8 vCPU: 1.05 s -> 0.31 s per evaluation with H)
  int stride = 1; { const int cand[] = {997, 499, 251, 127, 61, 31, 13, 7, 3}; for (int q : cand) if (q < np && np % q != 0) { stride = q; break; } }
  auto gcd = [](int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; };
  if (np > 0 && gcd(stride, np) != 1) stride = 1;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : cost)
  for (int q = 0; q < np; ++q) {
    const int p = int((int64_t(q) * stride) % np);
    const int f0 = P.pairFrames[2 * p], f1 = P.pairFrames[2 * p + 1];
    const int64_t beg = P.offsets[p], end = P.offsets[p + 1];
    if (beg == end) continue;
    const bool needJ = wantG || wantH;
    // local accumulation into thread-private copies of the touched blocks would cost n^2 each;
    // instead lock the blocks for the duration of the pair.
    std::vector<std::pair<int, int>> keys;
    if (wantH) {
      keys.push_back({f0, f0}); keys.push_back({f1, f1}); keys.push_back({std::max(f0, f1), std::min(f0, f1)});
      if (c.intr_opt == RCVD_INTR_SHARED) { keys.push_back({0, 0}); keys.push_back({std::max(f0, 0), 0}); keys.push_back({std::max(f1, 0), 0}); }
      std::sort(keys.begin(), keys.end()); keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
      for (auto& k : keys) omp_set_lock(&locks[lockId[k]]);
    }
    std::vector<double> gl;  // local gradient for the (up to 3) frames
    int gf[3] = {f0, f1, 0}; const int ngf = (c.intr_opt == RCVD_INTR_SHARED) ? 3 : 2;
    if (wantG) gl.assign(size_t(3) * n, 0.0);
    double J[3 * kMaxP]; int cols[kMaxP]; double r[3];
    double* tbl[3][3] = {{nullptr}}; bool tblRowIsA[3][3] = {{false}};
    if (wantH) {
      const int fr3[3] = {f0, f1, 0};
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
        if (!(c.intr_opt == RCVD_INTR_SHARED) && (a == 2 || b == 2)) continue;
        auto it = P.H.find({std::max(fr3[a], fr3[b]), std::min(fr3[a], fr3[b])});
        if (it == P.H.end()) continue;
        tbl[a][b] = it->second.data(); tblRowIsA[a][b] = fr3[a] > fr3[b];
      }
    }
    for (int64_t ci = beg; ci < end; ++ci) {
      const int Pn = evalStatic(P, f0, f1, &P.records[size_t(ci) * 6], P.jacMode, needJ, r, J, cols);
      const double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      double rho[3]; lossEval(c, s, rho);
      cost += 0.5 * rho[0];
      if (!needJ) continue;
      const double sc = correctorScale(s, rho);
      for (int i = 0; i < 3; ++i) r[i] *= sc;
      for (int i = 0; i < 3; ++i) for (int j = 0; j < Pn; ++j) { double& v = J[i * kMaxP + j]; v = constCol[cols[j] % n] ? 0.0 : v * sc; }
      if (wantG) {
        for (int j = 0; j < Pn; ++j) {
          const int fr = cols[j] / n; const int slot = (fr == f0) ? 0 : (fr == f1 ? 1 : 2);
          gl[size_t(slot) * n + cols[j] % n] += J[j] * r[0] + J[kMaxP + j] * r[1] + J[2 * kMaxP + j] * r[2];
        }
      }
      if (wantH) {
        // block pointer table for the (<= 3) frames of this residual block
        int slot[kMaxP];
        for (int a = 0; a < Pn; ++a) { const int fr = cols[a] / n; slot[a] = (fr == f0) ? 0 : (fr == f1 ? 1 : 2); }
        for (int a = 0; a < Pn; ++a) {
          const double ja0 = J[a], ja1 = J[kMaxP + a], ja2 = J[2 * kMaxP + a];
          if (ja0 == 0.0 && ja1 == 0.0 && ja2 == 0.0) continue;
          const int sa = slot[a], la = cols[a] % n;
          for (int b = 0; b <= a; ++b) {
            const double v = ja0 * J[b] + ja1 * J[kMaxP + b] + ja2 * J[2 * kMaxP + b];
            if (v == 0.0) continue;
            const int sb = slot[b], lb = cols[b] % n;
            if (sa == sb) {
              double* B = tbl[sa][sa];
              B[size_t(la) * n + lb] += v; if (la != lb) B[size_t(lb) * n + la] += v;
            } else if (tblRowIsA[sa][sb]) tbl[sa][sb][size_t(la) * n + lb] += v;
            else tbl[sa][sb][size_t(lb) * n + la] += v;
          }
        }
      }
    }
    if (wantH) for (auto& k : keys) omp_unset_lock(&locks[lockId[k]]);
    if (wantG) for (int q = 0; q < ngf; ++q) {
      if (q == 2 && (f0 == 0 || f1 == 0)) { /* slot 2 unused when frame 0 is already f0/f1 */ }
      omp_set_lock(&glocks[gf[q]]);
      for (int i = 0; i < n; ++i) P.g[size_t(gf[q]) * n + i] += gl[size_t(q) * n + i];
      omp_unset_lock(&glocks[gf[q]]);
    }
  }
  for (auto& l : locks) omp_destroy_lock(&l);
  for (auto& l : glocks) omp_destroy_lock(&l);
  if (costStatic) *costStatic = cost;
  // regularisers (serial over frames; cheap)
  auto sink = [&](Row& row) {
    cost += 0.5 * row.r * row.r;
    if (!(wantG || wantH)) return;
    for (int i = 0; i < row.n; ++i) if (constCol[row.col[i] % n]) row.d[i] = 0.0;
    if (wantG) for (int i = 0; i < row.n; ++i) P.g[row.col[i]] += row.d[i] * row.r;
    if (wantH) for (int i = 0; i < row.n; ++i) { if (row.d[i] == 0.0) continue; for (int j = 0; j <= i; ++j) if (row.d[j] != 0.0) addH(P, row.col[i], row.col[j], row.d[i] * row.d[j]); }
  };
  for (int f = 0; f < P.N; ++f) regulariserRows(P, f, P.jacMode, sink);
  positionRows(P, sink);
  {
    double J[3 * kMaxPT]; int cols[kMaxPT]; double r[3];
    for (size_t t = 0; t < P.tripCenters.size(); ++t) for (int64_t ci = P.tripOffsets[t]; ci < P.tripOffsets[t + 1]; ++ci) {
      const float* rec = &P.tripRecords[size_t(ci) * 10];
      const double w = double(rec[9]), sw = std::sqrt(w);
      const int Pn = evalTriplet(P, P.tripCenters[t], rec, wantG || wantH, r, J, cols);
      cost += 0.5 * w * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      if (!(wantG || wantH)) continue;
      for (int i = 0; i < 3; ++i) { r[i] *= sw; for (int j = 0; j < Pn; ++j) { double& v = J[i * kMaxPT + j]; v = constCol[cols[j] % n] ? 0.0 : v * sw; } }
      if (wantG) for (int j = 0; j < Pn; ++j) P.g[cols[j]] += J[j] * r[0] + J[kMaxPT + j] * r[1] + J[2 * kMaxPT + j] * r[2];
      if (wantH) for (int a = 0; a < Pn; ++a) for (int b = 0; b <= a; ++b) {
        const double v = J[a] * J[b] + J[kMaxPT + a] * J[kMaxPT + b] + J[2 * kMaxPT + a] * J[2 * kMaxPT + b];
        if (v != 0.0) addH(P, cols[a], cols[b], v);
      }
    }
  }
  return cost;
}

// Which parameters are part of the (reduced) Ceres program: referenced by at least
// one residual block and not constant.
static void activeMask(const Problem& P, std::vector<uint8_t>& act) {
  const rcvd_config& c = P.cfg; const Layout& L = P.L; const int n = L.nf;
  act.assign(P.U(), 0);
  const int np = int(P.pairFrames.size() / 2);
  int cols[kMaxP];
  for (int p = 0; p < np; ++p) {
    const int f0 = P.pairFrames[2 * p], f1 = P.pairFrames[2 * p + 1];
    for (int64_t ci = P.offsets[p]; ci < P.offsets[p + 1]; ++ci) {
      ObsData o0, o1; makeObs(c, &P.records[size_t(ci) * 6], o0); makeObs(c, &P.records[size_t(ci) * 6 + 3], o1);
      const int Pn = staticCols(P, f0, f1, o0, o1, cols);
      for (int j = 0; j < Pn; ++j) act[cols[j]] = 1;
    }
  }
  Problem& Pm = const_cast<Problem&>(P);
  auto sink = [&](Row& row) { for (int i = 0; i < row.n; ++i) act[row.col[i]] = 1; };
  for (int f = 0; f < P.N; ++f) regulariserRows(Pm, f, 0, sink);
  positionRows(Pm, sink);
  for (size_t t = 0; t < P.tripCenters.size(); ++t) for (int64_t ci = P.tripOffsets[t]; ci < P.tripOffsets[t + 1]; ++ci) {
    ObsData o[3]; for (int i = 0; i < 3; ++i) makeObs(c, &P.tripRecords[size_t(ci) * 10 + 3 * i], o[i]);
    int tc[kMaxPT]; const int Pn = tripletCols(P, P.tripCenters[t], o, tc);
    for (int j = 0; j < Pn; ++j) act[tc[j]] = 1;
  }
  for (int i = 0; i < P.U(); ++i) if (isConstLocal(P, i % n)) act[i] = 0;
}

// ---------------------------------------------------------------------------
// Block-sparse Cholesky (restating what SPARSE_NORMAL_CHOLESKY computes: an
// exact factorisation of J_s^T J_s + D^2; lib/PoseOptimizer.cpp:956).
// ---------------------------------------------------------------------------
struct BlockChol {
  int N = 0, n = 0;
  std::vector<int> order, pos;                 // elimination order over frames
  std::vector<std::vector<int>> cstruct;       // for each frame k: later-eliminated neighbours (after fill), sorted by pos
  std::map<std::pair<int, int>, std::vector<double>> L;  // key (r, c): pos[r] >= pos[c]; block rows=r, cols=c
  void symbolic(int N_, int n_, const std::map<std::pair<int, int>, std::vector<double>>& H) {
    N = N_; n = n_;
    std::vector<std::set<int>> adj(N);
    for (auto& kv : H) if (kv.first.first != kv.first.second) { adj[kv.first.first].insert(kv.first.second); adj[kv.first.second].insert(kv.first.first); }
    order.clear(); pos.assign(N, -1); cstruct.assign(N, {});
    std::vector<uint8_t> done(N, 0);
    std::vector<std::vector<int>> raw(N);
    for (int it = 0; it < N; ++it) {
      int best = -1; size_t bd = SIZE_MAX;
      for (int f = 0; f < N; ++f) if (!done[f] && adj[f].size() < bd) { bd = adj[f].size(); best = f; }
      done[best] = 1; pos[best] = it; order.push_back(best);
      std::vector<int> nb(adj[best].begin(), adj[best].end());
      raw[best] = nb;
      for (int a : nb) adj[a].erase(best);
      for (size_t i = 0; i < nb.size(); ++i) for (size_t j = i + 1; j < nb.size(); ++j) { adj[nb[i]].insert(nb[j]); adj[nb[j]].insert(nb[i]); }
    }
    for (int f = 0; f < N; ++f) { cstruct[f] = raw[f]; std::sort(cstruct[f].begin(), cstruct[f].end(), [&](int a, int b) { return pos[a] < pos[b]; }); }
    L.clear();
    for (int f = 0; f < N; ++f) { L[{f, f}].assign(size_t(n) * n, 0.0); for (int r : cstruct[f]) L[{r, f}].assign(size_t(n) * n, 0.0); }
  }
  // loads A = S H S + diag(D2) into L storage
  void load(const std::map<std::pair<int, int>, std::vector<double>>& H, const double* S, const double* D2) {
    for (auto& kv : L) std::fill(kv.second.begin(), kv.second.end(), 0.0);
    for (auto& kv : H) {
      const int hr = kv.first.first, hc = kv.first.second;  // hr >= hc by id
      const double* src = kv.second.data();
      if (hr == hc) {
        double* dst = L[{hr, hr}].data();
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) dst[size_t(i) * n + j] = src[size_t(i) * n + j] * S[hr * n + i] * S[hr * n + j];
        for (int i = 0; i < n; ++i) dst[size_t(i) * n + i] += D2[hr * n + i];
      } else if (pos[hr] > pos[hc]) {
        double* dst = L[{hr, hc}].data();
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) dst[size_t(i) * n + j] = src[size_t(i) * n + j] * S[hr * n + i] * S[hc * n + j];
      } else {
        double* dst = L[{hc, hr}].data();
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) dst[size_t(j) * n + i] = src[size_t(i) * n + j] * S[hr * n + i] * S[hc * n + j];
      }
    }
  }
  static bool potrf(double* A, int n) {
    for (int j = 0; j < n; ++j) {
      double* Aj = A + size_t(j) * n;
      double d = Aj[j];
      for (int p = 0; p < j; ++p) d -= Aj[p] * Aj[p];
      if (!(d > 0.0) || !std::isfinite(d)) return false;
      d = std::sqrt(d); Aj[j] = d;
      const double di = 1.0 / d;
      for (int i = j + 1; i < n; ++i) {
        double* Ai = A + size_t(i) * n;
        double s = Ai[j];
#pragma omp simd reduction(- : s)
        for (int p = 0; p < j; ++p) s -= Ai[p] * Aj[p];
        Ai[j] = s * di;
      }
      for (int p = j + 1; p < n; ++p) Aj[p] = 0.0;
    }
    return true;
  }
  // X <- X * Lkk^{-T} for rows [r0, r1): each row x solves x Lkk^T = a  (serial; parallelism is over strips)
  static void trsmRows(double* X, const double* Lkk, int n, int r0, int r1) {
    for (int i = r0; i < r1; ++i) {
      double* x = X + size_t(i) * n;
      for (int j = 0; j < n; ++j) {
        const double* Lj = Lkk + size_t(j) * n;
        double s = x[j];
#pragma omp simd reduction(- : s)
        for (int p = 0; p < j; ++p) s -= x[p] * Lj[p];
        x[j] = s / Lj[j];
      }
    }
  }
  // C[r0:r1, :] -= A[r0:r1, :] * B^T  (n x n row-major), 4x4 register tiles of dot products, serial
  static void gemmNTRows(double* C, const double* A, const double* B, int n, int r0, int r1) {
    for (int i0 = r0; i0 < r1; i0 += 4) for (int j0 = 0; j0 < n; j0 += 4) {
      const int im = std::min(4, r1 - i0), jm = std::min(4, n - j0);
      double acc[4][4] = {{0}};
      if (im == 4 && jm == 4) {
        const double *a0 = A + size_t(i0) * n, *a1 = a0 + n, *a2 = a1 + n, *a3 = a2 + n;
        const double *b0 = B + size_t(j0) * n, *b1 = b0 + n, *b2 = b1 + n, *b3 = b2 + n;
        double c00 = 0, c01 = 0, c02 = 0, c03 = 0, c10 = 0, c11 = 0, c12 = 0, c13 = 0, c20 = 0, c21 = 0, c22 = 0, c23 = 0, c30 = 0, c31 = 0, c32 = 0, c33 = 0;
#pragma omp simd reduction(+ : c00, c01, c02, c03, c10, c11, c12, c13, c20, c21, c22, c23, c30, c31, c32, c33)
        for (int p = 0; p < n; ++p) {
          c00 += a0[p] * b0[p]; c01 += a0[p] * b1[p]; c02 += a0[p] * b2[p]; c03 += a0[p] * b3[p];
          c10 += a1[p] * b0[p]; c11 += a1[p] * b1[p]; c12 += a1[p] * b2[p]; c13 += a1[p] * b3[p];
          c20 += a2[p] * b0[p]; c21 += a2[p] * b1[p]; c22 += a2[p] * b2[p]; c23 += a2[p] * b3[p];
          c30 += a3[p] * b0[p]; c31 += a3[p] * b1[p]; c32 += a3[p] * b2[p]; c33 += a3[p] * b3[p];
        }
        acc[0][0] = c00; acc[0][1] = c01; acc[0][2] = c02; acc[0][3] = c03; acc[1][0] = c10; acc[1][1] = c11; acc[1][2] = c12; acc[1][3] = c13;
        acc[2][0] = c20; acc[2][1] = c21; acc[2][2] = c22; acc[2][3] = c23; acc[3][0] = c30; acc[3][1] = c31; acc[3][2] = c32; acc[3][3] = c33;
      } else {
        for (int i = 0; i < im; ++i) for (int j = 0; j < jm; ++j) { double sacc = 0; for (int p = 0; p < n; ++p) sacc += A[size_t(i0 + i) * n + p] * B[size_t(j0 + j) * n + p]; acc[i][j] = sacc; }
      }
      for (int i = 0; i < im; ++i) for (int j = 0; j < jm; ++j) C[size_t(i0 + i) * n + j0 + j] -= acc[i][j];
    }
  }
  // Level-scheduled right-looking factorisation; OpenMP over (block task x 16-row strip) work items.
  bool factor() {
    std::vector<int> lvl(N, 0); int nl = 0;
    for (int k : order) { for (int a : cstruct[k]) lvl[a] = std::max(lvl[a], lvl[k] + 1); nl = std::max(nl, lvl[k] + 1); }
    std::vector<std::vector<int>> lf(nl);
    for (int k : order) lf[lvl[k]].push_back(k);
    const int strip = 16, ns = (n + strip - 1) / strip;
    bool ok = true;
    for (int l = 0; l < nl && ok; ++l) {
      const std::vector<int>& fr = lf[l];
      std::vector<char> good(fr.size(), 1);
#pragma omp parallel for schedule(dynamic, 1)
      for (size_t q = 0; q < fr.size(); ++q) good[q] = potrf(L[{fr[q], fr[q]}].data(), n) ? 1 : 0;
      for (char gch : good) if (!gch) ok = false;
      if (!ok) break;
      struct TItem { double* X; const double* Lkk; };
      std::vector<TItem> titems;
      for (int k : fr) for (int r : cstruct[k]) titems.push_back({L[{r, k}].data(), L[{k, k}].data()});
      const long nt = (long)titems.size() * ns;
#pragma omp parallel for schedule(dynamic, 1)
      for (long w = 0; w < nt; ++w) { const TItem& t = titems[w / ns]; const int r0 = int(w % ns) * strip; trsmRows(t.X, t.Lkk, n, r0, std::min(n, r0 + strip)); }
      // updates grouped by target block so that no two work items write the same rows
      std::map<std::pair<int, int>, std::vector<std::pair<const double*, const double*>>> upd;
      for (int k : fr) { const auto& st = cstruct[k];
        for (size_t a = 0; a < st.size(); ++a) for (size_t b = 0; b <= a; ++b) upd[{st[a], st[b]}].push_back({L[{st[a], k}].data(), L[{st[b], k}].data()}); }
      struct UItem { double* C; const std::vector<std::pair<const double*, const double*>>* src; };
      std::vector<UItem> uitems;
      for (auto& kv : upd) uitems.push_back({L[kv.first].data(), &kv.second});
      const long nu = (long)uitems.size() * ns;
#pragma omp parallel for schedule(dynamic, 1)
      for (long w = 0; w < nu; ++w) {
        const UItem& u = uitems[w / ns]; const int r0 = int(w % ns) * strip, r1 = std::min(n, r0 + strip);
        for (auto& pr : *u.src) gemmNTRows(u.C, pr.first, pr.second, n, r0, r1);
      }
    }
    return ok;
  }
  void solve(double* b) const {  // in place, b indexed by frame*n
    std::vector<double> tmp(n);
    for (int k : order) {
      const double* Lkk = L.find({k, k})->second.data();
      double* bk = b + size_t(k) * n;
      for (int i = 0; i < n; ++i) { double s = bk[i]; for (int p = 0; p < i; ++p) s -= Lkk[size_t(i) * n + p] * bk[p]; bk[i] = s / Lkk[size_t(i) * n + i]; }
      for (int r : cstruct[k]) {
        const double* Lrk = L.find({r, k})->second.data(); double* br = b + size_t(r) * n;
        for (int i = 0; i < n; ++i) { double s = 0; for (int p = 0; p < n; ++p) s += Lrk[size_t(i) * n + p] * bk[p]; br[i] -= s; }
      }
    }
    for (int it = N - 1; it >= 0; --it) {
      const int k = order[it];
      const double* Lkk = L.find({k, k})->second.data();
      double* bk = b + size_t(k) * n;
      for (int r : cstruct[k]) {
        const double* Lrk = L.find({r, k})->second.data(); const double* br = b + size_t(r) * n;
        for (int i = 0; i < n; ++i) { const double xr = br[i]; if (xr == 0.0) continue; for (int p = 0; p < n; ++p) bk[p] -= Lrk[size_t(i) * n + p] * xr; }
      }
      for (int i = n - 1; i >= 0; --i) { double s = bk[i]; for (int p = i + 1; p < n; ++p) s -= Lkk[size_t(p) * n + i] * bk[p]; bk[i] = s / Lkk[size_t(i) * n + i]; }
    }
  }
};

// y = H v for the lower-block symmetric storage.
static void symMatVec(const Problem& P, const double* v, double* y) {
  const int n = P.L.nf;
  std::fill(y, y + P.U(), 0.0);
  for (auto& kv : P.H) {
    const int r = kv.first.first, c = kv.first.second; const double* B = kv.second.data();
    for (int i = 0; i < n; ++i) { double s = 0; for (int j = 0; j < n; ++j) s += B[size_t(i) * n + j] * v[c * n + j]; y[r * n + i] += s; }
    if (r != c) for (int i = 0; i < n; ++i) { const double vi = v[r * n + i]; if (vi == 0.0) continue; for (int j = 0; j < n; ++j) y[c * n + j] += B[size_t(i) * n + j] * vi; }
  }
}

// ---------------------------------------------------------------------------
// Polynomial helpers for the Armijo line search (ceres/polynomial.cc restated).
// ---------------------------------------------------------------------------
struct Sample { double x, value, gradient; bool valueValid, gradValid; };
static double evalPoly(const std::vector<double>& p, double x) { double v = 0; for (double c : p) v = v * x + c; return v; }  // highest degree first
// Solve small dense system by Gaussian elimination with full pivoting.
static bool solveDense(std::vector<std::vector<double>> A, std::vector<double> b, std::vector<double>& x) {
  const int n = int(b.size()); std::vector<int> perm(n); for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) {
    int pr = k, pc = k; double best = 0;
    for (int i = k; i < n; ++i) for (int j = k; j < n; ++j) if (std::fabs(A[i][j]) > best) { best = std::fabs(A[i][j]); pr = i; pc = j; }
    if (best == 0) return false;
    std::swap(A[k], A[pr]); std::swap(b[k], b[pr]);
    for (int i = 0; i < n; ++i) std::swap(A[i][k], A[i][pc]);
    std::swap(perm[k], perm[pc]);
    for (int i = k + 1; i < n; ++i) { const double f = A[i][k] / A[k][k]; for (int j = k; j < n; ++j) A[i][j] -= f * A[k][j]; b[i] -= f * b[k]; }
  }
  std::vector<double> y(n);
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int j = i + 1; j < n; ++j) s -= A[i][j] * y[j]; y[i] = s / A[i][i]; }
  x.assign(n, 0.0); for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
  return true;
}
static std::vector<double> interpolatingPoly(const std::vector<Sample>& s) {
  int nc = 0; for (auto& q : s) { if (q.valueValid) ++nc; if (q.gradValid) ++nc; }
  const int deg = nc - 1;
  std::vector<std::vector<double>> A; std::vector<double> b;
  for (auto& q : s) {
    if (q.valueValid) { std::vector<double> row(nc); for (int j = 0; j <= deg; ++j) row[j] = std::pow(q.x, deg - j); A.push_back(row); b.push_back(q.value); }
    if (q.gradValid) { std::vector<double> row(nc); for (int j = 0; j < deg; ++j) row[j] = (deg - j) * std::pow(q.x, deg - j - 1); row[deg] = 0; A.push_back(row); b.push_back(q.gradient); }
  }
  std::vector<double> p; solveDense(A, b, p); return p;
}
// real parts of all roots (Durand-Kerner), mirroring FindPolynomialRoots' real output.
static std::vector<double> realRoots(std::vector<double> p) {
  while (!p.empty() && p[0] == 0.0) p.erase(p.begin());
  std::vector<double> out; const int d = int(p.size()) - 1; if (d < 1) return out;
  if (d == 1) { out.push_back(-p[1] / p[0]); return out; }
  if (d == 2) { const double a = p[0], b = p[1], c = p[2], D = b * b - 4 * a * c; if (D >= 0) { const double sq = std::sqrt(D); out.push_back((-b + sq) / (2 * a)); out.push_back((-b - sq) / (2 * a)); } else { out.push_back(-b / (2 * a)); out.push_back(-b / (2 * a)); } return out; }
  std::vector<std::pair<double, double>> z(d);
  for (int i = 0; i < d; ++i) { const double ang = 2 * M_PI * i / d + 0.4; z[i] = {0.9 * std::cos(ang), 0.9 * std::sin(ang)}; }
  auto cmul = [](std::pair<double, double> a, std::pair<double, double> b) { return std::make_pair(a.first * b.first - a.second * b.second, a.first * b.second + a.second * b.first); };
  auto cdiv = [](std::pair<double, double> a, std::pair<double, double> b) { const double dd = b.first * b.first + b.second * b.second; return std::make_pair((a.first * b.first + a.second * b.second) / dd, (a.second * b.first - a.first * b.second) / dd); };
  for (int it = 0; it < 500; ++it) {
    double mx = 0;
    for (int i = 0; i < d; ++i) {
      std::pair<double, double> v = {p[0], 0.0};
      for (int j = 1; j <= d; ++j) { v = cmul(v, z[i]); v.first += p[j]; }
      std::pair<double, double> den = {p[0], 0.0};
      for (int j = 0; j < d; ++j) if (j != i) den = cmul(den, {z[i].first - z[j].first, z[i].second - z[j].second});
      auto dl = cdiv(v, den); z[i].first -= dl.first; z[i].second -= dl.second;
      mx = std::max(mx, std::fabs(dl.first) + std::fabs(dl.second));
    }
    if (mx < 1e-14) break;
  }
  for (auto& r : z) out.push_back(r.first);
  return out;
}
static double minimizeInterpolating(const std::vector<Sample>& s, double xmin, double xmax) {
  const std::vector<double> poly = interpolatingPoly(s);
  double ox = (xmin + xmax) / 2.0, ov = evalPoly(poly, ox);
  const double vmin = evalPoly(poly, xmin); if (vmin < ov) { ov = vmin; ox = xmin; }
  const double vmax = evalPoly(poly, xmax); if (vmax < ov) { ov = vmax; ox = xmax; }
  if (poly.size() <= 2) return ox;
  std::vector<double> der; const int deg = int(poly.size()) - 1;
  for (int j = 0; j < deg; ++j) der.push_back((deg - j) * poly[j]);
  for (double r : realRoots(der)) { if (r < xmin || r > xmax) continue; const double v = evalPoly(poly, r); if (v < ov) { ov = v; ox = r; } }
  return ox;
}

// ---------------------------------------------------------------------------
// Trust-region minimizer with Ceres semantics (see header comment and
// SURVEY.md section 8c for the list of restated rules).
// ---------------------------------------------------------------------------
struct SolveTimes { double eval = 0, linear = 0, cost = 0; };

static void plusProject(const Problem& P, const std::vector<double>& x, const double* delta, std::vector<double>& out) {
  const int n = P.L.nf; const int U = P.U();
  out.resize(U);
  for (int i = 0; i < U; ++i) out[i] = x[i] + delta[i];
  if (P.cfg.depth_lower_bound && P.L.nd > 0) {
    // SetParameterLowerBound(block, 0, 0.0) for every depth block of in-range frames (:1108-1115)
    for (int f = 0; f < P.N; ++f) if (P.inRange[f]) for (int b = 0; b < P.L.G; ++b) { double& v = out[size_t(f) * n + P.L.offD + b * P.L.k]; v = std::max(v, 0.0); }
  }
}

static int solve(Problem& P, const rcvd_solve_options& o, rcvd_solve_summary& sum, SolveTimes& tm) {
  using clk = std::chrono::steady_clock;
  auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto tStart = clk::now();
  const int U = P.U(); const int n = P.L.nf;
  memset(&sum, 0, sizeof(sum));
  sum.num_constraints = P.offsets.empty() ? 0 : P.offsets.back();
  std::vector<uint8_t> act; activeMask(P, act);
  const bool constrained = P.cfg.depth_lower_bound && P.L.nd > 0;
  std::vector<double> x = P.x, zero(U, 0.0), cand;
  if (constrained) { plusProject(P, x, zero.data(), cand); x = cand; P.x = x; }
  auto maskedNorm = [&](const std::vector<double>& v) { double s = 0; for (int i = 0; i < U; ++i) if (act[i]) s += v[i] * v[i]; return std::sqrt(s); };
  double xNorm = maskedNorm(x);
  buildStructure(P);
  auto t0 = clk::now();
  double xCost = evaluate(P, true, true);
  tm.eval += ms(t0, clk::now());
  sum.initial_cost = xCost;
  std::vector<double> S(U, 1.0), D2(U), diag(U), gs(U), y(U), Hy(U), step(U), delta(U), gradient = P.g, xmin = x;
  auto Hdiag = [&](int i) { return P.H.find({i / n, i / n})->second[size_t(i % n) * n + i % n]; };
  if (o.jacobi_scaling) for (int i = 0; i < U; ++i) S[i] = 1.0 / (1.0 + std::sqrt(Hdiag(i)));
  auto gradMaxNorm = [&](const std::vector<double>& xx, const std::vector<double>& g) {
    std::vector<double> neg(U), pr; for (int i = 0; i < U; ++i) neg[i] = -g[i];
    plusProject(P, xx, neg.data(), pr);
    double m = 0; for (int i = 0; i < U; ++i) if (act[i]) m = std::max(m, std::fabs(xx[i] - pr[i])); return m;
  };
  double gmax = gradMaxNorm(x, gradient);
  BlockChol chol; chol.symbolic(P.N, n, P.H);
  double radius = o.initial_radius, decrease = 2.0; bool reuseDiag = false;
  double minimumCost = xCost;
  int iter = 0, invalid = 0; bool stepSuccessful = true;
  sum.termination = RCVD_TERM_NO_CONVERGENCE;
  auto finish = [&](int term, const char* msg) { sum.termination = term; snprintf(sum.message, sizeof(sum.message), "%s", msg); };
  if (o.verbose) fprintf(stderr, "[oracle] iter      cost      cost_change  |gradient|   |step|    tr_ratio  tr_radius\n");
  if (o.verbose) fprintf(stderr, "[oracle] %4d %.6e %10.2e %10.2e %10.2e %10.2e %10.2e\n", 0, xCost, 0.0, gmax, 0.0, 0.0, radius);
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (stepSuccessful) { ++sum.num_successful_steps; if (xCost < minimumCost || iter == 0) { minimumCost = xCost; xmin = x; } }
    else ++sum.num_unsuccessful_steps;
    if (iter >= o.max_iterations) { finish(RCVD_TERM_NO_CONVERGENCE, "Maximum number of iterations reached."); break; }
    if (stepSuccessful && gmax <= o.gradient_tolerance) { finish(RCVD_TERM_CONVERGENCE, "Gradient tolerance reached."); break; }
    if (radius <= o.min_radius) { finish(RCVD_TERM_CONVERGENCE, "Minimum trust region radius reached."); break; }
    ++iter; stepSuccessful = false;
    // ComputeTrustRegionStep (LevenbergMarquardtStrategy::ComputeStep)
    auto tl = clk::now();
    if (!reuseDiag) for (int i = 0; i < U; ++i) diag[i] = std::min(std::max(S[i] * S[i] * Hdiag(i), o.min_lm_diagonal), o.max_lm_diagonal);
    for (int i = 0; i < U; ++i) { D2[i] = diag[i] / radius; gs[i] = S[i] * gradient[i]; }
    chol.load(P.H, S.data(), D2.data());
    bool ok = chol.factor();
    reuseDiag = true;
    double modelChange = 0.0; bool valid = false;
    if (ok) {
      y = gs; chol.solve(y.data());
      for (int i = 0; i < U; ++i) if (!std::isfinite(y[i])) ok = false;
    }
    if (ok) {
      // model_cost_change = -(Js step)^T (f + Js step / 2), step = -y  ->  g_s.y - y^T H_s y / 2
      std::vector<double> Sy(U); for (int i = 0; i < U; ++i) Sy[i] = S[i] * y[i];
      symMatVec(P, Sy.data(), Hy.data());
      double gy = 0, yHy = 0; for (int i = 0; i < U; ++i) { gy += gs[i] * y[i]; yHy += Sy[i] * Hy[i]; }
      modelChange = gy - 0.5 * yHy;
      valid = modelChange > 0.0;
      if (valid) { for (int i = 0; i < U; ++i) { step[i] = -y[i]; delta[i] = step[i] * S[i]; } invalid = 0; }
    }
    tm.linear += ms(tl, clk::now());
    if (!valid) {
      if (++invalid >= o.max_consecutive_invalid_steps) { finish(RCVD_TERM_FAILURE, "Number of consecutive invalid steps more than max_num_consecutive_invalid_steps."); break; }
      radius = radius / decrease; decrease *= 2.0; reuseDiag = true;  // StepIsInvalid -> StepRejected(0)
      if (o.verbose) fprintf(stderr, "[oracle] %4d invalid step, radius %.3e\n", iter, radius);
      continue;
    }
    auto tc = clk::now();
    // DoLineSearch for bounds-constrained problems (Armijo, cubic interpolation, <= 20 iterations)
    if (constrained) {
      double gd = 0; for (int i = 0; i < U; ++i) gd += gradient[i] * delta[i];
      double dirMax = 0; for (int i = 0; i < U; ++i) dirMax = std::max(dirMax, std::fabs(delta[i]));
      std::vector<double> saveX = P.x;
      auto lsEval = [&](double a, Sample& s) {
        std::vector<double> d(U), xx; for (int i = 0; i < U; ++i) d[i] = a * delta[i];
        plusProject(P, x, d.data(), xx); P.x = xx;
        s.x = a; s.value = evaluate(P, true, false); s.valueValid = std::isfinite(s.value);
        double gg = 0; for (int i = 0; i < U; ++i) gg += P.g[i] * delta[i];
        s.gradient = gg; s.gradValid = s.valueValid && std::isfinite(gg);
      };
      Sample init{0.0, xCost, gd, true, true}, prev{0, 0, 0, false, false}, cur;
      lsEval(1.0, cur);
      int lsIter = 0; bool success = true;
      while (!cur.valueValid || cur.value > xCost + 1e-4 * gd * cur.x) {
        if (++lsIter >= 20) { success = false; break; }
        double ss;
        const double lo = 1e-3 * cur.x, hi = 0.6 * cur.x;
        if (!cur.valueValid) ss = std::min(std::max(cur.x * 0.5, lo), hi);
        else { std::vector<Sample> sm{init, cur}; if (prev.valueValid) sm.push_back(prev); ss = minimizeInterpolating(sm, lo, hi); }
        if (ss * dirMax < 1e-9) { success = false; break; }
        prev = cur; lsEval(ss, cur);
      }
      P.x = saveX;
      if (success) for (int i = 0; i < U; ++i) delta[i] *= cur.x;
    }
    // ComputeCandidatePointAndEvaluateCost
    plusProject(P, x, delta.data(), cand);
    P.x = cand;
    double candCost = evaluate(P, false, false);
    if (!std::isfinite(candCost)) candCost = std::numeric_limits<double>::max();
    tm.cost += ms(tc, clk::now());
    // ParameterToleranceReached
    double stepNorm = 0; for (int i = 0; i < U; ++i) if (act[i]) stepNorm += (x[i] - cand[i]) * (x[i] - cand[i]); stepNorm = std::sqrt(stepNorm);
    if (stepNorm <= o.parameter_tolerance * (xNorm + o.parameter_tolerance)) { P.x = x; finish(RCVD_TERM_CONVERGENCE, "Parameter tolerance reached."); break; }
    // FunctionToleranceReached
    const double costChange = xCost - candCost;
    if (std::fabs(costChange) <= o.function_tolerance * xCost) { P.x = x; finish(RCVD_TERM_CONVERGENCE, "Function tolerance reached."); break; }
    // IsStepSuccessful
    const double relDecrease = (candCost >= std::numeric_limits<double>::max()) ? std::numeric_limits<double>::lowest() : costChange / modelChange;
    if (relDecrease > o.min_relative_decrease) {
      x = cand; xNorm = maskedNorm(x);
      auto te = clk::now();
      xCost = evaluate(P, true, true);
      tm.eval += ms(te, clk::now());
      gradient = P.g; gmax = gradMaxNorm(x, gradient);
      stepSuccessful = true;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relDecrease - 1.0, 3));
      radius = std::min(o.max_radius, radius); decrease = 2.0; reuseDiag = false;
    } else {
      P.x = x;
      radius = radius / decrease; decrease *= 2.0; reuseDiag = true;
    }
    if (o.verbose) fprintf(stderr, "[oracle] %4d %.6e %10.2e %10.2e %10.2e %10.2e %10.2e\n", iter, stepSuccessful ? xCost : candCost, costChange, gmax, stepNorm, relDecrease, radius);
  }
  P.x = xmin;   // user-visible parameters: last iterate accepted with minimum cost
  sum.iterations = iter;
  sum.final_cost = minimumCost;
  sum.total_ms = ms(tStart, clk::now());
  sum.eval_ms = tm.eval; sum.linear_ms = tm.linear; sum.cost_ms = tm.cost;
  return RCVD_OK;
}

}  // namespace orc

// ---------------------------------------------------------------------------
// C ABI of the oracle (mirrors include/rcvd.h so tests drive both identically).
// ---------------------------------------------------------------------------
using orc::Problem;
ORC_API const char* orc_last_error() { return orc::g_err.c_str(); }
ORC_API int32_t orc_frame_stride(const rcvd_config* c) { auto L = orc::makeLayout(*c); return L.ok ? L.nf : -1; }
ORC_API int32_t orc_problem_create(const rcvd_config* c, Problem** out) {
  auto L = orc::makeLayout(*c);
  if (!L.ok || c->num_frames <= 0) { orc::g_err = "unsupported configuration"; return RCVD_ERR_INVALID; }
  Problem* P = new Problem(); P->cfg = *c; P->L = L; P->N = c->num_frames;
  P->inRange.assign(P->N, 1); P->median.assign(P->N, 1.0); P->x.assign(size_t(P->N) * L.nf, 0.0);
  P->offsets.assign(1, 0); P->tripOffsets.assign(1, 0);
  orc::buildScaleLocs(*P);
  *out = P; return RCVD_OK;
}
ORC_API void orc_problem_destroy(Problem* P) { delete P; }
ORC_API int32_t orc_problem_set_frames(Problem* P, const uint8_t* inr, const double* med, const double* aw) {
  if (inr) P->inRange.assign(inr, inr + P->N);
  if (med) P->median.assign(med, med + P->N);
  if (aw) P->adaptive.assign(aw, aw + size_t(P->N) * P->cfg.depth_grid_x * P->cfg.depth_grid_y); else P->adaptive.clear();
  P->H.clear();
  return RCVD_OK;
}
ORC_API int32_t orc_problem_set_constraints(Problem* P, int32_t np, const int32_t* pf, const int64_t* off, const float* rec) {
  P->pairFrames.assign(pf, pf + 2 * size_t(np)); P->offsets.assign(off, off + np + 1);
  P->records.assign(rec, rec + size_t(off[np]) * 6); P->H.clear();
  return RCVD_OK;
}
ORC_API int32_t orc_problem_set_triplets(Problem* P, int32_t nt, const int32_t* centers, const int64_t* off, const float* rec) {
  P->tripCenters.assign(centers, centers + nt); P->tripOffsets.assign(off, off + nt + 1); P->tripRecords.assign(rec, rec + size_t(off[nt]) * 10); P->H.clear();
  return RCVD_OK;
}
// residuals and dense Jacobian (Jet) of the smoothness blocks, unweighted: r[3*T], J[3*T][U]
ORC_API int32_t orc_triplet_jacobian(Problem* P, double* r, double* J) {
  const int U = P->U(); double Jb[3 * orc::kMaxPT]; int cols[orc::kMaxPT];
  for (size_t t = 0; t < P->tripCenters.size(); ++t) for (int64_t ci = P->tripOffsets[t]; ci < P->tripOffsets[t + 1]; ++ci) {
    const int Pn = orc::evalTriplet(*P, P->tripCenters[t], &P->tripRecords[size_t(ci) * 10], true, r + 3 * ci, Jb, cols);
    if (J) for (int i = 0; i < 3; ++i) { double* row = J + size_t(3 * ci + i) * U; for (int j = 0; j < Pn; ++j) row[cols[j]] += Jb[i * orc::kMaxPT + j]; }
  }
  return RCVD_OK;
}
ORC_API int32_t orc_problem_set_state(Problem* P, const double* x) { P->x.assign(x, x + P->U()); return RCVD_OK; }
ORC_API int32_t orc_problem_get_state(Problem* P, double* x) { memcpy(x, P->x.data(), sizeof(double) * P->U()); return RCVD_OK; }
ORC_API int32_t orc_set_jacobian_mode(Problem* P, int32_t mode) { P->jacMode = mode; return RCVD_OK; }
ORC_API int32_t orc_set_threads(int32_t n) { omp_set_num_threads(n); return RCVD_OK; }
ORC_API int32_t orc_get_threads() { return omp_get_max_threads(); }
ORC_API int32_t orc_evaluate(Problem* P, double* cost, double* grad) {
  if (P->H.empty()) orc::buildStructure(*P);
  *cost = orc::evaluate(*P, grad != nullptr, false);
  if (grad) memcpy(grad, P->g.data(), sizeof(double) * P->U());
  return RCVD_OK;
}
ORC_API int32_t orc_normal_matrix_dense(Problem* P, double* Hd) {
  orc::buildStructure(*P);
  orc::evaluate(*P, true, true);
  const int U = P->U(), n = P->L.nf;
  std::fill(Hd, Hd + size_t(U) * U, 0.0);
  for (auto& kv : P->H) {
    const int r = kv.first.first, c = kv.first.second;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
      const double v = kv.second[size_t(i) * n + j];
      Hd[size_t(r * n + i) * U + c * n + j] = v;
      if (r != c) Hd[size_t(c * n + j) * U + r * n + i] = v;
    }
  }
  return RCVD_OK;
}
// residuals (unrobustified) and dense Jacobian of the static blocks, for Jacobian tests:
// r[3*C], J[3*C][U] row-major. mode: 0 analytic, 1 jet.
ORC_API int32_t orc_static_jacobian(Problem* P, int32_t mode, double* r, double* J) {
  const int U = P->U(); const int np = int(P->pairFrames.size() / 2);
  double Jb[3 * orc::kMaxP]; int cols[orc::kMaxP];
  for (int p = 0; p < np; ++p) for (int64_t ci = P->offsets[p]; ci < P->offsets[p + 1]; ++ci) {
    const int Pn = orc::evalStatic(*P, P->pairFrames[2 * p], P->pairFrames[2 * p + 1], &P->records[size_t(ci) * 6], mode, true, r + 3 * ci, Jb, cols);
    if (J) for (int i = 0; i < 3; ++i) { double* row = J + size_t(3 * ci + i) * U; for (int j = 0; j < Pn; ++j) row[cols[j]] += Jb[i * orc::kMaxP + j]; }
  }
  return RCVD_OK;
}
// regulariser rows: returns count; fills r[] and dense J[count][U] when non-null.
ORC_API int32_t orc_regulariser_jacobian(Problem* P, int32_t mode, double* r, double* J, int32_t cap) {
  const int U = P->U(); int cnt = 0;
  auto sink = [&](orc::Row& row) {
    if (cnt < cap) { if (r) r[cnt] = row.r; if (J) for (int i = 0; i < row.n; ++i) J[size_t(cnt) * U + row.col[i]] += row.d[i]; }
    ++cnt;
  };
  for (int f = 0; f < P->N; ++f) orc::regulariserRows(*P, f, mode, sink);
  orc::positionRows(*P, sink);
  return cnt;
}
ORC_API int32_t orc_active_mask(Problem* P, uint8_t* out) { std::vector<uint8_t> a; orc::activeMask(*P, a); memcpy(out, a.data(), a.size()); return RCVD_OK; }
ORC_API int32_t orc_gather_depth(const rcvd_config* c, float lx, float ly, int32_t* idx, double* w) { orc::Gather g; orc::gatherDepth(*c, lx, ly, g); for (int i = 0; i < g.n; ++i) { idx[i] = g.idx[i]; w[i] = g.w[i]; } return g.n; }
ORC_API int32_t orc_gather_spatial(const rcvd_config* c, float lx, float ly, int32_t* idx, double* w) { orc::Gather g; orc::gatherSpatial(*c, lx, ly, g); for (int i = 0; i < g.n; ++i) { idx[i] = g.idx[i]; w[i] = g.w[i]; } return g.n; }
ORC_API int32_t orc_solve(Problem* P, const rcvd_solve_options* o, rcvd_solve_summary* s) { orc::SolveTimes tm; return orc::solve(*P, *o, *s, tm); }
// One LM iteration's worth of CPU work at the current state with a fixed radius (for the
// cpu_baseline leg): evaluate+accumulate, factor+solve, candidate cost. Times in ms.
ORC_API int32_t orc_time_iteration(Problem* P, double radius, double* msEval, double* msLinear, double* msCost) {
  using clk = std::chrono::steady_clock;
  auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const int U = P->U(), n = P->L.nf;
  if (P->H.empty()) orc::buildStructure(*P);
  auto t0 = clk::now();
  orc::evaluate(*P, true, true);
  auto t1 = clk::now();
  std::vector<double> S(U), D2(U), y(U);
  for (int i = 0; i < U; ++i) { const double h = P->H.find({i / n, i / n})->second[size_t(i % n) * n + i % n]; S[i] = 1.0 / (1.0 + std::sqrt(h)); D2[i] = std::min(std::max(S[i] * S[i] * h, 1e-6), 1e32) / radius; y[i] = S[i] * P->g[i]; }
  orc::BlockChol chol; chol.symbolic(P->N, n, P->H); chol.load(P->H, S.data(), D2.data());
  const bool ok = chol.factor(); if (ok) chol.solve(y.data());
  auto t2 = clk::now();
  std::vector<double> save = P->x;
  if (ok) for (int i = 0; i < U; ++i) P->x[i] -= y[i] * S[i];
  orc::evaluate(*P, false, false);
  P->x = save;
  auto t3 = clk::now();
  *msEval = ms(t0, t1); *msLinear = ms(t1, t2); *msCost = ms(t2, t3);
  return ok ? RCVD_OK : RCVD_ERR_NUMERIC;
}
// Solve (S H S + diag(d2)) y = b with the block Cholesky -- exposed for linear-solver tests.
ORC_API int32_t orc_block_solve(Problem* P, const double* S, const double* D2, const double* b, double* y) {
  const int U = P->U();
  orc::BlockChol chol; chol.symbolic(P->N, P->L.nf, P->H); chol.load(P->H, S, D2);
  if (!chol.factor()) return RCVD_ERR_NUMERIC;
  memcpy(y, b, sizeof(double) * U); chol.solve(y);
  return RCVD_OK;
}
