// rcvd_filter.cuh -- flow-guided temporal depth filter.
//
// Restates DepthVideoProcessor::flowGuidedFilter (reference lib/Processor.cpp:315-590) for a whole consecutive frame
// range in one launch: one thread per output pixel follows the forward / backward optical-flow chains of every pixel of
// its spatial window through up to frameRadius frames (:469-519), optionally the far connections of the frame (:521-546),
// reprojects every visited location with that frame's depth and camera (DepthVideo::project, lib/DepthVideo.cpp:637-681)
// and measures it along the reference camera's forward axis (:449-451); the output is the exp(-3 max/min) weighted mean
// or weighted median of these depths (:551-585).  All arithmetic is float32 in the reference's operation order with
// explicit round-to-nearest intrinsics (no FMA contraction); tan(fov/2) is taken on the host like the reference does.
//
// Layout: depth[F][hd][wd] f32 (transformed depth of the source stream), cams[F][12] f32
// = {position xyz, quaternion xyzw, tan(hFov/2), tan(vFov/2), pad}, fwd/bwd flow [F][h][w][2] f32 + mask [F][h][w] u8
// (fwd[i]: frame i -> i+1, bwd[i]: frame i -> i-1; unused slots may hold anything), far[K]: {src, dst} + flow/mask.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace rcvd {

struct FilterArgs {
  const float* depth; const float* cams;
  const float* fwd_flow; const uint8_t* fwd_mask; const float* bwd_flow; const uint8_t* bwd_mask;
  const int* far_pairs; const int* far_begin; const float* far_flow; const uint8_t* far_mask;   // far_begin[F+1]: CSR by source frame
  float* out; float2* scratch; int max_samples;
  int F, first_out, num_out, last_frame;   // frames indexed 0..F-1 (index 0 = absolute frame `base`); outputs for first_out .. first_out+num_out-1
  int w, h, wd, hd;
  int frame_radius, spatial_radius, median;
  float inv_aspect;
};

struct FilterCam { float px, py, pz, qx, qy, qz, qw, th, tv; };

__device__ __forceinline__ FilterCam load_cam(const float* cams, int f) {
  const float* c = cams + (size_t)f * 12;
  FilterCam k; k.px = c[0]; k.py = c[1]; k.pz = c[2]; k.qx = c[3]; k.qy = c[4]; k.qz = c[5]; k.qw = c[6]; k.th = c[7]; k.tv = c[8];
  return k;
}
// Eigen::Quaternion * Vector3: uv = 2 (q.vec x v); v + w uv + q.vec x uv
__device__ __forceinline__ void quat_rotate(const FilterCam& k, float vx, float vy, float vz, float& ox, float& oy, float& oz) {
  float ux = __fsub_rn(__fmul_rn(k.qy, vz), __fmul_rn(k.qz, vy));
  float uy = __fsub_rn(__fmul_rn(k.qz, vx), __fmul_rn(k.qx, vz));
  float uz = __fsub_rn(__fmul_rn(k.qx, vy), __fmul_rn(k.qy, vx));
  ux = __fadd_rn(ux, ux); uy = __fadd_rn(uy, uy); uz = __fadd_rn(uz, uz);
  const float cx = __fsub_rn(__fmul_rn(k.qy, uz), __fmul_rn(k.qz, uy));
  const float cy = __fsub_rn(__fmul_rn(k.qz, ux), __fmul_rn(k.qx, uz));
  const float cz = __fsub_rn(__fmul_rn(k.qx, uy), __fmul_rn(k.qy, ux));
  ox = __fadd_rn(__fadd_rn(vx, __fmul_rn(k.qw, ux)), cx);
  oy = __fadd_rn(__fadd_rn(vy, __fmul_rn(k.qw, uy)), cy);
  oz = __fadd_rn(__fadd_rn(vz, __fmul_rn(k.qw, uz)), cz);
}

// addSample (lib/Processor.cpp:436-446): depth of location `loc` (flow-resolution pixels) of frame fi along the reference forward axis
__device__ __forceinline__ float sample_depth(const FilterArgs& a, float lx, float ly, int fi, float rpx, float rpy, float rpz, float rfx, float rfy, float rfz) {
  const float nx = __fdiv_rn(lx, (float)a.w);
  const float ny = __fmul_rn(__fdiv_rn(ly, (float)a.h), a.inv_aspect);
  int x = min(a.wd - 1, (int)__fadd_rn(__fmul_rn(nx, (float)a.wd), 0.5f));
  int y = min(a.hd - 1, (int)__fadd_rn(__fmul_rn(__fdiv_rn(ny, a.inv_aspect), (float)a.hd), 0.5f));
  x = max(x, 0); y = max(y, 0);   // the reference reads out of bounds for locations < -0.5 depth pixels; clamped here
  const float d = a.depth[((size_t)fi * a.hd + y) * a.wd + x];
  const FilterCam k = load_cam(a.cams, fi);
  float rx_, ry_, rz_, ux_, uy_, uz_, fx_, fy_, fz_;
  quat_rotate(k, 1.f, 0.f, 0.f, rx_, ry_, rz_);
  quat_rotate(k, 0.f, 1.f, 0.f, ux_, uy_, uz_);
  quat_rotate(k, 0.f, 0.f, -1.f, fx_, fy_, fz_);
  const float rx = __fadd_rn(-1.f, __fmul_rn(2.f, nx));
  const float ry = __fsub_rn(1.f, __fdiv_rn(__fmul_rn(2.f, ny), a.inv_aspect));
  const float sx = __fmul_rn(rx, k.th), sy = __fmul_rn(ry, k.tv);
  const float rayx = __fadd_rn(__fadd_rn(fx_, __fmul_rn(rx_, sx)), __fmul_rn(ux_, sy));
  const float rayy = __fadd_rn(__fadd_rn(fy_, __fmul_rn(ry_, sx)), __fmul_rn(uy_, sy));
  const float rayz = __fadd_rn(__fadd_rn(fz_, __fmul_rn(rz_, sx)), __fmul_rn(uz_, sy));
  const float px = __fadd_rn(k.px, __fmul_rn(rayx, d)), py = __fadd_rn(k.py, __fmul_rn(rayy, d)), pz = __fadd_rn(k.pz, __fmul_rn(rayz, d));
  const float dx = __fsub_rn(px, rpx), dy = __fsub_rn(py, rpy), dz = __fsub_rn(pz, rpz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, rfx), __fmul_rn(dy, rfy)), __fmul_rn(dz, rfz));
}

// one flow step (:475-494): returns false when the chain breaks (masked source pixel or target outside the image)
__device__ __forceinline__ bool flow_step(const FilterArgs& a, const float* flow, const uint8_t* mask, float& lx, float& ly) {
  const int ix = min((int)__fadd_rn(lx, 0.5f), a.w - 1), iy = min((int)__fadd_rn(ly, 0.5f), a.h - 1);
  if (ix < 0 || iy < 0) return false;   // cannot happen for chains that passed the previous bounds test; guards the first read
  const size_t p = (size_t)iy * a.w + ix;
  if (!mask[p]) return false;
  lx = __fadd_rn(lx, flow[2 * p]); ly = __fadd_rn(ly, flow[2 * p + 1]);
  const int jx = (int)__fadd_rn(lx, 0.5f), jy = (int)__fadd_rn(ly, 0.5f);
  return !(jx < 0 || jx >= a.w || jy < 0 || jy >= a.h);
}

__device__ __forceinline__ float sample_weight(float s, float ref) {
  const float value = __fdiv_rn(fmaxf(s, ref), fminf(s, ref));
  return expf(__fmul_rn(-value, 3.f));
}

template <bool MEDIAN>
__global__ void __launch_bounds__(128) k_flow_guided_filter(FilterArgs a) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 4 + (threadIdx.x >> 5);
  if (x >= a.w || y >= a.h) return;
  const int frame = a.first_out + blockIdx.z;
  const size_t plane = (size_t)a.w * a.h;
  const FilterCam rk = load_cam(a.cams, frame);
  float rfx, rfy, rfz; quat_rotate(rk, 0.f, 0.f, -1.f, rfx, rfy, rfz);
  const int f0 = max(0, frame - a.frame_radius), f1 = min(a.last_frame, frame + a.frame_radius);
  const int x0 = max(0, x - a.spatial_radius), x1 = min(a.w - 1, x + a.spatial_radius);
  const int y0 = max(0, y - a.spatial_radius), y1 = min(a.h - 1, y + a.spatial_radius);
  const float ref = sample_depth(a, (float)x, (float)y, frame, rk.px, rk.py, rk.pz, rfx, rfy, rfz);
  float2* mine = MEDIAN ? a.scratch + ((size_t)blockIdx.z * plane + (size_t)y * a.w + x) * a.max_samples : nullptr;
  int n = 0; float dsum = 0.f, wsum = 0.f;
  auto add = [&](float lx, float ly, int fi) {
    const float s = sample_depth(a, lx, ly, fi, rk.px, rk.py, rk.pz, rfx, rfy, rfz);
    const float wgt = sample_weight(s, ref);
    dsum = __fadd_rn(dsum, __fmul_rn(s, wgt)); wsum = __fadd_rn(wsum, wgt);
    if (MEDIAN) mine[n] = make_float2(s, wgt);
    ++n;
  };
  for (int wy = y0; wy <= y1; ++wy)
    for (int wx = x0; wx <= x1; ++wx) {
      add((float)wx, (float)wy, frame);
      float lx = (float)wx, ly = (float)wy;
      for (int fi = frame + 1; fi <= f1; ++fi) {          // forward chain, flow (fi-1 -> fi)
        if (!flow_step(a, a.fwd_flow + (size_t)(fi - 1) * plane * 2, a.fwd_mask + (size_t)(fi - 1) * plane, lx, ly)) break;
        add(lx, ly, fi);
      }
      lx = (float)wx; ly = (float)wy;
      for (int fi = frame - 1; fi >= f0; --fi) {          // backward chain, flow (fi+1 -> fi)
        if (!flow_step(a, a.bwd_flow + (size_t)(fi + 1) * plane * 2, a.bwd_mask + (size_t)(fi + 1) * plane, lx, ly)) break;
        add(lx, ly, fi);
      }
      if (a.far_begin) {
        for (int q = a.far_begin[frame]; q < a.far_begin[frame + 1]; ++q) {   // far connections; a masked pixel ends the list (:531-533 `break`)
          lx = (float)wx; ly = (float)wy;
          if (!flow_step(a, a.far_flow + (size_t)q * plane * 2, a.far_mask + (size_t)q * plane, lx, ly)) break;
          add(lx, ly, a.far_pairs[2 * q + 1]);
        }
      }
    }
  float result;
  if (MEDIAN) {
    // weighted median (:566-579): samples by ascending depth, first one whose cumulative weight reaches half the total.
    const float half = __fdiv_rn(wsum, 2.f);
    for (int i = 1; i < n; ++i) {               // insertion sort; equal depths give the same output whatever their order
      const float2 v = mine[i]; int j = i - 1;
      while (j >= 0 && mine[j].x > v.x) { mine[j + 1] = mine[j]; --j; }
      mine[j + 1] = v;
    }
    float cum = 0.f; result = a.out[(size_t)blockIdx.z * plane + (size_t)y * a.w + x];   // the reference leaves the pixel untouched if no prefix qualifies (NaN weights)
    for (int i = 0; i < n; ++i) { cum = __fadd_rn(cum, mine[i].y); if (cum >= half) { result = mine[i].x; break; } }
  } else {
    result = wsum > 0.f ? __fdiv_rn(dsum, wsum) : 0.f;
  }
  a.out[(size_t)blockIdx.z * plane + (size_t)y * a.w + x] = result;
}

}  // namespace rcvd
