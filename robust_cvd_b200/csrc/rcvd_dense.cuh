// rcvd_dense.cuh -- dense per-pixel transform application:
// DepthXform::apply (reference lib/DepthMapTransform.cpp:394-415), GridDepthXform::paramMap
// (:950-994) and SpatialXform::warp (:428-449).  Pixel -> NDC uses the dense map
// x = -1 + x*2/(w-1), y = 1 - y*2/(h-1) in float32 (:397-407), unlike the solver samples.
#pragma once
#include "rcvd_device.cuh"

namespace rcvd {

__device__ __forceinline__ void dense_loc(int x, int y, int w, int h, float& lx, float& ly) {
  const float xs = __fdiv_rn(2.f, __fsub_rn((float)w, 1.f));
  const float ys = __fdiv_rn(2.f, __fsub_rn((float)h, 1.f));
  lx = __fadd_rn(-1.f, __fmul_rn((float)x, xs));
  ly = __fsub_rn(1.f, __fmul_rn((float)y, ys));
}

// MODE 0: apply (dst f32), 1: paramMap (out f64 x k), 2: warp (out f32 x 2)
template <int MODE>
__global__ void __launch_bounds__(256) k_dense(rcvd_config cfg, Layout L, const double* __restrict__ params /* frame-style vector */,
                                               const float* __restrict__ src, void* __restrict__ out, int h, int w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  const int x = i % w, y = i / w;
  float lx, ly; dense_loc(x, y, w, h, lx, ly);
  Gather g;
  if (MODE == 2) {
    gather_spatial(cfg, lx, ly, g);
    double u[2]; warp_value(L, g, params, u);
    reinterpret_cast<float2*>(out)[i] = make_float2((float)u[0], (float)u[1]);
  } else {
    gather_depth(cfg, lx, ly, g);
    if (MODE == 0) {
      reinterpret_cast<float*>(out)[i] = (float)depth_value(cfg, L, g, src[i], params);
    } else {
      double* o = reinterpret_cast<double*>(out) + (size_t)i * L.k;
      for (int d = 0; d < L.k; ++d) o[d] = 0.0;
      for (int q = 0; q < g.n; ++q) for (int d = 0; d < L.k; ++d) o[d] += params[L.offD + g.idx[q] * L.k + d] * g.w[q];
    }
  }
}

}  // namespace rcvd
