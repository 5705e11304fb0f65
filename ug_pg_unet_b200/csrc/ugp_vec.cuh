// Shared device helpers of the HBM-bound kernels: 8-channel bf16 vectors, the
// PyTorch bilinear (align_corners=True) index rule and the single definition of
// BatchNorm+ReLU on a stored conv output.
#pragma once

#include "ugp_common.cuh"

namespace ugp {

// PyTorch's align_corners=True source index: src = dst * (in-1)/(out-1).
__device__ __forceinline__ void bilinear_src(int dst, int in, int out, int* i0,
                                             int* i1, float* lambda) {
  const float scale = (out > 1) ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  const float src = scale * static_cast<float>(dst);
  int a = static_cast<int>(src);
  if (a > in - 1) a = in - 1;
  *i0 = a;
  *i1 = a + ((a < in - 1) ? 1 : 0);
  *lambda = src - static_cast<float>(a);
}

// Same rule with the scale (in-1)/(out-1) precomputed (on the host, in float: IEEE division, identical
// to the device's): no division and one conversion less per call.  The conversions run on the XU pipe
// (16 lanes / clk / SM), which the first upsample kernels saturated (ncu r02: 120 % of peak).
__device__ __forceinline__ void bilinear_src_scaled(int dst, int in, float scale, int* i0, int* i1, float* lambda) {
  const float src = scale * static_cast<float>(dst);
  int a = static_cast<int>(src);
  if (a > in - 1) a = in - 1;
  *i0 = a;
  *i1 = a + ((a < in - 1) ? 1 : 0);
  *lambda = src - static_cast<float>(a);
}

// The one definition of train/eval BatchNorm + ReLU on a stored bf16 conv
// output; forward, max-pool and backward all call this so the recomputed
// activation (and therefore the arg-max / ReLU mask) is bit-identical.
__device__ __forceinline__ float bn_relu_val(float z, float sc, float sh) {
  return fmaxf(fmaf(z, sc, sh), 0.f);
}
__device__ __forceinline__ float round_bf16(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

struct Vec8 {
  float v[8];
};
__device__ __forceinline__ Vec8 unpack8(const uint4 u) {
  Vec8 r;
  r.v[0] = bf16_lo(u.x); r.v[1] = bf16_hi(u.x);
  r.v[2] = bf16_lo(u.y); r.v[3] = bf16_hi(u.y);
  r.v[4] = bf16_lo(u.z); r.v[5] = bf16_hi(u.z);
  r.v[6] = bf16_lo(u.w); r.v[7] = bf16_hi(u.w);
  return r;
}
__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  Vec8 r;
  r.v[0] = bf16_lo(u.x); r.v[1] = bf16_hi(u.x);
  r.v[2] = bf16_lo(u.y); r.v[3] = bf16_hi(u.y);
  r.v[4] = bf16_lo(u.z); r.v[5] = bf16_hi(u.z);
  r.v[6] = bf16_lo(u.w); r.v[7] = bf16_hi(u.w);
  return r;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const Vec8& r) {
  uint4 u;
  u.x = pack_bf16x2(r.v[0], r.v[1]);
  u.y = pack_bf16x2(r.v[2], r.v[3]);
  u.z = pack_bf16x2(r.v[4], r.v[5]);
  u.w = pack_bf16x2(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

static inline int grid_for(long long total, int block, int max_blocks_per_sm = 16) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(device_sm_count()) * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace ugp
