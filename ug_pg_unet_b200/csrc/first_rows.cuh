// One first-layer GEMM row: the 3x3 neighbourhood of pixel (b, yo, xo) of the (optionally bilinearly
// resized, align_corners=True) input image as 64 bf16 values, k = (r*3+s)*Cin + ci, zero beyond 9*Cin and
// outside the image - the implicit im2col of nn.Conv2d(n_channels, 64, 3, padding=1) at
// UG_unet_parts.py:10 inside InConv with the resize of uncertainty_guided_trainer.py:208 / UG_unet.py:36-41
// folded in.  Shared by the packer (first_layer.cu: rows to HBM) and by the fused first-layer kernels
// (conv_first.cu: rows straight into the swizzled tcgen05 operand tile), so both see identical bits.
//
// Input: fp32 NCHW, or the loaders' uint8 NHWC bytes with ToTensor's /255 and the optional per-channel
// Normalize(mean, std) as one fma per sample (MoNuSegImprove/monuseg_dataset.py:137-187,
// Herlev/herlev_dataset.py:200-228).
#pragma once

#include "ugp_common.cuh"
#include "ugp_vec.cuh"

namespace ugp {

struct ByteNorm {
  float scale[8], bias[8];   // value = byte * scale[c] + bias[c]
};

template <int kCin, bool kResize, bool kU8>
__device__ __forceinline__ void first_layer_row(const void* __restrict__ xin, int Hin, int Win, int H, int W,
                                                int b, int yo, int xo, const ByteNorm& norm, uint4 (&u)[8]) {
  const long long plane = static_cast<long long>(Hin) * Win;
  const float* img = static_cast<const float*>(xin) + static_cast<long long>(b) * kCin * plane;       // fp32 NCHW
  const uint8_t* imgb = static_cast<const uint8_t*>(xin) + static_cast<long long>(b) * kCin * plane;  // uint8 NHWC
  float vals[64];
#pragma unroll
  for (int k = 9 * kCin; k < 64; ++k) vals[k] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yy = yo + r - 1;
    int y0 = 0, y1 = 0;
    float ly = 0.f;
    if (kResize) bilinear_src(yy < 0 ? 0 : (yy >= H ? H - 1 : yy), Hin, H, &y0, &y1, &ly);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int xx = xo + s - 1;
      const bool inside = yy >= 0 && yy < H && xx >= 0 && xx < W;
      int x0 = 0, x1 = 0;
      float lx = 0.f;
      if (kResize) bilinear_src(xx < 0 ? 0 : (xx >= W ? W - 1 : xx), Win, W, &x0, &x1, &lx);
#pragma unroll
      for (int ci = 0; ci < kCin; ++ci) {
        float v = 0.f;
        if (inside && kU8) {
          const float sc = norm.scale[ci], bi = norm.bias[ci];
          if (!kResize) {
            v = fmaf(static_cast<float>(imgb[(static_cast<long long>(yy) * Win + xx) * kCin + ci]), sc, bi);
          } else {
            const float v00 = fmaf(static_cast<float>(imgb[(static_cast<long long>(y0) * Win + x0) * kCin + ci]), sc, bi);
            const float v01 = fmaf(static_cast<float>(imgb[(static_cast<long long>(y0) * Win + x1) * kCin + ci]), sc, bi);
            const float v10 = fmaf(static_cast<float>(imgb[(static_cast<long long>(y1) * Win + x0) * kCin + ci]), sc, bi);
            const float v11 = fmaf(static_cast<float>(imgb[(static_cast<long long>(y1) * Win + x1) * kCin + ci]), sc, bi);
            v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
          }
        } else if (inside) {
          const float* p = img + ci * plane;
          if (!kResize) {
            v = __ldg(p + static_cast<long long>(yy) * Win + xx);
          } else {
            const float v00 = __ldg(p + static_cast<long long>(y0) * Win + x0);
            const float v01 = __ldg(p + static_cast<long long>(y0) * Win + x1);
            const float v10 = __ldg(p + static_cast<long long>(y1) * Win + x0);
            const float v11 = __ldg(p + static_cast<long long>(y1) * Win + x1);
            v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
          }
        }
        vals[(r * 3 + s) * kCin + ci] = v;
      }
    }
  }
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    u[g].x = pack_bf16x2(vals[g * 8 + 0], vals[g * 8 + 1]);
    u[g].y = pack_bf16x2(vals[g * 8 + 2], vals[g * 8 + 3]);
    u[g].z = pack_bf16x2(vals[g * 8 + 4], vals[g * 8 + 5]);
    u[g].w = pack_bf16x2(vals[g * 8 + 6], vals[g * 8 + 7]);
  }
}

// The same row in two phases for the unresized case, so that a producer thread can keep the loads of its
// NEXT row in flight while it packs the current one (conv_first.cu): `first_layer_fetch` issues the 9*Cin
// loads (raw fp32 bits, or the byte zero-extended) and returns the 9-bit in-image mask of the taps;
// `first_layer_pack` applies the byte normalisation, the zero padding and the bf16 rounding.  Same
// arithmetic as first_layer_row<kCin, false, kU8>: identical bits.
template <int kCin, bool kU8>
__device__ __forceinline__ uint32_t first_layer_fetch(const void* __restrict__ xin, int H, int W, int b, int yo, int xo,
                                                      uint32_t (&raw)[9 * kCin]) {
  const long long plane = static_cast<long long>(H) * W;
  const float* img = static_cast<const float*>(xin) + static_cast<long long>(b) * kCin * plane;
  const uint8_t* imgb = static_cast<const uint8_t*>(xin) + static_cast<long long>(b) * kCin * plane;
  uint32_t mask = 0;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yy = yo + r - 1;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int xx = xo + s - 1;
      const bool inside = yy >= 0 && yy < H && xx >= 0 && xx < W;
      if (inside) mask |= 1u << (r * 3 + s);
      // clamped address, unconditional load: no branch between the loads, all of them issue back to back
      const int yc = yy < 0 ? 0 : (yy >= H ? H - 1 : yy);
      const int xc = xx < 0 ? 0 : (xx >= W ? W - 1 : xx);
      const long long o = static_cast<long long>(yc) * W + xc;
#pragma unroll
      for (int ci = 0; ci < kCin; ++ci) {
        if (kU8) raw[(r * 3 + s) * kCin + ci] = imgb[o * kCin + ci];
        else     raw[(r * 3 + s) * kCin + ci] = __float_as_uint(__ldg(img + ci * plane + o));
      }
    }
  }
  return mask;
}

template <int kCin, bool kU8>
__device__ __forceinline__ void first_layer_pack(const uint32_t (&raw)[9 * kCin], uint32_t mask, const ByteNorm& norm,
                                                 uint4 (&u)[8]) {
  float vals[64];
#pragma unroll
  for (int k = 9 * kCin; k < 64; ++k) vals[k] = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const bool inside = (mask >> t) & 1u;
#pragma unroll
    for (int ci = 0; ci < kCin; ++ci) {
      const uint32_t w = raw[t * kCin + ci];
      const float v = kU8 ? fmaf(static_cast<float>(w), norm.scale[ci], norm.bias[ci]) : __uint_as_float(w);
      vals[t * kCin + ci] = inside ? v : 0.f;
    }
  }
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    u[g].x = pack_bf16x2(vals[g * 8 + 0], vals[g * 8 + 1]);
    u[g].y = pack_bf16x2(vals[g * 8 + 2], vals[g * 8 + 3]);
    u[g].z = pack_bf16x2(vals[g * 8 + 4], vals[g * 8 + 5]);
    u[g].w = pack_bf16x2(vals[g * 8 + 6], vals[g * 8 + 7]);
  }
}

// Host: ToTensor (/255) + Normalize(mean, std) of `Cin` channels as scale / bias (nullable HOST arrays).
int make_byte_norm(ByteNorm* out, int Cin, const float* mean, const float* stddev, const char* who);

}  // namespace ugp
