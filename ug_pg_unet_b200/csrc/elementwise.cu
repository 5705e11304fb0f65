// HBM-bound kernels around the convolutions: input packing, BatchNorm
// statistics / apply / backward, ReLU, 2x2 max-pool, bilinear x2 upsample and
// the weight (un)packers.  Activations are NHWC bf16, 16-byte (8-channel)
// vector accesses, fp32 math.
//
// Reference call sites: UG_unet_parts.py:11-15 (BatchNorm2d,
// ReLU), :49 (MaxPool2d(2)), :78 (F.interpolate x2 bilinear align_corners=True).
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

// ----------------------------------------------------------------------------
// First-layer packing: fp32 NCHW image (optionally bilinearly resized,
// align_corners=True) -> bf16 [B,H,W,64] rows holding the 3x3 neighbourhood,
// k = (r*3+s)*Cin + ci, zero beyond 9*Cin and outside the image (pad 1).
// ----------------------------------------------------------------------------
__global__ void im2col_first_kernel(const float* __restrict__ x, int B, int Cin,
                                    int Hin, int Win, int H, int W,
                                    __nv_bfloat16* __restrict__ out) {
  const long long total = static_cast<long long>(B) * H * W * 8;
  const bool resize = (Hin != H) || (Win != W);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx & 7);
    const long long pix = idx >> 3;
    const int xo = static_cast<int>(pix % W);
    const int yo = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    Vec8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float val = 0.f;
      if (k < 9 * Cin) {
        const int tap = k / Cin;
        const int ci = k - tap * Cin;
        const int yy = yo + tap / 3 - 1;
        const int xx = xo + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const float* plane = x + (static_cast<long long>(b) * Cin + ci) * Hin * Win;
          if (!resize) {
            val = __ldg(plane + static_cast<long long>(yy) * Win + xx);
          } else {
            int y0, y1, x0, x1;
            float ly, lx;
            bilinear_src(yy, Hin, H, &y0, &y1, &ly);
            bilinear_src(xx, Win, W, &x0, &x1, &lx);
            const float v00 = __ldg(plane + static_cast<long long>(y0) * Win + x0);
            const float v01 = __ldg(plane + static_cast<long long>(y0) * Win + x1);
            const float v10 = __ldg(plane + static_cast<long long>(y1) * Win + x0);
            const float v11 = __ldg(plane + static_cast<long long>(y1) * Win + x1);
            val = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) +
                  ly * ((1.f - lx) * v10 + lx * v11);
          }
        }
      }
      r.v[j] = val;
    }
    store8(out + pix * 64 + g * 8, r);
  }
}

// ----------------------------------------------------------------------------
// BatchNorm statistics -> per-channel scale/shift (+ running stats update).
// stats = [sum[C], sumsq[C]] of the stored conv outputs; zeroed after use.
// ----------------------------------------------------------------------------
__global__ void bn_finalize_kernel(float* __restrict__ stats, int C, float count,
                                   const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var,
                                   long long* __restrict__ num_batches_tracked,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean,
                                   float* __restrict__ save_rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
  if (c >= C) return;
  const float s = stats[c], s2 = stats[C + c];
  stats[c] = 0.f;
  stats[C + c] = 0.f;
  const float mean = s / count;
  float var = s2 / count - mean * mean;
  var = fmaxf(var, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  save_mean[c] = mean;
  save_rstd[c] = rstd;
  if (running_mean) {
    const float unbiased = (count > 1.f) ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// Eval mode: scale/shift from running statistics.
__global__ void bn_eval_coeffs_kernel(int C, const float* __restrict__ gamma,
                                      const float* __restrict__ beta,
                                      const float* __restrict__ running_mean,
                                      const float* __restrict__ running_var,
                                      float eps, float* __restrict__ scale,
                                      float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - running_mean[c] * sc;
}

// ----------------------------------------------------------------------------
// a = relu(z*scale + shift), optional fused 2x2 max-pool of a.
// One thread = one 2x2 window x 8 channels when pooling, else one pixel x 8 ch.
// ----------------------------------------------------------------------------
template <bool kPool>
__global__ void bn_relu_apply_kernel(const __nv_bfloat16* __restrict__ z,
                                     const float* __restrict__ scale,
                                     const float* __restrict__ shift, int B, int H,
                                     int W, int C, __nv_bfloat16* __restrict__ a,
                                     __nv_bfloat16* __restrict__ pooled) {
  const int cg = C >> 3;
  if (!kPool) {
    const long long total = static_cast<long long>(B) * H * W * cg;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
         idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int c0 = static_cast<int>(idx % cg) * 8;
      Vec8 v = load8(z + idx * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v.v[j] = bn_relu_val(v.v[j], __ldg(scale + c0 + j), __ldg(shift + c0 + j));
      store8(a + idx * 8, v);
    }
  } else {
    const int Ho = H >> 1, Wo = W >> 1;
    const long long total = static_cast<long long>(B) * Ho * Wo * cg;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
         idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int c0 = static_cast<int>(idx % cg) * 8;
      const long long op = idx / cg;
      const int xo = static_cast<int>(op % Wo);
      const int yo = static_cast<int>((op / Wo) % Ho);
      const int b = static_cast<int>(op / (static_cast<long long>(Wo) * Ho));
      float sc[8], sh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + c0 + j); sh[j] = __ldg(shift + c0 + j); }
      Vec8 best;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long pix = (static_cast<long long>(b) * H + (2 * yo + (i >> 1))) * W + 2 * xo + (i & 1);
        Vec8 v = load8(z + pix * C + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v.v[j] = round_bf16(bn_relu_val(v.v[j], sc[j], sh[j]));
          best.v[j] = (i == 0) ? v.v[j] : fmaxf(best.v[j], v.v[j]);
        }
        store8(a + pix * C + c0, v);
      }
      store8(pooled + op * C + c0, best);
    }
  }
}

// Stand-alone 2x2/stride-2 max pool on NHWC bf16 with PyTorch's index
// convention (plane-local h*W+w, first maximum wins, scan order (0,0),(0,1),
// (1,0),(1,1)); indices are optional.  Reference: UG_unet_parts.py:49.
__global__ void maxpool2x2_kernel(const __nv_bfloat16* __restrict__ x, int B, int H,
                                  int W, int C, __nv_bfloat16* __restrict__ out,
                                  int* __restrict__ indices) {
  const int cg = C >> 3;
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * cg;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % cg) * 8;
    const long long op = idx / cg;
    const int xo = static_cast<int>(op % Wo);
    const int yo = static_cast<int>((op / Wo) % Ho);
    const int b = static_cast<int>(op / (static_cast<long long>(Wo) * Ho));
    Vec8 best;
    int arg[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int yy = 2 * yo + (i >> 1), xx = 2 * xo + (i & 1);
      const long long pix = (static_cast<long long>(b) * H + yy) * W + xx;
      const Vec8 v = load8(x + pix * C + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (i == 0 || v.v[j] > best.v[j]) { best.v[j] = v.v[j]; arg[j] = yy * W + xx; }
      }
    }
    store8(out + op * C + c0, best);
    if (indices) {
#pragma unroll
      for (int j = 0; j < 8; ++j) indices[op * C + c0 + j] = arg[j];
    }
  }
}

// Backward of the stand-alone max pool: routes each pooled gradient to the
// first maximum of its window, zero elsewhere.
__global__ void maxpool2x2_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                      const __nv_bfloat16* __restrict__ dout, int B,
                                      int H, int W, int C,
                                      __nv_bfloat16* __restrict__ dx) {
  const int cg = C >> 3;
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * cg;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % cg) * 8;
    const long long op = idx / cg;
    const int xo = static_cast<int>(op % Wo);
    const int yo = static_cast<int>((op / Wo) % Ho);
    const int b = static_cast<int>(op / (static_cast<long long>(Wo) * Ho));
    Vec8 v[4];
    int arg[8];
    float best[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long pix = (static_cast<long long>(b) * H + (2 * yo + (i >> 1))) * W + 2 * xo + (i & 1);
      v[i] = load8(x + pix * C + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (i == 0 || v[i].v[j] > best[j]) { best[j] = v[i].v[j]; arg[j] = i; }
    }
    const Vec8 g = load8(dout + op * C + c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long pix = (static_cast<long long>(b) * H + (2 * yo + (i >> 1))) * W + 2 * xo + (i & 1);
      Vec8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = (arg[j] == i) ? g.v[j] : 0.f;
      store8(dx + pix * C + c0, o);
    }
  }
}

// ----------------------------------------------------------------------------
// Bilinear x2 upsample (align_corners=True) on NHWC bf16 and its adjoint.
// ----------------------------------------------------------------------------
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int B, int h,
                                  int w, int C, __nv_bfloat16* __restrict__ out) {
  const int cg = C >> 3;
  const int H = 2 * h, W = 2 * w;
  const long long total = static_cast<long long>(B) * H * W * cg;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % cg) * 8;
    const long long op = idx / cg;
    const int X = static_cast<int>(op % W);
    const int Y = static_cast<int>((op / W) % H);
    const int b = static_cast<int>(op / (static_cast<long long>(W) * H));
    int y0, y1, x0, x1;
    float ly, lx;
    bilinear_src(Y, h, H, &y0, &y1, &ly);
    bilinear_src(X, w, W, &x0, &x1, &lx);
    const __nv_bfloat16* base = x + static_cast<long long>(b) * h * w * C + c0;
    const Vec8 v00 = load8(base + (static_cast<long long>(y0) * w + x0) * C);
    const Vec8 v01 = load8(base + (static_cast<long long>(y0) * w + x1) * C);
    const Vec8 v10 = load8(base + (static_cast<long long>(y1) * w + x0) * C);
    const Vec8 v11 = load8(base + (static_cast<long long>(y1) * w + x1) * C);
    Vec8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o.v[j] = (1.f - ly) * ((1.f - lx) * v00.v[j] + lx * v01.v[j]) +
               ly * ((1.f - lx) * v10.v[j] + lx * v11.v[j]);
    store8(out + op * C + c0, o);
  }
}

// Adjoint in gather form (deterministic, no atomics): every low-res pixel sums
// the contributions of the (at most 5x5) high-res pixels that read it.
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int B,
                                      int h, int w, int C,
                                      __nv_bfloat16* __restrict__ dx) {
  const int cg = C >> 3;
  const int H = 2 * h, W = 2 * w;
  const long long total = static_cast<long long>(B) * h * w * cg;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % cg) * 8;
    const long long ip = idx / cg;
    const int x = static_cast<int>(ip % w);
    const int y = static_cast<int>((ip / w) % h);
    const int b = static_cast<int>(ip / (static_cast<long long>(w) * h));
    float wy[5], wx[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int Y = 2 * y - 2 + i;
      float wgt = 0.f;
      if (Y >= 0 && Y < H) {
        int i0, i1; float l;
        bilinear_src(Y, h, H, &i0, &i1, &l);
        if (i0 == y) wgt += 1.f - l;
        if (i1 == y) wgt += l;
      }
      wy[i] = wgt;
      const int X = 2 * x - 2 + i;
      wgt = 0.f;
      if (X >= 0 && X < W) {
        int i0, i1; float l;
        bilinear_src(X, w, W, &i0, &i1, &l);
        if (i0 == x) wgt += 1.f - l;
        if (i1 == x) wgt += l;
      }
      wx[i] = wgt;
    }
    Vec8 acc;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
    const __nv_bfloat16* base = dout + static_cast<long long>(b) * H * W * C + c0;
    for (int i = 0; i < 5; ++i) {
      if (wy[i] == 0.f) continue;
      const int Y = 2 * y - 2 + i;
      for (int k = 0; k < 5; ++k) {
        if (wx[k] == 0.f) continue;
        const int X = 2 * x - 2 + k;
        const Vec8 g = load8(base + (static_cast<long long>(Y) * W + X) * C);
        const float wgt = wy[i] * wx[k];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] = fmaf(wgt, g.v[j], acc.v[j]);
      }
    }
    store8(dx + ip * C + c0, acc);
  }
}

// ----------------------------------------------------------------------------
// BatchNorm(train) + ReLU (+ max-pool) backward.
//   g  = (dA + unpool(dP)) * [a > 0]          a = relu(z*scale+shift)
//   pass 1: sums[0][c] = sum g*xhat (= dgamma), sums[1][c] = sum g (= dbeta)
//           (xhat = (z-mean)*rstd; the order matches BatchNorm2d's weight, bias)
//   pass 2: dz = gamma*rstd * (g - sums1/n - xhat*sums0/n); dbias += sum dz
// Thread = 2x2 window x 8 channels (so the pool arg-max can be recomputed).
// ----------------------------------------------------------------------------
struct BnBwdArgs {
  const __nv_bfloat16* z;
  const __nv_bfloat16* dA;   // nullable (full-resolution gradient)
  const __nv_bfloat16* dP;   // nullable (gradient of the 2x2 max-pooled tensor)
  const float* scale;
  const float* shift;
  const float* mean;
  const float* rstd;
  const float* gamma;
  float* sums;               // [2][C]: dgamma, dbeta
  float* dbias;              // [C] or null (pass 2)
  __nv_bfloat16* dz;         // pass 2 output
  int B, H, W, C;
  float inv_count;
};

template <bool kApply>
__global__ void __launch_bounds__(256) bn_relu_bwd_kernel(const BnBwdArgs a) {
  // blockDim.x = 256; thread's channel group is fixed: (threadIdx.x % cg) when
  // cg divides 256 (C in {64,...,512} -> cg in {8,...,64}).
  const int cg = a.C >> 3;
  const int Ho = a.H >> 1, Wo = a.W >> 1;
  const long long windows = static_cast<long long>(a.B) * Ho * Wo;
  const int my_cg = threadIdx.x % cg;
  const int c0 = my_cg * 8;
  const int rows_per_block = blockDim.x / cg;
  const int my_row = threadIdx.x / cg;

  float sc[8], sh[8], mu[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(a.scale + c0 + j);
    sh[j] = __ldg(a.shift + c0 + j);
    mu[j] = __ldg(a.mean + c0 + j);
    rs[j] = __ldg(a.rstd + c0 + j);
  }
  float k0[8], k1[8], k2[8];
  if (kApply) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      k0[j] = __ldg(a.gamma + c0 + j) * rs[j];
      k1[j] = __ldg(a.sums + a.C + c0 + j) * a.inv_count;   // mean of g
      k2[j] = __ldg(a.sums + c0 + j) * a.inv_count;         // mean of g*xhat
    }
  }
  float acc0[8], acc1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

  for (long long win = static_cast<long long>(blockIdx.x) * rows_per_block + my_row;
       win < windows; win += static_cast<long long>(gridDim.x) * rows_per_block) {
    const int xo = static_cast<int>(win % Wo);
    const int yo = static_cast<int>((win / Wo) % Ho);
    const int b = static_cast<int>(win / (static_cast<long long>(Wo) * Ho));
    Vec8 zv[4];
    float best[8];
    int arg[8];
    long long pix[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pix[i] = (static_cast<long long>(b) * a.H + (2 * yo + (i >> 1))) * a.W + 2 * xo + (i & 1);
      zv[i] = load8(a.z + pix[i] * a.C + c0);
    }
    Vec8 gp;
    if (a.dP) {
      gp = load8(a.dP + win * a.C + c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float av = round_bf16(bn_relu_val(zv[i].v[j], sc[j], sh[j]));
          if (i == 0 || av > best[j]) { best[j] = av; arg[j] = i; }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      Vec8 g;
      if (a.dA) {
        g = load8(a.dA + pix[i] * a.C + c0);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g.v[j] = 0.f;
      }
      Vec8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float gj = g.v[j];
        if (a.dP && arg[j] == i) gj += gp.v[j];
        const float pre = fmaf(zv[i].v[j], sc[j], sh[j]);
        gj = (pre > 0.f) ? gj : 0.f;
        const float xhat = (zv[i].v[j] - mu[j]) * rs[j];
        if (!kApply) {
          acc0[j] += gj;
          acc1[j] = fmaf(gj, xhat, acc1[j]);
        } else {
          const float d = k0[j] * (gj - k1[j] - xhat * k2[j]);
          o.v[j] = d;
          acc0[j] += d;
        }
      }
      if (kApply) store8(a.dz + pix[i] * a.C + c0, o);
    }
  }

  // block reduction over the rows sharing a channel group
  __shared__ float red[2][256][9];
  if (!kApply || a.dbias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[0][threadIdx.x][j] = acc0[j];
      red[1][threadIdx.x][j] = acc1[j];
    }
    __syncthreads();
    if (my_row == 0) {
      for (int r = 1; r < rows_per_block; ++r) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc0[j] += red[0][r * cg + my_cg][j];
          acc1[j] += red[1][r * cg + my_cg][j];
        }
      }
      if (!kApply) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          atomicAdd(a.sums + c0 + j, acc1[j]);
          atomicAdd(a.sums + a.C + c0 + j, acc0[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(a.dbias + c0 + j, acc0[j]);
      }
    }
  }
}

// ----------------------------------------------------------------------------
// Weight packing. fp32 OIHW master weights -> bf16 GEMM operands.
//   fwd  pack: Wf[tap][co][ci]      (tap = r*3+s)
//   dgrad pack: Wd[tap][ci][co] = W[co][ci][2-r][2-s]
//   first layer (9*Cin <= 64): Wf[co][k], k = (r*3+s)*Cin + ci, zero padded to 64
// ----------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin,
                                    __nv_bfloat16* __restrict__ wf,
                                    __nv_bfloat16* __restrict__ wd) {
  const long long total = static_cast<long long>(Cout) * Cin * 9;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    // idx enumerates the fwd pack: (tap, co, ci), ci fastest
    const int ci = static_cast<int>(idx % Cin);
    const int co = static_cast<int>((idx / Cin) % Cout);
    const int tap = static_cast<int>(idx / (static_cast<long long>(Cin) * Cout));
    const float v = w[(static_cast<long long>(co) * Cin + ci) * 9 + tap];
    const __nv_bfloat16 bv = __float2bfloat16_rn(v);
    wf[idx] = bv;
    if (wd) wd[(static_cast<long long>(8 - tap) * Cin + ci) * Cout + co] = bv;
  }
}

__global__ void pack_weights_first_kernel(const float* __restrict__ w, int Cout,
                                          int Cin, __nv_bfloat16* __restrict__ wf) {
  const int total = Cout * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += gridDim.x * blockDim.x) {
    const int k = idx & 63;
    const int co = idx >> 6;
    float v = 0.f;
    if (k < 9 * Cin) {
      const int tap = k / Cin, ci = k - tap * Cin;
      v = w[(co * Cin + ci) * 9 + tap];
    }
    wf[idx] = __float2bfloat16_rn(v);
  }
}

// Packed fp32 gradient [tap][Cin][Cout] -> OIHW fp32 (accumulate optional).
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, int Cout, int Cin,
                                    int first_layer, float* __restrict__ grad,
                                    int accumulate) {
  const long long total = static_cast<long long>(Cout) * Cin * 9;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    // idx enumerates OIHW: (co, ci, tap), tap fastest
    const int tap = static_cast<int>(idx % 9);
    const int ci = static_cast<int>((idx / 9) % Cin);
    const int co = static_cast<int>(idx / (9LL * Cin));
    float v;
    if (first_layer)
      v = dwp[static_cast<long long>(tap * Cin + ci) * Cout + co];
    else
      v = dwp[(static_cast<long long>(tap) * Cin + ci) * Cout + co];
    grad[idx] = accumulate ? grad[idx] + v : v;
  }
}

// ----------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------
static inline int grid_for(long long total, int block, int max_blocks_per_sm = 16) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(device_sm_count()) * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

int im2col_first(const float* x, int B, int Cin, int Hin, int Win, int H, int W,
                 void* out, cudaStream_t stream) {
  UGP_REQUIRE(x && out, "im2col_first: null pointer");
  UGP_REQUIRE(Cin >= 1 && 9 * Cin <= 64, "im2col_first: Cin %d unsupported (9*Cin must be <= 64)", Cin);
  const long long total = static_cast<long long>(B) * H * W * 8;
  im2col_first_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      x, B, Cin, Hin, Win, H, W, static_cast<__nv_bfloat16*>(out));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int bn_finalize(float* stats, int C, long long count, const float* gamma,
                const float* beta, float eps, float momentum, float* running_mean,
                float* running_var, long long* nbt, float* scale, float* shift,
                float* save_mean, float* save_rstd, cudaStream_t stream) {
  UGP_REQUIRE(stats && gamma && beta && scale && shift && save_mean && save_rstd,
              "bn_finalize: null pointer");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(
      stats, C, static_cast<float>(count), gamma, beta, eps, momentum, running_mean,
      running_var, nbt, scale, shift, save_mean, save_rstd);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* rm,
                   const float* rv, float eps, float* scale, float* shift,
                   cudaStream_t stream) {
  UGP_REQUIRE(gamma && beta && rm && rv && scale && shift, "bn_eval_coeffs: null pointer");
  bn_eval_coeffs_kernel<<<(C + 127) / 128, 128, 0, stream>>>(C, gamma, beta, rm, rv, eps,
                                                             scale, shift);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int bn_relu_apply(const void* z, const float* scale, const float* shift, int B, int H,
                  int W, int C, void* a, void* pooled, cudaStream_t stream) {
  UGP_REQUIRE(z && scale && shift && a, "bn_relu_apply: null pointer");
  UGP_REQUIRE(C % 8 == 0, "bn_relu_apply: C must be a multiple of 8");
  if (pooled) {
    UGP_REQUIRE(H % 2 == 0 && W % 2 == 0, "bn_relu_apply: pooling needs even H, W");
    const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
    bn_relu_apply_kernel<true><<<grid_for(total, 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(z), scale, shift, B, H, W, C,
        static_cast<__nv_bfloat16*>(a), static_cast<__nv_bfloat16*>(pooled));
  } else {
    const long long total = static_cast<long long>(B) * H * W * (C / 8);
    bn_relu_apply_kernel<false><<<grid_for(total, 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(z), scale, shift, B, H, W, C,
        static_cast<__nv_bfloat16*>(a), nullptr);
  }
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int maxpool2x2(const void* x, int B, int H, int W, int C, void* out, int* indices,
               cudaStream_t stream) {
  UGP_REQUIRE(x && out, "maxpool2x2: null pointer");
  UGP_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2x2: bad shape");
  const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
  maxpool2x2_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), B, H, W, C, static_cast<__nv_bfloat16*>(out),
      indices);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int maxpool2x2_bwd(const void* x, const void* dout, int B, int H, int W, int C, void* dx,
                   cudaStream_t stream) {
  UGP_REQUIRE(x && dout && dx, "maxpool2x2_bwd: null pointer");
  UGP_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool2x2_bwd: bad shape");
  const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
  maxpool2x2_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dout), B, H,
      W, C, static_cast<__nv_bfloat16*>(dx));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int upsample2x(const void* x, int B, int h, int w, int C, void* out, cudaStream_t stream) {
  UGP_REQUIRE(x && out, "upsample2x: null pointer");
  UGP_REQUIRE(C % 8 == 0, "upsample2x: C must be a multiple of 8");
  const long long total = static_cast<long long>(B) * 4 * h * w * (C / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), B, h, w, C, static_cast<__nv_bfloat16*>(out));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int upsample2x_bwd(const void* dout, int B, int h, int w, int C, void* dx,
                   cudaStream_t stream) {
  UGP_REQUIRE(dout && dx, "upsample2x_bwd: null pointer");
  UGP_REQUIRE(C % 8 == 0, "upsample2x_bwd: C must be a multiple of 8");
  const long long total = static_cast<long long>(B) * h * w * (C / 8);
  upsample2x_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(dout), B, h, w, C, static_cast<__nv_bfloat16*>(dx));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int bn_relu_bwd(const void* z, const void* dA, const void* dP, const float* scale,
                const float* shift, const float* mean, const float* rstd,
                const float* gamma, int B, int H, int W, int C, float* sums,
                float* dbias, void* dz, cudaStream_t stream) {
  UGP_REQUIRE(z && (dA || dP) && scale && shift && mean && rstd && gamma && sums && dz,
              "bn_relu_bwd: null pointer");
  UGP_REQUIRE(C % 64 == 0 && C <= 2048 && 256 % (C / 8) == 0,
              "bn_relu_bwd: C %d unsupported", C);
  UGP_REQUIRE(H % 2 == 0 && W % 2 == 0, "bn_relu_bwd: needs even H, W");
  BnBwdArgs a;
  a.z = static_cast<const __nv_bfloat16*>(z);
  a.dA = static_cast<const __nv_bfloat16*>(dA);
  a.dP = static_cast<const __nv_bfloat16*>(dP);
  a.scale = scale; a.shift = shift; a.mean = mean; a.rstd = rstd; a.gamma = gamma;
  a.sums = sums; a.dbias = dbias; a.dz = static_cast<__nv_bfloat16*>(dz);
  a.B = B; a.H = H; a.W = W; a.C = C;
  a.inv_count = 1.f / static_cast<float>(static_cast<long long>(B) * H * W);
  const long long windows = static_cast<long long>(B) * (H / 2) * (W / 2);
  const int rows_per_block = 256 / (C / 8);
  long long blocks = (windows + rows_per_block - 1) / rows_per_block;
  const long long cap = static_cast<long long>(device_sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  UGP_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, stream));
  bn_relu_bwd_kernel<false><<<static_cast<int>(blocks), 256, 0, stream>>>(a);
  UGP_CHECK_CUDA(cudaGetLastError());
  bn_relu_bwd_kernel<true><<<static_cast<int>(blocks), 256, 0, stream>>>(a);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int pack_weights(const float* w, int Cout, int Cin, void* wf, void* wd,
                 cudaStream_t stream) {
  UGP_REQUIRE(w && wf, "pack_weights: null pointer");
  if (Cin % 64 == 0) {
    const long long total = static_cast<long long>(Cout) * Cin * 9;
    pack_weights_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
        w, Cout, Cin, static_cast<__nv_bfloat16*>(wf), static_cast<__nv_bfloat16*>(wd));
  } else {
    UGP_REQUIRE(9 * Cin <= 64, "pack_weights: Cin %d unsupported", Cin);
    pack_weights_first_kernel<<<grid_for(Cout * 64, 256), 256, 0, stream>>>(
        w, Cout, Cin, static_cast<__nv_bfloat16*>(wf));
  }
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int unpack_wgrad(const float* dwp, int Cout, int Cin, float* grad, int accumulate,
                 cudaStream_t stream) {
  UGP_REQUIRE(dwp && grad, "unpack_wgrad: null pointer");
  const long long total = static_cast<long long>(Cout) * Cin * 9;
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      dwp, Cout, Cin, Cin % 64 != 0, grad, accumulate);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
