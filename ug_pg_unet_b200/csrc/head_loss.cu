// Multi-scale logit head, uncertainty maps and the fused uncertainty-weighted
// loss.  All HBM-bound; fp32 math.
//
// Reference call sites: OutConv 1x1 (UG_unet_parts.py:84-91), multi-scale sum
// of bilinearly upsampled logits (UG_unet.py:294-303), sigmoid-margin
// uncertainty (UG_unet.py:36-57), BCE-with-logits(pos_weight) x (1+alpha*A)
// -> mean (UG_unet.py:61-94, uncertainty_guided_trainer.py:64-65), metrics
// (uncertainty_guided_trainer.py:90-123), softmax-entropy weighting for the
// classifier (Herlev/train_herlev.py:244-281).
#include "ugp_common.cuh"

namespace ugp {

constexpr int kMaxClasses = 8;

__device__ __forceinline__ void bilinear_src_f(int dst, int in, int out, int* i0,
                                               int* i1, float* lambda) {
  const float scale = (out > 1) ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  const float src = scale * static_cast<float>(dst);
  int a = static_cast<int>(src);
  if (a > in - 1) a = in - 1;
  *i0 = a;
  *i1 = a + ((a < in - 1) ? 1 : 0);
  *lambda = src - static_cast<float>(a);
}

// Same rule with the scale (in-1)/(out-1) computed once on the host (IEEE float division, identical to the
// device's): the per-pixel kernels below were bound by the float division and the 64-bit index divisions
// of their address arithmetic, not by memory (21 us for 9 MB).
__device__ __forceinline__ void bilinear_src_s(int dst, int in, float scale, int* i0, int* i1, float* lambda) {
  const float src = scale * static_cast<float>(dst);
  int a = static_cast<int>(src);
  if (a > in - 1) a = in - 1;
  *i0 = a;
  *i1 = a + ((a < in - 1) ? 1 : 0);
  *lambda = src - static_cast<float>(a);
}
static inline float host_bilinear_scale(int in, int out) {
  return (out > 1) ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
}

// ----------------------------------------------------------------------------
// 1x1 OutConv forward: x bf16 [M][C] (NHWC) -> l fp32 [B][nc][h*w] (NCHW).
// C/8 lanes cooperate on one pixel (16-byte loads), shuffle-reduced.
// ----------------------------------------------------------------------------
template <int kLanes, int kNC>   // lanes per pixel = C/8 in {8,16,32}; kNC = classes rounded up
__global__ void __launch_bounds__(256)
outconv_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                   const float* __restrict__ bias, int nc, long long M, int hw,
                   float* __restrict__ out) {
  constexpr int C = kLanes * 8;
  const int sub = threadIdx.x % kLanes;
  float wr[kNC][8];
#pragma unroll
  for (int k = 0; k < kNC; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[k][j] = (k < nc) ? __ldg(w + k * C + sub * 8 + j) : 0.f;
  const long long gthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long pix_per_iter = gthreads / kLanes;
  constexpr int kU = 4;   // independent 16-byte loads in flight per thread
  const long long iters = (M + pix_per_iter * kU - 1) / (pix_per_iter * kU);
  long long pix = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) / kLanes;
  for (long long it = 0; it < iters; ++it, pix += pix_per_iter * kU) {
    uint4 u[kU];
#pragma unroll
    for (int q = 0; q < kU; ++q) {
      const long long pp = pix + q * pix_per_iter;
      u[q] = (pp < M) ? *reinterpret_cast<const uint4*>(x + pp * C + sub * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int q = 0; q < kU; ++q) {
      const long long pp = pix + q * pix_per_iter;
      const bool live = pp < M;      // uniform over the kLanes lanes of a pixel; all lanes join the shuffles
      float acc[kNC];
#pragma unroll
      for (int k = 0; k < kNC; ++k) acc[k] = 0.f;
      const float xv[8] = {bf16_lo(u[q].x), bf16_hi(u[q].x), bf16_lo(u[q].y), bf16_hi(u[q].y),
                           bf16_lo(u[q].z), bf16_hi(u[q].z), bf16_lo(u[q].w), bf16_hi(u[q].w)};
#pragma unroll
      for (int k = 0; k < kNC; ++k)
        if (k < nc) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k] = fmaf(xv[j], wr[k][j], acc[k]);
        }
#pragma unroll
      for (int k = 0; k < kNC; ++k)
        if (k < nc) {
#pragma unroll
          for (int o = kLanes / 2; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        }
      if (live && sub == 0) {
        const long long b = pp / hw, r = pp % hw;
        for (int k = 0; k < nc; ++k) out[(b * nc + k) * hw + r] = acc[k] + __ldg(bias + k);
      }
    }
  }
}

// 1x1 OutConv backward: dx[p][c] (+)= sum_k dl[k][p]*w[k][c];
// dw[k][c] += sum_p dl[k][p]*x[p][c]; db[k] += sum_p dl[k][p].
template <int kLanes, int kNC>
__global__ void __launch_bounds__(256)
outconv_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                   const float* __restrict__ dl, int nc, long long M, int hw,
                   int accumulate, __nv_bfloat16* __restrict__ dx, float* __restrict__ dw,
                   float* __restrict__ db) {
  constexpr int C = kLanes * 8;
  const int sub = threadIdx.x % kLanes;
  float wr[kNC][8], dwr[kNC][8], dbr[kNC];
#pragma unroll
  for (int k = 0; k < kNC; ++k) {
    dbr[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wr[k][j] = (k < nc) ? __ldg(w + k * C + sub * 8 + j) : 0.f;
      dwr[k][j] = 0.f;
    }
  }
  const long long gthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long pix_per_iter = gthreads / kLanes;
  constexpr int kU = 2;   // independent pixels in flight per thread
  for (long long pix0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) / kLanes;
       pix0 < M; pix0 += pix_per_iter * kU) {
    uint4 u[kU], dprev[kU];
    float dlv[kU][kNC];
#pragma unroll
    for (int q = 0; q < kU; ++q) {
      const long long pix = pix0 + q * pix_per_iter;
      if (pix < M) {
        u[q] = *reinterpret_cast<const uint4*>(x + pix * C + sub * 8);
        if (accumulate) dprev[q] = *reinterpret_cast<const uint4*>(dx + pix * C + sub * 8);
        const long long b = pix / hw, r = pix % hw;
#pragma unroll
        for (int k = 0; k < kNC; ++k) dlv[q][k] = (k < nc) ? __ldg(dl + (b * nc + k) * hw + r) : 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < kU; ++q) {
      const long long pix = pix0 + q * pix_per_iter;
      if (pix >= M) continue;
      const float xv[8] = {bf16_lo(u[q].x), bf16_hi(u[q].x), bf16_lo(u[q].y), bf16_hi(u[q].y),
                           bf16_lo(u[q].z), bf16_hi(u[q].z), bf16_lo(u[q].w), bf16_hi(u[q].w)};
      float g[8];
      if (accumulate) {
        const uint4 d = dprev[q];
        g[0] = bf16_lo(d.x); g[1] = bf16_hi(d.x); g[2] = bf16_lo(d.y); g[3] = bf16_hi(d.y);
        g[4] = bf16_lo(d.z); g[5] = bf16_hi(d.z); g[6] = bf16_lo(d.w); g[7] = bf16_hi(d.w);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < kNC; ++k)
        if (k < nc) {
          const float d = dlv[q][k];
          if (sub == 0) dbr[k] += d;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            g[j] = fmaf(d, wr[k][j], g[j]);
            dwr[k][j] = fmaf(d, xv[j], dwr[k][j]);
          }
        }
      uint4 o;
      o.x = pack_bf16x2(g[0], g[1]); o.y = pack_bf16x2(g[2], g[3]);
      o.z = pack_bf16x2(g[4], g[5]); o.w = pack_bf16x2(g[6], g[7]);
      *reinterpret_cast<uint4*>(dx + pix * C + sub * 8) = o;
    }
  }
  // block reduction of dw / db through shared memory
  __shared__ float red[kNC][C];
  __shared__ float redb[kNC];
  for (int i = threadIdx.x; i < kNC * C; i += blockDim.x) (&red[0][0])[i] = 0.f;
  if (threadIdx.x < kNC) redb[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kNC; ++k)
    if (k < nc) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&red[k][sub * 8 + j], dwr[k][j]);
      if (sub == 0) atomicAdd(&redb[k], dbr[k]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < nc * C; i += blockDim.x) atomicAdd(dw + i, (&red[0][0])[i]);
  if (threadIdx.x < nc) atomicAdd(db + threadIdx.x, redb[threadIdx.x]);
}

// ----------------------------------------------------------------------------
// logits[b,k,Y,X] = sum_s bilinear_up(l_s)(Y,X)  (align_corners=True; the
// finest map is added as is).  Up to 4 scales; map s has size (R>>sh[s])^2.
// ----------------------------------------------------------------------------
struct HeadMaps {
  const float* l[4];
  float* dl[4];
  int h[4], w[4];   // map sizes (rectangular images: H != W)
  float sy[4], sx[4];   // (h-1)/(RH-1), (w-1)/(RW-1), computed on the host
  int n;
};

// grid = (ceil(RW / 256), RH, planes): no index division per pixel
__global__ void head_combine_fwd_kernel(const HeadMaps m, int planes, int RH, int RW,
                                        float* __restrict__ logits) {
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.z, Y = blockIdx.y;
  if (X >= RW) return;
  // PyTorch evaluates x6_out + x7_out + x8_out + x9_out left to right,
  // coarsest first (UG_unet.py:303).
  float acc = 0.f;
  for (int s = 0; s < m.n; ++s) {
    const int sh = m.h[s], sw = m.w[s];
    const float* src = m.l[s] + static_cast<long long>(plane) * sh * sw;
    float v;
    if (sh == RH && sw == RW) {
      v = __ldg(src + static_cast<long long>(Y) * RW + X);
    } else {
      int y0, y1, x0, x1;
      float ly, lx;
      bilinear_src_s(Y, sh, m.sy[s], &y0, &y1, &ly);
      bilinear_src_s(X, sw, m.sx[s], &x0, &x1, &lx);
      v = (1.f - ly) * ((1.f - lx) * __ldg(src + y0 * sw + x0) + lx * __ldg(src + y0 * sw + x1)) +
          ly * ((1.f - lx) * __ldg(src + y1 * sw + x0) + lx * __ldg(src + y1 * sw + x1));
    }
    acc = (s == 0) ? v : acc + v;
  }
  logits[(static_cast<long long>(plane) * RH + Y) * RW + X] = acc;
}

// Adjoint of one upsampled scale in gather form: dl[y,x] = sum over the
// high-res pixels whose interpolation stencil touches (y,x).
__global__ void head_combine_bwd_kernel(const float* __restrict__ dlogits, int planes,
                                        int RH, int RW, int sh, int sw, float* __restrict__ dl) {
  const long long total = static_cast<long long>(planes) * sh * sw;
  const float invy = (sh > 1) ? static_cast<float>(RH - 1) / static_cast<float>(sh - 1) : 0.f;
  const float invx = (sw > 1) ? static_cast<float>(RW - 1) / static_cast<float>(sw - 1) : 0.f;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % sw);
    const int y = static_cast<int>((idx / sw) % sh);
    const long long plane = idx / (static_cast<long long>(sh) * sw);
    const float* g = dlogits + plane * RH * RW;
    if (sh == RH && sw == RW) { dl[idx] = g[static_cast<long long>(y) * RW + x]; continue; }
    int Ylo = static_cast<int>(floorf((y - 1) * invy)) - 1, Yhi = static_cast<int>(ceilf((y + 1) * invy)) + 1;
    int Xlo = static_cast<int>(floorf((x - 1) * invx)) - 1, Xhi = static_cast<int>(ceilf((x + 1) * invx)) + 1;
    if (Ylo < 0) Ylo = 0;
    if (Xlo < 0) Xlo = 0;
    if (Yhi > RH - 1) Yhi = RH - 1;
    if (Xhi > RW - 1) Xhi = RW - 1;
    float acc = 0.f;
    for (int Y = Ylo; Y <= Yhi; ++Y) {
      int i0, i1; float l;
      bilinear_src_f(Y, sh, RH, &i0, &i1, &l);
      float wy = 0.f;
      if (i0 == y) wy += 1.f - l;
      if (i1 == y) wy += l;
      if (wy == 0.f) continue;
      float row = 0.f;
      for (int X = Xlo; X <= Xhi; ++X) {
        int j0, j1; float lx;
        bilinear_src_f(X, sw, RW, &j0, &j1, &lx);
        float wx = 0.f;
        if (j0 == x) wx += 1.f - lx;
        if (j1 == x) wx += lx;
        if (wx != 0.f) row = fmaf(wx, __ldg(g + static_cast<long long>(Y) * RW + X), row);
      }
      acc = fmaf(wy, row, acc);
    }
    dl[idx] = acc;
  }
}

// ----------------------------------------------------------------------------
// Sigmoid-margin uncertainty: A = 1 - 2*|up(sigmoid(z_prev)) - 0.5|
// (UG_unet.py:44-57).  z_prev fp32 [planes][h][w] -> A fp32 [planes][H][W].
// ----------------------------------------------------------------------------
__global__ void uncertainty_margin_kernel(const float* __restrict__ zprev, int planes,
                                          int h, int w, int H, int W, float sy, float sx,
                                          float* __restrict__ A) {
  // grid = (ceil(W / 256), H, planes)
  const int X = blockIdx.x * blockDim.x + threadIdx.x;
  const int plane = blockIdx.z, Y = blockIdx.y;
  if (X >= W) return;
  const float* src = zprev + static_cast<long long>(plane) * h * w;
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_src_s(Y, h, sy, &y0, &y1, &ly);
  bilinear_src_s(X, w, sx, &x0, &x1, &lx);
  const float p00 = 1.f / (1.f + expf(-__ldg(src + y0 * w + x0)));
  const float p01 = 1.f / (1.f + expf(-__ldg(src + y0 * w + x1)));
  const float p10 = 1.f / (1.f + expf(-__ldg(src + y1 * w + x0)));
  const float p11 = 1.f / (1.f + expf(-__ldg(src + y1 * w + x1)));
  const float p = (1.f - ly) * ((1.f - lx) * p00 + lx * p01) + ly * ((1.f - lx) * p10 + lx * p11);
  A[(static_cast<long long>(plane) * H + Y) * W + X] = 1.f - 2.f * fabsf(p - 0.5f);
}

// Softmax-entropy uncertainty per sample: u = H(softmax(z)) / ln K
// (Herlev/train_herlev.py:253-257).  One warp per sample.
__global__ void softmax_entropy_kernel(const float* __restrict__ logits, int B, int K,
                                       float* __restrict__ u) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* z = logits + static_cast<long long>(warp) * K;
  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) mx = fmaxf(mx, z[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f;
  for (int k = lane; k < K; k += 32) se += expf(z[k] - mx);
  se = warp_sum(se);
  float ent = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float p = expf(z[k] - mx) / se;
    ent -= p * logf(p + 1e-8f);
  }
  ent = warp_sum(ent);
  if (lane == 0) u[warp] = ent / logf(static_cast<float>(K));
}

// ----------------------------------------------------------------------------
// Fused loss: BCE-with-logits(pos_weight) * (1 + alpha*A) -> mean, plus every
// metric the trainer step reads, plus dlogits, in one pass.
//   out[0] = sum l*w   out[1] = sum l   out[2] = sum A   out[3] = sum A^2
//   out[4] = #(pred != target)          out[5 + 3*b + {0,1,2}] = per-image
//   sum(pred*t), sum(pred), sum(t)      (pred = z > 0  <=>  sigmoid(z) > 0.5)
// ----------------------------------------------------------------------------
template <typename T>   // target storage: float (the reference's mask tensors) or uint8 (loader bytes)
__global__ void __launch_bounds__(256)
bce_uncertainty_loss_kernel(const float* __restrict__ z, const T* __restrict__ t,
                            const float* __restrict__ A, float alpha, float pos_weight,
                            int B, long long per_image, float grad_scale,
                            float* __restrict__ dlogits, double* __restrict__ out) {
  // gridDim.y = B; blocks of one image stride over its pixels
  const int b = blockIdx.y;
  const float* zb = z + b * per_image;
  const T* tb = t + b * per_image;
  const float* Ab = A ? A + b * per_image : nullptr;
  float* db = dlogits ? dlogits + b * per_image : nullptr;
  float s_lw = 0.f, s_l = 0.f, s_a = 0.f, s_a2 = 0.f, s_ne = 0.f, s_pt = 0.f, s_p = 0.f, s_t = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       i < per_image; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float zz = zb[i], y = static_cast<float>(tb[i]);
    const float a = Ab ? Ab[i] : 0.f;
    const float wgt = Ab ? 1.f + alpha * a : 1.f;
    // softplus(-z) = log1p(exp(-|z|)) + max(-z, 0)
    const float sp = log1pf(expf(-fabsf(zz))) + fmaxf(-zz, 0.f);
    const float coef = 1.f + (pos_weight - 1.f) * y;
    const float l = (1.f - y) * zz + coef * sp;
    s_lw = fmaf(l, wgt, s_lw);
    s_l += l;
    s_a += a;
    s_a2 = fmaf(a, a, s_a2);
    const float pred = zz > 0.f ? 1.f : 0.f;
    s_ne += (pred != y) ? 1.f : 0.f;
    s_pt = fmaf(pred, y, s_pt);
    s_p += pred;
    s_t += y;
    if (db) {
      const float sig_neg = 1.f / (1.f + expf(zz));   // sigmoid(-z)
      db[i] = grad_scale * wgt * ((1.f - y) - coef * sig_neg);
    }
  }
  float vals[8] = {s_lw, s_l, s_a, s_a2, s_ne, s_pt, s_p, s_t};
  __shared__ float red[8][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float r = warp_sum(vals[k]);
    if (lane == 0) red[k][wid] = r;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float r = 0.f;
    for (int i = 0; i < 8; ++i) r += red[threadIdx.x][i];
    const int k = threadIdx.x;
    if (k < 5) atomicAdd(out + k, static_cast<double>(r));
    else atomicAdd(out + 5 + 3 * b + (k - 5), static_cast<double>(r));
  }
}

// Weighted cross entropy for the classifier: loss = mean(CE_i * w_i),
// w_i = 1 + alpha*u_i (Herlev/train_herlev.py:274-281); dlogits fused.
//   out[0] = sum CE*w, out[1] = sum CE, out[2] = #correct
__global__ void ce_weighted_loss_kernel(const float* __restrict__ logits,
                                        const long long* __restrict__ labels,
                                        const float* __restrict__ u, float alpha, int B,
                                        int K, float grad_scale,
                                        float* __restrict__ dlogits,
                                        double* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* z = logits + static_cast<long long>(warp) * K;
  const int y = static_cast<int>(labels[warp]);
  float mx = -INFINITY;
  int am = 0;
  for (int k = lane; k < K; k += 32)
    if (z[k] > mx) { mx = z[k]; am = k; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
  }
  float se = 0.f;
  for (int k = lane; k < K; k += 32) se += expf(z[k] - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  const float wgt = u ? 1.f + alpha * u[warp] : 1.f;
  const float ce = lse - z[y];
  if (dlogits)
    for (int k = lane; k < K; k += 32)
      dlogits[static_cast<long long>(warp) * K + k] =
          grad_scale * wgt * (expf(z[k] - lse) - (k == y ? 1.f : 0.f));
  if (lane == 0) {
    atomicAdd(out + 0, static_cast<double>(ce * wgt));
    atomicAdd(out + 1, static_cast<double>(ce));
    atomicAdd(out + 2, (am == y) ? 1.0 : 0.0);
  }
}

// ----------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------
static inline int grid_cap(long long blocks, int per_sm) {
  const long long cap = static_cast<long long>(device_sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

int outconv_fwd(const void* x, const float* w, const float* bias, int B, int hw, int C,
                int nc, float* out, cudaStream_t stream) {
  UGP_REQUIRE(x && w && bias && out, "outconv_fwd: null pointer");
  UGP_REQUIRE(nc >= 1 && nc <= kMaxClasses, "outconv_fwd: num_classes %d unsupported (max %d)", nc, kMaxClasses);
  const long long M = static_cast<long long>(B) * hw;
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  const int lanes = C / 8;
  const int grid = grid_cap((M * lanes + 255) / 256, 8);
  UGP_REQUIRE(C == 64 || C == 128 || C == 256, "outconv_fwd: C %d unsupported (64/128/256)", C);
#define UGP_OC_FWD(L, N) outconv_fwd_kernel<L, N><<<grid, 256, 0, stream>>>(xp, w, bias, nc, M, hw, out)
#define UGP_OC_FWD_C(N) \
  do { if (C == 64) UGP_OC_FWD(8, N); else if (C == 128) UGP_OC_FWD(16, N); else UGP_OC_FWD(32, N); } while (0)
  if (nc == 1) UGP_OC_FWD_C(1); else if (nc == 2) UGP_OC_FWD_C(2); else if (nc <= 4) UGP_OC_FWD_C(4); else UGP_OC_FWD_C(8);
#undef UGP_OC_FWD_C
#undef UGP_OC_FWD
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int outconv_bwd(const void* x, const float* w, const float* dl, int B, int hw, int C,
                int nc, int accumulate, void* dx, float* dw, float* db,
                cudaStream_t stream) {
  UGP_REQUIRE(x && w && dl && dx && dw && db, "outconv_bwd: null pointer");
  UGP_REQUIRE(nc >= 1 && nc <= kMaxClasses, "outconv_bwd: num_classes %d unsupported", nc);
  const long long M = static_cast<long long>(B) * hw;
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* dxp = static_cast<__nv_bfloat16*>(dx);
  const int lanes = C / 8;
  const int grid = grid_cap((M * lanes + 255) / 256, 8);
  UGP_REQUIRE(C == 64 || C == 128 || C == 256, "outconv_bwd: C %d unsupported (64/128/256)", C);
#define UGP_OC_BWD(L, N) \
  outconv_bwd_kernel<L, N><<<grid, 256, 0, stream>>>(xp, w, dl, nc, M, hw, accumulate, dxp, dw, db)
#define UGP_OC_BWD_C(N) \
  do { if (C == 64) UGP_OC_BWD(8, N); else if (C == 128) UGP_OC_BWD(16, N); else UGP_OC_BWD(32, N); } while (0)
  if (nc == 1) UGP_OC_BWD_C(1); else if (nc == 2) UGP_OC_BWD_C(2); else if (nc <= 4) UGP_OC_BWD_C(4); else UGP_OC_BWD_C(8);
#undef UGP_OC_BWD_C
#undef UGP_OC_BWD
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int head_combine_fwd_hw(const float* const* maps, const int* heights, const int* widths, int n, int planes,
                        int RH, int RW, float* logits, cudaStream_t stream) {
  UGP_REQUIRE(n >= 1 && n <= 4 && logits && maps && heights && widths, "head_combine_fwd: bad arguments");
  HeadMaps m;
  m.n = n;
  for (int i = 0; i < 4; ++i) {
    m.l[i] = i < n ? maps[i] : nullptr;
    m.dl[i] = nullptr;
    m.h[i] = i < n ? heights[i] : 0;
    m.w[i] = i < n ? widths[i] : 0;
    UGP_REQUIRE(i >= n || (m.h[i] >= 1 && m.h[i] <= RH && m.w[i] >= 1 && m.w[i] <= RW),
                "head_combine_fwd: map %d is %dx%d, logits %dx%d", i, m.h[i], m.w[i], RH, RW);
  }
  const long long total = static_cast<long long>(planes) * RH * RW;
  for (int i = 0; i < n; ++i) { m.sy[i] = host_bilinear_scale(m.h[i], RH); m.sx[i] = host_bilinear_scale(m.w[i], RW); }
  UGP_REQUIRE(planes <= 65535 && RH <= 65535, "head_combine_fwd: too many planes / rows");
  (void)total;
  head_combine_fwd_kernel<<<dim3((RW + 255) / 256, RH, planes), 256, 0, stream>>>(m, planes, RH, RW, logits);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int head_combine_fwd(const float* const* maps, const int* sizes, int n, int planes, int R,
                     float* logits, cudaStream_t stream) {
  return head_combine_fwd_hw(maps, sizes, sizes, n, planes, R, R, logits, stream);
}

// Separable row form for the coarse scales: one block per (plane, coarse row y).  Stage 1 folds the
// <= 2f+4 fine rows that touch y into one fine-resolution row t[X] (coalesced loads, thread = X);
// stage 2 folds t along X for every coarse x.  Fixed summation order (deterministic); the weights come
// from the same bilinear_src_f as the forward kernel, so this is its exact adjoint.
constexpr int kHeadMaxTaps = 64;
__global__ void __launch_bounds__(256)
head_combine_bwd_rows_kernel(const float* __restrict__ dlogits, int RH, int RW, int sh, int sw, float sy, float sx,
                             float invy, float invx, float* __restrict__ dl) {
  extern __shared__ float t[];            // RW floats folded row + RW (j0, lx) table entries
  int* tj0 = reinterpret_cast<int*>(t + RW);
  float* tlx = t + 2 * RW;
  __shared__ float wys[kHeadMaxTaps];
  const int plane = blockIdx.x / sh, y = blockIdx.x - plane * sh;
  int Ylo = static_cast<int>(floorf((y - 1) * invy)) - 1, Yhi = static_cast<int>(ceilf((y + 1) * invy)) + 1;
  if (Ylo < 0) Ylo = 0;
  if (Yhi > RH - 1) Yhi = RH - 1;
  const int ny = Yhi - Ylo + 1;           // host guarantees <= kHeadMaxTaps
  if (threadIdx.x < ny) {
    int i0, i1;
    float l;
    bilinear_src_s(Ylo + threadIdx.x, sh, sy, &i0, &i1, &l);
    wys[threadIdx.x] = (i0 == y ? 1.f - l : 0.f) + (i1 == y ? l : 0.f);
  }
  __syncthreads();
  const float* g = dlogits + (static_cast<long long>(plane) * RH + Ylo) * RW;
  for (int X = threadIdx.x; X < RW; X += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < ny; ++k) acc = fmaf(wys[k], __ldg(g + static_cast<long long>(k) * RW + X), acc);
    t[X] = acc;
    int j0, j1;
    float lx;
    bilinear_src_s(X, sw, sx, &j0, &j1, &lx);      // source columns / weight of fine column X, once per block
    tj0[X] = j0;
    tlx[X] = (j1 == j0) ? 0.f : lx;
  }
  __syncthreads();
  for (int x = threadIdx.x; x < sw; x += blockDim.x) {
    int Xlo = static_cast<int>(floorf((x - 1) * invx)) - 1, Xhi = static_cast<int>(ceilf((x + 1) * invx)) + 1;
    if (Xlo < 0) Xlo = 0;
    if (Xhi > RW - 1) Xhi = RW - 1;
    float acc = 0.f;
    for (int X = Xlo; X <= Xhi; ++X) {
      const int j0 = tj0[X];
      const float lx = tlx[X];
      // fine column X reads coarse columns j0 (weight 1 - lx) and j0 + 1 (weight lx; lx = 0 at the right edge)
      const float wx = (j0 == x ? 1.f - lx : 0.f) + (j0 + 1 == x ? lx : 0.f);
      acc = fmaf(wx, t[X], acc);
    }
    dl[(static_cast<long long>(plane) * sh + y) * sw + x] = acc;
  }
}

int head_combine_bwd_hw(const float* dlogits, int planes, int RH, int RW, int sh, int sw, float* dl,
                        cudaStream_t stream) {
  UGP_REQUIRE(dlogits && dl && sh >= 1 && sh <= RH && sw >= 1 && sw <= RW, "head_combine_bwd: bad arguments");
  const long long total = static_cast<long long>(planes) * sh * sw;
  const int taps = (sh > 1) ? 2 * ((RH - 1) / (sh - 1) + 1) + 4 : RH;
  if ((sh != RH || sw != RW) && taps <= kHeadMaxTaps && RW * 12 <= 48 * 1024 &&
      static_cast<long long>(planes) * sh < (1LL << 31)) {
    const float invy = (sh > 1) ? static_cast<float>(RH - 1) / static_cast<float>(sh - 1) : 0.f;
    const float invx = (sw > 1) ? static_cast<float>(RW - 1) / static_cast<float>(sw - 1) : 0.f;
    head_combine_bwd_rows_kernel<<<planes * sh, 256, 3 * RW * sizeof(float), stream>>>(
        dlogits, RH, RW, sh, sw, host_bilinear_scale(sh, RH), host_bilinear_scale(sw, RW), invy, invx, dl);
  } else {
    head_combine_bwd_kernel<<<grid_cap((total + 255) / 256, 16), 256, 0, stream>>>(dlogits, planes, RH, RW, sh,
                                                                                 sw, dl);
  }
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int head_combine_bwd(const float* dlogits, int planes, int R, int sz, float* dl,
                     cudaStream_t stream) {
  return head_combine_bwd_hw(dlogits, planes, R, R, sz, sz, dl, stream);
}

int uncertainty_margin(const float* zprev, int planes, int h, int w, int H, int W, float* A,
                       cudaStream_t stream) {
  UGP_REQUIRE(zprev && A, "uncertainty_margin: null pointer");
  const long long total = static_cast<long long>(planes) * H * W;
  (void)total;
  UGP_REQUIRE(planes <= 65535 && H <= 65535, "uncertainty_margin: too many planes / rows");
  uncertainty_margin_kernel<<<dim3((W + 255) / 256, H, planes), 256, 0, stream>>>(
      zprev, planes, h, w, H, W, host_bilinear_scale(h, H), host_bilinear_scale(w, W), A);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int softmax_entropy(const float* logits, int B, int K, float* u, cudaStream_t stream) {
  UGP_REQUIRE(logits && u && K >= 2, "softmax_entropy: bad arguments");
  softmax_entropy_kernel<<<(B * 32 + 127) / 128, 128, 0, stream>>>(logits, B, K, u);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

template <typename T>
static int bce_uncertainty_loss_t(const float* z, const T* t, const float* A, float alpha,
                                  float pos_weight, int B, long long per_image, float grad_scale,
                                  float* dlogits, double* out, cudaStream_t stream) {
  UGP_REQUIRE(z && t && out, "bce_uncertainty_loss: null pointer");
  UGP_REQUIRE(B >= 1 && B <= 65535, "bce_uncertainty_loss: bad batch %d", B);
  UGP_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (5 + 3 * B), stream));
  long long bx = (per_image + 256 * 4 - 1) / (256 * 4);
  const long long cap = (static_cast<long long>(device_sm_count()) * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<unsigned>(bx), B);
  bce_uncertainty_loss_kernel<T><<<grid, 256, 0, stream>>>(z, t, A, alpha, pos_weight, B, per_image,
                                                          grad_scale, dlogits, out);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int bce_uncertainty_loss(const float* z, const float* t, const float* A, float alpha,
                         float pos_weight, int B, long long per_image, float grad_scale,
                         float* dlogits, double* out, cudaStream_t stream) {
  return bce_uncertainty_loss_t<float>(z, t, A, alpha, pos_weight, B, per_image, grad_scale, dlogits, out,
                                       stream);
}

int bce_uncertainty_loss_u8(const float* z, const unsigned char* t, const float* A, float alpha,
                            float pos_weight, int B, long long per_image, float grad_scale,
                            float* dlogits, double* out, cudaStream_t stream) {
  return bce_uncertainty_loss_t<unsigned char>(z, t, A, alpha, pos_weight, B, per_image, grad_scale, dlogits,
                                               out, stream);
}

int ce_weighted_loss(const float* logits, const long long* labels, const float* u, float alpha,
                     int B, int K, float grad_scale, float* dlogits, double* out,
                     cudaStream_t stream) {
  UGP_REQUIRE(logits && labels && out && K >= 2, "ce_weighted_loss: bad arguments");
  UGP_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 3, stream));
  ce_weighted_loss_kernel<<<(B * 32 + 127) / 128, 128, 0, stream>>>(logits, labels, u, alpha, B, K,
                                                                    grad_scale, dlogits, out);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
