// HBM-bound kernels around the convolutions: input packing, BatchNorm
// statistics / apply / backward, ReLU, 2x2 max-pool, bilinear x2 upsample and
// the weight (un)packers.  Activations are NHWC bf16, 16-byte (8-channel)
// vector accesses, fp32 math.
//
// Reference call sites: UG_unet_parts.py:11-15 (BatchNorm2d,
// ReLU), :49 (MaxPool2d(2)), :78 (F.interpolate x2 bilinear align_corners=True).
#include "ugp_common.cuh"
#include "ugp_vec.cuh"

namespace ugp {

int g_upsample_smem_kb = 72;   // shared-memory budget of a bilinear x2 block (tuning switch `upsample_smem_kb`)

// ----------------------------------------------------------------------------
// BatchNorm statistics -> per-channel scale/shift (+ running stats update).
// stats = [parts][2][C]: per-CTA partial {sum, sum of squares} of the stored conv outputs, summed here
// in slot order (deterministic) and zeroed after use.
// ----------------------------------------------------------------------------
// Block = 32 channels x 32 slot groups: group g sums slots g, g+32, ... (all loads of a thread in
// flight at once), the group sums are then added in a fixed order -> deterministic, ~5 dependent steps.
constexpr int kFinGroups = 32;
__global__ void __launch_bounds__(32 * kFinGroups)
bn_finalize_kernel(float* __restrict__ stats, int parts, int C, float count,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                   long long* __restrict__ num_batches_tracked, float* __restrict__ scale,
                   float* __restrict__ shift, float* __restrict__ save_mean,
                   float* __restrict__ save_rstd) {
  __shared__ float red[2][kFinGroups][33];
  pdl_wait();      // launched while the convolution that writes `stats` drains (programmatic dependent launch)
  const int cx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked) *num_batches_tracked += 1;
  float s = 0.f, s2 = 0.f;
  // the kernel is one chain of dependent round trips (slots -> sums -> coefficients): everything the last
  // step needs is requested up front, together with the slots
  float g_c = 0.f, b_c = 0.f, rm_c = 0.f, rv_c = 0.f;
  if (g == 0 && c < C) {
    g_c = gamma[c];
    b_c = beta[c];
    if (running_mean) { rm_c = running_mean[c]; rv_c = running_var[c]; }
  }
  if (c < C) {
#pragma unroll 5
    for (int q = g; q < parts; q += kFinGroups) {
      const float* sl = stats + static_cast<size_t>(q) * 2 * C;
      s += __ldcg(sl + c);
      s2 += __ldcg(sl + C + c);
    }
    for (int q = g; q < parts; q += kFinGroups) {
      float* sl = stats + static_cast<size_t>(q) * 2 * C;
      sl[c] = 0.f;
      sl[C + c] = 0.f;
    }
  }
  red[0][g][cx] = s;
  red[1][g][cx] = s2;
  __syncthreads();
  if (g != 0 || c >= C) return;
  s = 0.f; s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kFinGroups; ++i) { s += red[0][i][cx]; s2 += red[1][i][cx]; }
  const float mean = s / count;
  float var = s2 / count - mean * mean;
  var = fmaxf(var, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float sc = g_c * rstd;
  scale[c] = sc;
  shift[c] = b_c - mean * sc;
  save_mean[c] = mean;
  save_rstd[c] = rstd;
  if (running_mean) {
    const float unbiased = (count > 1.f) ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * rm_c + momentum * mean;
    running_var[c] = (1.f - momentum) * rv_c + momentum * unbiased;
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
  pdl_wait();      // scale / shift come from bn_finalize, launched just before (programmatic dependent launch)
  const int cg = C >> 3;
  // the grid stride (gridDim.x * 256) is a multiple of cg (256 % cg == 0, checked by the launcher), so
  // a thread keeps its 8-channel group for the whole loop: scale / shift live in registers
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long first = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int c0 = static_cast<int>(first % cg) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + c0 + j); sh[j] = __ldg(shift + c0 + j); }
  if (!kPool) {
    const long long total = static_cast<long long>(B) * H * W * cg;
    for (long long idx = first; idx < total; idx += 2 * stride) {
      // two independent 16-byte loads in flight per thread
      const bool two = idx + stride < total;
      Vec8 v = load8(z + idx * 8);
      Vec8 v2 = two ? load8(z + (idx + stride) * 8) : v;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v.v[j] = bn_relu_val(v.v[j], sc[j], sh[j]);
        v2.v[j] = bn_relu_val(v2.v[j], sc[j], sh[j]);
      }
      store8(a + idx * 8, v);
      if (two) store8(a + (idx + stride) * 8, v2);
    }
  } else {
    const int Ho = H >> 1, Wo = W >> 1;
    const long long total = static_cast<long long>(B) * Ho * Wo * cg;
    for (long long idx = first; idx < total; idx += stride) {
      const long long op = idx / cg;
      const int xo = static_cast<int>(op % Wo);
      const int yo = static_cast<int>((op / Wo) % Ho);
      const int b = static_cast<int>(op / (static_cast<long long>(Wo) * Ho));
      Vec8 v[4];
      long long pix[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pix[i] = (static_cast<long long>(b) * H + (2 * yo + (i >> 1))) * W + 2 * xo + (i & 1);
        v[i] = load8(z + pix[i] * C + c0);
      }
      Vec8 best;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[i].v[j] = round_bf16(bn_relu_val(v[i].v[j], sc[j], sh[j]));
          best.v[j] = (i == 0) ? v[i].v[j] : fmaxf(best.v[j], v[i].v[j]);
        }
        store8(a + pix[i] * C + c0, v[i]);
      }
      store8(pooled + op * C + c0, best);
    }
  }
}

// a = relu(z*scale + shift) AND the 1x1 head on it (OutConv, UG_unet_parts.py:84-91 applied to a decoder
// output, UG_unet.py:294-297): lmap[b][k][p] = bh[k] + sum_c wh[k][c] * a[p][c] (fp32 NCHW), computed from
// the fp32 activation while it is in registers - the head no longer re-reads the activation from HBM.
// One thread = one pixel x 8 channels; the C/8 <= 32 lanes of a pixel are adjacent, the channel sum is a
// butterfly over them.
template <int kNC>
__global__ void __launch_bounds__(256)
bn_relu_apply_head_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
                          const float* __restrict__ shift, long long M, int C, int HW,
                          const float* __restrict__ wh, const float* __restrict__ bh,
                          __nv_bfloat16* __restrict__ a, float* __restrict__ lmap) {
  pdl_wait();      // scale / shift come from bn_finalize, launched just before (programmatic dependent launch)
  const int cg = C >> 3;                       // power of two, 8..32 (checked by the launcher)
  const int cg_shift = 31 - __clz(cg);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long total = M * cg;
  const int lane = threadIdx.x & 31;
  const int g = lane & (cg - 1);               // blockDim and the grid stride are multiples of 32 >= cg
  const int c0 = g * 8;
  float sc[8], sh[8], w[kNC][8], bias[kNC];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + c0 + j); sh[j] = __ldg(shift + c0 + j); }
#pragma unroll
  for (int k = 0; k < kNC; ++k) {
    bias[k] = __ldg(bh + k);
#pragma unroll
    for (int j = 0; j < 8; ++j) w[k][j] = __ldg(wh + k * C + c0 + j);
  }
  // warp-uniform loop bounds: every lane takes part in the shuffles, out-of-range lanes carry zeros
  for (long long base = blockIdx.x * static_cast<long long>(blockDim.x) + (threadIdx.x - lane); base < total;
       base += 2 * stride) {
    const long long i1 = base + lane, i2 = base + stride + lane;
    const bool v1 = i1 < total, v2 = i2 < total;
    Vec8 x1, x2;
    if (v1) x1 = load8(z + i1 * 8);
    if (v2) x2 = load8(z + i2 * 8);
    float d1[kNC], d2[kNC];
#pragma unroll
    for (int k = 0; k < kNC; ++k) { d1[k] = 0.f; d2[k] = 0.f; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x1.v[j] = v1 ? bn_relu_val(x1.v[j], sc[j], sh[j]) : 0.f;
      x2.v[j] = v2 ? bn_relu_val(x2.v[j], sc[j], sh[j]) : 0.f;
#pragma unroll
      for (int k = 0; k < kNC; ++k) { d1[k] = fmaf(w[k][j], x1.v[j], d1[k]); d2[k] = fmaf(w[k][j], x2.v[j], d2[k]); }
    }
    if (v1) store8(a + i1 * 8, x1);
    if (v2) store8(a + i2 * 8, x2);
#pragma unroll
    for (int k = 0; k < kNC; ++k) {
      for (int off = cg >> 1; off > 0; off >>= 1) {
        d1[k] += __shfl_xor_sync(0xffffffffu, d1[k], off);
        d2[k] += __shfl_xor_sync(0xffffffffu, d2[k], off);
      }
    }
    if (g == 0) {
      // [B][nc][HW]: (b*nc + k)*HW + q = pixel + (b*(nc-1) + k)*HW; one class needs no image index at all
      if (v1) {
        const long long pix = i1 >> cg_shift;
        const int b = (kNC == 1) ? 0 : static_cast<int>(pix / HW);
#pragma unroll
        for (int k = 0; k < kNC; ++k) lmap[pix + static_cast<long long>(b * (kNC - 1) + k) * HW] = d1[k] + bias[k];
      }
      if (v2) {
        const long long pix = i2 >> cg_shift;
        const int b = (kNC == 1) ? 0 : static_cast<int>(pix / HW);
#pragma unroll
        for (int k = 0; k < kNC; ++k) lmap[pix + static_cast<long long>(b * (kNC - 1) + k) * HW] = d2[k] + bias[k];
      }
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
// Bilinear x2 upsample (align_corners=True) on NHWC bf16 and its adjoint
// (F.interpolate / nn.Upsample at UG_unet_parts.py:76-79 and UG_unet.py:296-301).
// ----------------------------------------------------------------------------
// Both kernels were ALU / issue bound in their first form (ncu, r02: 75 % / 68 % issue utilisation at 1.9 /
// 2.3 TB/s of DRAM traffic): per 16-byte output the forward did four 16-byte loads, 32 bf16 unpacks, two
// integer divisions and the float arithmetic of two bilinear_src calls; the adjoint gathered 16-25 taps
// per output.  Now the interpolation is SEPARATED through shared memory (fp32): forward = vertical mix of
// the two source rows once per source pixel, then a two-tap horizontal lerp per output pixel; adjoint =
// vertical gather of the <= 5 contributing rows once per high-res column, then the <= 5-tap horizontal
// gather.  A thread owns one 4-channel group for the whole kernel and walks pixels with a constant stride
// (no division in the loops); a warp's 16-byte shared-memory accesses cover 512 contiguous bytes (no bank
// conflicts); index / weight tables are computed once per block with PyTorch's own arithmetic (bilinear_src).
//
// One block per output row (b, Y) and channel slab; slab width Cs (a multiple of 4 with 256 % (Cs/4) == 0)
// is chosen by the launcher so that the staged row fits in shared memory.
constexpr int kUpThreads = 256;

__device__ __forceinline__ float4 unpack4(const uint2 u) {
  return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
}

__global__ void __launch_bounds__(kUpThreads)
upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int B, int h, int w, int C, int Cs, float sy, float sx,
                  __nv_bfloat16* __restrict__ out) {
  extern __shared__ float up_smem[];
  float* mix = up_smem;                                   // [w][Cs] vertical mix, fp32
  int* tx0 = reinterpret_cast<int*>(mix + w * Cs);        // [2w] left source column * Cs
  float* tlx = reinterpret_cast<float*>(tx0 + 2 * w);     // [2w] weight of the right source column
  const int cg4 = Cs >> 2;
  const int c = (threadIdx.x % cg4) * 4;                  // this thread's channel group, fixed
  const int xr = threadIdx.x / cg4, xstep = kUpThreads / cg4;
  const int H = 2 * h, W = 2 * w;
  const int row = blockIdx.x;
  const int cs0 = blockIdx.y * Cs;
  const int b = row / H, Y = row - b * H;
  int y0, y1;
  float ly;
  bilinear_src_scaled(Y, h, sy, &y0, &y1, &ly);
  for (int X = threadIdx.x; X < W; X += kUpThreads) {
    int x0, x1;
    float lx;
    bilinear_src_scaled(X, w, sx, &x0, &x1, &lx);
    tx0[X] = x0 * Cs;
    tlx[X] = (x1 == x0) ? 0.f : lx;       // right edge: both taps are the last column
  }
  const __nv_bfloat16* r0 = x + (static_cast<long long>(b) * h + y0) * w * C + cs0 + c;
  const __nv_bfloat16* r1 = x + (static_cast<long long>(b) * h + y1) * w * C + cs0 + c;
  for (int xs = xr; xs < w; xs += 4 * xstep) {
    // four source pixels per trip: eight independent 8-byte loads in flight per thread
    uint2 a0[4], a1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int xx = xs + u * xstep;
      if (xx < w) {
        a0[u] = *reinterpret_cast<const uint2*>(r0 + xx * C);
        a1[u] = *reinterpret_cast<const uint2*>(r1 + xx * C);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int xx = xs + u * xstep;
      if (xx < w) {
        const float4 v0 = unpack4(a0[u]), v1 = unpack4(a1[u]);
        *reinterpret_cast<float4*>(mix + xx * Cs + c) =
            make_float4(fmaf(ly, v1.x - v0.x, v0.x), fmaf(ly, v1.y - v0.y, v0.y), fmaf(ly, v1.z - v0.z, v0.z),
                        fmaf(ly, v1.w - v0.w, v0.w));
      }
    }
  }
  __syncthreads();
  __nv_bfloat16* o = out + static_cast<long long>(row) * W * C + cs0 + c;
  const float* mc = mix + c;
#pragma unroll 4
  for (int X = xr; X < W; X += xstep) {
    const int o0 = tx0[X];
    const float lx = tlx[X];
    const float4 p = *reinterpret_cast<const float4*>(mc + o0);
    const float4 q = *reinterpret_cast<const float4*>(mc + o0 + ((lx != 0.f) ? Cs : 0));
    uint2 ov;
    ov.x = pack_bf16x2(fmaf(lx, q.x - p.x, p.x), fmaf(lx, q.y - p.y, p.y));
    ov.y = pack_bf16x2(fmaf(lx, q.z - p.z, p.z), fmaf(lx, q.w - p.w, p.w));
    *reinterpret_cast<uint2*>(o + X * C) = ov;
  }
}

// weight of high-res index I = 2*i - 2 + k (k = 0..4) on low-res index i; 0 outside the image
__device__ __forceinline__ float up2_adjoint_weight(int I, int i, int n_lo, int n_hi, float scale) {
  if (I < 0 || I >= n_hi) return 0.f;
  int i0, i1;
  float l;
  bilinear_src_scaled(I, n_lo, scale, &i0, &i1, &l);
  float wgt = 0.f;
  if (i0 == i) wgt += 1.f - l;
  if (i1 == i) wgt += l;
  return wgt;
}

// Adjoint in gather form (deterministic, no atomics): every low-res pixel sums the contributions of
// the (at most 5x5, typically 4x4) high-res pixels that read it.  One block per low-res row and
// channel slab: vertical gather into shared memory, then the horizontal gather.
__global__ void __launch_bounds__(kUpThreads)
upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dout, int B, int h, int w, int C, int Cs, float sy, float sx,
                      __nv_bfloat16* __restrict__ dx) {
  extern __shared__ float up_smem[];
  float* vs = up_smem;                           // [W][Cs] vertically gathered row, fp32
  float* twx = vs + 2 * w * Cs;                  // [w][5] horizontal weights of taps 2x-2 .. 2x+2
  float* twy = twx + w * 5;                      // [5] vertical weights of rows 2y-2 .. 2y+2 (block-uniform)
  const int cg4 = Cs >> 2;
  const int c = (threadIdx.x % cg4) * 4;
  const int xr = threadIdx.x / cg4, xstep = kUpThreads / cg4;
  const int H = 2 * h, W = 2 * w;
  const int row = blockIdx.x;
  const int cs0 = blockIdx.y * Cs;
  const int b = row / h, y = row - b * h;
  if (threadIdx.x < 5) twy[threadIdx.x] = up2_adjoint_weight(2 * y - 2 + static_cast<int>(threadIdx.x), y, h, H, sy);
  for (int t = threadIdx.x; t < w * 5; t += kUpThreads) {
    const int xx = t / 5, k = t - xx * 5;
    twx[t] = up2_adjoint_weight(2 * xx - 2 + k, xx, w, W, sx);
  }
  __syncthreads();
  float wy[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) wy[i] = twy[i];
  // rows 2y-2 .. 2y+2 clamped into the image (a clamped row has weight 0)
  const __nv_bfloat16* rp[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    int Yr = 2 * y - 2 + i;
    Yr = Yr < 0 ? 0 : (Yr >= H ? H - 1 : Yr);
    rp[i] = dout + (static_cast<long long>(b) * H + Yr) * W * C + cs0 + c;
  }
  constexpr int kCols = 4;
  for (int X = xr; X < W; X += kCols * xstep) {
    // four high-res columns per trip: twenty independent 8-byte loads in flight per thread (the kernel is
    // bound by DRAM latency x bytes in flight: 3 blocks / SM of staging leave only 768 threads per SM)
    uint2 r[kCols][5];
#pragma unroll
    for (int u = 0; u < kCols; ++u) {
      const int Xu = (X + u * xstep < W) ? X + u * xstep : X;
#pragma unroll
      for (int i = 0; i < 5; ++i) r[u][i] = *reinterpret_cast<const uint2*>(rp[i] + Xu * C);
    }
#pragma unroll
    for (int u = 0; u < kCols; ++u) {
      float4 sa = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float4 ga = unpack4(r[u][i]);
        sa.x = fmaf(wy[i], ga.x, sa.x); sa.y = fmaf(wy[i], ga.y, sa.y);
        sa.z = fmaf(wy[i], ga.z, sa.z); sa.w = fmaf(wy[i], ga.w, sa.w);
      }
      if (X + u * xstep < W) *reinterpret_cast<float4*>(vs + (X + u * xstep) * Cs + c) = sa;
    }
  }
  __syncthreads();
  __nv_bfloat16* o = dx + static_cast<long long>(row) * w * C + cs0 + c;
  const float* vc = vs + c;
#pragma unroll 2
  for (int xx = xr; xx < w; xx += xstep) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float wgt = twx[xx * 5 + k];
      int Xs = 2 * xx - 2 + k;
      Xs = Xs < 0 ? 0 : (Xs >= W ? W - 1 : Xs);       // out-of-image taps carry weight 0
      const float4 p = *reinterpret_cast<const float4*>(vc + Xs * Cs);
      acc.x = fmaf(wgt, p.x, acc.x); acc.y = fmaf(wgt, p.y, acc.y);
      acc.z = fmaf(wgt, p.z, acc.z); acc.w = fmaf(wgt, p.w, acc.w);
    }
    uint2 ov;
    ov.x = pack_bf16x2(acc.x, acc.y);
    ov.y = pack_bf16x2(acc.z, acc.w);
    *reinterpret_cast<uint2*>(o + xx * C) = ov;
  }
}

// ----------------------------------------------------------------------------
// host launchers
// ----------------------------------------------------------------------------


int bn_finalize(float* stats, int parts, int C, long long count, const float* gamma,
                const float* beta, float eps, float momentum, float* running_mean,
                float* running_var, long long* nbt, float* scale, float* shift,
                float* save_mean, float* save_rstd, cudaStream_t stream) {
  UGP_REQUIRE(stats && gamma && beta && scale && shift && save_mean && save_rstd,
              "bn_finalize: null pointer");
  UGP_REQUIRE(parts >= 1, "bn_finalize: parts must be >= 1");
  UGP_CHECK_CUDA(launch_k(bn_finalize_kernel, dim3((C + 31) / 32), dim3(32 * kFinGroups), 0, stream,
                          stats, parts, C, static_cast<float>(count), gamma, beta, eps, momentum, running_mean,
                          running_var, nbt, scale, shift, save_mean, save_rstd));
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
  UGP_REQUIRE(C % 8 == 0 && 256 % (C / 8) == 0, "bn_relu_apply: C must be 8 * a power of two <= 2048 (got %d)", C);
  if (pooled) {
    UGP_REQUIRE(H % 2 == 0 && W % 2 == 0, "bn_relu_apply: pooling needs even H, W");
    const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
    UGP_CHECK_CUDA(launch_k(bn_relu_apply_kernel<true>, dim3(grid_for(total, 256)), dim3(256), 0, stream,
                            static_cast<const __nv_bfloat16*>(z), scale, shift, B, H, W, C,
                            static_cast<__nv_bfloat16*>(a), static_cast<__nv_bfloat16*>(pooled)));
  } else {
    const long long total = static_cast<long long>(B) * H * W * (C / 8);
    UGP_CHECK_CUDA(launch_k(bn_relu_apply_kernel<false>, dim3(grid_for(total, 256)), dim3(256), 0, stream,
                            static_cast<const __nv_bfloat16*>(z), scale, shift, B, H, W, C,
                            static_cast<__nv_bfloat16*>(a), static_cast<__nv_bfloat16*>(nullptr)));
  }
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int bn_relu_apply_head(const void* z, const float* scale, const float* shift, int B, int H, int W, int C,
                       const float* wh, const float* bh, int nc, void* a, float* lmap, cudaStream_t stream) {
  UGP_REQUIRE(z && scale && shift && wh && bh && a && lmap, "bn_relu_apply_head: null pointer");
  UGP_REQUIRE(C == 64 || C == 128 || C == 256, "bn_relu_apply_head: C must be 64, 128 or 256 (got %d)", C);
  UGP_REQUIRE(nc == 1 || nc == 2, "bn_relu_apply_head: 1 or 2 classes (got %d)", nc);
  const long long M = static_cast<long long>(B) * H * W;
  const long long total = M * (C / 8);
  const int grid = grid_for(total, 256);
  if (nc == 1)
    UGP_CHECK_CUDA(launch_k(bn_relu_apply_head_kernel<1>, dim3(grid), dim3(256), 0, stream,
                            static_cast<const __nv_bfloat16*>(z), scale, shift, M, C, H * W, wh, bh,
                            static_cast<__nv_bfloat16*>(a), lmap));
  else
    UGP_CHECK_CUDA(launch_k(bn_relu_apply_head_kernel<2>, dim3(grid), dim3(256), 0, stream,
                            static_cast<const __nv_bfloat16*>(z), scale, shift, M, C, H * W, wh, bh,
                            static_cast<__nv_bfloat16*>(a), lmap));
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

// Channel-slab width of the separable upsample kernels: the widest Cs that divides C, keeps
// `pixels * Cs * 4` bytes of staging (+ the tables) inside the shared-memory budget and lets a block of
// kUpThreads threads own whole 4-channel groups (kUpThreads % (Cs / 4) == 0).
static int upsample_slab(int C, int pixels, int table_bytes, int budget) {
  for (int Cs = C; Cs >= 4; Cs -= 4) {
    if (C % Cs != 0 || kUpThreads % (Cs / 4) != 0) continue;
    if (static_cast<long long>(pixels) * Cs * 4 + table_bytes <= budget) return Cs;
  }
  return 0;
}

int upsample2x(const void* x, int B, int h, int w, int C, void* out, cudaStream_t stream) {
  UGP_REQUIRE(x && out, "upsample2x: null pointer");
  UGP_REQUIRE(C % 4 == 0, "upsample2x: C must be a multiple of 4");
  UGP_REQUIRE(B > 0 && h > 0 && w > 0 && static_cast<long long>(B) * 2 * h < (1LL << 31) &&
                  static_cast<long long>(2 * w) * C < (1LL << 28),
              "upsample2x: shape out of range");
  const int tables = 2 * w * 8;
  const float sy = (2 * h > 1) ? static_cast<float>(h - 1) / static_cast<float>(2 * h - 1) : 0.f;
  const float sx = (2 * w > 1) ? static_cast<float>(w - 1) / static_cast<float>(2 * w - 1) : 0.f;
  // staging budget: small enough that a block can share an SM with a ~200 KB tensor-core CTA of another
  // stream (the step overlaps these kernels with weight-gradient / uncertainty-forward convolutions whose
  // persistent CTAs hold every SM: a 64 KB block then waits for a whole convolution to drain)
  const int Cs = upsample_slab(C, w, tables, g_upsample_smem_kb * 1024);
  UGP_REQUIRE(Cs > 0 && C / Cs <= 65535, "upsample2x: row of %d pixels does not fit the staging buffer", w);
  const int smem = w * Cs * 4 + tables;
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(upsample2x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  upsample2x_kernel<<<dim3(B * 2 * h, C / Cs), kUpThreads, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(x), B, h, w, C, Cs, sy, sx, static_cast<__nv_bfloat16*>(out));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int upsample2x_bwd(const void* dout, int B, int h, int w, int C, void* dx,
                   cudaStream_t stream) {
  UGP_REQUIRE(dout && dx, "upsample2x_bwd: null pointer");
  UGP_REQUIRE(C % 4 == 0, "upsample2x_bwd: C must be a multiple of 4");
  UGP_REQUIRE(B > 0 && h > 0 && w > 0 && static_cast<long long>(B) * h < (1LL << 31) &&
                  static_cast<long long>(2 * w) * C < (1LL << 28),
              "upsample2x_bwd: shape out of range");
  const int tables = (w * 5 + 8) * 4;
  const float sy = (2 * h > 1) ? static_cast<float>(h - 1) / static_cast<float>(2 * h - 1) : 0.f;
  const float sx = (2 * w > 1) ? static_cast<float>(w - 1) / static_cast<float>(2 * w - 1) : 0.f;
  const int Cs = upsample_slab(C, 2 * w, tables, g_upsample_smem_kb * 1024);
  UGP_REQUIRE(Cs > 0 && C / Cs <= 65535, "upsample2x_bwd: row of %d pixels does not fit the staging buffer", 2 * w);
  const int smem = 2 * w * Cs * 4 + tables;
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(upsample2x_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  upsample2x_bwd_kernel<<<dim3(B * h, C / Cs), kUpThreads, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(dout), B, h, w, C, Cs, sy, sx, static_cast<__nv_bfloat16*>(dx));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}


// ----------------------------------------------------------------------------
// Global average pool over H*W of an NHWC bf16 feature map -> fp32 [B, C], and its adjoint
// (nn.AdaptiveAvgPool2d(1) + Flatten in front of the Herlev classifier head,
// Herlev/train_herlev.py:66-68).  Deterministic: fixed-order tree, no atomics.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
global_avg_pool_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, float* __restrict__ out) {
  // block = (image b, 64-channel slab); thread = (pixel lane 0..31, 8-channel group 0..7)
  __shared__ float red[32][65];
  const int slabs = C >> 6;
  const int b = blockIdx.x / slabs;
  const int c0 = (blockIdx.x - b * slabs) * 64;
  const int cgi = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const __nv_bfloat16* base = x + static_cast<long long>(b) * HW * C + c0 + cgi * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int p = pl; p < HW; p += 32) {
    const Vec8 v = load8(base + static_cast<long long>(p) * C);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v.v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[pl][cgi * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    out[static_cast<long long>(b) * C + c0 + threadIdx.x] = s / static_cast<float>(HW);
  }
}

__global__ void __launch_bounds__(256)
global_avg_pool_bwd_kernel(const float* __restrict__ dy, int HW, int C, __nv_bfloat16* __restrict__ dx) {
  const int cg = C >> 3;
  const int b = blockIdx.y;
  const float inv = 1.f / static_cast<float>(HW);
  const int items = HW * cg;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += gridDim.x * blockDim.x) {
    const int c0 = (i % cg) * 8;
    Vec8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = __ldg(dy + static_cast<long long>(b) * C + c0 + j) * inv;
    store8(dx + (static_cast<long long>(b) * HW) * C + static_cast<long long>(i) * 8, o);
  }
}

int global_avg_pool(const void* x, int B, int HW, int C, float* out, cudaStream_t stream) {
  UGP_REQUIRE(x && out && B > 0 && HW > 0, "global_avg_pool: bad arguments");
  UGP_REQUIRE(C % 64 == 0, "global_avg_pool: C %d must be a multiple of 64", C);
  global_avg_pool_kernel<<<B * (C / 64), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), HW, C, out);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int global_avg_pool_bwd(const float* dy, int B, int HW, int C, void* dx, cudaStream_t stream) {
  UGP_REQUIRE(dy && dx && B > 0 && B <= 65535 && HW > 0, "global_avg_pool_bwd: bad arguments");
  UGP_REQUIRE(C % 8 == 0, "global_avg_pool_bwd: C %d must be a multiple of 8", C);
  int gx = (HW * (C / 8) + 255) / 256;
  if (gx > 64) gx = 64;
  dim3 grid(gx, B);
  global_avg_pool_bwd_kernel<<<grid, 256, 0, stream>>>(dy, HW, C, static_cast<__nv_bfloat16*>(dx));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
