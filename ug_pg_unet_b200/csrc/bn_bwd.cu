// Backward of conv-output -> BatchNorm2d(train) -> ReLU (-> MaxPool2d(2)).
// Reference: autograd of UG_unet_parts.py:11-12,14-15 (+ :49 for the pool).
//
//   g    = (dA + unpool(dP)) * [a > 0]            a = relu(z*scale + shift)
//   pass 1 (reduce):  sums[0][c] = sum g*xhat (= dgamma), sums[1][c] = sum g (= dbeta)
//   pass 2 (apply):   dz = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat));  dbias = sum dz
//
// HBM-bound: pass 1 reads dA and z, pass 2 reads them again and writes dz (5
// bf16 units per conv-output element).  The plain variant
// streams pixels (4 independent 16-byte loads per array in flight per thread);
// the pooled variant walks 2x2 windows so the max-pool arg-max can be recomputed
// from z with exactly the forward arithmetic.
#include "ugp_common.cuh"
#include "ugp_vec.cuh"

namespace ugp {

int g_bn_skip_reduce = 0;   // what-if timing only (results wrong): skip the reduction pass of the plain variant
int g_bn_bwd_blocks_per_sm = 2;   // persistent blocks per SM of the backward kernels (tuning switch)
int g_bn_min_pixels = 8;   // plain variant: pixels per thread below which the grid shrinks (tuning switch)
int g_bn_pool_vec4 = 1;   // pooled backward: 1 = 4-channel threads, two windows in flight (software pipeline); 0 = 8-channel threads

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
  // fused 1x1 head (OutConv on the activation this BatchNorm produces, UG_unet.py:294-303): the gradient
  // of the activation gets  + sum_k dl[b][k][p] * wh[k][c]  and the head's weight / bias gradients are
  // reduced in the same pass - the head's own backward kernel (and its 2 bytes/element of dx) disappear
  const float* dl;           // [B][nc][H*W] fp32 gradient of the head's output map (nullable)
  const float* wh;           // [nc][C] head weights
  float* dwh;                // [nc][C] (+ [nc] bias gradient behind it), zeroed by the launcher
  int HW;
  int hw_shift;              // log2(HW) when HW is a power of two, else -1
  // raw != nullptr: the reduction pass already ran inside the convolution that produced dA
  // (ugp_conv_dgrad_bnstats_bf16): raw[0][c] = sum g*[a > 0], raw[1][c] = sum g*[a > 0]*z.  The apply pass
  // derives dgamma / dbeta from them and its first block publishes them in `sums`.
  const float* raw;
};

constexpr int kBnThreads = 256;
// independent 16-byte loads per array in flight per thread: 8 in the reduce pass, 4 in the apply
// pass (which also needs the three per-channel coefficient sets in registers)
template <bool kApply> struct BnUnroll { static constexpr int v = kApply ? 4 : 8; };

__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Block reduction over the rows that share a channel group, then one vector
// reduction per four channels per block.
template <int kVals>
__device__ __forceinline__ void block_reduce_channels(float (*red)[kBnThreads][9], const float* acc0,
                                                      const float* acc1, int cg, int my_cg,
                                                      int my_row, int rows_per_block, float* out0,
                                                      float* out1, int c0) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][threadIdx.x][j] = acc0[j];
    if (kVals > 1) red[1][threadIdx.x][j] = acc1[j];
  }
  __syncthreads();
  if (my_row == 0) {
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a0[j] = acc0[j]; a1[j] = (kVals > 1) ? acc1[j] : 0.f; }
    for (int r = 1; r < rows_per_block; ++r) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0[j] += red[0][r * cg + my_cg][j];
        if (kVals > 1) a1[j] += red[1][r * cg + my_cg][j];
      }
    }
    // two 16-byte vector reductions per sum instead of eight scalar ones: the L2 atomic unit serialises
    // same-address operations, and every block of the launch lands on the same 2*C addresses
    red_add_f32x4(out0 + c0, a0[0], a0[1], a0[2], a0[3]);
    red_add_f32x4(out0 + c0 + 4, a0[4], a0[5], a0[6], a0[7]);
    if (kVals > 1) {
      red_add_f32x4(out1 + c0, a1[0], a1[1], a1[2], a1[3]);
      red_add_f32x4(out1 + c0 + 4, a1[4], a1[5], a1[6], a1[7]);
    }
  }
}

// ---------------------------------------------------------------------------
// plain variant: no pooled gradient; one pixel x 8 channels per (thread, slot)
// ---------------------------------------------------------------------------
template <bool kApply, int kNC = 0>
__global__ void __launch_bounds__(kBnThreads, 2) bn_relu_bwd_plain_kernel(const BnBwdArgs a) {
  pdl_wait();      // the apply pass is launched while the reduction pass drains (programmatic dependent launch)
  const int cg = a.C >> 3;
  const int my_cg = threadIdx.x % cg;
  const int my_row = threadIdx.x / cg;
  const int rows_per_block = kBnThreads / cg;
  const int c0 = my_cg * 8;
  const long long M = static_cast<long long>(a.B) * a.H * a.W;

  float sc[8], sh[8];
  float k0[8], kz[8], kc[8];   // apply: dz = k0*g + kz*z + kc ; reduce: kz = mean
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(a.scale + c0 + j);
    sh[j] = __ldg(a.shift + c0 + j);
    const float mu = __ldg(a.mean + c0 + j);
    if (kApply) {
      const float rs = __ldg(a.rstd + c0 + j);
      const float gr = __ldg(a.gamma + c0 + j) * rs;
      float dbeta, dgamma;
      if (a.raw != nullptr) {
        dbeta = __ldg(a.raw + c0 + j);
        dgamma = rs * (__ldg(a.raw + a.C + c0 + j) - mu * dbeta);      // sum g*xhat = rstd*(sum g*z - mean*sum g)
        if (blockIdx.x == 0 && my_row == 0) { a.sums[c0 + j] = dgamma; a.sums[a.C + c0 + j] = dbeta; }
      } else {
        dbeta = __ldg(a.sums + a.C + c0 + j);
        dgamma = __ldg(a.sums + c0 + j);
      }
      const float mg = dbeta * a.inv_count;       // mean g
      const float mgx = dgamma * a.inv_count;     // mean g*xhat
      // dz = gr*(g - mg - (z-mu)*rs*mgx)
      k0[j] = gr;
      kz[j] = -gr * rs * mgx;
      kc[j] = -gr * mg - mu * kz[j];
    } else {
      kz[j] = mu;
      k0[j] = 0.f;
      kc[j] = 0.f;
    }
  }
  float acc0[8], acc1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

  // Software pipeline: the raw 16-byte words of trip t+1 are requested BEFORE trip t is consumed.  Without
  // it every warp of the launch sits in the same phase (all loading, then all computing): measured
  // (ncu, r02) 16 warps/SM at 45 % issue utilisation and 2.4 TB/s on the 67 MB layers, ~7 us per trip.
  constexpr int kHead = kNC > 0 ? kNC : 1;
  // pixels per trip (two trips in flight); the head variant carries 16*kNC more live registers
  constexpr int kU = kNC > 0 ? 2 : BnUnroll<kApply>::v / 2;
  const long long stride = static_cast<long long>(gridDim.x) * rows_per_block;
  const long long step = stride * kU;
  const bool has_dA = a.dA != nullptr;             // head only (the last decoder block): no tensor gradient
  float wh[kHead][8], hacc[kHead][8], hb[kHead];
#pragma unroll
  for (int k = 0; k < kHead; ++k) {
    hb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wh[k][j] = (kNC > 0) ? __ldg(a.wh + k * a.C + c0 + j) : 0.f;
      hacc[k][j] = 0.f;
    }
  }
  struct Raw {
    uint4 z[kU], g[kU];
    float dl[kU][kHead];
  };
  Raw rA, rB;
  auto fetch = [&](Raw& r, long long p) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long pp = p + u * stride;
      if (pp < M) {
        r.z[u] = *reinterpret_cast<const uint4*>(a.z + pp * a.C + c0);
        r.g[u] = has_dA ? *reinterpret_cast<const uint4*>(a.dA + pp * a.C + c0) : make_uint4(0u, 0u, 0u, 0u);
        if (kNC == 1) {
          r.dl[u][0] = __ldg(a.dl + pp);              // [B][1][HW]: the map is indexed by the pixel itself
        } else if (kNC > 1) {
          // [B][nc][HW]: (b*nc + k)*HW + q = pixel + (b*(nc-1) + k)*HW; no division for power-of-two maps
          const int pix = static_cast<int>(pp);
          const int b = (a.hw_shift >= 0) ? (pix >> a.hw_shift) : pix / a.HW;
#pragma unroll
          for (int k = 0; k < kHead; ++k)
            r.dl[u][k] = __ldg(a.dl + pp + static_cast<long long>(b * (kNC - 1) + k) * a.HW);
        }
      }
    }
  };
  auto consume = [&](const Raw& r, long long p) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long pp = p + u * stride;
      if (pp < M) {
        Vec8 o;
        const Vec8 zv = unpack8(r.z[u]);
        const Vec8 gv = unpack8(r.g[u]);
        if (kNC > 0 && !kApply && my_cg == 0) {
#pragma unroll
          for (int k = 0; k < kHead; ++k) hb[k] += r.dl[u][k];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = zv.v[j];
          const float act = fmaf(z, sc[j], sh[j]);          // > 0: the ReLU passes; then also the activation
          float gin = gv.v[j];
          if (kNC > 0) {
#pragma unroll
            for (int k = 0; k < kHead; ++k) {
              gin = fmaf(r.dl[u][k], wh[k][j], gin);
              if (!kApply) hacc[k][j] = fmaf(r.dl[u][k], fmaxf(act, 0.f), hacc[k][j]);
            }
          }
          const float g = (act > 0.f) ? gin : 0.f;
          if (!kApply) {
            acc0[j] = fmaf(g, z - kz[j], acc0[j]);   // sum g*(z-mean); * rstd at the end
            acc1[j] += g;
          } else {
            const float d = fmaf(k0[j], g, fmaf(kz[j], z, kc[j]));
            o.v[j] = d;
            acc0[j] += d;
          }
        }
        if (kApply) store8(a.dz + pp * a.C + c0, o);
      }
    }
  };
  long long p = static_cast<long long>(blockIdx.x) * rows_per_block + my_row;
  fetch(rA, p);
  while (p < M) {
    fetch(rB, p + step);
    consume(rA, p);
    p += step;
    if (p >= M) break;
    fetch(rA, p + step);
    consume(rB, p);
    p += step;
  }
  __shared__ float red[2][kBnThreads][9];
  if (!kApply) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc0[j] *= __ldg(a.rstd + c0 + j);
    block_reduce_channels<2>(red, acc0, acc1, cg, my_cg, my_row, rows_per_block, a.sums,
                             a.sums + a.C, c0);
    if (kNC > 0) {
      // head weight gradient dwh[k][c] = sum_p dl[k][p] * a[p][c]; bias gradient = sum_p dl[k][p]
#pragma unroll
      for (int k = 0; k < kHead; ++k) {
        __syncthreads();
        block_reduce_channels<1>(red, hacc[k], hacc[k], cg, my_cg, my_row, rows_per_block, a.dwh + k * a.C,
                                 nullptr, c0);
      }
      __syncthreads();
      if (my_cg == 0) {
#pragma unroll
        for (int k = 0; k < kHead; ++k) red[0][my_row][k] = hb[k];
      }
      __syncthreads();
      if (threadIdx.x < kHead) {
        float sum = 0.f;
        for (int r = 0; r < rows_per_block; ++r) sum += red[0][r][threadIdx.x];
        atomicAdd(a.dwh + kNC * a.C + threadIdx.x, sum);
      }
    }
  } else if (a.dbias) {
    block_reduce_channels<1>(red, acc0, acc1, cg, my_cg, my_row, rows_per_block, a.dbias, nullptr, c0);
  }
}

// ---------------------------------------------------------------------------
// pooled variant: thread = 2x2 window x 8 channels
// ---------------------------------------------------------------------------
template <bool kApply>
__global__ void __launch_bounds__(kBnThreads, 2) bn_relu_bwd_pool_kernel(const BnBwdArgs a) {
  pdl_wait();      // the apply pass is launched while the reduction pass drains (programmatic dependent launch)
  const int cg = a.C >> 3;
  const int Ho = a.H >> 1, Wo = a.W >> 1;
  const long long windows = static_cast<long long>(a.B) * Ho * Wo;
  const int my_cg = threadIdx.x % cg;
  const int c0 = my_cg * 8;
  const int rows_per_block = kBnThreads / cg;
  const int my_row = threadIdx.x / cg;

  float sc[8], sh[8], k0[8], kz[8], kc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = __ldg(a.scale + c0 + j);
    sh[j] = __ldg(a.shift + c0 + j);
    const float mu = __ldg(a.mean + c0 + j);
    if (kApply) {
      const float rs = __ldg(a.rstd + c0 + j);
      const float gr = __ldg(a.gamma + c0 + j) * rs;
      const float mg = __ldg(a.sums + a.C + c0 + j) * a.inv_count;
      const float mgx = __ldg(a.sums + c0 + j) * a.inv_count;
      k0[j] = gr;
      kz[j] = -gr * rs * mgx;
      kc[j] = -gr * mg - mu * kz[j];
    } else {
      kz[j] = mu;
      k0[j] = 0.f;
      kc[j] = 0.f;
    }
  }
  float acc0[8], acc1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

  for (long long win = static_cast<long long>(blockIdx.x) * rows_per_block + my_row;
       win < windows; win += static_cast<long long>(gridDim.x) * rows_per_block) {
    const int xo = static_cast<int>(win % Wo);
    const long long t2 = win / Wo;
    const int yo = static_cast<int>(t2 % Ho);
    const int b = static_cast<int>(t2 / Ho);
    // raw bf16x8 words stay packed until they are consumed (register footprint: 2 blocks / SM)
    uint32_t zr[4][4], gr[4][4], pr[4];
    long long pix[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pix[i] = (static_cast<long long>(b) * a.H + (2 * yo + (i >> 1))) * a.W + 2 * xo + (i & 1);
      const uint4 zq = *reinterpret_cast<const uint4*>(a.z + pix[i] * a.C + c0);
      zr[i][0] = zq.x; zr[i][1] = zq.y; zr[i][2] = zq.z; zr[i][3] = zq.w;
      if (a.dA) {
        const uint4 gq = *reinterpret_cast<const uint4*>(a.dA + pix[i] * a.C + c0);
        gr[i][0] = gq.x; gr[i][1] = gq.y; gr[i][2] = gq.z; gr[i][3] = gq.w;
      } else {
        gr[i][0] = gr[i][1] = gr[i][2] = gr[i][3] = 0u;
      }
    }
    {
      const uint4 pq = *reinterpret_cast<const uint4*>(a.dP + win * a.C + c0);
      pr[0] = pq.x; pr[1] = pq.y; pr[2] = pq.z; pr[3] = pq.w;
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      float dlo[4], dhi[4];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int j = 2 * w + hh;
        float zf[4];
        int arg = 0;
        float best = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          zf[i] = hh ? bf16_hi(zr[i][w]) : bf16_lo(zr[i][w]);
          const float av = round_bf16(bn_relu_val(zf[i], sc[j], sh[j]));
          if (i == 0 || av > best) { best = av; arg = i; }     // first max wins (MaxPool2d)
        }
        const float gpv = hh ? bf16_hi(pr[w]) : bf16_lo(pr[w]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float g = hh ? bf16_hi(gr[i][w]) : bf16_lo(gr[i][w]);
          if (arg == i) g += gpv;
          g = (fmaf(zf[i], sc[j], sh[j]) > 0.f) ? g : 0.f;
          if (!kApply) {
            acc0[j] = fmaf(g, zf[i] - kz[j], acc0[j]);
            acc1[j] += g;
          } else {
            const float d = fmaf(k0[j], g, fmaf(kz[j], zf[i], kc[j]));
            if (hh) dhi[i] = d; else dlo[i] = d;
            acc0[j] += d;
          }
        }
      }
      if (kApply) {
#pragma unroll
        for (int i = 0; i < 4; ++i) gr[i][w] = pack_bf16x2(dlo[i], dhi[i]);
      }
    }
    if (kApply) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(a.dz + pix[i] * a.C + c0) = make_uint4(gr[i][0], gr[i][1], gr[i][2], gr[i][3]);
    }
  }
  __shared__ float red[2][kBnThreads][9];
  if (!kApply) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc0[j] *= __ldg(a.rstd + c0 + j);
    block_reduce_channels<2>(red, acc0, acc1, cg, my_cg, my_row, rows_per_block, a.sums,
                             a.sums + a.C, c0);
  } else if (a.dbias) {
    block_reduce_channels<1>(red, acc0, acc1, cg, my_cg, my_row, rows_per_block, a.dbias, nullptr, c0);
  }
}

// ---------------------------------------------------------------------------
// pooled variant, 4 channels per thread (8-byte loads): half the register footprint of the 8-channel
// version, so 3 blocks (24 warps) per SM hide the DRAM latency of its load -> compute -> store trips.
// ---------------------------------------------------------------------------
constexpr int kPoolBlocksPerSM = 2;

struct PoolRaw4 {          // the raw words of one 2x2 window x 4 channels
  uint2 z[4], g[4], p;
};

template <bool kApply>
__global__ void __launch_bounds__(kBnThreads, kPoolBlocksPerSM) bn_relu_bwd_pool4_kernel(const BnBwdArgs a) {
  pdl_wait();      // the apply pass is launched while the reduction pass drains (programmatic dependent launch)
  const int cg = a.C >> 2;
  const int Ho = a.H >> 1, Wo = a.W >> 1;
  const long long windows = static_cast<long long>(a.B) * Ho * Wo;
  const int my_cg = threadIdx.x % cg;
  const int c0 = my_cg * 4;
  const int rows_per_block = kBnThreads / cg;
  const int my_row = threadIdx.x / cg;

  float sc[4], sh[4], k0[4], kz[4], kc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = __ldg(a.scale + c0 + j);
    sh[j] = __ldg(a.shift + c0 + j);
    const float mu = __ldg(a.mean + c0 + j);
    if (kApply) {
      const float rs = __ldg(a.rstd + c0 + j);
      const float gr = __ldg(a.gamma + c0 + j) * rs;
      const float mg = __ldg(a.sums + a.C + c0 + j) * a.inv_count;
      const float mgx = __ldg(a.sums + c0 + j) * a.inv_count;
      k0[j] = gr;
      kz[j] = -gr * rs * mgx;
      kc[j] = -gr * mg - mu * kz[j];
    } else {
      kz[j] = mu;
      k0[j] = 0.f;
      kc[j] = 0.f;
    }
  }
  float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};

  auto pixel = [&](long long win, int i) -> long long {
    const int xo = static_cast<int>(win % Wo);
    const long long t2 = win / Wo;
    const int yo = static_cast<int>(t2 % Ho);
    const int b = static_cast<int>(t2 / Ho);
    return (static_cast<long long>(b) * a.H + (2 * yo + (i >> 1))) * a.W + 2 * xo + (i & 1);
  };
  auto fetch = [&](PoolRaw4& r, long long win) {
    if (win >= windows) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long pix = pixel(win, i);
      r.z[i] = *reinterpret_cast<const uint2*>(a.z + pix * a.C + c0);
      r.g[i] = a.dA ? *reinterpret_cast<const uint2*>(a.dA + pix * a.C + c0) : make_uint2(0u, 0u);
    }
    r.p = *reinterpret_cast<const uint2*>(a.dP + win * a.C + c0);
  };
  auto consume = [&](PoolRaw4& r, long long win) {
    uint32_t out[4][2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      float dlo[4], dhi[4];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int j = 2 * w + hh;
        float zf[4];
        int arg = 0;
        float best = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t zw = w ? r.z[i].y : r.z[i].x;
          zf[i] = hh ? bf16_hi(zw) : bf16_lo(zw);
          const float av = round_bf16(bn_relu_val(zf[i], sc[j], sh[j]));
          if (i == 0 || av > best) { best = av; arg = i; }     // first max wins (MaxPool2d)
        }
        const uint32_t pw = w ? r.p.y : r.p.x;
        const float gpv = hh ? bf16_hi(pw) : bf16_lo(pw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t gw = w ? r.g[i].y : r.g[i].x;
          float g = hh ? bf16_hi(gw) : bf16_lo(gw);
          if (arg == i) g += gpv;
          g = (fmaf(zf[i], sc[j], sh[j]) > 0.f) ? g : 0.f;
          if (!kApply) {
            acc0[j] = fmaf(g, zf[i] - kz[j], acc0[j]);
            acc1[j] += g;
          } else {
            const float d = fmaf(k0[j], g, fmaf(kz[j], zf[i], kc[j]));
            if (hh) dhi[i] = d; else dlo[i] = d;
            acc0[j] += d;
          }
        }
      }
      if (kApply) {
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i][w] = pack_bf16x2(dlo[i], dhi[i]);
      }
    }
    if (kApply) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint2*>(a.dz + pixel(win, i) * a.C + c0) = make_uint2(out[i][0], out[i][1]);
    }
  };

  // software pipeline, two windows in flight per thread (see the plain kernel)
  const long long step = static_cast<long long>(gridDim.x) * rows_per_block;
  long long win = static_cast<long long>(blockIdx.x) * rows_per_block + my_row;
  PoolRaw4 ra, rb;
  fetch(ra, win);
  while (win < windows) {
    fetch(rb, win + step);
    consume(ra, win);
    win += step;
    if (win >= windows) break;
    fetch(ra, win + step);
    consume(rb, win);
    win += step;
  }

  // block reduction over the windows that share a channel group, one vector reduction per 4 channels per block
  __shared__ float red[2][kBnThreads][5];
  if (!kApply) {
#pragma unroll
    for (int j = 0; j < 4; ++j) acc0[j] *= __ldg(a.rstd + c0 + j);
  } else if (!a.dbias) {
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[0][threadIdx.x][j] = acc0[j]; red[1][threadIdx.x][j] = acc1[j]; }
  __syncthreads();
  if (my_row == 0) {
    for (int r = 1; r < rows_per_block; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc0[j] += red[0][r * cg + my_cg][j]; acc1[j] += red[1][r * cg + my_cg][j]; }
    }
    if (!kApply) {
      red_add_f32x4(a.sums + c0, acc0[0], acc0[1], acc0[2], acc0[3]);
      red_add_f32x4(a.sums + a.C + c0, acc1[0], acc1[1], acc1[2], acc1[3]);
    } else {
      red_add_f32x4(a.dbias + c0, acc0[0], acc0[1], acc0[2], acc0[3]);
    }
  }
}

int bn_relu_bwd_ex(const void* z, const void* dA, const void* dP, const float* scale,
                   const float* shift, const float* mean, const float* rstd,
                   const float* gamma, int B, int H, int W, int C, float* sums,
                   float* dbias, void* dz, const float* head_dl, const float* head_w, int nc,
                   float* head_grads, const float* raw, cudaStream_t stream) {
  UGP_REQUIRE(z && (dA || dP || head_dl) && scale && shift && mean && rstd && gamma && sums && dz,
              "bn_relu_bwd: null pointer");
  if (head_dl) {
    UGP_REQUIRE(head_w && head_grads && (nc == 1 || nc == 2) && dP == nullptr,
                "bn_relu_bwd: the fused head takes 1 or 2 classes and no pooled gradient (nc = %d)", nc);
    UGP_REQUIRE(static_cast<long long>(B) * H * W < (1LL << 31), "bn_relu_bwd: too many pixels for the fused head");
  }
  UGP_REQUIRE(C % 64 == 0 && C <= 2048 && kBnThreads % (C / 8) == 0,
              "bn_relu_bwd: C %d unsupported", C);
  BnBwdArgs a;
  a.z = static_cast<const __nv_bfloat16*>(z);
  a.dA = static_cast<const __nv_bfloat16*>(dA);
  a.dP = static_cast<const __nv_bfloat16*>(dP);
  a.scale = scale; a.shift = shift; a.mean = mean; a.rstd = rstd; a.gamma = gamma;
  a.sums = sums; a.dbias = dbias; a.dz = static_cast<__nv_bfloat16*>(dz);
  a.B = B; a.H = H; a.W = W; a.C = C;
  const long long M = static_cast<long long>(B) * H * W;
  a.inv_count = 1.f / static_cast<float>(M);
  a.dl = head_dl; a.wh = head_w; a.dwh = head_grads; a.HW = H * W;
  a.raw = raw;
  UGP_REQUIRE(raw == nullptr || (dA != nullptr && dP == nullptr && head_dl == nullptr),
              "bn_relu_bwd: precomputed sums go with the plain variant only");
  a.hw_shift = -1;
  for (int sft = 0; sft < 31; ++sft)
    if ((1 << sft) == a.HW) a.hw_shift = sft;
  if (head_dl) UGP_CHECK_CUDA(cudaMemsetAsync(head_grads, 0, sizeof(float) * (static_cast<size_t>(nc) * C + nc), stream));
  const int rows_per_block = kBnThreads / (C / 8);
  if (dbias == sums + 2 * C) {          // the engine's [3][C] block: one fill instead of two
    UGP_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 3 * C, stream));
  } else {
    UGP_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, stream));
    if (dbias) UGP_CHECK_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * C, stream));
  }
  // two resident blocks per SM keep ~64 KB of loads in flight per SM; more blocks only add
  // same-address atomic traffic in the final per-channel reduction
  const long long cap = static_cast<long long>(device_sm_count()) * g_bn_bwd_blocks_per_sm;
  if (dP) {
    UGP_REQUIRE(H % 2 == 0 && W % 2 == 0, "bn_relu_bwd: pooled gradient needs even H, W");
    if (g_bn_pool_vec4 != 0 && kBnThreads % (C / 4) == 0) {
      const int rows4 = kBnThreads / (C / 4);
      long long blocks = (M / 4 + rows4 * 2 - 1) / (rows4 * 2);
      const long long cap4 = static_cast<long long>(device_sm_count()) * kPoolBlocksPerSM;
      if (blocks > cap4) blocks = cap4;
      if (blocks < 1) blocks = 1;
      bn_relu_bwd_pool4_kernel<false><<<static_cast<int>(blocks), kBnThreads, 0, stream>>>(a);
      UGP_CHECK_CUDA(cudaGetLastError());
      UGP_CHECK_CUDA(launch_k(bn_relu_bwd_pool4_kernel<true>, dim3(static_cast<int>(blocks)), dim3(kBnThreads), 0, stream, a));
    } else {
      long long blocks = (M / 4 + rows_per_block * 4 - 1) / (rows_per_block * 4);
      if (blocks > cap) blocks = cap;
      if (blocks < 1) blocks = 1;
      bn_relu_bwd_pool_kernel<false><<<static_cast<int>(blocks), kBnThreads, 0, stream>>>(a);
      UGP_CHECK_CUDA(cudaGetLastError());
      UGP_CHECK_CUDA(launch_k(bn_relu_bwd_pool_kernel<true>, dim3(static_cast<int>(blocks)), dim3(kBnThreads), 0, stream, a));
    }
  } else {
    // at least g_bn_min_pixels pixels per thread: small tensors get few blocks, which keeps the
    // per-channel atomic tail short
    long long blocks = (M + rows_per_block * g_bn_min_pixels - 1) / (rows_per_block * g_bn_min_pixels);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (head_dl && nc == 1) {
      bn_relu_bwd_plain_kernel<false, 1><<<static_cast<int>(blocks), kBnThreads, 0, stream>>>(a);
      UGP_CHECK_CUDA(cudaGetLastError());
      UGP_CHECK_CUDA(launch_k(bn_relu_bwd_plain_kernel<true, 1>, dim3(static_cast<int>(blocks)), dim3(kBnThreads), 0, stream, a));
    } else if (head_dl) {
      bn_relu_bwd_plain_kernel<false, 2><<<static_cast<int>(blocks), kBnThreads, 0, stream>>>(a);
      UGP_CHECK_CUDA(cudaGetLastError());
      UGP_CHECK_CUDA(launch_k(bn_relu_bwd_plain_kernel<true, 2>, dim3(static_cast<int>(blocks)), dim3(kBnThreads), 0, stream, a));
    } else {
      if (!g_bn_skip_reduce && raw == nullptr)
        bn_relu_bwd_plain_kernel<false><<<static_cast<int>(blocks), kBnThreads, 0, stream>>>(a);
      UGP_CHECK_CUDA(cudaGetLastError());
      UGP_CHECK_CUDA(launch_k(bn_relu_bwd_plain_kernel<true>, dim3(static_cast<int>(blocks)), dim3(kBnThreads), 0, stream, a));
    }
  }
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int bn_relu_bwd(const void* z, const void* dA, const void* dP, const float* scale,
                const float* shift, const float* mean, const float* rstd,
                const float* gamma, int B, int H, int W, int C, float* sums,
                float* dbias, void* dz, cudaStream_t stream) {
  return bn_relu_bwd_ex(z, dA, dP, scale, shift, mean, rstd, gamma, B, H, W, C, sums, dbias, dz, nullptr, nullptr,
                        0, nullptr, nullptr, stream);
}

}  // namespace ugp
