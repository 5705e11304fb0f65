// Fused RMSprop over a flat fp32 parameter buffer.
// Reference: optim.RMSprop(lr, weight_decay=1e-4) at
// uncertainty_guided_trainer.py:84-88 (alpha .99, eps 1e-8, momentum 0, not
// centered): g += wd*p; v = alpha*v + (1-alpha)*g^2; p -= lr*g/(sqrt(v)+eps).
#include "ugp_common.cuh"

namespace ugp {

__global__ void rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g,
                               float* __restrict__ v, long long n, float lr, float alpha,
                               float eps, float wd, float gscale) {
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  for (long long i = tid; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pp.x, pp.y, pp.z, pp.w};
    const float ga[4] = {gg.x, gg.y, gg.z, gg.w};
    float va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = fmaf(wd, pa[j], ga[j] * gscale);
      va[j] = alpha * va[j] + (1.f - alpha) * gr * gr;
      pa[j] -= lr * gr / (sqrtf(va[j]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
  }
  for (long long i = (n4 << 2) + tid; i < n; i += stride) {
    const float gr = fmaf(wd, p[i], g[i] * gscale);
    const float vn = alpha * v[i] + (1.f - alpha) * gr * gr;
    v[i] = vn;
    p[i] -= lr * gr / (sqrtf(vn) + eps);
  }
}

int rmsprop_step(float* param, const float* grad, float* square_avg, long long n, float lr,
                 float alpha, float eps, float weight_decay, float grad_scale,
                 cudaStream_t stream) {
  UGP_REQUIRE(param && grad && square_avg && n > 0, "rmsprop_step: bad arguments");
  UGP_REQUIRE((reinterpret_cast<uintptr_t>(param) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(grad) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(square_avg) & 15) == 0,
              "rmsprop_step: buffers must be 16-byte aligned");
  long long blocks = ((n >> 2) + 255) / 256;
  const long long cap = static_cast<long long>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  rmsprop_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(param, grad, square_avg, n, lr,
                                                              alpha, eps, weight_decay,
                                                              grad_scale);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp

// ---------------------------------------------------------------------------
// Multi-tensor variant: one launch updates up to kRmsMaxTensors parameter
// tensors (pointer table passed by value in the kernel parameters), replacing
// the ~10 foreach kernels of torch.optim.RMSprop over the model's 80 tensors.
// ---------------------------------------------------------------------------
namespace ugp {

constexpr int kRmsMaxTensors = 96;

struct RmsTable {
  float* p[kRmsMaxTensors];
  const float* g[kRmsMaxTensors];
  float* v[kRmsMaxTensors];
  int n[kRmsMaxTensors];
};

__global__ void __launch_bounds__(256)
rmsprop_multi_kernel(const __grid_constant__ RmsTable t, float lr, float alpha, float eps, float wd,
                     float gscale) {
  const int k = blockIdx.y;
  float* __restrict__ p = t.p[k];
  const float* __restrict__ g = t.g[k];
  float* __restrict__ v = t.v[k];
  const int n = t.n[k];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gr = fmaf(wd, pv, g[i] * gscale);
    const float vn = alpha * v[i] + (1.f - alpha) * gr * gr;
    v[i] = vn;
    p[i] = pv - lr * gr / (sqrtf(vn) + eps);
  }
}

int rmsprop_multi(float* const* params, const float* const* grads, float* const* avgs,
                  const long long* sizes, int n_tensors, float lr, float alpha, float eps,
                  float weight_decay, float grad_scale, cudaStream_t stream) {
  UGP_REQUIRE(params && grads && avgs && sizes && n_tensors > 0, "rmsprop_multi: bad arguments");
  for (int base = 0; base < n_tensors; base += kRmsMaxTensors) {
    const int cnt = (n_tensors - base < kRmsMaxTensors) ? (n_tensors - base) : kRmsMaxTensors;
    RmsTable t;
    long long biggest = 0;
    for (int i = 0; i < cnt; ++i) {
      UGP_REQUIRE(sizes[base + i] > 0 && sizes[base + i] < (1LL << 31), "rmsprop_multi: bad size");
      t.p[i] = params[base + i];
      t.g[i] = grads[base + i];
      t.v[i] = avgs[base + i];
      t.n[i] = static_cast<int>(sizes[base + i]);
      if (sizes[base + i] > biggest) biggest = sizes[base + i];
    }
    int gx = static_cast<int>((biggest + 256 * 8 - 1) / (256 * 8));
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    dim3 grid(gx, cnt);
    rmsprop_multi_kernel<<<grid, 256, 0, stream>>>(t, lr, alpha, eps, weight_decay, grad_scale);
    UGP_CHECK_CUDA(cudaGetLastError());
  }
  return UGP_OK;
}

}  // namespace ugp

// ---------------------------------------------------------------------------
// Convolution weights: RMSprop straight from the PACKED fp32 weight gradient.
//
// conv_wgrad leaves dWp[tap][ci][co] (what the tensor-core kernel reduces into);
// the forward / input-gradient kernels read bf16 packs Wf[tap][co][ci] and
// Wd[8-tap][ci][co]; the master weights and the optimiser state stay fp32 OIHW
// (state_dict layout of nn.Conv2d / torch.optim.RMSprop).  One launch over all
// 3x3 layers does, per 32co x 32ci x 9 tile staged through shared memory:
//   g = dWp*gscale + wd*w ; v = alpha*v + (1-alpha)*g^2 ; w -= lr*g/(sqrt(v)+eps)
//   Wf, Wd <- bf16(w) ; dWp <- 0
// replacing unpack_wgrad + rmsprop + pack_weights (3 passes, 2 transposes).
// ---------------------------------------------------------------------------
namespace ugp {

extern int g_conv_debug;
constexpr int kCoTile = 32, kCiTile = 32;
constexpr int kCoRow = kCiTile * 9 + 1;      // 289 floats: conflict-free for all three access patterns
constexpr int kConvOptMaxItems = 40;

struct ConvOptItem {
  float* w;              // [Cout][Cin][3][3]
  float* v;              // square_avg, same layout
  float* dwp;            // [9][Cin][Cout]
  __nv_bfloat16* wf;     // [9][Cout][Cin]
  __nv_bfloat16* wd;     // [9][Cin][Cout], taps rotated; may be null
  int cout, cin;
  int tile_begin;        // first global tile index of this tensor
  int ci_tiles;
};
struct ConvOptTable {
  ConvOptItem it[kConvOptMaxItems];
  int n;
};

__global__ void __launch_bounds__(256)
rmsprop_conv_packed_kernel(const __grid_constant__ ConvOptTable t, float lr, float alpha, float eps,
                           float wd_coef, float gscale, int dbg) {
  extern __shared__ float tile[];             // [32][289]
  int k = 0;
  while (k + 1 < t.n && static_cast<int>(blockIdx.x) >= t.it[k + 1].tile_begin) ++k;
  const ConvOptItem& it = t.it[k];
  const int local = blockIdx.x - it.tile_begin;
  const int co0 = (local / it.ci_tiles) * kCoTile;
  const int ci0 = (local % it.ci_tiles) * kCiTile;
  const int Cout = it.cout, Cin = it.cin;

  // packed gradient -> tile[co][ci*9 + tap]; dWp is left zeroed for the next accumulation.
  // Loads are issued in batches of 6 before their dependent stores (the zeroing store may alias the
  // next load as far as the compiler knows, which would serialise one load per DRAM round trip).
  constexpr int kB1 = 6;
  for (int base = threadIdx.x; base < 9 * kCiTile * kCoTile; base += 256 * kB1) {
    float x[kB1];
    float* src[kB1];
#pragma unroll
    for (int u = 0; u < kB1; ++u) {
      const int i = base + u * 256;
      const int r = i % kCoTile;
      const int c = (i / kCoTile) % kCiTile;
      const int tap = i / (kCoTile * kCiTile);
      src[u] = it.dwp + (static_cast<long long>(tap) * Cin + ci0 + c) * Cout + co0 + r;
      x[u] = (dbg & 4) ? 0.f : __ldcs(src[u]);
    }
#pragma unroll
    for (int u = 0; u < kB1; ++u) {
      const int i = base + u * 256;
      const int r = i % kCoTile;
      const int c = (i / kCoTile) % kCiTile;
      const int tap = i / (kCoTile * kCiTile);
      tile[r * kCoRow + c * 9 + tap] = x[u];
      if (!(dbg & 4)) __stcs(src[u], 0.f);
    }
  }
  __syncthreads();
  // update in OIHW order (288-float runs per output channel, float4 accesses); the tile then holds
  // the new weights.  2304 float4 per tile = 9 per thread, in batches of 3.
  constexpr int kB2 = 3;
  if (!(dbg & 8))
  for (int base = threadIdx.x; base < kCoTile * kCiTile * 9 / 4; base += 256 * kB2) {
    float4 pw[kB2], pv[kB2];
#pragma unroll
    for (int u = 0; u < kB2; ++u) {
      const int i = base + u * 256;
      const int r = i / (kCiTile * 9 / 4), j = (i - r * (kCiTile * 9 / 4)) * 4;
      const long long off = (static_cast<long long>(co0 + r) * Cin + ci0) * 9 + j;
      pw[u] = *reinterpret_cast<const float4*>(it.w + off);
      pv[u] = *reinterpret_cast<const float4*>(it.v + off);
    }
#pragma unroll
    for (int u = 0; u < kB2; ++u) {
      const int i = base + u * 256;
      const int r = i / (kCiTile * 9 / 4), j = (i - r * (kCiTile * 9 / 4)) * 4;
      const long long off = (static_cast<long long>(co0 + r) * Cin + ci0) * 9 + j;
      float w4[4] = {pw[u].x, pw[u].y, pw[u].z, pw[u].w};
      float v4[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float gr = fmaf(wd_coef, w4[e], tile[r * kCoRow + j + e] * gscale);
        v4[e] = alpha * v4[e] + (1.f - alpha) * gr * gr;
        w4[e] = w4[e] - lr * gr / (sqrtf(v4[e]) + eps);
        tile[r * kCoRow + j + e] = w4[e];
      }
      *reinterpret_cast<float4*>(it.v + off) = make_float4(v4[0], v4[1], v4[2], v4[3]);
      *reinterpret_cast<float4*>(it.w + off) = make_float4(w4[0], w4[1], w4[2], w4[3]);
    }
  }
  __syncthreads();
  if (dbg & 16) return;
  // forward pack: rows (tap, co), 32 ci contiguous
  for (int i = threadIdx.x; i < 9 * kCoTile * (kCiTile / 2); i += 256) {
    const int cp = i % (kCiTile / 2);
    const int r = (i / (kCiTile / 2)) % kCoTile;
    const int tap = i / ((kCiTile / 2) * kCoTile);
    const uint32_t pk = pack_bf16x2(tile[r * kCoRow + (2 * cp) * 9 + tap], tile[r * kCoRow + (2 * cp + 1) * 9 + tap]);
    *reinterpret_cast<uint32_t*>(it.wf + (static_cast<long long>(tap) * Cout + co0 + r) * Cin + ci0 + 2 * cp) = pk;
  }
  if (it.wd) {
    // input-gradient pack: rows (8 - tap, ci), 32 co contiguous
    for (int i = threadIdx.x; i < 9 * kCiTile * (kCoTile / 2); i += 256) {
      const int rp = i % (kCoTile / 2);
      const int c = (i / (kCoTile / 2)) % kCiTile;
      const int tap = i / ((kCoTile / 2) * kCiTile);
      const uint32_t pk = pack_bf16x2(tile[(2 * rp) * kCoRow + c * 9 + tap], tile[(2 * rp + 1) * kCoRow + c * 9 + tap]);
      *reinterpret_cast<uint32_t*>(it.wd + (static_cast<long long>(8 - tap) * Cin + ci0 + c) * Cout + co0 + 2 * rp) = pk;
    }
  }
}

int rmsprop_conv_packed(float* const* w, float* const* v, float* const* dwp, void* const* wf,
                        void* const* wd, const int* cout, const int* cin, int n_tensors, float lr,
                        float alpha, float eps, float weight_decay, float grad_scale,
                        cudaStream_t stream) {
  UGP_REQUIRE(w && v && dwp && wf && wd && cout && cin && n_tensors > 0, "rmsprop_conv_packed: bad arguments");
  constexpr int kSmem = kCoTile * kCoRow * 4;
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(rmsprop_conv_packed_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  for (int base = 0; base < n_tensors; base += kConvOptMaxItems) {
    const int cnt = (n_tensors - base < kConvOptMaxItems) ? (n_tensors - base) : kConvOptMaxItems;
    ConvOptTable t;
    int tiles = 0;
    for (int i = 0; i < cnt; ++i) {
      const int co = cout[base + i], ci = cin[base + i];
      UGP_REQUIRE(co > 0 && ci > 0 && co % kCoTile == 0 && ci % kCiTile == 0,
                  "rmsprop_conv_packed: Cout %d / Cin %d must be multiples of 32", co, ci);
      UGP_REQUIRE(w[base + i] && v[base + i] && dwp[base + i] && wf[base + i],
                  "rmsprop_conv_packed: null tensor pointer");
      ConvOptItem& it = t.it[i];
      it.w = w[base + i]; it.v = v[base + i]; it.dwp = dwp[base + i];
      it.wf = static_cast<__nv_bfloat16*>(wf[base + i]);
      it.wd = static_cast<__nv_bfloat16*>(wd[base + i]);
      it.cout = co; it.cin = ci;
      it.tile_begin = tiles;
      it.ci_tiles = ci / kCiTile;
      tiles += (co / kCoTile) * it.ci_tiles;
    }
    t.n = cnt;
    rmsprop_conv_packed_kernel<<<tiles, 256, kSmem, stream>>>(t, lr, alpha, eps, weight_decay, grad_scale,
                                                              g_conv_debug);
    UGP_CHECK_CUDA(cudaGetLastError());
  }
  return UGP_OK;
}

}  // namespace ugp

// ---------------------------------------------------------------------------
// Multi-tensor Adam (torch.optim.Adam with L2 weight decay, amsgrad off), the optimiser of the Herlev
// trainer (Herlev/train_herlev.py:170-176): one launch per <= 64 tensors instead of torch's foreach chain.
//   g = grad*grad_scale + wd*p;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)        bc1 = 1 - b1^t, bc2 = 1 - b2^t
// ---------------------------------------------------------------------------
namespace ugp {

constexpr int kAdamMaxTensors = 64;

struct AdamTable {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  int n[kAdamMaxTensors];
};

__global__ void counter_add_kernel(long long* counter, long long inc) { *counter += inc; }

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamTable t, float step_size, float inv_sqrt_bc2, float b1, float b2,
                  float omb1, float omb2, float eps, float wd, float gscale,   // omb = 1 - beta (from double)
                  const long long* __restrict__ step_dev, double lr, double beta1, double beta2) {
  if (step_dev) {
    // step count kept on the device (CUDA-graph replays advance it without the host): the bias
    // corrections are formed here, in double like torch does on the host
    const double st = static_cast<double>(*step_dev);
    step_size = static_cast<float>(lr / (1.0 - pow(beta1, st)));
    inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(1.0 - pow(beta2, st)));
  }
  const int k = blockIdx.y;
  float* __restrict__ p = t.p[k];
  const float* __restrict__ g = t.g[k];
  float* __restrict__ m = t.m[k];
  float* __restrict__ v = t.v[k];
  const int n = t.n[k];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gr = fmaf(wd, pv, g[i] * gscale);
    const float mn = fmaf(b1, m[i], omb1 * gr);
    const float vn = fmaf(b2, v[i], omb2 * gr * gr);
    m[i] = mn;
    v[i] = vn;
    p[i] = pv - step_size * mn / (sqrtf(vn) * inv_sqrt_bc2 + eps);
  }
}

int adam_multi(float* const* params, const float* const* grads, float* const* exp_avgs,
               float* const* exp_avg_sqs, const long long* sizes, int n_tensors, double lr, double beta1,
               double beta2, double eps, double weight_decay, long long step, const long long* step_dev,
               float grad_scale, cudaStream_t stream) {
  UGP_REQUIRE(params && grads && exp_avgs && exp_avg_sqs && sizes && n_tensors > 0 && (step >= 1 || step_dev),
              "adam_multi: bad arguments");
  if (step < 1) step = 1;
  // hyper-parameters arrive as the optimiser's Python doubles: 1 - beta is formed in double like torch does
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  const float step_size = static_cast<float>(lr / bc1);
  const float inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  for (int base = 0; base < n_tensors; base += kAdamMaxTensors) {
    const int cnt = (n_tensors - base < kAdamMaxTensors) ? (n_tensors - base) : kAdamMaxTensors;
    AdamTable t;
    long long biggest = 0;
    for (int i = 0; i < cnt; ++i) {
      UGP_REQUIRE(sizes[base + i] > 0 && sizes[base + i] < (1LL << 31), "adam_multi: bad size");
      t.p[i] = params[base + i];
      t.g[i] = grads[base + i];
      t.m[i] = exp_avgs[base + i];
      t.v[i] = exp_avg_sqs[base + i];
      t.n[i] = static_cast<int>(sizes[base + i]);
      if (sizes[base + i] > biggest) biggest = sizes[base + i];
    }
    int gx = static_cast<int>((biggest + 256 * 8 - 1) / (256 * 8));
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    dim3 grid(gx, cnt);
    adam_multi_kernel<<<grid, 256, 0, stream>>>(t, step_size, inv_sqrt_bc2, static_cast<float>(beta1),
                                                static_cast<float>(beta2), static_cast<float>(1.0 - beta1),
                                                static_cast<float>(1.0 - beta2), static_cast<float>(eps),
                                                static_cast<float>(weight_decay), grad_scale, step_dev, lr, beta1,
                                                beta2);
    UGP_CHECK_CUDA(cudaGetLastError());
  }
  return UGP_OK;
}

int counter_add(long long* counter, long long inc, cudaStream_t stream) {
  UGP_REQUIRE(counter, "counter_add: null pointer");
  counter_add_kernel<<<1, 1, 0, stream>>>(counter, inc);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
