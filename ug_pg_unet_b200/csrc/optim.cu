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
