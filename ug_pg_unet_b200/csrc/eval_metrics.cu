// Evaluation-side reductions: what the reference's evaluators compute per image
// on the host with numpy after a `.cpu()` of every prediction
// (MoNuSegImprove/test_monuseg.py:186-201, 244-258, 264-297; Herlev/test_herlev.py:128-133, 196-204)
// done in one pass over the logits on the device.  HBM-bound: 4 B (+1 B target) read per pixel,
// 1 B written when the mask is requested.
#include "ugp_common.cuh"

namespace ugp {

// Segmentation: pred = sigmoid(z) > 0.5  <=>  z > 0 (test_monuseg.py:188-189, 246-247).
// out[b][4] (double) = {TP = sum pred*gt, sum pred, sum gt, sum sigmoid(z)}; every metric of
// calculate_metrics (test_monuseg.py:264-297) and predict_image's confidence follow from these and the
// pixel count.  mask (nullable): uint8 0/1 prediction.  gt nullable (prediction only).
template <typename T>
__global__ void __launch_bounds__(256)
seg_confusion_kernel(const float* __restrict__ z, const T* __restrict__ gt, long long per_image,
                     unsigned char* __restrict__ mask, double* __restrict__ out) {
  const int b = blockIdx.y;
  const float* zb = z + b * per_image;
  const T* gb = gt ? gt + b * per_image : nullptr;
  unsigned char* mb = mask ? mask + b * per_image : nullptr;
  float s_tp = 0.f, s_p = 0.f, s_g = 0.f, s_sig = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per_image;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float zz = zb[i];
    const float pred = zz > 0.f ? 1.f : 0.f;
    const float g = gb ? static_cast<float>(gb[i]) : 0.f;
    s_tp = fmaf(pred, g, s_tp);
    s_p += pred;
    s_g += g;
    s_sig += 1.f / (1.f + expf(-zz));
    if (mb) mb[i] = zz > 0.f ? 1 : 0;
  }
  float vals[4] = {s_tp, s_p, s_g, s_sig};
  __shared__ float red[4][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float r = warp_sum(vals[k]);
    if (lane == 0) red[k][wid] = r;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float r = 0.f;
    for (int i = 0; i < 8; ++i) r += red[threadIdx.x][i];
    atomicAdd(out + 4 * b + threadIdx.x, static_cast<double>(r));
  }
}

template <typename T>
static int seg_confusion_t(const float* z, const T* gt, int B, long long per_image, unsigned char* mask,
                           double* out, cudaStream_t stream) {
  UGP_REQUIRE(z && out, "seg_confusion: null pointer");
  UGP_REQUIRE(B >= 1 && B <= 65535 && per_image > 0, "seg_confusion: bad shape");
  UGP_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 4 * B, stream));
  long long bx = (per_image + 256 * 4 - 1) / (256 * 4);
  const long long cap = (static_cast<long long>(device_sm_count()) * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<unsigned>(bx), B);
  seg_confusion_kernel<T><<<grid, 256, 0, stream>>>(z, gt, per_image, mask, out);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int seg_confusion(const float* z, const void* gt, int gt_is_u8, int B, long long per_image,
                  unsigned char* mask, double* out, cudaStream_t stream) {
  if (gt_is_u8)
    return seg_confusion_t<unsigned char>(z, static_cast<const unsigned char*>(gt), B, per_image, mask, out,
                                          stream);
  return seg_confusion_t<float>(z, static_cast<const float*>(gt), B, per_image, mask, out, stream);
}

// Classification: probs = softmax(z), pred = argmax (first maximum), and when labels are given the
// K x K confusion matrix counts[label][pred] (Herlev/test_herlev.py:196-204 + sklearn
// confusion_matrix at :232).  One warp per sample.
__global__ void cls_eval_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B,
                                int K, float* __restrict__ probs, long long* __restrict__ preds,
                                unsigned long long* __restrict__ confusion) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* z = logits + static_cast<long long>(warp) * K;
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
  if (probs)
    for (int k = lane; k < K; k += 32) probs[static_cast<long long>(warp) * K + k] = expf(z[k] - mx) / se;
  if (lane == 0) {
    if (preds) preds[warp] = am;
    if (labels && confusion) {
      const long long y = labels[warp];
      if (y >= 0 && y < K) atomicAdd(confusion + y * K + am, 1ull);
    }
  }
}

int cls_eval(const float* logits, const long long* labels, int B, int K, float* probs, long long* preds,
             long long* confusion, cudaStream_t stream) {
  UGP_REQUIRE(logits && B >= 1 && K >= 2, "cls_eval: bad arguments");
  cls_eval_kernel<<<(B * 32 + 127) / 128, 128, 0, stream>>>(
      logits, labels, B, K, probs, preds, reinterpret_cast<unsigned long long*>(confusion));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
