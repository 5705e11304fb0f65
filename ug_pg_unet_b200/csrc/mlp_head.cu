// Herlev classification head (reference Herlev/train_herlev.py:66-77):
//   AdaptiveAvgPool2d(1) -> Flatten -> Dropout(.5) -> Linear(C,H1) -> ReLU -> Dropout(.3)
//   -> Linear(H1,H2) -> ReLU -> Dropout(.2) -> Linear(H2,K)
// as ONE forward kernel (global average pool of the NHWC bf16 feature map + the three Linear layers)
// and ONE backward kernel for the MLP (weight / bias gradients of all three layers + the gradient of the
// pooled vector); the broadcast of that gradient over the feature map is ugp_global_avg_pool_bwd.
// The head is M = B <= 64 rows against 0.4 M weights: a GEMM library call per layer is all launch
// latency, so the step keeps no cuBLAS call.
//
// Forward: grid = B x C/64 blocks; each block pools one 64-channel slab of one image (fixed-order tree,
// no atomics on the values), then takes a ticket; the LAST block of an image to finish runs that image's
// MLP (warp per output row, lanes stride the row with float4 loads, shuffle reduction).
// Backward: one thread-block cluster of 8 CTAs; three phases (layer 3, 2, 1) separated by cluster
// barriers, intermediate dh2 / dh1 in a caller-provided scratch buffer.
//
// Dropout (train mode): the caller passes U[B, C+H1+H2] uniform randoms (one torch.rand launch); element
// kept iff u >= p, scaled by 1/(1-p) - the same Bernoulli(1-p) mask distribution as nn.Dropout.  Null = eval.
#include "ugp_common.cuh"
#include "ugp_vec.cuh"

namespace ugp {

struct MlpHeadParams {
  int B, HW, C, H1, H2, K;
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  const float* u;          // [B, C+H1+H2] uniforms or null
  float p0, p1, p2;        // dropout probabilities in front of the three Linear layers
};

__device__ __forceinline__ float drop_scale(const float* u, int idx, float p) {
  if (u == nullptr || p <= 0.f) return 1.f;
  return (__ldg(u + idx) >= p) ? 1.f / (1.f - p) : 0.f;
}

// y[n] = act(b[n] + sum_k W[n][k] * x[k]) for n in [0, N): warp per row, x in shared memory (K % 4 == 0).
__device__ __forceinline__ void dense_rows(const float* __restrict__ W, const float* __restrict__ bias,
                                           const float* x_s, int N, int K, bool relu, float* y_s,
                                           float* y_g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int n = warp; n < N; n += nwarps) {
    const float4* row = reinterpret_cast<const float4*>(W + static_cast<long long>(n) * K);
    float acc = 0.f;
    for (int k4 = lane; k4 < (K >> 2); k4 += 32) {
      const float4 w = __ldg(row + k4);
      const float4 xv = *reinterpret_cast<const float4*>(x_s + 4 * k4);
      acc = fmaf(w.x, xv.x, acc); acc = fmaf(w.y, xv.y, acc);
      acc = fmaf(w.z, xv.z, acc); acc = fmaf(w.w, xv.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float v = acc + __ldg(bias + n);
      if (relu) v = fmaxf(v, 0.f);
      if (y_s) y_s[n] = v;
      if (y_g) y_g[n] = v;
    }
  }
}

__global__ void __launch_bounds__(256)
gap_mlp_head_fwd_kernel(const __nv_bfloat16* __restrict__ x, const MlpHeadParams p, float* __restrict__ pooled,
                        float* __restrict__ h1, float* __restrict__ h2, float* __restrict__ logits,
                        unsigned int* __restrict__ tickets) {
  extern __shared__ float smem[];          // max(32*65, C + H1 + H2) floats
  __shared__ unsigned int s_ticket;
  const int slabs = p.C >> 6;
  const int b = blockIdx.x / slabs;
  const int c0 = (blockIdx.x - b * slabs) * 64;
  {
    // ---- global average pool of this 64-channel slab (as global_avg_pool_kernel) ----
    float (*red)[65] = reinterpret_cast<float (*)[65]>(smem);
    const int cgi = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const __nv_bfloat16* base = x + static_cast<long long>(b) * p.HW * p.C + c0 + cgi * 8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int px = pl; px < p.HW; px += 32) {
      const Vec8 v = load8(base + static_cast<long long>(px) * p.C);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v.v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[pl][cgi * 8 + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 64) {
      float s = 0.f;
      for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
      pooled[static_cast<long long>(b) * p.C + c0 + threadIdx.x] = s / static_cast<float>(p.HW);
    }
  }
  // ---- the last slab of image b to arrive runs the MLP of image b ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[b], 1u);
  __syncthreads();
  if (s_ticket != static_cast<unsigned int>(slabs - 1)) return;
  __threadfence();
  if (threadIdx.x == 0) tickets[b] = 0u;       // ready for the next launch (CUDA-graph replays)
  float* a0 = smem;                // [C]  pooled * dropout
  float* a1 = a0 + p.C;            // [H1] relu(fc1) * dropout
  float* a2 = a1 + p.H1;           // [H2] relu(fc2) * dropout
  const int uw = p.C + p.H1 + p.H2;
  const float* ub = p.u ? p.u + static_cast<long long>(b) * uw : nullptr;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x)
    a0[c] = __ldcg(pooled + static_cast<long long>(b) * p.C + c) * drop_scale(ub, c, p.p0);
  __syncthreads();
  dense_rows(p.w1, p.b1, a0, p.H1, p.C, true, a1, h1 + static_cast<long long>(b) * p.H1);
  __syncthreads();
  for (int i = threadIdx.x; i < p.H1; i += blockDim.x) a1[i] *= drop_scale(ub, p.C + i, p.p1);
  __syncthreads();
  dense_rows(p.w2, p.b2, a1, p.H2, p.H1, true, a2, h2 + static_cast<long long>(b) * p.H2);
  __syncthreads();
  for (int i = threadIdx.x; i < p.H2; i += blockDim.x) a2[i] *= drop_scale(ub, p.C + p.H1 + i, p.p2);
  __syncthreads();
  dense_rows(p.w3, p.b3, a2, p.K, p.H2, false, nullptr, logits + static_cast<long long>(b) * p.K);
}

int gap_mlp_head_fwd(const void* x, int B, int HW, int C, const float* w1, const float* b1, int H1,
                     const float* w2, const float* b2, int H2, const float* w3, const float* b3, int K,
                     const float* u, float p0, float p1, float p2, float* pooled, float* h1, float* h2,
                     float* logits, unsigned int* tickets, cudaStream_t stream) {
  UGP_REQUIRE(x && w1 && b1 && w2 && b2 && w3 && b3 && pooled && h1 && h2 && logits && tickets,
              "gap_mlp_head_fwd: null pointer");
  UGP_REQUIRE(B > 0 && HW > 0 && K > 0, "gap_mlp_head_fwd: bad sizes");
  UGP_REQUIRE(C % 64 == 0 && H1 % 4 == 0 && H2 % 4 == 0, "gap_mlp_head_fwd: C %% 64, H1 %% 4, H2 %% 4 required (got %d, %d, %d)", C, H1, H2);
  UGP_REQUIRE(p0 < 1.f && p1 < 1.f && p2 < 1.f && p0 >= 0.f && p1 >= 0.f && p2 >= 0.f, "gap_mlp_head_fwd: dropout p must be in [0, 1)");
  MlpHeadParams p{B, HW, C, H1, H2, K, w1, b1, w2, b2, w3, b3, u, p0, p1, p2};
  int floats = C + H1 + H2;
  if (floats < 32 * 65) floats = 32 * 65;
  const int smem = floats * 4;
  UGP_REQUIRE(smem <= 200 * 1024, "gap_mlp_head_fwd: layer widths too large");
  if (smem > 48 * 1024)
    UGP_CHECK_CUDA(cudaFuncSetAttribute(gap_mlp_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  gap_mlp_head_fwd_kernel<<<B * (C / 64), 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(x), p, pooled, h1,
                                                               h2, logits, tickets);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

// ------------------------------------------------------------------------------------------------ backward
struct MlpHeadBwdParams {
  MlpHeadParams f;
  const float *dlogits, *pooled, *h1, *h2;      // [B,K], [B,C], [B,H1], [B,H2] (h = post-ReLU, pre-dropout)
  float *dw1, *db1, *dw2, *db2, *dw3, *db3;     // overwritten
  float* dpooled;                               // [B,C]
  float* scratch;                               // [B, H1 + H2]: dh1, dh2
};

// One Linear layer's backward, executed by all T threads of the cluster (linear id `tid`):
//   dW[n][k] = sum_b dout[b][n] * a_in[b][k],  db[n] = sum_b dout[b][n],
//   din[b][k] = (sum_n dout[b][n] * W[n][k]) * dscale(b, k) * (relu_in ? h_in[b][k] > 0 : 1)
// with a_in[b][k] = h_in[b][k] * dscale(b, k).  dout [B][N] sits in shared memory (every CTA holds all
// of it); threads are laid out so that consecutive threads touch consecutive k (coalesced W / dW rows).
__device__ __forceinline__ void linear_bwd(const float* dout_s, const float* __restrict__ W,
                                           const float* __restrict__ h_in, const float* u, int u_off, int u_w,
                                           float pdrop, bool relu_in, int B, int N, int Kin, float* __restrict__ dW,
                                           float* __restrict__ db, float* __restrict__ din, int tid, int T) {
  const int parts = T / Kin > 0 ? T / Kin : 1;          // threads per input column
  const int k = tid % Kin, part = tid / Kin;
  if (tid < Kin * parts) {
    // weight gradient: this thread owns column k of rows n = part, part + parts, ...; the column of the
    // layer input is held in registers, 16 samples at a time
    for (int b0 = 0; b0 < B; b0 += 16) {
      float a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int b = b0 + j;
        a[j] = (b < B) ? __ldg(h_in + static_cast<long long>(b) * Kin + k) *
                             drop_scale(u ? u + static_cast<long long>(b) * u_w : nullptr, u_off + k, pdrop)
                       : 0.f;
      }
      const int nb = (B - b0 < 16) ? B - b0 : 16;
      for (int n = part; n < N; n += parts) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nb) acc = fmaf(dout_s[(b0 + j) * N + n], a[j], acc);
        float* dst = dW + static_cast<long long>(n) * Kin + k;
        *dst = (b0 == 0) ? acc : *dst + acc;
      }
    }
    // input gradient: this thread owns column k of samples b = part, part + parts, ...
    if (din != nullptr) {
      for (int b = part; b < B; b += parts) {
        float acc = 0.f;
        for (int n = 0; n < N; ++n) acc = fmaf(dout_s[b * N + n], __ldg(W + static_cast<long long>(n) * Kin + k), acc);
        float g = acc * drop_scale(u ? u + static_cast<long long>(b) * u_w : nullptr, u_off + k, pdrop);
        if (relu_in && !(__ldg(h_in + static_cast<long long>(b) * Kin + k) > 0.f)) g = 0.f;
        din[static_cast<long long>(b) * Kin + k] = g;
      }
    }
  }
  for (int n = tid; n < N; n += T) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dout_s[b * N + n];
    db[n] = s;
  }
}

constexpr int kHeadCluster = 8;

__global__ void __cluster_dims__(kHeadCluster, 1, 1) __launch_bounds__(256)
mlp_head_bwd_kernel(const MlpHeadBwdParams q) {
  extern __shared__ float dout_s[];     // B * max(K, H2, H1) floats
  const MlpHeadParams& p = q.f;
  const int T = kHeadCluster * blockDim.x;
  const int tid = static_cast<int>(cluster_ctarank()) * blockDim.x + threadIdx.x;
  const int uw = p.C + p.H1 + p.H2;
  float* dh1 = q.scratch;
  float* dh2 = q.scratch + static_cast<long long>(p.B) * p.H1;
  // ---- layer 3: dout = dlogits ----
  for (int i = threadIdx.x; i < p.B * p.K; i += blockDim.x) dout_s[i] = __ldg(q.dlogits + i);
  __syncthreads();
  linear_bwd(dout_s, p.w3, q.h2, p.u, p.C + p.H1, uw, p.p2, true, p.B, p.K, p.H2, q.dw3, q.db3, dh2, tid, T);
  __threadfence();
  cluster_sync_all();
  // ---- layer 2: dout = dh2 ----
  for (int i = threadIdx.x; i < p.B * p.H2; i += blockDim.x) dout_s[i] = __ldcg(dh2 + i);
  __syncthreads();
  linear_bwd(dout_s, p.w2, q.h1, p.u, p.C, uw, p.p1, true, p.B, p.H2, p.H1, q.dw2, q.db2, dh1, tid, T);
  __threadfence();
  cluster_sync_all();
  // ---- layer 1: dout = dh1; the input is the pooled vector (no ReLU in front of it) ----
  for (int i = threadIdx.x; i < p.B * p.H1; i += blockDim.x) dout_s[i] = __ldcg(dh1 + i);
  __syncthreads();
  linear_bwd(dout_s, p.w1, q.pooled, p.u, 0, uw, p.p0, false, p.B, p.H1, p.C, q.dw1, q.db1, q.dpooled, tid, T);
}

int mlp_head_bwd(const float* dlogits, int B, int C, const float* w1, int H1, const float* w2, int H2,
                 const float* w3, int K, const float* u, float p0, float p1, float p2, const float* pooled,
                 const float* h1, const float* h2, float* dw1, float* db1, float* dw2, float* db2, float* dw3,
                 float* db3, float* dpooled, float* scratch, cudaStream_t stream) {
  UGP_REQUIRE(dlogits && w1 && w2 && w3 && pooled && h1 && h2 && dw1 && db1 && dw2 && db2 && dw3 && db3 &&
                  dpooled && scratch, "mlp_head_bwd: null pointer");
  UGP_REQUIRE(B > 0 && B <= 64 && C > 0 && H1 > 0 && H2 > 0 && K > 0, "mlp_head_bwd: bad sizes (B <= 64)");
  int widest = K > H2 ? K : H2;
  if (H1 > widest) widest = H1;
  const int smem = B * widest * 4;
  UGP_REQUIRE(smem <= 200 * 1024, "mlp_head_bwd: B * max(H1, H2, K) too large for shared memory");
  MlpHeadBwdParams q;
  q.f = MlpHeadParams{B, 0, C, H1, H2, K, w1, nullptr, w2, nullptr, w3, nullptr, u, p0, p1, p2};
  q.dlogits = dlogits; q.pooled = pooled; q.h1 = h1; q.h2 = h2;
  q.dw1 = dw1; q.db1 = db1; q.dw2 = dw2; q.db2 = db2; q.dw3 = dw3; q.db3 = db3;
  q.dpooled = dpooled; q.scratch = scratch;
  if (smem > 48 * 1024)
    UGP_CHECK_CUDA(cudaFuncSetAttribute(mlp_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mlp_head_bwd_kernel<<<kHeadCluster, 256, smem, stream>>>(q);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
