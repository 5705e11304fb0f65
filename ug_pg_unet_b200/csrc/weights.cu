// Weight (un)packing between the fp32 OIHW master layout of nn.Conv2d
// (UG_unet_parts.py:10,13; state_dict compatible) and the bf16 GEMM operands:
//   fwd   pack: Wf[tap][co][ci]                      (tap = r*3+s), K-major B operand
//   dgrad pack: Wd[tap][ci][co] = W[co][ci][2-r][2-s]
//   first layer (9*Cin <= 64): Wf[co][k], k = (r*3+s)*Cin + ci, zero padded to 64
// and back for the packed fp32 weight gradient dWp[tap][ci][co] -> OIHW.
// Small tiles (8 co x 32 ci x 9 taps) through shared memory: the fp32 side is read / written as
// 288-float runs, and even the 64 x 64 layers get 16 blocks (these kernels are latency-, not
// bandwidth-bound: 53.6 MB of parameters in total).
#include "ugp_common.cuh"
#include "ugp_vec.cuh"

namespace ugp {

constexpr int kWCo = 8;                  // output channels per block
constexpr int kWCi = 32;                 // input channels per block
constexpr int kWRow = kWCi * 9 + 1;      // padded smem row (conflict-free transposed reads)
constexpr int kUCo = 32;                 // unpack: output channels per block (contiguous in dWp)
constexpr int kUCi = 8;                  // unpack: input channels per block

__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin,
                    __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  __shared__ float tile[kWCo][kWRow];
  const int co0 = blockIdx.y * kWCo, ci0 = blockIdx.x * kWCi;
  for (int i = threadIdx.x; i < kWCo * kWCi * 9; i += blockDim.x) {
    const int r = i / (kWCi * 9), c = i - r * (kWCi * 9);
    tile[r][c] = w[(static_cast<long long>(co0 + r) * Cin + ci0) * 9 + c];
  }
  __syncthreads();
  // forward pack: rows (tap, co), 32 ci contiguous
  for (int i = threadIdx.x; i < 9 * kWCo * (kWCi / 2); i += blockDim.x) {
    const int cp = i % (kWCi / 2);
    const int r = (i / (kWCi / 2)) % kWCo;
    const int tap = i / ((kWCi / 2) * kWCo);
    const uint32_t v = pack_bf16x2(tile[r][(2 * cp) * 9 + tap], tile[r][(2 * cp + 1) * 9 + tap]);
    *reinterpret_cast<uint32_t*>(wf + (static_cast<long long>(tap) * Cout + co0 + r) * Cin + ci0 + 2 * cp) = v;
  }
  if (wd) {
    // dgrad pack: rows (8 - tap, ci), 8 co contiguous
    for (int i = threadIdx.x; i < 9 * kWCi * (kWCo / 2); i += blockDim.x) {
      const int rp = i % (kWCo / 2);
      const int c = (i / (kWCo / 2)) % kWCi;
      const int tap = i / ((kWCo / 2) * kWCi);
      const uint32_t v = pack_bf16x2(tile[2 * rp][c * 9 + tap], tile[2 * rp + 1][c * 9 + tap]);
      *reinterpret_cast<uint32_t*>(wd + (static_cast<long long>(8 - tap) * Cin + ci0 + c) * Cout + co0 + 2 * rp) = v;
    }
  }
}

__global__ void pack_weights_first_kernel(const float* __restrict__ w, int Cout, int Cin,
                                          __nv_bfloat16* __restrict__ wf) {
  const int total = Cout * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
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

// dWp[tap][ci][co] (co contiguous) -> grad[co][ci][tap]; dWp is zeroed after it is read so the
// next step's wgrad can accumulate into it again without a separate fill.
__global__ void __launch_bounds__(256)
unpack_wgrad_kernel(float* __restrict__ dwp, int Cout, int Cin, float* __restrict__ grad,
                    int accumulate) {
  // tile of 32 co x 8 ci x 9 taps: 128-byte runs along co on the packed side, 288-byte runs along
  // (ci, tap) on the OIHW side
  __shared__ float tile[kUCo][kUCi * 9 + 1];   // [co][ci*9 + tap]
  const int co0 = blockIdx.y * kUCo, ci0 = blockIdx.x * kUCi;
  for (int i = threadIdx.x; i < 9 * kUCi * kUCo; i += blockDim.x) {
    const int r = i % kUCo;                     // co (contiguous in dWp)
    const int c = (i / kUCo) % kUCi;            // ci
    const int tap = i / (kUCo * kUCi);
    float* src = dwp + (static_cast<long long>(tap) * Cin + ci0 + c) * Cout + co0 + r;
    tile[r][c * 9 + tap] = *src;
    *src = 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kUCo * kUCi * 9; i += blockDim.x) {
    const int r = i / (kUCi * 9), c = i - r * (kUCi * 9);
    float* dst = grad + (static_cast<long long>(co0 + r) * Cin + ci0) * 9 + c;
    *dst = accumulate ? *dst + tile[r][c] : tile[r][c];
  }
}

// first layer: dWp[k][co] with k = tap*Cin + ci -> grad[co][ci][tap]
__global__ void unpack_wgrad_first_kernel(float* __restrict__ dwp, int Cout, int Cin,
                                          float* __restrict__ grad, int accumulate) {
  const int total = Cout * Cin * 9;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int tap = idx % 9;
    const int ci = (idx / 9) % Cin;
    const int co = idx / (9 * Cin);
    float* src = dwp + (tap * Cin + ci) * Cout + co;
    const float v = *src;
    *src = 0.f;
    grad[idx] = accumulate ? grad[idx] + v : v;
  }
}

int pack_weights(const float* w, int Cout, int Cin, void* wf, void* wd, cudaStream_t stream) {
  UGP_REQUIRE(w && wf, "pack_weights: null pointer");
  if (Cin % 64 == 0) {
    UGP_REQUIRE(Cout % kWCo == 0, "pack_weights: Cout %d must be a multiple of %d", Cout, kWCo);
    dim3 grid(Cin / kWCi, Cout / kWCo);
    pack_weights_kernel<<<grid, 256, 0, stream>>>(w, Cout, Cin, static_cast<__nv_bfloat16*>(wf),
                                                  static_cast<__nv_bfloat16*>(wd));
  } else {
    UGP_REQUIRE(9 * Cin <= 64, "pack_weights: Cin %d unsupported", Cin);
    pack_weights_first_kernel<<<grid_for(Cout * 64, 256), 256, 0, stream>>>(
        w, Cout, Cin, static_cast<__nv_bfloat16*>(wf));
  }
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int unpack_wgrad(float* dwp, int Cout, int Cin, float* grad, int accumulate, cudaStream_t stream) {
  UGP_REQUIRE(dwp && grad, "unpack_wgrad: null pointer");
  if (Cin % 64 == 0) {
    UGP_REQUIRE(Cout % kUCo == 0, "unpack_wgrad: Cout %d must be a multiple of %d", Cout, kUCo);
    dim3 grid(Cin / kUCi, Cout / kUCo);
    unpack_wgrad_kernel<<<grid, 256, 0, stream>>>(dwp, Cout, Cin, grad, accumulate);
  } else {
    unpack_wgrad_first_kernel<<<grid_for(static_cast<long long>(Cout) * Cin * 9, 256), 256, 0, stream>>>(
        dwp, Cout, Cin, grad, accumulate);
  }
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
