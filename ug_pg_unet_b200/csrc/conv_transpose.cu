// nn.ConvTranspose2d(Cin, Cout, kernel_size=2, stride=2) on the tcgen05 implicit-GEMM kernels
// (reference: UpSample.up, UG_unet_parts.py:57-67, used by the legacy UNet, UG_unet.py:430-462).
//
//   out[b, 2y+i, 2x+j, co] = bias[co] + sum_ci x[b, y, x, ci] * W[ci, co, i, j]
//
// With stride == kernel size every output pixel has exactly one contributing input pixel and tap, so
// the layer is a 1x1 GEMM  [B*H*W, Cin] x [Cin, (i, j, co)]  followed by a pixel shuffle.  The shuffle
// costs nothing: in the NHWC output [B, 2H, 2W, Cout] the (j, co) pair of one input pixel is a
// contiguous run of 2*Cout elements, consecutive x are 2*Cout apart and consecutive y are 4*W*Cout
// apart, i.e. each output-row phase i is a strided NHWC view with 2*Cout channels
// (make_tmap_nhwc_view_bf16) that the GEMM epilogue writes / the loaders read through TMA:
//   forward  one launch, N = 4*Cout, columns [0, 2*Cout) -> phase 0, the rest -> phase 1
//            (two launches when 4*Cout exceeds the epilogue's channel tables)
//   dgrad    dx = [phase 0 | phase 1] (K = 4*Cout walked as two sources) x Wd
//   wgrad    per phase: dWp[i][ci][(j, co)] += sum_pixels x[p, ci] * dout_i[p, (j, co)]
// Weight packs (bf16): Wf[(i*2+j)*Cout + co][ci]  (K-major B operand of the forward GEMM),
//                      Wd[ci][(i*2+j)*Cout + co]   (K-major B operand of the input-gradient GEMM).
#include "ugp_common.cuh"
#include "ugp_vec.cuh"
#include "conv_gemm.cuh"

namespace ugp {

int conv_gemm_bf16_ex(const void* src0, int c0, const void* src1, int c1, int B, int H, int W, int taps,
                      const void* wpack, int cout, void* dst0, void* dst1, int split_n, const float* bias,
                      const float* scale, const float* shift, int relu, float* stats,
                      const ConvView* src_view, const ConvView* dst_view, int bias_period,
                      cudaStream_t stream, const void* bwd_z = nullptr);
int conv_wgrad_bf16_ex(const void* src0, int c0, const void* src1, int c1, const void* dz, int cout, int B,
                       int H, int W, int taps, float* dw_packed, const ConvView* dz_view,
                       cudaStream_t stream);

static int check_tconv_shape(const char* who, int cin, int cout) {
  UGP_REQUIRE(cin > 0 && cin % 64 == 0 && cin <= kMaxCout, "%s: cin %d unsupported (multiple of 64, <= %d)",
              who, cin, kMaxCout);
  UGP_REQUIRE(cout > 0 && cout % 64 == 0 && 2 * cout <= kMaxCout,
              "%s: cout %d unsupported (multiple of 64, <= %d)", who, cout, kMaxCout / 2);
  return UGP_OK;
}

int conv_transpose2x2_fwd(const void* x, int B, int H, int W, int cin, const void* wf, int cout,
                          const float* bias, void* out, cudaStream_t stream) {
  UGP_REQUIRE(x && wf && out, "conv_transpose2x2_fwd: null pointer");
  int rc;
  if ((rc = check_tconv_shape("conv_transpose2x2_fwd", cin, cout))) return rc;
  const ConvView v{2LL * cout, 4LL * W * cout, 4LL * H * W * cout};
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  const long long phase = 2LL * W * cout;   // elements between output rows 2y and 2y+1
  if (4 * cout <= kMaxCout)
    return conv_gemm_bf16_ex(x, cin, nullptr, 0, B, H, W, 1, wf, 4 * cout, o, o + phase, 2 * cout, bias,
                             nullptr, nullptr, 0, nullptr, nullptr, &v, cout, stream);
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(wf);
  for (int i = 0; i < 2; ++i)
    if ((rc = conv_gemm_bf16_ex(x, cin, nullptr, 0, B, H, W, 1, w + static_cast<long long>(i) * 2 * cout * cin,
                                2 * cout, o + i * phase, nullptr, 2 * cout, bias, nullptr, nullptr, 0,
                                nullptr, nullptr, &v, cout, stream)))
      return rc;
  return UGP_OK;
}

int conv_transpose2x2_dgrad(const void* dout, int B, int H, int W, int cout, const void* wd, int cin,
                            void* dx, cudaStream_t stream) {
  UGP_REQUIRE(dout && wd && dx, "conv_transpose2x2_dgrad: null pointer");
  int rc;
  if ((rc = check_tconv_shape("conv_transpose2x2_dgrad", cin, cout))) return rc;
  const ConvView v{2LL * cout, 4LL * W * cout, 4LL * H * W * cout};
  const __nv_bfloat16* d = static_cast<const __nv_bfloat16*>(dout);
  return conv_gemm_bf16_ex(d, 2 * cout, d + 2LL * W * cout, 2 * cout, B, H, W, 1, wd, cin, dx, nullptr, cin,
                           nullptr, nullptr, nullptr, 0, nullptr, &v, nullptr, 0, stream);
}

int conv_transpose2x2_wgrad(const void* x, const void* dout, int B, int H, int W, int cin, int cout,
                            float* dw_packed, cudaStream_t stream) {
  UGP_REQUIRE(x && dout && dw_packed, "conv_transpose2x2_wgrad: null pointer");
  int rc;
  if ((rc = check_tconv_shape("conv_transpose2x2_wgrad", cin, cout))) return rc;
  const ConvView v{2LL * cout, 4LL * W * cout, 4LL * H * W * cout};
  const __nv_bfloat16* d = static_cast<const __nv_bfloat16*>(dout);
  for (int i = 0; i < 2; ++i)
    if ((rc = conv_wgrad_bf16_ex(x, cin, nullptr, 0, d + i * 2LL * W * cout, 2 * cout, B, H, W, 1,
                                 dw_packed + static_cast<long long>(i) * cin * 2 * cout, &v, stream)))
      return rc;
  return UGP_OK;
}

// ---- weight (un)packing: fp32 [Cin][Cout][2][2] (nn.ConvTranspose2d.weight) <-> GEMM operands -------
constexpr int kTT = 32;   // 32 ci x 32 co tile, all four taps

__global__ void __launch_bounds__(256)
pack_weights_tconv_kernel(const float* __restrict__ w, int Cin, int Cout, __nv_bfloat16* __restrict__ wf,
                          __nv_bfloat16* __restrict__ wd) {
  __shared__ float tile[4][kTT][kTT + 1];   // [tap][ci][co]
  const int ci0 = blockIdx.x * kTT, co0 = blockIdx.y * kTT;
  const float4* w4 = reinterpret_cast<const float4*>(w);
  for (int i = threadIdx.x; i < kTT * kTT; i += blockDim.x) {
    const int c = i / kTT, o = i - c * kTT;
    const float4 v = w4[static_cast<long long>(ci0 + c) * Cout + co0 + o];
    tile[0][c][o] = v.x; tile[1][c][o] = v.y; tile[2][c][o] = v.z; tile[3][c][o] = v.w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * kTT * (kTT / 2); i += blockDim.x) {
    const int hp = i % (kTT / 2);
    const int r = (i / (kTT / 2)) % kTT;
    const int tap = i / ((kTT / 2) * kTT);
    // forward pack: row (tap, co = r), ci pair contiguous
    *reinterpret_cast<uint32_t*>(wf + (static_cast<long long>(tap) * Cout + co0 + r) * Cin + ci0 + 2 * hp) =
        pack_bf16x2(tile[tap][2 * hp][r], tile[tap][2 * hp + 1][r]);
    // input-gradient pack: row ci = r, (tap, co pair) contiguous
    if (wd)
      *reinterpret_cast<uint32_t*>(wd + static_cast<long long>(ci0 + r) * 4 * Cout + tap * Cout + co0 + 2 * hp) =
          pack_bf16x2(tile[tap][r][2 * hp], tile[tap][r][2 * hp + 1]);
  }
}

// dWp[i][ci][j*Cout + co] (fp32, co contiguous) -> grad[ci][co][i][j]; dWp is zeroed after it is read.
__global__ void __launch_bounds__(256)
unpack_wgrad_tconv_kernel(float* __restrict__ dwp, int Cin, int Cout, float* __restrict__ grad,
                          int accumulate) {
  const long long total = static_cast<long long>(Cin) * Cout;
  float4* g4 = reinterpret_cast<float4*>(grad);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(idx / Cout), co = static_cast<int>(idx - static_cast<long long>(ci) * Cout);
    float* p00 = dwp + static_cast<long long>(ci) * 2 * Cout + co;
    float* p10 = p00 + static_cast<long long>(Cin) * 2 * Cout;
    float4 v = make_float4(p00[0], p00[Cout], p10[0], p10[Cout]);
    p00[0] = 0.f; p00[Cout] = 0.f; p10[0] = 0.f; p10[Cout] = 0.f;
    if (accumulate) {
      const float4 o = g4[idx];
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    g4[idx] = v;
  }
}

int pack_weights_tconv(const float* w, int cin, int cout, void* wf, void* wd, cudaStream_t stream) {
  UGP_REQUIRE(w && wf, "pack_weights_tconv: null pointer");
  UGP_REQUIRE(cin % kTT == 0 && cout % kTT == 0, "pack_weights_tconv: channels must be multiples of %d", kTT);
  dim3 grid(cin / kTT, cout / kTT);
  pack_weights_tconv_kernel<<<grid, 256, 0, stream>>>(w, cin, cout, static_cast<__nv_bfloat16*>(wf),
                                                      static_cast<__nv_bfloat16*>(wd));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int unpack_wgrad_tconv(float* dwp, int cin, int cout, float* grad, int accumulate, cudaStream_t stream) {
  UGP_REQUIRE(dwp && grad, "unpack_wgrad_tconv: null pointer");
  unpack_wgrad_tconv_kernel<<<grid_for(static_cast<long long>(cin) * cout, 256), 256, 0, stream>>>(
      dwp, cin, cout, grad, accumulate);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

// ---- per-channel sum of an NHWC bf16 tensor (the bias gradient of a layer with no BatchNorm behind it) --
constexpr int kCsThreads = 256;

__global__ void __launch_bounds__(kCsThreads)
channel_sum_kernel(const __nv_bfloat16* __restrict__ x, long long pixels, int C, float* __restrict__ out) {
  __shared__ float part[kCsThreads][9];
  const int nvec = C / 8;
  const int lanes = kCsThreads / nvec;          // pixel rows handled per pass (>= 1: C <= 2048)
  const int vec = threadIdx.x % nvec, lane = threadIdx.x / nvec;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (lane < lanes) {
    for (long long p = static_cast<long long>(blockIdx.x) * lanes + lane; p < pixels;
         p += static_cast<long long>(gridDim.x) * lanes) {
      const Vec8 v = load8(x + p * C + vec * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] += v.v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) part[threadIdx.x][k] = s[k];
  __syncthreads();
  if (threadIdx.x < nvec) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += part[l * nvec + threadIdx.x][k];
      atomicAdd(out + threadIdx.x * 8 + k, t);
    }
  }
}

int channel_sum(const void* x, long long pixels, int C, float* out, cudaStream_t stream) {
  UGP_REQUIRE(x && out && pixels > 0, "channel_sum: bad arguments");
  UGP_REQUIRE(C % 8 == 0 && C / 8 <= kCsThreads, "channel_sum: C %d unsupported (multiple of 8, <= %d)", C,
              8 * kCsThreads);
  const int lanes = kCsThreads / (C / 8);
  const int grid = grid_for((pixels + lanes - 1) / lanes * kCsThreads, kCsThreads, 4);
  channel_sum_kernel<<<grid, kCsThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), pixels, C, out);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
