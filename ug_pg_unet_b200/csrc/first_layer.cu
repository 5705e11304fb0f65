// First-layer input packing (replaces the implicit im2col of the 3-channel
// nn.Conv2d at UG_unet_parts.py:10 inside InConv, and folds in the bilinear
// align_corners=True resize of uncertainty_guided_trainer.py:208 /
// UG_unet.py:36-41): fp32 NCHW image -> bf16 [B,H,W,64] rows holding the 3x3
// neighbourhood, k = (r*3+s)*Cin + ci, zero beyond 9*Cin and outside the image.
// The 64-wide rows then feed the tcgen05 GEMM as a 1x1 convolution.
//
// The same packer also takes the loaders' native format — uint8 HWC image bytes [B,Hin,Win,Cin]
// (np.array(PIL image), MoNuSegImprove/monuseg_dataset.py:137-187, test_monuseg.py:178-181) — with the
// ToTensor /255 and the optional per-channel Normalize(mean, std) (Herlev/herlev_dataset.py:200-228)
// folded in as one fma per sample (v = byte * scale[c] + bias[c]): 4x fewer host->device bytes and no
// separate float image in HBM.
//
// HBM-bound on the 128 B/pixel output.  One thread = one pixel: neighbours are
// read through L1 (adjacent threads share them), the row is written as eight
// 16-byte stores.
#include "first_rows.cuh"

namespace ugp {

template <int kCin, bool kResize, bool kU8>
__global__ void __launch_bounds__(256)
im2col_first_kernel(const void* __restrict__ xin, int B, int Hin, int Win, int H, int W,
                    __nv_bfloat16* __restrict__ out, const ByteNorm norm) {
  __shared__ uint4 stage[256 * 8];
  const long long total = static_cast<long long>(B) * H * W;
  // block-uniform trip count: every thread reaches the barriers, tail pixels are clamped and not stored
  for (long long pix0 = blockIdx.x * static_cast<long long>(blockDim.x); pix0 < total;
       pix0 += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = pix0 + threadIdx.x;
    const long long pc = pix < total ? pix : total - 1;
    const int xo = static_cast<int>(pc % W);
    const int yo = static_cast<int>((pc / W) % H);
    const int b = static_cast<int>(pc / (static_cast<long long>(W) * H));
    uint4 u[8];
    first_layer_row<kCin, kResize, kU8>(xin, Hin, Win, H, W, b, yo, xo, norm, u);
    // Stage the 256 rows of this block through shared memory (16-byte chunks, XOR-swizzled by the row)
    // so that the global stores are fully coalesced: lane l writes chunk l of a contiguous 32 KB run
    // instead of eight 16-byte pieces 128 bytes apart.
    uint4* srow = stage + threadIdx.x * 8;
#pragma unroll
    for (int g = 0; g < 8; ++g) srow[g ^ (threadIdx.x & 7)] = u[g];
    __syncthreads();
    const long long base_pix = pix - threadIdx.x;            // first pixel of this block's trip
    uint4* dst = reinterpret_cast<uint4*>(out + base_pix * 64);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int L = k * 256 + threadIdx.x;                   // linear 16-byte chunk of the 32 KB run
      const int r = L >> 3;
      if (base_pix + r < total) dst[L] = stage[r * 8 + ((L & 7) ^ (r & 7))];
    }
    __syncthreads();
  }
}

template <int kCin, bool kU8>
static int launch_im2col(const void* x, int B, int Hin, int Win, int H, int W, void* out,
                         const ByteNorm& norm, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * H * W;
  const int grid = grid_for(total, 256);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (Hin != H || Win != W)
    im2col_first_kernel<kCin, true, kU8><<<grid, 256, 0, stream>>>(x, B, Hin, Win, H, W, o, norm);
  else
    im2col_first_kernel<kCin, false, kU8><<<grid, 256, 0, stream>>>(x, B, Hin, Win, H, W, o, norm);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

template <bool kU8>
static int dispatch_im2col(const void* x, int B, int Cin, int Hin, int Win, int H, int W, void* out,
                           const ByteNorm& norm, cudaStream_t stream) {
  switch (Cin) {
    case 1: return launch_im2col<1, kU8>(x, B, Hin, Win, H, W, out, norm, stream);
    case 2: return launch_im2col<2, kU8>(x, B, Hin, Win, H, W, out, norm, stream);
    case 3: return launch_im2col<3, kU8>(x, B, Hin, Win, H, W, out, norm, stream);
    case 4: return launch_im2col<4, kU8>(x, B, Hin, Win, H, W, out, norm, stream);
    case 5: return launch_im2col<5, kU8>(x, B, Hin, Win, H, W, out, norm, stream);
    case 6: return launch_im2col<6, kU8>(x, B, Hin, Win, H, W, out, norm, stream);
    case 7: return launch_im2col<7, kU8>(x, B, Hin, Win, H, W, out, norm, stream);
    default:
      UGP_REQUIRE(false, "im2col_first: Cin %d unsupported (1..7: 9*Cin must be <= 64)", Cin);
  }
  return UGP_OK;
}

int make_byte_norm(ByteNorm* out, int Cin, const float* mean, const float* stddev, const char* who) {
  UGP_REQUIRE(Cin >= 1 && Cin <= 7, "%s: Cin %d unsupported (1..7)", who, Cin);
  for (int c = 0; c < Cin; ++c) {   // HOST arrays (nullable): (byte/255 - mean) / std
    const float m = mean ? mean[c] : 0.f, sd = stddev ? stddev[c] : 1.f;
    UGP_REQUIRE(sd > 0.f, "%s: std[%d] must be positive", who, c);
    out->scale[c] = 1.f / (255.f * sd);
    out->bias[c] = -m / sd;
  }
  return UGP_OK;
}

int im2col_first(const float* x, int B, int Cin, int Hin, int Win, int H, int W, void* out,
                 cudaStream_t stream) {
  UGP_REQUIRE(x && out, "im2col_first: null pointer");
  return dispatch_im2col<false>(x, B, Cin, Hin, Win, H, W, out, ByteNorm{}, stream);
}

int im2col_first_u8(const unsigned char* x, int B, int Cin, int Hin, int Win, int H, int W,
                    const float* mean, const float* stddev, void* out, cudaStream_t stream) {
  UGP_REQUIRE(x && out, "im2col_first_u8: null pointer");
  UGP_REQUIRE(Cin >= 1 && Cin <= 7, "im2col_first_u8: Cin %d unsupported (1..7)", Cin);
  ByteNorm norm{};
  if (int rc = make_byte_norm(&norm, Cin, mean, stddev, "im2col_first_u8")) return rc;
  return dispatch_im2col<true>(x, B, Cin, Hin, Win, H, W, out, norm, stream);
}

// ----------------------------------------------------------------------------
// Joint horizontal / vertical flips of a batch of loader bytes on the device: the geometric part of the
// reference's train-time augmentation that is exactly reproducible (MoNuSegImprove/monuseg_dataset.py:
// 150-158: TF.hflip / TF.vflip applied to image AND mask with one decision per sample).  flags[b] bit 0 =
// horizontal, bit 1 = vertical.  x: [B][H][W][C] bytes (C = 3 image bytes, C = 1 a mask plane).  The
// arbitrary-angle PIL rotation and colour jitter of :160-178 stay on the host loader.
// ----------------------------------------------------------------------------
__global__ void flip_u8_kernel(const unsigned char* __restrict__ x, const unsigned char* __restrict__ flags, int H,
                               int W, int C, unsigned char* __restrict__ out) {
  const int b = blockIdx.y;
  const unsigned char f = flags[b];
  const long long plane = static_cast<long long>(H) * W * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < plane;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long pix = i / C;
    const int xx = static_cast<int>(pix % W), yy = static_cast<int>(pix / W);
    const int sx = (f & 1) ? W - 1 - xx : xx, sy = (f & 2) ? H - 1 - yy : yy;
    out[b * plane + i] = x[b * plane + (static_cast<long long>(sy) * W + sx) * C + c];
  }
}

int flip_u8(const unsigned char* x, const unsigned char* flags, int B, int H, int W, int C, unsigned char* out,
            cudaStream_t stream) {
  UGP_REQUIRE(x && flags && out && x != out, "flip_u8: null pointer / in-place not supported");
  UGP_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && C > 0, "flip_u8: bad shape");
  const long long plane = static_cast<long long>(H) * W * C;
  int gx = static_cast<int>((plane + 255) / 256);
  if (gx > 1024) gx = 1024;
  flip_u8_kernel<<<dim3(gx, B), 256, 0, stream>>>(x, flags, H, W, C, out);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
