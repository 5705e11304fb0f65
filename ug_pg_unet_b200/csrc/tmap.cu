// Host-side TMA descriptor construction and error plumbing.
//
// cuTensorMapEncodeTiled is resolved through the runtime
// (cudaGetDriverEntryPoint) so the library does not link libcuda directly.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "ugp_common.cuh"

namespace ugp {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_last_error; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t,
                                  void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encoder() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault,
                                &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

static int encode(CUtensorMap* out, const void* ptr, int rank,
                  const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
  EncodeTiledFn fn = resolve_encoder();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return UGP_ERR_DRIVER;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                  const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)",
                   (int)r, rank);
    return UGP_ERR_DRIVER;
  }
  return UGP_OK;
}

// NHWC bf16 activation viewed as a 4-D tensor (C, W, H, B); C innermost.
int make_tmap_nhwc_bf16(CUtensorMap* out, const void* ptr, int B, int H, int W,
                        int C, int boxC, int boxW, int boxH) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H,
                        (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2,
                           (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  return encode(out, ptr, 4, dims, strides, box);
}

// Strided NHWC view (C, W, H, B) with explicit pixel / row / image strides in ELEMENTS: one output-row
// phase of a ConvTranspose2d(k=2, s=2) result, where a "pixel" is the (j, co) pair of 2*Cout channels
// and consecutive rows of the view are two image rows apart.  Strides must be multiples of 8 elements.
int make_tmap_nhwc_view_bf16(CUtensorMap* out, const void* ptr, int B, int H, int W, int C,
                             long long pix_stride, long long row_stride, long long img_stride,
                             int boxC, int boxW, int boxH) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)pix_stride * 2, (cuuint64_t)row_stride * 2,
                           (cuuint64_t)img_stride * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  return encode(out, ptr, 4, dims, strides, box);
}

// fp32 NCHW image viewed as a 4-D tensor (W, H, C, B), W innermost, NO swizzle: a (boxW x boxH x boxC) patch
// lands in shared memory as plain [c][row][x] floats; coordinates outside the image read as zero (= the
// convolution's zero padding).  W must be a multiple of 4 and the base 16-byte aligned (TMA stride rules).
int make_tmap_nchw_f32(CUtensorMap* out, const void* ptr, int B, int C, int H, int W, int boxW, int boxH, int boxC) {
  EncodeTiledFn fn = resolve_encoder();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return UGP_ERR_DRIVER;
  }
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
  cuuint32_t box[4] = {(cuuint32_t)boxW, (cuuint32_t)boxH, (cuuint32_t)boxC, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (fp32 image) failed with CUresult %d", (int)r);
    return UGP_ERR_DRIVER;
  }
  return UGP_OK;
}

// Row-major bf16 matrix [rows][cols]; cols innermost.
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, long long rows,
                      int cols, int boxCols, int boxRows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)boxCols, (cuuint32_t)boxRows};
  return encode(out, ptr, 2, dims, strides, box);
}

int g_conv_halo_mode = 1;
int g_wgrad_halo_mode = 1;
extern int g_conv_pair_mode;   // conv_pair.cu
extern int g_conv_halo_resident;   // conv_halo.cu
extern int g_conv_halo_prefetch;
extern int g_bn_pool_vec4;         // bn_bwd.cu
extern int g_bn_skip_reduce;
extern int g_bn_min_pixels;
extern int g_conv_small_kn;
extern int g_conv_first_fused;
extern int g_conv_first_patch;
// Programmatic dependent launch of the chains conv -> bn_finalize -> BN/ReLU pass -> conv and BN-backward reduce ->
// apply -> dgrad conv (launch_k / pdl_wait, ugp_common.cuh): the next kernel's launch and prologue overlap the tail
// of the previous one.  Same box A/B at batch 16: 5.363 -> 5.329 ms/step (4 pairs); 0 = plain launches.
int g_pdl = 1;
extern int g_upsample_smem_kb;
extern int g_bn_bwd_blocks_per_sm;
int g_wgrad_waves = 1;
int g_wgrad_max_ctas = 0;   // > 0: the split-K grids of the weight-gradient kernels are sized for this many CTAs instead of all SMs
int g_conv_debug = 0;   // what-if timing switches (results are wrong): 1 = skip MMA issue, 2 = skip epilogue body

int set_option(const char* name, int value) {
  if (strcmp(name, "conv_halo") == 0) {
    g_conv_halo_mode = value;
    return UGP_OK;
  }
  if (strcmp(name, "wgrad_halo") == 0) {
    g_wgrad_halo_mode = value;
    return UGP_OK;
  }
  if (strcmp(name, "bn_pool_vec4") == 0) {
    g_bn_pool_vec4 = value;
    return UGP_OK;
  }
  if (strcmp(name, "bn_bwd_blocks_per_sm") == 0) {
    g_bn_bwd_blocks_per_sm = value < 1 ? 1 : (value > 2 ? 2 : value);
    return UGP_OK;
  }
  if (strcmp(name, "upsample_smem_kb") == 0) {
    g_upsample_smem_kb = value < 4 ? 4 : (value > 96 ? 96 : value);
    return UGP_OK;
  }
  if (strcmp(name, "pdl") == 0) {
    g_pdl = value;
    return UGP_OK;
  }
  if (strcmp(name, "conv_first_patch") == 0) {
    g_conv_first_patch = value;
    return UGP_OK;
  }
  if (strcmp(name, "conv_first_fused") == 0) {
    g_conv_first_fused = value;
    return UGP_OK;
  }
  if (strcmp(name, "conv_small_kn") == 0) {
    g_conv_small_kn = value;
    return UGP_OK;
  }
  if (strcmp(name, "bn_min_pixels") == 0) {
    g_bn_min_pixels = value < 1 ? 1 : value;
    return UGP_OK;
  }
  if (strcmp(name, "bn_skip_reduce") == 0) {
    g_bn_skip_reduce = value;
    return UGP_OK;
  }
  if (strcmp(name, "halo_prefetch") == 0) {
    g_conv_halo_prefetch = value;
    return UGP_OK;
  }
  if (strcmp(name, "halo_resident") == 0) {
    g_conv_halo_resident = value;
    return UGP_OK;
  }
  if (strcmp(name, "conv_pair") == 0) {
    g_conv_pair_mode = value;
    return UGP_OK;
  }
  if (strcmp(name, "wgrad_max_ctas") == 0) {
    g_wgrad_max_ctas = value < 0 ? 0 : value;
    return UGP_OK;
  }
  if (strcmp(name, "wgrad_waves") == 0) {
    g_wgrad_waves = value < 1 ? 1 : value;
    return UGP_OK;
  }
  if (strcmp(name, "conv_debug") == 0) {
    g_conv_debug = value;
    return UGP_OK;
  }
  set_last_error("unknown option '%s'", name);
  return UGP_ERR_INVALID;
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) !=
        cudaSuccess)
      sms = 148;
  }
  return sms;
}

}  // namespace ugp
