// Weight gradient of the 3x3 (pad 1) / 1x1 convolution on tcgen05 tensor cores.
//
// Replaces the autograd weight-gradient of nn.Conv2d (reference call sites
// UG_unet_parts.py:10,13):
//
//   dWp[tap][ci][co] += sum_{pixels p} X[p + shift(tap), ci] * dZ[p, co]
//
// GEMM view: M = (tap, ci), N = co, K = pixels.  Both operands are NHWC bf16
// activations whose contiguous dimension (channels) is the GEMM M / N
// dimension, so both UMMA operands are MN-major: a TMA box of 64 channels x 64
// pixels lands in smem as 64 K-rows of 128 B (SWIZZLE_128B), which is exactly
// the canonical MN-major layout (8-row K groups 1024 B apart = SBO, 64-channel
// blocks one box apart = LBO).
//
// An M tile is two 64-channel boxes; each box is (tap, 64-channel chunk), so
// for Cin = 64 one tile covers two taps.  A CTA owns up to kT M tiles (kT
// accumulators of 128 x kN in TMEM, kT * kN <= 512 columns) that all share the
// dZ tile of each 64-pixel K chunk: the operand bytes pulled from L2 per MMA
// drop by ~kT/2 versus one tile per CTA, which is what bounds this kernel.
// K (pixels) is split across CTAs and partial tiles are reduced with vector
// fp32 atomics into the packed fp32 gradient [tap][Cin][Cout] (Cout contiguous).
#include "ugp_common.cuh"

namespace ugp {

bool pick_box_general(int H, int W, int pixels, int* BW, int* BH);   // conv_gemm.cu
extern int g_conv_debug;
extern int g_wgrad_waves;   // CTA waves the split-K grid is sized for (default 1)
extern int g_wgrad_max_ctas;

struct WgradParams {
  int taps;
  int chunks0, chunks1;   // 64-channel chunks of source 0 / 1
  int H, W, BHp, BWp;     // K box of BWp x BHp <= 64 pixels
  int tiles_x, tiles_y;   // boxes per image row / column
  int box_bytes;          // BWp * BHp * 128: bytes one TMA box delivers (8 KB for a full 64-pixel box)
  int zero_fill;          // box < 64 pixels: the unused K rows of every stage buffer must be zero
  int boxes;              // taps * (chunks0 + chunks1)
  int m_tiles;            // ceil(boxes / 2)
  int m_groups;           // ceil(m_tiles / kT)
  int n_tiles;
  int cout;
  int kchunks_total;      // B * tiles_x * tiles_y
  int kchunks_per_split;
  int stage_tiles;        // M tiles a stage buffer holds = min(kT, m_tiles)
  int stages;             // pipeline depth: as many stages of that size as fit (2..8)
  float* dw;              // [boxes*64][cout] fp32, accumulated
  int debug;              // what-if timing only: 1 = skip MMA issue, 2 = skip the reduction stores
};

constexpr int kWgBoxBytes = 64 * 64 * 2;  // 64 pixels x 64 channels bf16 = 8 KB

template <int kN>
struct WgradCfg {
  static constexpr int kT = 512 / kN > 5 ? 5 : 512 / kN;            // 5 / 4 / 2 accumulators
  static constexpr int kStages = (kN == 256) ? 3 : 2;
  static constexpr int kABytes = 2 * kT * kWgBoxBytes;
  static constexpr int kBBytes = (kN / 64) * kWgBoxBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b,
                                           float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr),
               "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int kN>
__global__ void __launch_bounds__(192, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX0,
                  const __grid_constant__ CUtensorMap tmX1,
                  const __grid_constant__ CUtensorMap tmDZ,
                  const WgradParams p) {
  using Cfg = WgradCfg<kN>;
  constexpr int kT = Cfg::kT;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Stage ring sized from the tiles this launch really uses (p.stage_tiles <= kT M tiles per CTA): a layer
  // with one M tile (the im2col'ed first layer: K = 1 M pixels against a 64 x 64 result) gets 8 stages of
  // 24 KB in flight instead of 2 stages of 88 KB that are mostly empty - it is bound by load latency.
  const int stages = p.stages;
  const int a_bytes = 2 * p.stage_tiles * kWgBoxBytes;
  uint8_t* sA = smem;
  uint8_t* sB = sA + stages * a_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + stages * Cfg::kBBytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* done_bar = empty_bar + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_group = blockIdx.x / p.n_tiles;
  const int n_tile = blockIdx.x - m_group * p.n_tiles;
  const int tile0 = m_group * kT;                         // first M tile of this CTA
  const int nt = (p.m_tiles - tile0 < kT) ? (p.m_tiles - tile0) : kT;   // live M tiles
  const int kc_begin = blockIdx.y * p.kchunks_per_split;
  int kc_end = kc_begin + p.kchunks_per_split;
  if (kc_end > p.kchunks_total) kc_end = p.kchunks_total;
  const int n_k = kc_end - kc_begin;   // host guarantees > 0

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmDZ);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc(tmem_slot, 512);
  if (p.zero_fill) {
    // partial K boxes (general image sizes): TMA only writes the first BWp*BHp rows of each 64-row
    // box, the MMAs read all 64 -> the rest must contribute zero
    uint4* z4 = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < stages * (a_bytes + Cfg::kBBytes) / 16; i += blockDim.x) z4[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform (uniform registers)

  const int chunks = p.chunks0 + p.chunks1;

  if (warp_idx == 0) {
    const bool leader = elect_one() != 0;     // whole-warp loop, elected lane issues (uniform operands)
    {
      const uint32_t stage_bytes = (nt * 2 + kN / 64) * p.box_bytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        const int t2 = kc / p.tiles_x;
        const int x0 = (kc - t2 * p.tiles_x) * p.BWp;
        const int b = t2 / p.tiles_y;
        const int y0 = (t2 - b * p.tiles_y) * p.BHp;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        uint8_t* a_dst = sA + stage * a_bytes;
        for (int i = 0; i < 2 * nt; ++i) {
          // the boxes of this CTA; an odd box count repeats the last box
          int id = tile0 * 2 + i;
          if (id >= p.boxes) id = p.boxes - 1;
          const int tap = id / chunks;
          const int ch = id - tap * chunks;
          const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
          if (leader) {
            if (ch < p.chunks0)
              tma_load_4d(&tmX0, &full_bar[stage], a_dst + i * kWgBoxBytes, ch * 64, x0 + dx, y0 + dy, b);
            else
              tma_load_4d(&tmX1, &full_bar[stage], a_dst + i * kWgBoxBytes, (ch - p.chunks0) * 64,
                          x0 + dx, y0 + dy, b);
          }
        }
        uint8_t* b_dst = sB + stage * Cfg::kBBytes;
        if (leader) {
          for (int j = 0; j < kN / 64; ++j)
            tma_load_4d(&tmDZ, &full_bar[stage], b_dst + j * kWgBoxBytes,
                        n_tile * kN + j * 64, x0, y0, b);
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_idx == 1) {
    // whole-warp issue loop (uniform-datapath operands), one elected lane executes tcgen05 (conv_halo.cu)
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = umma_idesc_bf16(128, kN, 1, 1);
    // MN-major SWIZZLE_128B descriptor template: LBO = one box (next 64 channels), SBO = 1024 B
    const uint64_t desc0 = umma_desc_sw128(0, kWgBoxBytes, 1024);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < n_k; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after_sync();
      {
        const uint32_t a_addr = smem_u32(sA + stage * a_bytes);
        const uint32_t b_addr = smem_u32(sB + stage * Cfg::kBBytes);
        for (int t = 0; t < nt; ++t) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // 16 pixels (K) per MMA = 16 rows of 128 B
            const uint64_t a_desc = desc0 + ((a_addr + t * 2 * kWgBoxBytes + k * 2048) >> 4);
            const uint64_t b_desc = desc0 + ((b_addr + k * 2048) >> 4);
            const uint32_t acc = (i > 0 || k > 0) ? 1u : 0u;
            if (leader && !(p.debug & 1)) umma_bf16(tmem_base + t * kN, a_desc, b_desc, idesc, acc);
          }
        }
        if (leader) {
          umma_commit(&empty_bar[stage]);
          if (i == n_k - 1) umma_commit(done_bar);
        }
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else {
    const int q = warp_idx & 3;
    const int row = q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
    for (int t = 0; t < nt; ++t) {
      const int box_id = (tile0 + t) * 2 + (row >> 6);
      float* dst = p.dw + (static_cast<size_t>(box_id) * 64 + (row & 63)) * p.cout + n_tile * kN;
      const bool live = box_id < p.boxes;
#pragma unroll 1
      for (int c = 0; c < kN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * kN + c * 32, v);
        tmem_ld_wait();
        if (live && !(p.debug & 2)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(dst + c * 32 + j, __uint_as_float(v[j]),
                       __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                       __uint_as_float(v[j + 3]));
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int kN>
static int launch_wgrad(const CUtensorMap& tmX0, const CUtensorMap& tmX1,
                        const CUtensorMap& tmDZ, WgradParams p, cudaStream_t stream) {
  using Cfg = WgradCfg<kN>;
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<kN>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    configured = true;
  }
  p.m_groups = (p.m_tiles + Cfg::kT - 1) / Cfg::kT;
  p.stage_tiles = p.m_tiles < Cfg::kT ? p.m_tiles : Cfg::kT;
  p.stages = (Cfg::kSmemBytes - 256 - 1024) / (2 * p.stage_tiles * kWgBoxBytes + Cfg::kBBytes);
  if (p.stages > 8) p.stages = 8;
  const int items = p.m_groups * p.n_tiles;
  // split K (pixels) so that the grid is ONE full wave of CTAs (one CTA per SM: 512 TMEM columns):
  // more splits only add fp32 reduction traffic and a ragged last wave
  int sms = device_sm_count();
  if (g_wgrad_max_ctas > 0 && g_wgrad_max_ctas < sms) sms = g_wgrad_max_ctas;
  int splits = g_wgrad_waves * sms / items;
  if (splits < 1) splits = 1;
  // keep at least 8 K chunks (512 pixels) per split
  int max_splits = p.kchunks_total / 8;
  if (max_splits < 1) max_splits = 1;
  if (splits > max_splits) splits = max_splits;
  p.kchunks_per_split = (p.kchunks_total + splits - 1) / splits;
  splits = (p.kchunks_total + p.kchunks_per_split - 1) / p.kchunks_per_split;
  dim3 grid(items, splits);
  conv_wgrad_kernel<kN><<<grid, 192, Cfg::kSmemBytes, stream>>>(tmX0, tmX1, tmDZ, p);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

// conv_wgrad_halo.cu
bool conv_wgrad_halo_eligible(int H, int W, int cout, int taps);
int conv_wgrad_halo_bf16(const void* src0, int c0, const void* src1, int c1, const void* dz, int cout,
                         int B, int H, int W, float* dw_packed, cudaStream_t stream);
extern int g_wgrad_halo_mode;

// dz_view (nullable): dz is a strided NHWC view (a ConvTranspose2d output-row phase, conv_transpose.cu).
int conv_wgrad_bf16_ex(const void* src0, int c0, const void* src1, int c1,
                       const void* dz, int cout, int B, int H, int W, int taps,
                       float* dw_packed, const ConvView* dz_view, cudaStream_t stream) {
  UGP_REQUIRE(src0 && dz && dw_packed, "conv_wgrad: null pointer");
  UGP_REQUIRE(taps == 9 || taps == 1, "conv_wgrad: taps must be 9 or 1");
  UGP_REQUIRE(c0 > 0 && c0 % 64 == 0 && c1 >= 0 && c1 % 64 == 0,
              "conv_wgrad: source channels must be multiples of 64");
  UGP_REQUIRE(cout % 64 == 0 && cout <= 1024, "conv_wgrad: cout %d unsupported", cout);
  if (!dz_view && g_wgrad_halo_mode != 0 && conv_wgrad_halo_eligible(H, W, cout, taps))
    return conv_wgrad_halo_bf16(src0, c0, src1, c1, dz, cout, B, H, W, dw_packed, stream);
  int BWp = W < 64 ? W : 64;
  int BHp = 64 / BWp;
  if (!(W % BWp == 0 && BWp * BHp == 64 && H % BHp == 0)) {
    // general image sizes: the largest box <= 64 pixels whose sides divide W and H
    BWp = BHp = 0;
    UGP_REQUIRE(H > 0 && W > 0 && pick_box_general(H, W, 64, &BWp, &BHp),
                "conv_wgrad: unsupported spatial size %dx%d", H, W);
  }
  int kn = 256;
  while (kn > 64 && cout % kn != 0) kn >>= 1;

  WgradParams p;
  p.taps = taps;
  p.chunks0 = c0 / 64;
  p.chunks1 = c1 / 64;
  p.H = H; p.W = W; p.BHp = BHp; p.BWp = BWp;
  p.boxes = taps * (p.chunks0 + p.chunks1);
  p.m_tiles = (p.boxes + 1) / 2;
  p.m_groups = 0;
  p.n_tiles = cout / kn;
  p.cout = cout;
  p.tiles_x = W / BWp;
  p.tiles_y = H / BHp;
  p.box_bytes = BWp * BHp * 128;
  p.zero_fill = BWp * BHp < 64;
  p.kchunks_total = B * p.tiles_x * p.tiles_y;
  p.kchunks_per_split = 0;
  p.dw = dw_packed;
  p.debug = g_conv_debug;

  CUtensorMap tmX0, tmX1, tmDZ;
  int rc;
  if ((rc = make_tmap_nhwc_bf16(&tmX0, src0, B, H, W, c0, 64, BWp, BHp))) return rc;
  if (c1 > 0) {
    if ((rc = make_tmap_nhwc_bf16(&tmX1, src1, B, H, W, c1, 64, BWp, BHp))) return rc;
  } else {
    tmX1 = tmX0;
  }
  if (dz_view) {
    if ((rc = make_tmap_nhwc_view_bf16(&tmDZ, dz, B, H, W, cout, dz_view->pix, dz_view->row, dz_view->img,
                                       64, BWp, BHp))) return rc;
  } else {
    if ((rc = make_tmap_nhwc_bf16(&tmDZ, dz, B, H, W, cout, 64, BWp, BHp))) return rc;
  }
  switch (kn) {
    case 256: return launch_wgrad<256>(tmX0, tmX1, tmDZ, p, stream);
    case 128: return launch_wgrad<128>(tmX0, tmX1, tmDZ, p, stream);
    default:  return launch_wgrad<64>(tmX0, tmX1, tmDZ, p, stream);
  }
}

int conv_wgrad_bf16(const void* src0, int c0, const void* src1, int c1,
                    const void* dz, int cout, int B, int H, int W, int taps,
                    float* dw_packed, cudaStream_t stream) {
  return conv_wgrad_bf16_ex(src0, c0, src1, c1, dz, cout, B, H, W, taps, dw_packed, nullptr, stream);
}

}  // namespace ugp
