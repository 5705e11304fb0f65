// 3x3 (pad 1) / 1x1 convolution as an implicit GEMM on tcgen05 tensor cores.
//
// Replaces the ATen/cuDNN call sites nn.Conv2d forward and its input-gradient
// (reference: UG_unet_parts.py:10,13 and their autograd), see SURVEY.md K1/K2.
//
//   D[m, n] = sum_{tap, c} A[pixel(m) + shift(tap), c] * Wp[tap][n][c]
//
//   * A: NHWC bf16 activations, read by TMA as 4-D boxes (64 ch x BW x BH x 1).
//     The tap shift is applied to the TMA coordinates; out-of-image pixels are
//     zero-filled by the TMA unit, so there is no im2col and no padding copy.
//     Up to two sources are walked back to back in K (the skip-concat of
//     UG_unet_parts.py:80 is never materialised).
//   * Wp: packed bf16 weights [tap][Cout][Cin] (K-major), TMA 2-D boxes.
//   * D accumulates in TMEM (fp32), two accumulator stages so the epilogue of
//     tile i overlaps the main loop of tile i+1.
//   * Epilogue: +bias, optional per-channel affine (eval-mode BatchNorm fold)
//     and ReLU, bf16 convert, swizzled smem staging, TMA store; per-channel
//     sum / sum-of-squares of the stored (bf16-rounded) values for train-mode
//     BatchNorm statistics (UG_unet_parts.py:11,14).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer and TMEM
// owner, warps 2..5 = epilogue (one TMEM lane quarter each). Persistent CTAs,
// static round-robin over (m_tile, n_tile).
#include "ugp_common.cuh"
#include "conv_epilogue.cuh"
#include "conv_gemm.cuh"

namespace ugp {

template <int kN>
struct ConvCfg {
  // A pipeline stage carries kG consecutive K blocks (one 64-channel A box + one B box each)
  // behind ONE mbarrier: the single MMA-issuing thread pays ~100+ cycles per barrier wait, so a
  // stage must hold enough tensor work to hide it (N = 256: 4 MMAs x 128 cycles; N = 128: 8 x 64).
  static constexpr int kG = (kN == 128) ? 2 : 1;
  static constexpr int kStages = (kN == 256) ? 3 : (kN == 128 ? 3 : 6);
  static constexpr int kStagingBufs = (kN == 128) ? 1 : 2;
  static constexpr int kMaxCoutT = (kN == 256) ? kMaxCout : 512;
  static constexpr int kBTileBytes = kN * kBlockK * 2;
  static constexpr int kBlockBytes = kATileBytes + kBTileBytes;   // one K block
  static constexpr int kTmemCols = 2 * kN;  // 128 / 256 / 512: powers of two
  // dynamic smem: stages + staging buffers + stats/bias/scale/shift + barriers
  static constexpr int kSmemBytes = kStages * kG * kBlockBytes + kStagingBufs * kOutChunkBytes +
                                    5 * kMaxCoutT * 4 + kStatsScratchBytes + 256 + 1024;
};

template <int kN>
__global__ void __launch_bounds__(192, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0,
                 const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmO0,
                 const __grid_constant__ CUtensorMap tmO1,
                 const ConvGemmParams p) {
  using Cfg = ConvCfg<kN>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kG = Cfg::kG;
  constexpr int kMaxC = Cfg::kMaxCoutT;
  uint8_t* sA = smem;                                   // [stage][g] A boxes
  uint8_t* sB = sA + kStages * kG * kATileBytes;        // [stage][g] B boxes
  uint8_t* sOut = sB + kStages * kG * Cfg::kBTileBytes;
  float* sStats = reinterpret_cast<float*>(sOut + Cfg::kStagingBufs * kOutChunkBytes);  // [2][kMaxC]
  float* sBias = sStats + 2 * kMaxC;
  float* sScale = sBias + kMaxC;
  float* sShift = sScale + kMaxC;
  uint8_t* sScratch = reinterpret_cast<uint8_t*>(sShift + kMaxC);   // kStatsScratchBytes
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sScratch + kStatsScratchBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;
  const int k_blocks = p.taps * (p.chunks0 + p.chunks1);

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO0);
    tma_prefetch_desc(&tmO1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);      // one elected lane per epilogue warp arrives
    }
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  pdl_wait();      // everything above overlaps the previous kernel's tail (programmatic dependent launch)
  if (warp_idx >= 2) {
    const int et = threadIdx.x - 64;
    for (int i = et; i < p.cout; i += 128) {
      sStats[i] = 0.f;
      sStats[kMaxC + i] = 0.f;
      sBias[i] = p.bias ? p.bias[i % p.bias_period] : 0.f;
      sScale[i] = p.scale ? p.scale[i] : 1.f;
      sShift[i] = p.shift ? p.shift[i] : 0.f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // lane-0 broadcast: marks the value warp-uniform so the MMA operands stay in uniform registers
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    // whole-warp loop, elected lane issues (keeps tensor-map / barrier operands in uniform registers)
    const bool leader = elect_one() != 0;
    {
      int stage = 0;
      uint32_t phase = 0;
      const int chunks = p.chunks0 + p.chunks1;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles;
        const int n_tile = tile - m_tile * p.n_tiles;
        const int tx = m_tile % p.tiles_x;
        const int t2 = m_tile / p.tiles_x;
        const int ty = t2 % p.tiles_y;
        const int b = t2 / p.tiles_y;
        const int x0 = tx * p.BW, y0 = ty * p.BH;
        for (int kb0 = 0; kb0 < k_blocks; kb0 += kG) {
          const int g = (k_blocks - kb0 < kG) ? (k_blocks - kb0) : kG;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          // an A box carries valid_rows (= BW*BH <= 128) rows of 128 B
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], g * (p.valid_rows * 128 + Cfg::kBTileBytes));
          for (int i = 0; i < g; ++i) {
            const int kb = kb0 + i;
            const int tap = kb / chunks;
            const int c = kb - tap * chunks;
            const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
            const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
            uint8_t* a_dst = sA + (stage * kG + i) * kATileBytes;
            if (leader) {
              if (c < p.chunks0)
                tma_load_4d(&tmA0, &full_bar[stage], a_dst, c * kBlockK, x0 + dx, y0 + dy, b);
              else
                tma_load_4d(&tmA1, &full_bar[stage], a_dst, (c - p.chunks0) * kBlockK, x0 + dx,
                            y0 + dy, b);
              tma_load_2d(&tmW, &full_bar[stage], sB + (stage * kG + i) * Cfg::kBTileBytes,
                          c * kBlockK, tap * p.cout + n_tile * kN);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    // whole-warp issue loop (uniform-datapath operands), one elected lane executes tcgen05 (conv_halo.cu)
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      if (tile + static_cast<int>(gridDim.x) >= num_tiles) pdl_trigger();   // last tile of this CTA: the next kernel may be scheduled
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + as * kN;
      for (int kb0 = 0; kb0 < k_blocks; kb0 += kG) {
        const int g = (k_blocks - kb0 < kG) ? (k_blocks - kb0) : kG;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        {
          for (int i = 0; i < g; ++i) {
            const uint64_t a_desc =
                umma_desc_sw128(smem_u32(sA + (stage * kG + i) * kATileBytes), 16, 1024);
            const uint64_t b_desc =
                umma_desc_sw128(smem_u32(sB + (stage * kG + i) * Cfg::kBTileBytes), 16, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // +32 B per K=16 step inside the 128 B swizzle row
              const uint32_t acc = (kb0 + i > 0 || k > 0) ? 1u : 0u;
              if (leader) umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
            }
          }
          if (leader) {
            umma_commit(&empty_bar[stage]);
            if (kb0 + g >= k_blocks) umma_commit(&tfull_bar[as]);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int et = threadIdx.x - 64;   // 0..127
    const int q = warp_idx & 3;        // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;     // row of the 128-row tile
    const bool do_stats = p.stats != nullptr;
    EpilogueCtx ectx;
    ectx.bias_addr = smem_u32(sBias);
    ectx.scale_addr = smem_u32(sScale);
    ectx.shift_addr = smem_u32(sShift);
    ectx.stats_addr = smem_u32(sStats);
    ectx.scratch_addr = smem_u32(sScratch);
    ectx.stats_stride = kMaxC;
    ectx.has_bias = p.bias != nullptr;
    ectx.has_affine = p.scale != nullptr && p.bwd_z == nullptr;
    ectx.bwd_z = p.bwd_z;
    ectx.bwd_z_words = p.cout >> 1;
    ectx.relu = p.relu != 0;
    ectx.do_stats = do_stats;
    ectx.valid_rows = p.valid_rows;
    uint32_t chunk_counter = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      // 2-D destination: (channel, first pixel of the tile); strided 4-D view: (channel, x0, y0, b)
      int o_row = m_tile * kBlockM, o_y = -1, o_b = 0;
      if (p.out4d) {
        const int t2 = m_tile / p.tiles_x;
        o_row = (m_tile - t2 * p.tiles_x) * p.BW;
        o_b = t2 / p.tiles_y;
        o_y = (t2 - o_b * p.tiles_y) * p.BH;
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after_sync();
#pragma unroll 1
      for (int cc = 0; cc < kN / 64; ++cc, ++chunk_counter) {
        uint32_t v[64];
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kN + cc * 64;
        tmem_ld_32x32(taddr, v);
        tmem_ld_32x32(taddr + 32, v + 32);
        tmem_ld_wait();
        if (cc == kN / 64 - 1) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        const int n_base = n_tile * kN + cc * 64;
        const uint32_t buf_addr = smem_u32(sOut) + (chunk_counter % Cfg::kStagingBufs) * kOutChunkBytes;
        if (n_base < p.split_n)
          epilogue_chunk<Cfg::kStagingBufs>(ectx, v, n_base, buf_addr, row, et, &tmO0, n_base,
                                            o_row, o_y, o_b);
        else
          epilogue_chunk<Cfg::kStagingBufs>(ectx, v, n_base, buf_addr, row, et, &tmO1,
                                            n_base - p.split_n, o_row, o_y, o_b);
      }
    }
    if (et == 0) tma_store_wait_all();
    if (do_stats) {
      named_bar_sync(1, 128);
      if (p.bwd_z != nullptr) {
        // BatchNorm-backward sums of this CTA -> the layer's [2][C] totals (fp32 atomics, like the other
        // gradient reductions of the backward pass)
        for (int i = et; i < p.cout; i += 128) {
          atomicAdd(p.stats + i, sStats[i]);
          atomicAdd(p.stats + p.cout + i, sStats[kMaxC + i]);
        }
      } else {
        for (int i = et; i < p.cout; i += 128) {
          // this CTA's slot of the [SMs][2][C] partial-statistics buffer (bn_finalize sums the slots in order)
          p.stats[(static_cast<size_t>(blockIdx.x) * 2 + 0) * p.cout + i] = sStats[i];
          p.stats[(static_cast<size_t>(blockIdx.x) * 2 + 1) * p.cout + i] = sStats[kMaxC + i];
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int kN>
static int launch_conv(const CUtensorMap& tmA0, const CUtensorMap& tmA1,
                       const CUtensorMap& tmW, const CUtensorMap& tmO0,
                       const CUtensorMap& tmO1, const ConvGemmParams& p,
                       cudaStream_t stream) {
  using Cfg = ConvCfg<kN>;
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<kN>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    configured = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  UGP_CHECK_CUDA(launch_k(conv_gemm_kernel<kN>, dim3(grid), dim3(192), Cfg::kSmemBytes, stream, tmA0, tmA1, tmW, tmO0, tmO1, p));
  return UGP_OK;
}

// Picks the 128-pixel box (BW x BH) for an H x W image whose tiles are runs of consecutive NHWC pixels
// (full rows, or pieces of one row): the destination is then a plain 2-D [pixels, C] map.
static bool pick_box(int H, int W, int pixels, int* BW, int* BH) {
  int bw = W < pixels ? W : pixels;
  if (W % bw != 0) return false;
  int bh = pixels / bw;
  if (bw * bh != pixels || H % bh != 0) return false;
  *BW = bw;
  *BH = bh;
  return true;
}

// Any other image size (224 -> 112 -> 56 -> 28 ..., H != W): the box with the most pixels <= `pixels`
// whose sides divide W and H, so tiles never overhang the image.  BW*BH may be < `pixels`: the unused
// accumulator rows are masked in the epilogue; the tile is stored through a 4-D (C, x, y, b) map.
bool pick_box_general(int H, int W, int pixels, int* BW, int* BH) {
  int best = 0;
  for (int bw = 1; bw <= W && bw <= pixels; ++bw) {
    if (W % bw != 0) continue;
    for (int bh = pixels / bw; bh >= 1; --bh) {
      if (H % bh != 0 || bh > 256) continue;
      if (bw * bh > best || (bw * bh == best && bw > *BW)) { best = bw * bh; *BW = bw; *BH = bh; }
      break;
    }
  }
  return best > 0;
}

// conv_halo.cu
bool conv_halo_eligible(int H, int W, int cout, int split_n, int taps);
int conv_halo_bf16(const void* src0, int c0, const void* src1, int c1, int B, int H, int W,
                   const void* wpack, int cout, void* dst0, void* dst1, int split_n,
                   const float* bias, const float* scale, const float* shift, int relu, float* stats,
                   const void* bwd_z, cudaStream_t stream);
extern int g_conv_halo_mode;   // 0 = per-tap loads everywhere, 1 = halo staging for wide thin layers
int g_conv_small_kn = 1;       // 1 = 128-wide tiles when 256-wide ones would fill < 60 % of the SMs (16 x 16 layers: 38 -> 32 us)

// conv_pair.cu
bool conv_pair_eligible(int m_tiles, int cout, int split_n);
int launch_conv_pair(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmW,
                     const CUtensorMap& tmO0, const CUtensorMap& tmO1, const ConvGemmParams& p,
                     cudaStream_t stream);

// src_view / dst_view (nullable): the sources / destinations are strided NHWC views instead of dense
// tensors (ConvTranspose2d phases, conv_transpose.cu); bias_period > 0: bias[n % bias_period].
int conv_gemm_bf16_ex(const void* src0, int c0, const void* src1, int c1, int B,
                      int H, int W, int taps, const void* wpack, int cout,
                      void* dst0, void* dst1, int split_n, const float* bias,
                      const float* scale, const float* shift, int relu,
                      float* stats, const ConvView* src_view, const ConvView* dst_view,
                      int bias_period, cudaStream_t stream, const void* bwd_z) {
  UGP_REQUIRE(src0 && wpack && dst0, "conv_gemm: null pointer");
  UGP_REQUIRE(taps == 9 || taps == 1, "conv_gemm: taps must be 9 or 1 (got %d)", taps);
  UGP_REQUIRE(c0 > 0 && c0 % 64 == 0 && c1 >= 0 && c1 % 64 == 0,
              "conv_gemm: source channels must be multiples of 64 (got %d, %d)", c0, c1);
  UGP_REQUIRE(c1 == 0 || src1, "conv_gemm: second source missing");
  UGP_REQUIRE(cout % 64 == 0 && cout <= kMaxCout, "conv_gemm: cout %d unsupported", cout);
  if (dst1 == nullptr) split_n = cout;
  UGP_REQUIRE(split_n > 0 && split_n <= cout && split_n % 64 == 0,
              "conv_gemm: bad split_n %d", split_n);
  const bool viewed = src_view != nullptr || dst_view != nullptr;
  if (bwd_z != nullptr) {
    // BatchNorm-backward statistics mode: [2][cout] fp32 totals accumulated into `stats` (zeroed here)
    UGP_REQUIRE(stats && scale && shift && !viewed && split_n == cout && !relu,
                "conv_gemm: BatchNorm-backward statistics need stats / scale / shift and one dense destination");
    UGP_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * cout, stream));
  }
  if (!viewed && g_conv_halo_mode != 0 && conv_halo_eligible(H, W, cout, split_n, taps))
    return conv_halo_bf16(src0, c0, src1, c1, B, H, W, wpack, cout, dst0, dst1, split_n, bias, scale,
                          shift, relu, stats, bwd_z, stream);
  int BW = 0, BH = 0;
  const bool general = !pick_box(H, W, kBlockM, &BW, &BH);
  UGP_REQUIRE(bwd_z == nullptr || !general, "conv_gemm: BatchNorm-backward statistics need power-of-two tiles");
  if (general) {
    UGP_REQUIRE(H > 0 && W > 0 && pick_box_general(H, W, kBlockM, &BW, &BH),
                "conv_gemm: unsupported spatial size %dx%d", H, W);
    UGP_REQUIRE(dst_view == nullptr && src_view == nullptr, "conv_gemm: strided views need power-of-two tiles");
  }
  int kn = 256;
  while (kn > 64 && cout % kn != 0) kn >>= 1;   // the epilogue routes each 64-column chunk to dst0 / dst1
  // Few tiles (the 16 x 16 bottleneck at batch 16: 32 m tiles x 2 = 64 tiles on 148 SMs): 128-wide tiles
  // double the CTAs; each then streams the same K with half the weight bytes (tuning switch conv_small_kn)
  if (g_conv_small_kn != 0 && kn == 256 && cout <= 512 && !general &&
      B * (H * W / kBlockM) * (cout / 256) * 10 < device_sm_count() * 6)
    kn = 128;

  UGP_REQUIRE(kn == 256 || cout <= 512, "conv_gemm: cout %d with a %d-wide tile unsupported", cout, kn);

  ConvGemmParams p;
  p.taps = taps;
  p.chunks0 = c0 / 64;
  p.chunks1 = c1 / 64;
  p.H = H; p.W = W; p.BH = BH; p.BW = BW;
  p.tiles_x = W / BW;
  p.tiles_y = H / BH;
  p.m_tiles = B * p.tiles_x * p.tiles_y;
  p.n_tiles = cout / kn;
  p.cout = cout;
  p.split_n = split_n;
  p.relu = relu;
  p.bias_period = bias_period > 0 ? bias_period : cout;
  p.out4d = dst_view != nullptr || general;
  p.valid_rows = BW * BH;
  p.bias = bias; p.scale = scale; p.shift = shift; p.stats = stats;
  p.bwd_z = static_cast<const uint32_t*>(bwd_z);

  CUtensorMap tmA0, tmA1, tmW, tmO0, tmO1;
  int rc;
  if (src_view) {
    if ((rc = make_tmap_nhwc_view_bf16(&tmA0, src0, B, H, W, c0, src_view->pix, src_view->row,
                                       src_view->img, 64, BW, BH))) return rc;
    if (c1 > 0 && (rc = make_tmap_nhwc_view_bf16(&tmA1, src1, B, H, W, c1, src_view->pix, src_view->row,
                                                 src_view->img, 64, BW, BH))) return rc;
  } else {
    if ((rc = make_tmap_nhwc_bf16(&tmA0, src0, B, H, W, c0, 64, BW, BH))) return rc;
    if (c1 > 0 && (rc = make_tmap_nhwc_bf16(&tmA1, src1, B, H, W, c1, 64, BW, BH))) return rc;
  }
  if (c1 == 0) tmA1 = tmA0;
  // CTA pairs (cta_group::2) for the 256-wide tiles: each CTA stages half of the weight tile
  const bool pair = !viewed && !general && kn == 256 && conv_pair_eligible(p.m_tiles, cout, split_n);
  if ((rc = make_tmap_2d_bf16(&tmW, wpack, (long long)taps * cout, c0 + c1, 64, pair ? 128 : kn))) return rc;
  const long long M = (long long)B * H * W;
  if (dst_view) {
    if ((rc = make_tmap_nhwc_view_bf16(&tmO0, dst0, B, H, W, split_n, dst_view->pix, dst_view->row,
                                       dst_view->img, 64, BW, BH))) return rc;
    if (split_n < cout && (rc = make_tmap_nhwc_view_bf16(&tmO1, dst1, B, H, W, cout - split_n, dst_view->pix,
                                                         dst_view->row, dst_view->img, 64, BW, BH))) return rc;
  } else if (general) {
    if ((rc = make_tmap_nhwc_bf16(&tmO0, dst0, B, H, W, split_n, 64, BW, BH))) return rc;
    if (split_n < cout && (rc = make_tmap_nhwc_bf16(&tmO1, dst1, B, H, W, cout - split_n, 64, BW, BH))) return rc;
  } else {
    if ((rc = make_tmap_2d_bf16(&tmO0, dst0, M, split_n, 64, kBlockM))) return rc;
    if (split_n < cout && (rc = make_tmap_2d_bf16(&tmO1, dst1, M, cout - split_n, 64, kBlockM))) return rc;
  }
  if (split_n >= cout) tmO1 = tmO0;
  if (pair) return launch_conv_pair(tmA0, tmA1, tmW, tmO0, tmO1, p, stream);
  switch (kn) {
    case 256: return launch_conv<256>(tmA0, tmA1, tmW, tmO0, tmO1, p, stream);
    case 128: return launch_conv<128>(tmA0, tmA1, tmW, tmO0, tmO1, p, stream);
    default:  return launch_conv<64>(tmA0, tmA1, tmW, tmO0, tmO1, p, stream);
  }
}

int conv_gemm_bf16(const void* src0, int c0, const void* src1, int c1, int B,
                   int H, int W, int taps, const void* wpack, int cout,
                   void* dst0, void* dst1, int split_n, const float* bias,
                   const float* scale, const float* shift, int relu,
                   float* stats, cudaStream_t stream) {
  return conv_gemm_bf16_ex(src0, c0, src1, c1, B, H, W, taps, wpack, cout, dst0, dst1, split_n, bias,
                           scale, shift, relu, stats, nullptr, nullptr, 0, stream, nullptr);
}

// Input gradient of a 3x3 convolution whose input is relu(BatchNorm(z)) (the second convolution of a
// DoubleConv, UG_unet_parts.py:12-13): dA = conv(dz, wd) as conv_gemm_bf16, and in the same launch the
// reduction pass of that BatchNorm's backward - raw[0][c] = sum dA*[a > 0], raw[1][c] = sum dA*[a > 0]*z.
int conv_dgrad_bnstats_bf16(const void* dz, int c, int B, int H, int W, const void* wd, int cout, void* dA,
                            const void* z, const float* scale, const float* shift, float* raw,
                            cudaStream_t stream) {
  UGP_REQUIRE(z && scale && shift && raw, "conv_dgrad_bnstats: null pointer");
  return conv_gemm_bf16_ex(dz, c, nullptr, 0, B, H, W, 9, wd, cout, dA, nullptr, cout, nullptr, scale, shift, 0,
                           raw, nullptr, nullptr, 0, stream, z);
}

}  // namespace ugp
