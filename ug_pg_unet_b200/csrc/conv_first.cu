// The network's first convolution without an im2col matrix in HBM.
//
// nn.Conv2d(n_channels, 64, 3, padding=1) at UG_unet_parts.py:10 inside InConv has K = 9 * n_channels = 27:
// too thin for the implicit-GEMM kernels (their K blocks are 64 channels of ONE tap), so it runs as a GEMM
// on 64-wide rows holding the whole 3x3 neighbourhood (first_rows.cuh).  Materialising those rows costs
// more HBM traffic than the layer's output: 134 MB written by the packer, read again by the GEMM and a third
// time by the weight gradient, for a 12.6 MB image (batch 16, 256 x 256).  Here the rows never leave the SM:
//
//   forward   producer warps build the 128-pixel x 64 operand tile directly in shared memory, in the
//             SWIZZLE_128B K-major layout a TMA load of the packed matrix would have produced; one thread
//             issues tcgen05.mma against the resident 64 x 64 weight tile; the shared epilogue
//             (conv_epilogue.cuh) adds the bias, stores bf16 through TMA and accumulates the train-mode
//             BatchNorm statistics - or, in eval mode, applies the folded BatchNorm + ReLU.
//             HBM traffic: the image once, the output once.
//   wgrad     dWp[k][co] += sum_pixels row[p][k] * dZ[p][co] (autograd of the same Conv2d): the same rows,
//             built 64 pixels at a time as the MN-major A operand, against TMA-loaded dZ boxes; pixels are
//             split over all SMs and the 64 x 64 partials meet through fp32 vector atomics.
//             HBM traffic: the image once, dZ once.
//
// Both also fold the bilinear resize (uncertainty_guided_trainer.py:208, UG_unet.py:36-41) and the uint8
// loader format in, exactly as the packer does (the row builder is shared, so the bits are identical).
#include "conv_epilogue.cuh"
#include "conv_gemm.cuh"
#include "first_rows.cuh"

namespace ugp {

constexpr int kFirstN = 64;   // output channels of the weight-gradient kernel (the stage-4 network's first layer)
extern int g_conv_debug;
int g_conv_first_fused = 1;   // ugp_set_option("conv_first_fused", 0): callers fall back to packer + GEMM

// Two CTAs per SM (each: warp 0 weights, warp 1 MMA issue, 4 epilogue warps, 4 producer warps), so that two
// epilogues and two tiles' worth of row building overlap on every SM; 96 registers per thread.
constexpr int kFirstThreads = 64 + 128 + 128;
constexpr int kFirstStages = 3;
// Image patches (kPatch mode: fp32 NCHW input at the layer's resolution, power-of-two tile boxes): instead of
// 27 global loads with their address arithmetic per row, warp 0 brings the tile's (BW + 8) x (BH + 2) x 3
// fp32 neighbourhood (4 pixels of margin left and right: aligned box start) in with ONE TMA box per tile (out-of-image coordinates read as zero = the padding) and
// the producers read their 27 values with constant-offset shared loads: ~70 instead of ~600 instructions per
// row on a kernel that is instruction-issue bound.
constexpr int kFirstPatchBytes = 4992;     // >= 136 x 3 x 3 x 4 (W >= 128), 72 x 4 x 3 x 4, 40 x 6 x 3 x 4, 24 x 10 x 3 x 4
constexpr int kPatchLeft = 4;              // the box starts 4 pixels (16 bytes) left of the tile: TMA needs a 16-byte aligned
                                           // start in the innermost dimension (x0 - 1 faults with an illegal instruction)

template <int kN>
struct FirstCfg {
  static constexpr int kBBytes = kN * 128;
  static constexpr int kPatchStages = (kN == 64) ? 3 : 2;     // two CTAs of <= 113 KB must fit one SM
  static constexpr int kSmemBytes = kFirstStages * kATileBytes + kBBytes + 2 * kOutChunkBytes +
                                    kPatchStages * kFirstPatchBytes + 5 * kN * 4 + kStatsScratchBytes + 256 + 1024;
  static_assert(2 * (kSmemBytes + 1024) <= 227 * 1024, "two CTAs per SM");
};

struct FirstParams {
  const void* x;
  int B, Hin, Win, H, W;
  int m_tiles;            // B * H * W / 128
  int ksteps;             // K = 16 MMAs that hold non-zero columns: ceil(9 * Cin / 16)
  int relu;
  int stat_parts;         // slots of the partial-statistics buffer (= SM count)
  int debug;              // what-if timing only (conv_debug): 1 = no TMA store, 4 = producers skip the image loads
  int BW, bw_shift, BH, PW;   // kPatch: tile box BW x BH pixels (BW = 1 << bw_shift), patch row pitch PW floats
  int patch_bytes;            // PW * (BH + 2) * Cin * 4
  const float* bias;
  const float* scale;
  const float* shift;
  float* stats;
  ByteNorm norm;
};

__device__ __forceinline__ void store_row_sw128(uint32_t rowp, int prow, const uint4 (&u)[8]) {
#pragma unroll
  for (int g = 0; g < 8; ++g) st_shared_v4(rowp + ((g ^ (prow & 7)) << 4), u[g].x, u[g].y, u[g].z, u[g].w);
}

template <int kCin>
__device__ __forceinline__ void patch_row(uint32_t pa, int plane, int pw, uint4 (&u)[8]) {
  // pa: shared address of this pixel's top-left neighbour in channel 0; plane / pw: channel / row pitch in floats
  float vals[64];
#pragma unroll
  for (int k = 9 * kCin; k < 64; ++k) vals[k] = 0.f;
#pragma unroll
  for (int ci = 0; ci < kCin; ++ci)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const uint32_t ra = pa + static_cast<uint32_t>((ci * plane + r * pw) << 2);
#pragma unroll
      for (int sx = 0; sx < 3; ++sx) vals[(r * 3 + sx) * kCin + ci] = __uint_as_float(ld_shared_b32(ra + sx * 4));
    }
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    u[g].x = pack_bf16x2(vals[g * 8 + 0], vals[g * 8 + 1]);
    u[g].y = pack_bf16x2(vals[g * 8 + 2], vals[g * 8 + 3]);
    u[g].z = pack_bf16x2(vals[g * 8 + 4], vals[g * 8 + 5]);
    u[g].w = pack_bf16x2(vals[g * 8 + 6], vals[g * 8 + 7]);
  }
}

template <int kN, int kCin, bool kResize, bool kU8, bool kPatch>
__global__ void __launch_bounds__(kFirstThreads, 2)
conv_first_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO,
                      const __grid_constant__ CUtensorMap tmX, const FirstParams p) {
  static_assert(!kPatch || (!kResize && !kU8), "patch mode: fp32 input at the layer's resolution");
  using Cfg = FirstCfg<kN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                     // [stage] 128 rows x 128 B
  uint8_t* sB = sA + kFirstStages * kATileBytes;          // kN rows (cout) x 128 B
  uint8_t* sOut = sB + Cfg::kBBytes;
  uint8_t* sPatch = sOut + 2 * kOutChunkBytes;            // [patch stage] fp32 [c][row][x]
  float* sStats = reinterpret_cast<float*>(sPatch + Cfg::kPatchStages * kFirstPatchBytes);   // [2][kN]
  float* sBias = sStats + 2 * kN;
  float* sScale = sBias + kN;
  float* sShift = sScale + kN;
  uint8_t* sScratch = reinterpret_cast<uint8_t*>(sShift + kN);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sScratch + kStatsScratchBytes);
  uint64_t* a_empty = a_full + kFirstStages;
  uint64_t* tfull_bar = a_empty + kFirstStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_full = tempty_bar + 2;
  uint64_t* p_full = w_full + 1;
  uint64_t* p_empty = p_full + Cfg::kPatchStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + Cfg::kPatchStages);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < kFirstStages; ++i) {
      mbar_init(&a_full[i], 4);        // lane 0 of each of the 4 producer warps
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < Cfg::kPatchStages; ++i) {
      mbar_init(&p_full[i], 1);
      mbar_init(&p_empty[i], 4);       // lane 0 of each producer warp, after its 27 shared loads
    }
    if (kPatch) tma_prefetch_desc(&tmX);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc(tmem_slot, 2 * kN);
  pdl_wait();      // everything above overlaps the previous kernel's tail (programmatic dependent launch)
  if (warp_idx >= 2 && warp_idx < 6) {
    for (int i = threadIdx.x - 64; i < kN; i += 128) {
      sStats[i] = 0.f;
      sStats[kN + i] = 0.f;
      sBias[i] = p.bias ? p.bias[i] : 0.f;
      sScale[i] = p.scale ? p.scale[i] : 1.f;
      sShift[i] = p.shift ? p.shift[i] : 0.f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp_idx == 0) {
    // ===================== weight tile: loaded once, resident; image patches: one TMA box per tile ==========
    const bool leader = elect_one() != 0;
    if (leader) {
      mbar_arrive_expect_tx(w_full, Cfg::kBBytes);
      tma_load_2d(&tmW, w_full, sB, 0, 0);
    }
    if constexpr (kPatch) {
      const int hw = p.H * p.W;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
        const int ps = it % Cfg::kPatchStages;
        mbar_wait(&p_empty[ps], ((it / Cfg::kPatchStages) & 1) ^ 1);
        if (leader) {
          const int pix0 = tile * kBlockM;
          const int b = pix0 / hw;
          const int rem = pix0 - b * hw;
          const int y0 = rem / p.W;
          const int x0 = rem - y0 * p.W;
          mbar_arrive_expect_tx(&p_full[ps], p.patch_bytes);
          tma_load_4d(&tmX, &p_full[ps], sPatch + ps * kFirstPatchBytes, x0 - kPatchLeft, y0 - 1, 0, b);
        }
        __syncwarp();
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kN, 0, 0);
    mbar_wait(w_full, 0);
    const uint64_t b_desc = umma_desc_sw128(smem_u32(sB), 16, 1024);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const int stage = it % kFirstStages;
      mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
      mbar_wait(&a_full[stage], (it / kFirstStages) & 1);
      tc_fence_after_sync();
      const uint64_t a_desc = umma_desc_sw128(smem_u32(sA + stage * kATileBytes), 16, 1024);
      for (int k = 0; k < p.ksteps; ++k)      // +32 B per K = 16 step inside the 128 B swizzle row
        if (leader) umma_bf16(tmem_base + as * kN, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0 ? 1u : 0u);
      if (leader) {
        umma_commit(&a_empty[stage]);
        umma_commit(&tfull_bar[as]);
      }
      __syncwarp();
    }
  } else if (warp_idx < 6) {
    // ===================== epilogue (warps 2..5: TMEM lane quarters 2, 3, 0, 1) =====================
    const int et = threadIdx.x - 64;
    const int q = warp_idx & 3;
    const int row = q * 32 + lane;
    const bool do_stats = p.stats != nullptr;
    EpilogueCtx ectx;
    ectx.bias_addr = smem_u32(sBias);
    ectx.scale_addr = smem_u32(sScale);
    ectx.shift_addr = smem_u32(sShift);
    ectx.stats_addr = smem_u32(sStats);
    ectx.scratch_addr = smem_u32(sScratch);
    ectx.stats_stride = kN;
    ectx.has_bias = p.bias != nullptr;
    ectx.has_affine = p.scale != nullptr;
    ectx.relu = p.relu != 0;
    ectx.do_stats = do_stats;
    ectx.valid_rows = kBlockM;
    ectx.whatif_no_store = (p.debug & 1) != 0;
    uint32_t chunk_counter = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait(&tfull_bar[as], (it >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int cc = 0; cc < kN / 64; ++cc, ++chunk_counter) {
        uint32_t v[64];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kN + cc * 64;
        tmem_ld_32x32(taddr, v);
        tmem_ld_32x32(taddr + 32, v + 32);
        tmem_ld_wait();
        if (cc == kN / 64 - 1) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
        const uint32_t buf_addr = smem_u32(sOut) + (chunk_counter & 1) * kOutChunkBytes;
        epilogue_chunk<2, false>(ectx, v, cc * 64, buf_addr, row, et, &tmO, cc * 64, tile * kBlockM);
      }
    }
    if (et == 0) tma_store_wait_all();
    if (do_stats) {
      named_bar_sync(1, 128);
      // Slot of the [SMs][2][C] partial-statistics buffer (zero on entry; bn_finalize sums the slots in order).
      // CTAs b and b + SMs share slot b: two fp32 atomic adds into a zero commute exactly, so the statistics
      // stay run-to-run identical.
      float* slot = p.stats + static_cast<size_t>(blockIdx.x % p.stat_parts) * 2 * kN;
      for (int i = et; i < kN; i += 128) {
        atomicAdd(slot + i, sStats[i]);
        atomicAdd(slot + kN + i, sStats[kN + i]);
      }
    }
  } else {
    // ===================== producers: one thread builds one row of every tile of this CTA =====================
    const int prow = threadIdx.x - 192;     // 0..127
    const int hw = p.H * p.W;
    const uint32_t rowp0 = smem_u32(sA) + prow * 128;
    auto emit = [&](const uint4 (&u)[8], int it) {
      const int stage = it % kFirstStages;
      mbar_wait(&a_empty[stage], ((it / kFirstStages) & 1) ^ 1);
      store_row_sw128(rowp0 + stage * kATileBytes, prow, u);
      fence_proxy_async_smem();         // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[stage]);
    };
    if constexpr (kPatch) {
      const int tx = prow & (p.BW - 1), ty = prow >> p.bw_shift;
      const int plane = (p.BH + 2) * p.PW;
      const uint32_t base = smem_u32(sPatch) + static_cast<uint32_t>((ty * p.PW + tx + kPatchLeft - 1) << 2);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
        const int ps = it % Cfg::kPatchStages;
        mbar_wait(&p_full[ps], (it / Cfg::kPatchStages) & 1);
        uint4 u[8];
        patch_row<kCin>(base + ps * kFirstPatchBytes, plane, p.PW, u);
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_empty[ps]);        // the patch stage may be refilled
        emit(u, it);
      }
    } else if constexpr (kResize) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
        const int pix = tile * kBlockM + prow;
        const int b = pix / hw;
        const int rem = pix - b * hw;
        const int yo = rem / p.W;
        uint4 u[8];
        first_layer_row<kCin, true, kU8>(p.x, p.Hin, p.Win, p.H, p.W, b, yo, rem - yo * p.W, p.norm, u);
        emit(u, it);
      }
    } else {
      // two rows in flight per thread: the loads of the next tile's row are issued before this one is packed
      uint32_t rawA[9 * kCin], rawB[9 * kCin];
      uint32_t mA = 0, mB = 0;
      auto fetch = [&](int tile, uint32_t (&raw)[9 * kCin]) -> uint32_t {
        const int pix = tile * kBlockM + prow;
        const int b = pix / hw;
        const int rem = pix - b * hw;
        const int yo = rem / p.W;
        return first_layer_fetch<kCin, kU8>(p.x, p.H, p.W, b, yo, rem - yo * p.W, raw);
      };
      int tile = blockIdx.x, it = 0;
      if (p.debug & 4) {
        for (; tile < p.m_tiles; tile += gridDim.x, ++it) {
          uint4 u[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) u[g] = make_uint4(tile, it, g, prow);
          emit(u, it);
        }
      }
      if (tile < p.m_tiles) mA = fetch(tile, rawA);
      while (tile < p.m_tiles) {
        uint4 u[8];
        const int t1 = tile + gridDim.x;
        if (t1 < p.m_tiles) mB = fetch(t1, rawB);
        first_layer_pack<kCin, kU8>(rawA, mA, p.norm, u);
        emit(u, it);
        if (t1 >= p.m_tiles) break;
        const int t2 = t1 + gridDim.x;
        if (t2 < p.m_tiles) mA = fetch(t2, rawA);
        first_layer_pack<kCin, kU8>(rawB, mB, p.norm, u);
        emit(u, it + 1);
        tile = t2;
        it += 2;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 2 * kN);
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
constexpr int kFwStages = 4;
constexpr int kFwBox = 64 * 128;                                  // 64 pixels x 64 values bf16
constexpr int kFwStageBytes = 3 * kFwBox;                         // A box twice (M = 128) + dZ box
constexpr int kFwPatchBytes = 2688;                               // >= 72 x 3 x 3 x 4 (W >= 64), 40 x 4 x 3 x 4, 24 x 6 x 3 x 4
constexpr int kFwSmemBytes = kFwStages * (kFwStageBytes + kFwPatchBytes) + 256 + 1024;
constexpr int kFwThreads = 64 + 256;                              // warp 0 dZ loads, warp 1 MMA, 8 producer warps

struct FirstWgradParams {
  const void* x;
  int B, Hin, Win, H, W;
  int kchunks_total;      // B * H * W / 64
  int kchunks_per_split;
  float* dw;              // [64 (k)][64 (cout)] fp32, accumulated
  ByteNorm norm;
  int BW, bw_shift, BH, PW;   // kPatch: chunk box BW x BH = 64 pixels, patch row pitch PW floats
  int patch_bytes;
};

__device__ __forceinline__ void red_add_f32x4_first(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int kCin, bool kResize, bool kU8, bool kPatch>
__global__ void __launch_bounds__(kFwThreads, 2)
conv_first_wgrad_kernel(const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ CUtensorMap tmX,
                        const FirstWgradParams p) {
  static_assert(!kPatch || (!kResize && !kU8), "patch mode: fp32 input at the layer's resolution");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sPatch = smem + kFwStages * kFwStageBytes;     // [stage] fp32 [c][row][x]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sPatch + kFwStages * kFwPatchBytes);
  uint64_t* empty_bar = full_bar + kFwStages;
  uint64_t* done_bar = empty_bar + kFwStages;
  uint64_t* p_full = done_bar + 1;
  uint64_t* p_empty = p_full + kFwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + kFwStages);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kc_begin = blockIdx.x * p.kchunks_per_split;
  int kc_end = kc_begin + p.kchunks_per_split;
  if (kc_end > p.kchunks_total) kc_end = p.kchunks_total;
  const int n_k = kc_end - kc_begin;      // host guarantees > 0

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmDZ);
    for (int i = 0; i < kFwStages; ++i) {
      mbar_init(&full_bar[i], 3);      // the dZ box (expect_tx arrival) + the 2 producer warps of the chunk
      mbar_init(&empty_bar[i], 1);
      mbar_init(&p_full[i], 1);
      mbar_init(&p_empty[i], 2);       // lane 0 of the chunk's 2 producer warps, after their shared loads
    }
    mbar_init(done_bar, 1);
    if (kPatch) tma_prefetch_desc(&tmX);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc(tmem_slot, kFirstN);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp_idx == 0) {
    // ===================== dZ boxes: 64 pixels x 64 channels by TMA =====================
    const bool leader = elect_one() != 0;
    for (int i = 0; i < n_k; ++i) {
      const int stage = i % kFwStages;
      mbar_wait(&empty_bar[stage], ((i / kFwStages) & 1) ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&full_bar[stage], kFwBox);
        tma_load_2d(&tmDZ, &full_bar[stage], smem + stage * kFwStageBytes + 2 * kFwBox, 0, (kc_begin + i) * 64);
      }
      if constexpr (kPatch) {
        mbar_wait(&p_empty[stage], ((i / kFwStages) & 1) ^ 1);
        if (leader) {
          const int hw = p.H * p.W;
          const int pix0 = (kc_begin + i) * 64;
          const int b = pix0 / hw;
          const int rem = pix0 - b * hw;
          const int y0 = rem / p.W;
          const int x0 = rem - y0 * p.W;
          mbar_arrive_expect_tx(&p_full[stage], p.patch_bytes);
          tma_load_4d(&tmX, &p_full[stage], sPatch + stage * kFwPatchBytes, x0 - kPatchLeft, y0 - 1, 0, b);
        }
      }
      __syncwarp();
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer: D[k][co] += rows^T . dZ, 16 pixels per instruction =====================
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = umma_idesc_bf16(128, kFirstN, 1, 1);
    const uint64_t desc0 = umma_desc_sw128(0, kFwBox, 1024);   // MN-major: next 64 M rows one box further
    for (int i = 0; i < n_k; ++i) {
      const int stage = i % kFwStages;
      mbar_wait(&full_bar[stage], (i / kFwStages) & 1);
      tc_fence_after_sync();
      const uint32_t a_addr = smem_u32(smem + stage * kFwStageBytes);
      const uint32_t b_addr = a_addr + 2 * kFwBox;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t a_desc = desc0 + ((a_addr + k * 2048) >> 4);
        const uint64_t b_desc = desc0 + ((b_addr + k * 2048) >> 4);
        if (leader) umma_bf16(tmem_base, a_desc, b_desc, idesc, (i > 0 || k > 0) ? 1u : 0u);
      }
      if (leader) {
        umma_commit(&empty_bar[stage]);
        if (i == n_k - 1) umma_commit(done_bar);
      }
      __syncwarp();
    }
  } else {
    // ===================== producers: four pairs of warps take every fourth 64-pixel chunk =====================
    const int pt = threadIdx.x - 64;        // 0..255
    const int pair = pt >> 6;
    const int prow = pt & 63;
    const int hw = p.H * p.W;
    const uint32_t rowp0 = smem_u32(smem) + prow * 128;
    auto emit = [&](const uint4 (&u)[8], int i) {
      const int stage = i % kFwStages;
      mbar_wait(&empty_bar[stage], ((i / kFwStages) & 1) ^ 1);
      store_row_sw128(rowp0 + stage * kFwStageBytes, prow, u);
      store_row_sw128(rowp0 + stage * kFwStageBytes + kFwBox, prow, u);   // accumulator rows 64..127 (unused copy)
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
    };
    if constexpr (kPatch) {
      const int tx = prow & (p.BW - 1), ty = prow >> p.bw_shift;
      const int plane = (p.BH + 2) * p.PW;
      const uint32_t base = smem_u32(sPatch) + static_cast<uint32_t>((ty * p.PW + tx + kPatchLeft - 1) << 2);
      for (int i = pair; i < n_k; i += 4) {
        const int stage = i % kFwStages;
        mbar_wait(&p_full[stage], (i / kFwStages) & 1);
        uint4 u[8];
        patch_row<kCin>(base + stage * kFwPatchBytes, plane, p.PW, u);
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_empty[stage]);
        emit(u, i);
      }
    } else if constexpr (kResize) {
      for (int i = pair; i < n_k; i += 4) {
        const int pix = (kc_begin + i) * 64 + prow;
        const int b = pix / hw;
        const int rem = pix - b * hw;
        const int yo = rem / p.W;
        uint4 u[8];
        first_layer_row<kCin, true, kU8>(p.x, p.Hin, p.Win, p.H, p.W, b, yo, rem - yo * p.W, p.norm, u);
        emit(u, i);
      }
    } else {
      uint32_t rawA[9 * kCin], rawB[9 * kCin];
      uint32_t mA = 0, mB = 0;
      auto fetch = [&](int i, uint32_t (&raw)[9 * kCin]) -> uint32_t {
        const int pix = (kc_begin + i) * 64 + prow;
        const int b = pix / hw;
        const int rem = pix - b * hw;
        const int yo = rem / p.W;
        return first_layer_fetch<kCin, kU8>(p.x, p.H, p.W, b, yo, rem - yo * p.W, raw);
      };
      int i = pair;
      if (i < n_k) mA = fetch(i, rawA);
      while (i < n_k) {
        uint4 u[8];
        if (i + 4 < n_k) mB = fetch(i + 4, rawB);
        first_layer_pack<kCin, kU8>(rawA, mA, p.norm, u);
        emit(u, i);
        if (i + 4 >= n_k) break;
        if (i + 8 < n_k) mA = fetch(i + 8, rawA);
        first_layer_pack<kCin, kU8>(rawB, mB, p.norm, u);
        emit(u, i + 4);
        i += 8;
      }
    }
    // ===================== reduction of this CTA's 64 x 64 partial (warps 2..5: lane quarters 2, 3, 0, 1) =====
    if (warp_idx < 6) {
      const int q = warp_idx & 3;
      const int row = q * 32 + lane;       // accumulator row = k index; rows 64..127 repeat rows 0..63
      mbar_wait(done_bar, 0);
      tc_fence_after_sync();
      if (row < 64) {
#pragma unroll 1
        for (int c = 0; c < kFirstN / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
          tmem_ld_wait();
          float* dst = p.dw + row * kFirstN + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_f32x4_first(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                __uint_as_float(v[j + 3]));
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, kFirstN);
  }
}

// ------------------------------------------------------------------------------------------------ host
bool conv_first_supported(int Cin, int cout, int B, int H, int W) {
  const long long total = static_cast<long long>(B) * H * W;
  return g_conv_first_fused != 0 && Cin == 3 && (cout == 64 || cout == 128) && H > 0 && W > 0 &&
         total % kBlockM == 0 && total < (1ll << 30);
}
bool conv_first_wgrad_supported(int Cin, int cout, int B, int H, int W) {
  return cout == kFirstN && conv_first_supported(Cin, cout, B, H, W);
}

int g_conv_first_patch = 1;   // ugp_set_option("conv_first_patch", 0): per-row global loads instead of TMA image patches

// Patch mode applies to fp32 images at the layer's resolution whose `pixels`-pixel tiles are power-of-two
// boxes inside one image (W >= pixels: pieces of a row; W < pixels: whole rows); fills the box geometry.
static bool patch_geometry(const void* x, bool u8, bool resize, int Cin, int H, int W, int pixels, int* BW, int* shift,
                           int* BH, int* PW, int* bytes) {
  if (g_conv_first_patch == 0 || u8 || resize || Cin != 3) return false;
  if (reinterpret_cast<uintptr_t>(x) % 16 != 0 || W % 4 != 0 || W < 16) return false;
  const int bw = W < pixels ? W : pixels;
  if ((bw & (bw - 1)) != 0 || W % bw != 0) return false;
  const int bh = pixels / bw;
  if (H % bh != 0) return false;
  int sh = 0;
  while ((1 << sh) < bw) ++sh;
  *BW = bw; *shift = sh; *BH = bh; *PW = bw + 2 * kPatchLeft;
  *bytes = (*PW) * (bh + 2) * Cin * 4;
  return true;
}

template <int kN, bool kResize, bool kU8, bool kPatch>
static int launch_first_fwd(const CUtensorMap& tmW, const CUtensorMap& tmO, const CUtensorMap& tmX, const FirstParams& p,
                            cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(conv_first_fwd_kernel<kN, 3, kResize, kU8, kPatch>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, FirstCfg<kN>::kSmemBytes));
    configured = true;
  }
  const int cap = 2 * device_sm_count();
  const int grid = p.m_tiles < cap ? p.m_tiles : cap;
  UGP_CHECK_CUDA(launch_k(conv_first_fwd_kernel<kN, 3, kResize, kU8, kPatch>, dim3(grid), dim3(kFirstThreads),
                          FirstCfg<kN>::kSmemBytes, stream, tmW, tmO, tmX, p));
  return UGP_OK;
}

template <int kN>
static int dispatch_first_fwd(bool resize, bool u8, bool patch, const CUtensorMap& tmW, const CUtensorMap& tmO,
                              const CUtensorMap& tmX, const FirstParams& p, cudaStream_t stream) {
  if (patch) return launch_first_fwd<kN, false, false, true>(tmW, tmO, tmX, p, stream);
  if (u8) return resize ? launch_first_fwd<kN, true, true, false>(tmW, tmO, tmX, p, stream)
                        : launch_first_fwd<kN, false, true, false>(tmW, tmO, tmX, p, stream);
  return resize ? launch_first_fwd<kN, true, false, false>(tmW, tmO, tmX, p, stream)
                : launch_first_fwd<kN, false, false, false>(tmW, tmO, tmX, p, stream);
}

int conv_first_fwd(const void* x, int is_u8, int B, int Cin, int Hin, int Win, int H, int W, const float* mean,
                   const float* stddev, const void* wpack, int cout, void* dst, const float* bias,
                   const float* scale, const float* shift, int relu, float* stats, cudaStream_t stream) {
  UGP_REQUIRE(x && wpack && dst, "conv_first_fwd: null pointer");
  UGP_REQUIRE(conv_first_supported(Cin, cout, B, H, W),
              "conv_first_fwd: needs Cin = 3, cout = 64 or 128 and B*H*W %% 128 == 0 (got Cin %d, cout %d, %d x %d x %d)",
              Cin, cout, B, H, W);
  FirstParams p{};
  p.x = x; p.B = B; p.Hin = Hin; p.Win = Win; p.H = H; p.W = W;
  const long long total = static_cast<long long>(B) * H * W;
  p.m_tiles = static_cast<int>(total / kBlockM);
  p.ksteps = (9 * Cin + 15) / 16;
  p.relu = relu; p.bias = bias; p.scale = scale; p.shift = shift; p.stats = stats;
  p.stat_parts = device_sm_count();
  p.debug = g_conv_debug;
  if (is_u8) {
    if (int rc = make_byte_norm(&p.norm, Cin, mean, stddev, "conv_first_fwd")) return rc;
  }
  CUtensorMap tmW, tmO;
  int rc;
  if ((rc = make_tmap_2d_bf16(&tmW, wpack, cout, 64, 64, cout))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmO, dst, total, cout, 64, kBlockM))) return rc;
  const bool resize = Hin != H || Win != W;
  CUtensorMap tmX = tmW;    // unused outside patch mode
  const bool patch = patch_geometry(x, is_u8 != 0, resize, Cin, H, W, kBlockM, &p.BW, &p.bw_shift, &p.BH, &p.PW,
                                    &p.patch_bytes);
  if (patch) {
    UGP_REQUIRE(p.patch_bytes <= kFirstPatchBytes, "conv_first_fwd: patch of %d bytes does not fit", p.patch_bytes);
    if ((rc = make_tmap_nchw_f32(&tmX, x, B, Cin, H, W, p.PW, p.BH + 2, Cin))) return rc;
  }
  return cout == 64 ? dispatch_first_fwd<64>(resize, is_u8 != 0, patch, tmW, tmO, tmX, p, stream)
                    : dispatch_first_fwd<128>(resize, is_u8 != 0, patch, tmW, tmO, tmX, p, stream);
}

template <bool kResize, bool kU8, bool kPatch>
static int launch_first_wgrad(const CUtensorMap& tmDZ, const CUtensorMap& tmX, const FirstWgradParams& p, int grid,
                              cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(conv_first_wgrad_kernel<3, kResize, kU8, kPatch>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kFwSmemBytes));
    configured = true;
  }
  conv_first_wgrad_kernel<3, kResize, kU8, kPatch><<<grid, kFwThreads, kFwSmemBytes, stream>>>(tmDZ, tmX, p);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

int conv_first_wgrad(const void* x, int is_u8, int B, int Cin, int Hin, int Win, int H, int W, const float* mean,
                     const float* stddev, const void* dz, int cout, float* dw_packed, cudaStream_t stream) {
  UGP_REQUIRE(x && dz && dw_packed, "conv_first_wgrad: null pointer");
  UGP_REQUIRE(conv_first_wgrad_supported(Cin, cout, B, H, W),
              "conv_first_wgrad: needs Cin = 3, cout = 64 and B*H*W %% 128 == 0 (got Cin %d, cout %d, %d x %d x %d)",
              Cin, cout, B, H, W);
  FirstWgradParams p{};
  p.x = x; p.B = B; p.Hin = Hin; p.Win = Win; p.H = H; p.W = W;
  const long long total = static_cast<long long>(B) * H * W;
  p.kchunks_total = static_cast<int>(total / 64);
  p.dw = dw_packed;
  if (is_u8) {
    if (int rc = make_byte_norm(&p.norm, Cin, mean, stddev, "conv_first_wgrad")) return rc;
  }
  // pixels split over one wave of CTAs (two per SM), at least 8 chunks (512 pixels) each
  int splits = 2 * device_sm_count();
  int max_splits = p.kchunks_total / 8;
  if (max_splits < 1) max_splits = 1;
  if (splits > max_splits) splits = max_splits;
  p.kchunks_per_split = (p.kchunks_total + splits - 1) / splits;
  splits = (p.kchunks_total + p.kchunks_per_split - 1) / p.kchunks_per_split;
  CUtensorMap tmDZ;
  if (int rc = make_tmap_2d_bf16(&tmDZ, dz, total, cout, 64, 64)) return rc;
  const bool resize = Hin != H || Win != W;
  CUtensorMap tmX = tmDZ;   // unused outside patch mode
  if (patch_geometry(x, is_u8 != 0, resize, Cin, H, W, 64, &p.BW, &p.bw_shift, &p.BH, &p.PW, &p.patch_bytes)) {
    UGP_REQUIRE(p.patch_bytes <= kFwPatchBytes, "conv_first_wgrad: patch of %d bytes does not fit", p.patch_bytes);
    if (int rc = make_tmap_nchw_f32(&tmX, x, B, Cin, H, W, p.PW, p.BH + 2, Cin)) return rc;
    return launch_first_wgrad<false, false, true>(tmDZ, tmX, p, splits, stream);
  }
  if (is_u8) return resize ? launch_first_wgrad<true, true, false>(tmDZ, tmX, p, splits, stream)
                           : launch_first_wgrad<false, true, false>(tmDZ, tmX, p, splits, stream);
  return resize ? launch_first_wgrad<true, false, false>(tmDZ, tmX, p, splits, stream)
                : launch_first_wgrad<false, false, false>(tmDZ, tmX, p, splits, stream);
}

}  // namespace ugp
