// 3x3 / 1x1 implicit-GEMM convolution on CTA PAIRS (tcgen05 cta_group::2) for the layers with
// Cout (or, for the input gradient, Cin) a multiple of 256.
//
// Same math, operands and epilogue as conv_gemm.cu (forward + dgrad of nn.Conv2d,
// UG_unet_parts.py:10,13); what changes is who feeds the tensor core.  With both operands in
// shared memory one SM fetches 64 B/clk of operands, so a 128x256x16 MMA costs 64 + 128 cycles
// of fetch for 128 cycles of math (measured).  A CTA pair computes a 256x256 tile:
// each CTA stages its own 128 pixels of A and only HALF of the weight tile (128 of the 256 output
// channels); the leader CTA issues tcgen05.mma.cta_group::2, which reads A and the B halves from
// both CTAs' shared memory and writes 128 accumulator rows into each CTA's TMEM.  Per-CTA fetch
// drops to 64 + 64 cycles per MMA (= the math time) and the L2 -> SM weight traffic halves.
//
// Protocol (s = smem stage, a = TMEM accumulator stage):
//   full[s]   lives in the leader; the leader's producer arms it with the bytes of BOTH CTAs and
//             both CTAs' TMA loads complete on it (cp.async.bulk.tensor ... .cta_group::2).
//   empty[s]  one per CTA; the leader's tcgen05.commit multicasts the arrive to both.
//   tfull[a]  one per CTA (multicast commit); each CTA's epilogue drains its own 128 rows.
//   tempty[a] lives in the leader; one elected lane per epilogue warp of both CTAs arrives (8).
#include "ugp_common.cuh"
#include "conv_epilogue.cuh"
#include "conv_gemm.cuh"

namespace ugp {

constexpr int kPairN = 256;                          // N tile of the pair
constexpr int kPairStages = 5;
constexpr int kPairBHalfBytes = (kPairN / 2) * kBlockK * 2;          // 16 KB: this CTA's half of B
constexpr int kPairStageBytes = kATileBytes + kPairBHalfBytes;       // 32 KB per CTA per K block
constexpr int kPairStagingBufs = 2;
constexpr int kPairTmemCols = 2 * kPairN;                            // 512: two accumulator stages
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kPairStagingBufs * kOutChunkBytes +
                               5 * kMaxCout * 4 + kStatsScratchBytes + 256 + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO0,
                 const __grid_constant__ CUtensorMap tmO1, const ConvGemmParams p) {
  constexpr int kStages = kPairStages;
  constexpr int kMaxC = kMaxCout;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // [stage] 128 pixels x 64 channels
  uint8_t* sB = sA + kStages * kATileBytes;             // [stage] this CTA's 128 output channels x 64
  uint8_t* sOut = sB + kStages * kPairBHalfBytes;
  float* sStats = reinterpret_cast<float*>(sOut + kPairStagingBufs * kOutChunkBytes);  // [2][kMaxC]
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
  const uint32_t rank = cluster_ctarank();               // 0 = leader (issues the MMAs)
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int pair_m_tiles = p.m_tiles >> 1;
  const int num_tiles = pair_m_tiles * p.n_tiles;        // 256 x 256 tiles
  const int k_blocks = p.taps * (p.chunks0 + p.chunks1);

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO0);
    tma_prefetch_desc(&tmO1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);      // used in the leader only
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);    // used in the leader only: 4 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc_pair(tmem_slot, kPairTmemCols);
  pdl_wait();      // everything above overlaps the previous kernel's tail (programmatic dependent launch)
  if (warp_idx >= 2) {
    const int et = threadIdx.x - 64;
    for (int i = et; i < p.cout; i += 128) {
      sStats[i] = 0.f;
      sStats[kMaxC + i] = 0.f;
      sBias[i] = p.bias ? p.bias[i] : 0.f;
      sScale[i] = p.scale ? p.scale[i] : 1.f;
      sShift[i] = p.shift ? p.shift[i] : 0.f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                  // the peer's barriers exist and both TMEM halves are allocated
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform (uniform registers)

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const bool leader = elect_one() != 0;     // whole-warp loop, elected lane issues (uniform operands)
    {
      int stage = 0;
      uint32_t phase = 0;
      const int chunks = p.chunks0 + p.chunks1;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        const int pm = tile / p.n_tiles;
        const int n_tile = tile - pm * p.n_tiles;
        const int m_tile = 2 * pm + static_cast<int>(rank);
        const int tx = m_tile % p.tiles_x;
        const int t2 = m_tile / p.tiles_x;
        const int ty = t2 % p.tiles_y;
        const int b = t2 / p.tiles_y;
        const int x0 = tx * p.BW, y0 = ty * p.BH;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader && rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), 0);
          const int tap = kb / chunks;
          const int c = kb - tap * chunks;
          const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
          const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
          uint8_t* a_dst = sA + stage * kATileBytes;
          if (leader) {
            if (c < p.chunks0)
              tma_load_4d_pair(&tmA0, bar, a_dst, c * kBlockK, x0 + dx, y0 + dy, b);
            else
              tma_load_4d_pair(&tmA1, bar, a_dst, (c - p.chunks0) * kBlockK, x0 + dx, y0 + dy, b);
            tma_load_2d_pair(&tmW, bar, sB + stage * kPairBHalfBytes, c * kBlockK,
                             tap * p.cout + n_tile * kPairN + static_cast<int>(rank) * (kPairN / 2));
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      // whole-warp issue loop (uniform-datapath operands), one elected lane executes tcgen05 (conv_halo.cu)
      const bool leader = elect_one() != 0;
      constexpr uint32_t idesc = umma_idesc_bf16(2 * kBlockM, kPairN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * kPairN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          {
            const uint64_t a_desc = umma_desc_sw128(smem_u32(sA + stage * kATileBytes), 16, 1024);
            const uint64_t b_desc = umma_desc_sw128(smem_u32(sB + stage * kPairBHalfBytes), 16, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
              if (leader) umma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, acc);
            }
            if (leader) {
              umma_commit_pair(&empty_bar[stage], 3);
              if (kb == k_blocks - 1) umma_commit_pair(&tfull_bar[as], 3);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int et = threadIdx.x - 64;   // 0..127
    const int q = warp_idx & 3;        // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
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
    const uint32_t tempty_leader[2] = {mapa_u32(smem_u32(&tempty_bar[0]), 0),
                                       mapa_u32(smem_u32(&tempty_bar[1]), 0)};
    uint32_t chunk_counter = 0;
    int it = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++it) {
      if (tile + num_pairs >= num_tiles && et == 0) pdl_trigger();   // last tile of this CTA: the next kernel may be scheduled
      const int pm = tile / p.n_tiles;
      const int n_tile = tile - pm * p.n_tiles;
      const int m_tile = 2 * pm + static_cast<int>(rank);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after_sync();
#pragma unroll 1
      for (int cc = 0; cc < kPairN / 64; ++cc, ++chunk_counter) {
        uint32_t v[64];
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kPairN + cc * 64;
        tmem_ld_32x32(taddr, v);
        tmem_ld_32x32(taddr + 32, v + 32);
        tmem_ld_wait();
        if (cc == kPairN / 64 - 1) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader[as]);
        }
        const int n_base = n_tile * kPairN + cc * 64;
        const uint32_t buf_addr = smem_u32(sOut) + (chunk_counter % kPairStagingBufs) * kOutChunkBytes;
        if (n_base < p.split_n)
          epilogue_chunk<kPairStagingBufs>(ectx, v, n_base, buf_addr, row, et, &tmO0, n_base,
                                           m_tile * kBlockM);
        else
          epilogue_chunk<kPairStagingBufs>(ectx, v, n_base, buf_addr, row, et, &tmO1,
                                           n_base - p.split_n, m_tile * kBlockM);
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

  // nobody leaves (or frees TMEM) while the peer may still read this CTA's smem / signal its barriers
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp_idx == 1) {
    tc_fence_after_sync();
    tmem_dealloc_pair(tmem_base, kPairTmemCols);
  }
}

int g_conv_pair_mode = 3;   // bit 0: CTA pairs for the 256-wide generic tiles; bit 1: for the halo kernels

bool conv_pair_eligible(int m_tiles, int cout, int split_n) {
  return (g_conv_pair_mode & 1) != 0 && m_tiles % 2 == 0 && cout % kPairN == 0 && split_n % 64 == 0 &&
         cout <= kMaxCout;
}

int launch_conv_pair(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmW,
                     const CUtensorMap& tmO0, const CUtensorMap& tmO1, const ConvGemmParams& p,
                     cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kPairSmemBytes));
    configured = true;
  }
  const int tiles = (p.m_tiles / 2) * p.n_tiles;
  const int max_pairs = device_sm_count() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  UGP_CHECK_CUDA(launch_k(conv_pair_kernel, dim3(2 * pairs), dim3(192), kPairSmemBytes, stream, tmA0, tmA1, tmW, tmO0, tmO1, p));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
