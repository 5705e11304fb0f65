// 3x3 (pad 1) convolution for the wide, thin layers (W a multiple of 128, few
// channels): the ones whose implicit GEMM is bound by L2 -> SM operand traffic,
// not by the tensor pipe (UG_unet_parts.py:10,13 at the 256^2 / 128^2 levels:
// inc, down1, up3, up4 and their input-gradients).
//
// Same math and epilogue as conv_gemm.cu, different operand staging:
//   * a CTA tile is 2 image rows x 128 pixels (two 128 x N accumulators);
//   * per 64-channel chunk ONE TMA box brings the 4 x 130-pixel halo of the
//     tile into shared memory (66.5 KB, SWIZZLE_128B).  All nine taps read
//     their shifted 128-row A operand out of that halo through the UMMA
//     descriptor start address (+128 B per pixel, +130*128 B per row): the
//     activation is fetched from L2 once per tile instead of nine times;
//   * the weight tile of a (chunk, tap) is shared by both accumulators.
// L2 -> SM bytes per MAC drop ~3x versus the per-tap TMA loads of conv_gemm.cu.
//
// Warp roles (224 threads): warp 0 TMA producer, warps 1 and 6 MMA issuers (one accumulator row
// each; warp 1 owns TMEM), warps 2..5 epilogue.
#include "ugp_common.cuh"
#include "conv_epilogue.cuh"

namespace ugp {

struct ConvHaloParams {
  int chunks0, chunks1;
  int H, W;
  int tiles_x, tiles_y;      // W/128, H/2
  int m_tiles, n_tiles;
  int cout;
  int split_n;
  int relu;
  int debug;                 // what-if timing only: 1 = skip MMA issue, 2 = skip epilogue body
  int resident;              // 1: all weights of the layer stay in shared memory for the whole kernel
  int b_region_bytes;        // shared memory reserved for weights (slots, or the resident set)
  int staging_bufs;          // epilogue staging buffers (2, or 1 when the resident weights need the room)
  int prefetch;              // L2-prefetch the halo of the tile `prefetch` tiles ahead (0 = off)
  const float* bias;
  const float* scale;
  const float* shift;
  float* stats;
  const uint32_t* bwd_z;     // BatchNorm-backward statistics mode (conv_epilogue.cuh)
};

extern int g_conv_debug;
extern int g_conv_halo_resident;
extern int g_conv_halo_prefetch;

constexpr int kHaloRows = 2;                               // image rows per CTA tile
constexpr int kHaloW = 130;                                // 128 pixels + 1 halo pixel each side
constexpr int kHaloBytes = (kHaloRows + 2) * kHaloW * 128;  // 66,560 B = 65 * 1024
constexpr int kHaloStages = 2;
constexpr int kHaloMaxCout = 256;

template <int kN, bool kPair = false>
struct HaloCfg {
  // One mbarrier wait of the single MMA-issuing thread costs ~100+ cycles, so a B slot must carry
  // enough tensor work to hide it: a slot holds kTapsPerSlot weight tiles (one kernel row of taps
  // for N = 64: 3 taps x 2 rows x 4 = 24 MMAs of 32 cycles per wait; N = 128: 8 MMAs of 64 cycles).
  static constexpr int kTapsPerSlot = (kN == 64) ? 3 : 1;
  // a pair's weight slots are half the size and their refill crosses SMs (longer round trip): twice the depth
  static constexpr int kBStages = ((kN == 64) ? 2 : 3) * (kPair ? 2 : 1);
  static constexpr int kRowsB = kPair ? kN / 2 : kN;      // weight rows (output channels) this CTA stages
  static constexpr int kTapBytes = kRowsB * 128;
  static constexpr int kBBytes = kTapsPerSlot * kTapBytes;
  // accumulator stages: all 512 TMEM columns (4 stages at N = 64, 2 at N = 128) so that the MMA warp can run
  // several short tiles ahead of the epilogue
  static constexpr int kAccStages = 512 / (kHaloRows * kN);
  static constexpr int kTmemCols = kAccStages * kHaloRows * kN;     // 512 (per CTA, also in a pair)
  static constexpr int kFixedBytes = kHaloStages * kHaloBytes + 5 * kHaloMaxCout * 4 + kStatsScratchBytes + 256 + 1024;
  static constexpr int kSmemBytes = kFixedBytes + kBStages * kBBytes + 2 * 16384;    // streaming weights
  static constexpr int kSmemLimit = 232448;                                            // 227 KB per CTA
  // plans the weight region: resident when the layer has one N tile and everything fits
  static void plan(ConvHaloParams* p, int chunks) {
    const int wbytes = chunks * 9 * kTapBytes;
    p->resident = 0; p->b_region_bytes = kBStages * kBBytes; p->staging_bufs = 2;
    if (p->n_tiles != 1) return;
    if (kFixedBytes + wbytes + 2 * 16384 <= kSmemLimit) { p->resident = 1; p->b_region_bytes = wbytes; }
    else if (kFixedBytes + wbytes + 16384 <= kSmemLimit) { p->resident = 1; p->b_region_bytes = wbytes; p->staging_bufs = 1; }
  }
  static int smem_bytes(const ConvHaloParams& p) { return kFixedBytes + p.b_region_bytes + p.staging_bufs * 16384; }
};

// kPair: the kernel is launched as clusters of 2 CTAs (cta_group::2).  CTA `rank` of pair `q` owns
// m tile 2q + rank (its own halo, accumulators and output rows) and stages rows
// [rank * kN/2, (rank + 1) * kN/2) of every weight tile; the leader (rank 0) issues one M = 256 MMA
// per (tap, k, row) that reads A and the two B halves from both CTAs' shared memory.  Per-CTA operand
// fetch per MMA drops from 4 KB + kN*32 B to 4 KB + kN*16 B.  Barrier protocol as in
// conv_pair.cu: hfull / bfull / tempty live in the leader, hempty / bempty / tfull in both CTAs.
template <int kN, bool kPair>
__device__ __forceinline__ void conv_halo_body(const CUtensorMap& tmA0, const CUtensorMap& tmA1,
                                               const CUtensorMap& tmW, const CUtensorMap& tmO0,
                                               const CUtensorMap& tmO1, const ConvHaloParams& p) {
  using Cfg = HaloCfg<kN, kPair>;
  constexpr int kBStages = Cfg::kBStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sHalo = smem;
  uint8_t* sB = sHalo + kHaloStages * kHaloBytes;
  uint8_t* sOut = sB + p.b_region_bytes;
  float* sStats = reinterpret_cast<float*>(sOut + p.staging_bufs * 16384);     // [2][kHaloMaxCout]
  float* sBias = sStats + 2 * kHaloMaxCout;
  float* sScale = sBias + kHaloMaxCout;
  float* sShift = sScale + kHaloMaxCout;
  uint8_t* sScratch = reinterpret_cast<uint8_t*>(sShift + kHaloMaxCout);   // kStatsScratchBytes
  uint64_t* hfull = reinterpret_cast<uint64_t*>(sScratch + kStatsScratchBytes);
  uint64_t* hempty = hfull + kHaloStages;
  uint64_t* bfull = hempty + kHaloStages;
  uint64_t* bempty = bfull + kBStages;
  uint64_t* tfull = bempty + kBStages;
  uint64_t* tempty = tfull + Cfg::kAccStages;
  uint64_t* wfull = tempty + Cfg::kAccStages;          // resident weights landed (leader's barrier in a pair)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  // Resident mode (one N tile, small Cin): the layer's weights are loaded once per CTA and stay in
  // shared memory; per tile only the halo is fetched.  On the 64 -> 64 layers the weights (74 KB) are
  // otherwise re-fetched for every 256-pixel tile and are half of the kernel's L2 traffic.
  const bool resident = p.resident != 0;
  // work items: (m tile, n tile) for single CTAs, (pair of m tiles, n tile) for pairs
  const int worker = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_workers = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int num_tiles = (kPair ? p.m_tiles >> 1 : p.m_tiles) * p.n_tiles;
  const int chunks = p.chunks0 + p.chunks1;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO0);
    tma_prefetch_desc(&tmO1);
    // consumer-side barriers take one tcgen05.commit from EACH of the two MMA-issuing warps
    for (int i = 0; i < kHaloStages; ++i) { mbar_init(&hfull[i], 1); mbar_init(&hempty[i], kHaloRows); }
    for (int i = 0; i < kBStages; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], kHaloRows); }
    for (int i = 0; i < Cfg::kAccStages; ++i) { mbar_init(&tfull[i], kHaloRows); mbar_init(&tempty[i], kPair ? 8 : 4); }   // one elected lane per epilogue warp (and CTA)
    mbar_init(wfull, 1);
    fence_mbar_init();
  }
  if (warp_idx == 1) {
    if (kPair) tmem_alloc_pair(tmem_slot, Cfg::kTmemCols);
    else tmem_alloc(tmem_slot, Cfg::kTmemCols);
  }
  pdl_wait();      // everything above overlaps the previous kernel's tail (programmatic dependent launch)
  if (warp_idx >= 2 && warp_idx < 6) {
    const int et = threadIdx.x - 64;
    for (int i = et; i < p.cout; i += 128) {
      sStats[i] = 0.f;
      sStats[kHaloMaxCout + i] = 0.f;
      sBias[i] = p.bias ? p.bias[i] : 0.f;
      sScale[i] = p.scale ? p.scale[i] : 1.f;
      sShift[i] = p.shift ? p.shift[i] : 0.f;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();       // the peer's barriers exist, both TMEM halves are allocated
  tc_fence_after_sync();
  // lane-0 broadcast: marks the value warp-uniform so the MMA operands stay in uniform registers
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one() != 0;     // whole-warp loop, elected lane issues (uniform operands)
    {
      int hs = 0, bs = 0;
      uint32_t hph = 0, bph = 0;
      if (resident && worker < num_tiles) {
        const uint32_t wbytes = static_cast<uint32_t>(chunks) * 9 * Cfg::kTapBytes;
        const uint32_t wbar = kPair ? mapa_u32(smem_u32(wfull), 0) : 0u;
        if (leader) {
          if (!kPair || rank == 0) mbar_arrive_expect_tx(wfull, (kPair ? 2 : 1) * wbytes);
          for (int c = 0; c < chunks; ++c)
            for (int tap = 0; tap < 9; ++tap) {
              uint8_t* dst = sB + (c * 9 + tap) * Cfg::kTapBytes;
              if (kPair) tma_load_2d_pair(&tmW, wbar, dst, c * 64, tap * p.cout + static_cast<int>(rank) * Cfg::kRowsB);
              else tma_load_2d(&tmW, wfull, dst, c * 64, tap * p.cout);
            }
        }
      }
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        const int m_tile = kPair ? 2 * (tile / p.n_tiles) + static_cast<int>(rank) : tile / p.n_tiles;
        const int n_tile = tile % p.n_tiles;
        const int tx = m_tile % p.tiles_x;
        const int t2 = m_tile / p.tiles_x;
        const int ty = t2 % p.tiles_y;
        const int b = t2 / p.tiles_y;
        const int x0 = tx * 128, y0 = ty * kHaloRows;
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(&hempty[hs], hph ^ 1);
          const CUtensorMap* tmA = (c < p.chunks0) ? &tmA0 : &tmA1;
          const int cc0 = ((c < p.chunks0) ? c : c - p.chunks0) * 64;
          if (kPair) {
            // the leader arms its barrier with the bytes of both CTAs; both CTAs' loads complete on it
            const uint32_t hbar = mapa_u32(smem_u32(&hfull[hs]), 0);
            if (leader) {
              if (rank == 0) mbar_arrive_expect_tx(&hfull[hs], 2 * kHaloBytes);
              // one box per halo row: several smaller TMA requests in flight instead of one 66 KB box
#pragma unroll
              for (int j = 0; j < kHaloRows + 2; ++j)
                tma_load_4d_pair(tmA, hbar, sHalo + hs * kHaloBytes + j * (kHaloW * 128), cc0, x0 - 1,
                                 y0 - 1 + j, b);
            }
          } else if (leader) {
            if (p.debug & 4) {            // what-if: no activation loads (stale smem), barrier only
              mbar_arrive(&hfull[hs]);
            } else {
              mbar_arrive_expect_tx(&hfull[hs], kHaloBytes);
#pragma unroll
              for (int j = 0; j < kHaloRows + 2; ++j)
                tma_load_4d(tmA, &hfull[hs], sHalo + hs * kHaloBytes + j * (kHaloW * 128), cc0, x0 - 1,
                            y0 - 1 + j, b);
            }
          }
          // the halo of the tile this CTA loads two tiles from now: DRAM -> L2 ahead of time, so that its
          // TMA load finds it in L2 (a DRAM round trip is longer than one short tile of MMAs)
          if (leader && p.prefetch) {
            const int ptile = tile + p.prefetch * num_workers;
            if (ptile < num_tiles) {
              const int pm = kPair ? 2 * (ptile / p.n_tiles) + static_cast<int>(rank) : ptile / p.n_tiles;
              const int ptx = pm % p.tiles_x, pt2 = pm / p.tiles_x;
              const int px0 = ptx * 128, py0 = (pt2 % p.tiles_y) * kHaloRows, pb = pt2 / p.tiles_y;
#pragma unroll
              for (int j = 0; j < kHaloRows + 2; ++j) tma_prefetch_4d(tmA, cc0, px0 - 1, py0 - 1 + j, pb);
            }
          }
          if (++hs == kHaloStages) { hs = 0; hph ^= 1; }
          for (int tap = 0; !resident && tap < 9; tap += Cfg::kTapsPerSlot) {
            mbar_wait(&bempty[bs], bph ^ 1);
            if (kPair) {
              const uint32_t bar = mapa_u32(smem_u32(&bfull[bs]), 0);
              if (leader) {
                if (rank == 0) mbar_arrive_expect_tx(&bfull[bs], 2 * Cfg::kBBytes);
#pragma unroll
                for (int t = 0; t < Cfg::kTapsPerSlot; ++t)
                  tma_load_2d_pair(&tmW, bar, sB + bs * Cfg::kBBytes + t * Cfg::kTapBytes, c * 64,
                                   (tap + t) * p.cout + n_tile * kN + static_cast<int>(rank) * Cfg::kRowsB);
              }
            } else if (leader) {
              mbar_arrive_expect_tx(&bfull[bs], Cfg::kBBytes);
#pragma unroll
              for (int t = 0; t < Cfg::kTapsPerSlot; ++t)
                tma_load_2d(&tmW, &bfull[bs], sB + bs * Cfg::kBBytes + t * Cfg::kTapBytes, c * 64,
                            (tap + t) * p.cout + n_tile * kN);
            }
            if (++bs == kBStages) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1 || warp_idx == 6) {
    // ===================== MMA issuers (the leader CTA of a pair) =====================
    // TWO issuing warps, one per accumulator row: the tensor pipe queues only ~2 MMAs, so with N = 64
    // (48 cycles per MMA) every barrier wait / descriptor setup of a single issuer is dead time
    // (tools/t_mma_gap.py: a gap between groups of 24 MMAs adds 1:1 to the group time).  The rows use
    // different accumulators, so the two streams need no ordering; each warp commits for itself.
    const int my_r = (warp_idx == 1) ? 0 : 1;
    // descriptor template: start address 0, SBO 1024, SWIZZLE_128B; a shifted A window only changes the
    // start-address field (absolute-address swizzle, base_offset stays 0).  All bases are computed ONCE,
    // broadcast from lane 0 (-> uniform registers) and combined with uniform integer offsets in the
    // loops: no shuffle / R2UR sits between the MMAs of a tile.
    const uint64_t desc0 = umma_desc_sw128(0, 16, 1024);
    const uint64_t a_stage0 = __shfl_sync(0xffffffffu, desc0 + ((smem_u32(sHalo) + my_r * (kHaloW * 128)) >> 4), 0);
    const uint64_t b_region0 = __shfl_sync(0xffffffffu, desc0 + (smem_u32(sB) >> 4), 0);
    const uint32_t d_row0 = __shfl_sync(0xffffffffu, tmem_base + my_r * kN, 0);
    // The whole warp runs the issue loop so that descriptors, TMEM addresses and barrier addresses are
    // computed on the uniform datapath; one elected lane executes the tcgen05 instructions.  (Issuing
    // from inside `if (lane == 0)` made ptxas wrap every MMA in an ELECT / R2UR waterfall: 87 cycles
    // per instruction instead of max(N/2, 32 + N/4), tools/t_mma.py.)
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = umma_idesc_bf16(kPair ? 256 : 128, kN, 0, 0);
    int hs = 0, bs = 0;
    uint32_t hph = 0, bph = 0;
    int it = 0;
    if (resident && rank == 0 && worker < num_tiles) mbar_wait(wfull, 0);
    for (int tile = worker; rank == 0 && tile < num_tiles; tile += num_workers, ++it) {
      const int as = it % Cfg::kAccStages;
      const uint32_t aphase = (it / Cfg::kAccStages) & 1;
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after_sync();
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&hfull[hs], hph);
        for (int tap = 0; tap < 9; tap += Cfg::kTapsPerSlot) {
          if (!resident) mbar_wait(&bfull[bs], bph);
          tc_fence_after_sync();
          {
            const int dy0 = tap / 3, dx0 = tap % 3;     // kTapsPerSlot = 3: dx0 = 0; = 1: t = 0
            const uint64_t b_base =
                b_region0 + static_cast<uint32_t>((resident ? (c * 9 + tap) * Cfg::kTapBytes : bs * Cfg::kBBytes) >> 4);
            const uint64_t a_base =
                a_stage0 + static_cast<uint32_t>((hs * kHaloBytes + (dy0 * kHaloW + dx0) * 128) >> 4);
            const uint32_t d_base = d_row0 + as * (kHaloRows * kN);
            const uint32_t acc0 = (c > 0 || tap > 0) ? 1u : 0u;
            if (leader && !(p.debug & 1)) {
#pragma unroll
              for (int t = 0; t < Cfg::kTapsPerSlot; ++t) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  // halo-relative window (+1 folded in): +8 descriptor units per pixel, +2 per K step
                  const uint64_t a_desc = a_base + (t * 8 + 2 * k);
                  const uint64_t b_desc = b_base + (t * (Cfg::kTapBytes >> 4) + 2 * k);
                  const uint32_t acc = (t > 0 || k > 0) ? 1u : acc0;
                  if (kPair) umma_bf16_pair(d_base, a_desc, b_desc, idesc, acc);
                  else umma_bf16(d_base, a_desc, b_desc, idesc, acc);
                }
              }
            }
            const bool last_tap = tap + Cfg::kTapsPerSlot >= 9;
            if (leader) {
              if (kPair) {
                if (!resident) umma_commit_pair(&bempty[bs], 3);
                if (last_tap) {
                  umma_commit_pair(&hempty[hs], 3);
                  if (c == chunks - 1) umma_commit_pair(&tfull[as], 3);
                }
              } else {
                if (!resident) umma_commit(&bempty[bs]);
                if (last_tap) {
                  umma_commit(&hempty[hs]);
                  if (c == chunks - 1) umma_commit(&tfull[as]);
                }
              }
            }
          }
          __syncwarp();
          if (!resident && ++bs == kBStages) { bs = 0; bph ^= 1; }
        }
        if (++hs == kHaloStages) { hs = 0; hph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
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
    ectx.stats_stride = kHaloMaxCout;
    ectx.has_bias = p.bias != nullptr;
    ectx.has_affine = p.scale != nullptr && p.bwd_z == nullptr;
    ectx.bwd_z = p.bwd_z;
    ectx.bwd_z_words = p.cout >> 1;
    ectx.relu = p.relu != 0;
    ectx.do_stats = do_stats;
    const uint32_t tempty_leader0 = kPair ? mapa_u32(smem_u32(&tempty[0]), 0) : 0u;
    auto release_accumulator = [&](int as) {
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_cluster(tempty_leader0 + as * 8);
        else mbar_arrive(&tempty[as]);
      }
    };
    uint32_t chunk_counter = 0;
    int it = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
      if (tile + num_workers >= num_tiles && et == 0) pdl_trigger();   // last tile of this CTA: the next kernel may be scheduled
      const int m_tile = kPair ? 2 * (tile / p.n_tiles) + static_cast<int>(rank) : tile / p.n_tiles;
      const int n_tile = tile % p.n_tiles;
      const int tx = m_tile % p.tiles_x;
      const int t2 = m_tile / p.tiles_x;
      const int ty = t2 % p.tiles_y;
      const int b = t2 / p.tiles_y;
      const int as = it % Cfg::kAccStages;
      const uint32_t aphase = (it / Cfg::kAccStages) & 1;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after_sync();
      if (p.debug & 2) {
        release_accumulator(as);
        continue;
      }
#pragma unroll 1
      for (int r = 0; r < kHaloRows; ++r) {
        const int out_row0 = ((b * p.H + ty * kHaloRows + r) * p.W) + tx * 128;   // pixel index
#pragma unroll 1
        for (int cc = 0; cc < kN / 64; ++cc, ++chunk_counter) {
          uint32_t v[64];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 as * (kHaloRows * kN) + r * kN + cc * 64;
          tmem_ld_32x32(taddr, v);
          tmem_ld_32x32(taddr + 32, v + 32);
          tmem_ld_wait();
          if (r == kHaloRows - 1 && cc == kN / 64 - 1) release_accumulator(as);
          const int n_base = n_tile * kN + cc * 64;
          const CUtensorMap* tmo = (n_base < p.split_n) ? &tmO0 : &tmO1;
          const int ocol = (n_base < p.split_n) ? n_base : n_base - p.split_n;
          if (p.staging_bufs == 2)
            epilogue_chunk<2>(ectx, v, n_base, smem_u32(sOut) + (chunk_counter & 1) * 16384, row, et, tmo, ocol,
                              out_row0);
          else
            epilogue_chunk<1>(ectx, v, n_base, smem_u32(sOut), row, et, tmo, ocol, out_row0);
        }
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
          atomicAdd(p.stats + p.cout + i, sStats[kHaloMaxCout + i]);
        }
      } else {
        for (int i = et; i < p.cout; i += 128) {
          // this CTA's slot of the [SMs][2][C] partial-statistics buffer (bn_finalize sums the slots in order)
          p.stats[(static_cast<size_t>(blockIdx.x) * 2 + 0) * p.cout + i] = sStats[i];
          p.stats[(static_cast<size_t>(blockIdx.x) * 2 + 1) * p.cout + i] = sStats[kHaloMaxCout + i];
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();       // the peer may still read this CTA's smem / signal its barriers
  if (warp_idx == 1) {
    tc_fence_after_sync();
    if (kPair) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int kN>
__global__ void __launch_bounds__(224, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO0,
                 const __grid_constant__ CUtensorMap tmO1, const ConvHaloParams p) {
  conv_halo_body<kN, false>(tmA0, tmA1, tmW, tmO0, tmO1, p);
}

template <int kN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(224, 1)
conv_halo_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                      const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO0,
                      const __grid_constant__ CUtensorMap tmO1, const ConvHaloParams p) {
  conv_halo_body<kN, true>(tmA0, tmA1, tmW, tmO0, tmO1, p);
}

extern int g_conv_pair_mode;   // conv_pair.cu
int g_conv_halo_resident = 1;  // 0 = always stream the weights through the slot ring
int g_conv_halo_prefetch = 0;  // tiles of L2 prefetch distance for the halo loads (measured: no gain, off)

template <int kN>
static int launch_halo_pair(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmW,
                            const CUtensorMap& tmO0, const CUtensorMap& tmO1, ConvHaloParams p,
                            cudaStream_t stream) {
  using Cfg = HaloCfg<kN, true>;
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_pair_kernel<kN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemLimit));
    configured = true;
  }
  Cfg::plan(&p, p.chunks0 + p.chunks1);
  if (g_conv_halo_resident == 0) { p.resident = 0; p.b_region_bytes = Cfg::kBStages * Cfg::kBBytes; p.staging_bufs = 2; }
  const int tiles = (p.m_tiles / 2) * p.n_tiles;
  const int max_pairs = device_sm_count() / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  UGP_CHECK_CUDA(launch_k(conv_halo_pair_kernel<kN>, dim3(2 * pairs), dim3(224), Cfg::smem_bytes(p), stream, tmA0, tmA1, tmW, tmO0, tmO1, p));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

template <int kN>
static int launch_halo(const CUtensorMap& tmA0, const CUtensorMap& tmA1, const CUtensorMap& tmW,
                       const CUtensorMap& tmO0, const CUtensorMap& tmO1, ConvHaloParams p,
                       cudaStream_t stream) {
  using Cfg = HaloCfg<kN>;
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(conv_halo_kernel<kN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemLimit));
    configured = true;
  }
  Cfg::plan(&p, p.chunks0 + p.chunks1);
  if (g_conv_halo_resident == 0) { p.resident = 0; p.b_region_bytes = Cfg::kBStages * Cfg::kBBytes; p.staging_bufs = 2; }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  UGP_CHECK_CUDA(launch_k(conv_halo_kernel<kN>, dim3(grid), dim3(224), Cfg::smem_bytes(p), stream, tmA0, tmA1, tmW, tmO0, tmO1, p));
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

// Returns 1 when the shape is one this kernel covers (the caller then must not run conv_gemm).
bool conv_halo_eligible(int H, int W, int cout, int split_n, int taps) {
  if (taps != 9 || W % 128 != 0 || H % kHaloRows != 0 || cout > kHaloMaxCout) return false;
  int kn = 128;
  while (kn > 64 && cout % kn != 0) kn >>= 1;   // the epilogue routes each 64-column chunk to dst0 / dst1
  return cout % kn == 0;
}

int conv_halo_bf16(const void* src0, int c0, const void* src1, int c1, int B, int H, int W,
                   const void* wpack, int cout, void* dst0, void* dst1, int split_n,
                   const float* bias, const float* scale, const float* shift, int relu, float* stats,
                   const void* bwd_z, cudaStream_t stream) {
  if (dst1 == nullptr) split_n = cout;
  int kn = 128;
  while (kn > 64 && cout % kn != 0) kn >>= 1;   // the epilogue routes each 64-column chunk to dst0 / dst1
  ConvHaloParams p;
  p.chunks0 = c0 / 64;
  p.chunks1 = c1 / 64;
  p.H = H; p.W = W;
  p.tiles_x = W / 128;
  p.tiles_y = H / kHaloRows;
  p.m_tiles = B * p.tiles_x * p.tiles_y;
  p.n_tiles = cout / kn;
  p.cout = cout;
  p.split_n = split_n;
  p.relu = relu;
  p.debug = g_conv_debug;
  p.prefetch = g_conv_halo_prefetch;
  p.bias = bias; p.scale = scale; p.shift = shift; p.stats = stats;
  p.bwd_z = static_cast<const uint32_t*>(bwd_z);

  CUtensorMap tmA0, tmA1, tmW, tmO0, tmO1;
  int rc;
  if ((rc = make_tmap_nhwc_bf16(&tmA0, src0, B, H, W, c0, 64, kHaloW, 1))) return rc;
  if (c1 > 0) {
    if ((rc = make_tmap_nhwc_bf16(&tmA1, src1, B, H, W, c1, 64, kHaloW, 1))) return rc;
  } else {
    tmA1 = tmA0;
  }
  // pairs pay a cross-SM handshake per tile: worth it from two 64-channel chunks per tile on (measured)
  const bool pair = (g_conv_pair_mode & 2) != 0 && p.m_tiles % 2 == 0 &&
                    (p.chunks0 + p.chunks1 >= 2 || (g_conv_pair_mode & 4) != 0);
  if ((rc = make_tmap_2d_bf16(&tmW, wpack, 9LL * cout, c0 + c1, 64, pair ? kn / 2 : kn))) return rc;
  const long long M = static_cast<long long>(B) * H * W;
  if ((rc = make_tmap_2d_bf16(&tmO0, dst0, M, split_n, 64, 128))) return rc;
  if (split_n < cout) {
    if ((rc = make_tmap_2d_bf16(&tmO1, dst1, M, cout - split_n, 64, 128))) return rc;
  } else {
    tmO1 = tmO0;
  }
  if (pair) {
    if (kn == 128) return launch_halo_pair<128>(tmA0, tmA1, tmW, tmO0, tmO1, p, stream);
    return launch_halo_pair<64>(tmA0, tmA1, tmW, tmO0, tmO1, p, stream);
  }
  if (kn == 128) return launch_halo<128>(tmA0, tmA1, tmW, tmO0, tmO1, p, stream);
  return launch_halo<64>(tmA0, tmA1, tmW, tmO0, tmO1, p, stream);
}

}  // namespace ugp
