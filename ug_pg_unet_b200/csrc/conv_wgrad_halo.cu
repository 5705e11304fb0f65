// Weight gradient for the wide layers with few output channels (W a multiple of
// 64, Cout <= 128): the shapes where conv_wgrad.cu is bound by L2 -> SM operand
// traffic because every tap re-fetches (a shifted copy of) the same pixels.
//
//   dWp[tap][ci][co] += sum_p X[p + shift(tap), ci] * dZ[p, co]
//
// A CTA owns one 64-channel input chunk, one 64-wide Cout tile and a range of
// 128-pixel K chunks (a 64-pixel segment of TWO image rows each).  Per K chunk ONE TMA box
// brings the 4 x 66-pixel halo of the segment (33 KB) plus the 2 x 64-pixel dZ box (16 KB);
// all nine taps are MN-major A operands read out of that halo: a tap is a
// start-address offset (+128 B per pixel, +66*128 B per row) and the two taps
// that share a 128-row accumulator are one "leading byte offset" apart, so no
// data is duplicated in shared memory.  Five accumulators (tap pairs 0/1, 2/3,
// 4/5, 6/7, 7/8) x 64 columns live in TMEM.  L2 bytes per 128 pixels: 49 KB instead
// of 176 KB with per-tap loads (66 KB with one-row chunks).
#include "ugp_common.cuh"

namespace ugp {

extern int g_conv_debug;
extern int g_wgrad_waves;
extern int g_wgrad_max_ctas;

struct WgradHaloParams {
  int chunks0, chunks1;
  int H, W;
  int n_tiles;
  int cout;
  int cin_total;
  int kchunks_total;      // B*H*W / 128
  int kchunks_per_split;
  float* dw;              // [9][cin_total][cout] fp32, accumulated
  int debug;
};

constexpr int kWhW = 66;                                   // 64 pixels + 1 halo pixel each side
constexpr int kWhRows = 2;                                 // image rows per K chunk
constexpr int kWhHaloBytes = (kWhRows + 2) * kWhW * 128;   // 33,792 B
constexpr int kWhHaloStride = 34 * 1024;                   // 1024-aligned slot
constexpr int kWhDzBytes = kWhRows * 64 * 128;             // 16 KB
constexpr int kWhStages = 4;
constexpr int kWhSmemBytes = kWhStages * (kWhHaloStride + kWhDzBytes) + 256 + 1024;

__device__ __forceinline__ void red_add_v4_h(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

__device__ __forceinline__ uint32_t tap_offset_bytes(int tap) {
  return static_cast<uint32_t>(((tap / 3) * kWhW + tap % 3) * 128);
}

__global__ void __launch_bounds__(192, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                       const __grid_constant__ CUtensorMap tmDZ, const WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sD = sX + kWhStages * kWhHaloStride;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sD + kWhStages * kWhDzBytes);
  uint64_t* empty_bar = full_bar + kWhStages;
  uint64_t* done_bar = empty_bar + kWhStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x / p.n_tiles;
  const int n_tile = blockIdx.x - chunk * p.n_tiles;
  const int kc_begin = blockIdx.y * p.kchunks_per_split;
  int kc_end = kc_begin + p.kchunks_per_split;
  if (kc_end > p.kchunks_total) kc_end = p.kchunks_total;
  const int n_k = kc_end - kc_begin;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmDZ);
    for (int i = 0; i < kWhStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp_idx == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform (uniform registers)

  if (warp_idx == 0) {
    const bool leader = elect_one() != 0;     // whole-warp loop, elected lane issues (uniform operands)
    {
      const int xsegs = p.W / 64, ypairs = p.H / kWhRows;
      int stage = 0;
      uint32_t phase = 0;
      for (int kc = kc_begin; kc < kc_end; ++kc) {
        // K chunk -> (image, row pair, 64-pixel segment)
        const int xs = kc % xsegs;
        const int t2 = kc / xsegs;
        const int b = t2 / ypairs;
        const int y0 = (t2 - b * ypairs) * kWhRows;
        const int x0 = xs * 64;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(&full_bar[stage], kWhHaloBytes + kWhDzBytes);
          if (chunk < p.chunks0)
            tma_load_4d(&tmX0, &full_bar[stage], sX + stage * kWhHaloStride, chunk * 64, x0 - 1, y0 - 1, b);
          else
            tma_load_4d(&tmX1, &full_bar[stage], sX + stage * kWhHaloStride, (chunk - p.chunks0) * 64,
                        x0 - 1, y0 - 1, b);
          tma_load_4d(&tmDZ, &full_bar[stage], sD + stage * kWhDzBytes, n_tile * 64, x0, y0, b);
        }
        if (++stage == kWhStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp_idx == 1) {
    // whole-warp issue loop (uniform-datapath operands), one elected lane executes tcgen05 (conv_halo.cu)
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
    // tap pairs per accumulator and the byte distance between the pair (descriptor LBO)
    const int first_tap[5] = {0, 2, 4, 6, 7};
    uint64_t a_tmpl[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const uint32_t lbo = tap_offset_bytes(first_tap[t] + 1) - tap_offset_bytes(first_tap[t]);
      a_tmpl[t] = umma_desc_sw128(tap_offset_bytes(first_tap[t]), lbo, 1024);
    }
    const uint64_t b_tmpl = umma_desc_sw128(0, 16, 1024);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < n_k; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after_sync();
      {
        const uint32_t x_addr = smem_u32(sX + stage * kWhHaloStride);
        const uint32_t d_addr = smem_u32(sD + stage * kWhDzBytes);
#pragma unroll
        for (int t = 0; t < 5; ++t) {
#pragma unroll
          for (int r = 0; r < kWhRows; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // 16 pixels (K) per MMA: halo row r (+66 pixels), 16-pixel group k; dZ row r (+8 KB)
              const uint64_t a_desc = a_tmpl[t] + ((x_addr + r * (kWhW * 128) + k * 2048) >> 4);
              const uint64_t b_desc = b_tmpl + ((d_addr + r * 8192 + k * 2048) >> 4);
              const uint32_t acc = (i > 0 || r > 0 || k > 0) ? 1u : 0u;
              if (leader && !(p.debug & 1)) umma_bf16(tmem_base + t * 64, a_desc, b_desc, idesc, acc);
            }
          }
        }
        if (leader) {
          umma_commit(&empty_bar[stage]);
          if (i == n_k - 1) umma_commit(done_bar);
        }
      }
      __syncwarp();
      if (++stage == kWhStages) { stage = 0; phase ^= 1; }
    }
  } else {
    const int q = warp_idx & 3;
    const int row = q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int t = 0; t < 5; ++t) {
      const int tap = (t < 4) ? 2 * t + (row >> 6) : 7 + (row >> 6);
      const bool live = (t < 4) || (row >= 64);        // rows 0..63 of the last tile repeat tap 7
      float* dst = p.dw + (static_cast<size_t>(tap) * p.cin_total + chunk * 64 + (row & 63)) * p.cout +
                   n_tile * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 64 + c * 32, v);
        tmem_ld_wait();
        if (live && !(p.debug & 2)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4_h(dst + c * 32 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                         __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
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

bool conv_wgrad_halo_eligible(int H, int W, int cout, int taps) {
  return taps == 9 && W % 64 == 0 && H % kWhRows == 0 && cout % 64 == 0 && cout <= 128;
}

int conv_wgrad_halo_bf16(const void* src0, int c0, const void* src1, int c1, const void* dz, int cout,
                         int B, int H, int W, float* dw_packed, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kWhSmemBytes));
    configured = true;
  }
  WgradHaloParams p;
  p.chunks0 = c0 / 64;
  p.chunks1 = c1 / 64;
  p.H = H; p.W = W;
  p.n_tiles = cout / 64;
  p.cout = cout;
  p.cin_total = c0 + c1;
  p.kchunks_total = static_cast<int>((static_cast<long long>(B) * H * W) / (64 * kWhRows));
  p.dw = dw_packed;
  p.debug = g_conv_debug;
  const int items = (p.chunks0 + p.chunks1) * p.n_tiles;
  int wave = device_sm_count();
  if (g_wgrad_max_ctas > 0 && g_wgrad_max_ctas < wave) wave = g_wgrad_max_ctas;
  int splits = g_wgrad_waves * wave / items;
  int max_splits = p.kchunks_total / 4;
  if (max_splits < 1) max_splits = 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.kchunks_per_split = (p.kchunks_total + splits - 1) / splits;
  splits = (p.kchunks_total + p.kchunks_per_split - 1) / p.kchunks_per_split;

  CUtensorMap tmX0, tmX1, tmDZ;
  int rc;
  if ((rc = make_tmap_nhwc_bf16(&tmX0, src0, B, H, W, c0, 64, kWhW, kWhRows + 2))) return rc;
  if (c1 > 0) {
    if ((rc = make_tmap_nhwc_bf16(&tmX1, src1, B, H, W, c1, 64, kWhW, kWhRows + 2))) return rc;
  } else {
    tmX1 = tmX0;
  }
  if ((rc = make_tmap_nhwc_bf16(&tmDZ, dz, B, H, W, cout, 64, 64, kWhRows))) return rc;
  dim3 grid(items, splits);
  conv_wgrad_halo_kernel<<<grid, 192, kWhSmemBytes, stream>>>(tmX0, tmX1, tmDZ, p);
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp
