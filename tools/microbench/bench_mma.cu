// Micro-benchmark of the tcgen05.mma issue / operand-fetch cost model: one thread
// issues a long chain of 128 x N x 16 (or, for a CTA pair, 256 x N x 16) bf16 MMAs whose operands are
// already resident in shared memory and reports SM cycles per MMA.  Not on the product path.
// Built as its own library (tools/microbench/libugp_microbench.so) by ug_pg_unet_b200/build.py; the product
// library and its public header do not contain it.
#include <stdarg.h>

#include "../../ug_pg_unet_b200/csrc/ugp_common.cuh"

namespace ugp {

struct MmaBenchParams {
  int n;        // MMA N
  int iters;    // MMAs issued = iters * 4 (four K = 16 steps of one 64-wide K block)
  int nacc;     // accumulators cycled through (1, 2 or 4; nacc * n <= 512)
  int a_tiles;  // distinct A tiles cycled through (1..4)
  int gap;      // issue gap (in ~100-cycle units) inserted every 24 MMAs
  int a_shift;  // A window start in 128-byte rows past the 1024-byte swizzle-atom boundary (0..7)
  long long* out;   // [gridDim.x] cycles
};

template <bool kPair>
__device__ __forceinline__ void mma_bench_body(const MmaBenchParams& p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                      // 4 x 16 KB (+ 1 KB so a shifted window stays inside)
  uint8_t* sB = smem + 4 * 16384 + 1024;   // up to 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < (4 * 16384 + 1024 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { if (kPair) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();
  tc_fence_after_sync();
  // broadcast from lane 0: tells the compiler the value is warp-uniform (uniform-register operand)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(kPair ? 256 : 128, p.n, 0, 0);
    long long t0 = 0, t1 = 0;
    uint32_t gapv = threadIdx.x;
    // the whole warp runs the issue loop (uniform control flow -> descriptors live in uniform
    // registers); one elected lane executes the tcgen05 instructions
    const bool leader = elect_one() != 0;
    {
      t0 = clock64();
      for (int i = 0; i < p.iters; ++i) {
        // nacc >= 8 encodes an issue gap: every 6 iterations (24 MMAs) the warp burns ~(nacc / 8) * 100
        // cycles of dependent shuffles before it issues again (is the gap hidden by queued MMAs?)
        if (p.gap > 0 && i % 6 == 0) {
          for (int gI = 0; gI < p.gap * 4; ++gI) gapv = __shfl_sync(0xffffffffu, gapv * 3 + 1, (gapv >> 3) & 31);
        }
        const uint64_t a_desc = umma_desc_sw128(smem_u32(sA + (i & (p.a_tiles - 1)) * 16384 + (p.a_shift & 7) * 128), 16, 1024);
        const uint64_t b_desc = umma_desc_sw128(smem_u32(sB), 16, 1024);
        const uint32_t d = tmem_base + (i & (p.nacc - 1)) * p.n;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // a_shift >= 8: switch the accumulator on EVERY MMA (k parity) instead of every 4 MMAs
          const uint32_t dk = (p.a_shift >= 8) ? tmem_base + (k & 1) * p.n : d;
          if (leader) {
            if (kPair) umma_bf16_pair(dk, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
            else umma_bf16(dk, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
          }
        }
      }
      if (leader) { if (kPair) umma_commit_pair(bar, 1); else umma_commit(bar); }
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (lane == 0) {
      t1 = clock64();
      p.out[blockIdx.x] = t1 - t0 + (gapv == 0x7fffffffu ? 1 : 0);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (kPair) cluster_sync_all();
  if (warp == 0) { tc_fence_after_sync(); if (kPair) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(const MmaBenchParams p) { mma_bench_body<false>(p); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_bench_pair_kernel(const MmaBenchParams p) {
  mma_bench_body<true>(p);
}

int bench_mma(int n, int pair, int iters, int nacc, int a_tiles, int a_shift, int ctas, long long* out_cycles,
              cudaStream_t stream) {
  UGP_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && nacc >= 1 && nacc * n <= 512 && a_tiles >= 1 && a_tiles <= 4 && a_shift >= 0 && a_shift < 4096 &&
                  iters > 0 && ctas > 0 && out_cycles,
              "bench_mma: bad arguments");
  constexpr int kSmem = 4 * 16384 + 1024 + 32768 + 1024 + 256;
  MmaBenchParams p{n, iters, nacc, a_tiles, a_shift >> 4, a_shift & 15, out_cycles};
  if (pair) {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(mma_bench_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    mma_bench_pair_kernel<<<2 * ((ctas + 1) / 2), 128, kSmem, stream>>>(p);
  } else {
    UGP_CHECK_CUDA(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    mma_bench_kernel<<<ctas, 128, kSmem, stream>>>(p);
  }
  UGP_CHECK_CUDA(cudaGetLastError());
  return UGP_OK;
}

}  // namespace ugp

namespace ugp {
// the microbenchmark library carries its own error sink (the product's lives in capi.cu)
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  fputc('\n', stderr);
  va_end(ap);
}
}  // namespace ugp

/* SM cycles for a chain of iters*4 tcgen05 MMAs of shape 128 x n x 16 (pair != 0: 256 x n x 16 on a CTA
 * pair, cta_group::2) with operands resident in shared memory, on `ctas` CTAs at once; out_cycles[cta]
 * (leaders only for pairs).  a_shift: start of the A window in 128-byte rows past a 1024-byte swizzle-atom
 * boundary (the halo kernels' shifted taps). */
extern "C" int ugp_bench_mma(int n, int pair, int iters, int nacc, int a_tiles, int a_shift, int ctas,
                             long long* out_cycles, void* stream) {
  return ugp::bench_mma(n, pair, iters, nacc, a_tiles, a_shift, ctas, out_cycles, static_cast<cudaStream_t>(stream));
}
