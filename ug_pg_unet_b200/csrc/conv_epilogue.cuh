// Shared epilogue of the tcgen05 convolution kernels: one 128-row x 64-column
// accumulator chunk (already in registers) -> +bias, optional affine + ReLU,
// bf16, swizzled smem staging, TMA store, and per-channel sum / sum-of-squares of
// the stored values (train-mode BatchNorm statistics, UG_unet_parts.py:11,14).
// Executed by the 128 epilogue threads (warps 2..5).
#pragma once

#include "ugp_common.cuh"

namespace ugp {

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_b32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void red_shared_add(uint32_t addr, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

constexpr int kStatsScratchBytes = 4 * 32 * 16;

struct EpilogueCtx {
  uint32_t bias_addr;    // smem byte addresses of the per-channel arrays
  uint32_t scale_addr;
  uint32_t shift_addr;
  uint32_t stats_addr;   // [2][stats_stride] floats: this CTA's running per-channel sum / sum of squares
  uint32_t stats_stride; // in floats
  uint32_t scratch_addr; // 2 KB: the four row-quarter partials of one chunk (kStatsScratchBytes)
  bool has_bias, has_affine, relu, do_stats;
  int valid_rows = 128;  // tile rows >= valid_rows hold no pixel (BW*BH < 128 boxes): stored / counted as zero
  // BatchNorm-backward statistics mode (the input gradient of a DoubleConv's second convolution IS the
  // activation gradient of its first BatchNorm): instead of sum / sum of squares of the stored values the
  // statistics pass accumulates  S1 = sum g*[a > 0]  and  S2 = sum g*[a > 0]*z  per channel, with z the
  // stored conv output of that BatchNorm (read here as bf16x2 words, row-major [pixels][C/2]) and
  // a = z*scale + shift (scale / shift in the smem arrays, NOT applied to the output).  This is the whole
  // reduction pass of ugp_bn_relu_bwd: that kernel then only runs its apply pass.
  const uint32_t* bwd_z = nullptr;
  int bwd_z_words = 0;   // words per pixel = C / 2
  bool whatif_no_store = false;   // what-if timing only (wrong results): the tile is staged but never stored
};

// v: 64 fp32 accumulator values of this thread's row (columns n_base .. n_base+63).
// buf_addr: smem byte address of the 16 KB staging buffer (1024-aligned, SWIZZLE_128B rows).
// kBwdZ = false compiles the BatchNorm-backward statistics mode out (kernels that never run it: fewer registers).
template <int kStagingBufs, bool kBwdZ = true>
__device__ __forceinline__ void epilogue_chunk(const EpilogueCtx& e, const uint32_t* v, int n_base,
                                               uint32_t buf_addr, int row, int et,
                                               const CUtensorMap* tm_out, int out_col, int out_row,
                                               int out_y = -1, int out_b = 0) {
  // BN-backward statistics: request this thread's 32 z words (one word column over a quarter of the rows)
  // now, so that they arrive while the accumulator is converted and staged
  uint32_t zw[kBwdZ ? 32 : 1];
  if (kBwdZ && e.bwd_z != nullptr) {
    const uint32_t* zp = e.bwd_z + (static_cast<size_t>(out_row) + (et >> 5) * 32) * e.bwd_z_words + (n_base >> 1) + (et & 31);
#pragma unroll
    for (int i = 0; i < 32; ++i) zw[i] = __ldg(zp + static_cast<size_t>(i) * e.bwd_z_words);
  }
  uint32_t packed[32];
  const bool live_row = row < e.valid_rows;
#pragma unroll
  for (int j = 0; j < 64; j += 4) {
    float f0 = __uint_as_float(v[j + 0]), f1 = __uint_as_float(v[j + 1]);
    float f2 = __uint_as_float(v[j + 2]), f3 = __uint_as_float(v[j + 3]);
    if (e.has_bias) {
      const float4 bi = ld_shared_f4(e.bias_addr + (n_base + j) * 4);
      f0 += bi.x; f1 += bi.y; f2 += bi.z; f3 += bi.w;
    }
    if (e.has_affine) {
      const float4 sc = ld_shared_f4(e.scale_addr + (n_base + j) * 4);
      const float4 sh = ld_shared_f4(e.shift_addr + (n_base + j) * 4);
      f0 = fmaf(f0, sc.x, sh.x); f1 = fmaf(f1, sc.y, sh.y);
      f2 = fmaf(f2, sc.z, sh.z); f3 = fmaf(f3, sc.w, sh.w);
    }
    if (e.relu) {
      f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f);
      f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f);
    }
    packed[j / 2] = live_row ? pack_bf16x2(f0, f1) : 0u;
    packed[j / 2 + 1] = live_row ? pack_bf16x2(f2, f3) : 0u;
  }
  // the TMA store that last read this staging buffer must be done (folding this wait into the
  // previous chunk's statistics barrier was measured slower: it puts the store's read latency on
  // every thread's path instead of hiding it behind the next chunk's TMEM load and conversion)
  if (et == 0) tma_store_wait_read<kStagingBufs - 1>();
  named_bar_sync(1, 128);
  const uint32_t rowp = buf_addr + row * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    st_shared_v4(rowp + ((j ^ (row & 7)) << 4), packed[4 * j], packed[4 * j + 1], packed[4 * j + 2],
                 packed[4 * j + 3]);
  fence_proxy_async_smem();
  named_bar_sync(2, 128);
  if (et == 0 && !e.whatif_no_store) {
    if (out_y < 0) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                       reinterpret_cast<uint64_t>(tm_out)),
                   "r"(buf_addr), "r"(out_col), "r"(out_row)
                   : "memory");
    } else {
      // strided 4-D destination (C, x, y, b): the tile's BW x BH pixel box of a ConvTranspose2d phase view
      asm volatile(
          "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
              reinterpret_cast<uint64_t>(tm_out)),
          "r"(buf_addr), "r"(out_col), "r"(out_row), "r"(out_y), "r"(out_b)
          : "memory");
    }
    tma_store_commit();
  }
  if (e.do_stats) {
    // column sums over the bf16-rounded tile: a thread owns one 32-bit word column (2 channels)
    // over a quarter of the rows; reads are conflict-free (32 lanes = one 128 B row)
    const int wcol = et & 31;
    const int r0 = (et >> 5) * 32;
    const uint32_t colp = buf_addr + ((wcol & 3) << 2);
    const int chunk = wcol >> 2;
    float s0a = 0.f, s1a = 0.f, q0a = 0.f, q1a = 0.f, s0b = 0.f, s1b = 0.f, q0b = 0.f, q1b = 0.f;
    if (!kBwdZ || e.bwd_z == nullptr) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {      // r0 is a multiple of 8: (r0 + i) & 7 == i & 7
        const uint32_t wa = ld_shared_b32(colp + (r0 + i) * 128 + ((chunk ^ (i & 7)) << 4));
        const uint32_t wb = ld_shared_b32(colp + (r0 + i + 1) * 128 + ((chunk ^ ((i + 1) & 7)) << 4));
        const float a0 = bf16_lo(wa), a1 = bf16_hi(wa), b0 = bf16_lo(wb), b1 = bf16_hi(wb);
        s0a += a0; s1a += a1; q0a = fmaf(a0, a0, q0a); q1a = fmaf(a1, a1, q1a);
        s0b += b0; s1b += b1; q0b = fmaf(b0, b0, q0b); q1b = fmaf(b1, b1, q1b);
      }
    } else {
      // BN-backward mode: g = the staged (bf16-rounded) gradient, masked by the ReLU of z*scale + shift
      float2 sc, sh;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(sc.x), "=f"(sc.y) : "r"(e.scale_addr + (n_base + 2 * wcol) * 4) : "memory");
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(sh.x), "=f"(sh.y) : "r"(e.shift_addr + (n_base + 2 * wcol) * 4) : "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const uint32_t wa = ld_shared_b32(colp + (r0 + i) * 128 + ((chunk ^ (i & 7)) << 4));
        const float z0 = bf16_lo(zw[kBwdZ ? i : 0]), z1 = bf16_hi(zw[kBwdZ ? i : 0]);
        const float g0 = (fmaf(z0, sc.x, sh.x) > 0.f) ? bf16_lo(wa) : 0.f;
        const float g1 = (fmaf(z1, sc.y, sh.y) > 0.f) ? bf16_hi(wa) : 0.f;
        if (i & 1) { s0b += g0; s1b += g1; q0b = fmaf(g0, z0, q0b); q1b = fmaf(g1, z1, q1b); }
        else       { s0a += g0; s1a += g1; q0a = fmaf(g0, z0, q0a); q1a = fmaf(g1, z1, q1a); }
      }
    }
    // Deterministic reduction (no atomics): the four row quarters meet in scratch and are summed in
    // a fixed order by the first warp, which owns channel pair (c, c+1) of the CTA's running sums.
    // Run-to-run identical statistics keep the whole bf16 forward bit-reproducible.
    st_shared_v4(e.scratch_addr + ((et >> 5) * 32 + wcol) * 16, __float_as_uint(s0a + s0b),
                 __float_as_uint(s1a + s1b), __float_as_uint(q0a + q0b), __float_as_uint(q1a + q1b));
    named_bar_sync(3, 128);
    if (et < 32) {
      const float4 p0 = ld_shared_f4(e.scratch_addr + (0 * 32 + wcol) * 16);
      const float4 p1 = ld_shared_f4(e.scratch_addr + (1 * 32 + wcol) * 16);
      const float4 p2 = ld_shared_f4(e.scratch_addr + (2 * 32 + wcol) * 16);
      const float4 p3 = ld_shared_f4(e.scratch_addr + (3 * 32 + wcol) * 16);
      const int c = n_base + 2 * wcol;
      const uint32_t sa = e.stats_addr + c * 4, qa = e.stats_addr + (e.stats_stride + c) * 4;
      float2 s, q;
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(s.x), "=f"(s.y) : "r"(sa) : "memory");
      asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(q.x), "=f"(q.y) : "r"(qa) : "memory");
      s.x += ((p0.x + p1.x) + p2.x) + p3.x;
      s.y += ((p0.y + p1.y) + p2.y) + p3.y;
      q.x += ((p0.z + p1.z) + p2.z) + p3.z;
      q.y += ((p0.w + p1.w) + p2.w) + p3.w;
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sa), "f"(s.x), "f"(s.y) : "memory");
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(qa), "f"(q.x), "f"(q.y) : "memory");
    }
  }
}

}  // namespace ugp
