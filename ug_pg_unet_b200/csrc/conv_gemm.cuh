// Parameters and tile constants shared by the tcgen05 implicit-GEMM convolution kernels
// (conv_gemm.cu: one CTA per 128-pixel tile; conv_pair.cu: CTA pairs, cta_group::2).
#pragma once

#include "ugp_common.cuh"

namespace ugp {

struct ConvGemmParams {
  int taps;
  int chunks0, chunks1;
  int H, W, BH, BW;
  int tiles_x, tiles_y;
  int m_tiles, n_tiles;
  int cout;
  int split_n;
  int relu;
  int bias_period;   // bias[n % bias_period] (= cout for a plain convolution)
  int out4d;         // destinations are 4-D views (C, x, y, b): ConvTranspose2d phases, non-contiguous tiles
  int valid_rows;    // BW * BH <= 128: accumulator rows that are real pixels (general image sizes)
  const float* bias;
  const float* scale;
  const float* shift;
  float* stats;
  const uint32_t* bwd_z;   // non-null: BatchNorm-backward statistics mode (conv_epilogue.cuh), stats = fp32 [2][cout] accumulated
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = 128 B rows
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kOutChunkBytes = kBlockM * 64 * 2;    // 16 KB (64 output channels)
constexpr int kMaxCout = 1024;

}  // namespace ugp
