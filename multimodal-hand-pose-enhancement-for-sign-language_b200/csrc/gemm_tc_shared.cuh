// Definitions shared by the bf16 tcgen05 tap-GEMM kernels (k_gemm_tc.cu: one tile per CTA; k_gemm_persist.cu: the
// persistent multi-tile kernel): tile / ring configuration and the specialised epilogues.
#pragma once
#include "gemm_epilogue.cuh"
#include "ptx_sm100.cuh"
#include "tc_plans.h"

namespace b2h {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KB

// BN = 256: one CTA per SM, 4 stages.  BN <= 128: two co-resident CTAs per SM (<= 113 KB each) so that one
// CTA's epilogue overlaps the other's MMA main loop on the shared tensor core.
template <int BN>
struct FpropCfg {
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int OCC = (BN == 256) ? 1 : 2;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
  // epilogue staging (reuses the pipeline buffers): 128 rows x (BN * 4 bytes + 16)
  static constexpr int EPI_PITCH_MAX = BN * 4 + 16;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int EPI_BYTES = 128 * EPI_PITCH_MAX;
  static constexpr int MAIN_BYTES = PIPE_BYTES > EPI_BYTES ? PIPE_BYTES : EPI_BYTES;
  static constexpr int SMEM_BYTES = MAIN_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 4 * BN * 4 /*bias, pivot | scale, shift, pooled-BWDSUM scale*/;
  // tap-merged main loop: the A ring holds (tl + ntaps - 1) x tb rows of 64 channels once per channel chunk (all
  // taps read it through shifted descriptors), the B ring one (BN x 64) weight tile per (chunk, tap)
  static constexpr int SA = 2;
  static constexpr int A_STAGE = 24 * 1024;
  static constexpr int SB = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
  static_assert(SA * A_STAGE + SB * B_BYTES <= MAIN_BYTES, "merged rings must fit the pipeline buffers");
  // CTA pairs (BN = 256, opt-in B2H_PAIR=1): a CTA stages half of the B tile, so the same shared memory holds twice
  // as many k-blocks.  Measured on the B200 (profiles/pair_r02.md): parity-green, but 1-5 % SLOWER than single CTAs
  // with shallow and with deep rings alike -- the 128 x 256 tile is not bound by operand delivery (ncu: L2 at 17 %,
  // tensor pipe at 33 % of a multi-wave launch) but by what a CTA does around its main loop (launch, pipeline fill,
  // epilogue), which is what the persistent kernel of k_gemm_persist.cu removes.
  static constexpr int PAIR_B_BYTES = B_BYTES / 2;
  static constexpr int PAIR_STAGE_BYTES = TC_A_BYTES + PAIR_B_BYTES;
  static constexpr int PAIR_STAGES = 6;
  static constexpr int PAIR_SB = 8;
  static_assert(BN != 256 || (PAIR_STAGES * PAIR_STAGE_BYTES <= MAIN_BYTES &&
                              SA * A_STAGE + PAIR_SB * PAIR_B_BYTES <= MAIN_BYTES && PAIR_SB <= 8 && PAIR_STAGES <= 8),
                "pair rings must fit the pipeline buffers / barrier slots");
};

#ifndef B2H_PAIR_DEFAULT
#define B2H_PAIR_DEFAULT 0
#endif
constexpr int TC_THREADS = 64 + 256;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (2 per TMEM sub-partition)

// specialised epilogues (everything else falls back to EPI_GENERIC)
enum {
  EPI_GENERIC = 0, EPI_BIAS_LEAKY = 1, EPI_BIAS_RELU = 2, EPI_BIAS_F32 = 3, EPI_MASK = 4, EPI_PLAIN = 5,
  EPI_BIAS_LEAKY_BN = 6, EPI_BIAS_RELU_BN = 7   // + eval-mode BatchNorm folded to a per-channel affine
};
constexpr bool epi_has_bias(int k) {
  return k == EPI_BIAS_LEAKY || k == EPI_BIAS_RELU || k == EPI_BIAS_F32 || k == EPI_BIAS_LEAKY_BN || k == EPI_BIAS_RELU_BN;
}
constexpr bool epi_has_bn(int k) { return k == EPI_BIAS_LEAKY_BN || k == EPI_BIAS_RELU_BN; }

template <int KIND>
__device__ __forceinline__ void epi_fast8(const float* s_bias, const float* s_scale, const float* s_shift,
                                          const uint8_t* mask_row, int col, int nn, const uint32_t* acc_bits, float* v) {
  float4 b0 = make_float4(0, 0, 0, 0), b1 = b0;
  if (epi_has_bias(KIND)) {
    b0 = *reinterpret_cast<const float4*>(s_bias + col);
    b1 = *reinterpret_cast<const float4*>(s_bias + col + 4);
  }
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float sc[8], sh[8];
  if (epi_has_bn(KIND)) {
#pragma unroll
    for (int j = 0; j < 8; ++j) sc[j] = s_scale[col + j], sh[j] = s_shift[col + j];
  }
  uint32_t m0 = 0x01010101u, m1 = 0x01010101u;
  if (KIND == EPI_MASK) {
    m0 = *reinterpret_cast<const uint32_t*>(mask_row + nn);
    m1 = *reinterpret_cast<const uint32_t*>(mask_row + nn + 4);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float x = __uint_as_float(acc_bits[j]) + bb[j];
    if (KIND == EPI_BIAS_LEAKY || KIND == EPI_BIAS_LEAKY_BN) x = x > 0.f ? x : x * kLeakySlope;
    if (KIND == EPI_BIAS_RELU || KIND == EPI_BIAS_RELU_BN) x = x > 0.f ? x : 0.f;
    if (epi_has_bn(KIND)) x = fmaf(x, sc[j], sh[j]);
    if (KIND == EPI_MASK) {
      uint32_t byte = ((j < 4 ? m0 : m1) >> (8 * (j & 3))) & 0xFFu;
      x = byte ? 2.f * x : 0.f;
    }
    v[j] = x;
  }
}

// epilogue modes of the tap-GEMM kernels (k_gemm_tc.cu, k_gemm_tf32.cu): see gemm_tc_kernel
enum { MODE_PLAIN = 0, MODE_STATS = 1, MODE_BWDSUM = 2 };

struct BwdSumsDev {
  const float* mean;
  const float* invstd;
  double* accum;
  int C, Cs, groups;
  int up2;      // z row = tile row / 2 (x2 nearest up-sampling between the producer and this GEMM's rows)
  int zbytes;   // bytes of the z box
  int pool2;    // MaxPool1d(2) between the producer and this GEMM's rows: z rows 2l, 2l+1; the larger z*scale+shift counts
  const float* scale;
  const float* shift;
};

}  // namespace b2h
