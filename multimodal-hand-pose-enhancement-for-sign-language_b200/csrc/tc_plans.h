// Host-side launch plans of the tcgen05 kernels: encoded TMA tensor maps + tiling parameters.
// Built once per op (b2h_program_add) and replayed; one-shot calls build them per call.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "../../include/b2h_abi.h"

namespace b2h {

struct TcGemmParams {
  int B, Lo, Kc, stride;
  int tl, tb, n_lchunks;  // M tile = tb samples x tl rows (tl * tb = 128)
  int tl_log2;
  int ntaps;
  int tap_map[B2H_MAX_TAPS];    // 0: base / even-row view, 1: odd-row view
  int tap_coord[B2H_MAX_TAPS];  // row coordinate offset inside that view
  int tap_w[B2H_MAX_TAPS];      // tap index inside the packed weight (K offset = tap_w * Kc)
  // tap-merged main loop (stride 1, consecutive taps): tile rows ordered (row, sample), A box loaded once per chunk
  int merged, tb_log2, tap_lo, a_box_bytes;
};

struct alignas(64) TcGemmPlan {
  CUtensorMap tmA0, tmA1, tmB;
  CUtensorMap tmZ0, tmZ1;  // fuse_bwd: the producer layer's z (all rows / even rows, odd rows), not swizzled
  CUtensorMap tmO0, tmO1;  // persist: the output tensor (phase 0 / phase 1 rows) for the TMA stores of the epilogue
  TcGemmParams p;
  int BN, grid_x, grid_y;
  int epi;         // specialised epilogue id (EPI_*)
  int esz;         // operand element size: 2 = bf16 (kind::f16), 4 = fp32 (3xTF32)
  int pair;        // BN = 256 tiles run as CTA pairs (cta_group::2): tmB boxes are half tiles, grid_x is even
  int persist;     // several waves of BN = 256 tiles with a plain epilogue: the persistent kernel (k_gemm_persist.cu)
  int fuse_stats;  // BatchNorm statistics of the output produced by the epilogue
  int fuse_bwd;    // first pass of the producer's BatchNorm backward produced by the epilogue
  const float* bs_mean;
  const float* bs_invstd;
  double* bs_accum;
  int bs_C, bs_Cs, bs_groups, bs_up2, bs_zbytes, bs_pool2;
  const float* bs_scale;
  const float* bs_shift;
};

struct TcWgradParams {
  int Mpad, Npad, ntaps;
  int tl, tb, n_lchunks;  // k-block = tb samples x tl rows (tl * tb = 64)
  int total_kb, kb_per_split;
  int tap_map[B2H_MAX_TAPS];
  int tap_coord[B2H_MAX_TAPS];
  int direct, Mvalid, Nvalid;   // direct: one split, the epilogue writes dW[m][n][t] itself (no partial planes)
};

struct alignas(64) TcWgradPlan {
  CUtensorMap tmP, tmQ0, tmQ1;
  TcWgradParams p;
  int WN, grid_x, splits;
  int esz;
};

inline int wgrad_kblock_rows(int esz) { return esz == 4 ? 32 : 64; }
constexpr int kTf32MaxKbPerSplit = 24;   // 3xTF32 wgrad: at most 24 k-blocks (768 rows) per split-K slice
int make_map_3d(CUtensorMap* m, const void* base, int64_t C, int64_t L, int64_t B, int64_t row_pitch, int64_t sample_pitch,
                int box_c, int box_l, int box_b, int swizzle /* 0 none, 1 = 128B, 2 = 128B over 32-byte chunks */, int esz);
int make_map_2d(CUtensorMap* m, const void* base, int64_t K, int64_t N, int64_t row_pitch, int box_k, int box_n, int esz);
int plan_gemm_tc(const b2h_gemm_t& d, TcGemmPlan* plan, int esz);
int plan_wgrad_tc(const b2h_wgrad_t& d, TcWgradPlan* plan, int esz);
int64_t wgrad_tc_workspace_bytes(const b2h_wgrad_t& d, int esz);
// fp32 mode on the tensor cores (3xTF32, k_gemm_tf32.cu)
int run_gemm_tf32(const TcGemmPlan& plan, const b2h_gemm_t& d, cudaStream_t s);
int run_wgrad_tf32(const TcWgradPlan& plan, const b2h_wgrad_t& d, cudaStream_t s);
// persistent multi-tile form of the bf16 tap-GEMM (k_gemm_persist.cu)
bool persist_supports_epilogue(int kind);
int run_gemm_persist(const TcGemmPlan& plan, const b2h_gemm_t& d, int kind, cudaStream_t s);
int plan_gemm_bf16(const b2h_gemm_t& d, TcGemmPlan* plan);
int run_gemm_bf16(const TcGemmPlan& plan, const b2h_gemm_t& d, cudaStream_t s);
int plan_wgrad_bf16(const b2h_wgrad_t& d, TcWgradPlan* plan);
int run_wgrad_bf16(const TcWgradPlan& plan, const b2h_wgrad_t& d, cudaStream_t s);
int64_t wgrad_bf16_workspace_bytes(const b2h_wgrad_t& d);
int64_t wgrad_workspace_bytes(const b2h_wgrad_t& d, int dtype);
int64_t bn_partial_floats(int rows, int C, int groups);
int64_t l1_partial_floats(const b2h_l1_t& d);

}  // namespace b2h
