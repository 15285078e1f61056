// fp32 precision mode of the tap-GEMM family on the Blackwell tensor cores: 3xTF32.
//
// The reference computes in fp32 (modelZoo.py / train_gan.py never autocast) and the parity bar of this mode is
// 1e-5, which a single TF32 pass (10 mantissa bits, 1.4e-3) cannot meet.  Every fp32 operand x is
// therefore split into two TF32 numbers
//     hi = x with its 13 low mantissa bits dropped   (what tcgen05.mma.kind::tf32 reads from the raw fp32 word),
//     lo = x - hi                                    (exact in fp32, <= 13 significant bits; its own TF32 rounding
//                                                     leaves 2^-21 |x|)
// and the product is accumulated in fp32 in TMEM as  a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  (the a_lo*b_lo term, 2^-20 of
// the product at most, is dropped): three MMAs of half the bf16 rate per 8-deep k step.
//
// Accumulators.  The tensor core adds into its fp32 accumulator with truncation: measured on the B200, one
// accumulator fed by all three MMAs loses ~2^-24 of its value per k step, systematically (1.4e-5 at K = 1792).  So the
// 512 TMEM columns hold 512/BN accumulators: the two correction products go to their own accumulator (their rounding
// is relative to a 2^-10 times smaller sum) and the a_hi*b_hi products rotate over the other 512/BN - 1, k-block by
// k-block; the epilogue adds them in fp32 with round-to-nearest.  Each accumulator then sees 1/(3 * (512/BN - 1)) of
// the adds, of a partial sum.
//
// Pipeline per CTA: warp 0 = TMA producer (the same 128B-swizzled fp32 boxes as the bf16 kernels, 32 channels = 128
// bytes per row), warps 2-9 = split stage (read the raw stage, write the `lo` copy next to it — an element-wise pass,
// so the swizzle is irrelevant — then fence.proxy.async + mbarrier arrive), warp 1 = MMA issuer, and after the main
// loop warps 2-9 run the epilogue (tcgen05.ld -> bias / activation / eval-BN / dropout -> staged rows -> 16-byte
// coalesced stores).  Tap-merged main loop as in k_gemm_tc.cu: the A box of a stride-1 convolution is loaded and
// split ONCE per 32-channel chunk and read by every tap through a shifted descriptor.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "gemm_epilogue.cuh"
#include "ptx_sm100.cuh"
#include "bn_finalize.cuh"
#include "tc_plans.h"
#include "gemm_tc_shared.cuh"

namespace b2h {

using namespace ptx;

constexpr int T32_BM = 128;
constexpr int T32_BK = 32;                 // fp32 elements = 128 bytes = one swizzle row
constexpr int T32_A_BYTES = T32_BM * 128;  // 16 KB
constexpr int T32_THREADS = 64 + 256;      // warp 0 TMA, warp 1 MMA, warps 2..9 split + epilogue
constexpr int T32_SPLITTERS = 256;

// lo word of the split, pre-rounded so that the MMA's truncation of its 13 low bits rounds to nearest
__device__ __forceinline__ uint32_t tf32_lo(uint32_t x) {
  const float lo = __uint_as_float(x) - __uint_as_float(x & 0xFFFFE000u);
  return __float_as_uint(lo) + 0x1000u;
}

// raw[0, bytes) -> lo[0, bytes), 16 bytes per thread and iteration (bytes is a multiple of 16)
__device__ __forceinline__ void split_block(const uint8_t* raw, uint8_t* lo, int bytes, int tid) {
  const uint4* src = reinterpret_cast<const uint4*>(raw);
  uint4* dst = reinterpret_cast<uint4*>(lo);
  const int n = bytes >> 4;
#pragma unroll 4
  for (int i = tid; i < n; i += T32_SPLITTERS) {
    const uint4 v = src[i];
    dst[i] = make_uint4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  }
}

// the split copy is written with generic-proxy stores and read by the MMA through the async proxy
__device__ __forceinline__ void split_done(uint64_t* bar, int lane) {
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

template <int BN>
struct Tf32Cfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int RAW_BYTES = T32_A_BYTES + B_BYTES;   // [A raw | B raw | A lo | B lo]
  static constexpr int STAGE_BYTES = 2 * RAW_BYTES;
  static constexpr int STAGES = (BN == 128) ? 3 : 4;         // 192 KB either way: one CTA per SM
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  // tap-merged: A ring of SA x [raw 24 KB | lo 24 KB], B ring of SB x [raw | lo]
  static constexpr int SA = 2;
  static constexpr int A_STAGE = 24 * 1024;
  static constexpr int SB = (BN == 128) ? 3 : 4;
  static constexpr int MERGED_BYTES = SA * 2 * A_STAGE + SB * 2 * B_BYTES;
  static constexpr int EPI_PITCH = BN * 4 + 16;
  static constexpr int EPI_BYTES = 128 * EPI_PITCH;
  static constexpr int MAIN_BYTES = PIPE_BYTES > MERGED_BYTES ? PIPE_BYTES : MERGED_BYTES;
  static_assert(EPI_BYTES <= MAIN_BYTES, "the epilogue tile reuses the pipeline buffers");
  static constexpr int SMEM_BYTES = MAIN_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * BN * 4 /*pivot | invstd*/;
  // STATS / BWDSUM: per-warp partial sums and the z tile live behind the epilogue tile in the (idle) pipeline buffers
  static_assert(((EPI_BYTES + 8 * BN * 8 + 127) & ~127) + 128 * BN * 4 <= MAIN_BYTES, "z tile placement");
  static constexpr int NACC = 512 / BN - 1;        // accumulators of the a_hi*b_hi products; + 1 for the corrections
  static constexpr int SMALL_COL = NACC * BN;
};

// the three MMAs of one 8-deep k step: corrections -> `small`, main product -> `main`
__device__ __forceinline__ void umma_3xtf32(uint32_t main, uint32_t small, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi,
                                            uint64_t b_lo, uint32_t idesc, bool main_acc, bool small_acc) {
  umma_tf32(small, a_lo, b_hi, idesc, small_acc);
  umma_tf32(small, a_hi, b_lo, idesc, 1);
  umma_tf32(main, a_hi, b_hi, idesc, main_acc);
}

// sum of the accumulators of 32 columns [c, c+32) of this thread's TMEM lane: mains in order, then the corrections
__device__ __forceinline__ void tmem_sum_accumulators(uint32_t lane_base, int c, int width, int n_main, int small_col,
                                                      uint32_t* acc) {
  tmem_ld_32x32(lane_base + (uint32_t)c, acc);
  tmem_ld_wait();
#pragma unroll 1
  for (int a = 1; a <= n_main; ++a) {
    uint32_t t[32];
    tmem_ld_32x32(lane_base + (uint32_t)((a < n_main ? a * width : small_col) + c), t);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(t[j]));
  }
}

// MODE (as in gemm_tc_kernel): STATS = the store phase also accumulates the train-mode BatchNorm statistics of the
// tile (shifted sums of the fp32 values as stored -> fp64 atomics -> the last CTA finalises); BWDSUM (dgrad) = it
// accumulates the first pass of the producer layer's BatchNorm backward against a TMA-staged fp32 tile of z.
template <int BN, bool MERGED, int MODE = MODE_PLAIN>
__global__ void __launch_bounds__(T32_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmZ0,
                 const __grid_constant__ CUtensorMap tmZ1, TcGemmParams p, EpiParams e, b2h_bn_stats_t st, BwdSumsDev bs) {
  constexpr bool STATS = MODE == MODE_STATS;
  constexpr bool BWDSUM = MODE == MODE_BWDSUM;
  using Cfg = Tf32Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // barriers: B ring (or the joint ring when not merged): full / ready / empty;  merged A ring: a_full / a_ready / a_empty
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::MAIN_BYTES);
  uint64_t* ready_bar = full_bar + 4;
  uint64_t* empty_bar = ready_bar + 4;
  uint64_t* a_full = empty_bar + 4;
  uint64_t* a_ready = a_full + 2;
  uint64_t* a_empty = a_ready + 2;
  uint64_t* tmem_full_bar = a_empty + 2;
  uint64_t* z_bar = tmem_full_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(z_bar + 1);
  int* s_flag = reinterpret_cast<int*>(tmem_ptr + 1);
  float* s_piv = reinterpret_cast<float*>(smem + Cfg::MAIN_BYTES + 256);   // STATS: pivot;  BWDSUM: mean
  float* s_isd = s_piv + BN;                                               // BWDSUM: invstd

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, nt = blockIdx.y;
  const int bt = mt / p.n_lchunks, lc = mt - bt * p.n_lchunks;
  const int b0 = bt * p.tb, l0 = lc * p.tl;
  const int n0 = nt * BN;
  const int kpt = p.Kc / T32_BK;
  const int nkb = p.ntaps * kpt;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&ready_bar[i], T32_SPLITTERS / 32);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_ready[i], T32_SPLITTERS / 32);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    if (BWDSUM) {
      prefetch_tmap(&tmZ0);
      prefetch_tmap(&tmZ1);
      mbar_init(z_bar, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr uint32_t idesc = idesc_tf32(T32_BM, BN, 0, 0);
  constexpr int NACC = Cfg::NACC;
  const uint32_t tmem_small = tmem_base + Cfg::SMALL_COL;

  if (warp == 0) {
    if (lane == 0) {
      if (MERGED) {
        uint8_t* ringB = smem + Cfg::SA * 2 * Cfg::A_STAGE;
        int ib = 0;
        for (int kc = 0; kc < kpt; ++kc) {
          const int sa = kc % Cfg::SA;
          mbar_wait(&a_empty[sa], ((kc / Cfg::SA) & 1) ^ 1);
          mbar_arrive_expect_tx(&a_full[sa], (uint32_t)p.a_box_bytes);
          tma_load_3d(smem + sa * 2 * Cfg::A_STAGE, &tmA0, &a_full[sa], kc * T32_BK, b0, l0 + p.tap_lo);
          for (int t = 0; t < p.ntaps; ++t, ++ib) {
            const int sb = ib % Cfg::SB;
            mbar_wait(&empty_bar[sb], ((ib / Cfg::SB) & 1) ^ 1);
            mbar_arrive_expect_tx(&full_bar[sb], Cfg::B_BYTES);
            tma_load_2d(ringB + sb * 2 * Cfg::B_BYTES, &tmB, &full_bar[sb], p.tap_w[t] * p.Kc + kc * T32_BK, n0);
          }
        }
      } else {
        for (int kb = 0; kb < nkb; ++kb) {
          const int stage = kb % STAGES;
          mbar_wait(&empty_bar[stage], ((kb / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::RAW_BYTES);
          const int t = kb / kpt, kc = kb - t * kpt;
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          tma_load_3d(sA, p.tap_map[t] ? &tmA1 : &tmA0, &full_bar[stage], kc * T32_BK, l0 + p.tap_coord[t], b0);
          tma_load_2d(sA + T32_A_BYTES, &tmB, &full_bar[stage], p.tap_w[t] * p.Kc + kc * T32_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      if (MERGED) {
        const uint32_t ringB = smem_u32(smem + Cfg::SA * 2 * Cfg::A_STAGE);
        int ib = 0;
        for (int kc = 0; kc < kpt; ++kc) {
          const int sa = kc % Cfg::SA;
          mbar_wait(&a_ready[sa], (kc / Cfg::SA) & 1);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + sa * 2 * Cfg::A_STAGE);
          for (int t = 0; t < p.ntaps; ++t, ++ib) {
            const int sb = ib % Cfg::SB;
            mbar_wait(&ready_bar[sb], (ib / Cfg::SB) & 1);
            tc_fence_after();
            const uint32_t sB = ringB + sb * 2 * Cfg::B_BYTES;
            // tap t reads rows [t*tb, t*tb + 128) of the A box (t*tb rows = whole 8-row swizzle atoms)
            const uint64_t a_hi = smem_desc_sw128(sA + (uint32_t)(t * p.tb) * 128u, 16, 1024);
            const uint64_t a_lo = smem_desc_sw128(sA + Cfg::A_STAGE + (uint32_t)(t * p.tb) * 128u, 16, 1024);
            const uint64_t b_hi = smem_desc_sw128(sB, 16, 1024);
            const uint64_t b_lo = smem_desc_sw128(sB + Cfg::B_BYTES, 16, 1024);
            const uint32_t tmem_main = tmem_base + (uint32_t)(ib % NACC) * BN;
#pragma unroll
            for (int k = 0; k < T32_BK / 8; ++k) {   // 8 tf32 = 32 bytes along K: +2 in the >>4 address field
              const uint64_t o = (uint64_t)(k * 2);
              umma_3xtf32(tmem_main, tmem_small, a_hi + o, a_lo + o, b_hi + o, b_lo + o, idesc, ib >= NACC || k != 0,
                          (ib | k) != 0);
            }
            umma_commit(&empty_bar[sb]);
          }
          umma_commit(&a_empty[sa]);
        }
      } else {
        for (int kb = 0; kb < nkb; ++kb) {
          const int stage = kb % STAGES;
          mbar_wait(&ready_bar[stage], (kb / STAGES) & 1);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t a_hi = smem_desc_sw128(sA, 16, 1024);
          const uint64_t b_hi = smem_desc_sw128(sA + T32_A_BYTES, 16, 1024);
          const uint64_t a_lo = smem_desc_sw128(sA + Cfg::RAW_BYTES, 16, 1024);
          const uint64_t b_lo = smem_desc_sw128(sA + Cfg::RAW_BYTES + T32_A_BYTES, 16, 1024);
          const uint32_t tmem_main = tmem_base + (uint32_t)(kb % NACC) * BN;
#pragma unroll
          for (int k = 0; k < T32_BK / 8; ++k) {
            const uint64_t o = (uint64_t)(k * 2);
            umma_3xtf32(tmem_main, tmem_small, a_hi + o, a_lo + o, b_hi + o, b_lo + o, idesc, kb >= NACC || k != 0,
                        (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
        }
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int et = threadIdx.x - 64;   // 0..255
    // ---- split stage
    if (MERGED) {
      uint8_t* ringB = smem + Cfg::SA * 2 * Cfg::A_STAGE;
      int ib = 0;
      for (int kc = 0; kc < kpt; ++kc) {
        const int sa = kc % Cfg::SA;
        mbar_wait(&a_full[sa], (kc / Cfg::SA) & 1);
        uint8_t* sA = smem + sa * 2 * Cfg::A_STAGE;
        split_block(sA, sA + Cfg::A_STAGE, p.a_box_bytes, et);
        split_done(&a_ready[sa], lane);
        for (int t = 0; t < p.ntaps; ++t, ++ib) {
          const int sb = ib % Cfg::SB;
          mbar_wait(&full_bar[sb], (ib / Cfg::SB) & 1);
          uint8_t* sB = ringB + sb * 2 * Cfg::B_BYTES;
          split_block(sB, sB + Cfg::B_BYTES, Cfg::B_BYTES, et);
          split_done(&ready_bar[sb], lane);
        }
      }
    } else {
      for (int kb = 0; kb < nkb; ++kb) {
        const int stage = kb % STAGES;
        mbar_wait(&full_bar[stage], (kb / STAGES) & 1);
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        split_block(sA, sA + Cfg::RAW_BYTES, Cfg::RAW_BYTES, et);
        split_done(&ready_bar[stage], lane);
      }
    }
    // ---- epilogue: TMEM sub-partition = warp % 4, two warps per sub-partition split the tile columns
    const int sub = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int ph = n0 / e.half;           // a tile never straddles a sub-pixel phase
    const int nn0 = n0 - ph * e.half;     // first channel (within the phase) of this tile
    const int valid_cols = min(BN, e.Nvalid - nn0);
    constexpr int pitch = Cfg::EPI_PITCH;
    uint8_t* stage = smem + (size_t)sub * 32 * pitch;
    // a CTA past the last group's clips (ragged tiles) has no valid row: clamp its group index
    const int grp = (STATS || BWDSUM) ? min(b0 / (p.B / (STATS ? st.groups : bs.groups)), (STATS ? st.groups : bs.groups) - 1) : 0;
    if (STATS || BWDSUM) {
      for (int i = et; i < BN; i += 256) {
        if (STATS) {
          s_piv[i] = (st.running_mean && nn0 + i < st.C) ? st.running_mean[nn0 + i] : 0.f;
        } else {
          const bool in = nn0 + i < bs.C;
          s_piv[i] = in ? bs.mean[grp * bs.Cs + nn0 + i] : 0.f;
          s_isd[i] = in ? bs.invstd[grp * bs.Cs + nn0 + i] : 0.f;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    // z tile (BWDSUM): behind the staging tile and the per-warp partials; the pipeline buffers are idle by then
    uint8_t* zs = smem + (((size_t)Cfg::EPI_BYTES + (size_t)8 * BN * 8 + 127) & ~(size_t)127);
    DropCtx drop;
    drop.init(e.drop, e.drop_C);
    {
      const int r = sub * 32 + lane;  // tile row == TMEM lane
      const int bi = MERGED ? (r & (p.tb - 1)) : (r >> p.tl_log2), li = MERGED ? (r >> p.tb_log2) : (r & (p.tl - 1));
      const int b = b0 + bi, lo = l0 + li;
      const int64_t grow = (int64_t)b * e.Lo_actual + (int64_t)lo * e.nphase + ph;
      const uint64_t drop_row_base = (uint64_t)grow * (uint64_t)e.drop_C;
      // rows outside the tensor are computed but never stored: they must not index the dropout mask either
      if (!((b < p.B) && (lo < p.Lo) && (lo * e.nphase + ph < e.Lo_actual))) drop.mode = B2H_DROP_NONE;
      uint8_t* my = stage + (size_t)lane * pitch;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      if (BWDSUM && et == 0) {   // every MMA has retired: the stage buffers are free
        mbar_arrive_expect_tx(z_bar, (uint32_t)bs.zbytes);
        if (MERGED)   // z maps of a merged plan are (C, B, L) too
          tma_load_3d(zs, ph ? &tmZ1 : &tmZ0, z_bar, nn0, b0, bs.up2 ? (l0 >> 1) : l0);
        else
          tma_load_3d(zs, ph ? &tmZ1 : &tmZ0, z_bar, nn0, bs.up2 ? (l0 >> 1) : l0, b0);
      }
      constexpr int CH = BN / 2;  // columns per epilogue warp
      const int n_main = min(NACC, nkb);
#pragma unroll 1
      for (int c = chalf * CH; c < (chalf + 1) * CH; c += 32) {
        uint32_t acc[32];
        tmem_sum_accumulators(tmem_base + ((uint32_t)(sub * 32) << 16), c, BN, n_main, Cfg::SMALL_COL, acc);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float v[8];
          epi_finish8(e, drop, drop_row_base, nn0 + c + j, acc + j, v);
          float4* dst = reinterpret_cast<float4*>(my + (size_t)(c + j) * 4);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
    asm volatile("bar.sync %0, 64;" ::"r"(2 + sub) : "memory");   // both warps of this sub-partition are done
    if (STATS || BWDSUM) {
      // lane <-> fixed 16-byte column chunk (4 channels); the lanes left over take further rows of the same iteration
      constexpr int CHUNKS = BN / 4;
      constexpr int LPR = CHUNKS < 32 ? CHUNKS : 32;
      constexpr int RPI = 32 / LPR;
      const int ch = lane % LPR, rsub = lane / LPR;
      const bool ch_ok = ch * 4 < valid_cols;   // a ragged last chunk is stored whole (zeros in the padding)
      float piv[4], sc[4], a1[4], a2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        piv[i] = s_piv[ch * 4 + i], a1[i] = 0.f, a2[i] = 0.f;
        sc[i] = BWDSUM ? s_isd[ch * 4 + i] : 1.f;
      }
      if (BWDSUM) mbar_wait(z_bar, 0);
      if (ch_ok) {
#pragma unroll 1
        for (int rr = chalf * 16 + rsub; rr < chalf * 16 + 16; rr += RPI) {
          const int r = sub * 32 + rr;
          const int bi = MERGED ? (r & (p.tb - 1)) : (r >> p.tl_log2), li = MERGED ? (r >> p.tb_log2) : (r & (p.tl - 1));
          const int b = b0 + bi, lo = l0 + li;
          const int ris = lo * e.nphase + ph;
          if (b >= p.B || lo >= p.Lo || ris >= e.Lo_actual) continue;
          const int64_t grow = (int64_t)b * e.Lo_actual + ris;
          float* gdst = reinterpret_cast<float*>(e.out) + (size_t)grow * e.ldo + e.out_coff + nn0;
          const float4 u = *reinterpret_cast<const float4*>(stage + (size_t)rr * pitch + ch * 16);
          *reinterpret_cast<float4*>(gdst + ch * 4) = u;
          const float w[4] = {u.x, u.y, u.z, u.w};
          if (STATS) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float d0 = w[i] - piv[i];
              a1[i] += d0;
              a2[i] = fmaf(d0, d0, a2[i]);
            }
          } else {
            const int zr = !bs.up2 ? r : (MERGED ? (li >> 1) * p.tb + bi : bi * (p.tl >> 1) + (li >> 1));
            const float4 zq = *reinterpret_cast<const float4*>(zs + ((size_t)zr * BN + ch * 4) * 4);
            const float zw[4] = {zq.x, zq.y, zq.z, zq.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float h = (zw[i] - piv[i]) * sc[i];
              a1[i] += w[i];
              a2[i] = fmaf(w[i], h, a2[i]);
            }
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], off);
          a2[i] += __shfl_xor_sync(0xffffffffu, a2[i], off);
        }
      }
      // per-warp partials -> fixed-order sum over the 8 epilogue warps -> fp64 atomics
      float2* s_part = reinterpret_cast<float2*>(smem + (size_t)Cfg::EPI_BYTES);   // [8][BN]
      if (rsub == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) s_part[(warp - 2) * BN + ch * 4 + i] = make_float2(a1[i], a2[i]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int c = et; c < valid_cols; c += 256) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) {
          const float2 v = s_part[w8 * BN + c];
          t1 += v.x, t2 += v.y;
        }
        if (STATS) {
          bn_stats_accumulate(st, blockIdx.x % kCopies, grp, nn0 + c, t1, t2);
        } else {
          double* a = bs.accum + (((int64_t)(blockIdx.x % B2H_BWD_COPIES) * bs.groups + grp) * bs.C + nn0 + c) * 2;
          atomicAdd(a + 0, (double)t1);
          atomicAdd(a + 1, (double)t2);
        }
      }
      if (STATS) {
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          const uint32_t t = atomicAdd(st.ticket, 1u);
          const int last = (t == gridDim.x * gridDim.y - 1u);
          if (last) *st.ticket = 0u;
          *s_flag = last;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (*s_flag) {
          __threadfence();
          bn_stats_finalize(st, et, 256);
        }
      }
    } else
    if (valid_cols > 0) {
      const int row_bytes = valid_cols * 4;
      const int full16 = row_bytes >> 4;
#pragma unroll 1
      for (int rr = chalf * 16; rr < chalf * 16 + 16; ++rr) {
        const int r = sub * 32 + rr;
        const int bi = MERGED ? (r & (p.tb - 1)) : (r >> p.tl_log2), li = MERGED ? (r >> p.tb_log2) : (r & (p.tl - 1));
        const int b = b0 + bi, lo = l0 + li;
        const int ris = lo * e.nphase + ph;
        if (b >= p.B || lo >= p.Lo || ris >= e.Lo_actual) continue;
        const int64_t grow = (int64_t)b * e.Lo_actual + ris;
        uint8_t* gdst = reinterpret_cast<uint8_t*>(e.out) + ((size_t)grow * e.ldo + e.out_coff + nn0) * 4;
        const uint8_t* src = stage + (size_t)rr * pitch;
        for (int ch = lane; ch < full16; ch += 32)
          *reinterpret_cast<uint4*>(gdst + ch * 16) = *reinterpret_cast<const uint4*>(src + ch * 16);
        for (int bo = full16 * 16 + lane * 4; bo < row_bytes; bo += 128)   // ragged tail
          *reinterpret_cast<uint32_t*>(gdst + bo) = *reinterpret_cast<const uint32_t*>(src + bo);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// wgrad: D[m][n] = sum_rows P[row][m] * Q[shift_t(row)][n]; both operands MN-major (32-channel slabs of 32 rows)
// ---------------------------------------------------------------------------------------------
constexpr int W32_BM = 128;
constexpr int W32_BK = 32;                  // rows per k-block
constexpr int W32_SLAB = W32_BK * 128;      // (32 rows x 32 channels) fp32 = 4 KB

template <int WN>
struct Wgrad32Cfg {
  static constexpr int A_BYTES = (W32_BM / 32) * W32_SLAB;   // 16 KB
  static constexpr int B_BYTES = (WN / 32) * W32_SLAB;
  static constexpr int RAW_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGE_BYTES = 2 * RAW_BYTES;
  static constexpr int STAGES = (WN == 128) ? 3 : 4;
  static constexpr int EPI_PITCH = WN * 4 + 16;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int EPI_BYTES = 128 * EPI_PITCH;
  static constexpr int MAIN_BYTES = PIPE_BYTES > EPI_BYTES ? PIPE_BYTES : EPI_BYTES;
  static constexpr int SMEM_BYTES = MAIN_BYTES + 1024 + 256;
  static constexpr int NACC = 512 / WN - 1;
  static constexpr int SMALL_COL = NACC * WN;
};

template <int WN>
__global__ void __launch_bounds__(T32_THREADS, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ0,
                  const __grid_constant__ CUtensorMap tmQ1, TcWgradParams p, float* __restrict__ partial) {
  using Cfg = Wgrad32Cfg<WN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::MAIN_BYTES);
  uint64_t* ready_bar = full_bar + 4;
  uint64_t* empty_bar = ready_bar + 4;
  uint64_t* tmem_full_bar = empty_bar + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.Npad / WN;
  const int m0 = (blockIdx.x / n_tiles) * W32_BM, n0 = (blockIdx.x % n_tiles) * WN;
  const int t = blockIdx.y, split = blockIdx.z;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.total_kb);
  const int nkb = kb_end - kb_begin;  // >= 1 by construction
  // Mpad may be 64: only the slabs that exist are loaded; the accumulator rows of the missing slabs hold garbage
  // that is never stored (rows of D are independent)
  const int a_slabs = min(W32_BM / 32, (p.Mpad - m0) / 32);
  const bool dead_tap = p.tap_map[t] < 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmP);
    prefetch_tmap(&tmQ0);
    prefetch_tmap(&tmQ1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&ready_bar[i], T32_SPLITTERS / 32);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (dead_tap) {
    if (warp >= 2 && warp < 6) {
      const int m = m0 + (warp & 3) * 32 + lane;
      if (p.direct) {
        if (m < p.Mvalid)
          for (int c = 0; c < WN && n0 + c < p.Nvalid; ++c) partial[((int64_t)m * p.Nvalid + n0 + c) * p.ntaps + t] = 0.f;
      } else if (m < p.Mpad) {
        float* dst = partial + (((int64_t)split * p.ntaps + t) * p.Mpad + m) * p.Npad + n0;
        for (int c = 0; c < WN; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* tq = p.tap_map[t] ? &tmQ1 : &tmQ0;
      for (int i = 0; i < nkb; ++i) {
        const int stage = i % STAGES;
        mbar_wait(&empty_bar[stage], ((i / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], a_slabs * W32_SLAB + Cfg::B_BYTES);
        const int kb = kb_begin + i;
        const int bt = kb / p.n_lchunks, lc = kb - bt * p.n_lchunks;
        const int b0 = bt * p.tb, r0 = lc * p.tl;
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
#pragma unroll
        for (int j = 0; j < W32_BM / 32; ++j)
          if (j < a_slabs) tma_load_3d(sA + j * W32_SLAB, &tmP, &full_bar[stage], m0 + j * 32, r0, b0);
#pragma unroll
        for (int j = 0; j < WN / 32; ++j)
          tma_load_3d(sB + j * W32_SLAB, tq, &full_bar[stage], n0 + j * 32, r0 + p.tap_coord[t], b0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32(W32_BM, WN, 1, 1);
      for (int i = 0; i < nkb; ++i) {
        const int stage = i % STAGES;
        mbar_wait(&ready_bar[stage], (i / STAGES) & 1);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        // MN-major fp32: 128B swizzle over 32-byte chunks, K atoms of 4 rows.  LBO = stride between 32-channel
        // slabs, SBO = stride between 4-row groups
        const uint64_t a_hi = smem_desc_sw128_base32(sA, W32_SLAB, 512);
        const uint64_t b_hi = smem_desc_sw128_base32(sA + Cfg::A_BYTES, W32_SLAB, 512);
        const uint64_t a_lo = smem_desc_sw128_base32(sA + Cfg::RAW_BYTES, W32_SLAB, 512);
        const uint64_t b_lo = smem_desc_sw128_base32(sA + Cfg::RAW_BYTES + Cfg::A_BYTES, W32_SLAB, 512);
        const uint32_t tmem_main = tmem_base + (uint32_t)(i % Cfg::NACC) * WN;
#pragma unroll
        for (int k = 0; k < W32_BK / 8; ++k) {   // 8 rows (K) = two 4-row atoms = 1024 bytes: +64 in the >>4 field
          const uint64_t o = (uint64_t)(k * 64);
          umma_3xtf32(tmem_main, tmem_base + Cfg::SMALL_COL, a_hi + o, a_lo + o, b_hi + o, b_lo + o, idesc,
                      i >= Cfg::NACC || k != 0, (i | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int et = threadIdx.x - 64;
    for (int i = 0; i < nkb; ++i) {
      const int stage = i % STAGES;
      mbar_wait(&full_bar[stage], (i / STAGES) & 1);
      uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
      // (slabs that were not loaded are split too: their garbage stays in accumulator rows that are never stored)
      split_block(sA, sA + Cfg::RAW_BYTES, Cfg::RAW_BYTES, et);
      split_done(&ready_bar[stage], lane);
    }
    // epilogue: thread == accumulator row -> padded smem row; then coalesced row copies
    const int sub = warp & 3;
    const int chalf = (warp - 2) >> 2;
    constexpr int pitch = Cfg::EPI_PITCH;
    uint8_t* stage = smem + (size_t)sub * 32 * pitch;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    {
      uint8_t* my = stage + (size_t)lane * pitch;
      constexpr int CH = WN / 2;
      const int n_main = min(Cfg::NACC, nkb);
#pragma unroll 1
      for (int c = chalf * CH; c < (chalf + 1) * CH; c += 32) {
        uint32_t v[32];
        tmem_sum_accumulators(tmem_base + ((uint32_t)(sub * 32) << 16), c, WN, n_main, Cfg::SMALL_COL, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(my + (size_t)(c + j) * 4) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    asm volatile("bar.sync %0, 64;" ::"r"(2 + sub) : "memory");
    if (p.direct) {
      // one split: this tile IS the gradient of tap t -> PyTorch layout dW[m][n][t] (`partial` is dW here)
#pragma unroll 1
      for (int rr = chalf * 16; rr < chalf * 16 + 16; ++rr) {
        const int m = m0 + sub * 32 + rr;
        if (m >= p.Mvalid) break;
        const float* src = reinterpret_cast<const float*>(stage + (size_t)rr * pitch);
        float* dst = partial + ((int64_t)m * p.Nvalid + n0) * p.ntaps + t;
        for (int n = lane; n < WN && n0 + n < p.Nvalid; n += 32) dst[(int64_t)n * p.ntaps] = src[n];
      }
    } else {
#pragma unroll 1
      for (int rr = chalf * 16; rr < chalf * 16 + 16; ++rr) {
        const int m = m0 + sub * 32 + rr;
        if (m >= p.Mpad) break;
        uint8_t* gdst = reinterpret_cast<uint8_t*>(partial + (((int64_t)split * p.ntaps + t) * p.Mpad + m) * p.Npad + n0);
        const uint8_t* src = stage + (size_t)rr * pitch;
        for (int ch = lane; ch < WN / 4; ch += 32)
          *reinterpret_cast<uint4*>(gdst + ch * 16) = *reinterpret_cast<const uint4*>(src + ch * 16);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------
template <int BN, bool MERGED, int MODE>
static int launch_tf32(const TcGemmPlan& plan, const EpiParams& e, cudaStream_t s, const b2h_bn_stats_t& st) {
  using Cfg = Tf32Cfg<BN>;
  B2H_CARVE(gemm_tf32_kernel<BN, MERGED, MODE>);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t er = cudaFuncSetAttribute(gemm_tf32_kernel<BN, MERGED, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::SMEM_BYTES);
    if (er != cudaSuccess) return cuda_fail(er, "gemm_tf32 smem attribute");
    attr_set = true;
  }
  BwdSumsDev bs;
  memset(&bs, 0, sizeof(bs));
  if (MODE == MODE_BWDSUM) {
    bs.mean = plan.bs_mean;
    bs.invstd = plan.bs_invstd;
    bs.accum = plan.bs_accum;
    bs.C = plan.bs_C;
    bs.Cs = plan.bs_Cs;
    bs.groups = plan.bs_groups;
    bs.up2 = plan.bs_up2;
    bs.zbytes = plan.bs_zbytes;
  }
  const bool z = MODE == MODE_BWDSUM;
  dim3 grid(plan.grid_x, plan.grid_y);
  launch(gemm_tf32_kernel<BN, MERGED, MODE>, grid, T32_THREADS, Cfg::SMEM_BYTES, s, plan.tmA0, plan.tmA1, plan.tmB,
         z ? plan.tmZ0 : plan.tmA0, z ? plan.tmZ1 : plan.tmA0, plan.p, e, st, bs);
  B2H_LAUNCH_CHECK("gemm_tf32");
  return B2H_OK;
}

template <int BN, int MODE>
static int launch_tf32_m(const TcGemmPlan& plan, const EpiParams& e, cudaStream_t s, const b2h_bn_stats_t& st) {
  return plan.p.merged ? launch_tf32<BN, true, MODE>(plan, e, s, st) : launch_tf32<BN, false, MODE>(plan, e, s, st);
}

template <int BN>
static int launch_tf32_mode(const TcGemmPlan& plan, const EpiParams& e, cudaStream_t s, const b2h_bn_stats_t& st) {
  if (plan.fuse_stats) return launch_tf32_m<BN, MODE_STATS>(plan, e, s, st);
  if (plan.fuse_bwd) return launch_tf32_m<BN, MODE_BWDSUM>(plan, e, s, st);
  return launch_tf32_m<BN, MODE_PLAIN>(plan, e, s, st);
}

int run_gemm_tf32(const TcGemmPlan& plan, const b2h_gemm_t& d, cudaStream_t s) {
  EpiParams e = make_epi(d);
  int rc = plan.BN == 128 ? launch_tf32_mode<128>(plan, e, s, d.stats) : launch_tf32_mode<64>(plan, e, s, d.stats);
  if (rc) return rc;
  // shapes the epilogue cannot cover: separate passes with the same results
  if (d.stats.z && !plan.fuse_stats) {
    B2H_CHECK_ARG(d.stats.z == d.out && d.out_coff == 0, B2H_ERR_ARG, "gemm: stats must describe the output tensor");
    rc = launch_bn_stats(d.stats, B2H_F32, s);
    if (rc) return rc;
  }
  if (d.bwd_sums.z && !plan.fuse_bwd) return launch_bwd_sums_separate(d, B2H_F32, s);
  return B2H_OK;
}

template <int WN>
static int launch_wg32(const TcWgradPlan& plan, float* partial, cudaStream_t s) {
  using Cfg = Wgrad32Cfg<WN>;
  B2H_CARVE(wgrad_tf32_kernel<WN>);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t er = cudaFuncSetAttribute(wgrad_tf32_kernel<WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (er != cudaSuccess) return cuda_fail(er, "wgrad_tf32 smem attribute");
    attr_set = true;
  }
  dim3 grid(plan.grid_x, plan.p.ntaps, plan.splits);
  launch(wgrad_tf32_kernel<WN>, grid, T32_THREADS, Cfg::SMEM_BYTES, s, plan.tmP, plan.tmQ0, plan.tmQ1, plan.p, partial);
  B2H_LAUNCH_CHECK("wgrad_tf32");
  return B2H_OK;
}

int run_wgrad_tf32(const TcWgradPlan& plan, const b2h_wgrad_t& d, cudaStream_t s) {
  float* out = plan.p.direct ? d.dW : d.partial;
  int rc = plan.WN == 128 ? launch_wg32<128>(plan, out, s) : launch_wg32<64>(plan, out, s);
  if (rc || plan.p.direct) return rc;
  return launch_wgrad_reduce(d, plan.splits, s);
}

}  // namespace b2h
