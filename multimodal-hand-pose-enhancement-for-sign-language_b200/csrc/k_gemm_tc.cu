// bf16 precision mode of the tap-GEMM family on the Blackwell tensor cores (sm_100a):
//   * operands staged by TMA (cp.async.bulk.tensor) into 128B-swizzled shared memory; the temporal
//     taps of the convolution are row-shifted 3-D boxes (C, L, B) whose out-of-bounds rows are
//     zero filled by the TMA unit == the conv zero padding, per sample, with no im2col;
//   * tcgen05.mma (kind::f16, BF16 x BF16 -> FP32) issued by one thread, accumulators in TMEM;
//   * warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer / TMEM owner, warps 2-5 =
//     epilogue (tcgen05.ld -> bias / activation / eval-BN / dropout -> global);
//   * wgrad contracts over the row dimension, i.e. both operands are MN-major in shared memory:
//     the same TMA boxes, described to the MMA with MN-major descriptors (no transposes).
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

// ---------------------------------------------------------------------------------------------
// device: fprop-like kernel
// ---------------------------------------------------------------------------------------------
// STATS: the epilogue also produces the train-mode BatchNorm statistics of the tile it stores (per-column
// shifted sums of the bf16-rounded outputs -> fp64 atomics -> the last CTA finalises), see bn_finalize.cuh.
// BWDSUM (dgrad): the store phase also accumulates the first pass of the producer layer's BatchNorm backward,
// sum(g) and sum(g * zhat) per column, against a tile of the producer's z that one TMA brings into the idle
// pipeline buffers while the accumulator is being staged (b2h_bwd_sums_t).
// MERGED (stride-1 convolutions with consecutive taps): the M tile is ordered (row-in-sample, sample) with
// tb >= 8 samples, so the A box — tl + ntaps - 1 rows of tb samples, loaded ONCE per 64-channel chunk — serves
// every tap through a descriptor shifted by tap * tb rows (a multiple of the 8-row swizzle atom): A traffic and
// TMA issue drop by the tap count.
// PAIR (BN = 256 only, opt-in B2H_PAIR=1): CTA pairs (thread-block clusters of two neighbouring M tiles on one TPC) run
// ONE tcgen05.mma.cta_group::2 of M = 256: each CTA stages its own 128 rows of A and HALF of the B tile (128 of the
// 256 weight rows), the leader (cluster rank 0) issues the MMAs for both, each CTA's TMEM receives its own 128 x 256
// accumulator.  The TMA loads of both CTAs count their bytes on the leader's `full` barrier; the leader's
// tcgen05.commit multicasts to the `empty` / `tmem_full` barriers of both.  See gemm_tc_shared.cuh for what it measured.
// GADD (dgrad of a skip connection, b2h_gemm_t.grad_add, KIND = EPI_MASK): every epilogue thread requests its row of the
// other consumer's gradient BEFORE it waits for the accumulator -- the loads complete under the main loop, the column
// loop (unrolled: the prefetched chunks are named registers) adds them.  Without it (other kinds) the loads sit inside
// the column loop, one L2 round trip per 8 columns: measured +8 us per launch.
template <int BN, int KIND, int MODE, bool MERGED, bool PAIR = false, bool GADD = false>
__global__ void __launch_bounds__(TC_THREADS, FpropCfg<BN>::OCC)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmZ0,
               const __grid_constant__ CUtensorMap tmZ1, TcGemmParams p, EpiParams e, b2h_bn_stats_t st,
               BwdSumsDev bs) {
  constexpr bool STATS = MODE == MODE_STATS;
  constexpr bool BWDSUM = MODE == MODE_BWDSUM;
  using Cfg = FpropCfg<BN>;
  // joint (A + B) ring of the plain loop / B ring of the tap-merged loop: depth and strides
  constexpr int STAGES = PAIR ? Cfg::PAIR_STAGES : Cfg::STAGES;
  constexpr int STAGE_BYTES = PAIR ? Cfg::PAIR_STAGE_BYTES : Cfg::STAGE_BYTES;
  constexpr int SB = PAIR ? Cfg::PAIR_SB : Cfg::SB;
  constexpr int B_STRIDE = PAIR ? Cfg::PAIR_B_BYTES : Cfg::B_BYTES;
  constexpr int NBAR = STAGES > SB ? STAGES : SB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::MAIN_BYTES);
  uint64_t* empty_bar = full_bar + NBAR;
  uint64_t* tmem_full_bar = empty_bar + NBAR;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_bias = reinterpret_cast<float*>(smem + Cfg::MAIN_BYTES + 256);
  float* s_piv = s_bias + BN;     // STATS / BWDSUM: pivot / mean;  *_BN kinds: folded BN scale
  float* s_shift = s_piv + BN;    // *_BN kinds: folded BN shift;  pooled BWDSUM: the producer's shift
  float* s_pscale = s_shift + BN; // pooled BWDSUM: the producer's scale
  int* s_flag = reinterpret_cast<int*>(tmem_ptr + 1);
  uint64_t* z_bar = reinterpret_cast<uint64_t*>(tmem_ptr + 2);
  uint64_t* a_full = z_bar + 1;            // MERGED: the A ring (full_bar / empty_bar are the B ring)
  uint64_t* a_empty = a_full + Cfg::SA;

  static_assert(!PAIR || BN == 256, "CTA pairs run the 256-column tile");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, nt = blockIdx.y;
  const int bt = mt / p.n_lchunks, lc = mt - bt * p.n_lchunks;
  const int b0 = bt * p.tb, l0 = lc * p.tl;
  const int n0 = nt * BN;
  const int kpt = p.Kc / TC_BK;
  const int nkb = p.ntaps * kpt;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  constexpr int B_LOAD_BYTES = PAIR ? Cfg::B_BYTES / 2 : Cfg::B_BYTES;   // what THIS CTA loads of the B tile
  const int nB = n0 + (PAIR ? (int)cta_rank * (BN / 2) : 0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB);
    for (int i = 0; i < NBAR; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    if (MERGED) {
      for (int i = 0; i < Cfg::SA; ++i) {
        mbar_init(&a_full[i], 1);
        mbar_init(&a_empty[i], 1);
      }
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
    if (PAIR) {
      tmem_alloc2(tmem_ptr, BN);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_ptr, BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (PAIR)
    cluster_sync_all();   // the peer's TMA / commits must find this CTA's barriers initialised
  else
    __syncthreads();
  tc_fence_after();
  pdl_sync();  // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail
  const uint32_t tmem_base = *tmem_ptr;

  if (MERGED && warp == 0) {
    if (lane == 0) {
      uint8_t* ringB = smem + Cfg::SA * Cfg::A_STAGE;
      int ib = 0;
      for (int kc = 0; kc < kpt; ++kc) {
        const int sa = kc % Cfg::SA;
        mbar_wait(&a_empty[sa], ((kc / Cfg::SA) & 1) ^ 1);
        // tensor map dims (C, B, L): rows land as [row-in-sample][sample], zero outside [0, La)
        if (PAIR) {
          if (leader) mbar_arrive_expect_tx(&a_full[sa], 2u * (uint32_t)p.a_box_bytes);
          tma_load_3d_pair(smem + sa * Cfg::A_STAGE, &tmA0, &a_full[sa], kc * TC_BK, b0, l0 + p.tap_lo);
        } else {
          mbar_arrive_expect_tx(&a_full[sa], (uint32_t)p.a_box_bytes);
          tma_load_3d(smem + sa * Cfg::A_STAGE, &tmA0, &a_full[sa], kc * TC_BK, b0, l0 + p.tap_lo);
        }
        for (int t = 0; t < p.ntaps; ++t, ++ib) {
          const int sb = ib % SB;
          mbar_wait(&empty_bar[sb], ((ib / SB) & 1) ^ 1);
          if (PAIR) {
            if (leader) mbar_arrive_expect_tx(&full_bar[sb], Cfg::B_BYTES);   // both halves
            tma_load_2d_pair(ringB + sb * B_STRIDE, &tmB, &full_bar[sb], p.tap_w[t] * p.Kc + kc * TC_BK, nB);
          } else {
            mbar_arrive_expect_tx(&full_bar[sb], Cfg::B_BYTES);
            tma_load_2d(ringB + sb * B_STRIDE, &tmB, &full_bar[sb], p.tap_w[t] * p.Kc + kc * TC_BK, n0);
          }
        }
      }
    }
  } else if (MERGED && warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = idesc_bf16(PAIR ? 2 * TC_BM : TC_BM, BN, 0, 0);
      const uint32_t ringB = smem_u32(smem + Cfg::SA * Cfg::A_STAGE);
      int ib = 0;
      for (int kc = 0; kc < kpt; ++kc) {
        const int sa = kc % Cfg::SA;
        mbar_wait(&a_full[sa], (kc / Cfg::SA) & 1);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + sa * Cfg::A_STAGE);
        for (int t = 0; t < p.ntaps; ++t, ++ib) {
          const int sb = ib % SB;
          mbar_wait(&full_bar[sb], (ib / SB) & 1);
          tc_fence_after();
          // tap t reads the rows [t*tb, t*tb + 128) of the A box: t*tb rows = a whole number of 8-row atoms
          const uint64_t adesc = smem_desc_sw128(sA + (uint32_t)(t * p.tb) * 128u, 16, 1024);
          const uint64_t bdesc = smem_desc_sw128(ringB + sb * B_STRIDE, 16, 1024);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            if (PAIR)
              umma_bf16_pair(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | t | k) != 0);
            else
              umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | t | k) != 0);
          }
          if (PAIR) umma_commit_pair(&empty_bar[sb]); else umma_commit(&empty_bar[sb]);
        }
        if (PAIR) umma_commit_pair(&a_empty[sa]); else umma_commit(&a_empty[sa]);
      }
      if (PAIR) umma_commit_pair(tmem_full_bar); else umma_commit(tmem_full_bar);
    }
  } else if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int stage = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const int t = kb / kpt, kc = kb - t * kpt;
        uint8_t* sA = smem + stage * STAGE_BYTES;
        uint8_t* sB = sA + TC_A_BYTES;
        if (PAIR) {
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * TC_A_BYTES + Cfg::B_BYTES);   // A of both, B halves
          tma_load_3d_pair(sA, p.tap_map[t] ? &tmA1 : &tmA0, &full_bar[stage], kc * TC_BK, l0 + p.tap_coord[t], b0);
          tma_load_2d_pair(sB, &tmB, &full_bar[stage], p.tap_w[t] * p.Kc + kc * TC_BK, nB);
        } else {
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_3d(sA, p.tap_map[t] ? &tmA1 : &tmA0, &full_bar[stage], kc * TC_BK, l0 + p.tap_coord[t], b0);
          tma_load_2d(sB, &tmB, &full_bar[stage], p.tap_w[t] * p.Kc + kc * TC_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = idesc_bf16(PAIR ? 2 * TC_BM : TC_BM, BN, 0, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int stage = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t sB = sA + TC_A_BYTES;
        const uint64_t adesc = smem_desc_sw128(sA, 16, 1024);
        const uint64_t bdesc = smem_desc_sw128(sB, 16, 1024);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the >>4 address field
          if (PAIR)
            umma_bf16_pair(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          else
            umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        if (PAIR) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
      }
      if (PAIR) umma_commit_pair(tmem_full_bar); else umma_commit(tmem_full_bar);
    }
  } else {
    // epilogue warps 2..9: TMEM sub-partition = warp % 4 (hardware rule), two warps per sub-partition split
    // the tile columns.
    // stage 1: thread == tile row: tcgen05.ld -> bias/act/BN/dropout -> own row of a padded smem tile
    // stage 2: the two warps of a sub-partition copy its 32 rows out with row-contiguous 16-byte accesses
    const int sub = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int ph = n0 / e.half;           // a tile never straddles a sub-pixel phase
    const int nn0 = n0 - ph * e.half;     // first channel (within the phase) of this tile
    const int valid_cols = min(BN, e.Nvalid - nn0);
    const int esz = e.out_f32 ? 4 : 2;
    const int pitch = BN * esz + 16;
    uint8_t* stage = smem + (size_t)sub * 32 * pitch;
    const int et = threadIdx.x - 64;      // 0..255
    if (epi_has_bias(KIND)) {
      for (int i = et; i < BN; i += 256) {
        s_bias[i] = e.bias[nn0 + i];
        if (epi_has_bn(KIND)) s_piv[i] = e.post_scale[nn0 + i], s_shift[i] = e.post_shift[nn0 + i];
        if (STATS) s_piv[i] = (st.running_mean && nn0 + i < st.C) ? st.running_mean[nn0 + i] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    // (a pair's padding CTA past the last M tile has no valid row: its sums are zeros, its group index is clamped)
    const int grp = (STATS || BWDSUM) ? min(b0 / (p.B / (STATS ? st.groups : bs.groups)), (STATS ? st.groups : bs.groups) - 1) : 0;
    if (BWDSUM) {
      for (int i = et; i < BN; i += 256) {
        const bool in = nn0 + i < bs.C;
        s_piv[i] = in ? bs.mean[grp * bs.Cs + nn0 + i] : 0.f;
        s_bias[i] = in ? bs.invstd[grp * bs.Cs + nn0 + i] : 0.f;
        if (BN == 64 && bs.pool2) {
          s_pscale[i] = in ? bs.scale[grp * bs.Cs + nn0 + i] : 0.f;
          s_shift[i] = in ? bs.shift[grp * bs.Cs + nn0 + i] : 0.f;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    // z tile (BWDSUM): behind the staging tile and the per-warp partials; the pipeline buffers are idle by then
    uint8_t* zs = smem + (((size_t)128 * pitch + (size_t)8 * BN * 8 + 127) & ~(size_t)127);
    DropCtx drop;
    drop.init(e.drop, e.drop_C);
    {
      const int r = sub * 32 + lane;  // tile row == TMEM lane
      const int bi = MERGED ? (r & (p.tb - 1)) : (r >> p.tl_log2), li = MERGED ? (r >> p.tb_log2) : (r & (p.tl - 1));
      const int b = b0 + bi, lo = l0 + li;
      const int64_t grow = (int64_t)b * e.Lo_actual + (int64_t)lo * e.nphase + ph;
      const uint64_t drop_row_base = (uint64_t)grow * (uint64_t)e.drop_C;
      const bool row_in = (b < p.B) && (lo < p.Lo) && (lo * e.nphase + ph < e.Lo_actual);
      const uint8_t* mask_row = (KIND == EPI_MASK && row_in) ? e.drop.mask + drop_row_base : nullptr;
      if (KIND == EPI_GENERIC && !row_in) drop.mode = B2H_DROP_NONE;   // never index the mask with a row outside the tensor
      uint8_t* my = stage + (size_t)lane * pitch;
      constexpr int CH = BN / 2;  // columns per epilogue warp
      uint4 ga[GADD ? CH / 8 : 1];
      if (GADD) {
        const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(e.grad_add) + grow * e.ld_grad_add + nn0 + chalf * CH;
#pragma unroll
        for (int q = 0; q < CH / 8; ++q)
          ga[q] = row_in ? *reinterpret_cast<const uint4*>(gp + q * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      if (BWDSUM && et == 0) {   // every MMA has retired: the stage buffers are free
        mbar_arrive_expect_tx(z_bar, (uint32_t)bs.zbytes);
        const int zl0 = bs.up2 ? (l0 >> 1) : ((BN == 64 && bs.pool2) ? 2 * l0 : l0);   // first z frame of the box
        if (MERGED)   // z maps of a merged plan are (C, B, L) too
          tma_load_3d(zs, ph ? &tmZ1 : &tmZ0, z_bar, nn0, b0, zl0);
        else
          tma_load_3d(zs, ph ? &tmZ1 : &tmZ0, z_bar, nn0, zl0, b0);
      }
#pragma unroll(GADD ? CH / 32 : 1)
      for (int ci = 0; ci < CH / 32; ++ci) {
        const int c = chalf * CH + ci * 32;
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)c, acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float v[8];
          if (KIND == EPI_GENERIC) {
            epi_finish8(e, drop, drop_row_base, nn0 + c + j, acc + j, v);
          } else if (KIND == EPI_MASK) {
            if (row_in && nn0 + c + j + 8 <= e.drop_C) {
              epi_fast8<EPI_MASK>(s_bias, s_piv, s_shift, mask_row, c + j, nn0 + c + j, acc + j, v);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int cc = nn0 + c + j + k;
                const bool keep = row_in && cc < e.drop_C && mask_row[cc];
                v[k] = keep ? 2.f * __uint_as_float(acc[j + k]) : 0.f;
              }
            }
          } else {
            epi_fast8<KIND>(s_bias, s_piv, s_shift, nullptr, c + j, nn0 + c + j, acc + j, v);
          }
          if (GADD) {
            // skip connection: the other consumer's gradient of the same tensor (bf16 rows, zero in the channel padding)
            const uint4 q = ga[ci * 4 + j / 8];
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              v[2 * k] += __uint_as_float(w[k] << 16);
              v[2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u);
            }
          } else if ((KIND == EPI_MASK || KIND == EPI_PLAIN || KIND == EPI_GENERIC) && e.grad_add && row_in) {
            const F8 o = load8<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(e.grad_add) +
                                              grow * e.ld_grad_add + nn0 + c + j);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += o.v[k];
          }
          if (KIND == EPI_BIAS_F32 || (KIND == EPI_GENERIC && e.out_f32)) {
            float4* dst = reinterpret_cast<float4*>(my + (size_t)(c + j) * 4);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
          } else {
            __nv_bfloat162 q0 = __floats2bfloat162_rn(v[0], v[1]), q1 = __floats2bfloat162_rn(v[2], v[3]);
            __nv_bfloat162 q2 = __floats2bfloat162_rn(v[4], v[5]), q3 = __floats2bfloat162_rn(v[6], v[7]);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&q0);
            u.y = *reinterpret_cast<uint32_t*>(&q1);
            u.z = *reinterpret_cast<uint32_t*>(&q2);
            u.w = *reinterpret_cast<uint32_t*>(&q3);
            *reinterpret_cast<uint4*>(my + (size_t)(c + j) * 2) = u;
          }
        }
      }
    }
    // both warps of this sub-partition have written their column halves
    asm volatile("bar.sync %0, 64;" ::"r"(2 + sub) : "memory");
    if (STATS || BWDSUM) {
      // lane <-> fixed 16-byte column chunk; the lanes left over take further rows of the same iteration
      constexpr int CHUNKS = BN / 8;
      constexpr int LPR = CHUNKS < 32 ? CHUNKS : 32;
      constexpr int RPI = 32 / LPR;
      const int ch = lane % LPR, rsub = lane / LPR;
      const bool ch_ok = ch * 8 < valid_cols;   // a ragged last chunk is stored whole (zeros in the padding)
      // STATS: piv = pivot of the shifted sums.  BWDSUM: piv = mean, sc = invstd of the producer layer
      float piv[8], sc[8], a1[8], a2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        piv[i] = s_piv[ch * 8 + i], a1[i] = 0.f, a2[i] = 0.f;
        sc[i] = BWDSUM ? s_bias[ch * 8 + i] : 1.f;
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
          uint8_t* gdst = reinterpret_cast<uint8_t*>(e.out) + ((size_t)grow * e.ldo + e.out_coff + nn0) * 2;
          const uint4 u = *reinterpret_cast<const uint4*>(stage + (size_t)rr * pitch + ch * 16);
          *reinterpret_cast<uint4*>(gdst + ch * 16) = u;
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
          if (STATS) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
              const float d0 = f.x - piv[2 * i], d1 = f.y - piv[2 * i + 1];
              a1[2 * i] += d0, a1[2 * i + 1] += d1;
              a2[2 * i] = fmaf(d0, d0, a2[2 * i]), a2[2 * i + 1] = fmaf(d1, d1, a2[2 * i + 1]);
            }
          } else if (BN == 64 && bs.pool2) {   // (64-column tiles only: plan_gemm_tc)
            // MaxPool1d(2) between the producer and this GEMM: the gradient row belongs to the z row of the pair
            // (frames 2 li, 2 li + 1 of the box) with the larger z*scale + shift, the first one on ties -- the same
            // fmaf and comparison as bn_apply / bn_bwd
            const int zr0 = MERGED ? (2 * li) * p.tb + bi : bi * (2 * p.tl) + 2 * li;
            const int zr1 = MERGED ? zr0 + p.tb : zr0 + 1;
            const uint4 zq0 = *reinterpret_cast<const uint4*>(zs + ((size_t)zr0 * BN + ch * 8) * 2);
            const uint4 zq1 = *reinterpret_cast<const uint4*>(zs + ((size_t)zr1 * BN + ch * 8) * 2);
            const uint32_t zw0[4] = {zq0.x, zq0.y, zq0.z, zq0.w}, zw1[4] = {zq1.x, zq1.y, zq1.z, zq1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
              const float2 za = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zw0[i]));
              const float2 zb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zw1[i]));
              const float s0 = s_pscale[ch * 8 + 2 * i], s1 = s_pscale[ch * 8 + 2 * i + 1];
              const float t0 = s_shift[ch * 8 + 2 * i], t1 = s_shift[ch * 8 + 2 * i + 1];
              const float zx = fmaf(zb.x, s0, t0) > fmaf(za.x, s0, t0) ? zb.x : za.x;
              const float zy = fmaf(zb.y, s1, t1) > fmaf(za.y, s1, t1) ? zb.y : za.y;
              const float h0 = (zx - piv[2 * i]) * sc[2 * i], h1 = (zy - piv[2 * i + 1]) * sc[2 * i + 1];
              a1[2 * i] += f.x, a1[2 * i + 1] += f.y;
              a2[2 * i] = fmaf(f.x, h0, a2[2 * i]), a2[2 * i + 1] = fmaf(f.y, h1, a2[2 * i + 1]);
            }
          } else {
            const int zr = !bs.up2 ? r : (MERGED ? (li >> 1) * p.tb + bi : bi * (p.tl >> 1) + (li >> 1));
            const uint4 zq = *reinterpret_cast<const uint4*>(zs + ((size_t)zr * BN + ch * 8) * 2);
            const uint32_t zw[4] = {zq.x, zq.y, zq.z, zq.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
              const float2 zf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&zw[i]));
              const float h0 = (zf.x - piv[2 * i]) * sc[2 * i], h1 = (zf.y - piv[2 * i + 1]) * sc[2 * i + 1];
              a1[2 * i] += f.x, a1[2 * i + 1] += f.y;
              a2[2 * i] = fmaf(f.x, h0, a2[2 * i]), a2[2 * i + 1] = fmaf(f.y, h1, a2[2 * i + 1]);
            }
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], off);
          a2[i] += __shfl_xor_sync(0xffffffffu, a2[i], off);
        }
      }
      // per-warp partials -> fixed-order sum over the 8 epilogue warps -> fp64 atomics
      float2* s_part = reinterpret_cast<float2*>(smem + (size_t)128 * pitch);   // [8][BN], pipeline buffers are idle
      if (rsub == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_part[(warp - 2) * BN + ch * 8 + i] = make_float2(a1[i], a2[i]);
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
      const int row_bytes = valid_cols * esz;
      const int full16 = row_bytes >> 4;  // 16-byte chunks that are entirely valid
#pragma unroll 1
      for (int rr = chalf * 16; rr < chalf * 16 + 16; ++rr) {
        const int r = sub * 32 + rr;
        const int bi = MERGED ? (r & (p.tb - 1)) : (r >> p.tl_log2), li = MERGED ? (r >> p.tb_log2) : (r & (p.tl - 1));
        const int b = b0 + bi, lo = l0 + li;
        const int ris = lo * e.nphase + ph;
        if (b >= p.B || lo >= p.Lo || ris >= e.Lo_actual) continue;
        const int64_t grow = (int64_t)b * e.Lo_actual + ris;
        uint8_t* gdst = reinterpret_cast<uint8_t*>(e.out) + ((size_t)grow * e.ldo + e.out_coff + nn0) * esz;
        const uint8_t* src = stage + (size_t)rr * pitch;
        for (int ch = lane; ch < full16; ch += 32)
          *reinterpret_cast<uint4*>(gdst + ch * 16) = *reinterpret_cast<const uint4*>(src + ch * 16);
        // ragged tail (Nvalid not a multiple of the vector width): element-wise
        const int tail0 = full16 * 16;
        for (int bo = tail0 + lane * esz; bo < row_bytes; bo += 32 * esz) {
          if (esz == 4)
            *reinterpret_cast<uint32_t*>(gdst + bo) = *reinterpret_cast<const uint32_t*>(src + bo);
          else
            *reinterpret_cast<uint16_t*>(gdst + bo) = *reinterpret_cast<const uint16_t*>(src + bo);
        }
      }
    }
  }
  tc_fence_before();
  if (PAIR)
    cluster_sync_all();   // the leader's MMAs read the peer's shared memory; both TMEM halves are released together
  else
    __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, BN); else tmem_dealloc(tmem_base, BN);
  }
}

// ---------------------------------------------------------------------------------------------
// device: wgrad kernel.  D[m][n] = sum_rows P[row][m] * Q[shift_t(row)][n]; both operands MN-major.
// ---------------------------------------------------------------------------------------------
constexpr int WG_BM = 128;
constexpr int WG_BK = 64;  // rows per k-block
constexpr int WG_SLAB = WG_BK * 128;  // one (64 rows x 64 channels) box = 8 KB

template <int WN>
struct WgradCfg {
  static constexpr int A_BYTES = (WG_BM / 64) * WG_SLAB;  // 16 KB
  static constexpr int B_BYTES = (WN / 64) * WG_SLAB;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OCC = (WN == 256) ? 1 : 2;
  static constexpr int STAGES = (WN == 256) ? 4 : (WN == 128 ? 3 : 4);
  static constexpr int EPI_PITCH = WN * 4 + 16;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int EPI_BYTES = 128 * EPI_PITCH;
  static constexpr int MAIN_BYTES = PIPE_BYTES > EPI_BYTES ? PIPE_BYTES : EPI_BYTES;
  static constexpr int SMEM_BYTES = MAIN_BYTES + 1024 + 256;
};

template <int WN>
__global__ void __launch_bounds__(192, WgradCfg<WN>::OCC)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ0,
                const __grid_constant__ CUtensorMap tmQ1, TcWgradParams p, float* __restrict__ partial) {
  using Cfg = WgradCfg<WN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::MAIN_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.Npad / WN;
  const int m0 = (blockIdx.x / n_tiles) * WG_BM, n0 = (blockIdx.x % n_tiles) * WN;
  const int t = blockIdx.y, split = blockIdx.z;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(kb_begin + p.kb_per_split, p.total_kb);
  const int nkb = kb_end - kb_begin;  // >= 1 by construction
  // Mpad may be 64 (discriminator layers): only the slabs that exist are loaded; the accumulator rows
  // of the missing slab hold garbage that is never stored (rows of D are independent)
  const int a_slabs = min(WG_BM / 64, (p.Mpad - m0) / 64);
  // a tap whose strided row view of Q is empty (e.g. odd rows of a 1-row input) contributes zeros
  const bool dead_tap = p.tap_map[t] < 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmP);
    prefetch_tmap(&tmQ0);
    prefetch_tmap(&tmQ1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, WN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();  // barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail
  const uint32_t tmem_base = *tmem_ptr;

  if (dead_tap) {
    if (warp >= 2) {
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
        const uint32_t phase = (i / STAGES) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], a_slabs * WG_SLAB + Cfg::B_BYTES);
        const int kb = kb_begin + i;
        const int bt = kb / p.n_lchunks, lc = kb - bt * p.n_lchunks;
        const int b0 = bt * p.tb, r0 = lc * p.tl;
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
#pragma unroll
        for (int j = 0; j < WG_BM / 64; ++j)
          if (j < a_slabs) tma_load_3d(sA + j * WG_SLAB, &tmP, &full_bar[stage], m0 + j * 64, r0, b0);
#pragma unroll
        for (int j = 0; j < WN / 64; ++j)
          tma_load_3d(sB + j * WG_SLAB, tq, &full_bar[stage], n0 + j * 64, r0 + p.tap_coord[t], b0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(WG_BM, WN, 1, 1);
      for (int i = 0; i < nkb; ++i) {
        const int stage = i % STAGES;
        const uint32_t phase = (i / STAGES) & 1;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sB = sA + Cfg::A_BYTES;
        // MN-major, 128B swizzle: LBO = stride between 64-channel slabs, SBO = stride between 8-row groups
        const uint64_t adesc = smem_desc_sw128(sA, WG_SLAB, 1024);
        const uint64_t bdesc = smem_desc_sw128(sB, WG_SLAB, 1024);
#pragma unroll
        for (int k = 0; k < WG_BK / 16; ++k) {
          // advance 16 rows (K) = 2 swizzle atoms of 1024 bytes: +128 in the >>4 address field
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (i | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // stage 1: thread == accumulator row -> padded smem row; stage 2: coalesced row copies to the partial plane
    const int sub = warp & 3;
    constexpr int pitch = Cfg::EPI_PITCH;
    uint8_t* stage = smem + (size_t)sub * 32 * pitch;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    {
      uint8_t* my = stage + (size_t)lane * pitch;
#pragma unroll 1
      for (int c = 0; c < WN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(my + (size_t)(c + j) * 4) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    __syncwarp();
    if (p.direct) {
      // one split: this tile IS the gradient of tap t -> PyTorch layout dW[m][n][t] (`partial` is dW here)
#pragma unroll 1
      for (int rr = 0; rr < 32; ++rr) {
        const int m = m0 + sub * 32 + rr;
        if (m >= p.Mvalid) break;
        const float* src = reinterpret_cast<const float*>(stage + (size_t)rr * pitch);
        float* dst = partial + ((int64_t)m * p.Nvalid + n0) * p.ntaps + t;
        for (int n = lane; n < WN && n0 + n < p.Nvalid; n += 32) dst[(int64_t)n * p.ntaps] = src[n];
      }
    } else
#pragma unroll 1
    for (int rr = 0; rr < 32; ++rr) {
      const int m = m0 + sub * 32 + rr;
      if (m >= p.Mpad) break;
      uint8_t* gdst = reinterpret_cast<uint8_t*>(partial + (((int64_t)split * p.ntaps + t) * p.Mpad + m) * p.Npad + n0);
      const uint8_t* src = stage + (size_t)rr * pitch;
      for (int ch = lane; ch < WN / 4; ch += 32)
        *reinterpret_cast<uint4*>(gdst + ch * 16) = *reinterpret_cast<const uint4*>(src + ch * 16);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, WN);
  }
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps and launch plans
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// tensor map over a (C, L, B) view of a bf16 (esz = 2) or fp32 (esz = 4) tensor: element (c, l, b) at
// base + (b*sample_pitch + l*row_pitch + c) * esz bytes
int make_map_3d(CUtensorMap* m, const void* base, int64_t C, int64_t L, int64_t B, int64_t row_pitch,
                int64_t sample_pitch, int box_c, int box_l, int box_b, int swizzle, int esz) {
  EncodeTiledFn fn = get_encode_fn();
  B2H_CHECK_ARG(fn != nullptr, B2H_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)row_pitch * esz, (cuuint64_t)sample_pitch * esz};
  cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_l, (cuuint32_t)box_b};
  cuuint32_t estr[3] = {1, 1, 1};
  B2H_CHECK_ARG(((uintptr_t)base % 16) == 0 && strides[0] % 16 == 0 && strides[1] % 16 == 0, B2H_ERR_ALIGN,
                "tensor map: base/strides must be 16-byte aligned");
  CUresult r = fn(m, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                               : (swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2H_CHECK_ARG(r == CUDA_SUCCESS, B2H_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: %d (C=%lld L=%lld B=%lld)", (int)r,
                (long long)C, (long long)L, (long long)B);
  return B2H_OK;
}

int make_map_2d(CUtensorMap* m, const void* base, int64_t K, int64_t N, int64_t row_pitch, int box_k, int box_n, int esz) {
  EncodeTiledFn fn = get_encode_fn();
  B2H_CHECK_ARG(fn != nullptr, B2H_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)row_pitch * esz};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_n};
  cuuint32_t estr[2] = {1, 1};
  B2H_CHECK_ARG(((uintptr_t)base % 16) == 0 && strides[0] % 16 == 0, B2H_ERR_ALIGN, "tensor map: alignment");
  CUresult r = fn(m, esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2H_CHECK_ARG(r == CUDA_SUCCESS, B2H_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
  return B2H_OK;
}

// rows-per-sample chunk (power of two <= cap) that wastes the fewest padded rows; ties -> larger
static int choose_tl(int L, int cap) {
  int best = 1;
  int64_t best_pad = (int64_t)L;
  for (int tl = 1; tl <= cap; tl <<= 1) {
    int64_t pad = (int64_t)ceil_div(L, tl) * tl;
    if (pad <= best_pad) {
      best_pad = pad;
      best = tl;
    }
  }
  return best;
}

// (map index, coordinate offset) of the strided row  lo*stride + off  in the even/odd row views
static void tap_view(int stride, int off, int* map, int* coord) {
  if (stride == 1) {
    *map = 0;
    *coord = off;
  } else {
    int par = ((off % 2) + 2) % 2;
    *map = par;
    *coord = (off - par) / 2;
  }
}

// strided views of a [B][L][ld] tensor: view 0 = rows 0,2,4.. (or all rows if stride 1), view 1 = rows 1,3,5..
// (boxes of 128 bytes of channels: 64 bf16 / 32 fp32)
static int make_row_views(CUtensorMap* m0, CUtensorMap* m1, bool* has1, const void* base, int C, int L, int B, int ld,
                          int stride, int box_l, int box_b, int esz, int swz = 1) {
  const uint8_t* p = reinterpret_cast<const uint8_t*>(base);
  const int box_c = 128 / esz;
  int rc;
  if (stride == 1) {
    rc = make_map_3d(m0, p, C, L, B, ld, (int64_t)L * ld, box_c, box_l, box_b, swz, esz);
    if (rc) return rc;
    *m1 = *m0;
    *has1 = false;
    return B2H_OK;
  }
  int Le = (L + 1) / 2, Lod = L / 2;
  rc = make_map_3d(m0, p, C, Le, B, 2 * (int64_t)ld, (int64_t)L * ld, box_c, box_l, box_b, swz, esz);
  if (rc) return rc;
  if (Lod > 0) {
    rc = make_map_3d(m1, p + (size_t)ld * esz, C, Lod, B, 2 * (int64_t)ld, (int64_t)L * ld, box_c, box_l, box_b, swz, esz);
    if (rc) return rc;
    *has1 = true;
  } else {
    *m1 = *m0;
    *has1 = false;
  }
  return B2H_OK;
}

static int epi_kind(const b2h_gemm_t& d);

// B2H_PAIR=0 runs every tile as a single CTA (cta_group::1)
static bool pair_mode_enabled() {
  static const bool on = [] {
    const char* e = getenv("B2H_PAIR");
    return e ? atoi(e) != 0 : B2H_PAIR_DEFAULT != 0;
  }();
  return on;
}

int plan_gemm_bf16(const b2h_gemm_t& d, TcGemmPlan* plan) { return plan_gemm_tc(d, plan, 2); }

// esz = 2: bf16 operands (kind::f16);  esz = 4: fp32 operands, 3xTF32 (k_gemm_tf32.cu).  A k-block is 128 bytes of
// channels either way: 64 bf16 / 32 fp32.
int plan_gemm_tc(const b2h_gemm_t& d, TcGemmPlan* plan, int esz) {
  TcGemmParams& p = plan->p;
  const int bke = 128 / esz;
  plan->esz = esz;
  p.B = d.B;
  p.Lo = d.Lo;
  p.Kc = d.Kc;
  p.stride = d.stride;
  const bool ncl = d.out_f32 == 2;   // the op writes the (B, Nvalid, Lo_actual) fp32 NCL tensor itself
  B2H_CHECK_ARG(ncl || (d.ldo % (16 / esz) == 0 && d.out_coff % (16 / esz) == 0 && ((uintptr_t)d.out % 16) == 0 &&
                        (d.out_f32 == 0 || d.ldo % 4 == 0)),
                B2H_ERR_ALIGN, "gemm_tc: out/ldo/out_coff must allow 16-byte row stores (ldo=%d coff=%d)", d.ldo,
                d.out_coff);
  B2H_CHECK_ARG(!ncl || (esz == 2 && d.nphase == 1 && d.stride == 1 && d.out_coff == 0 && d.Npad % 256 == 0 &&
                         d.Lo_actual % 4 == 0 && d.Lo_actual == d.Lo && ((uintptr_t)d.out % 16) == 0 && d.bias &&
                         d.act == B2H_ACT_NONE && !d.post_scale && d.drop.mode == B2H_DROP_NONE && !d.stats.z &&
                         !d.bwd_sums.z),
                B2H_ERR_ARG, "gemm: NCL output (out_f32 = 2) needs bf16 operands, a stride-1 single-phase op with bias "
                "only, Npad %% 256 == 0 and Lo %% 4 == 0");
  int rc;
  // tap-merged main loop: stride 1, taps forming a run of consecutive row offsets, and a tile of tl <= 16 rows x
  // tb >= 8 samples whose A box (tl + ntaps - 1 rows) fits the A stage
  p.merged = 0;
  int order[B2H_MAX_TAPS];
  for (int t = 0; t < d.ntaps; ++t) order[t] = t;
  std::sort(order, order + d.ntaps, [&](int a, int b) { return d.tap_off[a] < d.tap_off[b]; });
  const bool pool2 = d.out_pool2 != 0;     // MaxPool1d(2) in the epilogue: the persistent 256-column kernel only
  const bool resid = d.resid != nullptr || pool2;   // residual add in the epilogue: likewise
  B2H_CHECK_ARG(!(pool2 && d.resid), B2H_ERR_ARG, "gemm: out_pool2 and resid exclude each other");
  B2H_CHECK_ARG(!resid || (esz == 2 && d.nphase == 1 && d.out_coff == 0 && d.Npad % 256 == 0 && !d.out_f32 &&
                           d.Lo_actual == d.Lo && (pool2 || (d.ld_resid % 8 == 0 && ((uintptr_t)d.resid % 16) == 0)) &&
                           d.drop.mode == B2H_DROP_NONE && !d.stats.z && !d.bwd_sums.z),
                B2H_ERR_ARG, "gemm: a residual needs bf16 operands, one phase, Npad %% 256 == 0, no statistics / dropout");
  B2H_CHECK_ARG(!d.grad_add || (esz == 2 && !d.out_f32 && !resid && !ncl && d.out_coff == 0 && !d.bias && !d.post_scale &&
                                d.act == B2H_ACT_NONE && !d.stats.z && d.ld_grad_add % 8 == 0 &&
                                d.ld_grad_add >= d.Npad / d.nphase && ((uintptr_t)d.grad_add % 16) == 0),
                B2H_ERR_ARG, "gemm: grad_add is for bf16 dgrad ops (no bias / activation / statistics, out_coff = 0, "
                "16-byte aligned rows of at least Npad / nphase channels)");
  bool run = d.stride == 1 && d.ntaps >= 2 && !ncl && !getenv("B2H_NO_TAP_MERGE");
  for (int t = 1; t < d.ntaps && run; ++t) run = d.tap_off[order[t]] == d.tap_off[order[t - 1]] + 1;
  if (run) {
    const int h = d.ntaps - 1;
    int best = 0;
    int64_t best_pad = 0;
    for (int tl = 1; tl <= 16; tl <<= 1) {
      if ((tl + h) * (TC_BM / tl) * 128 > FpropCfg<64>::A_STAGE) continue;
      const int64_t pad = (int64_t)ceil_div(d.Lo, tl) * tl;
      if (!best || pad <= best_pad) best = tl, best_pad = pad;
    }
    // (do not trade more than 1/8 of padded rows for the merged loop)
    if (best && best_pad * 8 <= (int64_t)ceil_div(d.Lo, choose_tl(d.Lo, TC_BM)) * choose_tl(d.Lo, TC_BM) * 9) {
      p.merged = 1;
      p.tl = best;
      p.tb = TC_BM / best;
      p.tap_lo = d.tap_off[order[0]];
      p.a_box_bytes = (best + h) * p.tb * 128;
      p.ntaps = d.ntaps;
      for (int t = 0; t < d.ntaps; ++t) p.tap_w[t] = order[t], p.tap_map[t] = 0, p.tap_coord[t] = d.tap_off[order[t]];
      // dims (C, B, L): the box lands as [row][sample][64 channels]
      rc = make_map_3d(&plan->tmA0, d.A, d.Kc, d.B, d.La, (int64_t)d.La * d.lda, d.lda, bke, p.tb, best + h, true, esz);
      if (rc) return rc;
      plan->tmA1 = plan->tmA0;
    }
  }
  if (!p.merged) {
    p.tl = choose_tl(d.Lo, TC_BM);
    p.tb = TC_BM / p.tl;
    bool has1 = false;
    rc = make_row_views(&plan->tmA0, &plan->tmA1, &has1, d.A, d.Kc, d.La, d.B, d.lda, d.stride, p.tl, p.tb, esz);
    if (rc) return rc;
    // taps that only ever read the (empty) odd view contribute nothing: drop them
    p.ntaps = 0;
    for (int t = 0; t < d.ntaps; ++t) {
      int map, coord;
      tap_view(d.stride, d.tap_off[t], &map, &coord);
      if (map == 1 && !has1) continue;
      p.tap_map[p.ntaps] = map;
      p.tap_coord[p.ntaps] = coord;
      p.tap_w[p.ntaps] = t;
      p.ntaps++;
    }
    B2H_CHECK_ARG(p.ntaps >= 1, B2H_ERR_SHAPE, "gemm_bf16: no tap reads inside the input (La=%d)", d.La);
    p.tap_lo = 0;
    p.a_box_bytes = 0;
  }
  p.tl_log2 = 0;
  while ((1 << p.tl_log2) < p.tl) ++p.tl_log2;
  p.tb_log2 = 0;
  while ((1 << p.tb_log2) < p.tb) ++p.tb_log2;
  p.n_lchunks = ceil_div(d.Lo, p.tl);
  const int m_tiles = ceil_div(d.B, p.tb) * p.n_lchunks;
  // tile width: the largest BN (dividing one phase) that minimises the estimated wave time
  const int half = d.Npad / d.nphase;
  const int nkb = p.ntaps * (d.Kc / bke);
  int best_bn = 64;
  double best_t = 1e30;
  const int sms = sm_count();
  // (3xTF32: three MMAs of half the rate per 32-channel k-block, tiles up to 128 columns: two accumulator slots)
  const int bn_max = esz == 4 ? 128 : 256;
  const double kcost = esz == 4 ? 6.0 : 2.0;
  for (int bn = 64; bn <= bn_max; bn <<= 1) {
    if (half % bn) continue;
    int64_t tiles = (int64_t)m_tiles * (d.Npad / bn);
    double waves = (double)((tiles + sms - 1) / sms);
    double t = waves * (3000.0 + (double)nkb * kcost * bn + 12.0 * bn);
    if (t < best_t * 0.999) {
      best_t = t;
      best_bn = bn;
    }
  }
  if (ncl || resid) best_bn = 256;
  // pooled backward sums (b2h_bwd_sums_t.rowmap = POOL2): the z box is twice the tile's height -- with 64-column tiles
  // it fits the idle pipeline buffers behind the staging tile (32 KB)
  const bool pool_bwd = esz == 2 && d.bwd_sums.z && d.bwd_sums.rowmap == B2H_ROW_POOL2 && !getenv("B2H_NO_FUSED_BWD") &&
                        !getenv("B2H_NO_POOL_BWDSUM");
  if (pool_bwd) best_bn = 64;
  if (const char* f = getenv("B2H_FORCE_BN"); f && !ncl && !resid && !pool_bwd) {  // tuning aid
    int bn = atoi(f);
    if ((bn == 64 || bn == 128 || bn == 256) && bn <= bn_max && half % bn == 0) best_bn = bn;
  }
  if (getenv("B2H_DEBUG_PLAN"))
    fprintf(stderr, "[b2h] gemm plan: B=%d Lo=%d Kc=%d Npad=%d ntaps=%d stride=%d nphase=%d -> merged=%d tl=%d tb=%d BN=%d tiles=%dx%d\n",
            d.B, d.Lo, d.Kc, d.Npad, d.ntaps, d.stride, d.nphase, p.merged, p.tl, p.tb, best_bn, m_tiles, d.Npad / best_bn);
  plan->BN = best_bn;
  plan->epi = epi_kind(d);
  // CTA pairs (cta_group::2) for the 256-column bf16 tile: two neighbouring M tiles share one MMA of M = 256, each
  // CTA stages half of the B tile.  An odd tile count is padded with a CTA whose rows are all outside the tensor.
  plan->pair = (esz == 2 && best_bn == 256 && m_tiles >= 2 && !ncl && !resid && pair_mode_enabled()) ? 1 : 0;
  plan->grid_x = plan->pair ? (m_tiles + 1) / 2 * 2 : m_tiles;
  plan->grid_y = d.Npad / best_bn;
  // launches of more than one wave of 256-column tiles with a plain epilogue (batched inference): persistent CTAs
  // with a double-buffered accumulator, so launch / pipeline fill / epilogue of a tile overlap the next tile's MMAs
  plan->persist = (esz == 2 && best_bn == 256 && !plan->pair && !d.stats.z && !d.bwd_sums.z &&
                   (int64_t)plan->grid_x * plan->grid_y > sms && persist_supports_epilogue(plan->epi) &&
                   !getenv("B2H_NO_PERSIST")) ? 1 : 0;
  if (d.grad_add) plan->persist = 0;   // (the added gradient is read by the one-tile kernel's epilogue)
  if (ncl) {
    // NCL output exists only in the persistent kernel: (T, C, B) fp32 map, boxes of tl frames x 32 channels x tb clips
    plan->persist = 1;
    rc = make_map_3d(&plan->tmO0, d.out, d.Lo_actual, d.Nvalid, d.B, d.Lo_actual, (int64_t)d.Nvalid * d.Lo_actual, p.tl, 32,
                     p.tb, 0, 4);
    if (rc) return rc;
    plan->tmO1 = plan->tmO0;
  }
  if (pool2)   // pairs of rows of one clip must sit in one warp: tile rows (frame, clip) with <= 16 clips, or clip-major
    B2H_CHECK_ARG(p.tl >= 2 && (!p.merged || p.tb <= 16) && d.Lo_actual >= 2, B2H_ERR_SHAPE,
                  "gemm: out_pool2 needs >= 2 frames per tile and clip (tl=%d tb=%d merged=%d)", p.tl, p.tb, p.merged);
  if (resid) {
    B2H_CHECK_ARG(persist_supports_epilogue(plan->epi) && plan->epi != EPI_BIAS_F32, B2H_ERR_ARG,
                  "gemm: a residual needs one of the specialised bf16 epilogues (bias + activation [+ folded BN])");
    plan->persist = 1;
  }
  if (plan->persist && !d.out_f32) {
    // output tensor maps of the TMA-store epilogue: one per sub-pixel phase (rows ph, ph + nphase, ...), boxes of 64
    // channels x the M tile in the tile's own row order; channels >= Nvalid and rows outside the tensor are clipped.
    // (resid_up2: the two "phases" are the even / odd output rows that GEMM row l is written to)
    const int ophases = (d.resid && d.resid_up2) ? 2 : d.nphase;
    const int orows = (d.resid && d.resid_up2) ? 2 * d.Lo_actual : (pool2 ? d.Lo_actual / 2 : d.Lo_actual);
    const __nv_bfloat16* o = reinterpret_cast<const __nv_bfloat16*>(d.out) + d.out_coff;
    for (int ph = 0; ph < ophases && !rc; ++ph) {
      CUtensorMap* mo = ph ? &plan->tmO1 : &plan->tmO0;
      const int rows = std::min(d.Lo, (orows - ph + ophases - 1) / ophases);
      // (pooled output: the box holds half the tile's rows, pairs of one clip merged)
      const int obox_l = pool2 ? p.tl / 2 : p.tl;
      if (rows <= 0) {   // a phase without rows (one-row outputs): not a multi-wave shape anyway
        plan->persist = 0;
        break;
      }
      const int64_t row_pitch = (int64_t)ophases * d.ldo, sample_pitch = (int64_t)orows * d.ldo;
      if (p.merged)
        rc = make_map_3d(mo, o + (int64_t)ph * d.ldo, d.Nvalid, d.B, rows, sample_pitch, row_pitch, 64, p.tb, obox_l, 1, 2);
      else
        rc = make_map_3d(mo, o + (int64_t)ph * d.ldo, d.Nvalid, rows, d.B, row_pitch, sample_pitch, 64, obox_l, p.tb, 1, 2);
    }
    if (ophases == 1) plan->tmO1 = plan->tmO0;
    if (rc) return rc;
  }
  rc = make_map_2d(&plan->tmB, d.W, (int64_t)d.ntaps * d.Kc, d.Npad, (int64_t)d.ntaps * d.Kc, bke,
                   plan->pair ? best_bn / 2 : best_bn, esz);
  if (rc) return rc;
  plan->fuse_stats = 0;
  plan->fuse_bwd = 0;
  if (esz == 4 && getenv("B2H_TF32_NO_FUSE")) return B2H_OK;   // (statistics / backward sums as separate passes)
  if (d.stats.z) {
    const b2h_bn_stats_t& st = d.stats;
    B2H_CHECK_ARG(st.z == d.out && d.out_coff == 0 && st.ld == d.ldo && st.C == d.Nvalid && st.groups >= 1 &&
                      d.B % st.groups == 0 && st.rows_per_group == (d.B / st.groups) * d.Lo_actual,
                  B2H_ERR_ARG, "gemm: stats must describe the output tensor of the op");
    const int kind = epi_kind(d);
    // (a ragged last 16-byte chunk is stored whole: the weight rows / bias beyond Nvalid are zero, so the
    // pad columns of z receive zeros; its statistics cover the valid columns only)
    plan->fuse_stats = (kind == EPI_BIAS_LEAKY || kind == EPI_BIAS_RELU) && d.ldo >= ((d.Nvalid + 7) & ~7) &&
                       (d.B / st.groups) % p.tb == 0 && st.partial && ((uintptr_t)st.partial % 16) == 0 &&
                       st.ticket && st.Cs >= st.C && !getenv("B2H_NO_FUSED_STATS");
  }
  plan->fuse_bwd = 0;
  if (d.bwd_sums.z) {
    const b2h_bwd_sums_t& bs = d.bwd_sums;
    const int kind = epi_kind(d);
    const bool up2 = bs.rowmap == B2H_ROW_UP2;
    const bool pool2z = bs.rowmap == B2H_ROW_POOL2;
    bool ok = (kind == EPI_MASK || kind == EPI_PLAIN) && !plan->fuse_stats && bs.C == d.Nvalid && d.out_coff == 0 &&
              d.ldo >= ((d.Nvalid + 7) & ~7) && bs.groups >= 1 && d.B % bs.groups == 0 &&
              (d.B / bs.groups) % p.tb == 0 && bs.accum && ((uintptr_t)bs.accum % 16) == 0 && bs.mean && bs.invstd &&
              bs.ld % 8 == 0 && bs.ld >= d.Npad / d.nphase && !getenv("B2H_NO_FUSED_BWD");
    if (up2)
      ok = ok && d.nphase == 1 && p.tl >= 2 && d.Lo_actual == 2 * bs.Lz;
    else if (pool2z)
      ok = ok && pool_bwd && best_bn == 64 && d.nphase == 1 && bs.Lz == 2 * d.Lo_actual && bs.scale && bs.shift &&
           ((128 * (64 * 2 + 16) + 64 * 64 + 127) & ~127) + 2 * 128 * 64 * 2 <= FpropCfg<64>::MAIN_BYTES;
    else
      ok = ok && bs.rowmap == B2H_ROW_IDENT && d.Lo_actual == bs.Lz;
    if (ok) {
      const uint8_t* z = reinterpret_cast<const uint8_t*>(bs.z);   // (act dtype: esz bytes per element)
      const size_t zrow = (size_t)bs.ld * esz;
      const int box_l = up2 ? p.tl / 2 : (pool2z ? 2 * p.tl : p.tl);
      const int64_t zs_b = (int64_t)bs.Lz * bs.ld;   // sample pitch of z
      if (p.merged && d.nphase == 1) {   // (C, B, L) like the A operand of a merged plan
        rc = make_map_3d(&plan->tmZ0, z, bs.ld, d.B, bs.Lz, zs_b, bs.ld, best_bn, p.tb, box_l, false, esz);
        plan->tmZ1 = plan->tmZ0;
      } else if (p.merged) {
        const int Le = (bs.Lz + 1) / 2, Lod = bs.Lz / 2;
        rc = make_map_3d(&plan->tmZ0, z, bs.ld, d.B, Le, zs_b, 2 * (int64_t)bs.ld, best_bn, p.tb, box_l, false, esz);
        if (!rc && Lod > 0)
          rc = make_map_3d(&plan->tmZ1, z + zrow, bs.ld, d.B, Lod, zs_b, 2 * (int64_t)bs.ld, best_bn, p.tb, box_l, false, esz);
        else
          plan->tmZ1 = plan->tmZ0;
      } else if (d.nphase == 1) {
        rc = make_map_3d(&plan->tmZ0, z, bs.ld, bs.Lz, d.B, bs.ld, (int64_t)bs.Lz * bs.ld, best_bn, box_l, p.tb, false, esz);
        plan->tmZ1 = plan->tmZ0;
      } else {   // output row 2*lo + ph <-> the even / odd rows of z
        const int Le = (bs.Lz + 1) / 2, Lod = bs.Lz / 2;
        rc = make_map_3d(&plan->tmZ0, z, bs.ld, Le, d.B, 2 * (int64_t)bs.ld, (int64_t)bs.Lz * bs.ld, best_bn, box_l, p.tb,
                         false, esz);
        if (!rc && Lod > 0)
          rc = make_map_3d(&plan->tmZ1, z + zrow, bs.ld, Lod, d.B, 2 * (int64_t)bs.ld, (int64_t)bs.Lz * bs.ld, best_bn,
                           box_l, p.tb, false, esz);
        else   // no odd rows: the odd-phase tiles have no valid row and ignore what they load
          plan->tmZ1 = plan->tmZ0;
      }
      if (rc) return rc;
      plan->fuse_bwd = 1;
      plan->bs_mean = bs.mean;
      plan->bs_invstd = bs.invstd;
      plan->bs_accum = bs.accum;
      plan->bs_C = bs.C;
      plan->bs_Cs = bs.Cs;
      plan->bs_groups = bs.groups;
      plan->bs_up2 = up2 ? 1 : 0;
      plan->bs_pool2 = pool2z ? 1 : 0;
      plan->bs_scale = bs.scale;
      plan->bs_shift = bs.shift;
      plan->bs_zbytes = best_bn * esz * box_l * p.tb;
    }
  }
  return B2H_OK;
}

static int epi_kind(const b2h_gemm_t& d) {
  if (getenv("B2H_GENERIC_EPI")) return EPI_GENERIC;
  const bool nodrop = d.drop.mode == B2H_DROP_NONE;
  if (d.post_scale && d.bias && nodrop && !d.out_f32 && d.act == B2H_ACT_LEAKY) return EPI_BIAS_LEAKY_BN;
  if (d.post_scale && d.bias && nodrop && !d.out_f32 && d.act == B2H_ACT_RELU) return EPI_BIAS_RELU_BN;
  if (d.post_scale) return EPI_GENERIC;
  if (d.bias && nodrop && !d.out_f32 && d.act == B2H_ACT_LEAKY) return EPI_BIAS_LEAKY;
  if (d.bias && nodrop && !d.out_f32 && d.act == B2H_ACT_RELU) return EPI_BIAS_RELU;
  if (d.bias && nodrop && d.out_f32 && d.act == B2H_ACT_NONE) return EPI_BIAS_F32;
  if (!d.bias && !d.out_f32 && d.act == B2H_ACT_NONE && d.drop.mode == B2H_DROP_MASK && d.drop_C % 4 == 0 &&
      ((uintptr_t)d.drop.mask % 4) == 0)
    return EPI_MASK;
  if (!d.bias && !d.out_f32 && d.act == B2H_ACT_NONE && nodrop) return EPI_PLAIN;
  return EPI_GENERIC;
}

template <int BN, int KIND, int MODE, bool MERGED>
static int launch_fprop_m(const TcGemmPlan& plan, const EpiParams& e, cudaStream_t s, const b2h_bn_stats_t& st);

template <int BN, int KIND, int MODE = MODE_PLAIN>
static int launch_fprop(const TcGemmPlan& plan, const EpiParams& e, cudaStream_t s,
                        const b2h_bn_stats_t& st = b2h_bn_stats_t()) {
  return plan.p.merged ? launch_fprop_m<BN, KIND, MODE, true>(plan, e, s, st)
                       : launch_fprop_m<BN, KIND, MODE, false>(plan, e, s, st);
}

template <int BN, int KIND, int MODE, bool MERGED>
static int launch_fprop_m(const TcGemmPlan& plan, const EpiParams& e, cudaStream_t s, const b2h_bn_stats_t& st) {
  using Cfg = FpropCfg<BN>;
  static_assert(((128 * (BN * 2 + 16) + 64 * BN + 127) & ~127) + 128 * BN * 2 <= Cfg::MAIN_BYTES, "z tile placement");
  B2H_CARVE(gemm_tc_kernel<BN, KIND, MODE, MERGED>);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t er = cudaFuncSetAttribute(gemm_tc_kernel<BN, KIND, MODE, MERGED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::SMEM_BYTES);
    if (er != cudaSuccess) return cuda_fail(er, "gemm_tc smem attribute");
    if (BN == 256) {
      B2H_CARVE(gemm_tc_kernel<256, KIND, MODE, MERGED, true>);
      er = cudaFuncSetAttribute(gemm_tc_kernel<256, KIND, MODE, MERGED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                FpropCfg<256>::SMEM_BYTES);
      if (er != cudaSuccess) return cuda_fail(er, "gemm_tc (pair) smem attribute");
    }
    attr_set = true;
  }
  dim3 grid(plan.grid_x, plan.grid_y);
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
    bs.pool2 = plan.bs_pool2;
    bs.scale = plan.bs_scale;
    bs.shift = plan.bs_shift;
  }
  const bool z = MODE == MODE_BWDSUM;
  if constexpr (KIND == EPI_MASK && (MODE == MODE_BWDSUM || MODE == MODE_PLAIN)) {
    if (e.grad_add && !(BN == 256 && plan.pair) && !getenv("B2H_NO_GADD_PREFETCH")) {
      static bool gattr = false;
      if (!gattr) {
        B2H_CARVE(gemm_tc_kernel<BN, KIND, MODE, MERGED, false, true>);
        cudaError_t er = cudaFuncSetAttribute(gemm_tc_kernel<BN, KIND, MODE, MERGED, false, true>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (er != cudaSuccess) return cuda_fail(er, "gemm_tc (grad_add) smem attribute");
        gattr = true;
      }
      launch(gemm_tc_kernel<BN, KIND, MODE, MERGED, false, true>, grid, TC_THREADS, Cfg::SMEM_BYTES, s, plan.tmA0, plan.tmA1,
             plan.tmB, z ? plan.tmZ0 : plan.tmA0, z ? plan.tmZ1 : plan.tmA0, plan.p, e, st, bs);
      B2H_LAUNCH_CHECK("gemm_tc");
      return B2H_OK;
    }
  }
  if (BN == 256 && plan.pair)
    launch_cluster(gemm_tc_kernel<256, KIND, MODE, MERGED, true>, grid, TC_THREADS, FpropCfg<256>::SMEM_BYTES, s, 2u,
                   plan.tmA0, plan.tmA1, plan.tmB, z ? plan.tmZ0 : plan.tmA0, z ? plan.tmZ1 : plan.tmA0, plan.p, e, st, bs);
  else
    launch(gemm_tc_kernel<BN, KIND, MODE, MERGED>, grid, TC_THREADS, Cfg::SMEM_BYTES, s, plan.tmA0, plan.tmA1, plan.tmB,
           z ? plan.tmZ0 : plan.tmA0, z ? plan.tmZ1 : plan.tmA0, plan.p, e, st, bs);
  B2H_LAUNCH_CHECK("gemm_tc");
  return B2H_OK;
}

template <int BN>
static int launch_fprop_kind(const TcGemmPlan& plan, const EpiParams& e, int kind, cudaStream_t s,
                             const b2h_bn_stats_t& st) {
  if (plan.fuse_stats) {
    if (kind == EPI_BIAS_LEAKY) return launch_fprop<BN, EPI_BIAS_LEAKY, MODE_STATS>(plan, e, s, st);
    return launch_fprop<BN, EPI_BIAS_RELU, MODE_STATS>(plan, e, s, st);
  }
  if (plan.fuse_bwd) {
    if (kind == EPI_MASK) return launch_fprop<BN, EPI_MASK, MODE_BWDSUM>(plan, e, s);
    return launch_fprop<BN, EPI_PLAIN, MODE_BWDSUM>(plan, e, s);
  }
  switch (kind) {
    case EPI_BIAS_LEAKY: return launch_fprop<BN, EPI_BIAS_LEAKY>(plan, e, s);
    case EPI_BIAS_RELU: return launch_fprop<BN, EPI_BIAS_RELU>(plan, e, s);
    case EPI_BIAS_F32: return launch_fprop<BN, EPI_BIAS_F32>(plan, e, s);
    case EPI_MASK: return launch_fprop<BN, EPI_MASK>(plan, e, s);
    case EPI_PLAIN: return launch_fprop<BN, EPI_PLAIN>(plan, e, s);
    case EPI_BIAS_LEAKY_BN: return launch_fprop<BN, EPI_BIAS_LEAKY_BN>(plan, e, s);
    case EPI_BIAS_RELU_BN: return launch_fprop<BN, EPI_BIAS_RELU_BN>(plan, e, s);
    default: return launch_fprop<BN, EPI_GENERIC>(plan, e, s);
  }
}

int run_gemm_bf16(const TcGemmPlan& plan, const b2h_gemm_t& d, cudaStream_t s) {
  EpiParams e = make_epi(d);
  const int kind = epi_kind(d);
  if (plan.persist && persist_supports_epilogue(kind)) return run_gemm_persist(plan, d, kind, s);
  int rc;
  switch (plan.BN) {
    case 256: rc = launch_fprop_kind<256>(plan, e, kind, s, d.stats); break;
    case 128: rc = launch_fprop_kind<128>(plan, e, kind, s, d.stats); break;
    default: rc = launch_fprop_kind<64>(plan, e, kind, s, d.stats); break;
  }
  if (rc) return rc;
  // shapes the epilogue cannot cover: separate passes with the same results
  if (d.stats.z && !plan.fuse_stats) rc = launch_bn_stats(d.stats, B2H_BF16, s);
  if (!rc && d.bwd_sums.z && !plan.fuse_bwd) rc = launch_bwd_sums_separate(d, B2H_BF16, s);
  return rc;
}

int plan_wgrad_bf16(const b2h_wgrad_t& d, TcWgradPlan* plan) { return plan_wgrad_tc(d, plan, 2); }

// esz = 4 (3xTF32): k-blocks of 32 rows (both operands are staged twice: as loaded and as the low-order split)
int plan_wgrad_tc(const b2h_wgrad_t& d, TcWgradPlan* plan, int esz) {
  TcWgradParams& p = plan->p;
  plan->esz = esz;
  const int wk = wgrad_kblock_rows(esz);
  p.Mpad = d.Mpad;
  p.Npad = d.Npad;
  p.ntaps = d.ntaps;
  p.tl = choose_tl(d.Lp, wk);
  p.tb = wk / p.tl;
  p.n_lchunks = ceil_div(d.Lp, p.tl);
  p.total_kb = ceil_div(d.B, p.tb) * p.n_lchunks;
  // MN-major fp32 operands: the 128-byte swizzle over 32-byte chunks (smem_desc_sw128_base32)
  const int swz = esz == 4 ? 2 : 1;
  int rc = make_map_3d(&plan->tmP, d.P, d.Mpad, d.Lp, d.B, d.ldp, (int64_t)d.Lp * d.ldp, 128 / esz, p.tl, p.tb, swz, esz);
  if (rc) return rc;
  bool has1 = false;
  rc = make_row_views(&plan->tmQ0, &plan->tmQ1, &has1, d.Q, d.Npad, d.Lq, d.B, d.ldq, d.stride, p.tl, p.tb, esz, swz);
  if (rc) return rc;
  for (int t = 0; t < d.ntaps; ++t) {
    tap_view(d.stride, d.tap_off[t], &p.tap_map[t], &p.tap_coord[t]);
    if (p.tap_map[t] == 1 && !has1) p.tap_map[t] = -1;  // empty odd-row view: the tap's gradient is zero
  }
  int wn = 64;
  if (d.Npad % 256 == 0 && esz == 2)
    wn = 256;
  else if (d.Npad % 128 == 0)
    wn = 128;
  if (const char* e = getenv("B2H_WGRAD_MAX_WN")) wn = std::min(wn, std::max(64, atoi(e)));   // tuning aid
  // prefer more tiles over wider tiles when the grid would not fill the machine
  const int sms = sm_count();
  while (wn > 64 && (int64_t)ceil_div(d.Mpad, WG_BM) * (d.Npad / wn) * d.ntaps * std::max(1, p.total_kb / 8) < sms) wn >>= 1;
  if (const char* e = getenv("B2H_FORCE_WN")) {   // tests: run a given tile width at any problem size
    const int f = atoi(e);
    if ((f == 64 || f == 128 || (f == 256 && esz == 2)) && d.Npad % f == 0) wn = f;
  }
  plan->WN = wn;
  int64_t tiles = (int64_t)ceil_div(d.Mpad, WG_BM) * (d.Npad / wn) * d.ntaps;
  // split-K target: CTAs per launch.  The weight gradients run beside the dgrad chain, so the launch does not
  // have to fill the machine on its own; fewer, longer CTAs cost less SM time (fixed per-CTA overhead, fp32
  // partial planes) at the price of a longer launch.
  int target = sms;
  if (const char* e = getenv("B2H_WGRAD_CTAS")) target = std::min(sms, std::max(1, atoi(e)));
  // WN = 256 runs one CTA per SM: a grid of target+1 CTAs would take two waves, so round the split count down
  int splits = d.splits > 0 ? d.splits
               : (int)std::max<int64_t>(1, (wn == 256 || esz == 4) ? target / tiles : (target + tiles - 1) / tiles);
  // 3xTF32: the tensor core adds into its fp32 accumulator with truncation, a relative loss of ~2^-24 per add that
  // grows with the length of the k range: bound the k-blocks of one split (the partial planes are summed in fp32 RN)
  if (esz == 4) splits = std::max(splits, ceil_div(p.total_kb, kTf32MaxKbPerSplit));
  if (splits > p.total_kb) splits = p.total_kb;
  if (splits > 64) splits = 64;
  p.kb_per_split = ceil_div(p.total_kb, splits);
  plan->splits = ceil_div(p.total_kb, p.kb_per_split);
  B2H_CHECK_ARG(plan->splits == 1 || d.partial_bytes <= 0 ||
                    (int64_t)plan->splits * d.ntaps * d.Mpad * d.Npad * (int64_t)sizeof(float) <= d.partial_bytes,
                B2H_ERR_ARG, "wgrad: workspace of %lld bytes is too small for %d splits x %d taps x %d x %d",
                (long long)d.partial_bytes, plan->splits, d.ntaps, d.Mpad, d.Npad);
  p.direct = plan->splits == 1 ? 1 : 0;
  p.Mvalid = d.Mvalid;
  p.Nvalid = d.Nvalid;
  plan->grid_x = ceil_div(d.Mpad, WG_BM) * (d.Npad / wn);
  return B2H_OK;
}

template <int WN>
static int launch_wg(const TcWgradPlan& plan, float* partial, cudaStream_t s) {
  B2H_CARVE(wgrad_tc_kernel<WN>);
  using Cfg = WgradCfg<WN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t er = cudaFuncSetAttribute(wgrad_tc_kernel<WN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (er != cudaSuccess) return cuda_fail(er, "wgrad_tc smem attribute");
    attr_set = true;
  }
  dim3 grid(plan.grid_x, plan.p.ntaps, plan.splits);
  launch(wgrad_tc_kernel<WN>, grid, 192, Cfg::SMEM_BYTES, s, plan.tmP, plan.tmQ0, plan.tmQ1, plan.p, partial);
  B2H_LAUNCH_CHECK("wgrad_tc");
  return B2H_OK;
}

int run_wgrad_bf16(const TcWgradPlan& plan, const b2h_wgrad_t& d, cudaStream_t s) {
  float* out = plan.p.direct ? d.dW : d.partial;   // one split: straight into dW, no reduce launch
  int rc;
  switch (plan.WN) {
    case 256: rc = launch_wg<256>(plan, out, s); break;
    case 128: rc = launch_wg<128>(plan, out, s); break;
    default: rc = launch_wg<64>(plan, out, s); break;
  }
  if (rc || plan.p.direct) return rc;
  return launch_wgrad_reduce(d, plan.splits, s);
}

int64_t wgrad_bf16_workspace_bytes(const b2h_wgrad_t& d) { return wgrad_tc_workspace_bytes(d, 2); }

int64_t wgrad_tc_workspace_bytes(const b2h_wgrad_t& d, int esz) {
  // upper bound independent of the plan: splits <= 64 but never more than total k-blocks
  const int wk = wgrad_kblock_rows(esz);
  int tl = choose_tl(d.Lp, wk);
  int total_kb = ceil_div(d.B, wk / tl) * ceil_div(d.Lp, tl);
  int sms = sm_count();
  int64_t tiles_min = (int64_t)ceil_div(d.Mpad, WG_BM) * std::max(1, d.Npad / 256) * d.ntaps;
  int64_t splits = d.splits > 0 ? d.splits : std::max<int64_t>(1, (sms + tiles_min - 1) / tiles_min);
  if (esz == 4) splits = std::max<int64_t>(splits, ceil_div(total_kb, kTf32MaxKbPerSplit));
  splits = std::min<int64_t>(std::min<int64_t>(splits, total_kb), 64);
  return (splits + 1) * (int64_t)d.ntaps * d.Mpad * d.Npad * (int64_t)sizeof(float);
}

}  // namespace b2h
