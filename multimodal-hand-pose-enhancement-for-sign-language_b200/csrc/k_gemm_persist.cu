// Persistent form of the bf16 tap-GEMM for launches of several waves (batched inference: 4096 x 64 frames = 2048
// tiles of 128 x 256 on 148 SMs).
//
// What the one-tile-per-CTA kernel of k_gemm_tc.cu leaves on the table there (ncu, profiles/ncu_r02_infer.md): a
// 128 x 256 x 768 tile needs 3.1 us of tensor-pipe time but a CTA takes 11.7 us for it -- CTA launch and prologue
// (barrier init, TMEM allocation), pipeline fill (the first TMA round trip), then the epilogue (TMEM -> registers ->
// global) all run with the tensor pipe idle: 33 % active, L2 at 17 %, DRAM at 15 %.  Here one CTA per SM stays
// resident and walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...:
//   * the TMA producer warp runs ahead across tile boundaries (the operand rings never drain),
//   * the accumulator is double-buffered in TMEM (2 x 256 columns): the MMA warp starts tile i+1 while the eight
//     epilogue warps drain tile i, handing slots back through tmem_empty barriers,
//   * the epilogue never waits for global memory: with per-thread st.global the store phase alone took 60 % of such a
//     launch (skipping the stores: 100.6 -> 39.4 us for the first layer at 4096 x 64, profiles/persist_r02.md) -- the
//     LSU's store queue back-pressures the epilogue warps, which hold the TMEM slot the MMA warp is waiting for.  The
//     bf16 tile is written to a 128B-swizzled staging buffer (4 boxes of 128 rows x 64 channels), the TMEM slot is
//     released at once, and ONE thread issues 4 TMA stores (cp.async.bulk.tensor, UTMASTG) that drain in the
//     background; rows / channels outside the tensor are clipped by the tensor map.  (fp32 BLC outputs are twice the
//     staging size and keep per-thread stores.)
//   * NCL: the output layer of a batched inference writes the (B, C, T) fp32 result of the network itself
//     (b2h_gemm_t.out_f32 = 2): a thread owns one frame of one clip, so it writes its 32 channels of a chunk as column
//     [clip][channel][frame] of the staging buffer (lanes = consecutive frames: conflict-free) and TMA stores of
//     (tl frames x 32 channels x tb clips) boxes produce the transposed layout, two rounds of 128 channels per tile --
//     instead of fp32 BLC to HBM, a to_ncl pass reading it back and writing NCL (264 MB each way at 4096 x 64).
// Same tile decomposition, operands, descriptors and epilogue arithmetic as gemm_tc_kernel<256, KIND, MODE_PLAIN, MERGED>:
// results are bit-identical (tests/test_gpu_replay.py compares both against the restatement).
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "gemm_tc_shared.cuh"

namespace b2h {

using namespace ptx;

constexpr int PBN = 256;

// shared memory of the persistent kernel: operand rings (3 joint stages of 48 KB, or the tap-merged A ring 2 x 24 KB +
// B ring 3 x 32 KB = 144 KB) + the output staging buffer (4 swizzled boxes of 128 rows x 128 bytes = 64 KB)
struct PersistCfg {
  static constexpr int STAGES = 3;
  static constexpr int SB = 3;
  static constexpr int RING_BYTES = STAGES * FpropCfg<PBN>::STAGE_BYTES;
  static_assert(FpropCfg<PBN>::SA * FpropCfg<PBN>::A_STAGE + SB * FpropCfg<PBN>::B_BYTES <= RING_BYTES, "merged rings");
  static constexpr int BOX_BYTES = 128 * 128;
  static constexpr int STAGING_BYTES = 4 * BOX_BYTES;
  static constexpr int MAIN_BYTES = RING_BYTES + STAGING_BYTES;
  static constexpr int SMEM_BYTES = MAIN_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 3 * PBN * 4 /*bias | scale | shift*/;
  static_assert(RING_BYTES % 1024 == 0 && SMEM_BYTES <= 232448, "shared memory budget");
};

// epi_fast8 with the bias / folded-BN vectors read straight from global memory (warp-uniform addresses, L1-resident
// after the first tile): the tiles of a persistent CTA have different column offsets and its warps are not in step
template <int KIND>
__device__ __forceinline__ void epi_global8(const EpiParams& e, int nn, const uint32_t* acc_bits, float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float x = __uint_as_float(acc_bits[j]);
    if (epi_has_bias(KIND)) x += __ldg(e.bias + nn + j);
    if (KIND == EPI_BIAS_LEAKY || KIND == EPI_BIAS_LEAKY_BN) x = x > 0.f ? x : x * kLeakySlope;
    if (KIND == EPI_BIAS_RELU || KIND == EPI_BIAS_RELU_BN) x = x > 0.f ? x : 0.f;
    if (epi_has_bn(KIND)) x = fmaf(x, __ldg(e.post_scale + nn + j), __ldg(e.post_shift + nn + j));
    v[j] = x;
  }
}

template <int KIND, bool MERGED, bool NCL = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO0,
                       const __grid_constant__ CUtensorMap tmO1, TcGemmParams p, EpiParams e, int m_tiles, int n_tiles,
                       int dbg /* timing experiments only: 1 = no global stores, 2 = no TMEM loads either */) {
  using Cfg = FpropCfg<PBN>;
  using PC = PersistCfg;
  constexpr int STAGES = PC::STAGES;
  constexpr bool TMA_OUT = KIND != EPI_BIAS_F32;   // bf16 tiles leave through the staging buffer + TMA stores
  static_assert(!NCL || (KIND == EPI_BIAS_F32 && !MERGED), "NCL output: fp32, rows of a warp = consecutive frames");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + PC::RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + PC::MAIN_BYTES);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* a_full = empty_bar + 4;        // MERGED: the A ring (full_bar / empty_bar are the B ring)
  uint64_t* a_empty = a_full + 2;
  uint64_t* tmem_full = a_empty + 2;       // [2] accumulator slot written
  uint64_t* tmem_empty = tmem_full + 2;    // [2] accumulator slot drained (one arrival per epilogue warp)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  // bias / folded-BN scale / shift of the tile's 256 columns (bf16 tiles: reloaded when the column offset changes)
  float* s_bias = reinterpret_cast<float*>(smem + PC::MAIN_BYTES + 256);
  float* s_scale = s_bias + PBN;
  float* s_shift = s_scale + PBN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kpt = p.Kc / TC_BK;
  const int nkb = p.ntaps * kpt;
  const int total = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA0);
    prefetch_tmap(&tmA1);
    prefetch_tmap(&tmB);
    if (TMA_OUT || NCL) {
      prefetch_tmap(&tmO0);
      prefetch_tmap(&tmO1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
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

  if (warp == 0) {
    if (lane == 0) {
      int ia = 0, ib = 0;   // running ring counters: they continue across tiles
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
        const int bt = mt / p.n_lchunks, lc = mt - bt * p.n_lchunks;
        const int b0 = bt * p.tb, l0 = lc * p.tl, n0 = nt * PBN;
        if (MERGED) {
          uint8_t* ringB = smem + Cfg::SA * Cfg::A_STAGE;
          for (int kc = 0; kc < kpt; ++kc, ++ia) {
            const int sa = ia % Cfg::SA;
            mbar_wait(&a_empty[sa], ((ia / Cfg::SA) & 1) ^ 1);
            mbar_arrive_expect_tx(&a_full[sa], (uint32_t)p.a_box_bytes);
            tma_load_3d(smem + sa * Cfg::A_STAGE, &tmA0, &a_full[sa], kc * TC_BK, b0, l0 + p.tap_lo);
            for (int t = 0; t < p.ntaps; ++t, ++ib) {
              const int sb = ib % PC::SB;
              mbar_wait(&empty_bar[sb], ((ib / PC::SB) & 1) ^ 1);
              mbar_arrive_expect_tx(&full_bar[sb], Cfg::B_BYTES);
              tma_load_2d(ringB + sb * Cfg::B_BYTES, &tmB, &full_bar[sb], p.tap_w[t] * p.Kc + kc * TC_BK, n0);
            }
          }
        } else {
          for (int kb = 0; kb < nkb; ++kb, ++ib) {
            const int stage = ib % STAGES;
            mbar_wait(&empty_bar[stage], ((ib / STAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            const int t = kb / kpt, kc = kb - t * kpt;
            uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
            tma_load_3d(sA, p.tap_map[t] ? &tmA1 : &tmA0, &full_bar[stage], kc * TC_BK, l0 + p.tap_coord[t], b0);
            tma_load_2d(sA + TC_A_BYTES, &tmB, &full_bar[stage], p.tap_w[t] * p.Kc + kc * TC_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(TC_BM, PBN, 0, 0);
      int ia = 0, ib = 0, it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int slot = it & 1;
        mbar_wait(&tmem_empty[slot], ((it >> 1) & 1) ^ 1);   // the epilogue has drained this slot's previous tile
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)slot * PBN;
        if (MERGED) {
          const uint32_t ringB = smem_u32(smem + Cfg::SA * Cfg::A_STAGE);
          for (int kc = 0; kc < kpt; ++kc, ++ia) {
            const int sa = ia % Cfg::SA;
            mbar_wait(&a_full[sa], (ia / Cfg::SA) & 1);
            tc_fence_after();
            const uint32_t sA = smem_u32(smem + sa * Cfg::A_STAGE);
            for (int t = 0; t < p.ntaps; ++t, ++ib) {
              const int sb = ib % PC::SB;
              mbar_wait(&full_bar[sb], (ib / PC::SB) & 1);
              tc_fence_after();
              const uint64_t adesc = smem_desc_sw128(sA + (uint32_t)(t * p.tb) * 128u, 16, 1024);
              const uint64_t bdesc = smem_desc_sw128(ringB + sb * Cfg::B_BYTES, 16, 1024);
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k)
                umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | t | k) != 0);
              umma_commit(&empty_bar[sb]);
            }
            umma_commit(&a_empty[sa]);
          }
        } else {
          for (int kb = 0; kb < nkb; ++kb, ++ib) {
            const int stage = ib % STAGES;
            mbar_wait(&full_bar[stage], (ib / STAGES) & 1);
            tc_fence_after();
            const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint64_t adesc = smem_desc_sw128(sA, 16, 1024);
            const uint64_t bdesc = smem_desc_sw128(sA + TC_A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            umma_commit(&empty_bar[stage]);
          }
        }
        umma_commit(&tmem_full[slot]);
      }
    }
  } else {
    // epilogue warps 2..9: TMEM sub-partition = warp % 4, two warps per sub-partition split the tile columns
    const int sub = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;   // 0..255
    const int r = sub * 32 + lane;     // tile row == TMEM lane
    int it = 0, s_nn0 = -1;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
      const int bt = mt / p.n_lchunks, lc = mt - bt * p.n_lchunks;
      const int b0 = bt * p.tb, l0 = lc * p.tl, n0 = nt * PBN;
      const int ph = n0 / e.half;           // a tile never straddles a sub-pixel phase
      const int nn0 = n0 - ph * e.half;     // first channel (within the phase) of this tile
      const int valid_cols = min(PBN, e.Nvalid - nn0);
      const int bi = MERGED ? (r & (p.tb - 1)) : (r >> p.tl_log2), li = MERGED ? (r >> p.tb_log2) : (r & (p.tl - 1));
      const int b = b0 + bi, lo = l0 + li;
      const int ris = lo * e.nphase + ph;
      const bool row_in = (b < p.B) && (lo < p.Lo) && (ris < e.Lo_actual);
      const int64_t grow = (int64_t)b * e.Lo_actual + ris;
      float* grow_f32 = reinterpret_cast<float*>(e.out) + (size_t)grow * e.ldo + e.out_coff + nn0;
      const int slot = it & 1;
      if (NCL) {
        // two rounds of 128 channels: warp (sub, chalf) drains channels round*128 + chalf*64 + [0, 64) of its 32 rows
        const int tl = p.tl;
#pragma unroll 1
        for (int round = 0; round < 2; ++round) {
          if (round == 0 && nn0 != s_nn0) {            // (all warps are past the previous tile's arithmetic)
            s_bias[et] = e.bias[nn0 + et];
            s_nn0 = nn0;
          }
          if (et == 0) tma_store_wait_read();          // the previous round's boxes have left the staging buffer
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (round == 0) {
            mbar_wait(&tmem_full[slot], (it >> 1) & 1);
            tc_fence_after();
          }
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int box = chalf * 2 + cc;            // 32-channel box of this round
            const int c = round * 128 + box * 32;
            uint32_t acc[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(slot * PBN + c), acc);
            tmem_ld_wait();
            if (c >= valid_cols) continue;
            float* col0 = reinterpret_cast<float*>(staging + box * PC::BOX_BYTES) + (size_t)(bi * 32) * tl + li;
#pragma unroll
            for (int j = 0; j < 32; ++j)   // [clip][channel][frame]: the lanes of a warp are consecutive frames
              col0[(size_t)j * tl] = __uint_as_float(acc[j]) + s_bias[c + j];
          }
          if (round == 1) {   // every TMEM read of this tile is done: hand the slot back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[slot]);
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (et == 0) {
#pragma unroll
            for (int bx = 0; bx < 4; ++bx) {
              const int c = round * 128 + bx * 32;
              if (c < valid_cols) tma_store_3d(&tmO0, staging + bx * PC::BOX_BYTES, l0, nn0 + c, b0);
            }
            tma_store_commit();
          }
        }
        continue;
      }
      const int n_rounds = (TMA_OUT && e.resid && e.resid_up2) ? 2 : 1;   // up2: output rows 2l (round 0) and 2l+1 (round 1)
#pragma unroll 1
      for (int rnd = 0; rnd < n_rounds; ++rnd) {
      if (TMA_OUT) {
        // (every epilogue warp is past the previous tile's arithmetic here: its bar.sync before the TMA issue)
        if (nn0 != s_nn0) {
          if (epi_has_bias(KIND)) s_bias[et] = e.bias[nn0 + et];
          if (epi_has_bn(KIND)) s_scale[et] = e.post_scale[nn0 + et], s_shift[et] = e.post_shift[nn0 + et];
          s_nn0 = nn0;
        }
        // the staging buffer is free once the previous tile's TMA stores have read it
        if (et == 0) tma_store_wait_read();
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (rnd == 0) {
        mbar_wait(&tmem_full[slot], (it >> 1) & 1);
        tc_fence_after();
      }
      // residual row of this thread's output row (this round's, when up-sampling)
      const __nv_bfloat16* rrow = nullptr;
      if (TMA_OUT && e.resid && row_in)
        rrow = reinterpret_cast<const __nv_bfloat16*>(e.resid) +
               (e.resid_up2 ? ((int64_t)b * 2 * e.Lo_actual + 2 * lo + rnd) : grow) * e.ld_resid + nn0;
      constexpr int CH = PBN / 2;  // columns per epilogue warp
#pragma unroll 1
      for (int c = chalf * CH; c < (chalf + 1) * CH; c += 32) {
        uint32_t acc[32];
        if (dbg & 2) continue;
        tmem_ld_32x32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(slot * PBN + c), acc);
        tmem_ld_wait();
        if (!TMA_OUT && !row_in) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (c + j >= valid_cols) break;
          float v[8];
          if (TMA_OUT)
            epi_fast8<KIND>(s_bias, s_scale, s_shift, nullptr, c + j, nn0 + c + j, acc + j, v);
          else
            epi_global8<KIND>(e, nn0 + c + j, acc + j, v);
          if (TMA_OUT) {
            // box = 64 channels; 16-byte chunk g of row r sits at chunk g ^ (r & 7) of its 128-byte line (128B swizzle)
            const int col = c + j;
            if (e.out_pool2) {
              // MaxPool1d(2): the partner frame of this thread's row sits in the same warp (lane ^ 1 in clip-major
              // tiles, lane ^ tb in the tap-merged (frame, clip) order with tb <= 16)
              const int pm = MERGED ? p.tb : 1;
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], __shfl_xor_sync(0xffffffffu, v[k], pm));
            }
            if (rrow) {
              const uint4 rq = *reinterpret_cast<const uint4*>(rrow + col);
              const uint32_t rw[4] = {rq.x, rq.y, rq.z, rq.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 rf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rw[k]));
                v[2 * k] += rf.x, v[2 * k + 1] += rf.y;
              }
            }
            __nv_bfloat162 q0 = __floats2bfloat162_rn(v[0], v[1]), q1 = __floats2bfloat162_rn(v[2], v[3]);
            __nv_bfloat162 q2 = __floats2bfloat162_rn(v[4], v[5]), q3 = __floats2bfloat162_rn(v[6], v[7]);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&q0);
            u.y = *reinterpret_cast<uint32_t*>(&q1);
            u.z = *reinterpret_cast<uint32_t*>(&q2);
            u.w = *reinterpret_cast<uint32_t*>(&q3);
            if (e.out_pool2) {
              if ((li & 1) == 0) {   // even frames carry the pair's maximum to row (frame / 2) of the half-height box
                const int rp = MERGED ? ((li >> 1) * p.tb + bi) : (bi * (p.tl >> 1) + (li >> 1));
                uint8_t* dst = staging + (col >> 6) * PC::BOX_BYTES + rp * 128 + ((((col & 63) >> 3) ^ (rp & 7)) << 4);
                *reinterpret_cast<uint4*>(dst) = u;
              }
            } else {
              uint8_t* dst = staging + (col >> 6) * PC::BOX_BYTES + r * 128 + ((((col & 63) >> 3) ^ (r & 7)) << 4);
              *reinterpret_cast<uint4*>(dst) = u;
            }
          } else {
            const int nv = min(8, valid_cols - (c + j));
            if ((dbg & 1) && v[0] != 123456.f) continue;   // (keeps the arithmetic alive)
            float* dst = grow_f32 + c + j;
            if (nv == 8) {
              reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
              reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                if (k < nv) dst[k] = v[k];
            }
          }
        }
      }
      if (rnd == n_rounds - 1) {   // this warp has read its part of the slot: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[slot]);
      }
      if (TMA_OUT) {
        fence_proxy_async_smem();   // the staged tile: generic-proxy writes -> async-proxy (TMA) reads
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0 && !(dbg & 1)) {
          const CUtensorMap* mo = (n_rounds == 2 ? rnd : ph) ? &tmO1 : &tmO0;
#pragma unroll
          for (int bx = 0; bx < 4; ++bx) {
            if (bx * 64 < valid_cols) {
              const int lrow = e.out_pool2 ? (l0 >> 1) : l0;   // (pooled output: half the rows)
              if (MERGED)   // output maps of a merged plan are (C, B, L) like its A operand
                tma_store_3d(mo, staging + bx * PC::BOX_BYTES, nn0 + bx * 64, b0, lrow);
              else
                tma_store_3d(mo, staging + bx * PC::BOX_BYTES, nn0 + bx * 64, lrow, b0);
            }
          }
          tma_store_commit();
        }
      }
      }   // rounds
    }
    if ((TMA_OUT || NCL) && et == 0) tma_store_wait_all();
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
template <int KIND, bool MERGED, bool NCL = false>
static int launch_persist(const TcGemmPlan& plan, const EpiParams& e, cudaStream_t s) {
  constexpr int SMEM = PersistCfg::SMEM_BYTES;
  B2H_CARVE(gemm_tc_persist_kernel<KIND, MERGED, NCL>);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t er = cudaFuncSetAttribute(gemm_tc_persist_kernel<KIND, MERGED, NCL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (er != cudaSuccess) return cuda_fail(er, "gemm_tc_persist smem attribute");
    attr_set = true;
  }
  const int total = plan.grid_x * plan.grid_y;
  const int ctas = std::min(total, sm_count());
  static const int dbg = getenv("B2H_PERSIST_DBG") ? atoi(getenv("B2H_PERSIST_DBG")) : 0;
  launch(gemm_tc_persist_kernel<KIND, MERGED, NCL>, dim3(ctas), TC_THREADS, SMEM, s, plan.tmA0, plan.tmA1, plan.tmB, plan.tmO0,
         plan.tmO1, plan.p, e, plan.grid_x, plan.grid_y, dbg);
  B2H_LAUNCH_CHECK("gemm_tc_persist");
  return B2H_OK;
}

template <int KIND>
static int launch_persist_m(const TcGemmPlan& plan, const EpiParams& e, cudaStream_t s) {
  return plan.p.merged ? launch_persist<KIND, true>(plan, e, s) : launch_persist<KIND, false>(plan, e, s);
}

// the epilogue kinds the persistent kernel covers (the others stay with the one-tile-per-CTA kernel)
bool persist_supports_epilogue(int kind) {
  return kind == EPI_BIAS_LEAKY || kind == EPI_BIAS_RELU || kind == EPI_BIAS_F32 || kind == EPI_BIAS_LEAKY_BN ||
         kind == EPI_BIAS_RELU_BN || kind == EPI_PLAIN;
}

int run_gemm_persist(const TcGemmPlan& plan, const b2h_gemm_t& d, int kind, cudaStream_t s) {
  EpiParams e = make_epi(d);
  switch (kind) {
    case EPI_BIAS_LEAKY: return launch_persist_m<EPI_BIAS_LEAKY>(plan, e, s);
    case EPI_BIAS_RELU: return launch_persist_m<EPI_BIAS_RELU>(plan, e, s);
    case EPI_BIAS_F32:
      if (d.out_f32 == 2) {
        B2H_CHECK_ARG(!plan.p.merged, B2H_ERR_ARG, "gemm_tc_persist: NCL output needs the plain (unmerged) tile order");
        return launch_persist<EPI_BIAS_F32, false, true>(plan, e, s);
      }
      return launch_persist_m<EPI_BIAS_F32>(plan, e, s);
    case EPI_BIAS_LEAKY_BN: return launch_persist_m<EPI_BIAS_LEAKY_BN>(plan, e, s);
    case EPI_BIAS_RELU_BN: return launch_persist_m<EPI_BIAS_RELU_BN>(plan, e, s);
    case EPI_PLAIN: return launch_persist_m<EPI_PLAIN>(plan, e, s);
    default: set_error("gemm_tc_persist: unsupported epilogue %d", kind); return B2H_ERR_ARG;
  }
}

}  // namespace b2h
