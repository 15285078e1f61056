// BatchNorm1d statistics / apply / backward and column sums (HBM/L2-bound elementwise + reduction kernels).
// Layout: rows x ld, channels contiguous, 8 channels per thread (one 16 B access in bf16), CTA = (TXp, TY).
//
// Reference semantics restated (modelZoo.py:192-198, SURVEY.md section 8a "PyTorch semantics"):
//   train: y = (z - mean_b) / sqrt(var_b(biased) + eps) * gamma + beta;
//          running = (1-m)*running + m*batch (running_var from the UNBIASED batch variance);
//   eval:  y = (z - running_mean) / sqrt(running_var + eps) * gamma + beta.
// Cross-CTA sums are combined with fp64 atomics on zero-initialised accumulators (the order of fp64 additions
// only perturbs bits far below the fp32 results written out).
// Both are consumed as a folded per-channel affine y = z*scale + shift ([groups][Cs] arrays, zero in the
// channel padding) that bn_stats (train) or bn_fold (eval) writes.
//
// Code size matters here: these kernels run for a few microseconds, and a launch whose straight-line code
// does not fit the instruction caches is bound by instruction fetch.  Hence small loop bodies, rolled loops
// and out-of-line Philox.
#include <algorithm>
#include <string.h>

#include "bn_finalize.cuh"

namespace b2h {

static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

__device__ __forceinline__ F8 zero8() {
  F8 r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = 0.f;
  return r;
}
__device__ __forceinline__ F8 ld8f(const float* p) { return load8<float>(p); }

// ---------------------------------------------------------------------------------------------
// bn_stats: thread = 8 channels x 4 consecutive rows (all loads issued before use), shifted sums around a pivot
// common to the whole launch (the running mean), per-CTA fp32 partials combined with fp64 atomics; the last
// CTA (ticket) turns the accumulators into mean / invstd / scale / shift, updates the running statistics and
// re-zeroes the accumulators for the next launch.
// ---------------------------------------------------------------------------------------------
constexpr int kRowThreads = 256;

// block-level sum over ty of one F8 per thread, result in ty == 0.  (Loops over the 8 lanes of an F8 are always
// fully unrolled: a dynamically indexed register array would be demoted to local memory.)
__device__ __forceinline__ void block_sum4(float4* sm, float4& v, int tx, int ty, int TXp, int TY) {
  __syncthreads();
  sm[ty * TXp + tx] = v;
  __syncthreads();
#pragma unroll 1
  for (int off = TY >> 1; off > 0; off >>= 1) {
    if (ty < off) {
      float4 a = sm[ty * TXp + tx], b = sm[(ty + off) * TXp + tx];
      sm[ty * TXp + tx] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    __syncthreads();
  }
  v = sm[tx];
}
__device__ __forceinline__ void block_sum8(float4* sm, F8& v, int tx, int ty, int TXp, int TY) {
  float4 lo = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]), hi = make_float4(v.v[4], v.v[5], v.v[6], v.v[7]);
  block_sum4(sm, lo, tx, ty, TXp, TY);
  block_sum4(sm, hi, tx, ty, TXp, TY);
  v.v[0] = lo.x, v.v[1] = lo.y, v.v[2] = lo.z, v.v[3] = lo.w, v.v[4] = hi.x, v.v[5] = hi.y, v.v[6] = hi.z, v.v[7] = hi.w;
}
constexpr int kRPT = 4;   // rows per thread of the reduction kernels (all loads issued before use)

template <typename T>
__global__ void __launch_bounds__(kRowThreads) bn_stats_kernel(b2h_bn_stats_t d, int ctas_per_group) {
  pdl_sync();
  __shared__ float4 s_red[kRowThreads];
  const int tx = threadIdx.x, ty = threadIdx.y, TXp = blockDim.x, TY = blockDim.y;
  const int g = blockIdx.y;
  const int c0 = tx * 8;
  const int rpg = d.rows_per_group;
  const int row0 = (blockIdx.x * TY + ty) * kRPT;   // within the group
  const T* z = reinterpret_cast<const T*>(d.z) + (int64_t)g * rpg * d.ld + c0;
  double* accum = reinterpret_cast<double*>(d.partial);   // [groups][C][2], zero between launches
  F8 s1 = zero8(), s2 = zero8();
  if (c0 < d.C && row0 < rpg) {
    F8 piv = zero8();
    if (d.running_mean) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c0 + i < d.C) piv.v[i] = __ldg(d.running_mean + c0 + i);
    }
    F8 v[kRPT];
#pragma unroll
    for (int u = 0; u < kRPT; ++u) v[u] = (row0 + u < rpg) ? load8<T>(z + (int64_t)(row0 + u) * d.ld) : piv;
#pragma unroll
    for (int u = 0; u < kRPT; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dx = v[u].v[i] - piv.v[i];
        s1.v[i] += dx;
        s2.v[i] = fmaf(dx, dx, s2.v[i]);
      }
    }
  }
  block_sum8(s_red, s1, tx, ty, TXp, TY);
  block_sum8(s_red, s2, tx, ty, TXp, TY);
  if (ty == 0 && c0 < d.C) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (c0 + i < d.C) {
        bn_stats_accumulate(d, blockIdx.x % kCopies, g, c0 + i, s1.v[i], s2.v[i]);
      }
    }
  }
  if (!last_block_done(d.ticket, gridDim.x * gridDim.y)) return;
  bn_stats_finalize(d, ty * TXp + tx, kRowThreads);
}

int64_t bn_partial_floats(int rows, int C, int groups) {
  (void)rows;
  return (int64_t)128 * (groups > 0 ? groups : 1) * C * 2;
}

struct RowGrid {
  int txp, ty, ctas;
};
static RowGrid row_grid(int Cwork, int rows) {
  RowGrid r;
  r.txp = std::min(pow2_ceil(ceil_div(Cwork, 8)), kRowThreads / 2);
  r.ty = kRowThreads / r.txp;
  r.ctas = ceil_div(ceil_div(rows, kRPT), r.ty);
  return r;
}

int launch_bn_stats(const b2h_bn_stats_t& d, int dtype, cudaStream_t s) {
  B2H_CARVE(bn_stats_kernel<__nv_bfloat16>);
  B2H_CARVE(bn_stats_kernel<float>);
  B2H_CHECK_ARG(d.C > 0 && d.C <= 512 && d.groups >= 1 && d.rows_per_group > 0 && d.Cs >= d.C, B2H_ERR_SHAPE,
                "bn_stats: bad shape C=%d Cs=%d groups=%d rows=%d", d.C, d.Cs, d.groups, d.rows_per_group);
  B2H_CHECK_ARG(d.ld % 8 == 0 && d.ld >= ((d.C + 7) & ~7), B2H_ERR_ALIGN, "bn_stats: ld=%d C=%d", d.ld, d.C);
  B2H_CHECK_ARG(((uintptr_t)d.partial % 16) == 0, B2H_ERR_ALIGN, "bn_stats: workspace must be 16-byte aligned");
  RowGrid rg = row_grid(d.C, d.rows_per_group);
  dim3 grid(rg.ctas, d.groups), block(rg.txp, rg.ty);
  if (dtype == B2H_BF16)
    launch(bn_stats_kernel<__nv_bfloat16>, grid, block, 0, s, d, rg.ctas);
  else
    launch(bn_stats_kernel<float>, grid, block, 0, s, d, rg.ctas);
  B2H_LAUNCH_CHECK("bn_stats");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// bn_apply: out = dropout( BN0(src0) [+ BN1(src1)] ), zero fill of the channel padding
// ---------------------------------------------------------------------------------------------
// Thread (tx, ty) owns channels 8*tx.. of 4 consecutive rows (half a Philox block): ONE Philox call per thread,
// no integer division in the row loop for the regular row maps, 4 rows of loads in flight.
struct SrcRows {
  const void* z;
  int ld, rowmap, L_src;
  bool regular;   // source row is a pure function of the output row (no per-clip wrap needed)
};
template <bool SIMPLE>
__device__ __forceinline__ int64_t src_row_of(const SrcRows& s, int row, int L) {
  if (s.rowmap == B2H_ROW_IDENT) return row;
  if (SIMPLE) return row >> 1;
  if (s.regular) return s.rowmap == B2H_ROW_UP2 ? (row >> 1) : 2 * (int64_t)row;
  const int b = row / L, l = row - b * L;
  if (s.rowmap == B2H_ROW_UP2) return (int64_t)b * s.L_src + (l >> 1);
  if (s.rowmap == B2H_ROW_POOL2) return (int64_t)b * s.L_src + 2 * l;
  return b;  // BCAST
}

// SIMPLE: no pooled source, every source IDENT or a regular UP2, the 4 rows of a thread in one group (the common
// case; keeps the pooling / per-row division code and its registers out of the kernel).
template <typename T, bool SIMPLE>
__global__ void __launch_bounds__(256, 2) bn_apply_kernel(b2h_bn_apply_t d) {
  pdl_sync();
  const int tx = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
  const int c0 = tx * 8;
  if (c0 >= d.Cfill) return;
  const int rows = d.B * d.L;
  const int rpg = rows / d.groups;
  const int row0 = (blockIdx.x * TY + ty) * 4;   // 4 rows per thread = half of a Philox block of 8 rows
  if (row0 >= rows) return;
  DropCtx drop;
  drop.init(d.drop, d.drop_C);
  T* out = reinterpret_cast<T*>(d.out) + d.out_coff + c0;
  const bool live = c0 < d.C;
  const bool two = d.nsrc > 1;
  const int nvalid = min(8, d.C - c0);
  SrcRows sr[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const b2h_bn_src_t& src = d.src[k < d.nsrc ? k : 0];
    sr[k].z = reinterpret_cast<const T*>(src.z) + src.coff + c0;
    sr[k].ld = src.ld;
    sr[k].rowmap = src.rowmap;
    sr[k].L_src = src.L_src;
    sr[k].regular = (src.rowmap == B2H_ROW_UP2 && (d.L & 1) == 0 && src.L_src * 2 == d.L) ||
                    (src.rowmap == B2H_ROW_POOL2 && src.L_src == 2 * d.L);
  }
  const bool pool0 = !SIMPLE && sr[0].rowmap == B2H_ROW_POOL2;   // (a pooled SECOND source is rejected by the launcher)
  // folded affine of the group (a block of 8 rows never straddles groups when rows_per_group % 8 == 0)
  F8 s0 = zero8(), t0 = zero8(), s1 = zero8(), t1 = zero8();
  int gcur = -1;
  uint2 philox_bits = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
  if (live && drop.mode == B2H_DROP_PHILOX) philox_bits = drop.philox_rows8((uint32_t)row0 >> 3, (uint32_t)(d.drop_coff + c0) >> 3);
  {
    F8 za[4], zb[4], zc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = row0 + u;
      if (row < rows && live) {
        const T* p0 = reinterpret_cast<const T*>(sr[0].z) + src_row_of<SIMPLE>(sr[0], row, d.L) * sr[0].ld;
        za[u] = load8<T>(p0);
        if (pool0) zb[u] = load8<T>(p0 + sr[0].ld);
        if (two) {
          const T* p1 = reinterpret_cast<const T*>(sr[1].z) + src_row_of<SIMPLE>(sr[1], row, d.L) * sr[1].ld;
          zc[u] = load8<T>(p1);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = row0 + u;
      const int j = row & 7;
      if (row >= rows) break;
      F8 y = zero8();
      if (live) {
        const int g = d.groups == 1 ? 0 : (SIMPLE ? row0 / rpg : row / rpg);
        if (g != gcur) {
          gcur = g;
          s0 = ld8f(d.src[0].scale + g * d.src[0].Cs + d.src[0].coff + c0);
          t0 = ld8f(d.src[0].shift + g * d.src[0].Cs + d.src[0].coff + c0);
          if (two) {
            s1 = ld8f(d.src[1].scale + g * d.src[1].Cs + d.src[1].coff + c0);
            t1 = ld8f(d.src[1].shift + g * d.src[1].Cs + d.src[1].coff + c0);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) y.v[i] = fmaf(za[u].v[i], s0.v[i], t0.v[i]);
        if (pool0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float y2 = fmaf(zb[u].v[i], s0.v[i], t0.v[i]);
            y.v[i] = y2 > y.v[i] ? y2 : y.v[i];   // MaxPool1d: the second element wins only if strictly greater
          }
        }
        if (two) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            y.v[i] += fmaf(zc[u].v[i], s1.v[i], t1.v[i]);
          }
        }
        if (drop.mode != B2H_DROP_NONE) {
          uint32_t bits;
          if (drop.mode == B2H_DROP_PHILOX)
            bits = ((j < 4 ? philox_bits.x : philox_bits.y) >> ((j & 3) * 8)) & 0xFFu;
          else
            bits = drop.keep8_rc((uint32_t)row, (uint32_t)(d.drop_coff + c0));
#pragma unroll
          for (int i = 0; i < 8; ++i) y.v[i] = ((bits >> i) & 1u) ? 2.f * y.v[i] : 0.f;
          drop.save8((uint64_t)row * d.drop_C + d.drop_coff + c0, bits, nvalid);
        }
      }
      store8<T>(out + (int64_t)row * d.out_ld, y);
    }
  }
}

int launch_bn_apply(const b2h_bn_apply_t& d, int dtype, cudaStream_t s) {
  B2H_CARVE(bn_apply_kernel<__nv_bfloat16, true>);
  B2H_CARVE(bn_apply_kernel<__nv_bfloat16, false>);
  B2H_CARVE(bn_apply_kernel<float, true>);
  B2H_CARVE(bn_apply_kernel<float, false>);
  B2H_CHECK_ARG(d.nsrc >= 1 && d.nsrc <= 2 && d.C > 0 && d.Cfill >= d.C && d.Cfill <= 1024 && d.groups >= 1,
                B2H_ERR_SHAPE, "bn_apply: bad shape C=%d Cfill=%d nsrc=%d", d.C, d.Cfill, d.nsrc);
  B2H_CHECK_ARG(d.Cfill % 8 == 0 && d.out_ld % 8 == 0 && d.out_coff % 8 == 0 && d.drop_coff % 8 == 0, B2H_ERR_ALIGN,
                "bn_apply: alignment Cfill=%d ld=%d coff=%d", d.Cfill, d.out_ld, d.out_coff);
  B2H_CHECK_ARG((d.B * d.L) % d.groups == 0, B2H_ERR_SHAPE, "bn_apply: rows not divisible by groups");
  B2H_CHECK_ARG(d.nsrc == 1 || d.src[1].rowmap != B2H_ROW_POOL2, B2H_ERR_SHAPE,
                "bn_apply: a pooled source must be the first source");
  for (int i = 0; i < d.nsrc; ++i) {
    B2H_CHECK_ARG(d.src[i].ld % 8 == 0 && d.src[i].coff % 8 == 0 && d.src[i].Cs % 8 == 0, B2H_ERR_ALIGN,
                  "bn_apply: src alignment");
    B2H_CHECK_ARG(d.src[i].coff + ((d.C + 7) & ~7) <= d.src[i].Cs, B2H_ERR_SHAPE, "bn_apply: source channel range");
  }
  const int txp = std::min(pow2_ceil(ceil_div(d.Cfill, 8)), 128);
  const int ty = 256 / txp;
  const int rows = d.B * d.L;
  dim3 grid(ceil_div(ceil_div(rows, 4), ty)), block(txp, ty);
  bool simple = d.groups == 1 || (rows / d.groups) % 4 == 0;
  for (int i = 0; i < d.nsrc; ++i) {
    const b2h_bn_src_t& sc = d.src[i];
    simple = simple && (sc.rowmap == B2H_ROW_IDENT ||
                        (sc.rowmap == B2H_ROW_UP2 && (d.L & 1) == 0 && sc.L_src * 2 == d.L));
  }
  if (dtype == B2H_BF16) {
    if (simple)
      launch(bn_apply_kernel<__nv_bfloat16, true>, grid, block, 0, s, d);
    else
      launch(bn_apply_kernel<__nv_bfloat16, false>, grid, block, 0, s, d);
  } else {
    if (simple)
      launch(bn_apply_kernel<float, true>, grid, block, 0, s, d);
    else
      launch(bn_apply_kernel<float, false>, grid, block, 0, s, d);
  }
  B2H_LAUNCH_CHECK("bn_apply");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// bn_bwd: two passes (reduce, apply); dy is recomputed from the gradient sources in both.
// Thread = 8 channels x 4 consecutive rows, all loads issued before use; per-CTA partial sums are combined with
// fp64 atomics over a few accumulator copies (see the kernel comment below for who sums / re-zeroes them).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ F8 load_g8(const b2h_grad_src_t& gs, int64_t row, int c0) {
  if (gs.f32) return load8<float>(reinterpret_cast<const float*>(gs.g) + row * gs.ld + gs.coff + c0);
  return load8<T>(reinterpret_cast<const T*>(gs.g) + row * gs.ld + gs.coff + c0);
}

// IDENT or regular UP2 (even consumer length == 2*L): the source row is a pure function of `row`
template <typename T>
__device__ __forceinline__ void add_grad_simple8(const b2h_grad_src_t& gs, int row, int c0, F8& dy) {
  if (gs.rowmap == B2H_ROW_IDENT) {
    const F8 g = load_g8<T>(gs, row, c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) dy.v[i] += g.v[i];
  } else {
    const F8 g0 = load_g8<T>(gs, 2 * (int64_t)row, c0), g1 = load_g8<T>(gs, 2 * (int64_t)row + 1, c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) dy.v[i] += g0.v[i] + g1.v[i];
  }
}

// contribution of ONE gradient source to dy(row, c0..c0+7); zown = z(row); sc/sh = forward affine of this layer.
// `reg`: the source row is a pure function of `row` (IDENT, or UP2 with an even consumer length == 2*L)
template <typename T>
__device__ __forceinline__ void add_grad_src8(const b2h_bn_bwd_t& d, const b2h_grad_src_t& gs, bool reg, const F8& sc,
                                              const F8& sh, const F8& zown, const F8& zpart, int row, int c0, F8& dy) {
  if (gs.rowmap == B2H_ROW_IDENT) {
    const F8 g = load_g8<T>(gs, row, c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) dy.v[i] += g.v[i];
    return;
  }
  if (gs.rowmap == B2H_ROW_UP2 && reg) {
    const F8 g0 = load_g8<T>(gs, 2 * (int64_t)row, c0), g1 = load_g8<T>(gs, 2 * (int64_t)row + 1, c0);
#pragma unroll
    for (int i = 0; i < 8; ++i) dy.v[i] += g0.v[i] + g1.v[i];
    return;
  }
  if (gs.rowmap == B2H_ROW_POOL2 && reg) {
    // regular pooling (L even, L_src == L/2): pooled row = row/2, pair partner = row^1 (a row this thread holds
    // already: zpart) — no per-row division, no second z load
    const F8& zp = zpart;
    const F8 g = load_g8<T>(gs, (int64_t)(row >> 1), c0);
    const bool even = (row & 1) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float yo = fmaf(zown.v[i], sc.v[i], sh.v[i]), yp = fmaf(zp.v[i], sc.v[i], sh.v[i]);
      const bool sel = even ? !(yp > yo) : (yo > yp);
      if (sel) dy.v[i] += g.v[i];
    }
    return;
  }
  const int b = row / d.L, l = row - b * d.L;
  if (gs.rowmap == B2H_ROW_UP2) {
    // the consumer read this tensor at row l' / 2 for l' in [0, L_src)
    const int64_t base = (int64_t)b * gs.L_src;
    if (2 * l < gs.L_src) {
      const F8 g = load_g8<T>(gs, base + 2 * l, c0);
#pragma unroll
      for (int i = 0; i < 8; ++i) dy.v[i] += g.v[i];
    }
    if (2 * l + 1 < gs.L_src) {
      const F8 g = load_g8<T>(gs, base + 2 * l + 1, c0);
#pragma unroll
      for (int i = 0; i < 8; ++i) dy.v[i] += g.v[i];
    }
    return;
  }
  // POOL2: dy(b, t) = g(b, t/2) if t is the (first) argmax of its pair
  const int lp = l >> 1;
  if (lp < gs.L_src) {
    const T* z = reinterpret_cast<const T*>(d.bn.z);
    const F8 zp = load8<T>(z + ((int64_t)b * d.L + (l ^ 1)) * d.bn.ld + d.bn.coff + c0);
    const F8 g = load_g8<T>(gs, (int64_t)b * gs.L_src + lp, c0);
    const bool even = (l & 1) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float yo = fmaf(zown.v[i], sc.v[i], sh.v[i]), yp = fmaf(zp.v[i], sc.v[i], sh.v[i]);
      const bool sel = even ? !(yp > yo) : (yo > yp);
      if (sel) dy.v[i] += g.v[i];
    }
  }
}

// Pass 1 only accumulates (fp64 atomics into kCopiesBwd1 copies, no ticket: its CTAs retire as soon as their
// reds are issued) — or is skipped altogether when the dgrad GEMMs that wrote the gradient sources produced the
// sums (b2h_gemm_t.bwd_sums -> d.accum).  Pass 2 starts by summing those copies for the channels of its CTA,
// writes dpre, accumulates the bias gradient the same way, and its LAST CTA (ticket) writes dgamma / dbeta /
// dbias / sums and re-zeroes both accumulator regions.  SIMPLE: every gradient source is IDENT or a regular UP2
// (no per-row div / pooling branches in the code).
constexpr int kCopiesBwd1 = B2H_BWD_COPIES;

template <typename T, int PASS, bool SIMPLE>
__global__ void __launch_bounds__(kRowThreads, 2) bn_bwd_kernel(b2h_bn_bwd_t d) {
  pdl_sync();
  __shared__ float4 s_red[kRowThreads];
  __shared__ float2 s_m[PASS == 2 ? 512 : 1];
  const int tx = threadIdx.x, ty = threadIdx.y, TXp = blockDim.x, TY = blockDim.y;
  const int tid = ty * TXp + tx;
  const int g = blockIdx.y;
  const int c0 = tx * 8;
  const int rows = d.B * d.L;
  const int rpg = rows / d.groups;
  const int r0 = (blockIdx.x * TY + ty) * kRPT;   // within the group
  const T* z = reinterpret_cast<const T*>(d.bn.z) + d.bn.coff + c0;
  T* dpre = reinterpret_cast<T*>(d.dpre) + c0;
  // accumulators, zero between launches: region 1 [kCopiesBwd1][groups][C][2] (sum dy, sum dy*zhat),
  // region 2 [kCopies][groups][C] (sum dpre)
  double* accum1 = d.accum ? d.accum : reinterpret_cast<double*>(d.partial);
  double* accum2 = reinterpret_cast<double*>(d.partial) + (int64_t)kCopiesBwd1 * d.groups * d.C * 2;
  if (PASS == 2) {
    const double inv_n = 1.0 / (double)rpg;
    for (int c = tid; c < d.C; c += kRowThreads) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int k = 0; k < kCopiesBwd1; ++k) {   // fixed order over the copies
        const double2 v = __ldcg(reinterpret_cast<const double2*>(accum1 + (((int64_t)k * d.groups + g) * d.C + c) * 2));
        a += v.x, b += v.y;
      }
      s_m[c] = make_float2((float)(a * inv_n), (float)(b * inv_n));
    }
    __syncthreads();
  }
  F8 acc_a = zero8(), acc_b = zero8();
  if (c0 < d.Cfill && r0 < rpg) {
    const bool live = c0 < d.C;
    const bool two = d.ngsrc > 1;
    const bool reg0 = d.gsrc[0].rowmap == B2H_ROW_IDENT ||
                      (d.gsrc[0].rowmap == B2H_ROW_UP2 && (d.gsrc[0].L_src & 1) == 0 && d.gsrc[0].L_src == 2 * d.L) ||
                      (d.gsrc[0].rowmap == B2H_ROW_POOL2 && (d.L & 1) == 0 && 2 * d.gsrc[0].L_src == d.L);
    const bool reg1 = d.gsrc[1].rowmap == B2H_ROW_IDENT ||
                      (d.gsrc[1].rowmap == B2H_ROW_UP2 && (d.gsrc[1].L_src & 1) == 0 && d.gsrc[1].L_src == 2 * d.L) ||
                      (d.gsrc[1].rowmap == B2H_ROW_POOL2 && (d.L & 1) == 0 && 2 * d.gsrc[1].L_src == d.L);
    // all per-channel arrays are zero in the channel padding -> no per-element guards below
    F8 sc = zero8(), sh = zero8(), mean8 = zero8(), istd8 = zero8(), mdy = zero8(), mdyz = zero8();
    if (live) {
      const int o = g * d.bn.Cs + c0;
      sc = ld8f(d.bn.scale + o);
      sh = ld8f(d.bn.shift + o);
      mean8 = ld8f(d.bn.mean + o);
      istd8 = ld8f(d.bn.invstd + o);
      if (PASS == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (c0 + i < d.C) {
            const float2 sm = s_m[c0 + i];
            mdy.v[i] = sm.x;
            mdyz.v[i] = sm.y;
          }
        }
      }
    }
    F8 zo[kRPT], dy[kRPT];
#pragma unroll
    for (int u = 0; u < kRPT; ++u) {
      zo[u] = dy[u] = zero8();
      if (r0 + u < rpg && live) zo[u] = load8<T>(z + (int64_t)(g * rpg + r0 + u) * d.bn.ld);
    }
#pragma unroll
    for (int u = 0; u < kRPT; ++u) {
      if (r0 + u < rpg && live) {
        const int row = g * rpg + r0 + u;
        if (SIMPLE) {
          add_grad_simple8<T>(d.gsrc[0], row, c0, dy[u]);
          if (two) add_grad_simple8<T>(d.gsrc[1], row, c0, dy[u]);
        } else {
          // (regular pooling: rows per group and r0 are even, so the pair partner row^1 is zo[u^1])
          add_grad_src8<T>(d, d.gsrc[0], reg0, sc, sh, zo[u], zo[u ^ 1], row, c0, dy[u]);
          if (two) add_grad_src8<T>(d, d.gsrc[1], reg1, sc, sh, zo[u], zo[u ^ 1], row, c0, dy[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kRPT; ++u) {
      if (r0 + u < rpg) {
        const int row = g * rpg + r0 + u;
        F8 out = zero8();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float zh = (zo[u].v[i] - mean8.v[i]) * istd8.v[i];
          if (PASS == 1) {
            acc_a.v[i] += dy[u].v[i];
            acc_b.v[i] = fmaf(dy[u].v[i], zh, acc_b.v[i]);
          } else {
            const float dz = sc.v[i] * (dy[u].v[i] - mdy.v[i] - zh * mdyz.v[i]);   // scale == 0 beyond C
            const float dp = dz * act_bwd(zo[u].v[i], d.act);
            out.v[i] = dp;
            acc_a.v[i] += dp;
          }
        }
        if (PASS == 2) store8<T>(dpre + (int64_t)row * d.ld_dpre, out);
      }
    }
  }
  if (PASS == 2 && d.defer == 1) return;   // dpre is all this launch owes the chain; b2h_colsum(bn_accum) finishes the rest
  block_sum8(s_red, acc_a, tx, ty, TXp, TY);
  if (PASS == 1) block_sum8(s_red, acc_b, tx, ty, TXp, TY);
  if (ty == 0 && c0 < d.C) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (c0 + i < d.C) {
        if (PASS == 1) {
          double* a = accum1 + (((int64_t)(blockIdx.x % kCopiesBwd1) * d.groups + g) * d.C + c0 + i) * 2;
          atomicAdd(a + 0, (double)acc_a.v[i]);
          atomicAdd(a + 1, (double)acc_b.v[i]);
        } else {
          atomicAdd(accum2 + ((int64_t)(blockIdx.x % kCopies) * d.groups + g) * d.C + c0 + i, (double)acc_a.v[i]);
        }
      }
    }
  }
  if (PASS == 1) return;
  if (d.defer == 2) return;   // sums of dpre accumulated; no ticket, no serial tail: b2h_colsum(src = NULL) finishes
  if (!last_block_done(d.ticket, gridDim.x * gridDim.y)) return;
#pragma unroll 1
  for (int c = tid; c < d.C; c += kRowThreads) {
    double tot_a = 0.0, tot_b = 0.0, tot_c = 0.0;
#pragma unroll 1
    for (int gg = 0; gg < d.groups; ++gg) {
      double ta = 0.0, tb = 0.0;
#pragma unroll
      for (int k = 0; k < kCopiesBwd1; ++k) {
        double2* acc = reinterpret_cast<double2*>(accum1 + (((int64_t)k * d.groups + gg) * d.C + c) * 2);
        const double2 v = __ldcg(acc);
        *acc = make_double2(0.0, 0.0);
        ta += v.x, tb += v.y;
      }
#pragma unroll
      for (int k = 0; k < kCopies; ++k) {
        double* acc = accum2 + ((int64_t)k * d.groups + gg) * d.C + c;
        tot_c += __ldcg(acc);
        *acc = 0.0;
      }
      if (d.sums) {
        d.sums[((int64_t)gg * d.C + c) * 2 + 0] = (float)ta;
        d.sums[((int64_t)gg * d.C + c) * 2 + 1] = (float)tb;
      }
      tot_a += ta, tot_b += tb;
    }
    if (d.dbeta) d.dbeta[c] = (float)tot_a;
    if (d.dgamma) d.dgamma[c] = (float)tot_b;
    if (d.dbias) d.dbias[c] = (float)tot_c;
  }
}

static bool grad_src_simple(const b2h_grad_src_t& gs, int L) {
  return gs.rowmap == B2H_ROW_IDENT || (gs.rowmap == B2H_ROW_UP2 && (gs.L_src & 1) == 0 && gs.L_src == 2 * L);
}

template <typename T>
static void launch_bn_bwd_t(const b2h_bn_bwd_t& d, dim3 grid, dim3 block, bool simple, int pass, cudaStream_t s) {
  if (pass == 1) {
    if (simple)
      launch(bn_bwd_kernel<T, 1, true>, grid, block, 0, s, d);
    else
      launch(bn_bwd_kernel<T, 1, false>, grid, block, 0, s, d);
  } else {
    if (simple)
      launch(bn_bwd_kernel<T, 2, true>, grid, block, 0, s, d);
    else
      launch(bn_bwd_kernel<T, 2, false>, grid, block, 0, s, d);
  }
}

static int launch_bn_bwd_passes(const b2h_bn_bwd_t& d, int dtype, int first_pass, int last_pass, cudaStream_t s) {
  B2H_CARVE(bn_bwd_kernel<__nv_bfloat16, 1, true>);
  B2H_CARVE(bn_bwd_kernel<__nv_bfloat16, 2, true>);
  B2H_CARVE(bn_bwd_kernel<__nv_bfloat16, 1, false>);
  B2H_CARVE(bn_bwd_kernel<__nv_bfloat16, 2, false>);
  B2H_CARVE(bn_bwd_kernel<float, 1, true>);
  B2H_CARVE(bn_bwd_kernel<float, 2, true>);
  B2H_CARVE(bn_bwd_kernel<float, 1, false>);
  B2H_CARVE(bn_bwd_kernel<float, 2, false>);
  B2H_CHECK_ARG(d.C > 0 && d.C <= 512 && d.Cfill >= d.C && d.Cfill <= 1024 && d.groups >= 1 && d.ngsrc >= 1 &&
                    d.ngsrc <= 2,
                B2H_ERR_SHAPE, "bn_bwd: bad shape C=%d Cfill=%d ngsrc=%d", d.C, d.Cfill, d.ngsrc);
  B2H_CHECK_ARG(d.Cfill % 8 == 0 && d.ld_dpre % 8 == 0 && d.bn.ld % 8 == 0 && d.bn.coff == 0 && d.bn.Cs % 8 == 0 &&
                    d.bn.Cs >= d.Cfill && ((uintptr_t)d.partial % 16) == 0 && ((uintptr_t)d.accum % 16) == 0,
                B2H_ERR_ALIGN, "bn_bwd: alignment / padded per-channel arrays");
  for (int i = 0; i < d.ngsrc; ++i)
    B2H_CHECK_ARG(d.gsrc[i].ld % 8 == 0 && d.gsrc[i].coff % 8 == 0, B2H_ERR_ALIGN, "bn_bwd: grad source alignment");
  B2H_CHECK_ARG((d.B * d.L) % d.groups == 0, B2H_ERR_SHAPE, "bn_bwd: rows not divisible by groups");
  B2H_CHECK_ARG(d.bn.mean && d.bn.invstd && d.bn.scale && d.bn.shift, B2H_ERR_ARG,
                "bn_bwd: needs the batch statistics of the forward pass");
  const int rpg = d.B * d.L / d.groups;
  RowGrid rg = row_grid(d.Cfill, rpg);
  dim3 grid(rg.ctas, d.groups), block(rg.txp, rg.ty);
  bool simple = grad_src_simple(d.gsrc[0], d.L) && (d.ngsrc < 2 || grad_src_simple(d.gsrc[1], d.L));
  for (int pass = first_pass; pass <= last_pass; ++pass) {
    if (dtype == B2H_BF16)
      launch_bn_bwd_t<__nv_bfloat16>(d, grid, block, simple, pass, s);
    else
      launch_bn_bwd_t<float>(d, grid, block, simple, pass, s);
    B2H_LAUNCH_CHECK(pass == 1 ? "bn_bwd pass 1" : "bn_bwd pass 2");
  }
  return B2H_OK;
}

int launch_bn_bwd(const b2h_bn_bwd_t& d, int dtype, cudaStream_t s) {
  B2H_CHECK_ARG(d.defer >= 0 && d.defer <= 2 && (!d.defer || (d.accum && d.dpre && !d.first_pass_only)), B2H_ERR_ARG,
                "bn_bwd: defer (0, 1, 2) needs the first-pass accumulators and dpre");
  B2H_CHECK_ARG(d.defer != 2 || d.partial, B2H_ERR_ARG, "bn_bwd: defer = 2 accumulates the sums of dpre in `partial`");
  B2H_CHECK_ARG(!d.first_pass_only || (d.accum && !d.dpre), B2H_ERR_ARG,
                "bn_bwd: first_pass_only accumulates into `accum` and writes no dpre");
  if (d.first_pass_only) {
    b2h_bn_bwd_t f = d;
    f.partial = reinterpret_cast<float*>(d.accum);   // (unused by the first pass when accum is given)
    f.ld_dpre = 8;
    return launch_bn_bwd_passes(f, dtype, 1, 1, s);
  }
  // with `accum` the first pass was produced by the GEMMs that wrote the gradient sources
  return launch_bn_bwd_passes(d, dtype, d.accum ? 2 : 1, 2, s);
}

// b2h_gemm_t.bwd_sums for GEMMs whose epilogue cannot produce it (fp32 path, unsupported tilings): the first
// pass of bn_bwd restricted to this GEMM's output as the only gradient source.
int launch_bwd_sums_separate(const b2h_gemm_t& g, int dtype, cudaStream_t s) {
  const b2h_bwd_sums_t& bs = g.bwd_sums;
  const bool pooled = bs.rowmap == B2H_ROW_POOL2;
  B2H_CHECK_ARG(bs.rowmap == B2H_ROW_IDENT || bs.rowmap == B2H_ROW_UP2 || (pooled && bs.scale && bs.shift), B2H_ERR_ARG,
                "bwd_sums: rowmap (POOL2 needs the producer's scale / shift)");
  B2H_CHECK_ARG(bs.accum && bs.mean && bs.invstd && bs.C == g.Nvalid && !g.out_f32, B2H_ERR_ARG,
                "bwd_sums: must describe the layer that produced the rows this GEMM differentiates");
  b2h_bn_bwd_t d;
  memset(&d, 0, sizeof(d));
  d.gsrc[0].g = g.out;
  d.gsrc[0].ld = g.ldo;
  d.gsrc[0].coff = g.out_coff;
  d.gsrc[0].rowmap = bs.rowmap;
  d.gsrc[0].L_src = g.Lo_actual;
  d.ngsrc = 1;
  d.bn.z = bs.z;
  d.bn.ld = bs.ld;
  d.bn.Cs = bs.Cs;
  d.bn.mean = bs.mean;
  d.bn.invstd = bs.invstd;
  d.bn.scale = pooled ? bs.scale : bs.mean;   // (only read by the pooling path)
  d.bn.shift = pooled ? bs.shift : bs.mean;
  d.ld_dpre = 8;
  d.Cfill = (bs.C + 7) & ~7;
  d.B = g.B;
  d.L = bs.Lz;
  d.C = bs.C;
  d.groups = bs.groups;
  d.accum = bs.accum;
  d.partial = reinterpret_cast<float*>(bs.accum);
  B2H_CHECK_ARG(grad_src_simple(d.gsrc[0], d.L) || (pooled && (d.L & 1) == 0 && 2 * g.Lo_actual == d.L), B2H_ERR_SHAPE,
                "bwd_sums: Lo_actual=%d does not match Lz=%d", g.Lo_actual, bs.Lz);
  return launch_bn_bwd_passes(d, dtype, 1, 1, s);
}

// ---------------------------------------------------------------------------------------------
// colsum: out[c] = sum over rows of src[row][c] (bias gradients of layers without BN); same scheme
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kRowThreads) colsum_kernel(b2h_colsum_t d) {
  pdl_sync();
  __shared__ float4 s_red[kRowThreads];
  const int tx = threadIdx.x, ty = threadIdx.y, TXp = blockDim.x, TY = blockDim.y;
  const int c0 = tx * 8;
  const int r0 = (blockIdx.x * TY + ty) * kRPT;
  double* accum = reinterpret_cast<double*>(d.partial);   // [C], zero between launches
  F8 acc = zero8();
  if (c0 < d.C && r0 < d.rows) {
    const T* src = reinterpret_cast<const T*>(d.src) + c0;
    F8 v[kRPT];
#pragma unroll
    for (int u = 0; u < kRPT; ++u) v[u] = (r0 + u < d.rows) ? load8<T>(src + (int64_t)(r0 + u) * d.ld) : zero8();
#pragma unroll
    for (int u = 0; u < kRPT; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] += v[u].v[i];
  }
  block_sum8(s_red, acc, tx, ty, TXp, TY);
  if (ty == 0 && c0 < d.C) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c0 + i < d.C) atomicAdd(accum + (int64_t)(blockIdx.x % kCopies) * d.C + c0 + i, (double)acc.v[i]);
  }
  if (!last_block_done(d.ticket, gridDim.x)) return;
#pragma unroll 1
  for (int c = ty * TXp + tx; c < d.C; c += kRowThreads) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kCopies; ++k) {
      t += __ldcg(accum + (int64_t)k * d.C + c);
      accum[(int64_t)k * d.C + c] = 0.0;
    }
    d.out[c] = (float)t;
    if (d.bn_accum) {   // deferred BatchNorm backward: dbeta / dgamma from the first-pass sums, accumulators re-zeroed
      double tot_a = 0.0, tot_b = 0.0;
#pragma unroll 1
      for (int gg = 0; gg < d.bn_groups; ++gg) {
        double ta = 0.0, tb = 0.0;
#pragma unroll
        for (int k = 0; k < B2H_BWD_COPIES; ++k) {   // fixed order over the copies, as bn_bwd's own tail
          double2* acc = reinterpret_cast<double2*>(d.bn_accum + (((int64_t)k * d.bn_groups + gg) * d.C + c) * 2);
          const double2 v = __ldcg(acc);
          *acc = make_double2(0.0, 0.0);
          ta += v.x, tb += v.y;
        }
        tot_a += ta, tot_b += tb;
      }
      if (d.dbeta) d.dbeta[c] = (float)tot_a;
      if (d.dgamma) d.dgamma[c] = (float)tot_b;
    }
  }
}

// Tail of a BatchNorm backward whose bn_bwd launch ran with defer = 2 (it accumulated the sums of dpre in its own
// `partial` and skipped the ticket / last-CTA pass): one CTA sums the accumulator copies in the same fixed order as
// bn_bwd's own tail, writes dbeta / dgamma / dbias and re-zeroes both regions.  Runs beside the backward chain.
__global__ void __launch_bounds__(256) bn_bwd_finish_kernel(b2h_colsum_t d) {
  pdl_sync();
  double* accum1 = d.bn_accum;
  double* accum2 = reinterpret_cast<double*>(d.partial) + (int64_t)B2H_BWD_COPIES * d.bn_groups * d.C * 2;
#pragma unroll 1
  for (int c = threadIdx.x; c < d.C; c += 256) {
    double tot_a = 0.0, tot_b = 0.0, tot_c = 0.0;
#pragma unroll 1
    for (int gg = 0; gg < d.bn_groups; ++gg) {
      double ta = 0.0, tb = 0.0;
#pragma unroll
      for (int k = 0; k < B2H_BWD_COPIES; ++k) {
        double2* acc = reinterpret_cast<double2*>(accum1 + (((int64_t)k * d.bn_groups + gg) * d.C + c) * 2);
        const double2 v = __ldcg(acc);
        *acc = make_double2(0.0, 0.0);
        ta += v.x, tb += v.y;
      }
#pragma unroll
      for (int k = 0; k < kCopies; ++k) {
        double* acc = accum2 + ((int64_t)k * d.bn_groups + gg) * d.C + c;
        tot_c += __ldcg(acc);
        *acc = 0.0;
      }
      tot_a += ta, tot_b += tb;
    }
    if (d.dbeta) d.dbeta[c] = (float)tot_a;
    if (d.dgamma) d.dgamma[c] = (float)tot_b;
    if (d.out) d.out[c] = (float)tot_c;
  }
}

int launch_colsum(const b2h_colsum_t& d, int dtype, cudaStream_t s) {
  if (!d.src) {   // no rows to sum: the finishing launch of a bn_bwd with defer = 2
    B2H_CARVE(bn_bwd_finish_kernel);
    B2H_CHECK_ARG(d.bn_accum && d.partial && d.bn_groups >= 1 && d.C > 0 && d.C <= 512 &&
                      ((uintptr_t)d.bn_accum % 16) == 0 && ((uintptr_t)d.partial % 8) == 0,
                  B2H_ERR_ARG, "colsum: src = NULL finishes a deferred bn_bwd and needs bn_accum, partial, bn_groups");
    launch(bn_bwd_finish_kernel, 1, 256, 0, s, d);
    B2H_LAUNCH_CHECK("bn_bwd_finish");
    return B2H_OK;
  }
  B2H_CARVE(colsum_kernel<__nv_bfloat16>);
  B2H_CARVE(colsum_kernel<float>);
  B2H_CHECK_ARG(d.C > 0 && d.C <= 512 && d.rows > 0 && d.ld % 8 == 0 && d.ld >= ((d.C + 7) & ~7), B2H_ERR_SHAPE,
                "colsum: bad shape");
  B2H_CHECK_ARG(((uintptr_t)d.partial % 8) == 0, B2H_ERR_ALIGN, "colsum: workspace must be 8-byte aligned");
  B2H_CHECK_ARG(!d.bn_accum || (d.bn_groups >= 1 && ((uintptr_t)d.bn_accum % 16) == 0), B2H_ERR_ARG,
                "colsum: bn_accum needs bn_groups >= 1 and a 16-byte aligned accumulator");
  RowGrid rg = row_grid(d.C, d.rows);
  dim3 grid(rg.ctas), block(rg.txp, rg.ty);
  if (dtype == B2H_BF16 && !d.f32)
    launch(colsum_kernel<__nv_bfloat16>, grid, block, 0, s, d);
  else
    launch(colsum_kernel<float>, grid, block, 0, s, d);
  B2H_LAUNCH_CHECK("colsum");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// bn_fold: eval-mode BN as a per-channel scale/shift (consumed by bn_apply sources and GEMM epilogues)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bn_fold_body(const b2h_bn_fold_t& d, int c) {
  if (c >= d.Cpad) return;
  float s = 0.f, t = 0.f;
  if (c < d.C) {
    float invstd = 1.0f / sqrtf(d.running_var[c] + d.eps);
    s = invstd * (d.gamma ? d.gamma[c] : 1.f);
    t = (d.beta ? d.beta[c] : 0.f) - d.running_mean[c] * s;
  }
  d.scale[c] = s;
  d.shift[c] = t;
}
__global__ void bn_fold_kernel(b2h_bn_fold_t d) {
  pdl_sync();
  bn_fold_body(d, blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void bn_fold_multi_kernel(b2h_bn_fold_multi_t m) {
  pdl_sync();
  __shared__ b2h_bn_fold_t d;
  if (threadIdx.x < sizeof(b2h_bn_fold_t) / 4)
    reinterpret_cast<uint32_t*>(&d)[threadIdx.x] = reinterpret_cast<const uint32_t*>(m.descs + blockIdx.y)[threadIdx.x];
  __syncthreads();
  bn_fold_body(d, blockIdx.x * blockDim.x + threadIdx.x);
}

int launch_bn_fold(const b2h_bn_fold_t& d, cudaStream_t s) {
  B2H_CARVE(bn_fold_kernel);
  B2H_CHECK_ARG(d.C > 0 && d.Cpad >= d.C, B2H_ERR_SHAPE, "bn_fold: bad shape");
  launch(bn_fold_kernel, ceil_div(d.Cpad, 128), 128, 0, s, d);
  B2H_LAUNCH_CHECK("bn_fold");
  return B2H_OK;
}

int launch_bn_fold_multi(const b2h_bn_fold_multi_t& m, cudaStream_t s) {
  B2H_CARVE(bn_fold_multi_kernel);
  static_assert(sizeof(b2h_bn_fold_t) % 4 == 0 && sizeof(b2h_bn_fold_t) / 4 <= 128, "descriptor staging");
  B2H_CHECK_ARG(m.descs && m.n > 0 && m.n <= 65535 && m.max_cpad > 0, B2H_ERR_ARG, "bn_fold_multi: bad args");
  launch(bn_fold_multi_kernel, dim3(ceil_div(m.max_cpad, 128), m.n), 128, 0, s, m);
  B2H_LAUNCH_CHECK("bn_fold_multi");
  return B2H_OK;
}

}  // namespace b2h
