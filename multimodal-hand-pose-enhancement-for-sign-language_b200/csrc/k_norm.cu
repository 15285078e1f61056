// BatchNorm1d statistics / apply / backward and column sums (HBM/L2-bound elementwise + reduction
// kernels).  Layout: rows x ld, channels contiguous, 4 channels per thread (8 B bf16 / 16 B fp32
// vector access), CTA = (TXp, TY) threads covering kBnChunkRows rows.
//
// Reference semantics restated (modelZoo.py:192-198, SURVEY.md section 8a "PyTorch semantics"):
//   train: y = (z - mean_b) / sqrt(var_b(biased) + eps) * gamma + beta;
//          running = (1-m)*running + m*batch (running_var from the UNBIASED batch variance);
//   eval:  y = (z - running_mean) / sqrt(running_var + eps) * gamma + beta.
#include <algorithm>

#include "b2h_common.cuh"

namespace b2h {

// Thread geometry shared by the kernels below: a CTA is (TXp, TY) threads, thread (tx, ty) owns channels
// 4*tx .. 4*tx+3 and rows ty, ty+TY, ... of the CTA's row chunk.  Reduction kernels use 512 threads and at
// most kMaxChunks chunks per group so that the ordered final merge by the last CTA stays short.
constexpr int kRedThreads = 512;
constexpr int kMaxChunks = 64;
struct RedShape {
  int txp, ty, rows_per_chunk, nchunks;
};
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
static RedShape red_shape(int Cwork, int rows_per_group, int threads) {
  RedShape r;
  r.txp = std::min(pow2_ceil(ceil_div(Cwork, 4)), threads / 2);
  r.ty = threads / r.txp;
  int R = std::max(r.ty * 4, ceil_div(ceil_div(rows_per_group, kMaxChunks), r.ty) * r.ty);
  r.rows_per_chunk = R;
  r.nchunks = ceil_div(rows_per_group, R);
  return r;
}

// sum the per-thread accumulators over ty (fixed pairing tree -> deterministic); result valid for ty == 0
__device__ __forceinline__ void reduce_over_ty(float4* sm, float4& v, int tx, int ty, int TXp, int TY) {
  __syncthreads();
  sm[ty * TXp + tx] = v;
  __syncthreads();
  for (int off = TY >> 1; off > 0; off >>= 1) {
    if (ty < off) {
      float4 a = sm[ty * TXp + tx], b = sm[(ty + off) * TXp + tx];
      sm[ty * TXp + tx] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    __syncthreads();
  }
  v = sm[tx];
}

// lane-strided sum over chunks of the float2 partial[(chunk*groups + g)*C + c], 8 independent loads in flight
template <typename F>
__device__ __forceinline__ void lane_chunk_loop(const float* partial, int nchunks, int groups, int g, int C, int c,
                                                int lane, int NL, F&& f) {
  int ch = lane;
  for (; ch + 7 * NL < nchunks; ch += 8 * NL) {
    float2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      v[u] = __ldcg(reinterpret_cast<const float2*>(partial + (((int64_t)(ch + u * NL) * groups + g) * C + c) * 2));
#pragma unroll
    for (int u = 0; u < 8; ++u) f(ch + u * NL, v[u]);
  }
  for (; ch < nchunks; ch += NL)
    f(ch, __ldcg(reinterpret_cast<const float2*>(partial + (((int64_t)ch * groups + g) * C + c) * 2)));
}

// ordered sum over chunks of partial[(chunk*groups + g)*C + c][which] by the last CTA:
// thread (c, lane) takes chunks lane, lane+NL, ...; lanes are then added in order by lane 0.
struct FinalLanes {
  int Cp2, NL, c, lane;
  __device__ __forceinline__ FinalLanes(int C, int tid) {
    Cp2 = 1;
    while (Cp2 < C) Cp2 <<= 1;
    NL = kRedThreads / Cp2;
    if (NL < 1) NL = 1;
    c = tid % Cp2;
    lane = tid / Cp2;
  }
};

struct Affine4 {
  float4 s, t;
};

// per-channel scale/shift of a BN source for channels c0..c0+3 (guarded by C)
__device__ __forceinline__ Affine4 bn_affine(const b2h_bn_src_t& src, int g, int C, int c0) {
  Affine4 a;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = c0 + i;
    float s = 0.f, t = 0.f;
    if (c < C) {
      const int cs = src.coff + c;  // channel index inside the source layer
      float gamma = src.gamma ? src.gamma[cs] : 1.f;
      float beta = src.beta ? src.beta[cs] : 0.f;
      float mean, invstd;
      if (src.use_running) {
        mean = src.running_mean[cs];
        invstd = 1.0f / sqrtf(src.running_var[cs] + src.eps);
      } else {
        mean = src.mean[g * src.C_total + cs];
        invstd = src.invstd[g * src.C_total + cs];
      }
      s = invstd * gamma;
      t = beta - mean * s;
    }
    f4(a.s, i) = s;
    f4(a.t, i) = t;
  }
  return a;
}

__device__ __forceinline__ float4 fma4(float4 z, const Affine4& a) {
  return make_float4(fmaf(z.x, a.s.x, a.t.x), fmaf(z.y, a.s.y, a.t.y), fmaf(z.z, a.s.z, a.t.z),
                     fmaf(z.w, a.s.w, a.t.w));
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// BN(src) evaluated at output position (b, l) for channels c0..c0+3
template <typename T>
__device__ __forceinline__ float4 bn_src_eval(const b2h_bn_src_t& src, const Affine4& a, int b, int l, int c0) {
  const T* z = reinterpret_cast<const T*>(src.z);
  const int64_t base = (int64_t)b * src.L_src;
  switch (src.rowmap) {
    case B2H_ROW_UP2:
      return fma4(load4<T>(z + (base + (l >> 1)) * src.ld + src.coff + c0), a);
    case B2H_ROW_POOL2: {
      float4 y0 = fma4(load4<T>(z + (base + 2 * l) * src.ld + src.coff + c0), a);
      float4 y1 = fma4(load4<T>(z + (base + 2 * l + 1) * src.ld + src.coff + c0), a);
      // MaxPool1d: the second element replaces the first only if strictly greater
      return make_float4(y1.x > y0.x ? y1.x : y0.x, y1.y > y0.y ? y1.y : y0.y, y1.z > y0.z ? y1.z : y0.z,
                         y1.w > y0.w ? y1.w : y0.w);
    }
    case B2H_ROW_BCAST:
      return fma4(load4<T>(z + (int64_t)b * src.ld + src.coff + c0), a);
    default:
      return fma4(load4<T>(z + (base + l) * src.ld + src.coff + c0), a);
  }
}

// ---------------------------------------------------------------------------------------------
// bn_stats: per-chunk shifted sums (pivot = first row of the chunk) -> (mean, M2) per chunk ->
// ordered, division-free merge in double by the last CTA (deterministic)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kRedThreads) bn_stats_kernel(b2h_bn_stats_t d, int nchunks, int R) {
  __shared__ float4 s_red[kRedThreads];
  __shared__ double s_dbl[kRedThreads];
  __shared__ double s_mean[kRedThreads];
  const int tx = threadIdx.x, ty = threadIdx.y, TXp = blockDim.x, TY = blockDim.y;
  const int chunk = blockIdx.x, g = blockIdx.y;
  const int c0 = tx * 4;
  const int rpg = d.rows_per_group;
  const int r_begin = chunk * R;
  const int r_end = min(r_begin + R, rpg);
  const T* z = reinterpret_cast<const T*>(d.z) + (int64_t)g * rpg * d.ld + c0;
  float4 s1 = make_float4(0, 0, 0, 0), s2 = s1, piv = s1;
  if (c0 < d.C) {
    piv = load4<T>(z + (int64_t)r_begin * d.ld);
    for (int r = r_begin + ty; r < r_end; r += 4 * TY) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int rr = r + u * TY;
        v[u] = rr < r_end ? load4<T>(z + (int64_t)rr * d.ld) : piv;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float dx = f4(v[u], i) - f4(piv, i);
          f4(s1, i) += dx;
          f4(s2, i) = fmaf(dx, dx, f4(s2, i));
        }
      }
    }
  }
  reduce_over_ty(s_red, s1, tx, ty, TXp, TY);
  reduce_over_ty(s_red, s2, tx, ty, TXp, TY);
  if (ty == 0 && c0 < d.C) {
    const float n = (float)(r_end - r_begin);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = c0 + i;
      if (c < d.C) {
        float m = f4(s1, i) / n;
        float* p = d.partial + (((int64_t)chunk * d.groups + g) * d.C + c) * 2;
        p[0] = f4(piv, i) + m;
        p[1] = fmaxf(f4(s2, i) - f4(s1, i) * m, 0.f);
      }
    }
  }
  if (!last_block_done(d.ticket, gridDim.x * gridDim.y)) return;
  const int tid = ty * TXp + tx;
  const FinalLanes fl(d.C, tid);
  const bool act = fl.c < d.C && fl.lane < fl.NL;
  for (int gg = 0; gg < d.groups; ++gg) {
    // single pass in double: S0 = sum n_c mean_c, S1 = sum (M2_c + n_c mean_c^2); M2 = S1 - S0^2 / N
    double s0 = 0.0, s1 = 0.0;
    if (act)
      lane_chunk_loop(d.partial, nchunks, d.groups, gg, d.C, fl.c, fl.lane, fl.NL, [&](int ch, float2 v) {
        const double nb = (double)(min(ch * R + R, rpg) - ch * R), m = (double)v.x;
        s0 += nb * m;
        s1 += (double)v.y + nb * m * m;
      });
    __syncthreads();
    s_dbl[tid] = s0;
    s_mean[tid] = s1;
    __syncthreads();
    if (act && fl.lane == 0) {
      double t0 = 0.0, t1 = 0.0;
      for (int l = 0; l < fl.NL; ++l) {
        t0 += s_dbl[l * fl.Cp2 + fl.c];
        t1 += s_mean[l * fl.Cp2 + fl.c];
      }
      const int c = fl.c;
      const double mean = t0 / (double)rpg;
      double m2 = t1 - t0 * mean;
      if (m2 < 0.0) m2 = 0.0;
      const double var_b = m2 / (double)rpg;
      d.mean[gg * d.C + c] = (float)mean;
      d.invstd[gg * d.C + c] = (float)(1.0 / sqrt(var_b + (double)d.eps));
      if (d.running_mean && (gg == 0 || d.update_all_groups)) {
        double var_u = rpg > 1 ? m2 / (double)(rpg - 1) : var_b;
        float mom = d.momentum;
        d.running_mean[c] = (1.f - mom) * d.running_mean[c] + mom * (float)mean;
        d.running_var[c] = (1.f - mom) * d.running_var[c] + mom * (float)var_u;
      }
    }
  }
  if (tid == 0 && d.running_mean && d.num_batches_tracked)
    *d.num_batches_tracked += d.update_all_groups ? d.groups : 1;
}

int64_t bn_partial_floats(int rows, int C, int groups) {
  (void)rows;
  return (int64_t)128 * (groups > 0 ? groups : 1) * C * 2;
}

int launch_bn_stats(const b2h_bn_stats_t& d, int dtype, cudaStream_t s) {
  B2H_CARVE(bn_stats_kernel<__nv_bfloat16>);
  B2H_CARVE(bn_stats_kernel<float>);
  B2H_CHECK_ARG(d.C > 0 && d.C <= 512 && d.groups >= 1 && d.rows_per_group > 0, B2H_ERR_SHAPE,
                "bn_stats: bad shape C=%d groups=%d rows=%d", d.C, d.groups, d.rows_per_group);
  B2H_CHECK_ARG(d.ld % 4 == 0 && d.ld >= ((d.C + 3) & ~3), B2H_ERR_ALIGN, "bn_stats: ld=%d C=%d", d.ld, d.C);
  RedShape rs = red_shape(d.C, d.rows_per_group, kRedThreads);
  dim3 grid(rs.nchunks, d.groups), block(rs.txp, rs.ty);
  if (dtype == B2H_BF16)
    bn_stats_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(d, rs.nchunks, rs.rows_per_chunk);
  else
    bn_stats_kernel<float><<<grid, block, 0, s>>>(d, rs.nchunks, rs.rows_per_chunk);
  B2H_LAUNCH_CHECK("bn_stats");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// bn_apply: out = dropout( BN0(src0) [+ BN1(src1)] ), zero fill of the channel padding.
// 256 threads, 4 rows per thread with all loads issued before the first store.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(b2h_bn_apply_t d) {
  const int tx = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
  const int c0 = tx * 4;
  if (c0 >= d.Cfill) return;
  const int rows = d.B * d.L;
  const int rpg = rows / d.groups;
  const int r0 = blockIdx.x * (TY * 4) + ty;
  DropCtx drop;
  drop.init(d.drop);
  T* out = reinterpret_cast<T*>(d.out);
  const bool live = c0 < d.C;
  float4 y[4];
  Affine4 a0, a1;
  int gcur = -1;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int row = r0 + u * TY;
    y[u] = make_float4(0, 0, 0, 0);
    if (row < rows && live) {
      const int b = row / d.L, l = row - b * d.L;
      const int g = row / rpg;
      if (g != gcur) {
        a0 = bn_affine(d.src[0], g, d.C, c0);
        if (d.nsrc > 1) a1 = bn_affine(d.src[1], g, d.C, c0);
        gcur = g;
      }
      y[u] = bn_src_eval<T>(d.src[0], a0, b, l, c0);
      if (d.nsrc > 1) y[u] = add4(y[u], bn_src_eval<T>(d.src[1], a1, b, l, c0));
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int row = r0 + u * TY;
    if (row >= rows) continue;
    if (live) {
      if (drop.mode != B2H_DROP_NONE) {
        const uint64_t di = (uint64_t)row * d.drop_C + d.drop_coff + c0;
        float4 m = drop.scale4(di);
        y[u].x *= m.x, y[u].y *= m.y, y[u].z *= m.z, y[u].w *= m.w;
        drop.save4(di, m, min(4, d.C - c0));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c0 + i >= d.C) f4(y[u], i) = 0.f;
    }
    store4<T>(out + (int64_t)row * d.out_ld + d.out_coff + c0, y[u]);
  }
}

int launch_bn_apply(const b2h_bn_apply_t& d, int dtype, cudaStream_t s) {
  B2H_CARVE(bn_apply_kernel<__nv_bfloat16>);
  B2H_CARVE(bn_apply_kernel<float>);
  B2H_CHECK_ARG(d.nsrc >= 1 && d.nsrc <= 2 && d.C > 0 && d.Cfill >= d.C && d.Cfill <= 1024 && d.groups >= 1,
                B2H_ERR_SHAPE, "bn_apply: bad shape C=%d Cfill=%d nsrc=%d", d.C, d.Cfill, d.nsrc);
  B2H_CHECK_ARG(d.Cfill % 4 == 0 && d.out_ld % 4 == 0 && d.out_coff % 4 == 0, B2H_ERR_ALIGN,
                "bn_apply: alignment Cfill=%d ld=%d coff=%d", d.Cfill, d.out_ld, d.out_coff);
  B2H_CHECK_ARG((d.B * d.L) % d.groups == 0, B2H_ERR_SHAPE, "bn_apply: rows not divisible by groups");
  for (int i = 0; i < d.nsrc; ++i) {
    B2H_CHECK_ARG(d.src[i].ld % 4 == 0 && d.src[i].coff % 4 == 0, B2H_ERR_ALIGN, "bn_apply: src alignment");
    B2H_CHECK_ARG(d.src[i].coff + d.C <= d.src[i].C_total, B2H_ERR_SHAPE, "bn_apply: source channel range");
  }
  int txp = std::min(pow2_ceil(ceil_div(d.Cfill, 4)), 256);
  int ty = 256 / txp;
  dim3 grid(ceil_div(d.B * d.L, ty * 4)), block(txp, ty);
  if (dtype == B2H_BF16)
    bn_apply_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(d);
  else
    bn_apply_kernel<float><<<grid, block, 0, s>>>(d);
  B2H_LAUNCH_CHECK("bn_apply");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// bn_bwd: two passes (reduce, apply); dy is recomputed from the gradient sources in both
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float4 load_g4(const b2h_grad_src_t& gs, int64_t row, int c0) {
  if (gs.f32) return load4<float>(reinterpret_cast<const float*>(gs.g) + row * gs.ld + gs.coff + c0);
  return load4<T>(reinterpret_cast<const T*>(gs.g) + row * gs.ld + gs.coff + c0);
}

// dy(b, l, c0..c0+3) of this layer's BN output; zown = z(b,l), aff = forward affine of this layer
template <typename T>
__device__ __forceinline__ float4 bn_bwd_dy(const b2h_bn_bwd_t& d, const Affine4& aff, float4 zown, int b, int l,
                                            int c0) {
  float4 dy = make_float4(0, 0, 0, 0);
  for (int s = 0; s < d.ngsrc; ++s) {
    const b2h_grad_src_t& gs = d.gsrc[s];
    const int64_t base = (int64_t)b * gs.L_src;
    if (gs.rowmap == B2H_ROW_UP2) {
      // consumer read this tensor at row l' / 2 for l' in [0, L_src)
      int l0 = 2 * l;
      if (l0 < gs.L_src) dy = add4(dy, load_g4<T>(gs, base + l0, c0));
      if (l0 + 1 < gs.L_src) dy = add4(dy, load_g4<T>(gs, base + l0 + 1, c0));
    } else if (gs.rowmap == B2H_ROW_POOL2) {
      int lp = l >> 1;
      if (lp < gs.L_src) {
        const T* z = reinterpret_cast<const T*>(d.bn.z);
        int lpart = l ^ 1;
        float4 zp = load4<T>(z + ((int64_t)b * d.L + lpart) * d.bn.ld + d.bn.coff + c0);
        float4 yo = fma4(zown, aff), yp = fma4(zp, aff);
        float4 g = load_g4<T>(gs, base + lp, c0);
        bool even = (l & 1) == 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          bool sel = even ? !(f4(yp, i) > f4(yo, i)) : (f4(yo, i) > f4(yp, i));
          if (sel) f4(dy, i) += f4(g, i);
        }
      }
    } else {
      dy = add4(dy, load_g4<T>(gs, base + l, c0));
    }
  }
  return dy;
}

template <typename T, int PASS>
__global__ void __launch_bounds__(kRedThreads) bn_bwd_kernel(b2h_bn_bwd_t d, int nchunks, int R) {
  __shared__ float4 s_red[kRedThreads];
  __shared__ double s_dbl[2][kRedThreads];
  const int tx = threadIdx.x, ty = threadIdx.y, TXp = blockDim.x, TY = blockDim.y;
  const int chunk = blockIdx.x, g = blockIdx.y;
  const int c0 = tx * 4;
  const int rows = d.B * d.L;
  const int rpg = rows / d.groups;
  const int r_begin = chunk * R;
  const int r_end = min(r_begin + R, rpg);
  const T* z = reinterpret_cast<const T*>(d.bn.z);
  T* dpre = reinterpret_cast<T*>(d.dpre);
  float4 acc_a = make_float4(0, 0, 0, 0), acc_b = make_float4(0, 0, 0, 0);
  if (c0 < d.Cfill) {
    Affine4 aff;
    float4 mean4 = make_float4(0, 0, 0, 0), istd4 = mean4, sg4 = mean4, mdy = mean4, mdyz = mean4;
    const bool live = c0 < d.C;
    if (live) {
      aff = bn_affine(d.bn, g, d.C, c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int c = c0 + i;
        if (c < d.C) {
          f4(mean4, i) = d.bn.mean[g * d.C + c];
          f4(istd4, i) = d.bn.invstd[g * d.C + c];
          f4(sg4, i) = f4(istd4, i) * (d.bn.gamma ? d.bn.gamma[c] : 1.f);
          if (PASS == 2) {
            const float* sm = d.sums + ((int64_t)g * d.C + c) * 2;
            f4(mdy, i) = sm[0] / (float)rpg;
            f4(mdyz, i) = sm[1] / (float)rpg;
          }
        }
      }
    }
    for (int r = r_begin + ty; r < r_end; r += 2 * TY) {
      float4 zo[2], dy[2];
      int rowv[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int rr = r + u * TY;
        rowv[u] = rr < r_end ? g * rpg + rr : -1;
        zo[u] = dy[u] = make_float4(0, 0, 0, 0);
        if (rowv[u] >= 0 && live) {
          const int b = rowv[u] / d.L, l = rowv[u] - b * d.L;
          zo[u] = load4<T>(z + (int64_t)rowv[u] * d.bn.ld + d.bn.coff + c0);
          dy[u] = bn_bwd_dy<T>(d, aff, zo[u], b, l, c0);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (rowv[u] < 0) continue;
        float4 out = make_float4(0, 0, 0, 0);
        if (live) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (c0 + i < d.C) {
              float zh = (f4(zo[u], i) - f4(mean4, i)) * f4(istd4, i);
              if (PASS == 1) {
                f4(acc_a, i) += f4(dy[u], i);
                f4(acc_b, i) = fmaf(f4(dy[u], i), zh, f4(acc_b, i));
              } else {
                float dz = f4(sg4, i) * (f4(dy[u], i) - f4(mdy, i) - zh * f4(mdyz, i));
                float dp = dz * act_bwd(f4(zo[u], i), d.act);
                f4(out, i) = dp;
                f4(acc_a, i) += dp;
              }
            }
          }
        }
        if (PASS == 2) store4<T>(dpre + (int64_t)rowv[u] * d.ld_dpre + c0, out);
      }
    }
  }
  reduce_over_ty(s_red, acc_a, tx, ty, TXp, TY);
  if (PASS == 1) reduce_over_ty(s_red, acc_b, tx, ty, TXp, TY);
  if (ty == 0 && c0 < d.C) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = c0 + i;
      if (c < d.C) {
        float* p = d.partial + (((int64_t)chunk * d.groups + g) * d.C + c) * 2;
        p[0] = f4(acc_a, i);
        p[1] = f4(acc_b, i);
      }
    }
  }
  if (!last_block_done(d.ticket, gridDim.x * gridDim.y)) return;
  const int tid = ty * TXp + tx;
  const FinalLanes fl(d.C, tid);
  const bool act = fl.c < d.C && fl.lane < fl.NL;
  double tot_a = 0.0, tot_b = 0.0;
  for (int gg = 0; gg < d.groups; ++gg) {
    double sa = 0.0, sb = 0.0;
    if (act)
      lane_chunk_loop(d.partial, nchunks, d.groups, gg, d.C, fl.c, fl.lane, fl.NL, [&](int, float2 v) {
        sa += (double)v.x;
        sb += (double)v.y;
      });
    __syncthreads();
    s_dbl[0][tid] = sa;
    s_dbl[1][tid] = sb;
    __syncthreads();
    if (act && fl.lane == 0) {
      double ta = 0.0, tb = 0.0;
      for (int l = 0; l < fl.NL; ++l) {
        ta += s_dbl[0][l * fl.Cp2 + fl.c];
        tb += s_dbl[1][l * fl.Cp2 + fl.c];
      }
      if (PASS == 1) {
        d.sums[((int64_t)gg * d.C + fl.c) * 2 + 0] = (float)ta;
        d.sums[((int64_t)gg * d.C + fl.c) * 2 + 1] = (float)tb;
      }
      tot_a += ta;
      tot_b += tb;
    }
  }
  if (act && fl.lane == 0) {
    if (PASS == 1) {
      if (d.dbeta) d.dbeta[fl.c] = (float)tot_a;
      if (d.dgamma) d.dgamma[fl.c] = (float)tot_b;
    } else {
      if (d.dbias) d.dbias[fl.c] = (float)tot_a;
    }
  }
}

int launch_bn_bwd(const b2h_bn_bwd_t& d, int dtype, cudaStream_t s) {
  B2H_CARVE(bn_bwd_kernel<__nv_bfloat16, 1>);
  B2H_CARVE(bn_bwd_kernel<__nv_bfloat16, 2>);
  B2H_CARVE(bn_bwd_kernel<float, 1>);
  B2H_CARVE(bn_bwd_kernel<float, 2>);
  B2H_CHECK_ARG(d.C > 0 && d.C <= 512 && d.Cfill >= d.C && d.Cfill <= 1024 && d.groups >= 1 && d.ngsrc >= 1 &&
                    d.ngsrc <= 2,
                B2H_ERR_SHAPE, "bn_bwd: bad shape C=%d Cfill=%d ngsrc=%d", d.C, d.Cfill, d.ngsrc);
  B2H_CHECK_ARG(d.Cfill % 4 == 0 && d.ld_dpre % 4 == 0 && d.bn.ld % 4 == 0 && d.bn.coff % 4 == 0, B2H_ERR_ALIGN,
                "bn_bwd: alignment");
  B2H_CHECK_ARG((d.B * d.L) % d.groups == 0, B2H_ERR_SHAPE, "bn_bwd: rows not divisible by groups");
  B2H_CHECK_ARG(!d.bn.use_running, B2H_ERR_ARG, "bn_bwd: backward is only defined for batch statistics");
  B2H_CHECK_ARG(d.bn.coff == 0 && d.bn.C_total == d.C, B2H_ERR_SHAPE, "bn_bwd: bn source must cover the whole layer");
  int rpg = d.B * d.L / d.groups;
  RedShape rs = red_shape(d.Cfill, rpg, kRedThreads);
  dim3 grid(rs.nchunks, d.groups), block(rs.txp, rs.ty);
  if (dtype == B2H_BF16) {
    bn_bwd_kernel<__nv_bfloat16, 1><<<grid, block, 0, s>>>(d, rs.nchunks, rs.rows_per_chunk);
    B2H_LAUNCH_CHECK("bn_bwd pass 1");
    bn_bwd_kernel<__nv_bfloat16, 2><<<grid, block, 0, s>>>(d, rs.nchunks, rs.rows_per_chunk);
  } else {
    bn_bwd_kernel<float, 1><<<grid, block, 0, s>>>(d, rs.nchunks, rs.rows_per_chunk);
    B2H_LAUNCH_CHECK("bn_bwd pass 1");
    bn_bwd_kernel<float, 2><<<grid, block, 0, s>>>(d, rs.nchunks, rs.rows_per_chunk);
  }
  B2H_LAUNCH_CHECK("bn_bwd pass 2");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// colsum: out[c] = sum over rows of src[row][c] (bias gradients of layers without BN)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kRedThreads) colsum_kernel(b2h_colsum_t d, int nchunks, int R) {
  __shared__ float4 s_red[kRedThreads];
  __shared__ double s_dbl[kRedThreads];
  const int tx = threadIdx.x, ty = threadIdx.y, TXp = blockDim.x, TY = blockDim.y;
  const int c0 = tx * 4;
  const int r_begin = blockIdx.x * R;
  const int r_end = min(r_begin + R, d.rows);
  float4 acc = make_float4(0, 0, 0, 0);
  if (c0 < d.C)
    for (int r = r_begin + ty; r < r_end; r += TY)
      acc = add4(acc, load4<T>(reinterpret_cast<const T*>(d.src) + (int64_t)r * d.ld + c0));
  reduce_over_ty(s_red, acc, tx, ty, TXp, TY);
  if (ty == 0 && c0 < d.C) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (c0 + i < d.C) d.partial[(int64_t)blockIdx.x * d.C + c0 + i] = f4(acc, i);
  }
  if (!last_block_done(d.ticket, gridDim.x)) return;
  const int tid = ty * TXp + tx;
  const FinalLanes fl(d.C, tid);
  const bool act = fl.c < d.C && fl.lane < fl.NL;
  double sa = 0.0;
  if (act)
    for (int ch = fl.lane; ch < nchunks; ch += fl.NL) sa += (double)__ldcg(d.partial + (int64_t)ch * d.C + fl.c);
  __syncthreads();
  s_dbl[tid] = sa;
  __syncthreads();
  if (act && fl.lane == 0) {
    double t = 0.0;
    for (int l = 0; l < fl.NL; ++l) t += s_dbl[l * fl.Cp2 + fl.c];
    d.out[fl.c] = (float)t;
  }
}

int launch_colsum(const b2h_colsum_t& d, int dtype, cudaStream_t s) {
  B2H_CARVE(colsum_kernel<__nv_bfloat16>);
  B2H_CARVE(colsum_kernel<float>);
  B2H_CHECK_ARG(d.C > 0 && d.C <= 512 && d.rows > 0 && d.ld % 4 == 0, B2H_ERR_SHAPE, "colsum: bad shape");
  RedShape rs = red_shape(d.C, d.rows, kRedThreads);
  dim3 grid(rs.nchunks), block(rs.txp, rs.ty);
  if (dtype == B2H_BF16 && !d.f32)
    colsum_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(d, rs.nchunks, rs.rows_per_chunk);
  else
    colsum_kernel<float><<<grid, block, 0, s>>>(d, rs.nchunks, rs.rows_per_chunk);
  B2H_LAUNCH_CHECK("colsum");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// bn_fold: eval-mode BN as a per-channel scale/shift for the GEMM epilogue
// ---------------------------------------------------------------------------------------------
__global__ void bn_fold_kernel(b2h_bn_fold_t d) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
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

int launch_bn_fold(const b2h_bn_fold_t& d, cudaStream_t s) {
  B2H_CARVE(bn_fold_kernel);
  B2H_CHECK_ARG(d.C > 0 && d.Cpad >= d.C, B2H_ERR_SHAPE, "bn_fold: bad shape");
  bn_fold_kernel<<<ceil_div(d.Cpad, 128), 128, 0, s>>>(d);
  B2H_LAUNCH_CHECK("bn_fold");
  return B2H_OK;
}

}  // namespace b2h
