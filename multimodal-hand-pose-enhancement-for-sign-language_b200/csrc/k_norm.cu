// BatchNorm1d statistics / apply / backward and column sums (HBM/L2-bound elementwise + reduction
// kernels).  Layout: rows x ld, channels contiguous, 4 channels per thread (8 B bf16 / 16 B fp32
// vector access), CTA = (TXp, TY) threads covering kBnChunkRows rows.
//
// Reference semantics restated (modelZoo.py:192-198, SURVEY.md section 8a "PyTorch semantics"):
//   train: y = (z - mean_b) / sqrt(var_b(biased) + eps) * gamma + beta;
//          running = (1-m)*running + m*batch (running_var from the UNBIASED batch variance);
//   eval:  y = (z - running_mean) / sqrt(running_var + eps) * gamma + beta.
#include "b2h_common.cuh"

namespace b2h {

struct BlockShape {
  int txp, ty;
};
static BlockShape block_shape(int Cwork) {
  int tx = ceil_div(Cwork, 4);
  int txp = 1;
  while (txp < tx) txp <<= 1;
  if (txp > 256) txp = 256;
  return {txp, 256 / txp};
}

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
// bn_stats: chunked Welford + ordered Chan merge (deterministic), finalised by the last CTA
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(b2h_bn_stats_t d, int nchunks) {
  __shared__ float4 s_mean[256];
  __shared__ float4 s_m2[256];
  __shared__ int s_n[256];
  const int tx = threadIdx.x, ty = threadIdx.y, TXp = blockDim.x, TY = blockDim.y;
  const int chunk = blockIdx.x, g = blockIdx.y;
  const int c0 = tx * 4;
  const int rpg = d.rows_per_group;
  const int r_begin = chunk * kBnChunkRows;
  const int r_end = min(r_begin + kBnChunkRows, rpg);
  const T* z = reinterpret_cast<const T*>(d.z);
  float4 mean = make_float4(0, 0, 0, 0), m2 = make_float4(0, 0, 0, 0);
  int n = 0;
  if (c0 < d.C) {
    for (int r = r_begin + ty; r < r_end; r += TY) {
      float4 v = load4<T>(z + ((int64_t)g * rpg + r) * d.ld + c0);
      ++n;
      float inv = 1.0f / (float)n;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float delta = f4(v, i) - f4(mean, i);
        f4(mean, i) += delta * inv;
        f4(m2, i) += delta * (f4(v, i) - f4(mean, i));
      }
    }
  }
  const int sidx = ty * TXp + tx;
  s_mean[sidx] = mean;
  s_m2[sidx] = m2;
  s_n[sidx] = n;
  __syncthreads();
  if (ty == 0 && c0 < d.C) {
    float na = (float)n;
    for (int j = 1; j < TY; ++j) {
      int nbj = s_n[j * TXp + tx];
      if (nbj == 0) continue;
      float nb = (float)nbj;
      float4 mb = s_mean[j * TXp + tx], qb = s_m2[j * TXp + tx];
      float nab = na + nb;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float delta = f4(mb, i) - f4(mean, i);
        f4(mean, i) += delta * (nb / nab);
        f4(m2, i) += f4(qb, i) + delta * delta * (na * nb / nab);
      }
      na = nab;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = c0 + i;
      if (c < d.C) {
        float* p = d.partial + (((int64_t)chunk * d.groups + g) * d.C + c) * 2;
        p[0] = f4(mean, i);
        p[1] = f4(m2, i);
      }
    }
  }
  if (!last_block_done(d.ticket, gridDim.x * gridDim.y)) return;
  // ordered merge over chunks, in double
  const int tid = ty * TXp + tx;
  for (int c = tid; c < d.C; c += 256) {
    for (int gg = 0; gg < d.groups; ++gg) {
      double na = 0.0, ma = 0.0, qa = 0.0;
      for (int ch = 0; ch < nchunks; ++ch) {
        int rb = ch * kBnChunkRows;
        double nb = (double)(min(rb + kBnChunkRows, rpg) - rb);
        const float* p = d.partial + (((int64_t)ch * d.groups + gg) * d.C + c) * 2;
        double mb = (double)__ldcg(p), qb = (double)__ldcg(p + 1);
        double nab = na + nb, delta = mb - ma;
        ma += delta * (nb / nab);
        qa += qb + delta * delta * (na * nb / nab);
        na = nab;
      }
      double var_b = qa / na;
      d.mean[gg * d.C + c] = (float)ma;
      d.invstd[gg * d.C + c] = (float)(1.0 / sqrt(var_b + (double)d.eps));
      if (d.running_mean && (gg == 0 || d.update_all_groups)) {
        double var_u = na > 1.0 ? qa / (na - 1.0) : var_b;
        float mom = d.momentum;
        d.running_mean[c] = (1.f - mom) * d.running_mean[c] + mom * (float)ma;
        d.running_var[c] = (1.f - mom) * d.running_var[c] + mom * (float)var_u;
      }
    }
  }
  if (tid == 0 && d.running_mean && d.num_batches_tracked)
    *d.num_batches_tracked += d.update_all_groups ? d.groups : 1;
}

int64_t bn_partial_floats(int rows, int C, int groups) {
  int rpg = rows / (groups > 0 ? groups : 1);
  return (int64_t)bn_nchunks(rpg) * groups * C * 2;
}

int launch_bn_stats(const b2h_bn_stats_t& d, int dtype, cudaStream_t s) {
  B2H_CHECK_ARG(d.C > 0 && d.C <= 1024 && d.groups >= 1 && d.rows_per_group > 0, B2H_ERR_SHAPE,
                "bn_stats: bad shape C=%d groups=%d rows=%d", d.C, d.groups, d.rows_per_group);
  B2H_CHECK_ARG(d.ld % 4 == 0 && d.ld >= ((d.C + 3) & ~3), B2H_ERR_ALIGN, "bn_stats: ld=%d C=%d", d.ld, d.C);
  BlockShape bs = block_shape(d.C);
  int nchunks = bn_nchunks(d.rows_per_group);
  dim3 grid(nchunks, d.groups), block(bs.txp, bs.ty);
  if (dtype == B2H_BF16)
    bn_stats_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(d, nchunks);
  else
    bn_stats_kernel<float><<<grid, block, 0, s>>>(d, nchunks);
  B2H_LAUNCH_CHECK("bn_stats");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// bn_apply: out = dropout( BN0(src0) [+ BN1(src1)] ), zero fill of the channel padding
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(b2h_bn_apply_t d) {
  const int tx = threadIdx.x, ty = threadIdx.y, TY = blockDim.y;
  const int c0 = tx * 4;
  if (c0 >= d.Cfill) return;
  const int rows = d.B * d.L;
  const int rpg = rows / d.groups;
  const int r_begin = blockIdx.x * kBnChunkRows;
  const int r_end = min(r_begin + kBnChunkRows, rows);
  DropCtx drop;
  drop.init(d.drop);
  T* out = reinterpret_cast<T*>(d.out);
  Affine4 a0, a1;
  int gcur = -1;
  for (int row = r_begin + ty; row < r_end; row += TY) {
    const int b = row / d.L, l = row - b * d.L;
    const int g = row / rpg;
    float4 y = make_float4(0, 0, 0, 0);
    if (c0 < d.C) {
      if (g != gcur) {
        a0 = bn_affine(d.src[0], g, d.C, c0);
        if (d.nsrc > 1) a1 = bn_affine(d.src[1], g, d.C, c0);
        gcur = g;
      }
      y = bn_src_eval<T>(d.src[0], a0, b, l, c0);
      if (d.nsrc > 1) y = add4(y, bn_src_eval<T>(d.src[1], a1, b, l, c0));
      if (drop.mode != B2H_DROP_NONE) {
        float4 m = drop.scale4((uint64_t)row * d.drop_C + d.drop_coff + c0);
        y.x *= m.x, y.y *= m.y, y.z *= m.z, y.w *= m.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c0 + i >= d.C) f4(y, i) = 0.f;
    }
    store4<T>(out + (int64_t)row * d.out_ld + d.out_coff + c0, y);
  }
}

int launch_bn_apply(const b2h_bn_apply_t& d, int dtype, cudaStream_t s) {
  B2H_CHECK_ARG(d.nsrc >= 1 && d.nsrc <= 2 && d.C > 0 && d.Cfill >= d.C && d.Cfill <= 1024 && d.groups >= 1,
                B2H_ERR_SHAPE, "bn_apply: bad shape C=%d Cfill=%d nsrc=%d", d.C, d.Cfill, d.nsrc);
  B2H_CHECK_ARG(d.Cfill % 4 == 0 && d.out_ld % 4 == 0 && d.out_coff % 4 == 0, B2H_ERR_ALIGN,
                "bn_apply: alignment Cfill=%d ld=%d coff=%d", d.Cfill, d.out_ld, d.out_coff);
  B2H_CHECK_ARG((d.B * d.L) % d.groups == 0, B2H_ERR_SHAPE, "bn_apply: rows not divisible by groups");
  for (int i = 0; i < d.nsrc; ++i) {
    B2H_CHECK_ARG(d.src[i].ld % 4 == 0 && d.src[i].coff % 4 == 0, B2H_ERR_ALIGN, "bn_apply: src alignment");
    B2H_CHECK_ARG(d.src[i].coff + d.C <= d.src[i].C_total, B2H_ERR_SHAPE, "bn_apply: source channel range");
  }
  BlockShape bs = block_shape(d.Cfill);
  dim3 grid(ceil_div(d.B * d.L, kBnChunkRows)), block(bs.txp, bs.ty);
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
__global__ void __launch_bounds__(256) bn_bwd_kernel(b2h_bn_bwd_t d, int nchunks) {
  __shared__ float4 s_a[256];
  __shared__ float4 s_b[256];
  const int tx = threadIdx.x, ty = threadIdx.y, TXp = blockDim.x, TY = blockDim.y;
  const int chunk = blockIdx.x, g = blockIdx.y;
  const int c0 = tx * 4;
  const int rows = d.B * d.L;
  const int rpg = rows / d.groups;
  const int r_begin = chunk * kBnChunkRows;
  const int r_end = min(r_begin + kBnChunkRows, rpg);
  const T* z = reinterpret_cast<const T*>(d.bn.z);
  T* dpre = reinterpret_cast<T*>(d.dpre);
  float4 acc_a = make_float4(0, 0, 0, 0), acc_b = make_float4(0, 0, 0, 0);
  if (c0 < d.Cfill) {
    Affine4 aff;
    float4 mean4 = make_float4(0, 0, 0, 0), istd4 = mean4, sg4 = mean4, mdy = mean4, mdyz = mean4;
    if (c0 < d.C) {
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
    for (int r = r_begin + ty; r < r_end; r += TY) {
      const int row = g * rpg + r;
      const int b = row / d.L, l = row - b * d.L;
      float4 out = make_float4(0, 0, 0, 0);
      if (c0 < d.C) {
        float4 zo = load4<T>(z + (int64_t)row * d.bn.ld + d.bn.coff + c0);
        float4 dy = bn_bwd_dy<T>(d, aff, zo, b, l, c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c0 + i < d.C) {
            float zh = (f4(zo, i) - f4(mean4, i)) * f4(istd4, i);
            if (PASS == 1) {
              f4(acc_a, i) += f4(dy, i);
              f4(acc_b, i) += f4(dy, i) * zh;
            } else {
              float dz = f4(sg4, i) * (f4(dy, i) - f4(mdy, i) - zh * f4(mdyz, i));
              float dp = dz * act_bwd(f4(zo, i), d.act);
              f4(out, i) = dp;
              f4(acc_a, i) += dp;
            }
          }
        }
      }
      if (PASS == 2) store4<T>(dpre + (int64_t)row * d.ld_dpre + c0, out);
    }
  }
  const int sidx = ty * TXp + tx;
  s_a[sidx] = acc_a;
  s_b[sidx] = acc_b;
  __syncthreads();
  if (ty == 0 && c0 < d.C) {
    for (int j = 1; j < TY; ++j) {
      acc_a = add4(acc_a, s_a[j * TXp + tx]);
      acc_b = add4(acc_b, s_b[j * TXp + tx]);
    }
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
  for (int c = tid; c < d.C; c += 256) {
    double tot_a = 0.0, tot_b = 0.0;
    for (int gg = 0; gg < d.groups; ++gg) {
      double sa = 0.0, sb = 0.0;
      for (int ch = 0; ch < nchunks; ++ch) {
        const float* p = d.partial + (((int64_t)ch * d.groups + gg) * d.C + c) * 2;
        sa += (double)__ldcg(p);
        sb += (double)__ldcg(p + 1);
      }
      if (PASS == 1) {
        d.sums[((int64_t)gg * d.C + c) * 2 + 0] = (float)sa;
        d.sums[((int64_t)gg * d.C + c) * 2 + 1] = (float)sb;
      }
      tot_a += sa;
      tot_b += sb;
    }
    if (PASS == 1) {
      if (d.dbeta) d.dbeta[c] = (float)tot_a;
      if (d.dgamma) d.dgamma[c] = (float)tot_b;
    } else {
      if (d.dbias) d.dbias[c] = (float)tot_a;
    }
  }
}

int launch_bn_bwd(const b2h_bn_bwd_t& d, int dtype, cudaStream_t s) {
  B2H_CHECK_ARG(d.C > 0 && d.Cfill >= d.C && d.Cfill <= 1024 && d.groups >= 1 && d.ngsrc >= 1 && d.ngsrc <= 2,
                B2H_ERR_SHAPE, "bn_bwd: bad shape C=%d Cfill=%d ngsrc=%d", d.C, d.Cfill, d.ngsrc);
  B2H_CHECK_ARG(d.Cfill % 4 == 0 && d.ld_dpre % 4 == 0 && d.bn.ld % 4 == 0 && d.bn.coff % 4 == 0, B2H_ERR_ALIGN,
                "bn_bwd: alignment");
  B2H_CHECK_ARG((d.B * d.L) % d.groups == 0, B2H_ERR_SHAPE, "bn_bwd: rows not divisible by groups");
  B2H_CHECK_ARG(!d.bn.use_running, B2H_ERR_ARG, "bn_bwd: backward is only defined for batch statistics");
  B2H_CHECK_ARG(d.bn.coff == 0 && d.bn.C_total == d.C, B2H_ERR_SHAPE, "bn_bwd: bn source must cover the whole layer");
  BlockShape bs = block_shape(d.Cfill);
  int rpg = d.B * d.L / d.groups;
  int nchunks = bn_nchunks(rpg);
  dim3 grid(nchunks, d.groups), block(bs.txp, bs.ty);
  if (dtype == B2H_BF16) {
    bn_bwd_kernel<__nv_bfloat16, 1><<<grid, block, 0, s>>>(d, nchunks);
    B2H_LAUNCH_CHECK("bn_bwd pass 1");
    bn_bwd_kernel<__nv_bfloat16, 2><<<grid, block, 0, s>>>(d, nchunks);
  } else {
    bn_bwd_kernel<float, 1><<<grid, block, 0, s>>>(d, nchunks);
    B2H_LAUNCH_CHECK("bn_bwd pass 1");
    bn_bwd_kernel<float, 2><<<grid, block, 0, s>>>(d, nchunks);
  }
  B2H_LAUNCH_CHECK("bn_bwd pass 2");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// colsum: out[c] = sum over rows of src[row][c] (bias gradients of layers without BN)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(b2h_colsum_t d, int nchunks) {
  __shared__ float4 s_a[256];
  const int tx = threadIdx.x, ty = threadIdx.y, TXp = blockDim.x, TY = blockDim.y;
  const int c0 = tx * 4;
  const int r_begin = blockIdx.x * kBnChunkRows;
  const int r_end = min(r_begin + kBnChunkRows, d.rows);
  float4 acc = make_float4(0, 0, 0, 0);
  if (c0 < d.C)
    for (int r = r_begin + ty; r < r_end; r += TY)
      acc = add4(acc, load4<T>(reinterpret_cast<const T*>(d.src) + (int64_t)r * d.ld + c0));
  s_a[ty * TXp + tx] = acc;
  __syncthreads();
  if (ty == 0 && c0 < d.C) {
    for (int j = 1; j < TY; ++j) acc = add4(acc, s_a[j * TXp + tx]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (c0 + i < d.C) d.partial[(int64_t)blockIdx.x * d.C + c0 + i] = f4(acc, i);
  }
  if (!last_block_done(d.ticket, gridDim.x)) return;
  for (int c = ty * TXp + tx; c < d.C; c += 256) {
    double sa = 0.0;
    for (int ch = 0; ch < nchunks; ++ch) sa += (double)__ldcg(d.partial + (int64_t)ch * d.C + c);
    d.out[c] = (float)sa;
  }
}

int launch_colsum(const b2h_colsum_t& d, int dtype, cudaStream_t s) {
  B2H_CHECK_ARG(d.C > 0 && d.C <= 1024 && d.rows > 0 && d.ld % 4 == 0, B2H_ERR_SHAPE, "colsum: bad shape");
  BlockShape bs = block_shape(d.C);
  int nchunks = ceil_div(d.rows, kBnChunkRows);
  dim3 grid(nchunks), block(bs.txp, bs.ty);
  if (dtype == B2H_BF16 && !d.f32)
    colsum_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(d, nchunks);
  else
    colsum_kernel<float><<<grid, block, 0, s>>>(d, nchunks);
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
  B2H_CHECK_ARG(d.C > 0 && d.Cpad >= d.C, B2H_ERR_SHAPE, "bn_fold: bad shape");
  bn_fold_kernel<<<ceil_div(d.Cpad, 128), 128, 0, s>>>(d);
  B2H_LAUNCH_CHECK("bn_fold");
  return B2H_OK;
}

}  // namespace b2h
