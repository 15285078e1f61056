// Boundary layout changes (NCL fp32 <-> BLC act dtype), calc_motion, losses, fused Adam, weight
// repacking and 6D->rotation-matrix.  All HBM/L2-bound: coalesced on both sides via 32x32 smem
// tile transposes, float4 access where the layout allows, warp-shuffle + ordered final reductions.
#include "b2h_common.cuh"

namespace b2h {

// ---------------------------------------------------------------------------------------------
// prep: fp32 source (NCL / rows / broadcast rows / calc_motion of NCL) -> dropout -> BLC act dtype
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) prep_ncl_kernel(b2h_prep_t d) {
  pdl_sync();
  // tile: 32 channels x 32 time steps of sample blockIdx.z
  __shared__ float tile[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const bool motion = d.kind == B2H_SRC_MOTION;
  const int Ls = motion ? d.L + 1 : d.L;  // source time length
  const float* src = d.src + (int64_t)b * d.C * Ls;
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j, l = l0 + tx;
    float v = 0.f;
    if (c < d.C && l < d.L) {
      v = src[(int64_t)c * Ls + l];
      if (motion) v = src[(int64_t)c * Ls] - v;  // x[:, :, :1] - x[:, :, :-1]  (train_gan.py:210)
    }
    tile[j][tx] = v;
  }
  __syncthreads();
  DropCtx drop;
  drop.init(d.drop, d.C);
  // write phase: each thread stores 4 consecutive channels of one time step (8 B bf16 / 16 B fp32)
  const int tid = ty * 32 + tx;
  const int l = l0 + (tid >> 3), c = c0 + (tid & 7) * 4;
  if (l < d.L && c < d.Cfill) {
    const int64_t row = (int64_t)b * d.L + l;
    float4 v = make_float4(0, 0, 0, 0);
    if (c < d.C) {
      v = make_float4(tile[(tid & 7) * 4 + 0][tid >> 3], tile[(tid & 7) * 4 + 1][tid >> 3],
                      tile[(tid & 7) * 4 + 2][tid >> 3], tile[(tid & 7) * 4 + 3][tid >> 3]);
      if (drop.mode != B2H_DROP_NONE) {
        const uint64_t idx = (uint64_t)row * d.C + c;
        if (c + 3 < d.C) {
          float4 m = drop.scale4(idx);
          v.x *= m.x, v.y *= m.y, v.z *= m.z, v.w *= m.w;
          drop.save4(idx, m, 4);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c + k < d.C) {
              float mk = drop.scale1(idx + k);
              f4(v, k) *= mk;
              if (drop.save) drop.save[idx + k] = mk != 0.f;
            }
        }
      }
    }
    const int64_t o = row * d.ld + c;
    if (d.out_f32)
      store4<float>(reinterpret_cast<float*>(d.out) + o, v);
    else
      store4<T>(reinterpret_cast<T*>(d.out) + o, v);
  }
}

// The same op for sources whose rows are 16-byte aligned (source length % 4 == 0; Cfill, ld % 8 == 0): one CTA = 64
// source frames x 32 channels of one clip; every thread requests its two float4 of the source (channel t/8, 4
// consecutive frames, + 32 frames) before anything else, the tile is transposed through shared memory (pitch 33:
// conflict-free both ways), and a thread writes 8 channels of one frame (16 B in bf16) with ONE Philox call for its
// 8 keep flags (the scalar kernel spends one call per 4).  Same values, same flags.
template <typename T>
__global__ void __launch_bounds__(256) prep_ncl_vec_kernel(b2h_prep_t d) {
  pdl_sync();
  __shared__ float tile[64][33];
  const int t = threadIdx.x;
  const int b = blockIdx.z, l0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
  const bool motion = d.kind == B2H_SRC_MOTION;
  const int Ls = motion ? d.L + 1 : d.L;
  const int ch = c0 + (t >> 3), f4 = (t & 7) * 4;
  const float* src = d.src + ((int64_t)b * d.C + ch) * Ls;
  float4 x[2];
  float first = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int l = l0 + f4 + 32 * h;
    x[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ch < d.C && l < Ls) x[h] = *reinterpret_cast<const float4*>(src + l);
  }
  if (motion && ch < d.C) first = src[0];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int f = f4 + 32 * h;
    // x[:, :, :1] - x[:, :, :-1]  (train_gan.py:210); frames past the end hold zeros and are never written out
    tile[f + 0][t >> 3] = motion ? first - x[h].x : x[h].x;
    tile[f + 1][t >> 3] = motion ? first - x[h].y : x[h].y;
    tile[f + 2][t >> 3] = motion ? first - x[h].z : x[h].z;
    tile[f + 3][t >> 3] = motion ? first - x[h].w : x[h].w;
  }
  __syncthreads();
  const int r = t >> 2, c8 = (t & 3) * 8;
  const int l = l0 + r, c = c0 + c8;
  if (l >= d.L || c >= d.Cfill) return;
  const int64_t row = (int64_t)b * d.L + l;
  F8 v;
#pragma unroll
  for (int k = 0; k < 8; ++k) v.v[k] = (c + k < d.C) ? tile[r][c8 + k] : 0.f;
  if (c < d.C) {
    DropCtx drop;
    drop.init(d.drop, d.C);
    if (drop.mode != B2H_DROP_NONE) {
      const uint32_t bits = drop.keep8_rc((uint32_t)row, (uint32_t)c);
#pragma unroll
      for (int k = 0; k < 8; ++k) v.v[k] *= ((bits >> k) & 1u) ? 2.f : 0.f;
      drop.save8((uint64_t)row * d.C + c, bits, min(8, d.C - c));
    }
  }
  const int64_t o = row * d.ld + c;
  if (d.out_f32)
    store8<float>(reinterpret_cast<float*>(d.out) + o, v);
  else
    store8<T>(reinterpret_cast<T*>(d.out) + o, v);
}

template <typename T>
__global__ void __launch_bounds__(256) prep_rows_kernel(b2h_prep_t d) {
  pdl_sync();
  // one thread per 4 channels; rows = B*L
  const int64_t rows = (int64_t)d.B * d.L;
  const int nq = d.Cfill / 4;
  DropCtx drop;
  drop.init(d.drop, d.C);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * nq;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / nq;
    int c0 = (int)(i - row * nq) * 4;
    int64_t srow = d.kind == B2H_SRC_BCAST ? row / d.L : row;
    const float* sp = d.src + srow * d.src_ld + c0;
    float4 v = make_float4(0, 0, 0, 0);
    if (c0 + 3 < d.C && (d.src_ld % 4 == 0)) {
      v = *reinterpret_cast<const float4*>(sp);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c0 + k < d.C) f4(v, k) = sp[k];
    }
    if (drop.mode != B2H_DROP_NONE && c0 < d.C) {
      uint64_t idx = (uint64_t)row * d.C + c0;
      if (c0 + 3 < d.C) {
        float4 m = drop.scale4(idx);
        v.x *= m.x, v.y *= m.y, v.z *= m.z, v.w *= m.w;
        drop.save4(idx, m, 4);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c0 + k < d.C) {
            float mk = drop.scale1(idx + k);
            f4(v, k) *= mk;
            if (drop.save) drop.save[idx + k] = mk != 0.f;
          }
      }
    }
    if (d.out_f32)
      store4<float>(reinterpret_cast<float*>(d.out) + row * d.ld + c0, v);
    else
      store4<T>(reinterpret_cast<T*>(d.out) + row * d.ld + c0, v);
  }
}

int launch_prep(const b2h_prep_t& d, int dtype, cudaStream_t s) {
  B2H_CARVE(prep_ncl_kernel<__nv_bfloat16>);
  B2H_CARVE(prep_ncl_kernel<float>);
  B2H_CARVE(prep_rows_kernel<__nv_bfloat16>);
  B2H_CARVE(prep_rows_kernel<float>);
  B2H_CHECK_ARG(d.B > 0 && d.L > 0 && d.C > 0 && d.Cfill >= d.C && d.ld >= d.Cfill, B2H_ERR_SHAPE,
                "prep: bad shape B=%d L=%d C=%d Cfill=%d ld=%d", d.B, d.L, d.C, d.Cfill, d.ld);
  if (d.kind == B2H_SRC_NCL || d.kind == B2H_SRC_MOTION) {
    B2H_CHECK_ARG(d.B <= 65535, B2H_ERR_SHAPE, "prep: B too large for grid.z");
    B2H_CHECK_ARG(d.Cfill % 4 == 0 && d.ld % 4 == 0, B2H_ERR_ALIGN, "prep: Cfill/ld must be multiples of 4");
    static const bool no_vec = getenv("B2H_NO_PREP_VEC") != nullptr;
    const int Ls = d.kind == B2H_SRC_MOTION ? d.L + 1 : d.L;
    const int esz_out = (d.out_f32 || dtype != B2H_BF16) ? 4 : 2;
    if (!no_vec && Ls % 4 == 0 && d.Cfill % 8 == 0 && d.ld % 8 == 0 && (uintptr_t)d.src % 16 == 0 &&
        (uintptr_t)d.out % 16 == 0 && esz_out * d.ld % 16 == 0) {
      B2H_CARVE(prep_ncl_vec_kernel<__nv_bfloat16>);
      B2H_CARVE(prep_ncl_vec_kernel<float>);
      dim3 vgrid(ceil_div(d.L, 64), ceil_div(d.Cfill, 32), d.B);
      if (dtype == B2H_BF16)
        launch(prep_ncl_vec_kernel<__nv_bfloat16>, vgrid, 256, 0, s, d);
      else
        launch(prep_ncl_vec_kernel<float>, vgrid, 256, 0, s, d);
      B2H_LAUNCH_CHECK("prep");
      return B2H_OK;
    }
    dim3 grid(ceil_div(d.L, 32), ceil_div(d.Cfill, 32), d.B), block(32, 8);
    if (dtype == B2H_BF16)
      launch(prep_ncl_kernel<__nv_bfloat16>, grid, block, 0, s, d);
    else
      launch(prep_ncl_kernel<float>, grid, block, 0, s, d);
  } else {
    B2H_CHECK_ARG(d.Cfill % 4 == 0 && d.ld % 4 == 0, B2H_ERR_ALIGN, "prep: Cfill/ld must be multiples of 4");
    int64_t work = (int64_t)d.B * d.L * (d.Cfill / 4);
    int blocks = (int)std::min<int64_t>(ceil_div64(work, 256), (int64_t)sm_count() * 16);
    if (dtype == B2H_BF16)
      launch(prep_rows_kernel<__nv_bfloat16>, blocks, 256, 0, s, d);
    else
      launch(prep_rows_kernel<float>, blocks, 256, 0, s, d);
  }
  B2H_LAUNCH_CHECK("prep");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// to_ncl: BLC -> (B, C, L) fp32
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) to_ncl_kernel(b2h_to_ncl_t d) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int b = blockIdx.z, l0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const T* src = reinterpret_cast<const T*>(d.src) + (int64_t)b * d.L * d.ld;
  for (int j = ty; j < 32; j += 8) {
    int l = l0 + j, c = c0 + tx;
    tile[j][tx] = (l < d.L && c < d.C) ? to_f<T>(src[(int64_t)l * d.ld + c]) : 0.f;
  }
  __syncthreads();
  float* dst = d.dst + (int64_t)b * d.C * d.L;
  for (int j = ty; j < 32; j += 8) {
    int c = c0 + j, l = l0 + tx;
    if (c < d.C && l < d.L) dst[(int64_t)c * d.L + l] = tile[tx][j];
  }
}

int launch_to_ncl(const b2h_to_ncl_t& d, int dtype, cudaStream_t s) {
  B2H_CARVE(to_ncl_kernel<__nv_bfloat16>);
  B2H_CARVE(to_ncl_kernel<float>);
  B2H_CHECK_ARG(d.B > 0 && d.B <= 65535 && d.L > 0 && d.C > 0 && d.ld >= d.C, B2H_ERR_SHAPE, "to_ncl: bad shape");
  dim3 grid(ceil_div(d.L, 32), ceil_div(d.C, 32), d.B), block(32, 8);
  if (dtype == B2H_BF16 && !d.src_f32)
    launch(to_ncl_kernel<__nv_bfloat16>, grid, block, 0, s, d);
  else
    launch(to_ncl_kernel<float>, grid, block, 0, s, d);
  B2H_LAUNCH_CHECK("to_ncl");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// Regression criterion forward + backward (LOSSES, utils/constants.py:53-58; default nn.L1Loss): one pass over out, gt
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid (L/32, B/spc, channel tiles): one 32 x 32 (frames x channels) tile per CTA and clip -- the kernel is bound by the
// latency of a CTA's dependent phases (load, transpose, store), not by bytes, so the tiles of a clip run side by side
// instead of one after the other (8 tiles per CTA: 25.6 us; one: see profiles/)
template <typename T>
__global__ void __launch_bounds__(256) l1_kernel(b2h_l1_t d, int cext, int spc) {
  pdl_sync();
  __shared__ float tile[32][33];
  __shared__ float tile_o[32][33];   // out_blc: the prediction tile, [frame][channel]
  __shared__ double s_part[256];
  __shared__ float s_col[1024];
  const int l0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * 32 + tx;
  if (d.dbias) {
    for (int c = tid; c < d.C; c += 256) s_col[c] = 0.f;
    __syncthreads();
  }
  const int64_t numel = (int64_t)d.B * d.C * d.L;
  const float gval = d.gscale / (float)numel;
  const uint32_t nblocks = gridDim.x * gridDim.y * gridDim.z;
  const uint32_t bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  float acc = 0.f;
  for (int si = 0; si < spc; ++si) {
  const int b = blockIdx.y * spc + si;
  for (int c0 = blockIdx.z * 32; c0 < min(cext, (int)blockIdx.z * 32 + 32); c0 += 32) {
    if (d.out_blc) {
      // the prediction comes as BLC rows (channels contiguous): coalesced 128-byte row reads, transposed through
      // shared memory; the NCL copy the reference returns is written below, in the same pass
      for (int j = ty; j < 32; j += 8) {
        const int l = l0 + j, c = c0 + tx;
        tile_o[j][tx] = (l < d.L && c < d.C) ? d.out_blc[((int64_t)b * d.L + l) * d.out_blc_ld + c] : 0.f;
      }
      __syncthreads();
    }
    for (int j = ty; j < 32; j += 8) {
      int c = c0 + j, l = l0 + tx;
      float sg = 0.f;
      if (c < d.C && l < d.L) {
        int64_t i = ((int64_t)b * d.C + c) * d.L + l;
        float o;
        if (d.out_blc) {
          o = tile_o[tx][j];
          const_cast<float*>(d.out)[i] = o;
        } else {
          o = d.out[i];
        }
        float diff = o - d.gt[i];
        if (d.kind == B2H_LOSS_L1) {
          acc += fabsf(diff);
          sg = diff > 0.f ? gval : (diff < 0.f ? -gval : 0.f);  // sign(diff)/numel, sign(0) = 0
        } else if (d.kind == B2H_LOSS_HUBER1) {                 // nn.HuberLoss(delta = 1)
          const float a = fabsf(diff);
          acc += a < 1.f ? 0.5f * diff * diff : a - 0.5f;
          sg = fminf(fmaxf(diff, -1.f), 1.f) * gval;
        } else {                                                // L2: d^2 ; ROBUST (alpha 2, scale 1/2): 2 d^2
          const float w = d.kind == B2H_LOSS_ROBUST ? 2.f : 1.f;
          acc += w * diff * diff;
          sg = 2.f * w * diff * gval;
        }
      }
      tile[j][tx] = sg;
      if (d.dbias) {   // column sum over this tile's 32 time steps, of the value as it is stored
        const float cs = warp_sum(to_f<T>(from_f<T>(sg)));
        if (tx == 0 && c < d.C) s_col[c] += cs;
      }
    }
    __syncthreads();
    if (d.dout) {
      // 4 consecutive channels of one time step per thread
      const int l = l0 + (tid >> 3), c = c0 + (tid & 7) * 4;
      if (l < d.L && c < d.Cfill) {
        float4 v = make_float4(tile[(tid & 7) * 4 + 0][tid >> 3], tile[(tid & 7) * 4 + 1][tid >> 3],
                               tile[(tid & 7) * 4 + 2][tid >> 3], tile[(tid & 7) * 4 + 3][tid >> 3]);
        store4<T>(reinterpret_cast<T*>(d.dout) + ((int64_t)b * d.L + l) * d.ld + c, v);
      }
    }
    __syncthreads();
  }
  }
  if (d.dbias) {
    for (int c = blockIdx.z * 32 + tid; c < min(d.C, (int)blockIdx.z * 32 + 32); c += 256)   // this CTA's channels
      atomicAdd(d.dbias_accum + (int64_t)(bid % 16) * d.C + c, (double)s_col[c]);
  }
  s_part[tid] = (double)acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_part[tid] += s_part[tid + o];
    __syncthreads();
  }
  if (tid == 0) d.partial[bid] = (float)s_part[0];
  if (!last_block_done(d.ticket, nblocks)) return;
  if (d.dbias) {
    for (int c = tid; c < d.C; c += 256) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        double* a = d.dbias_accum + (int64_t)k * d.C + c;
        t += __ldcg(a);
        *a = 0.0;
      }
      d.dbias[c] = (float)t;
    }
  }
  double t = 0.0;
  for (uint32_t k = tid; k < nblocks; k += 256) t += (double)__ldcg(d.partial + k);
  __syncthreads();
  s_part[tid] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_part[tid] += s_part[tid + o];
    __syncthreads();
  }
  // ROBUST: the constant of the negative log-likelihood, log(scale) + log Z(alpha) = log(1/2) + log sqrt(2 pi)
  if (tid == 0) d.loss[0] = (float)(s_part[0] / (double)numel + (d.kind == B2H_LOSS_ROBUST ? 0.22579135264472743 : 0.0));
}

// one element of the criterion: adds its loss term to acc, returns d loss / d out (see b2h_l1_t)
__device__ __forceinline__ float loss_elem(int kind, float diff, float gval, float& acc) {
  if (kind == B2H_LOSS_L1) {
    acc += fabsf(diff);
    return diff > 0.f ? gval : (diff < 0.f ? -gval : 0.f);  // sign(diff)/numel, sign(0) = 0
  }
  if (kind == B2H_LOSS_HUBER1) {                            // nn.HuberLoss(delta = 1)
    const float a = fabsf(diff);
    acc += a < 1.f ? 0.5f * diff * diff : a - 0.5f;
    return fminf(fmaxf(diff, -1.f), 1.f) * gval;
  }
  const float w = kind == B2H_LOSS_ROBUST ? 2.f : 1.f;      // L2: d^2 ; ROBUST (alpha 2, scale 1/2): 2 d^2
  acc += w * diff * diff;
  return 2.f * w * diff * gval;
}

// The generator step's form of the criterion (prediction given as the output layer's fp32 BLC rows, L % 4 == 0):
// one CTA = 64 frames x 32 channels of one clip, every global access 16 bytes wide and all four loads of a thread
// (2 x prediction rows, 2 x target rows = 64 B) issued before the first use.
//   load   thread (row = t/8 [+32], 4 channels)  : out_blc rows, 128 B per row and warp quarter
//          thread (channel = t/8, 4 frames [+32]): gt (NCL), 128 B per channel and warp quarter
//   pass 1 prediction tile -> shared [frame][channel] (pitch 33: conflict-free both ways)
//   pass 2 in NCL orientation: out (NCL, the tensor the reference returns) written as float4, loss terms, gradient
//          -> shared [frame][channel]; column sums of the gradient as stored (bias gradient of the output layer):
//          8 lanes share a channel -> shuffle -> one fp64 atomic per channel and CTA
//   pass 3 gradient rows out: thread = 8 channels of one frame (16 B in bf16)
template <typename T>
__global__ void __launch_bounds__(256) l1_vec_kernel(b2h_l1_t d, int cext) {
  pdl_sync();
  __shared__ float tile_o[64][33];
  __shared__ float tile_g[64][33];
  __shared__ double s_part[256];
  const int t = threadIdx.x;
  const int l0 = blockIdx.x * 64, b = blockIdx.y, c0 = blockIdx.z * 32;
  const int64_t numel = (int64_t)d.B * d.C * d.L;
  const float gval = d.gscale / (float)numel;
  const uint32_t nblocks = gridDim.x * gridDim.y * gridDim.z;
  const uint32_t bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  // ---- loads
  const int orow = t >> 3, oc = c0 + (t & 7) * 4;          // prediction: rows orow, orow + 32; channels oc..oc+3
  const int gch = c0 + (t >> 3), gf = (t & 7) * 4;          // target: channel gch; frames l0 + gf + 32 h
  float4 o4[2], g4[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int l = l0 + orow + 32 * h;
    o4[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < d.L && oc < d.out_blc_ld)
      o4[h] = *reinterpret_cast<const float4*>(d.out_blc + ((int64_t)b * d.L + l) * d.out_blc_ld + oc);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int l = l0 + gf + 32 * h;
    g4[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gch < d.C && l < d.L) g4[h] = *reinterpret_cast<const float4*>(d.gt + ((int64_t)b * d.C + gch) * d.L + l);
  }
  // ---- pass 1
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float* row = tile_o[orow + 32 * h] + (t & 7) * 4;
    row[0] = o4[h].x, row[1] = o4[h].y, row[2] = o4[h].z, row[3] = o4[h].w;
  }
  __syncthreads();
  // ---- pass 2
  float acc = 0.f, csum = 0.f;
  const int cl = t >> 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int f = gf + 32 * h, l = l0 + f;
    const bool in = gch < d.C && l < d.L;
    const float o[4] = {tile_o[f][cl], tile_o[f + 1][cl], tile_o[f + 2][cl], tile_o[f + 3][cl]};
    const float g[4] = {g4[h].x, g4[h].y, g4[h].z, g4[h].w};
    float sg[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sg[k] = in ? loss_elem(d.kind, o[k] - g[k], gval, acc) : 0.f;
      csum += to_f<T>(from_f<T>(sg[k]));
      tile_g[f + k][cl] = sg[k];
    }
    if (in)
      *reinterpret_cast<float4*>(const_cast<float*>(d.out) + ((int64_t)b * d.C + gch) * d.L + l) =
          make_float4(o[0], o[1], o[2], o[3]);
  }
  if (d.dbias) {
    csum += __shfl_xor_sync(0xffffffffu, csum, 1);
    csum += __shfl_xor_sync(0xffffffffu, csum, 2);
    csum += __shfl_xor_sync(0xffffffffu, csum, 4);
    if ((t & 7) == 0 && gch < d.C) atomicAdd(d.dbias_accum + (int64_t)(bid % 16) * d.C + gch, (double)csum);
  }
  __syncthreads();
  // ---- pass 3
  if (d.dout) {
    const int r = t >> 2, c8 = (t & 3) * 8, l = l0 + r;
    if (l < d.L && c0 + c8 < cext) {
      F8 v;
#pragma unroll
      for (int k = 0; k < 8; ++k) v.v[k] = tile_g[r][c8 + k];
      store8<T>(reinterpret_cast<T*>(d.dout) + ((int64_t)b * d.L + l) * d.ld + c0 + c8, v);
    }
  }
  // ---- loss: CTA sum -> partial -> the last CTA sums the partials in fixed order (and finishes the bias gradient)
  s_part[t] = (double)acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) s_part[t] += s_part[t + o];
    __syncthreads();
  }
  if (t == 0) d.partial[bid] = (float)s_part[0];
  if (!last_block_done(d.ticket, nblocks)) return;
  if (d.dbias) {
    for (int c = t; c < d.C; c += 256) {
      double tt = 0.0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        double* a = d.dbias_accum + (int64_t)k * d.C + c;
        tt += __ldcg(a);
        *a = 0.0;
      }
      d.dbias[c] = (float)tt;
    }
  }
  double tt = 0.0;
  for (uint32_t k = t; k < nblocks; k += 256) tt += (double)__ldcg(d.partial + k);
  __syncthreads();
  s_part[t] = tt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (t < o) s_part[t] += s_part[t + o];
    __syncthreads();
  }
  if (t == 0) d.loss[0] = (float)(s_part[0] / (double)numel + (d.kind == B2H_LOSS_ROBUST ? 0.22579135264472743 : 0.0));
}

// the vectorised form needs: the prediction as BLC rows, 16-byte aligned rows on both sides, a gradient to write
static bool l1_vec_ok(const b2h_l1_t& d, int cext) {
  static const bool off = getenv("B2H_NO_L1_VEC") != nullptr;
  return !off && d.out_blc && d.dout && d.L % 4 == 0 && d.out_blc_ld % 4 == 0 && cext % 8 == 0 && d.ld % 8 == 0 &&
         ((uintptr_t)d.out_blc % 16 == 0) && ((uintptr_t)d.gt % 16 == 0) && ((uintptr_t)d.out % 16 == 0) &&
         ((uintptr_t)d.dout % 16 == 0);
}

int launch_l1(const b2h_l1_t& d, int dtype, cudaStream_t s) {
  B2H_CARVE(l1_kernel<__nv_bfloat16>);
  B2H_CARVE(l1_kernel<float>);
  B2H_CHECK_ARG(d.B > 0 && d.B <= 65535 && d.C > 0 && d.L > 0, B2H_ERR_SHAPE, "l1: bad shape");
  B2H_CHECK_ARG(d.kind >= B2H_LOSS_L1 && d.kind <= B2H_LOSS_ROBUST, B2H_ERR_ARG, "l1: unknown loss kind");
  B2H_CHECK_ARG(!d.dout || (d.ld >= d.Cfill && d.Cfill >= d.C && d.Cfill % 4 == 0 && d.ld % 4 == 0), B2H_ERR_SHAPE,
                "l1: bad dout shape");
  B2H_CHECK_ARG(!d.dbias || (d.dout && d.dbias_accum), B2H_ERR_ARG, "l1: dbias needs dout and its workspace");
  B2H_CHECK_ARG(!d.dbias || d.C <= 1024, B2H_ERR_SHAPE, "l1: dbias supports up to 1024 channels");
  int cext = d.dout ? d.Cfill : d.C;
  if (l1_vec_ok(d, cext)) {
    B2H_CARVE(l1_vec_kernel<__nv_bfloat16>);
    B2H_CARVE(l1_vec_kernel<float>);
    dim3 grid(ceil_div(d.L, 64), d.B, ceil_div(cext, 32));
    if (dtype == B2H_BF16)
      launch(l1_vec_kernel<__nv_bfloat16>, grid, 256, 0, s, d, cext);
    else
      launch(l1_vec_kernel<float>, grid, 256, 0, s, d, cext);
    B2H_LAUNCH_CHECK("l1");
    return B2H_OK;
  }
  const int spc = 1;   // clips per CTA (more than one was measured slower: the kernel is latency-bound per CTA)
  dim3 grid(ceil_div(d.L, 32), d.B / spc, ceil_div(cext, 32)), block(32, 8);
  if (dtype == B2H_BF16)
    launch(l1_kernel<__nv_bfloat16>, grid, block, 0, s, d, cext, spc);
  else
    launch(l1_kernel<float>, grid, block, 0, s, d, cext, spc);
  B2H_LAUNCH_CHECK("l1");
  return B2H_OK;
}
int64_t l1_partial_floats(const b2h_l1_t& d) {
  return (int64_t)ceil_div(d.L, 32) * d.B * ceil_div(d.dout ? d.Cfill : d.C, 32);
}

// ---------------------------------------------------------------------------------------------
// MSE on discriminator scores (nn.MSELoss, train_gan.py:93): tiny, one CTA
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mse_kernel(b2h_mse_t d) {
  pdl_sync();
  __shared__ double s_red[256];
  double total = 0.0;
  double bsum = 0.0;
  for (int g = 0; g < d.groups; ++g) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < d.n; i += 256) {
      float diff = d.score[((int64_t)g * d.n + i) * d.ld] - d.target[g];
      acc += (double)diff * (double)diff;
      const float gr = 2.0f * diff / (float)d.n;
      if (d.dscore) d.dscore[((int64_t)g * d.n + i) * d.ld] = gr;
      if (d.dpre) {
        const int64_t o = ((int64_t)g * d.n + i) * d.dpre_ld;
        if (d.dpre_bf16) {
          const __nv_bfloat16 q = __float2bfloat16_rn(gr);
          reinterpret_cast<__nv_bfloat16*>(d.dpre)[o] = q;
          bsum += (double)__bfloat162float(q);
        } else {
          reinterpret_cast<float*>(d.dpre)[o] = gr;
          bsum += (double)gr;
        }
      }
    }
    s_red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
      __syncthreads();
    }
    total += s_red[0] / (double)d.n;
    __syncthreads();
  }
  if (d.dbias) {   // bias gradient of the score layer = column sum of the gradient rows just written
    s_red[threadIdx.x] = bsum;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) d.dbias[0] = (float)s_red[0];
  }
  if (threadIdx.x == 0) {
    d.loss[0] = (float)total;
    if (d.total) d.total[0] = (float)total + (d.add ? d.add[0] : 0.f);
  }
}

int launch_mse(const b2h_mse_t& d, cudaStream_t s) {
  B2H_CARVE(mse_kernel);
  B2H_CHECK_ARG(d.groups >= 1 && d.groups <= 2 && d.n > 0 && d.ld >= 1, B2H_ERR_SHAPE, "mse: bad shape");
  B2H_CHECK_ARG((!d.dpre || d.dpre_ld >= 1) && (!d.dbias || d.dpre), B2H_ERR_ARG, "mse: dpre / dbias");
  launch(mse_kernel, 1, 256, 0, s, d);
  B2H_LAUNCH_CHECK("mse");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// fused Adam over a flat parameter buffer (torch.optim.Adam semantics, train_gan.py:69,88)
//   m = m + (g - m)(1 - b1);  v = v*b2 + (1 - b2) g^2
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v)/sqrt(1 - b2^t) + eps)
// 28 B/param: reads p,g,m,v, writes p,m,v.
// ---------------------------------------------------------------------------------------------
// one thread: advance the step counter and evaluate the Python-double scalar prologue of
// torch/optim/adam.py (_single_tensor_adam) once, so the element-wise kernel stays tiny
__global__ void adam_step_kernel(b2h_adam_t d) {
  pdl_sync();
  const int64_t t = *d.step + 1;
  *d.step = t;
  const double bc1 = 1.0 - pow(d.beta1, (double)t);
  const double bc2 = 1.0 - pow(d.beta2, (double)t);
  d.scalars[0] = (float)(-(d.lr / bc1));   // -step_size
  d.scalars[1] = (float)sqrt(bc2);         // bias_correction2_sqrt
}

// one element of torch's _single_tensor_adam, in its operation order
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float gs, float w1, float w2, float b2,
                                          float bc2_sqrt, float eps, float neg_step) {
  const float gk = g * gs;
  const float mk = m + w1 * (gk - m);                  // exp_avg.lerp_(grad, 1 - beta1)
  const float vk = v * b2 + (w2 * gk) * gk;            // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(vk) / bc2_sqrt + eps;      // (sqrt(v) / sqrt(bc2)).add_(eps)
  p += (neg_step * mk) / denom;                        // addcdiv_(exp_avg, denom, value=-step_size)
  m = mk;
  v = vk;
}

// Thread = two float4 of each array per trip (the second one half a grid away), all eight loads issued before the
// first use; the components are named, not indexed (an indexed float4 lives in local memory: the round-1 kernel had a
// 64-byte stack frame and a store / reload per array on its dependency chain).
__global__ void __launch_bounds__(256, 4) adam_kernel(b2h_adam_t d) {
  pdl_sync();
  const float neg_step = d.scalars[0];
  const float bc2_sqrt = d.scalars[1];
  const float w1 = (float)(1.0 - d.beta1), w2 = (float)(1.0 - d.beta2);
  const float b2 = (float)d.beta2, eps = (float)d.eps, gs = d.gscale;
  const int64_t n4 = d.n / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float4* P4 = reinterpret_cast<float4*>(d.p);
  float4* M4 = reinterpret_cast<float4*>(d.m);
  float4* V4 = reinterpret_cast<float4*>(d.v);
  const float4* G4 = reinterpret_cast<const float4*>(d.g);
#define B2H_ADAM4(P, G, M, V)                                                  \
  adam_elem(P.x, G.x, M.x, V.x, gs, w1, w2, b2, bc2_sqrt, eps, neg_step);      \
  adam_elem(P.y, G.y, M.y, V.y, gs, w1, w2, b2, bc2_sqrt, eps, neg_step);      \
  adam_elem(P.z, G.z, M.z, V.z, gs, w1, w2, b2, bc2_sqrt, eps, neg_step);      \
  adam_elem(P.w, G.w, M.w, V.w, gs, w1, w2, b2, bc2_sqrt, eps, neg_step);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
    const int64_t j = i + stride;
    const bool two = j < n4;
    float4 p0 = P4[i], g0 = G4[i], m0 = M4[i], v0 = V4[i];
    float4 p1 = p0, g1 = g0, m1 = m0, v1 = v0;
    if (two) p1 = P4[j], g1 = G4[j], m1 = M4[j], v1 = V4[j];
    B2H_ADAM4(p0, g0, m0, v0)
    P4[i] = p0, M4[i] = m0, V4[i] = v0;
    if (two) {
      B2H_ADAM4(p1, g1, m1, v1)
      P4[j] = p1, M4[j] = m1, V4[j] = v1;
    }
  }
#undef B2H_ADAM4
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += stride) {
    float p = d.p[i], m = d.m[i], v = d.v[i];
    adam_elem(p, d.g[i], m, v, gs, w1, w2, b2, bc2_sqrt, eps, neg_step);
    d.p[i] = p, d.m[i] = m, d.v[i] = v;
  }
}

int launch_adam(const b2h_adam_t& d, cudaStream_t s) {
  B2H_CARVE(adam_step_kernel);
  B2H_CARVE(adam_kernel);
  B2H_CHECK_ARG(d.step && d.scalars && d.phase >= 0 && d.phase <= 2 && (d.phase == 1 || d.n > 0), B2H_ERR_ARG,
                "adam: bad args");
  if (d.phase != 2) {
    launch(adam_step_kernel, 1, 1, 0, s, d);
    B2H_LAUNCH_CHECK("adam_step");
    if (d.phase == 1) return B2H_OK;
  }
  B2H_CHECK_ARG(((uintptr_t)d.p % 16 == 0) && ((uintptr_t)d.g % 16 == 0) && ((uintptr_t)d.m % 16 == 0) &&
                    ((uintptr_t)d.v % 16 == 0),
                B2H_ERR_ALIGN, "adam: buffers must be 16-byte aligned");
  // 4 resident CTAs per SM (64 registers), two float4 of each array per thread and trip: 128 KB in flight per SM
  int blocks = (int)std::min<int64_t>(ceil_div64(d.n / 8 + 1, 256), (int64_t)sm_count() * 4);
  launch(adam_kernel, blocks, 256, 0, s, d);
  B2H_LAUNCH_CHECK("adam");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// weight repack (PyTorch layout fp32 -> GEMM operand layout [Opad*nphase][ntaps][Ipad], act dtype)
// ---------------------------------------------------------------------------------------------
// One thread writes 8 consecutive `i` of one (phase, o, tap) — a single 16 B store in bf16.  The lane-fast index
// is whichever of i / o is closer to contiguous in W, so the fp32 gathers of a warp share sectors: i for the
// forward layouts (W is [o][i][k]), o for the transposed dgrad layouts (W is [i][o][k]).
template <typename T>
__device__ __forceinline__ void pack_body(const b2h_pack_t& d, uint32_t first, uint32_t stride) {
  const uint32_t I8 = (uint32_t)d.Ipad >> 3, Opad = (uint32_t)d.Opad, ntaps = (uint32_t)d.ntaps;
  const uint32_t total = (uint32_t)d.nphase * Opad * ntaps * I8;
  const bool o_fast = d.o_stride < d.i_stride;
  T* out = reinterpret_cast<T*>(d.out);
  for (uint32_t idx = first; idx < total; idx += stride) {
    uint32_t i8, t, o, ph, r;
    if (o_fast) {
      o = idx % Opad, r = idx / Opad;
      t = r % ntaps, r /= ntaps;
      i8 = r % I8, ph = r / I8;
    } else {
      i8 = idx % I8, r = idx / I8;
      t = r % ntaps, r /= ntaps;
      o = r % Opad, ph = r / Opad;
    }
    const int k = d.tapmap[ph][t];
    F8 v;
#pragma unroll
    for (int j = 0; j < 8; ++j) v.v[j] = 0.f;
    if (k >= 0 && (int)o < d.O) {
      const float* w = d.W + (int64_t)o * d.o_stride + (int64_t)k * d.k_stride;
      const int i0 = (int)i8 * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (i0 + j < d.I) v.v[j] = __ldg(w + (int64_t)(i0 + j) * d.i_stride);
    }
    store8<T>(out + (((int64_t)ph * Opad + o) * ntaps + t) * d.Ipad + i8 * 8, v);
  }
  if (d.out_bias) {
    for (uint32_t o = first; o < Opad; o += stride) d.out_bias[o] = (d.bias && (int)o < d.O) ? d.bias[o] : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pack_kernel(b2h_pack_t d) {
  pdl_sync();
  pack_body<T>(d, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

template <typename T>
__global__ void __launch_bounds__(256) pack_multi_kernel(b2h_pack_multi_t m) {
  pdl_sync();
  __shared__ b2h_pack_t d;
  if (threadIdx.x < sizeof(b2h_pack_t) / 4)
    reinterpret_cast<uint32_t*>(&d)[threadIdx.x] = reinterpret_cast<const uint32_t*>(m.descs + blockIdx.y)[threadIdx.x];
  __syncthreads();
  pack_body<T>(d, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

int launch_pack_multi(const b2h_pack_multi_t& m, int dtype, cudaStream_t s) {
  B2H_CARVE(pack_multi_kernel<__nv_bfloat16>);
  B2H_CARVE(pack_multi_kernel<float>);
  B2H_CHECK_ARG(m.descs && m.n > 0 && m.n <= 65535 && m.max_elems > 0, B2H_ERR_ARG, "pack_multi: bad args");
  static_assert(sizeof(b2h_pack_t) % 4 == 0 && sizeof(b2h_pack_t) / 4 <= 256, "descriptor staging");
  int bx = (int)std::min<int64_t>(ceil_div64(m.max_elems, 256 * 8 * 2), 64);
  dim3 grid(std::max(bx, 1), m.n);
  if (dtype == B2H_BF16)
    launch(pack_multi_kernel<__nv_bfloat16>, grid, 256, 0, s, m);
  else
    launch(pack_multi_kernel<float>, grid, 256, 0, s, m);
  B2H_LAUNCH_CHECK("pack_multi");
  return B2H_OK;
}

int launch_pack(const b2h_pack_t& d, int dtype, cudaStream_t s) {
  B2H_CARVE(pack_kernel<__nv_bfloat16>);
  B2H_CARVE(pack_kernel<float>);
  B2H_CHECK_ARG(d.O > 0 && d.I > 0 && d.Opad >= d.O && d.Ipad >= d.I && d.ntaps >= 1 && d.ntaps <= B2H_MAX_TAPS &&
                    d.nphase >= 1 && d.nphase <= 2,
                B2H_ERR_SHAPE, "pack: bad shape");
  B2H_CHECK_ARG(d.Ipad % 8 == 0, B2H_ERR_ALIGN, "pack: Ipad must be a multiple of 8");
  int64_t total = (int64_t)d.nphase * d.Opad * d.ntaps * d.Ipad;
  B2H_CHECK_ARG(total < ((int64_t)1 << 31), B2H_ERR_SHAPE, "pack: operand too large");
  int blocks = (int)std::min<int64_t>(ceil_div64(total, 256 * 8), (int64_t)sm_count() * 8);
  if (dtype == B2H_BF16)
    launch(pack_kernel<__nv_bfloat16>, blocks, 256, 0, s, d);
  else
    launch(pack_kernel<float>, blocks, 256, 0, s, d);
  B2H_LAUNCH_CHECK("pack");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// rot6d -> rotation matrix (utils/conversion_utils.py:86-107), one joint per thread
//   x = a/(|a|+1e-6); z = x X b; z /= (|z|+1e-6); y = z X x; M = [x y z] as columns, row-major
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rot6d_kernel(b2h_rot6d_t d) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2* p = reinterpret_cast<const float2*>(d.r6d + i * 6);
    float2 p0 = p[0], p1 = p[1], p2 = p[2];
    float ax = p0.x, ay = p0.y, az = p1.x, bx = p1.y, by = p2.x, bz = p2.y;
    float na = sqrtf(ax * ax + ay * ay + az * az) + 1e-6f;
    float xx = ax / na, xy = ay / na, xz = az / na;
    float zx = xy * bz - xz * by, zy = xz * bx - xx * bz, zz = xx * by - xy * bx;
    float nz = sqrtf(zx * zx + zy * zy + zz * zz) + 1e-6f;
    zx /= nz, zy /= nz, zz /= nz;
    float yx = zy * xz - zz * xy, yy = zz * xx - zx * xz, yz = zx * xy - zy * xx;
    float* o = d.mat + i * 9;
    o[0] = xx, o[1] = yx, o[2] = zx;
    o[3] = xy, o[4] = yy, o[5] = zy;
    o[6] = xz, o[7] = yz, o[8] = zz;
  }
}

int launch_rot6d(const b2h_rot6d_t& d, cudaStream_t s) {
  B2H_CARVE(rot6d_kernel);
  B2H_CHECK_ARG(d.n > 0, B2H_ERR_SHAPE, "rot6d: n must be positive");
  B2H_CHECK_ARG((uintptr_t)d.r6d % 8 == 0, B2H_ERR_ALIGN, "rot6d: input must be 8-byte aligned");
  int blocks = (int)std::min<int64_t>(ceil_div64(d.n, 256), (int64_t)sm_count() * 8);
  launch(rot6d_kernel, blocks, 256, 0, s, d);
  B2H_LAUNCH_CHECK("rot6d");
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// fk: 6-D rotations -> rotation matrix -> axis-angle -> forward kinematics, one frame per thread
// (conversion_utils.py:33-41 log map as scipy's as_rotvec: unit quaternion, angle = 2 atan2(|v|, w), series for
// small angles; conversion_utils.py:117-137 Rodrigues step per bone)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void r6d_to_rotvec(const float* r, float* aa) {
  const float ax = r[0], ay = r[1], az = r[2], bx = r[3], by = r[4], bz = r[5];
  const float na = sqrtf(ax * ax + ay * ay + az * az) + 1e-6f;
  const float xx = ax / na, xy = ay / na, xz = az / na;
  float zx = xy * bz - xz * by, zy = xz * bx - xx * bz, zz = xx * by - xy * bx;
  const float nz = sqrtf(zx * zx + zy * zy + zz * zz) + 1e-6f;
  zx /= nz, zy /= nz, zz /= nz;
  const float yx = zy * xz - zz * xy, yy = zz * xx - zx * xz, yz = zx * xy - zy * xx;
  // M = [x y z] as columns
  const float m00 = xx, m01 = yx, m02 = zx, m10 = xy, m11 = yy, m12 = zy, m20 = xz, m21 = yz, m22 = zz;
  float w, qx, qy, qz;
  const float tr = m00 + m11 + m22;
  if (tr > 0.f) {
    const float s = sqrtf(tr + 1.f) * 2.f;
    w = 0.25f * s, qx = (m21 - m12) / s, qy = (m02 - m20) / s, qz = (m10 - m01) / s;
  } else if (m00 > m11 && m00 > m22) {
    const float s = sqrtf(1.f + m00 - m11 - m22) * 2.f;
    w = (m21 - m12) / s, qx = 0.25f * s, qy = (m01 + m10) / s, qz = (m02 + m20) / s;
  } else if (m11 > m22) {
    const float s = sqrtf(1.f + m11 - m00 - m22) * 2.f;
    w = (m02 - m20) / s, qx = (m01 + m10) / s, qy = 0.25f * s, qz = (m12 + m21) / s;
  } else {
    const float s = sqrtf(1.f + m22 - m00 - m11) * 2.f;
    w = (m10 - m01) / s, qx = (m02 + m20) / s, qy = (m12 + m21) / s, qz = 0.25f * s;
  }
  const float qn = rsqrtf(w * w + qx * qx + qy * qy + qz * qz);
  w *= qn, qx *= qn, qy *= qn, qz *= qn;
  if (w < 0.f) w = -w, qx = -qx, qy = -qy, qz = -qz;
  const float vn = sqrtf(qx * qx + qy * qy + qz * qz);
  const float angle = 2.f * atan2f(vn, w);
  const float a2 = angle * angle;
  const float scale = angle <= 1e-3f ? 2.f + a2 / 12.f + 7.f * a2 * a2 / 2880.f : angle / sinf(0.5f * angle);
  aa[0] = scale * qx, aa[1] = scale * qy, aa[2] = scale * qz;
}

__global__ void __launch_bounds__(128) fk_kernel(b2h_fk_t d) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n) return;
  float xyz[(B2H_FK_MAX_BONES + 1) * 3];
#pragma unroll
  for (int k = 0; k < 6; ++k) xyz[k] = d.root[k];
  const float* row = d.r6d + i * d.ld;
  for (int b = 1; b < d.nbones; ++b) {
    float r[6], aa[3];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int c = (b - 1) * 6 + k;
      float v = row[c];
      if (d.mean) v = fmaf(v, d.std[c], d.mean[c]);
      r[k] = v;
    }
    r6d_to_rotvec(r, aa);
    if (d.aa) {
      float* o = d.aa + i * (int64_t)(d.nbones - 1) * 3 + (b - 1) * 3;
      o[0] = aa[0], o[1] = aa[1], o[2] = aa[2];
    }
    const int j = d.joint[b], pb = d.before[b];
    const float jx = xyz[j * 3], jy = xyz[j * 3 + 1], jz = xyz[j * 3 + 2];
    float ux = jx - xyz[pb * 3], uy = jy - xyz[pb * 3 + 1], uz = jz - xyz[pb * 3 + 2];
    const float un = rsqrtf(ux * ux + uy * uy + uz * uz);
    ux *= un, uy *= un, uz *= un;
    const float th = sqrtf(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
    float vx = ux, vy = uy, vz = uz;
    if (th > 0.f) {
      const float ex = aa[0] / th, ey = aa[1] / th, ez = aa[2] / th;
      float sn, cs;
      sincosf(th, &sn, &cs);
      const float dot = ex * ux + ey * uy + ez * uz;
      const float cx = ey * uz - ez * uy, cy = ez * ux - ex * uz, cz = ex * uy - ey * ux;
      vx = ux * cs + cx * sn + ex * dot * (1.f - cs);
      vy = uy * cs + cy * sn + ey * dot * (1.f - cs);
      vz = uz * cs + cz * sn + ez * dot * (1.f - cs);
    }
    const float len = d.bone_len[b];
    xyz[(b + 1) * 3] = jx + len * vx;
    xyz[(b + 1) * 3 + 1] = jy + len * vy;
    xyz[(b + 1) * 3 + 2] = jz + len * vz;
  }
  float* o = d.xyz + i * (int64_t)(d.nbones + 1) * 3;
  for (int k = 0; k < (d.nbones + 1) * 3; ++k) o[k] = xyz[k];
}

int launch_fk(const b2h_fk_t& d, cudaStream_t s) {
  B2H_CARVE(fk_kernel);
  B2H_CHECK_ARG(d.n > 0 && d.nbones >= 2 && d.nbones <= B2H_FK_MAX_BONES && d.ld >= (d.nbones - 1) * 6,
                B2H_ERR_SHAPE, "fk: n=%lld nbones=%d ld=%d", (long long)d.n, d.nbones, d.ld);
  B2H_CHECK_ARG(d.r6d && d.xyz && ((d.mean == nullptr) == (d.std == nullptr)), B2H_ERR_ARG, "fk: null pointers");
  for (int b = 1; b < d.nbones; ++b)
    B2H_CHECK_ARG(d.joint[b] >= 0 && d.joint[b] <= b && d.before[b] >= 0 && d.before[b] <= b &&
                      d.joint[b] != d.before[b],
                  B2H_ERR_ARG, "fk: bone %d must hang off joints that are already placed", b);
  launch(fk_kernel, (unsigned)ceil_div64(d.n, 128), 128, 0, s, d);
  B2H_LAUNCH_CHECK("fk");
  return B2H_OK;
}

int launch_fill(const b2h_fill_t& d, cudaStream_t s) {
  B2H_CHECK_ARG(d.ptr && d.bytes >= 0, B2H_ERR_ARG, "fill: bad args");
  cudaError_t e = cudaMemsetAsync(d.ptr, d.value, (size_t)d.bytes, s);
  if (e != cudaSuccess) return cuda_fail(e, "fill");
  return B2H_OK;
}

}  // namespace b2h
