// Shared epilogue of the tap-GEMM kernels (fp32 FFMA path and bf16 tcgen05 path):
//   v = acc + bias -> activation -> (eval BN scale/shift) -> (dropout keep*2 of the dgrad site)
// and the sub-pixel (2-phase) output addressing used by ConvTranspose1d forward / strided dgrad.
#pragma once
#include <stdlib.h>

#include "b2h_common.cuh"

namespace b2h {

struct EpiParams {
  const float* bias;
  const float* post_scale;
  const float* post_shift;
  void* out;
  int Lo, Lo_actual, ldo, out_coff, Nvalid, nphase, half;  // half = Npad / nphase
  int act, out_f32;
  int drop_C;
  b2h_dropout_t drop;
  const void* resid;   // residual add (persistent kernel only, see b2h_gemm_t)
  int ld_resid, resid_up2, out_pool2;
  const void* grad_add;   // dgrad: the gradient another consumer of the same tensor wrote (bf16 one-tile kernel only)
  int ld_grad_add;
};

inline EpiParams make_epi(const b2h_gemm_t& d) {
  EpiParams e;
  e.bias = d.bias;
  e.post_scale = d.post_scale;
  e.post_shift = d.post_shift;
  e.out = d.out;
  e.Lo = d.Lo;
  e.Lo_actual = d.Lo_actual;
  e.ldo = d.ldo;
  e.out_coff = d.out_coff;
  e.Nvalid = d.Nvalid;
  e.nphase = d.nphase;
  e.half = d.Npad / d.nphase;
  e.act = d.act;
  e.out_f32 = d.out_f32;
  e.drop_C = d.drop_C;
  e.drop = d.drop;
  e.resid = d.resid;
  e.ld_resid = d.ld_resid;
  e.resid_up2 = d.resid_up2;
  e.out_pool2 = d.out_pool2;
  e.grad_add = d.grad_add;
  e.ld_grad_add = d.ld_grad_add;
  return e;
}

// Finish and store `NV` consecutive accumulator columns [n, n+NV) of GEMM row (b, lo).
// NV is 4 or 8; n is a multiple of NV, so a vector never straddles a phase boundary.
template <typename T, int NV>
__device__ __forceinline__ void epilogue_store(const EpiParams& e, const DropCtx& drop, int b, int lo, int n,
                                               const float* acc) {
  const int ph = n / e.half;
  const int nn = n - ph * e.half;
  const int row_in_sample = lo * e.nphase + ph;
  if (nn >= e.Nvalid || row_in_sample >= e.Lo_actual) return;
  const int64_t grow = (int64_t)b * e.Lo_actual + row_in_sample;
  float v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float x = acc[j];
    if (e.bias) x += e.bias[nn + j];
    x = act_fwd(x, e.act);
    if (e.post_scale) x = fmaf(x, e.post_scale[nn + j], e.post_shift[nn + j]);
    v[j] = x;
  }
  if (drop.mode != B2H_DROP_NONE) {
#pragma unroll
    for (int j = 0; j < NV; j += 4) {
      if (nn + j + 3 < e.drop_C) {
        float4 m = drop.scale4((uint64_t)grow * e.drop_C + nn + j);
        v[j] *= m.x, v[j + 1] *= m.y, v[j + 2] *= m.z, v[j + 3] *= m.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (nn + j + k < e.drop_C) v[j + k] *= drop.scale1((uint64_t)grow * e.drop_C + nn + j + k);
      }
    }
  }
  const int64_t o = grow * e.ldo + e.out_coff + nn;
  if (nn + NV <= e.Nvalid) {
    if (e.out_f32) {
      float* p = reinterpret_cast<float*>(e.out) + o;
#pragma unroll
      for (int j = 0; j < NV; j += 4) store4<float>(p + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    } else {
      T* p = reinterpret_cast<T*>(e.out) + o;
#pragma unroll
      for (int j = 0; j < NV; j += 4) store4<T>(p + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (nn + j < e.Nvalid) {
        if (e.out_f32)
          reinterpret_cast<float*>(e.out)[o + j] = v[j];
        else
          reinterpret_cast<T*>(e.out)[o + j] = from_f<T>(v[j]);
      }
    }
  }
}

// finish 8 accumulator columns [nn, nn+8) of one output row: bias -> act -> eval-BN -> dropout keep*2
__device__ __forceinline__ void epi_finish8(const EpiParams& e, const DropCtx& drop, uint64_t drop_row_base, int nn,
                                            const uint32_t* acc_bits, float* v) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float x = __uint_as_float(acc_bits[j]);
    if (e.bias) x += __ldg(e.bias + nn + j);
    x = act_fwd(x, e.act);
    if (e.post_scale) x = fmaf(x, __ldg(e.post_scale + nn + j), __ldg(e.post_shift + nn + j));
    v[j] = x;
  }
  if (drop.mode != B2H_DROP_NONE) {
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      if (nn + j + 3 < e.drop_C) {
        float4 m = drop.scale4(drop_row_base + nn + j);
        v[j] *= m.x, v[j + 1] *= m.y, v[j + 2] *= m.z, v[j + 3] *= m.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (nn + j + k < e.drop_C) v[j + k] *= drop.scale1(drop_row_base + nn + j + k);
      }
    }
  }
}


}  // namespace b2h
