// Shared device/host helpers for libb2h (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <utility>

#include "../../include/b2h_abi.h"

namespace b2h {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();
extern thread_local int64_t g_launch_count;

#define B2H_CHECK_ARG(cond, code, ...)   \
  do {                                   \
    if (!(cond)) {                       \
      ::b2h::set_error(__VA_ARGS__);     \
      return (code);                     \
    }                                    \
  } while (0)

#define B2H_LAUNCH_CHECK(what)                                   \
  do {                                                           \
    ::b2h::g_launch_count++;                                     \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ::b2h::cuda_fail(e__, what);  \
  } while (0)

// Every kernel asks for the maximum shared-memory carveout so that consecutive launches never force the SM
// to re-partition L1/shared memory (the tensor-core kernels need ~100-200 KB; the elementwise ones need none).
#define B2H_CARVE(...)                                                                                     \
  do {                                                                                                     \
    static bool done__ = false;                                                                            \
    if (!done__) {                                                                                         \
      done__ = true;                                                                                       \
      if (!getenv("B2H_NO_CARVEOUT")) {                                                                    \
        cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributePreferredSharedMemoryCarveout,                  \
                             (int)cudaSharedmemCarveoutMaxShared);                                         \
        cudaGetLastError();                                                                                \
      }                                                                                                    \
    }                                                                                                      \
  } while (0)

// Programmatic dependent launch: every kernel of the library is launched with the programmatic-stream-
// serialization attribute, so inside a stream (or a captured graph) kernel N+1 is scheduled while kernel N
// still runs.  Each kernel executes pdl_sync() before its first global-memory access: it blocks until the
// preceding grid has completed and flushed, then lets the next grid start its own launch/prologue.  What
// overlaps is launch latency, CTA scheduling and the per-CTA prologue (barrier init, TMEM alloc, descriptor
// prefetch) — no data is touched early.  B2H_NO_PDL=1 launches plainly (pdl_sync() is then a no-op).
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("B2H_NO_PDL") ? 0 : 1;
  return on == 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// same, as thread-block clusters of `cluster_x` CTAs along x (CTA pairs of the cta_group::2 kernels)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                  unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

constexpr float kLeakySlope = 0.2f;  // nn.LeakyReLU(0.2, True), modelZoo.py:195

// ---------------------------------------------------------------------------------------------
// 4-wide vector access in the activation dtype
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive channels: one 16-byte access in bf16, two in fp32
struct F8 {
  float v[8];
};
template <typename T>
__device__ __forceinline__ F8 load8(const T* p);
template <>
__device__ __forceinline__ F8 load8<float>(const float* p) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  F8 r;
  r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
  return r;
}
template <>
__device__ __forceinline__ F8 load8<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  F8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
  return r;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const F8& r);
template <>
__device__ __forceinline__ void store8<float>(float* p, const F8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const F8& r) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 q = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&q);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ float& f4(float4& v, int i) { return (&v.x)[i]; }
__device__ __forceinline__ float f4c(const float4& v, int i) { return (&v.x)[i]; }

// ---------------------------------------------------------------------------------------------
// Dropout: explicit keep-mask or Philox4x32-10 keyed by (seed, step, site, element index)
// ---------------------------------------------------------------------------------------------
// out of line and rolled on purpose: one small copy per kernel instead of ~1 KB per call site (these
// kernels are instruction-fetch bound when their straight-line code outgrows the instruction caches)
static __device__ __noinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll 2
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// Dropout keep flags of a site whose tensor is [rows][C].
//   MASK:   explicit uint8 flags mask[row*C + c]
//   PHILOX: one Philox4x32-10 call yields the 128 keep bits of a (16 rows x 8 channels) block:
//           counter = (row >> 4) * ceil(C/8) + (c >> 3), bit = (row & 15) * 8 + (c & 7),
//           keyed by (seed, step, site) -> a thread that owns 8 channels of 8 consecutive rows needs ONE call.
struct DropCtx {
  int mode;
  uint32_t site;
  const uint8_t* mask;
  uint8_t* save;
  uint2 key;      // seed
  uint32_t step;  // low 32 bits of the step counter
  uint32_t C, Cq; // row length of the site's tensor, ceil(C / 8)
  __device__ __forceinline__ void init(const b2h_dropout_t& d, int row_len) {
    mode = d.mode;
    site = (uint32_t)d.site;
    mask = d.mask;
    save = d.save;
    key = make_uint2(0u, 0u);
    step = 0u;
    C = (uint32_t)max(row_len, 1);
    Cq = (C + 7u) >> 3;
    if (mode == B2H_DROP_PHILOX) {
      uint64_t seed = d.state[0], st = d.state[1];
      key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
      step = (uint32_t)st;
    }
  }
  // PHILOX: the 64 keep bits of rows 8*rb8 .. 8*rb8+7 for channel octet cq (byte j = row 8*rb8+j)
  __device__ __forceinline__ uint2 philox_rows8(uint32_t rb8, uint32_t cq) const {
    const uint64_t c = (uint64_t)(rb8 >> 1) * Cq + cq;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), site, step), key);
    return (rb8 & 1u) ? make_uint2(r.z, r.w) : make_uint2(r.x, r.y);
  }
  // keep bits (bit i = channel c0 + i) of 8 channels of one row; c0 % 8 == 0
  __device__ __forceinline__ uint32_t keep8_rc(uint32_t row, uint32_t c0) const {
    if (mode == B2H_DROP_NONE) return 0xFFu;
    if (mode == B2H_DROP_PHILOX) {
      const uint2 w = philox_rows8(row >> 3, c0 >> 3);
      const uint32_t j = row & 7u;
      return ((j < 4u ? w.x : w.y) >> ((j & 3u) * 8u)) & 0xFFu;
    }
    const uint64_t idx = (uint64_t)row * C + c0;
    uint32_t b = 0;
    const uint32_t n = min(8u, C - c0);
    if (n == 8u && (idx & 3u) == 0u) {
      const uint32_t m0 = *reinterpret_cast<const uint32_t*>(mask + idx);
      const uint32_t m1 = *reinterpret_cast<const uint32_t*>(mask + idx + 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        b |= ((m0 >> (8 * i)) & 0xFFu) ? (1u << i) : 0u;
        b |= ((m1 >> (8 * i)) & 0xFFu) ? (16u << i) : 0u;
      }
      return b;
    }
#pragma unroll 1
    for (uint32_t i = 0; i < n; ++i) b |= mask[idx + i] ? (1u << i) : 0u;
    return b;
  }
  // generic element access by flat index idx = row*C + c (rare paths: one division)
  __device__ __forceinline__ float scale1(uint64_t idx) const {
    if (mode == B2H_DROP_NONE) return 1.f;
    if (mode == B2H_DROP_MASK) return mask[idx] ? 2.f : 0.f;
    const uint32_t row = (uint32_t)(idx / C), c = (uint32_t)(idx - (uint64_t)row * C);
    return ((keep8_rc(row, c & ~7u) >> (c & 7u)) & 1u) ? 2.f : 0.f;
  }
  __device__ __forceinline__ float4 scale4(uint64_t idx) const {
    if (mode == B2H_DROP_NONE) return make_float4(1.f, 1.f, 1.f, 1.f);
    if (mode == B2H_DROP_PHILOX && (C & 3u) == 0u && (idx & 3u) == 0u) {
      const uint32_t row = (uint32_t)(idx / C), c = (uint32_t)(idx - (uint64_t)row * C);
      const uint32_t b = keep8_rc(row, c & ~7u) >> (c & 7u);
      return make_float4((b & 1u) ? 2.f : 0.f, (b & 2u) ? 2.f : 0.f, (b & 4u) ? 2.f : 0.f, (b & 8u) ? 2.f : 0.f);
    }
    return make_float4(scale1(idx), scale1(idx + 1), scale1(idx + 2), scale1(idx + 3));
  }
  // record the keep flags of elements idx .. idx+n-1 (n <= 4) for the backward pass
  __device__ __forceinline__ void save4(uint64_t idx, float4 m, int n) const {
    if (!save) return;
    if (n == 4 && (idx & 3) == 0) {
      uchar4 k = make_uchar4(m.x != 0.f, m.y != 0.f, m.z != 0.f, m.w != 0.f);
      *reinterpret_cast<uchar4*>(save + idx) = k;
    } else {
      for (int i = 0; i < n; ++i) save[idx + i] = (&m.x)[i] != 0.f;
    }
  }
  // record the keep flags of n <= 8 elements (bit i = element idx + i)
  __device__ __forceinline__ void save8(uint64_t idx, uint32_t bits, int n) const {
    if (!save) return;
    if (n == 8 && (idx & 7) == 0) {
      uint2 w;
      w.x = (bits & 1u) | ((bits & 2u) << 7) | ((bits & 4u) << 14) | ((bits & 8u) << 21);
      w.y = ((bits >> 4) & 1u) | ((bits & 32u) << 3) | ((bits & 64u) << 10) | ((bits & 128u) << 17);
      *reinterpret_cast<uint2*>(save + idx) = w;
    } else {
      for (int i = 0; i < n; ++i) save[idx + i] = (bits >> i) & 1u;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// "last block done" ticket: the last CTA of a grid performs the ordered final reduction
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool last_block_done(uint32_t* ticket, uint32_t nblocks) {
  __shared__ int s_is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    uint32_t t = atomicAdd(ticket, 1u);
    s_is_last = (t == nblocks - 1u);
    if (s_is_last) *ticket = 0u;  // self reset for the next launch / graph replay
  }
  __syncthreads();
  if (s_is_last) __threadfence();
  return s_is_last != 0;
}

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == B2H_ACT_LEAKY) return v > 0.f ? v : v * kLeakySlope;
  if (act == B2H_ACT_RELU) return v > 0.f ? v : 0.f;
  return v;
}
// derivative expressed on the OUTPUT of the in-place activation (sign(out) == sign(pre))
__device__ __forceinline__ float act_bwd(float out, int act) {
  if (act == B2H_ACT_LEAKY) return out > 0.f ? 1.f : kLeakySlope;
  if (act == B2H_ACT_RELU) return out > 0.f ? 1.f : 0.f;
  return 1.f;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// launchers implemented in the .cu files
int launch_gemm_f32(const b2h_gemm_t& d, cudaStream_t s);
int launch_wgrad_f32(const b2h_wgrad_t& d, cudaStream_t s);
int launch_gemm_bf16(const b2h_gemm_t& d, void* cache, cudaStream_t s);
int launch_wgrad_bf16(const b2h_wgrad_t& d, void* cache, cudaStream_t s);
int wgrad_choose_splits(const b2h_wgrad_t& d, int dtype);
int launch_wgrad_reduce(const b2h_wgrad_t& d, int splits, cudaStream_t s);

int launch_bn_stats(const b2h_bn_stats_t& d, int dtype, cudaStream_t s);
int launch_bn_apply(const b2h_bn_apply_t& d, int dtype, cudaStream_t s);
int launch_bn_bwd(const b2h_bn_bwd_t& d, int dtype, cudaStream_t s);
int launch_fk(const b2h_fk_t& d, cudaStream_t s);
int launch_bwd_sums_separate(const b2h_gemm_t& g, int dtype, cudaStream_t s);
int launch_prep(const b2h_prep_t& d, int dtype, cudaStream_t s);
int launch_to_ncl(const b2h_to_ncl_t& d, int dtype, cudaStream_t s);
int launch_l1(const b2h_l1_t& d, int dtype, cudaStream_t s);
int launch_mse(const b2h_mse_t& d, cudaStream_t s);
int launch_colsum(const b2h_colsum_t& d, int dtype, cudaStream_t s);
int launch_adam(const b2h_adam_t& d, cudaStream_t s);
int launch_pack(const b2h_pack_t& d, int dtype, cudaStream_t s);
int launch_pack_multi(const b2h_pack_multi_t& d, int dtype, cudaStream_t s);
int launch_bn_fold(const b2h_bn_fold_t& d, cudaStream_t s);
int launch_bn_fold_multi(const b2h_bn_fold_multi_t& d, cudaStream_t s);
int launch_rot6d(const b2h_rot6d_t& d, cudaStream_t s);
int launch_fill(const b2h_fill_t& d, cudaStream_t s);
int launch_dp_adam(const b2h_dp_adam_t& d, cudaStream_t s);

constexpr int kBnChunkRows = 64;  // rows per CTA of the BN / colsum reductions
inline int bn_nchunks(int rows_per_group) { return ceil_div(rows_per_group, kBnChunkRows); }

}  // namespace b2h
