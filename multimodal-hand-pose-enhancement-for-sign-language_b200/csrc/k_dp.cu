// Data-parallel optimiser step fused with its collective (include/b2h_abi.h, b2h_dp_adam_t): reduce-scatter of the
// flat gradients over NVLink peer memory, Adam on the owned slice, all-gather of the updated parameters, in ONE
// kernel.  Per step and rank the wire carries n*4*(world-1)/world bytes in (gradient slices) and the same out
// (parameter slices) — a two-shot all-reduce's traffic — but the optimizer runs between the two shots on 1/world
// of the elements, there is no separate Adam launch and no NCCL launch / proxy latency on the critical path.
// With multicast addresses (NVLS) the reduction happens in the NVSwitch: one multimem.ld_reduce per 16 bytes
// instead of `world` peer loads, one multimem.st instead of `world` peer stores.
#include "b2h_common.cuh"

namespace b2h {

namespace {

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// signal protocol (one uint32 per (CTA, source rank) in the receiver's pad; 0 = empty, 1 = signalled):
// put: CAS 0 -> 1 with release semantics at system scope, spinning while the previous signal is unconsumed;
// wait: CAS 1 -> 0 with acquire semantics.  Self-resetting, so a pad is reusable by the next call at once.
__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// Barrier between CTA `blockIdx.x` of every rank.  bar.sync orders the CTA's earlier accesses before the signalling
// threads; release / acquire at system scope are cumulative, so after the barrier every thread of the CTA observes
// what every thread of the peer CTAs wrote before it (and the peers' earlier kernels in stream order).
__device__ __forceinline__ void peer_barrier(const b2h_dp_adam_t& d, uint64_t timeout_ns) {
  __syncthreads();
  if ((int)threadIdx.x < d.world) {
    const int q = threadIdx.x;
    const uint64_t t0 = globaltimer_ns();
    uint32_t* put = d.signal[q] + (size_t)blockIdx.x * d.world + d.rank;
    while (cas_release_sys(put, 0u, 1u) != 0u) {
      if (globaltimer_ns() - t0 > timeout_ns) __trap();
    }
    uint32_t* get = d.signal[d.rank] + (size_t)blockIdx.x * d.world + q;
    while (cas_acquire_sys(get, 1u, 0u) != 1u) {
      if (globaltimer_ns() - t0 > timeout_ns) __trap();
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_f4(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// NVLS: the switch returns the sum over all ranks' copies / stores to all ranks' copies
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

}  // namespace

// W = compile-time bound of the world size (2, 4, 8), U = elements (float4) per thread and iteration: small worlds
// have few peer loads per element, so several elements are fetched before the first add to keep the NVLink latency
// covered (W * U = 8 sixteen-byte loads in flight per thread in every configuration)
template <int W, int U>
__global__ void __launch_bounds__(512) dp_adam_kernel(b2h_dp_adam_t d) {
  pdl_sync();
  const uint64_t timeout_ns = (uint64_t)(d.timeout_ms > 0 ? d.timeout_ms : 10000) * 1000000ull;
  peer_barrier(d, timeout_ns);   // all gradients complete; nobody reads the old parameters any more

  const float neg_step = d.scalars[0];
  const float bc2_sqrt = d.scalars[1];
  const float w1 = (float)(1.0 - d.beta1), w2 = (float)(1.0 - d.beta2);
  const float b2 = (float)d.beta2, eps = (float)d.eps, gs = d.gscale;
  const int64_t n4 = d.n >> 2;
  const int64_t chunk = (n4 + d.world - 1) / d.world;                      // float4s per rank
  const int64_t lo = chunk * d.rank < n4 ? chunk * d.rank : n4;
  const int64_t hi = lo + chunk < n4 ? lo + chunk : n4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float* p_own = d.p[d.rank];
  for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * U) {
    float4 g[U];
    if (d.g_mc) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < hi) g[u] = multimem_ld_reduce_f4(d.g_mc + i * 4);
      }
    } else {
      float4 part[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * stride;
#pragma unroll
        for (int q = 0; q < W; ++q)      // all loads in flight before the first add
          if (q < d.world && i < hi) part[u][q] = ld_sys_f4(d.g[q] + i * 4);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < W; ++q)      // rank order: the same sum whoever owns the slice
          if (q < d.world) {
            g[u].x += part[u][q].x, g[u].y += part[u][q].y, g[u].z += part[u][q].z, g[u].w += part[u][q].w;
          }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= hi) break;
      float4 p = reinterpret_cast<const float4*>(p_own)[i];
      float4 m = reinterpret_cast<float4*>(d.m)[i];
      float4 v = reinterpret_cast<float4*>(d.v)[i];
#pragma unroll
      for (int k = 0; k < 4; ++k) {                     // the arithmetic of adam_kernel (k_misc.cu), same order
        float gk = f4(g[u], k) * gs;
        float mk = f4(m, k) + w1 * (gk - f4(m, k));
        float vk = f4(v, k) * b2 + (w2 * gk) * gk;
        float denom = sqrtf(vk) / bc2_sqrt + eps;
        f4(p, k) += (neg_step * mk) / denom;
        f4(m, k) = mk;
        f4(v, k) = vk;
      }
      reinterpret_cast<float4*>(d.m)[i] = m;
      reinterpret_cast<float4*>(d.v)[i] = v;
      if (d.p_mc) {
        multimem_st_f4(d.p_mc + i * 4, p);
      } else {
#pragma unroll
        for (int q = 0; q < W; ++q)
          if (q < d.world) st_sys_f4(d.p[q] + i * 4, p);
      }
    }
  }
  __threadfence_system();
  peer_barrier(d, timeout_ns);   // every rank's parameters complete; the gradients may be overwritten
}

int dp_adam_blocks(int64_t n, int world) {
  const int64_t chunk = ceil_div64(n / 4, world);
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(chunk, 512), B2H_DP_MAX_BLOCKS));
}

int launch_dp_adam(const b2h_dp_adam_t& d, cudaStream_t s) {
  B2H_CHECK_ARG(d.world >= 1 && d.world <= B2H_DP_MAX_PEERS && d.rank >= 0 && d.rank < d.world, B2H_ERR_ARG,
                "dp_adam: bad rank / world");
  B2H_CHECK_ARG(d.n > 0 && d.n % 4 == 0, B2H_ERR_SHAPE, "dp_adam: n must be a positive multiple of 4");
  B2H_CHECK_ARG(d.m && d.v && d.scalars, B2H_ERR_ARG, "dp_adam: null moments / scalars");
  B2H_CHECK_ARG(((uintptr_t)d.m % 16 == 0) && ((uintptr_t)d.v % 16 == 0) && ((uintptr_t)d.g_mc % 16 == 0) &&
                    ((uintptr_t)d.p_mc % 16 == 0),
                B2H_ERR_ALIGN, "dp_adam: buffers must be 16-byte aligned");
  for (int q = 0; q < d.world; ++q) {
    B2H_CHECK_ARG(d.p[q] && d.g[q] && d.signal[q], B2H_ERR_ARG, "dp_adam: null peer pointer");
    B2H_CHECK_ARG(((uintptr_t)d.p[q] % 16 == 0) && ((uintptr_t)d.g[q] % 16 == 0), B2H_ERR_ALIGN,
                  "dp_adam: buffers must be 16-byte aligned");
  }
  // few CTAs by design: the kernel spins on its peers, and two of them (generator / discriminator) may be
  // resident at once — together they must never be able to fill the GPU
  const int blocks = dp_adam_blocks(d.n, d.world);
  if (d.world <= 2)
    launch(dp_adam_kernel<2, 4>, blocks, 512, 0, s, d);
  else if (d.world <= 4)
    launch(dp_adam_kernel<4, 2>, blocks, 512, 0, s, d);
  else if (d.world <= 8)
    launch(dp_adam_kernel<8, 1>, blocks, 512, 0, s, d);
  else
    launch(dp_adam_kernel<B2H_DP_MAX_PEERS, 1>, blocks, 512, 0, s, d);
  B2H_LAUNCH_CHECK("dp_adam");
  return B2H_OK;
}

}  // namespace b2h
