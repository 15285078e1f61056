// Cross-CTA part of the train-mode BatchNorm statistics, shared by bn_stats_kernel and by the tcgen05 GEMM
// epilogue that produces the statistics of its own output tile (k_gemm_tc.cu).
#pragma once
#include "b2h_common.cuh"

namespace b2h {

// accumulator copies: CTA i adds into copy i % 16 (same-address fp64 atomics serialise in L2)
constexpr int kCopies = 16;

// accumulators: [kCopies][groups][C][2] doubles (shifted sum, shifted sum of squares), zero between launches
__device__ __forceinline__ void bn_stats_accumulate(const b2h_bn_stats_t& d, int copy, int g, int c, float s1,
                                                    float s2) {
  double* a = reinterpret_cast<double*>(d.partial) + (((int64_t)copy * d.groups + g) * d.C + c) * 2;
  atomicAdd(a + 0, (double)s1);
  atomicAdd(a + 1, (double)s2);
}

// Run by every thread (tid of nthreads) of the LAST CTA of the launch: accumulators -> mean / invstd / folded
// scale / shift, running-statistics update, accumulators re-zeroed for the next launch.
__device__ __forceinline__ void bn_stats_finalize(const b2h_bn_stats_t& d, int tid, int nthreads) {
  double* accum = reinterpret_cast<double*>(d.partial);
  const int rpg = d.rows_per_group;
#pragma unroll 1
  for (int c = tid; c < d.C; c += nthreads) {
    const double p = d.running_mean ? (double)d.running_mean[c] : 0.0;
#pragma unroll 1
    for (int gg = 0; gg < d.groups; ++gg) {
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int k = 0; k < kCopies; ++k) {   // fixed order over the copies
        double2* acc = reinterpret_cast<double2*>(accum + (((int64_t)k * d.groups + gg) * d.C + c) * 2);
        const double2 v = __ldcg(acc);
        *acc = make_double2(0.0, 0.0);
        t0 += v.x;
        t1 += v.y;
      }
      const double dm = t0 / (double)rpg;      // mean - pivot
      const double mean = p + dm;
      double m2 = t1 - t0 * dm;
      if (m2 < 0.0) m2 = 0.0;
      const double var_b = m2 / (double)rpg;
      const float invstd = (float)(1.0 / sqrt(var_b + (double)d.eps));
      const float scale = invstd * (d.gamma ? d.gamma[c] : 1.f);
      d.mean[gg * d.Cs + c] = (float)mean;
      d.invstd[gg * d.Cs + c] = invstd;
      d.scale[gg * d.Cs + c] = scale;
      d.shift[gg * d.Cs + c] = (d.beta ? d.beta[c] : 0.f) - (float)mean * scale;
      if (d.running_mean && (gg == 0 || d.update_all_groups)) {
        const double var_u = rpg > 1 ? m2 / (double)(rpg - 1) : var_b;
        const float mom = d.momentum;
        d.running_mean[c] = (1.f - mom) * d.running_mean[c] + mom * (float)mean;
        d.running_var[c] = (1.f - mom) * d.running_var[c] + mom * (float)var_u;
      }
    }
  }
  if (tid == 0 && d.running_mean && d.num_batches_tracked)
    *d.num_batches_tracked += d.update_all_groups ? d.groups : 1;
}

}  // namespace b2h
