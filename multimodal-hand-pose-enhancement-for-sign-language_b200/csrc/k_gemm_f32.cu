// fp32 precision mode of the tap-GEMM family: shared-memory tiled FFMA kernels.
// The 1e-5 parity bar of fp32 mode rules out single-pass TF32 (1.4e-3), so this
// mode keeps every multiply-accumulate in fp32 on the CUDA cores.  The bf16 mode (k_gemm_tc.cu)
// runs the same descriptors on the tcgen05 tensor cores.
#include "gemm_epilogue.cuh"

namespace b2h {

// ---------------------------------------------------------------------------------------------
// fprop-like: out[b, lo, n] = epi( sum_t sum_c A[b, lo*stride + off_t, c] * W[n, t, c] )
// CTA tile 128 (rows) x 64 (cols), BK = 16, 256 threads, 8x4 micro tile
// ---------------------------------------------------------------------------------------------
constexpr int F_BM = 128, F_BN = 64, F_BK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(b2h_gemm_t d, EpiParams e) {
  pdl_sync();
  __shared__ __align__(16) float As[F_BK][F_BM + 4];
  __shared__ __align__(16) float Bs[F_BK][F_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * F_BM, n0 = blockIdx.y * F_BN;
  const int M = d.B * d.Lo;
  const float* A = reinterpret_cast<const float*>(d.A);
  const float* W = reinterpret_cast<const float*>(d.W);
  const int Ktot = d.ntaps * d.Kc;

  // the two A rows this thread loads (fixed over the K loop)
  int a_row[2], a_b[2], a_lo[2], a_kq[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    int f = tid + j * 256;
    a_row[j] = f >> 2;
    a_kq[j] = f & 3;
    int m = m0 + a_row[j];
    if (m < M) {
      a_b[j] = m / d.Lo;
      a_lo[j] = m - a_b[j] * d.Lo;
    } else {
      a_b[j] = -1;
      a_lo[j] = 0;
    }
  }
  const int w_n = tid >> 2, w_kq = tid & 3;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < d.ntaps; ++t) {
    const float* a_ptr[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int li = a_lo[j] * d.stride + d.tap_off[t];
      a_ptr[j] = (a_b[j] >= 0 && li >= 0 && li < d.La) ? A + ((int64_t)a_b[j] * d.La + li) * d.lda + a_kq[j] * 4
                                                       : nullptr;
    }
    const float* w_ptr = W + (int64_t)(n0 + w_n) * Ktot + (int64_t)t * d.Kc + w_kq * 4;
    for (int kc = 0; kc < d.Kc; kc += F_BK) {
      float4 av[2], wv;
#pragma unroll
      for (int j = 0; j < 2; ++j)
        av[j] = a_ptr[j] ? *reinterpret_cast<const float4*>(a_ptr[j] + kc) : make_float4(0, 0, 0, 0);
      wv = *reinterpret_cast<const float4*>(w_ptr + kc);
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        As[a_kq[j] * 4 + 0][a_row[j]] = av[j].x;
        As[a_kq[j] * 4 + 1][a_row[j]] = av[j].y;
        As[a_kq[j] * 4 + 2][a_row[j]] = av[j].z;
        As[a_kq[j] * 4 + 3][a_row[j]] = av[j].w;
      }
      Bs[w_kq * 4 + 0][w_n] = wv.x;
      Bs[w_kq * 4 + 1][w_n] = wv.y;
      Bs[w_kq * 4 + 2][w_n] = wv.z;
      Bs[w_kq * 4 + 3][w_n] = wv.w;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < F_BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
  DropCtx drop;
  drop.init(e.drop, e.drop_C);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    int b = m / d.Lo, lo = m - b * d.Lo;
    epilogue_store<float, 4>(e, drop, b, lo, n0 + tx * 4, acc[i]);
  }
}

static int check_gemm(const b2h_gemm_t& d) {
  B2H_CHECK_ARG(d.B > 0 && d.La > 0 && d.Lo > 0 && d.ntaps >= 1 && d.ntaps <= B2H_MAX_TAPS, B2H_ERR_SHAPE,
                "gemm: bad shape B=%d La=%d Lo=%d ntaps=%d", d.B, d.La, d.Lo, d.ntaps);
  B2H_CHECK_ARG(d.Kc > 0 && d.Kc % 64 == 0 && d.Npad > 0 && d.Npad % 64 == 0, B2H_ERR_SHAPE,
                "gemm: Kc=%d / Npad=%d must be multiples of 64", d.Kc, d.Npad);
  B2H_CHECK_ARG(d.nphase == 1 || (d.nphase == 2 && (d.Npad / 2) % 64 == 0), B2H_ERR_SHAPE, "gemm: bad nphase");
  B2H_CHECK_ARG(d.lda >= d.Kc && d.lda % 8 == 0 && d.ldo % 4 == 0 && d.out_coff % 4 == 0, B2H_ERR_ALIGN,
                "gemm: lda=%d ldo=%d coff=%d alignment", d.lda, d.ldo, d.out_coff);
  B2H_CHECK_ARG(d.Nvalid > 0 && d.Nvalid <= d.Npad / d.nphase && d.Lo_actual > 0, B2H_ERR_SHAPE, "gemm: Nvalid/Lo_actual");
  B2H_CHECK_ARG(d.stride == 1 || d.stride == 2, B2H_ERR_SHAPE, "gemm: stride must be 1 or 2");
  B2H_CHECK_ARG(d.out_f32 == 0 || d.out_f32 == 1, B2H_ERR_ARG, "gemm: NCL output (out_f32 = 2) is a bf16-mode feature");
  B2H_CHECK_ARG((d.post_scale == nullptr) == (d.post_shift == nullptr), B2H_ERR_ARG, "gemm: post scale/shift");
  B2H_CHECK_ARG(!d.grad_add, B2H_ERR_ARG, "gemm: grad_add is a bf16-mode feature");
  return B2H_OK;
}

int launch_gemm_f32(const b2h_gemm_t& d, cudaStream_t s) {
  B2H_CARVE(gemm_f32_kernel);
  int rc = check_gemm(d);
  if (rc) return rc;
  dim3 grid(ceil_div(d.B * d.Lo, F_BM), d.Npad / F_BN);
  launch(gemm_f32_kernel, grid, 256, 0, s, d, make_epi(d));
  B2H_LAUNCH_CHECK("gemm_f32");
  if (d.stats.z) {   // fp32 path: the statistics of the output are a separate pass
    B2H_CHECK_ARG(d.stats.z == d.out && d.out_coff == 0, B2H_ERR_ARG, "gemm: stats must describe the output tensor");
    int rc2 = launch_bn_stats(d.stats, B2H_F32, s);
    if (rc2) return rc2;
  }
  if (d.bwd_sums.z) return launch_bwd_sums_separate(d, B2H_F32, s);
  return B2H_OK;
}

// ---------------------------------------------------------------------------------------------
// wgrad: partial[split][t][m][n] = sum over the split's rows of P[row, m] * Q[shift_t(row), n]
// CTA tile 64 x 64, 16 rows per step, 256 threads, 4x4 micro tile; ordered split reduction after.
// ---------------------------------------------------------------------------------------------
constexpr int W_BM = 64, W_BN = 64, W_BK = 16;

template <typename T>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(b2h_wgrad_t d, int splits, int rows_per_split) {
  pdl_sync();
  __shared__ __align__(16) float As[W_BK][W_BM];
  __shared__ __align__(16) float Bs[W_BK][W_BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n_tiles = d.Npad / W_BN;
  const int m0 = (blockIdx.x / n_tiles) * W_BM, n0 = (blockIdx.x % n_tiles) * W_BN;
  const int t = blockIdx.y, split = blockIdx.z;
  const int rows = d.B * d.Lp;
  const int r_begin = split * rows_per_split;
  const int r_end = min(r_begin + rows_per_split, rows);
  const T* P = reinterpret_cast<const T*>(d.P);
  const T* Q = reinterpret_cast<const T*>(d.Q);
  const int lrow = tid >> 4, lq = (tid & 15) * 4;  // this thread loads row lrow, columns lq..lq+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int r0 = r_begin; r0 < r_end; r0 += W_BK) {
    int row = r0 + lrow;
    float4 pv = make_float4(0, 0, 0, 0), qv = pv;
    if (row < r_end) {
      int b = row / d.Lp, r = row - b * d.Lp;
      pv = load4<T>(P + (int64_t)row * d.ldp + m0 + lq);
      int rq = r * d.stride + d.tap_off[t];
      if (rq >= 0 && rq < d.Lq) qv = load4<T>(Q + ((int64_t)b * d.Lq + rq) * d.ldq + n0 + lq);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lrow][lq]) = pv;
    *reinterpret_cast<float4*>(&Bs[lrow][lq]) = qv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < W_BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[4] = {a0.x, a0.y, a0.z, a0.w};
      float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* part = d.partial + (((int64_t)split * d.ntaps + t) * d.Mpad) * d.Npad;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    *reinterpret_cast<float4*>(part + (int64_t)m * d.Npad + n0 + tx * 4) =
        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

// dW[m][n][t] = sum_split partial[split][t][m][n]   (fixed order -> deterministic).
// A warp owns 32 consecutive n of one (t, m): lane = (split group sg = lane / 8, float4 column c8 = lane % 8);
// each lane sums the splits sg, sg+4, ... with independent float4 loads, two shuffles combine the 4 groups.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(b2h_wgrad_t d, int splits) {
  pdl_sync();
  const int lane = threadIdx.x & 31, sg = lane >> 3, c8 = lane & 7;
  const int n32s = (d.Nvalid + 31) >> 5;
  const int total = d.ntaps * d.Mvalid * n32s;
  const int64_t plane = (int64_t)d.Mpad * d.Npad;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
    const int nb = w % n32s;
    int r = w / n32s;
    const int m = r % d.Mvalid, t = r / d.Mvalid;
    const int n = nb * 32 + c8 * 4;   // < Npad (a multiple of 64): the loads stay inside the plane
    const float* p = d.partial + (int64_t)t * plane + (int64_t)m * d.Npad + n;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int sp = sg; sp < splits; sp += 4) {
      const float4 v = *reinterpret_cast<const float4*>(p + (int64_t)sp * d.ntaps * plane);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
#pragma unroll
    for (int off = 8; off < 32; off <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
      acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
    }
    // lane (sg, c8) writes column n + sg
    const float v = sg == 0 ? acc.x : sg == 1 ? acc.y : sg == 2 ? acc.z : acc.w;
    if (n + sg < d.Nvalid) d.dW[((int64_t)m * d.Nvalid + n + sg) * d.ntaps + t] = v;
  }
}

static int check_wgrad(const b2h_wgrad_t& d) {
  B2H_CHECK_ARG(d.B > 0 && d.Lp > 0 && d.Lq > 0 && d.ntaps >= 1 && d.ntaps <= B2H_MAX_TAPS, B2H_ERR_SHAPE,
                "wgrad: bad shape");
  B2H_CHECK_ARG(d.Mpad % 64 == 0 && d.Npad % 64 == 0 && d.Mvalid <= d.Mpad && d.Nvalid <= d.Npad && d.Mvalid > 0 &&
                    d.Nvalid > 0,
                B2H_ERR_SHAPE, "wgrad: pads must be multiples of 64");
  B2H_CHECK_ARG(d.ldp >= d.Mpad && d.ldq >= d.Npad && d.ldp % 8 == 0 && d.ldq % 8 == 0, B2H_ERR_ALIGN,
                "wgrad: ldp=%d ldq=%d", d.ldp, d.ldq);
  B2H_CHECK_ARG(d.stride == 1 || d.stride == 2, B2H_ERR_SHAPE, "wgrad: stride must be 1 or 2");
  B2H_CHECK_ARG(d.partial && d.dW, B2H_ERR_ARG, "wgrad: null output/workspace");
  B2H_CHECK_ARG(((uintptr_t)d.partial % 16) == 0, B2H_ERR_ALIGN, "wgrad: workspace must be 16-byte aligned");
  return B2H_OK;
}

int wgrad_choose_splits(const b2h_wgrad_t& d, int dtype) {
  if (d.splits > 0) return d.splits;
  const int rows = d.B * d.Lp;
  const int tile = dtype == B2H_BF16 ? 128 : 64;
  int64_t tiles = (int64_t)ceil_div(d.Mpad, tile) * ceil_div(d.Npad, tile) * d.ntaps;
  int target = 2 * sm_count();
  int splits = (int)std::max<int64_t>(1, target / std::max<int64_t>(tiles, 1));
  int max_splits = std::max(1, rows / 256);
  if (splits > max_splits) splits = max_splits;
  if (splits > 64) splits = 64;
  return splits;
}

int launch_wgrad_reduce(const b2h_wgrad_t& d, int splits, cudaStream_t s) {
  B2H_CARVE(wgrad_reduce_kernel);
  int64_t total_warps = (int64_t)d.ntaps * d.Mvalid * ((d.Nvalid + 31) / 32);
  int blocks = (int)std::min<int64_t>(ceil_div64(total_warps, 8), (int64_t)sm_count() * 8);
  launch(wgrad_reduce_kernel, blocks, 256, 0, s, d, splits);
  B2H_LAUNCH_CHECK("wgrad_reduce");
  return B2H_OK;
}

int launch_wgrad_f32(const b2h_wgrad_t& d, cudaStream_t s) {
  B2H_CARVE(wgrad_simt_kernel<float>);
  int rc = check_wgrad(d);
  if (rc) return rc;
  int splits = wgrad_choose_splits(d, B2H_F32);
  const int rows = d.B * d.Lp;
  int rows_per_split = ceil_div(ceil_div(rows, splits), W_BK) * W_BK;
  splits = ceil_div(rows, rows_per_split);
  B2H_CHECK_ARG(d.partial_bytes <= 0 ||
                    (int64_t)splits * d.ntaps * d.Mpad * d.Npad * (int64_t)sizeof(float) <= d.partial_bytes,
                B2H_ERR_ARG, "wgrad: workspace of %lld bytes is too small for %d splits", (long long)d.partial_bytes, splits);
  dim3 grid((d.Mpad / W_BM) * (d.Npad / W_BN), d.ntaps, splits);
  launch(wgrad_simt_kernel<float>, grid, 256, 0, s, d, splits, rows_per_split);
  B2H_LAUNCH_CHECK("wgrad_f32");
  return launch_wgrad_reduce(d, splits, s);
}

int64_t wgrad_workspace_bytes(const b2h_wgrad_t& d, int dtype) {
  b2h_wgrad_t t = d;
  int splits = wgrad_choose_splits(t, dtype);
  return (int64_t)(splits + 1) * d.ntaps * d.Mpad * d.Npad * sizeof(float);
}

}  // namespace b2h
