// C ABI of libb2h.so: error state, one-shot entry points and recorded programs.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "b2h_common.cuh"
#include "tc_plans.h"

namespace b2h {

static thread_local char g_err[512] = "";
thread_local int64_t g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return B2H_ERR_CUDA;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n = v;
    else
      return 148;  // B200
  }
  return n;
}

static int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  if (major != 10) {
    set_error("libb2h is built for sm_100a only; current device has compute capability %d.x", major);
    return B2H_ERR_ARCH;
  }
  return B2H_OK;
}

union OpDesc {
  b2h_gemm_t gemm;
  b2h_wgrad_t wgrad;
  b2h_bn_stats_t bn_stats;
  b2h_bn_apply_t bn_apply;
  b2h_bn_bwd_t bn_bwd;
  b2h_prep_t prep;
  b2h_to_ncl_t to_ncl;
  b2h_l1_t l1;
  b2h_mse_t mse;
  b2h_colsum_t colsum;
  b2h_adam_t adam;
  b2h_pack_t pack;
  b2h_bn_fold_t bn_fold;
  b2h_rot6d_t rot6d;
  b2h_fill_t fill;
  b2h_pack_multi_t pack_multi;
  b2h_bn_fold_multi_t bn_fold_multi;
  b2h_fk_t fk;
  b2h_dp_adam_t dp_adam;
  OpDesc() { memset(this, 0, sizeof(*this)); }
};

static size_t desc_size(int kind) {
  switch (kind) {
    case B2H_OP_GEMM: return sizeof(b2h_gemm_t);
    case B2H_OP_WGRAD: return sizeof(b2h_wgrad_t);
    case B2H_OP_BN_STATS: return sizeof(b2h_bn_stats_t);
    case B2H_OP_BN_APPLY: return sizeof(b2h_bn_apply_t);
    case B2H_OP_BN_BWD: return sizeof(b2h_bn_bwd_t);
    case B2H_OP_PREP: return sizeof(b2h_prep_t);
    case B2H_OP_TO_NCL: return sizeof(b2h_to_ncl_t);
    case B2H_OP_L1: return sizeof(b2h_l1_t);
    case B2H_OP_MSE: return sizeof(b2h_mse_t);
    case B2H_OP_COLSUM: return sizeof(b2h_colsum_t);
    case B2H_OP_ADAM: return sizeof(b2h_adam_t);
    case B2H_OP_PACK: return sizeof(b2h_pack_t);
    case B2H_OP_BN_FOLD: return sizeof(b2h_bn_fold_t);
    case B2H_OP_ROT6D: return sizeof(b2h_rot6d_t);
    case B2H_OP_FILL: return sizeof(b2h_fill_t);
    case B2H_OP_PACK_MULTI: return sizeof(b2h_pack_multi_t);
    case B2H_OP_BN_FOLD_MULTI: return sizeof(b2h_bn_fold_multi_t);
    case B2H_OP_FK: return sizeof(b2h_fk_t);
    case B2H_OP_DP_ADAM: return sizeof(b2h_dp_adam_t);
    default: return 0;
  }
}

struct alignas(64) Op {
  TcGemmPlan gplan;   // valid for bf16 GEMM ops
  TcWgradPlan wplan;  // valid for bf16 wgrad ops
  OpDesc d;
  int kind;
};

static int run_op(const Op& op, int dtype, cudaStream_t s) {
  switch (op.kind) {
    case B2H_OP_GEMM:
      if (dtype == B2H_BF16) return run_gemm_bf16(op.gplan, op.d.gemm, s);
      return op.gplan.esz == 4 ? run_gemm_tf32(op.gplan, op.d.gemm, s) : launch_gemm_f32(op.d.gemm, s);
    case B2H_OP_WGRAD:
      if (dtype == B2H_BF16) return run_wgrad_bf16(op.wplan, op.d.wgrad, s);
      return op.wplan.esz == 4 ? run_wgrad_tf32(op.wplan, op.d.wgrad, s) : launch_wgrad_f32(op.d.wgrad, s);
    case B2H_OP_BN_STATS: return launch_bn_stats(op.d.bn_stats, dtype, s);
    case B2H_OP_BN_APPLY: return launch_bn_apply(op.d.bn_apply, dtype, s);
    case B2H_OP_BN_BWD: return launch_bn_bwd(op.d.bn_bwd, dtype, s);
    case B2H_OP_PREP: return launch_prep(op.d.prep, dtype, s);
    case B2H_OP_TO_NCL: return launch_to_ncl(op.d.to_ncl, dtype, s);
    case B2H_OP_L1: return launch_l1(op.d.l1, dtype, s);
    case B2H_OP_MSE: return launch_mse(op.d.mse, s);
    case B2H_OP_COLSUM: return launch_colsum(op.d.colsum, dtype, s);
    case B2H_OP_ADAM: return launch_adam(op.d.adam, s);
    case B2H_OP_PACK: return launch_pack(op.d.pack, dtype, s);
    case B2H_OP_BN_FOLD: return launch_bn_fold(op.d.bn_fold, s);
    case B2H_OP_ROT6D: return launch_rot6d(op.d.rot6d, s);
    case B2H_OP_FILL: return launch_fill(op.d.fill, s);
    case B2H_OP_PACK_MULTI: return launch_pack_multi(op.d.pack_multi, dtype, s);
    case B2H_OP_BN_FOLD_MULTI: return launch_bn_fold_multi(op.d.bn_fold_multi, s);
    case B2H_OP_FK: return launch_fk(op.d.fk, s);
    case B2H_OP_DP_ADAM: return launch_dp_adam(op.d.dp_adam, s);
    default: set_error("unknown op kind %d", op.kind); return B2H_ERR_ARG;
  }
}

// fp32 mode runs its contractions on the tensor cores as 3xTF32 (k_gemm_tf32.cu); B2H_FP32_SIMT=1 selects the FFMA
// kernels of k_gemm_f32.cu instead (same contracts; kept for comparison)
static bool fp32_on_tensor_cores() {
  static const bool on = getenv("B2H_FP32_SIMT") == nullptr;
  return on;
}

// tensor-core GEMM-class ops need their TMA descriptors; the FFMA ones only validation at launch
static int prepare_op(Op& op, int dtype) {
  op.gplan.esz = 0;
  op.wplan.esz = 0;
  if (dtype != B2H_BF16 && !fp32_on_tensor_cores()) return B2H_OK;
  const int esz = dtype == B2H_BF16 ? 2 : 4;
  if (op.kind == B2H_OP_GEMM) return plan_gemm_tc(op.d.gemm, &op.gplan, esz);
  if (op.kind == B2H_OP_WGRAD) return plan_wgrad_tc(op.d.wgrad, &op.wplan, esz);
  return B2H_OK;
}

static int check_dtype(int dtype) {
  B2H_CHECK_ARG(dtype == B2H_F32 || dtype == B2H_BF16, B2H_ERR_ARG, "bad dtype %d", dtype);
  return B2H_OK;
}

}  // namespace b2h

struct b2h_program {
  int dtype;
  int64_t launches;
  std::vector<b2h::Op> ops;
};

using namespace b2h;

extern "C" {

int b2h_abi_version(void) { return B2H_ABI_VERSION; }
const char* b2h_last_error(void) { return g_err; }
int b2h_check_device(void) { return check_device(); }
int b2h_sm_count(void) { return sm_count(); }
int b2h_desc_size(int kind) { return (int)desc_size(kind); }

#define B2H_ONESHOT(kind_, field_, desc_)                      \
  int rc = check_dtype(dtype);                                 \
  if (rc) return rc;                                           \
  B2H_CHECK_ARG(desc_ != nullptr, B2H_ERR_ARG, "null descriptor"); \
  Op* op = new (std::nothrow) Op();                            \
  B2H_CHECK_ARG(op != nullptr, B2H_ERR_ARG, "out of memory"); \
  op->kind = kind_;                                            \
  op->d.field_ = *desc_;                                       \
  rc = prepare_op(*op, dtype);                                 \
  if (rc == B2H_OK) rc = run_op(*op, dtype, (cudaStream_t)s);  \
  delete op;                                                   \
  return rc;

int b2h_gemm(const b2h_gemm_t* d, int dtype, b2h_stream_t s) { B2H_ONESHOT(B2H_OP_GEMM, gemm, d) }
int b2h_wgrad(const b2h_wgrad_t* d, int dtype, b2h_stream_t s) { B2H_ONESHOT(B2H_OP_WGRAD, wgrad, d) }
int b2h_bn_stats(const b2h_bn_stats_t* d, int dtype, b2h_stream_t s) { B2H_ONESHOT(B2H_OP_BN_STATS, bn_stats, d) }
int b2h_bn_apply(const b2h_bn_apply_t* d, int dtype, b2h_stream_t s) { B2H_ONESHOT(B2H_OP_BN_APPLY, bn_apply, d) }
int b2h_bn_bwd(const b2h_bn_bwd_t* d, int dtype, b2h_stream_t s) { B2H_ONESHOT(B2H_OP_BN_BWD, bn_bwd, d) }
int b2h_prep(const b2h_prep_t* d, int dtype, b2h_stream_t s) { B2H_ONESHOT(B2H_OP_PREP, prep, d) }
int b2h_to_ncl(const b2h_to_ncl_t* d, int dtype, b2h_stream_t s) { B2H_ONESHOT(B2H_OP_TO_NCL, to_ncl, d) }
int b2h_l1(const b2h_l1_t* d, int dtype, b2h_stream_t s) { B2H_ONESHOT(B2H_OP_L1, l1, d) }
int b2h_colsum(const b2h_colsum_t* d, int dtype, b2h_stream_t s) { B2H_ONESHOT(B2H_OP_COLSUM, colsum, d) }
int b2h_pack(const b2h_pack_t* d, int dtype, b2h_stream_t s) { B2H_ONESHOT(B2H_OP_PACK, pack, d) }
int b2h_pack_multi(const b2h_pack_multi_t* d, int dtype, b2h_stream_t s) {
  B2H_ONESHOT(B2H_OP_PACK_MULTI, pack_multi, d)
}
int b2h_mse(const b2h_mse_t* d, b2h_stream_t s) {
  const int dtype = B2H_F32;
  B2H_ONESHOT(B2H_OP_MSE, mse, d)
}
int b2h_adam(const b2h_adam_t* d, b2h_stream_t s) {
  const int dtype = B2H_F32;
  B2H_ONESHOT(B2H_OP_ADAM, adam, d)
}
int b2h_bn_fold(const b2h_bn_fold_t* d, b2h_stream_t s) {
  const int dtype = B2H_F32;
  B2H_ONESHOT(B2H_OP_BN_FOLD, bn_fold, d)
}
int b2h_bn_fold_multi(const b2h_bn_fold_multi_t* d, b2h_stream_t s) {
  const int dtype = B2H_F32;
  B2H_ONESHOT(B2H_OP_BN_FOLD_MULTI, bn_fold_multi, d)
}
int b2h_dp_adam(const b2h_dp_adam_t* d, b2h_stream_t s) {
  const int dtype = B2H_F32;
  B2H_ONESHOT(B2H_OP_DP_ADAM, dp_adam, d)
}
int b2h_fk(const b2h_fk_t* d, b2h_stream_t s) {
  const int dtype = B2H_F32;
  B2H_ONESHOT(B2H_OP_FK, fk, d)
}
int b2h_rot6d_to_mat(const b2h_rot6d_t* d, b2h_stream_t s) {
  const int dtype = B2H_F32;
  B2H_ONESHOT(B2H_OP_ROT6D, rot6d, d)
}
int b2h_fill(const b2h_fill_t* d, b2h_stream_t s) {
  const int dtype = B2H_F32;
  B2H_ONESHOT(B2H_OP_FILL, fill, d)
}

int64_t b2h_wgrad_workspace_bytes(const b2h_wgrad_t* d, int dtype) {
  if (!d) return B2H_ERR_ARG;
  if (dtype == B2H_BF16) return wgrad_bf16_workspace_bytes(*d);
  // fp32 mode: enough for either implementation (3xTF32 or B2H_FP32_SIMT=1)
  return std::max(wgrad_tc_workspace_bytes(*d, 4), wgrad_workspace_bytes(*d, dtype));
}
int64_t b2h_bn_partial_floats(int rows, int C, int groups) { return bn_partial_floats(rows, C, groups); }
int64_t b2h_l1_partial_floats(const b2h_l1_t* d) { return d ? l1_partial_floats(*d) : (int64_t)B2H_ERR_ARG; }

b2h_program* b2h_program_create(int dtype) {
  if (check_dtype(dtype)) return nullptr;
  b2h_program* p = new (std::nothrow) b2h_program();
  if (!p) {
    set_error("out of memory");
    return nullptr;
  }
  p->dtype = dtype;
  p->launches = 0;
  return p;
}

void b2h_program_destroy(b2h_program* p) { delete p; }

int b2h_program_add(b2h_program* p, int kind, const void* desc) {
  B2H_CHECK_ARG(p && desc, B2H_ERR_ARG, "program_add: null argument");
  size_t n = desc_size(kind);
  B2H_CHECK_ARG(n > 0, B2H_ERR_ARG, "program_add: unknown op kind %d", kind);
  p->ops.emplace_back();
  Op& op = p->ops.back();
  op.kind = kind;
  memcpy(&op.d, desc, n);
  int rc = prepare_op(op, p->dtype);
  if (rc) {
    p->ops.pop_back();
    return rc;
  }
  return (int)p->ops.size() - 1;
}

int b2h_program_size(const b2h_program* p) { return p ? (int)p->ops.size() : B2H_ERR_ARG; }

int b2h_program_run(b2h_program* p, int first, int count, b2h_stream_t s) {
  B2H_CHECK_ARG(p != nullptr, B2H_ERR_ARG, "program_run: null program");
  int n = (int)p->ops.size();
  if (count < 0) count = n - first;
  B2H_CHECK_ARG(first >= 0 && first + count <= n, B2H_ERR_ARG, "program_run: range [%d, %d) outside [0, %d)", first,
                first + count, n);
  int64_t before = g_launch_count;
  for (int i = first; i < first + count; ++i) {
    int rc = run_op(p->ops[i], p->dtype, (cudaStream_t)s);
    if (rc) return rc;
  }
  p->launches = g_launch_count - before;
  return B2H_OK;
}

int64_t b2h_program_launches(const b2h_program* p) { return p ? p->launches : 0; }

int b2h_program_op_plan(const b2h_program* p, int idx, b2h_op_plan_t* out) {
  B2H_CHECK_ARG(p && out, B2H_ERR_ARG, "program_op_plan: null argument");
  B2H_CHECK_ARG(idx >= 0 && idx < (int)p->ops.size(), B2H_ERR_ARG, "program_op_plan: op %d outside [0, %d)", idx,
                (int)p->ops.size());
  const Op& op = p->ops[idx];
  memset(out, 0, sizeof(*out));
  out->kind = op.kind;
  if (op.gplan.esz && op.kind == B2H_OP_GEMM) {
    const TcGemmPlan& g = op.gplan;
    out->tensor_core = 1;
    out->tile_n = g.BN;
    out->merged = g.p.merged;
    out->fuse_stats = g.fuse_stats;
    out->fuse_bwd = g.fuse_bwd;
    out->epilogue = g.epi;
    out->grid[0] = g.grid_x, out->grid[1] = g.grid_y, out->grid[2] = 1;
  } else if (op.wplan.esz && op.kind == B2H_OP_WGRAD) {
    const TcWgradPlan& w = op.wplan;
    out->tensor_core = 1;
    out->tile_n = w.WN;
    out->splits = w.splits;
    out->grid[0] = w.grid_x, out->grid[1] = w.p.ntaps, out->grid[2] = w.splits;
  }
  return B2H_OK;
}
int64_t b2h_launch_count(void) { return g_launch_count; }

}  // extern "C"
