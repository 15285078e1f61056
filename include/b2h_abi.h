/*
 * b2h_abi.h -- C ABI of libb2h.so: the B200 (sm_100a) implementation of the Body2Hands-style
 * temporal pose regressor / discriminator GAN hot path.
 *
 * The reference (alvaro-budria/Multimodal-Hand-Pose-Enhancement-for-Sign-Language) has NO
 * FFI / plugin layer: its hot path is `torch.nn` modules (modelZoo.py) driven by train_gan.py
 * and inference.py.  This header is therefore the boundary a maintainer would bind with
 * `ctypes`; every entry point cites the reference lines it replaces.
 * File:line citations are relative to the reference repository root.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.  All pointers are DEVICE pointers
 *     unless the name says `host`.  Nothing allocates, nothing synchronises the host, every
 *     launch goes to the given stream -> every call is CUDA-graph capturable.
 *   - Return value: 0 = OK, negative = b2h_status; `b2h_last_error()` gives the message.
 *   - Activations inside the path are channels-last "BLC": [B][L][ld] with the channel dimension
 *     padded to a multiple of 64 (zero filled), in the *activation dtype* of the precision mode
 *     (B2H_F32 -> float, B2H_BF16 -> __nv_bfloat16).  Boundary tensors keep the reference layout
 *     "NCL" = (B, C, T) fp32 contiguous (modelZoo.py forward signatures).
 *   - A *program* is a recorded list of ops (built once per shape by the host-side mirror of
 *     modelZoo) that `b2h_program_run` replays on a stream with one call.
 */
#ifndef B2H_ABI_H_
#define B2H_ABI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2H_ABI_VERSION 1

typedef void* b2h_stream_t; /* cudaStream_t */

enum b2h_status {
  B2H_OK = 0,
  B2H_ERR_SHAPE = -1,
  B2H_ERR_ALIGN = -2,
  B2H_ERR_ARCH = -3,
  B2H_ERR_CUDA = -4,
  B2H_ERR_ARG = -5
};

enum b2h_dtype { B2H_F32 = 0, B2H_BF16 = 1 };
enum b2h_act { B2H_ACT_NONE = 0, B2H_ACT_LEAKY = 1 /* slope 0.2 */, B2H_ACT_RELU = 2 };
enum b2h_rowmap {
  B2H_ROW_IDENT = 0, /* src row (b, l)                                             */
  B2H_ROW_UP2 = 1,   /* fwd: src row (b, l/2)  [repeat_interleave(2)[:L], modelZoo.py:295-296]
                        bwd: dy(b, j) = g(b, 2j) + g(b, 2j+1)                      */
  B2H_ROW_POOL2 = 2, /* fwd: max over src rows (b, 2l), (b, 2l+1) [MaxPool1d(2,2), modelZoo.py:197]
                        bwd: dy(b, t) = g(b, t/2) if t is the (first) argmax       */
  B2H_ROW_BCAST = 3  /* fwd: src row (b) for every l (eval-mode text rows are identical) */
};
enum b2h_prep_src {
  B2H_SRC_NCL = 0,    /* fp32 (B, C, L): modelZoo input_ / loss gradients                     */
  B2H_SRC_ROWS = 1,   /* fp32 (B*L, C) row major: image feats (B, T, 2000)                    */
  B2H_SRC_BCAST = 2,  /* fp32 (B, C) repeated L times: text feats, process_text modelZoo.py:284-287 */
  B2H_SRC_MOTION = 3  /* fp32 NCL (B, C, L+1): out[b,l,c] = x[b,c,0] - x[b,c,l]  (calc_motion, train_gan.py:209-211) */
};
enum b2h_dropout_mode { B2H_DROP_NONE = 0, B2H_DROP_MASK = 1, B2H_DROP_PHILOX = 2 };

/* nn.Dropout(0.5) (modelZoo.py:193 and every block): y = x * keep * 2.
 * MASK: explicit keep-mask (parity tests).  PHILOX: Philox4x32-10 keyed by (seed, step, site, element). */
typedef struct {
  int32_t mode;
  int32_t site;
  const uint8_t* mask;   /* MASK: [rows][C] keep flags, C = drop_C of the op */
  const uint64_t* state; /* PHILOX: device {seed, step} */
  uint8_t* save;         /* forward ops, optional: the keep flags this op applied are also written here
                            ([rows][C]) so that the backward replays them in MASK mode */
} b2h_dropout_t;

/* train-mode BatchNorm1d statistics (declared here because a GEMM can produce them in its epilogue) */
/* batch statistics of z over (rows of one group): mean, biased var -> invstd, and the folded affine
 * scale = gamma*invstd, shift = beta - mean*scale; running update
 * running = (1-m)*running + m*batch (unbiased var), num_batches_tracked += 1. */
typedef struct {
  const void* z;
  int32_t ld, C, rows_per_group, groups;
  int32_t Cs;    /* stride of the [groups][Cs] outputs */
  float* mean;
  float* invstd;
  float* scale;
  float* shift;
  const float* gamma; /* [C] */
  const float* beta;  /* [C] */
  float* running_mean; /* may be NULL (no update) */
  float* running_var;
  int64_t* num_batches_tracked;
  float momentum, eps;
  float* partial;        /* workspace >= b2h_bn_partial_floats() */
  uint32_t* ticket;      /* workspace, zero-initialised, self-resetting */
  int32_t update_all_groups; /* 1: apply the running update once per group in order (two D forwards) */
} b2h_bn_stats_t;

/* First pass of the BatchNorm backward of the layer whose BN output a dgrad GEMM differentiates, produced by
 * that GEMM (tensor-core path: in its epilogue, from the tile it stores and a TMA-staged tile of z):
 *   accum[copy][group][c] += ( sum_rows out[row, c],  sum_rows out[row, c] * (z[map(row), c] - mean[c]) * invstd[c] )
 * over the rows this GEMM writes; several GEMMs (the consumers of one layer) may add into the same accumulators.
 * b2h_bn_bwd with the same `accum` then runs its second pass only. */
#define B2H_BWD_COPIES 8
typedef struct {
  const void* z;      /* producer's post-activation tensor [B][Lz][ld], act dtype; NULL = disabled */
  int32_t ld, Lz;     /* row pitch / rows per sample of z */
  int32_t rowmap;     /* B2H_ROW_IDENT: z row (b, l) (Lz == Lo_actual);  B2H_ROW_UP2: z row (b, l/2) (Lo_actual == 2*Lz);
                         B2H_ROW_POOL2 (Lz == 2*Lo_actual): the GEMM differentiates MaxPool1d(2) of the producer's BN
                         output -- gradient row (b, l) belongs to z row (b, 2l) or (b, 2l+1), whichever has the larger
                         z*scale + shift (the first on ties, as nn.MaxPool1d), and to no other row */
  int32_t C, Cs, groups;
  const float* mean;   /* [groups][Cs] batch statistics of the producer's forward */
  const float* invstd;
  double* accum;      /* [B2H_BWD_COPIES][groups][C][2], zero before the first contribution */
  const float* scale; /* [groups][Cs] forward affine of the producer (B2H_ROW_POOL2 only) */
  const float* shift;
} b2h_bwd_sums_t;

/* ------------------------------------------------------------------------------------------- */
/* tap-GEMM: every contraction of the path (Conv1d, its dgrad, ConvTranspose1d as a 2-phase     */
/* sub-pixel conv, its dgrad as a strided conv, Linear) is                                       */
/*   out[b, lo, n] = epi( sum_t sum_c A[b, lo*stride + tap_off[t], c] * W[n, t, c] )            */
/* with zero rows outside [0, La).  Replaces nn.Conv1d / nn.ConvTranspose1d / nn.Linear call    */
/* sites modelZoo.py:19-118,182-281,768-813 and their autograd backward.                         */
/* ------------------------------------------------------------------------------------------- */
#define B2H_MAX_TAPS 8
typedef struct {
  const void* A;     /* [B][La][lda], act dtype                                                */
  const void* W;     /* packed [Npad][ntaps][Kc], act dtype (b2h_pack_weight)                  */
  const float* bias; /* [Npad/nphase] packed fp32 (indexed by the channel within a phase) or NULL */
  void* out;         /* [B][Lo_actual][ldo] (+ out_coff); act dtype, or fp32 if out_f32        */
  int32_t B, La, Lo, lda, ldo, out_coff;
  int32_t Kc;     /* channels per tap, multiple of 64 (zero padded)                            */
  int32_t Npad;   /* multiple of 64                                                            */
  int32_t Nvalid; /* columns >= Nvalid (per phase) are not written                            */
  int32_t ntaps, stride;
  int32_t tap_off[B2H_MAX_TAPS];
  int32_t nphase;    /* 1, or 2: columns [p*Npad/2, (p+1)*Npad/2) are output row lo*2+p (sub-pixel) */
  int32_t Lo_actual; /* rows per sample of the real output tensor (= Lo*nphase unless ragged)  */
  int32_t act;       /* b2h_act, applied after bias                                            */
  const float* post_scale; /* eval-mode BatchNorm folded to y = v*scale + shift after act, or NULL */
  const float* post_shift;
  int32_t out_f32;   /* 0: `out` is BLC in the activation dtype; 1: BLC fp32; 2 (bf16 mode, output layer of a batched
                        inference): `out` is the reference's own (B, Nvalid, Lo_actual) fp32 NCL tensor and the op
                        writes it itself -- needs nphase = 1, stride = 1, bias only (no activation / BN / dropout /
                        statistics), Npad % 256 == 0, Lo_actual == Lo, Lo % 4 == 0; ldo / out_coff are ignored */
  b2h_dropout_t drop; /* dgrad: multiply by keep*2 of the dropout site that produced A_prev    */
  int32_t drop_C;     /* valid channel count of that site (mask row length)                    */
  b2h_bn_stats_t stats; /* optional (stats.z != NULL, must equal `out`): also compute the train-mode BatchNorm
                           statistics of the output, exactly as b2h_bn_stats(&stats) right after this op would
                           (on the tensor-core path inside the GEMM epilogue, without re-reading `out`) */
  b2h_bwd_sums_t bwd_sums; /* optional (bwd_sums.z != NULL), dgrad ops: see b2h_bwd_sums_t */
  /* Residual add in the epilogue (bf16 mode, eval plans of a batched inference; runs the persistent 256-column kernel:
   * nphase = 1, Npad % 256 == 0, no statistics): out[row] = epi(...)[row] + resid[row], both [B][Lo_actual][ld] in the
   * activation dtype.  resid_up2 = 1: this layer's result is up-sampled x2 (nearest) on the way -- GEMM row l produces
   * the output rows 2l and 2l+1, out[2l+k] = epi(...)[l] + resid[2l+k]; out and resid have 2*Lo_actual rows per sample.
   * (modelZoo.py:262-270: the skip connections of the decoder) */
  const void* resid;
  int32_t ld_resid;
  int32_t resid_up2;
  /* MaxPool1d(2) in the epilogue (same kernel, same restrictions; modelZoo.py:197 between the encoder and conv5):
   * out[l'] = max(epi(...)[2l'], epi(...)[2l'+1]) for l' < Lo_actual / 2 -- `out` has Lo_actual / 2 rows per sample. */
  int32_t out_pool2;
  int32_t reserved1;
  /* dgrad ops of the bf16 tap-GEMM (one tile per CTA): out[row] = epi(...)[row] + grad_add[row] -- the gradient that
   * ANOTHER consumer of the same tensor has already written (a skip connection: conv5's output feeds conv6 and, added
   * to skip4's, skip5: modelZoo.py:262-270), [B][Lo_actual][ld_grad_add] in the activation dtype, added in fp32 before
   * the one rounding of the stored value.  The producer's BatchNorm backward then reads ONE gradient source, and the
   * sums of b2h_bwd_sums_t (taken from the stored values) cover both consumers. */
  const void* grad_add;
  int32_t ld_grad_add;
  int32_t reserved2;
} b2h_gemm_t;

/* wgrad: dW[m][n][t] = sum_{b,r} P[b, r, m] * Q[b, r*stride + tap_off[t], n]   (PyTorch weight layout)
 * conv: P = dpre, Q = a (dW[n][c][k]);  convT: P = a, Q = dpre (dW[c][n][k]); Linear: ntaps = 1. */
typedef struct {
  const void* P; /* [B][Lp][ldp] act dtype */
  const void* Q; /* [B][Lq][ldq] act dtype */
  float* dW;     /* [Mvalid][Nvalid][ntaps] fp32 */
  float* partial; /* workspace: >= b2h_wgrad_workspace_bytes() */
  int64_t partial_bytes; /* size of `partial`; checked against the plan when > 0 (0 = caller vouches for it) */
  int32_t B, Lp, Lq, ldp, ldq;
  int32_t Mpad, Npad, Mvalid, Nvalid; /* pads are multiples of 64 */
  int32_t ntaps, stride;
  int32_t tap_off[B2H_MAX_TAPS];
  int32_t splits; /* split-K slices over the rows: 0 = let the library choose (fills the GPU; fp32 partial planes in
                     `partial` + an ordered reduce launch);  1 on the tensor-core path = split-free: one CTA per
                     output tile and tap walks all rows and writes dW itself (no workspace traffic, no reduce
                     launch, a fraction of the SM time, a longer launch — for callers that run it beside other work) */
} b2h_wgrad_t;

/* ------------------------------------------------------------------------------------------- */
/* BatchNorm1d pieces (modelZoo.py:196 etc.; block order conv -> LeakyReLU -> BN)    */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* z;  /* post-activation, pre-BN tensor [rows][ld], act dtype */
  int32_t ld, coff, rowmap, L_src;
  int32_t Cs;           /* row stride of the per-channel arrays below = channels of the source layer padded
                           to a multiple of 64; entries beyond the layer's channels are 0 */
  const float* scale;   /* [groups][Cs]  gamma * invstd              y = z * scale + shift          */
  const float* shift;   /* [groups][Cs]  beta - mean * scale                                         */
  const float* mean;    /* [groups][Cs]  batch mean          (backward only)                        */
  const float* invstd;  /* [groups][Cs]  1/sqrt(var_b + eps) (backward only)                        */
} b2h_bn_src_t;

/* out[b,l,coff+c] = dropout( BN0(src0) [+ BN1(src1)] ), zero fill up to Cfill */
typedef struct {
  b2h_bn_src_t src[2];
  int32_t nsrc;
  void* out;
  int32_t out_ld, out_coff;
  int32_t B, L, C, Cfill, groups;
  b2h_dropout_t drop;
  int32_t drop_C, drop_coff;
} b2h_bn_apply_t;

typedef struct {
  const void* g; /* gradient w.r.t. the consumer's pre-dropout input (mask already applied) */
  int32_t ld, coff, rowmap, L_src, f32;
} b2h_grad_src_t;

/* BN + activation backward. dy = sum of grad sources;
 *   dz = gamma*invstd*(dy - mean(dy) - zhat*mean(dy*zhat));  dpre = dz * act'(z)
 * also dgamma = sum(dy*zhat), dbeta = sum(dy), dbias = sum(dpre). */
typedef struct {
  b2h_grad_src_t gsrc[2];
  int32_t ngsrc;
  b2h_bn_src_t bn; /* z, mean, invstd, scale, shift of THIS layer (rowmap IDENT, coff 0) */
  /* for POOL2 grad sources the pooled tensor was max over BN(z) pairs of this layer */
  void* dpre;      /* [rows][ld_dpre] act dtype, zero filled up to Cfill */
  int32_t ld_dpre, Cfill;
  int32_t B, L, C, groups, act;
  float* dgamma; /* [C] (summed over groups) */
  float* dbeta;
  float* dbias;
  float* sums;    /* workspace [groups][C][2]  (sum dy, sum dy*zhat) */
  float* partial; /* workspace [nchunks][groups][C][2] (shared by both passes) */
  uint32_t* ticket;
  double* accum;  /* optional: the first-pass sums were already accumulated here by the GEMMs that produced the
                     gradient sources (b2h_gemm_t.bwd_sums); the op then runs its second pass only and re-zeroes it */
  int32_t defer;  /* 1 (needs `accum` and `dpre`): the op only writes dpre -- no reduction, no ticket, no serial tail on
                     the dependency chain of the backward pass.  dgamma / dbeta / dbias and the re-zeroing of `accum`
                     are left to a b2h_colsum over dpre with `bn_accum` = this accum, which only feeds the optimizer
                     and can run beside the chain.  dgamma / dbeta / dbias / sums / partial / ticket are ignored.
                     2: as 1, but the op still accumulates the sums of dpre (the bias gradient) in ITS OWN `partial`
                     (b2h_bn_partial_floats, not shared with another op in flight, zero before) and only skips the
                     ticket and the last-CTA pass; the finishing b2h_colsum then has src = NULL (nothing to re-read)
                     and the same `partial` */
  int32_t first_pass_only; /* 1 (needs `accum`, dpre == NULL): accumulate sum(dy), sum(dy*zhat) of the given gradient
                     sources into `accum` and stop -- the contribution of a consumer whose dgrad GEMM cannot carry
                     b2h_gemm_t.bwd_sums for this layer (it serves another producer), run beside the chain as soon as
                     that gradient exists */
} b2h_bn_bwd_t;

/* ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* src;
  void* out; /* [B][L][ld] act dtype */
  int32_t kind; /* b2h_prep_src */
  int32_t B, L, C, ld, Cfill;
  int32_t src_ld; /* ROWS/BCAST: row pitch of src in floats */
  b2h_dropout_t drop;
  int32_t out_f32;
} b2h_prep_t;

typedef struct {
  const void* src; /* [B][L][ld] */
  float* dst;      /* (B, C, L) fp32 */
  int32_t B, L, C, ld, src_f32;
} b2h_to_ncl_t;

/* Regression criterion of the generator step, forward + backward in one pass (LOSSES, utils/constants.py:53-58;
 * train_gan.py:286-292).  With d = out - gt, n = numel, g = gscale / n:
 *   B2H_LOSS_L1     nn.L1Loss():             loss = mean|d|            dout = sign(d) * g
 *   B2H_LOSS_L2     nn.MSELoss():            loss = mean d^2           dout = 2 d * g
 *   B2H_LOSS_HUBER1 nn.HuberLoss(delta=1.0): loss = mean(|d| < 1 ? d^2/2 : |d| - 1/2)
 *                                            dout = clamp(d, -1, 1) * g
 *   B2H_LOSS_ROBUST robust_loss AdaptiveLossFunction as train_gan.py:74-77,286-290 uses it: its latent alpha / scale
 *                   are never handed to the optimiser (train_gan.py:69), so they stay at their initial values
 *                   alpha = 2, scale = 1/2 (utils/robust_loss/adaptive.py:55-59) and the negative log-likelihood is
 *                   mean(2 d^2) + log(1/2) + log sqrt(2 pi);  dout = 4 d * g
 * dout is written BLC in the activation dtype. */
enum { B2H_LOSS_L1 = 0, B2H_LOSS_L2 = 1, B2H_LOSS_HUBER1 = 2, B2H_LOSS_ROBUST = 3 };
typedef struct {
  const float* out; /* NCL fp32 */
  const float* gt;  /* NCL fp32 */
  void* dout;       /* [B][L][ld] act dtype (may be NULL: forward only) */
  float* loss;      /* loss[0] */
  float* partial;
  uint32_t* ticket;
  int32_t B, C, L, ld, Cfill;
  float gscale;
  int32_t kind;        /* B2H_LOSS_* (0 = L1) */
  float* dbias;        /* optional: column sums of dout as stored = bias gradient of the output layer [C] */
  double* dbias_accum; /* workspace for dbias: [16][C] doubles, zero-initialised, self-resetting */
  const float* out_blc; /* optional: the prediction as the output layer's GEMM left it, fp32 [B][L][out_blc_ld].  When
                           given, the op reads it from there and WRITES `out` (the NCL fp32 tensor the reference returns)
                           in the same pass -- it replaces b2h_to_ncl + the re-read of its result */
  int32_t out_blc_ld;
  int32_t reserved0;
} b2h_l1_t;

/* nn.MSELoss(score, target) (train_gan.py:93,247,292) on (groups, n) scores, one target per group;
 * loss[0] = sum_g mean((s_g - t_g)^2); dscore = 2*(s - t)/n (NULL: forward only);
 * optionally total[0] = loss[0] + add[0].
 * Optionally the gradient goes straight into the backward of the (1-channel) score layer: dpre[(g*n + i)*dpre_ld]
 * = dscore in the activation dtype (column 0 of the layer's dpre rows; the padding columns stay as they are) and
 * dbias[0] = the sum of those stored values (the layer's bias gradient). */
typedef struct {
  const float* score; /* element (g, i) at score[(g*n + i)*ld] */
  float* dscore;      /* same addressing; may be NULL */
  float* loss;
  const float* add;
  float* total;
  int32_t groups, n, ld;
  float target[2];
  void* dpre;         /* optional */
  int32_t dpre_ld, dpre_bf16;
  float* dbias;       /* optional, needs dpre */
} b2h_mse_t;

typedef struct {
  const void* src; /* [rows][ld] act dtype */
  float* out;      /* [C] */
  float* partial;
  uint32_t* ticket;
  int32_t rows, ld, C, f32;
  /* optional: finish a deferred BatchNorm backward (b2h_bn_bwd_t.defer) -- src = its dpre, out = the bias gradient;
   * the last CTA also sums the first-pass accumulators [B2H_BWD_COPIES][bn_groups][C][2] in fixed order,
   * writes dbeta[c] = sum dy, dgamma[c] = sum dy*zhat (over the groups) and re-zeroes them.
   * src = NULL (defer = 2): one CTA; the sums of dpre are taken from the bn_bwd's own `partial` (same pointer here) */
  double* bn_accum;
  float* dgamma;
  float* dbeta;
  int32_t bn_groups, reserved0;
} b2h_colsum_t;

/* torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay=0) (train_gan.py:69,88) over a
 * flat buffer; step count lives on the device (graph replay); gscale folds the 1/world of DDP. */
typedef struct {
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t n;
  double lr, beta1, beta2, eps; /* Python-float hyper-parameters, as torch.optim.Adam holds them */
  float gscale;
  int64_t* step;  /* device; incremented by this op BEFORE use (t = ++step) */
  float* scalars; /* device workspace, 2 floats: the step's bias-correction scalars */
  int32_t phase;  /* 0: whole update (advance the step, then update n elements);  1: only advance the step and
                     refresh `scalars` (n ignored);  2: only update n elements with the current `scalars` — one
                     phase-1 call followed by phase-2 calls over disjoint ranges (gradient buckets) is phase 0 */
} b2h_adam_t;

/* weight repack: out[(ph*Opad + o)][t][i] = W[o*o_stride + i*i_stride + tapmap[ph][t]*k_stride]
 * (0 when tapmap < 0 or o >= O or i >= I); optional bias repack out_bias[o] = bias[o], o < Opad. */
typedef struct {
  const float* W;
  void* out; /* act dtype */
  int32_t O, I, Opad, Ipad, ntaps, nphase;
  int32_t o_stride, i_stride, k_stride;
  int32_t tapmap[2][B2H_MAX_TAPS];
  const float* bias;
  float* out_bias;
} b2h_pack_t;

/* many weight repacks in ONE launch: `descs` is a DEVICE array of n descriptors (the per-step repack of a
 * whole network after the optimizer step) */
typedef struct {
  const b2h_pack_t* descs;
  int32_t n;
  int64_t max_elems; /* largest nphase*Opad*ntaps*Ipad among the descriptors */
} b2h_pack_multi_t;

/* eval-mode BN folded to scale/shift: scale = gamma/sqrt(rv+eps), shift = beta - rm*scale, padded */
typedef struct {
  const float *gamma, *beta, *running_mean, *running_var;
  float* scale;
  float* shift;
  int32_t C, Cpad;
  float eps;
} b2h_bn_fold_t;

/* many BN folds in ONE launch: `descs` is a DEVICE array of n descriptors (all BN layers of an eval network) */
typedef struct {
  const b2h_bn_fold_t* descs;
  int32_t n;
  int32_t max_cpad; /* largest Cpad among the descriptors */
} b2h_bn_fold_multi_t;

/* 6D rotation -> 3x3 matrix, row-major 9 floats per joint (utils/conversion_utils.py:86-107) */
typedef struct {
  const float* r6d; /* [n][6] */
  float* mat;       /* [n][9] */
  int64_t n;
} b2h_rot6d_t;

/* inference post-processing behind save_results (utils/utils.py:388-427): per frame, the 6-D rotations of the
 * nbones-1 joints -> rotation matrix (conversion_utils.py:86-107) -> axis-angle (conversion_utils.py:33-41)
 * -> forward kinematics over the bone tree (conversion_utils.py:117-137): bone i starts at joint `joint[i]`,
 * ends at joint i+1, and is the Rodrigues rotation, by the i-th axis-angle, of the unit direction from joint
 * `before[i]` to joint `joint[i]`, scaled by bone_len[i].  Bone 0 is the root bone (joints 0 and 1 = `root`). */
#define B2H_FK_MAX_BONES 64
typedef struct {
  const float* r6d;  /* [n][ld] fp32, (nbones-1)*6 values per frame */
  int32_t ld;
  const float* mean; /* optional per-channel de-standardisation r6d*std + mean ((nbones-1)*6 values), or NULL */
  const float* std;
  float* aa;         /* optional out [n][(nbones-1)*3] axis-angles, or NULL */
  float* xyz;        /* out [n][(nbones+1)*3] */
  int64_t n;
  int32_t nbones;    /* <= B2H_FK_MAX_BONES */
  int8_t joint[B2H_FK_MAX_BONES];
  int8_t before[B2H_FK_MAX_BONES];
  float bone_len[B2H_FK_MAX_BONES];
  float root[6];
} b2h_fk_t;

typedef struct {
  void* ptr;
  int64_t bytes;
  int32_t value; /* byte value */
} b2h_fill_t;

/* ------------------------------------------------------------------------------------------- */
/* library                                                                                      */
/* ------------------------------------------------------------------------------------------- */
int b2h_abi_version(void);
const char* b2h_last_error(void);
/* 0 if a CUDA device of compute capability 10.x is current, else B2H_ERR_ARCH / B2H_ERR_CUDA */
int b2h_check_device(void);
int b2h_sm_count(void);
/* kernels launched by the calling thread through this library so far (one-shots and program runs) */
int64_t b2h_launch_count(void);
/* sizeof() of the descriptor struct of an op kind (b2h_op_kind), for binding self-checks */
int b2h_desc_size(int kind);

/* one-shot launches (unit tests, eager use). `dtype` is the activation dtype (b2h_dtype). */
int b2h_gemm(const b2h_gemm_t* d, int dtype, b2h_stream_t s);
int b2h_wgrad(const b2h_wgrad_t* d, int dtype, b2h_stream_t s);
int64_t b2h_wgrad_workspace_bytes(const b2h_wgrad_t* d, int dtype);
int b2h_bn_stats(const b2h_bn_stats_t* d, int dtype, b2h_stream_t s);
int64_t b2h_bn_partial_floats(int rows, int C, int groups); /* size of `partial` for stats / bwd */
int b2h_bn_apply(const b2h_bn_apply_t* d, int dtype, b2h_stream_t s);
int b2h_bn_bwd(const b2h_bn_bwd_t* d, int dtype, b2h_stream_t s);
int b2h_prep(const b2h_prep_t* d, int dtype, b2h_stream_t s);
int b2h_to_ncl(const b2h_to_ncl_t* d, int dtype, b2h_stream_t s);
int b2h_l1(const b2h_l1_t* d, int dtype, b2h_stream_t s);
int64_t b2h_l1_partial_floats(const b2h_l1_t* d);
int b2h_mse(const b2h_mse_t* d, b2h_stream_t s);
int b2h_colsum(const b2h_colsum_t* d, int dtype, b2h_stream_t s);
int b2h_adam(const b2h_adam_t* d, b2h_stream_t s);
int b2h_pack(const b2h_pack_t* d, int dtype, b2h_stream_t s);
int b2h_pack_multi(const b2h_pack_multi_t* d, int dtype, b2h_stream_t s);
int b2h_bn_fold(const b2h_bn_fold_t* d, b2h_stream_t s);
int b2h_bn_fold_multi(const b2h_bn_fold_multi_t* d, b2h_stream_t s);
int b2h_rot6d_to_mat(const b2h_rot6d_t* d, b2h_stream_t s);
int b2h_fill(const b2h_fill_t* d, b2h_stream_t s);
int b2h_fk(const b2h_fk_t* d, b2h_stream_t s);

/* Data-parallel optimiser step as ONE kernel over NVLink peer memory (the gradient all-reduce of
 * `train_gan.py`'s step under one process per GPU): reduce-scatter of the flat gradients, Adam on the owned slice,
 * all-gather of the updated parameters — replaces ncclAllReduce(grad) + b2h_adam(phase 2) of a bucket.
 *   entry barrier   every rank's gradients of the range are complete, and nobody still reads the old parameters
 *   rank r, slice r g = sum_q g[q][i] in rank order (peer loads; or one multimem.ld_reduce through the NVSwitch
 *                   when g_mc is given), Adam with the current `scalars` (b2h_adam phase 1 ran before) on the local
 *                   moments m / v, which only hold meaningful values for the owned slice, new parameter stored to
 *                   every rank's buffer (peer stores, or one multimem.st when p_mc is given)
 *   exit barrier    every rank's parameters are complete and its gradients may be overwritten
 * Every rank must call it with the same n / world and its own rank, in the same order on every rank; all pointers
 * are peer-mapped device addresses of symmetric allocations (cudaIpc / cuMem fabric handles or
 * torch.distributed._symmetric_memory).  `signal[q]` is rank q's signal pad for THIS call site:
 * B2H_DP_MAX_BLOCKS * world uint32, zero-initialised once, self-resetting.  A peer that never arrives makes the
 * kernel trap after `timeout_ms` (0: 10 s) instead of hanging the device.  CUDA-graph capturable. */
#define B2H_DP_MAX_PEERS 16
#define B2H_DP_MAX_BLOCKS 32
typedef struct {
  float* p[B2H_DP_MAX_PEERS];          /* every rank's parameter range [n] (own rank included) */
  const float* g[B2H_DP_MAX_PEERS];    /* every rank's gradient range [n] */
  uint32_t* signal[B2H_DP_MAX_PEERS];  /* every rank's signal pad */
  const float* g_mc;                   /* optional multicast address of the gradient ranges (NVLS), else NULL */
  float* p_mc;                         /* optional multicast address of the parameter ranges, else NULL */
  float* m;                            /* local Adam moments [n] */
  float* v;
  int64_t n;                           /* multiple of 4; ranges 16-byte aligned */
  int32_t rank, world;
  double beta1, beta2, eps;
  float gscale;                        /* 1 / world for the mean gradient of DDP */
  const float* scalars;                /* the step's bias-correction scalars written by b2h_adam phase 1 */
  int32_t timeout_ms;
} b2h_dp_adam_t;
int b2h_dp_adam(const b2h_dp_adam_t* d, b2h_stream_t s);

/* recorded programs: the whole-graph entry points (generator forward / step, discriminator step)
 * are programs built by the host-side mirror of modelZoo and replayed with ONE call. */
typedef struct b2h_program b2h_program;
enum b2h_op_kind {
  B2H_OP_GEMM = 1, B2H_OP_WGRAD, B2H_OP_BN_STATS, B2H_OP_BN_APPLY, B2H_OP_BN_BWD, B2H_OP_PREP,
  B2H_OP_TO_NCL, B2H_OP_L1, B2H_OP_MSE, B2H_OP_COLSUM, B2H_OP_ADAM, B2H_OP_PACK, B2H_OP_BN_FOLD,
  B2H_OP_ROT6D, B2H_OP_FILL, B2H_OP_PACK_MULTI, B2H_OP_BN_FOLD_MULTI, B2H_OP_FK, B2H_OP_DP_ADAM
};
b2h_program* b2h_program_create(int dtype);
void b2h_program_destroy(b2h_program* p);
/* appends a copy of the descriptor (`desc` points at the struct matching `kind`); returns the op
 * index or a negative b2h_status */
int b2h_program_add(b2h_program* p, int kind, const void* desc);
int b2h_program_size(const b2h_program* p);
/* replay ops [first, first+count) on the stream; count < 0 = to the end */
int b2h_program_run(b2h_program* p, int first, int count, b2h_stream_t s);
/* number of kernel launches the last run issued (bench.py's gpu_launches) */
int64_t b2h_program_launches(const b2h_program* p);
/* How the library launches op `idx` of a program: tile shape, split-K and epilogue fusions of the tensor-core
 * kernels (zeros for the other op kinds).  The parity tests use it to assert that every kernel template instance
 * of the benchmarked shapes is exercised against the oracle. */
typedef struct {
  int32_t kind;        /* b2h_op_kind */
  int32_t tensor_core; /* 1: the op runs a tcgen05 kernel */
  int32_t tile_n;      /* gemm: BN;  wgrad: WN */
  int32_t splits;      /* wgrad: split-K slices (1 = split-free, dW written by the GEMM itself) */
  int32_t merged;      /* gemm: tap-merged main loop */
  int32_t fuse_stats;  /* gemm: train-mode BatchNorm statistics produced by the epilogue */
  int32_t fuse_bwd;    /* gemm: first pass of the producer's BatchNorm backward produced by the epilogue */
  int32_t epilogue;    /* gemm: id of the specialised epilogue (0 = generic) */
  int32_t grid[3];
  int32_t reserved[5];
} b2h_op_plan_t;
int b2h_program_op_plan(const b2h_program* p, int idx, b2h_op_plan_t* out);

#ifdef __cplusplus
}
#endif
#endif /* B2H_ABI_H_ */
