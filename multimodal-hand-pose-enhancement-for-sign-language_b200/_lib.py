"""ctypes binding of libb2h.so (include/b2h_abi.h).  No fallbacks: if the library is missing or the
device is not a B200-class (sm_100) GPU, calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb2h.so")

MAX_TAPS = 8
F32, BF16 = 0, 1
ACT_NONE, ACT_LEAKY, ACT_RELU = 0, 1, 2
ROW_IDENT, ROW_UP2, ROW_POOL2, ROW_BCAST = 0, 1, 2, 3
SRC_NCL, SRC_ROWS, SRC_BCAST, SRC_MOTION = 0, 1, 2, 3
DROP_NONE, DROP_MASK, DROP_PHILOX = 0, 1, 2
# regression criterion of the generator step (b2h_l1_t.kind; --loss of train_gan.py, utils/constants.py:53-58)
LOSS_L1, LOSS_L2, LOSS_HUBER1, LOSS_ROBUST = 0, 1, 2, 3
LOSS_KINDS = {"L1": LOSS_L1, "L2": LOSS_L2, "Huber1": LOSS_HUBER1, "RobustLoss": LOSS_ROBUST}
(OP_GEMM, OP_WGRAD, OP_BN_STATS, OP_BN_APPLY, OP_BN_BWD, OP_PREP, OP_TO_NCL, OP_L1, OP_MSE, OP_COLSUM,
 OP_ADAM, OP_PACK, OP_BN_FOLD, OP_ROT6D, OP_FILL, OP_PACK_MULTI, OP_BN_FOLD_MULTI, OP_FK, OP_DP_ADAM) = range(1, 20)

i32, i64, f32, f64, vp = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p


class Dropout(C.Structure):
    _fields_ = [("mode", i32), ("site", i32), ("mask", vp), ("state", vp), ("save", vp)]


class BnStats(C.Structure):
    _fields_ = [("z", vp), ("ld", i32), ("C", i32), ("rows_per_group", i32), ("groups", i32), ("Cs", i32),
                ("mean", vp), ("invstd", vp), ("scale", vp), ("shift", vp), ("gamma", vp), ("beta", vp),
                ("running_mean", vp), ("running_var", vp), ("num_batches_tracked", vp),
                ("momentum", f32), ("eps", f32), ("partial", vp), ("ticket", vp), ("update_all_groups", i32)]


BWD_COPIES = 8


class BwdSums(C.Structure):
    _fields_ = [("z", vp), ("ld", i32), ("Lz", i32), ("rowmap", i32), ("C", i32), ("Cs", i32), ("groups", i32),
                ("mean", vp), ("invstd", vp), ("accum", vp), ("scale", vp), ("shift", vp)]


class Gemm(C.Structure):
    _fields_ = [("A", vp), ("W", vp), ("bias", vp), ("out", vp),
                ("B", i32), ("La", i32), ("Lo", i32), ("lda", i32), ("ldo", i32), ("out_coff", i32),
                ("Kc", i32), ("Npad", i32), ("Nvalid", i32), ("ntaps", i32), ("stride", i32),
                ("tap_off", i32 * MAX_TAPS), ("nphase", i32), ("Lo_actual", i32), ("act", i32),
                ("post_scale", vp), ("post_shift", vp), ("out_f32", i32), ("drop", Dropout), ("drop_C", i32),
                ("stats", BnStats), ("bwd_sums", BwdSums), ("resid", vp), ("ld_resid", i32), ("resid_up2", i32),
                ("out_pool2", i32), ("reserved1", i32), ("grad_add", vp), ("ld_grad_add", i32), ("reserved2", i32)]


class Wgrad(C.Structure):
    _fields_ = [("P", vp), ("Q", vp), ("dW", vp), ("partial", vp), ("partial_bytes", i64),
                ("B", i32), ("Lp", i32), ("Lq", i32), ("ldp", i32), ("ldq", i32),
                ("Mpad", i32), ("Npad", i32), ("Mvalid", i32), ("Nvalid", i32),
                ("ntaps", i32), ("stride", i32), ("tap_off", i32 * MAX_TAPS), ("splits", i32)]


class BnSrc(C.Structure):
    _fields_ = [("z", vp), ("ld", i32), ("coff", i32), ("rowmap", i32), ("L_src", i32), ("Cs", i32),
                ("scale", vp), ("shift", vp), ("mean", vp), ("invstd", vp)]


class BnApply(C.Structure):
    _fields_ = [("src", BnSrc * 2), ("nsrc", i32), ("out", vp), ("out_ld", i32), ("out_coff", i32),
                ("B", i32), ("L", i32), ("C", i32), ("Cfill", i32), ("groups", i32),
                ("drop", Dropout), ("drop_C", i32), ("drop_coff", i32)]


class GradSrc(C.Structure):
    _fields_ = [("g", vp), ("ld", i32), ("coff", i32), ("rowmap", i32), ("L_src", i32), ("f32", i32)]


class BnBwd(C.Structure):
    _fields_ = [("gsrc", GradSrc * 2), ("ngsrc", i32), ("bn", BnSrc), ("dpre", vp),
                ("ld_dpre", i32), ("Cfill", i32), ("B", i32), ("L", i32), ("C", i32), ("groups", i32),
                ("act", i32), ("dgamma", vp), ("dbeta", vp), ("dbias", vp), ("sums", vp), ("partial", vp),
                ("ticket", vp), ("accum", vp), ("defer", i32), ("first_pass_only", i32)]


class Prep(C.Structure):
    _fields_ = [("src", vp), ("out", vp), ("kind", i32), ("B", i32), ("L", i32), ("C", i32), ("ld", i32),
                ("Cfill", i32), ("src_ld", i32), ("drop", Dropout), ("out_f32", i32)]


class ToNcl(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("B", i32), ("L", i32), ("C", i32), ("ld", i32), ("src_f32", i32)]


class L1(C.Structure):
    _fields_ = [("out", vp), ("gt", vp), ("dout", vp), ("loss", vp), ("partial", vp), ("ticket", vp),
                ("B", i32), ("C", i32), ("L", i32), ("ld", i32), ("Cfill", i32), ("gscale", f32), ("kind", i32),
                ("dbias", vp), ("dbias_accum", vp), ("out_blc", vp), ("out_blc_ld", i32), ("reserved0", i32)]


class Mse(C.Structure):
    _fields_ = [("score", vp), ("dscore", vp), ("loss", vp), ("add", vp), ("total", vp),
                ("groups", i32), ("n", i32), ("ld", i32), ("target", f32 * 2),
                ("dpre", vp), ("dpre_ld", i32), ("dpre_bf16", i32), ("dbias", vp)]


class Colsum(C.Structure):
    _fields_ = [("src", vp), ("out", vp), ("partial", vp), ("ticket", vp),
                ("rows", i32), ("ld", i32), ("C", i32), ("f32", i32),
                ("bn_accum", vp), ("dgamma", vp), ("dbeta", vp), ("bn_groups", i32), ("reserved0", i32)]


class Adam(C.Structure):
    _fields_ = [("p", vp), ("g", vp), ("m", vp), ("v", vp), ("n", i64),
                ("lr", f64), ("beta1", f64), ("beta2", f64), ("eps", f64), ("gscale", f32), ("step", vp), ("scalars", vp), ("phase", i32)]


class Pack(C.Structure):
    _fields_ = [("W", vp), ("out", vp), ("O", i32), ("I", i32), ("Opad", i32), ("Ipad", i32),
                ("ntaps", i32), ("nphase", i32), ("o_stride", i32), ("i_stride", i32), ("k_stride", i32),
                ("tapmap", (i32 * MAX_TAPS) * 2), ("bias", vp), ("out_bias", vp)]


class PackMulti(C.Structure):
    _fields_ = [("descs", vp), ("n", i32), ("max_elems", i64)]


class BnFold(C.Structure):
    _fields_ = [("gamma", vp), ("beta", vp), ("running_mean", vp), ("running_var", vp),
                ("scale", vp), ("shift", vp), ("C", i32), ("Cpad", i32), ("eps", f32)]


class BnFoldMulti(C.Structure):
    _fields_ = [("descs", vp), ("n", i32), ("max_cpad", i32)]


class Rot6d(C.Structure):
    _fields_ = [("r6d", vp), ("mat", vp), ("n", i64)]


FK_MAX_BONES = 64


class Fk(C.Structure):
    _fields_ = [("r6d", vp), ("ld", i32), ("mean", vp), ("std", vp), ("aa", vp), ("xyz", vp), ("n", i64),
                ("nbones", i32), ("joint", C.c_int8 * FK_MAX_BONES), ("before", C.c_int8 * FK_MAX_BONES),
                ("bone_len", f32 * FK_MAX_BONES), ("root", f32 * 6)]


class Fill(C.Structure):
    _fields_ = [("ptr", vp), ("bytes", i64), ("value", i32)]


DP_MAX_PEERS, DP_MAX_BLOCKS = 16, 32


class DpAdam(C.Structure):
    """b2h_dp_adam_t: reduce-scatter(grad) + Adam + all-gather(param) over peer memory in one kernel."""
    _fields_ = [("p", vp * DP_MAX_PEERS), ("g", vp * DP_MAX_PEERS), ("signal", vp * DP_MAX_PEERS),
                ("g_mc", vp), ("p_mc", vp), ("m", vp), ("v", vp), ("n", i64), ("rank", i32), ("world", i32),
                ("beta1", f64), ("beta2", f64), ("eps", f64), ("gscale", f32), ("scalars", vp), ("timeout_ms", i32)]


class OpPlan(C.Structure):
    """b2h_op_plan_t: how the library launches one op of a program."""
    _fields_ = [("kind", i32), ("tensor_core", i32), ("tile_n", i32), ("splits", i32), ("merged", i32),
                ("fuse_stats", i32), ("fuse_bwd", i32), ("epilogue", i32), ("grid", i32 * 3), ("reserved", i32 * 5)]


OP_STRUCT = {OP_GEMM: Gemm, OP_WGRAD: Wgrad, OP_BN_STATS: BnStats, OP_BN_APPLY: BnApply, OP_BN_BWD: BnBwd,
             OP_PREP: Prep, OP_TO_NCL: ToNcl, OP_L1: L1, OP_MSE: Mse, OP_COLSUM: Colsum, OP_ADAM: Adam,
             OP_PACK: Pack, OP_BN_FOLD: BnFold, OP_ROT6D: Rot6d, OP_FILL: Fill, OP_PACK_MULTI: PackMulti,
             OP_BN_FOLD_MULTI: BnFoldMulti, OP_FK: Fk, OP_DP_ADAM: DpAdam}
KIND_OF = {v: k for k, v in OP_STRUCT.items()}

# every symbol include/b2h_abi.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "b2h_abi_version": (C.c_int, []),
    "b2h_last_error": (C.c_char_p, []),
    "b2h_check_device": (C.c_int, []),
    "b2h_sm_count": (C.c_int, []),
    "b2h_launch_count": (i64, []),
    "b2h_desc_size": (C.c_int, [C.c_int]),
    "b2h_gemm": (C.c_int, [C.POINTER(Gemm), C.c_int, vp]),
    "b2h_wgrad": (C.c_int, [C.POINTER(Wgrad), C.c_int, vp]),
    "b2h_wgrad_workspace_bytes": (i64, [C.POINTER(Wgrad), C.c_int]),
    "b2h_bn_stats": (C.c_int, [C.POINTER(BnStats), C.c_int, vp]),
    "b2h_bn_partial_floats": (i64, [C.c_int, C.c_int, C.c_int]),
    "b2h_bn_apply": (C.c_int, [C.POINTER(BnApply), C.c_int, vp]),
    "b2h_bn_bwd": (C.c_int, [C.POINTER(BnBwd), C.c_int, vp]),
    "b2h_prep": (C.c_int, [C.POINTER(Prep), C.c_int, vp]),
    "b2h_to_ncl": (C.c_int, [C.POINTER(ToNcl), C.c_int, vp]),
    "b2h_l1": (C.c_int, [C.POINTER(L1), C.c_int, vp]),
    "b2h_l1_partial_floats": (i64, [C.POINTER(L1)]),
    "b2h_mse": (C.c_int, [C.POINTER(Mse), vp]),
    "b2h_colsum": (C.c_int, [C.POINTER(Colsum), C.c_int, vp]),
    "b2h_adam": (C.c_int, [C.POINTER(Adam), vp]),
    "b2h_pack": (C.c_int, [C.POINTER(Pack), C.c_int, vp]),
    "b2h_pack_multi": (C.c_int, [C.POINTER(PackMulti), C.c_int, vp]),
    "b2h_bn_fold": (C.c_int, [C.POINTER(BnFold), vp]),
    "b2h_bn_fold_multi": (C.c_int, [C.POINTER(BnFoldMulti), vp]),
    "b2h_rot6d_to_mat": (C.c_int, [C.POINTER(Rot6d), vp]),
    "b2h_fill": (C.c_int, [C.POINTER(Fill), vp]),
    "b2h_fk": (C.c_int, [C.POINTER(Fk), vp]),
    "b2h_dp_adam": (C.c_int, [C.POINTER(DpAdam), vp]),
    "b2h_program_create": (vp, [C.c_int]),
    "b2h_program_destroy": (None, [vp]),
    "b2h_program_add": (C.c_int, [vp, C.c_int, vp]),
    "b2h_program_size": (C.c_int, [vp]),
    "b2h_program_run": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "b2h_program_launches": (i64, [vp]),
    "b2h_program_op_plan": (C.c_int, [vp, C.c_int, C.POINTER(OpPlan)]),
}

ONESHOT = {OP_GEMM: ("b2h_gemm", True), OP_WGRAD: ("b2h_wgrad", True), OP_BN_STATS: ("b2h_bn_stats", True),
           OP_BN_APPLY: ("b2h_bn_apply", True), OP_BN_BWD: ("b2h_bn_bwd", True), OP_PREP: ("b2h_prep", True),
           OP_TO_NCL: ("b2h_to_ncl", True), OP_L1: ("b2h_l1", True), OP_MSE: ("b2h_mse", False),
           OP_COLSUM: ("b2h_colsum", True), OP_ADAM: ("b2h_adam", False), OP_PACK: ("b2h_pack", True),
           OP_BN_FOLD: ("b2h_bn_fold", False), OP_ROT6D: ("b2h_rot6d_to_mat", False), OP_FILL: ("b2h_fill", False),
           OP_PACK_MULTI: ("b2h_pack_multi", True), OP_BN_FOLD_MULTI: ("b2h_bn_fold_multi", False), OP_FK: ("b2h_fk", False),
           OP_DP_ADAM: ("b2h_dp_adam", False)}


class B2HError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = False):
    """Load libb2h.so.  Raises (never falls back) if the library cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _build
            _build.build()
        else:
            raise B2HError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` "
                           "(the B200 path has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.b2h_abi_version() != 1:
        raise B2HError("libb2h.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc < 0:
        msg = load().b2h_last_error().decode(errors="replace")
        raise B2HError(f"{what} failed ({rc}): {msg}")
    return rc


def require_device():
    check(load().b2h_check_device(), "b2h_check_device")


def run_oneshot(desc, dtype: int, stream: int):
    lib = load()
    name, has_dtype = ONESHOT[KIND_OF[type(desc)]]
    fn = getattr(lib, name)
    rc = fn(C.byref(desc), dtype, stream) if has_dtype else fn(C.byref(desc), stream)
    check(rc, name)
