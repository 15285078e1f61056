"""Every record of every program of a trainer lowers to its C descriptor (include/b2h_abi.h through the ctypes mirrors
of b2h_b200/_lib.py) on CPU: unknown field names, wrong value types and table overflows are caught before a GPU is
involved.  (Launch-time validation — shapes, alignment, tensor maps — lives in the library and needs the device.)"""
import ctypes as C

import pytest
import torch

import b2h_b200  # noqa: F401
from b2h_b200 import _lib as L
from b2h_b200.program import lower
from b2h_b200.trainer import GanTrainer


def programs_of(tr):
    progs = {"g_loss": tr.g_loss_prog, "d_loss": tr.d_loss_prog, "g_buckets": tr._buckets["g"][1],
             "d_buckets": tr._buckets["d"][1]}
    for name in ("G_train", "G_eval", "D_train", "D_eval"):
        progs[name] = getattr(tr, name).prog
    return progs


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("variant,rf,cin,cout", [("v1", False, 36, 252), ("v1", True, 36, 252), ("b2h", True, 36, 252),
                                                 ("v2", True, 264, 24), ("v4", True, 36, 252),
                                                 ("v4_deeper", True, 90, 198)])
def test_every_record_lowers(variant, rf, cin, cout, precision):
    tr = GanTrainer(variant, cin, cout, rf, 4, 20, precision=precision, device="cpu", drop_mode="philox", loss="Huber1")
    lib = L.load()
    kinds = set()
    n = 0
    for name, prog in programs_of(tr).items():
        covered = sorted(prog.segments.values())
        assert covered and all(a[1] <= b[0] for a, b in zip(covered, covered[1:])), name   # segments do not overlap
        assert sum(b - a for a, b in covered) == len(prog.recs), name                       # and no op is outside one
        for rec in prog.recs:
            st = lower(rec)
            assert C.sizeof(st) == lib.b2h_desc_size(rec.kind), rec
            kinds.add(rec.kind)
            n += 1
            for fname, ctype in st._fields_:
                v = rec.f.get(fname)
                if ctype is L.vp and isinstance(v, torch.Tensor):
                    assert getattr(st, fname) == v.data_ptr(), (rec, fname)
                    assert v.data_ptr() % 4 == 0
    assert n > 100
    assert {L.OP_GEMM, L.OP_WGRAD, L.OP_BN_APPLY, L.OP_BN_BWD, L.OP_PREP, L.OP_TO_NCL, L.OP_L1, L.OP_MSE, L.OP_ADAM,
            L.OP_PACK_MULTI} <= kinds
    # the regression criterion reaches the descriptor
    l1 = [rec for rec in tr.g_loss_prog.recs if rec.kind == L.OP_L1]
    assert len(l1) == 1 and lower(l1[0]).kind == L.LOSS_HUBER1
