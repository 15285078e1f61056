"""Op records and recorded programs.

A `Program` is an ordered list of op records (kind + fields holding torch tensors / scalars) with
named segments.  `finalize()` lowers every record to its C descriptor (include/b2h_abi.h) inside a
native `b2h_program`; `run(segment)` replays a segment on the current CUDA stream with ONE C call,
which makes the whole generator / discriminator step a single capturable launch sequence.

The records are plain data so that tests can interpret the very same program with the CPU op
restatements in oracle/ops_emul.py (graph-logic check without a GPU).  The product path never does
that: `run()` always goes through libb2h.so and raises if it is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from contextlib import contextmanager
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L

# timing experiment only (tools/skip_exp.sh): ops whose tag matches are left out of every replay -- results invalid
_SKIP_RE = re.compile(os.environ["B2H_DBG_SKIP_OPS"]) if os.environ.get("B2H_DBG_SKIP_OPS") else None


class OpRec:
    __slots__ = ("kind", "f", "tag")

    def __init__(self, kind: int, tag: str, fields: dict):
        self.kind, self.tag, self.f = kind, tag, fields

    def __repr__(self):
        return f"OpRec({L.OP_STRUCT[self.kind].__name__}, {self.tag})"


def _ptr(v):
    if v is None:
        return None
    if isinstance(v, torch.Tensor):
        return v.data_ptr()
    return int(v)


def _fill_struct(st, fields: dict):
    """Recursively copy a dict of python values / tensors into a ctypes Structure."""
    for name, ctype in st._fields_:
        if name not in fields:
            continue
        v = fields[name]
        if isinstance(ctype, type) and issubclass(ctype, C.Structure):
            _fill_struct(getattr(st, name), v or {})
        elif isinstance(ctype, type) and issubclass(ctype, C.Array):
            arr = getattr(st, name)
            elem = ctype._type_
            if isinstance(elem, type) and issubclass(elem, C.Structure):
                for i, item in enumerate(v):
                    _fill_struct(arr[i], item)
            elif isinstance(elem, type) and issubclass(elem, C.Array):
                for i, row in enumerate(v):
                    for j, x in enumerate(row):
                        arr[i][j] = x
            else:
                for i, x in enumerate(v):
                    arr[i] = x
        elif ctype is L.vp:
            setattr(st, name, _ptr(v))
        else:
            setattr(st, name, v)
    return st


def lower(rec: OpRec):
    st = L.OP_STRUCT[rec.kind]()
    unknown = {k for k in rec.f if not k.startswith("_")} - {n for n, _ in st._fields_}
    if unknown:
        raise KeyError(f"{rec}: unknown fields {sorted(unknown)}")
    return _fill_struct(st, rec.f)


def no_drop():
    return {"mode": L.DROP_NONE, "site": 0, "mask": None, "state": None, "save": None}


class Program:
    def __init__(self, dtype: int, device: torch.device):
        self.dtype = dtype
        self.device = torch.device(device)
        self.recs: List[OpRec] = []
        self.segments: Dict[str, Tuple[int, int]] = {}
        self._handle = None
        self.segment_launches: Dict[str, int] = {}  # kernel launches of the last run of each segment

    @property
    def act_dtype(self):
        return torch.bfloat16 if self.dtype == L.BF16 else torch.float32

    def add(self, op_kind: int, tag: str = "", /, **fields) -> int:
        assert self._handle is None, "program already lowered: invalidate() it first"
        self.recs.append(OpRec(op_kind, tag, fields))
        return len(self.recs) - 1

    @contextmanager
    def segment(self, name: str):
        start = len(self.recs)
        yield
        assert name not in self.segments, name
        self.segments[name] = (start, len(self.recs))

    # ---- native execution -------------------------------------------------------------------
    def finalize(self):
        if self.device.type != "cuda":
            raise L.B2HError("programs execute on a CUDA device only (no CPU fallback)")
        lib = L.load()
        L.require_device()
        h = lib.b2h_program_create(self.dtype)
        if not h:
            raise L.B2HError("b2h_program_create failed: " + lib.b2h_last_error().decode())
        self._handle = C.c_void_p(h)
        for rec in self.recs:
            st = lower(rec)
            rc = lib.b2h_program_add(self._handle, rec.kind, C.byref(st))
            L.check(rc, f"b2h_program_add[{rec}]")
        return self

    def invalidate(self):
        """A record changed after the program was lowered: destroy the native program; the next run lowers it again
        (CUDA graphs that captured its launches keep the OLD arguments and must be re-captured by their owner)."""
        if self._handle is not None:
            L.load().b2h_program_destroy(self._handle)
            self._handle = None

    def run(self, segment: Optional[str] = None, stream: Optional[int] = None):
        if self._handle is None:
            self.finalize()
        first, end = self.segments[segment] if segment is not None else (0, len(self.recs))
        if end == first:
            return
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        if _SKIP_RE is not None:
            return self._run_skipping(first, end, stream)
        rc = L.load().b2h_program_run(self._handle, first, end - first, C.c_void_p(stream))
        L.check(rc, f"b2h_program_run[{segment}]")
        self.segment_launches[segment or "*"] = self.launches()

    def run_range(self, first: int, end: int, stream: Optional[int] = None):
        if self._handle is None:
            self.finalize()
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        if _SKIP_RE is not None:     # timing experiment (B2H_DBG_SKIP_OPS=regex): results are invalid by construction
            return self._run_skipping(first, end, stream)
        rc = L.load().b2h_program_run(self._handle, first, end - first, C.c_void_p(stream))
        L.check(rc, f"b2h_program_run[{first}:{end}]")

    def _run_skipping(self, first: int, end: int, stream: int):
        """B2H_DBG_SKIP_OPS: replay [first, end) without the ops whose tag matches -- what would the step cost if those
        launches were gone?  (tools/skip_exp.sh; never set in a run whose results are used)"""
        i = first
        while i < end:
            if _SKIP_RE.search(self.recs[i].tag):
                i += 1
                continue
            j = i
            while j < end and not _SKIP_RE.search(self.recs[j].tag):
                j += 1
            L.check(L.load().b2h_program_run(self._handle, i, j - i, C.c_void_p(stream)), f"b2h_program_run[{i}:{j}]")
            i = j

    def op_plan(self, idx: int) -> dict:
        """Launch plan of op `idx` as the library built it (b2h_program_op_plan): tile width, split-K, fusions."""
        if self._handle is None:
            self.finalize()
        st = L.OpPlan()
        L.check(L.load().b2h_program_op_plan(self._handle, idx, C.byref(st)), f"b2h_program_op_plan[{idx}]")
        return {"tag": self.recs[idx].tag, "kind": st.kind, "tensor_core": bool(st.tensor_core), "tile_n": st.tile_n,
                "splits": st.splits, "merged": bool(st.merged), "fuse_stats": bool(st.fuse_stats),
                "fuse_bwd": bool(st.fuse_bwd), "epilogue": st.epilogue, "grid": tuple(st.grid)}

    def tile_report(self) -> List[dict]:
        """op_plan() of every GEMM-class op (tap-GEMMs and weight gradients) of the program."""
        return [self.op_plan(i) for i, r in enumerate(self.recs) if r.kind in (L.OP_GEMM, L.OP_WGRAD)]

    def launches(self) -> int:
        return int(L.load().b2h_program_launches(self._handle)) if self._handle else 0

    def __del__(self):
        try:
            if self._handle is not None and L._lib is not None:
                L._lib.b2h_program_destroy(self._handle)
        except Exception:
            pass
