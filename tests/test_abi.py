"""The C-ABI library loads (no GPU needed) and exports every function include/b2h_abi.h declares; the ctypes
mirrors have the same size as the C structs; compute entry points refuse to run without a device."""
import ctypes
import os
import re

import b2h_b200  # noqa: F401
from b2h_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "b2h_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2h_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b2h_abi.h but not exported by libb2h.so"
        assert n in L.SYMBOLS, f"{n} has no ctypes prototype in _lib.py"
    assert lib.b2h_abi_version() == 1


def test_struct_sizes_match():
    lib = L.load()
    for kind, st in L.OP_STRUCT.items():
        assert lib.b2h_desc_size(kind) == ctypes.sizeof(st), st.__name__


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        return
    lib = L.load()
    assert lib.b2h_check_device() < 0
    assert len(lib.b2h_last_error()) > 0


def test_library_has_no_driver_link_dependency():
    # libcuda.so is resolved at run time (cudaGetDriverEntryPoint), so the library loads on a CPU-only box
    import subprocess
    out = subprocess.run(["ldd", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out
