"""Import alias: `import b2h_b200` == the package directory
`multimodal-hand-pose-enhancement-for-sign-language_b200/` (whose name is not a Python identifier).

Every submodule is registered under both names, so `b2h_b200.trainer` and the package's own relative imports are
the SAME module objects (one `_lib` handle, one `B2HError` class, one `Program` class)."""
import importlib
import os
import pkgutil
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_REAL = "multimodal-hand-pose-enhancement-for-sign-language_b200"
_pkg = importlib.import_module(_REAL)
sys.modules[__name__] = _pkg
for _m in pkgutil.iter_modules(_pkg.__path__):
    if not _m.ispkg and os.path.exists(os.path.join(_pkg.__path__[0], _m.name + ".py")):   # (libb2h.so is no module)
        sys.modules[f"{__name__}.{_m.name}"] = importlib.import_module(f"{_REAL}.{_m.name}")
