"""Import alias: `import b2h_b200` == the package directory
`multimodal-hand-pose-enhancement-for-sign-language_b200/` (whose name is not a Python identifier)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("multimodal-hand-pose-enhancement-for-sign-language_b200")
sys.modules[__name__] = _pkg
