import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (authoring container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(REFERENCE)
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    skip_ref = pytest.mark.skip(reason="/root/reference not mounted")
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(skip_gpu)


@pytest.fixture(scope="session")
def reference_modelzoo():
    """The real reference modelZoo, imported read-only from /root/reference."""
    if not os.path.isdir(REFERENCE):
        pytest.skip("/root/reference not mounted")
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_modelZoo", os.path.join(REFERENCE, "modelZoo.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
