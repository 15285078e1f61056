"""Procedural, order-independent parameter / input / mask generators shared by tools/make_golden.py
(which runs the REAL reference in the authoring container) and tests/test_golden.py (which replays
the fixtures against the oracle and the CUDA path on any box)."""
import zlib

import torch

CASES = [  # name, variant, require_feats, in_dim, out_dim, B, T
    ("v1_body", "v1", False, 36, 252, 2, 16),
    ("v1_text", "v1", True, 36, 252, 2, 16),
    ("b2h_image", "b2h", True, 36, 252, 2, 8),
    ("v2_text_finger1", "v2", True, 264, 24, 2, 16),
    ("v4_text", "v4", True, 36, 252, 2, 16),
    ("v4_deeper_text", "v4_deeper", True, 36, 252, 2, 16),
    ("v1_body_T192", "v1", False, 36, 252, 1, 192),
]
DISC_CASES = [("disc_T64", 252, 6, 64), ("disc_T192", 252, 4, 192)]


def _gen(tag: str):
    return torch.Generator().manual_seed(zlib.crc32(tag.encode()) & 0x7FFFFFFF)


def fill_state_dict(sd):
    """Deterministic values for every entry of a state_dict, keyed by name only."""
    out = {}
    for k, v in sd.items():
        g = _gen("param:" + k)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.tensor(3, dtype=torch.int64)
        elif k.endswith("running_var"):
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith("running_mean"):
            out[k] = torch.randn(v.shape, generator=g) * 0.2
        elif v.dim() >= 2:
            fan_in = v.shape[1] * (v.shape[2] if v.dim() == 3 else 1)
            out[k] = (torch.rand(v.shape, generator=g) * 2 - 1) / fan_in ** 0.5
        elif k.endswith("weight"):      # BN gamma (1-D): positive and negative scales
            out[k] = (torch.rand(v.shape, generator=g) * 1.5 + 0.25) * torch.where(
                torch.rand(v.shape, generator=g) < 0.2, -1.0, 1.0)
        else:                            # biases
            out[k] = torch.randn(v.shape, generator=g) * 0.1
    return out


def inputs(name, cin, cout, B, T, feats_kind):
    g = _gen("input:" + name)
    x = torch.randn(B, cin, T, generator=g)
    y = torch.randn(B, cout, T, generator=g)
    f = None
    if feats_kind == "text":
        f = torch.randn(B, 512, generator=g)
    elif feats_kind == "image":
        f = torch.randn(B, T, 2000, generator=g)
    return x, y, f


def mask_for(name, site, shape):
    return (torch.rand(shape, generator=_gen(f"mask:{name}:{site}")) < 0.5).to(torch.uint8)


def grad_digest(t):
    t = t.detach().double().reshape(-1)
    return torch.stack([t.sum(), t.abs().sum(), (t * torch.arange(1, t.numel() + 1, dtype=torch.float64)).sum()
                        / max(t.numel(), 1)])
