#!/usr/bin/env python
"""Generate tests/golden/losses.npz by running the REAL reference (/root/reference, read-only): every entry of
LOSSES (utils/constants.py:53-58) evaluated the way train_gan.py:286-292 does — "RobustLoss" through the real
robust_loss AdaptiveLossFunction at the alpha / scale it is constructed with (train_gan.py:74-77; never optimised) —
on procedural (out, gt) pairs: the loss value and its gradient with respect to `out`.

Run in the authoring container only:  python tools/make_golden_losses.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.golden_common import _gen  # noqa: E402

REF = "/root/reference"
B, C, T = 3, 252, 8


def main():
    sys.path.insert(0, os.path.join(REF, "utils"))
    try:
        import constants
    finally:
        sys.path.pop(0)
    g = _gen("losses")
    out = (torch.randn(B, C, T, generator=g) * 1.5)
    gt = torch.randn(B, C, T, generator=g)
    out[0, 0, :4] = gt[0, 0, :4]                          # exact zeros: sign(0) = 0, Huber's quadratic branch
    out[0, 1, 0], out[0, 1, 1] = gt[0, 1, 0] + 1.0, gt[0, 1, 1] - 1.0   # |d| = delta
    fix = {"out": out.numpy(), "gt": gt.numpy()}
    for name in ("L1", "L2", "Huber1", "RobustLoss"):
        o = out.clone().requires_grad_(True)
        crit = constants.LOSSES[name]
        if name == "RobustLoss":
            crit = crit(num_dims=C * T, float_dtype=torch.float32, device="cpu")
            loss = torch.mean(crit.lossfun(torch.reshape(o, (B, -1)) - torch.reshape(gt, (B, -1))))
            fix["robust_alpha"] = crit.alpha().detach().numpy()[0, :4]
            fix["robust_scale"] = crit.scale().detach().numpy()[0, :4]
        else:
            loss = crit(o, gt)
        grad, = torch.autograd.grad(loss, o)
        fix[name + "_loss"] = np.float64(loss.item())
        fix[name + "_grad"] = grad.numpy()
    path = os.path.join(ROOT, "tests", "golden", "losses.npz")
    np.savez_compressed(path, **fix)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in fix.items()})


if __name__ == "__main__":
    main()
