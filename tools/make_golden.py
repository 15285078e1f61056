#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (/root/reference, read-only) on procedural
inputs: eval outputs, train-mode outputs with replayed dropout masks, gradient digests, BN buffers
after one train forward, discriminator scores, calc_motion and np_rot6d_to_mat vectors.

Run in the authoring container only:  python tools/make_golden.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_models import REF_CLASS, ReplayDropout  # noqa: E402  (mask replay module only)
from tools import golden_common as GC  # noqa: E402

REF = "/root/reference"


def load_ref(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def install_mask_replay(model, case):
    """SURVEY 8c recipe: replace every nn.Dropout child of every nn.Sequential by a mask-replaying module."""
    store = {}
    shapes = {}
    for name, seq in model.named_modules():
        if isinstance(seq, torch.nn.Sequential):
            for i, child in enumerate(seq):
                if isinstance(child, torch.nn.Dropout):
                    rd = ReplayDropout(child.p)
                    rd.site = f"{name}.{i}"
                    rd.store = store
                    seq[i] = rd
    return store, shapes


def main():
    mz = load_ref("modelZoo.py", "_ref_modelZoo")
    sys.path.insert(0, os.path.join(REF, "utils"))
    conv = load_ref("utils/conversion_utils.py", "_ref_conv")
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, variant, rf, cin, cout, B, T in GC.CASES:
        m = getattr(mz, REF_CLASS[variant])()
        if variant == "b2h":
            m.build_net(cin, cout, require_image=rf)
        else:
            m.build_net(cin, cout, require_text=rf)
        m.load_state_dict(GC.fill_state_dict(m.state_dict()))
        kind = None if not rf else ("image" if variant == "b2h" else "text")
        x, y, f = GC.inputs(name, cin, cout, B, T, kind)
        m.eval()
        with torch.no_grad():
            out_eval = m(x, feats_=f)
        # train mode with replayed masks
        store, _ = install_mask_replay(m, name)
        m.eval()
        # trace the dropout input shapes once (eval mode leaves them untouched), then draw the masks
        shapes = {}
        hooks = [mod.register_forward_pre_hook(lambda mod, inp: shapes.__setitem__(mod.site, tuple(inp[0].shape)))
                 for mod in m.modules() if isinstance(mod, ReplayDropout)]
        with torch.no_grad():
            m(x, feats_=f)
        for h in hooks:
            h.remove()
        for site, shp in shapes.items():
            store[site] = GC.mask_for(name, site, shp)
        m.train()
        out_train = m(x, feats_=f)
        loss = torch.nn.functional.l1_loss(out_train, y)
        loss.backward()
        rec = {"out_eval": out_eval.numpy(), "out_train": out_train.detach().numpy(), "l1": np.float64(loss.item())}
        for k, p in m.named_parameters():
            if p.grad is not None:
                rec["grad:" + k] = GC.grad_digest(p.grad).numpy()
        for k, v in m.state_dict().items():
            if k.endswith(("running_mean", "running_var")):
                rec["buf:" + k] = v.numpy().copy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print("wrote", name, {k: getattr(v, "shape", None) for k, v in list(rec.items())[:3]})
    tg = load_ref("train_gan.py", "_ref_train_gan") if False else None  # train_gan imports wandb at module top
    for name, cin, B, T in GC.DISC_CASES:
        d = mz.regressor_fcn_bn_discriminator()
        d.build_net(cin)
        d.load_state_dict(GC.fill_state_dict(d.state_dict()))
        x, y, _ = GC.inputs(name, cin, cin, B, T, None)
        motion = x[:, :, :1] - x[:, :, :-1]            # train_gan.py:209-211 (restated; module import needs wandb)
        d.eval()
        with torch.no_grad():
            score_eval = d(motion)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), score_eval=score_eval.numpy(),
                            motion_digest=GC.grad_digest(motion).numpy())
        print("wrote", name, score_eval.shape)
    # np_rot6d_to_mat, row by row, on the reference's own function
    g = torch.Generator().manual_seed(7)
    r6d = torch.randn(256, 6, generator=g).numpy().astype(np.float64)
    mats = np.stack([conv.np_rot6d_to_mat(r6d[i:i + 1])[0] for i in range(r6d.shape[0])])
    aa = conv._rot6d_to_aa(r6d[:32])
    np.savez_compressed(os.path.join(out_dir, "rot6d.npz"), r6d=r6d, mat=mats, aa32=aa)
    print("wrote rot6d", mats.shape)


if __name__ == "__main__":
    main()
