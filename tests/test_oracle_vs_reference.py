"""Pin the oracle restatement (oracle/ref_models.py) against the REAL reference modules.

Runs only where /root/reference is mounted (the authoring container).  The same pins travel
to the GPU box as tests/golden/*.npz (see tools/make_golden.py, tests/test_golden.py).
"""
import numpy as np
import pytest
import torch

from oracle import ref_models as R

pytestmark = pytest.mark.reference

CASES = [  # (variant, require_feats, in_dim, out_dim, T)
    ("v1", False, 36, 252, 64), ("v1", True, 36, 252, 64), ("b2h", False, 36, 252, 32),
    ("b2h", True, 36, 252, 32), ("v2", False, 264, 24, 64), ("v2", True, 36, 252, 62),
    ("v4", False, 36, 252, 64), ("v4", True, 36, 252, 64), ("v4_deeper", False, 36, 252, 64),
    ("v4_deeper", True, 36, 252, 64), ("v1", False, 36, 252, 192), ("v1", False, 36, 252, 66),
]


def _build_ref(mz, variant, rf, cin, cout):
    m = getattr(mz, R.REF_CLASS[variant])()
    if variant == "b2h":
        m.build_net(cin, cout, require_image=rf)
    else:
        m.build_net(cin, cout, require_text=rf)
    return m


def _feats(variant, rf, B, T, g):
    if not rf:
        return None
    if variant == "b2h":
        return torch.randn(B, T, 2000, generator=g)
    return torch.randn(B, 512, generator=g)


@pytest.mark.parametrize("variant,rf,cin,cout,T", CASES)
def test_generator_matches_reference(reference_modelzoo, variant, rf, cin, cout, T):
    torch.manual_seed(23456)
    ref = _build_ref(reference_modelzoo, variant, rf, cin, cout)
    ora = R.build_generator(variant, cin, cout, rf)
    sd = ref.state_dict()
    osd = ora.state_dict()
    assert list(sd.keys()) == list(osd.keys())
    for k in sd:
        assert sd[k].shape == osd[k].shape, k
    ora.load_state_dict(sd)
    g = torch.Generator().manual_seed(1)
    B = 3
    x = torch.randn(B, cin, T, generator=g)
    f = _feats(variant, rf, B, T, g)
    # eval forward: bit-identical
    ref.eval(), ora.eval()
    with torch.no_grad():
        a, b = ref(x, feats_=f), ora(x, feats_=f)
    assert a.shape == (B, cout, T)
    assert torch.equal(a, b)
    # train forward/backward with the same torch RNG stream: bit-identical outputs, grads, BN buffers
    ref.train(), ora.train()
    torch.manual_seed(7)
    a = ref(x, feats_=f)
    a.abs().mean().backward()
    torch.manual_seed(7)
    b = ora(x, feats_=f)
    b.abs().mean().backward()
    assert torch.equal(a, b)
    for (k, p), (_, q) in zip(ref.named_parameters(), ora.named_parameters()):
        if p.grad is None:
            assert q.grad is None, k
        else:
            assert torch.equal(p.grad, q.grad), k
    for k, v in ref.state_dict().items():
        assert torch.equal(v, ora.state_dict()[k]), k


def test_discriminator_matches_reference(reference_modelzoo):
    torch.manual_seed(23456)
    ref = reference_modelzoo.regressor_fcn_bn_discriminator()
    ref.build_net(252)
    ora = R.build_discriminator(252)
    assert list(ref.state_dict().keys()) == list(ora.state_dict().keys())
    ora.load_state_dict(ref.state_dict())
    x = torch.randn(4, 252, 63)
    for T in (63, 191):
        x = torch.randn(4, 252, T)
        ref.eval(), ora.eval()
        with torch.no_grad():
            assert torch.equal(ref(x), ora(x))
    ref.train(), ora.train()
    torch.manual_seed(3)
    a = ref(x)
    torch.manual_seed(3)
    b = ora(x)
    assert torch.equal(a, b)
    assert a.shape == (4, 1, 2)


def test_mask_replay_equals_torch_dropout_scaling():
    d = R.ReplayDropout(0.5)
    d.site = "s"
    d.train()
    x = torch.randn(2, 3, 5)
    keep = (torch.rand(2, 3, 5) < 0.5).to(torch.uint8)
    d.store = {"s": keep}
    assert torch.equal(d(x), x * keep.float() * 2.0)


def test_mac_counts_match_survey():
    # hook-measured on the reference
    assert R.macs_per_clip(R.build_generator("v1", 36, 252), 64) == 81_369_600 or \
        abs(R.macs_per_clip(R.build_generator("v1", 36, 252), 64) - 81.370e6) < 1e3
    assert abs(R.macs_per_clip(R.build_generator("v1", 36, 252, True), 64, "text") - 214.310e6) < 1e3
    assert abs(R.macs_per_clip(R.build_generator("b2h", 36, 252, True), 64, "image") - 238.689e6) < 1e3
    assert abs(R.macs_per_clip(R.build_discriminator(252), 63) - 3.018e6) < 1e3


@pytest.mark.parametrize("loss", ["L1", "L2", "Huber1", "RobustLoss"])
def test_reg_criterion_matches_reference_losses(loss):
    """R.reg_criterion against LOSSES[--loss] evaluated exactly as train_gan.py:74-77,286-292 does, values and
    gradients; "RobustLoss" = the real AdaptiveLossFunction at its (never optimised) initial alpha / scale."""
    import sys
    sys.path.insert(0, "/root/reference/utils")
    try:
        import constants
    finally:
        sys.path.pop(0)
    g = torch.Generator().manual_seed(11)
    B, C, T = 4, 252, 16
    out = (torch.randn(B, C, T, generator=g) * 1.5).requires_grad_(True)
    gt = torch.randn(B, C, T, generator=g)
    crit = constants.LOSSES[loss]
    if loss == "RobustLoss":
        crit = crit(num_dims=C * T, float_dtype=torch.float32, device="cpu")
        assert float(crit.alpha().min()) == float(crit.alpha().max()) == 2.0
        assert abs(float(crit.scale().min()) - 0.5) < 1e-7 and abs(float(crit.scale().max()) - 0.5) < 1e-7
        ref = torch.mean(crit.lossfun(torch.reshape(out, (B, -1)) - torch.reshape(gt, (B, -1))))
    else:
        ref = crit(out, gt)
    ref_grad, = torch.autograd.grad(ref, out)
    ours = R.reg_criterion(loss, out, gt)
    our_grad, = torch.autograd.grad(ours, out)
    assert abs(float(ours) - float(ref)) <= 1e-6 * abs(float(ref))
    assert float((our_grad - ref_grad).abs().max()) <= 1e-6 * float(ref_grad.abs().max())


@pytest.fixture(scope="module")
def reference_train_gan():
    """The REAL /root/reference/train_gan.py imported unmodified (runs here under WANDB_MODE=disabled)."""
    import importlib.util
    import os
    import sys
    os.environ["WANDB_MODE"] = "disabled"
    ref = "/root/reference"
    added = [ref, os.path.join(ref, "utils"), os.path.join(ref, "viz")]
    saved = {k: sys.modules.pop(k) for k in ("modelZoo", "constants", "load_save_utils", "standardization_utils",
                                              "postprocess_utils") if k in sys.modules}
    sys.path[:0] = added
    try:
        spec = importlib.util.spec_from_file_location("_ref_train_gan", os.path.join(ref, "train_gan.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        import wandb
        wandb.init(mode="disabled")
    finally:
        for p in added:
            sys.path.remove(p)
        for k in ("modelZoo", "constants", "load_save_utils", "standardization_utils", "postprocess_utils"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    mod.device = torch.device("cpu")
    return mod


@pytest.mark.parametrize("variant,rf,label_smooth", [("v1", False, False), ("v1", True, True), ("b2h", True, False)])
def test_step_bodies_match_reference_train_gan(reference_train_gan, variant, rf, label_smooth):
    """oracle.generator_step / discriminator_step against the reference's own train_generator / train_discriminator
    (train_gan.py:215-308) over two batches each, same torch RNG stream (dropout included): bit-identical parameters,
    Adam moments and BN buffers, and the printed average losses."""
    import argparse
    tg = reference_train_gan
    mz = tg.modelZoo
    B, T, cin, cout, lr = 4, 32, 36, 252, 1e-3
    torch.manual_seed(23456)
    Gr = _build_ref(mz, variant, rf, cin, cout)
    Dr = mz.regressor_fcn_bn_discriminator()
    Dr.build_net(cout)
    Go, Do = R.build_generator(variant, cin, cout, rf), R.build_discriminator(cout)
    Go.load_state_dict(Gr.state_dict())
    Do.load_state_dict(Dr.state_dict())
    g = torch.Generator().manual_seed(5)
    X, Y = torch.randn(2 * B + 1, cin, T, generator=g).numpy(), torch.randn(2 * B + 1, cout, T, generator=g).numpy()
    F = _feats(variant, rf, 2 * B + 1, T, g)
    Fn = F.numpy() if F is not None else None
    args = argparse.Namespace(batch_size=B, num_epochs=3, log_step=1, epoch=1, loss="L1", disc_label_smooth=label_smooth,
                              require_text=rf and variant != "b2h", require_image=rf and variant == "b2h")
    opt = lambda m: torch.optim.Adam(m.parameters(), lr=lr, weight_decay=0)   # noqa: E731  (train_gan.py:69,88)
    gr_opt, dr_opt, go_opt, do_opt = opt(Gr), opt(Dr), opt(Go), opt(Do)
    # ---- reference: one generator epoch, one discriminator epoch (2 full batches each, the 9th clip is dropped)
    torch.manual_seed(11)
    tg.train_generator(args, Gr, Dr, torch.nn.L1Loss(), torch.nn.MSELoss(), gr_opt, X, Y, 1, train_feats=Fn)
    tg.train_discriminator(args, Gr, Dr, torch.nn.MSELoss(), dr_opt, X, Y, 2, train_feats=Fn)
    # ---- oracle restatement of the same step bodies
    torch.manual_seed(11)
    tx, ty = torch.from_numpy(X), torch.from_numpy(Y)
    for i in range(2):
        sl = slice(i * B, (i + 1) * B)
        R.generator_step(Go, Do, go_opt, tx[sl], ty[sl], F[sl] if F is not None else None)
    for i in range(2):
        sl = slice(i * B, (i + 1) * B)
        R.discriminator_step(Go, Do, do_opt, tx[sl], ty[sl], F[sl] if F is not None else None,
                             label_smooth=label_smooth)
    for ref, ora in ((Gr, Go), (Dr, Do)):
        for k, v in ref.state_dict().items():
            assert torch.equal(v, ora.state_dict()[k]), k
    for ro, oo in ((gr_opt, go_opt), (dr_opt, do_opt)):
        rs, os_ = ro.state_dict()["state"], oo.state_dict()["state"]
        assert rs.keys() == os_.keys()
        for i in rs:
            assert torch.equal(rs[i]["exp_avg"], os_[i]["exp_avg"]) and torch.equal(rs[i]["exp_avg_sq"], os_[i]["exp_avg_sq"])
