"""The drop-in boundary: modelZoo class names, build_net / forward signatures, state_dict layout and
default initialisation must equal the reference's (checked against the real reference where mounted,
against the oracle restatement everywhere)."""
import inspect

import pytest
import torch

import modelZoo
from b2h_b200 import _lib as L
from oracle import ref_models as R

CASES = [("v1", False), ("v1", True), ("b2h", False), ("b2h", True), ("v2", True), ("v4", True), ("v4_deeper", True),
         ("v4_deeper", False)]


def _build(mz, variant, rf, cin=36, cout=252):
    m = getattr(mz, R.REF_CLASS[variant])()
    if variant == "b2h":
        m.build_net(cin, cout, require_image=rf)
    else:
        m.build_net(cin, cout, require_text=rf)
    return m


@pytest.mark.parametrize("variant,rf", CASES)
def test_state_dict_layout_matches_oracle(variant, rf):
    ours = _build(modelZoo, variant, rf)
    ora = R.build_generator(variant, 36, 252, rf)
    a, b = ours.state_dict(), ora.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
    assert [k for k, _ in ours.named_parameters()] == [k for k, _ in ora.named_parameters()]
    ours.load_state_dict(b)
    ours.load_state_dict(b, strict=False)


def test_discriminator_layout():
    d = modelZoo.regressor_fcn_bn_discriminator()
    d.build_net(252)
    ora = R.build_discriminator(252)
    assert list(d.state_dict().keys()) == list(ora.state_dict().keys())
    for k, v in ora.state_dict().items():
        assert d.state_dict()[k].shape == v.shape


@pytest.mark.reference
@pytest.mark.parametrize("variant,rf", CASES)
def test_same_seed_same_init_as_reference(reference_modelzoo, variant, rf):
    torch.manual_seed(23456)
    ref = _build(reference_modelzoo, variant, rf)
    torch.manual_seed(23456)
    ours = _build(modelZoo, variant, rf)
    for (ka, va), (kb, vb) in zip(ref.state_dict().items(), ours.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    # forward signature parity (train_gan.py:280, inference.py:115)
    assert list(inspect.signature(ref.forward).parameters) == list(inspect.signature(ours.forward).parameters)
    assert list(inspect.signature(ref.build_net).parameters) == list(inspect.signature(ours.build_net).parameters)


def test_cpu_call_fails_loudly():
    m = _build(modelZoo, "v1", False)
    with pytest.raises(L.B2HError):
        m(torch.zeros(2, 36, 16))


@pytest.mark.gpu
@pytest.mark.parametrize("variant,rf", [("v1", False), ("v1", True), ("b2h", True), ("v4", True)])
def test_module_forward_backward_on_gpu(variant, rf):
    """Eval forward at 1e-5, and a train step driven by torch autograd + torch.optim.Adam on the module's own
    parameters (the unmodified train_gan.py flow) moves the parameters like the oracle does."""
    torch.manual_seed(0)
    ora = R.build_generator(variant, 36, 252, rf)
    ours = _build(modelZoo, variant, rf)
    ours.load_state_dict(ora.state_dict())
    ours.to("cuda")
    B, T = 8, 64
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 36, T, generator=g)
    y = torch.randn(B, 252, T, generator=g)
    f = None if not rf else (torch.randn(B, T, 2000, generator=g) if variant == "b2h" else torch.randn(B, 512, generator=g))
    fc = f.cuda() if f is not None else None
    ours.eval(), ora.eval()
    with torch.no_grad():
        ref = ora(x, feats_=f)
    out = ours(x.cuda(), feats_=fc)
    assert out.shape == ref.shape
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) < 1e-5
    # train mode through autograd; dropout differs (Philox vs torch RNG) so compare with dropout-free oracle
    # statistics only: loss decreases and gradients are finite and non-zero for every live parameter
    ours.train()
    opt = torch.optim.Adam(ours.parameters(), lr=1e-3)
    losses = []
    for _ in range(5):
        o = ours(x.cuda(), feats_=fc)
        loss = torch.nn.functional.l1_loss(o, y.cuda())
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
    live = [p for k, p in ours.named_parameters() if p.grad is not None]
    assert len(live) >= 30 and all(torch.isfinite(p.grad).all() for p in live)
    sd = ours.state_dict()
    assert int(sd["encoder.3.num_batches_tracked"]) == 5


@pytest.mark.gpu
def test_discriminator_module_on_gpu():
    torch.manual_seed(0)
    ora = R.build_discriminator(252)
    d = modelZoo.regressor_fcn_bn_discriminator()
    d.build_net(252)
    d.load_state_dict(ora.state_dict())
    d.to("cuda")
    x = torch.randn(16, 252, 63)
    d.eval(), ora.eval()
    with torch.no_grad():
        ref = ora(x)
    out = d(x.cuda())
    assert out.shape == ref.shape == (16, 1, 1)
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) < 1e-5
    d.train()
    sc = d(x.cuda())
    torch.nn.functional.mse_loss(sc, torch.ones_like(sc)).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in d.parameters())


@pytest.mark.gpu
def test_two_train_forwards_before_one_backward():
    """The reference's discriminator step (train_gan.py:240-249): D(fake_motion), D(real_motion), then ONE
    d_loss.backward().  Both forwards have the same shape; each must keep its own saved activations / batch
    statistics (a plan lease), so the gradients equal the oracle's (dropout off on both sides)."""
    torch.manual_seed(0)
    ora = R.build_discriminator(252)
    d = modelZoo.regressor_fcn_bn_discriminator()
    d.build_net(252)
    d.load_state_dict(ora.state_dict())
    d.drop_mode = "none"
    d.to("cuda")
    for m in ora.modules():
        if isinstance(m, R.ReplayDropout):
            m.p = 0.0
    g = torch.Generator().manual_seed(3)
    fake, real = torch.randn(16, 252, 63, generator=g), torch.randn(16, 252, 63, generator=g) * 1.5 + 0.3
    ora.train(), d.train()
    fs, rs = ora(fake), ora(real)
    (torch.nn.functional.mse_loss(fs, torch.zeros_like(fs)) + torch.nn.functional.mse_loss(rs, torch.ones_like(rs))).backward()
    fs2, rs2 = d(fake.cuda()), d(real.cuda())
    assert float((fs2.cpu() - fs).abs().max()) <= 1e-5 * float(fs.abs().max())
    assert float((rs2.cpu() - rs).abs().max()) <= 1e-5 * float(rs.abs().max())
    (torch.nn.functional.mse_loss(fs2, torch.zeros_like(fs2)) + torch.nn.functional.mse_loss(rs2, torch.ones_like(rs2))).backward()
    for (k, p), (_, q) in zip(ora.named_parameters(), d.named_parameters()):
        den = float(p.grad.abs().max()) + 1e-12
        assert float((q.grad.cpu() - p.grad).abs().max()) <= 5e-4 * den, k
    # both leases are back: a third and a fourth forward reuse the two plans
    assert sum(len(v) for v in d._plans.values()) == 2
    d(fake.cuda()), d(real.cuda())
    assert sum(len(v) for v in d._plans.values()) == 2


def test_plan_lease_bookkeeping():
    """Leases without a device: a leased plan is not handed out again until released; dropping the autograd node
    releases it."""
    from b2h_b200.modelzoo import _Lease

    class P:
        _leased = False
    a = P()
    l1 = _Lease(a)
    assert a._leased
    l1.release()
    assert not a._leased
    l2 = _Lease(a)
    del l2
    assert not a._leased
