"""Golden vectors produced by the REAL reference (tools/make_golden.py, run where /root/reference is
mounted).  CPU tests pin the oracle restatement against them on any box; GPU tests check the CUDA path
against the same vectors (fp32 mode, 1e-5)."""
import os

import numpy as np
import pytest
import torch

import b2h_b200  # noqa: F401
from oracle import ops_emul as E
from oracle import ref_models as R
from tools import golden_common as GC

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _masks(name, model, x, f):
    shapes = {k: tuple(v.shape) for k, v in R.make_masks(model, x, seed=0, feats=f).items()}
    return {site: GC.mask_for(name, site, shp) for site, shp in shapes.items()}


@pytest.mark.parametrize("name,variant,rf,cin,cout,B,T", GC.CASES)
def test_oracle_generator_matches_golden(name, variant, rf, cin, cout, B, T):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    G = R.build_generator(variant, cin, cout, rf)
    G.load_state_dict(GC.fill_state_dict(G.state_dict()))
    kind = None if not rf else ("image" if variant == "b2h" else "text")
    x, y, f = GC.inputs(name, cin, cout, B, T, kind)
    G.eval()
    with torch.no_grad():
        assert rel_err(G(x, feats_=f), gold["out_eval"]) < 1e-6
    masks = _masks(name, G, x, f)
    G.train()
    G.set_masks(masks)
    out = G(x, feats_=f)
    assert rel_err(out, gold["out_train"]) < 1e-6
    loss = torch.nn.functional.l1_loss(out, y)
    assert abs(loss.item() - float(gold["l1"])) < 1e-6
    loss.backward()
    for k, p in G.named_parameters():
        key = "grad:" + k
        if p.grad is None:
            assert key not in gold.files, k
        else:
            assert rel_err(GC.grad_digest(p.grad), gold[key]) < 1e-5, k
    for k, v in G.state_dict().items():
        if k.endswith(("running_mean", "running_var")):
            assert rel_err(v, gold["buf:" + k]) < 1e-6, k


@pytest.mark.parametrize("name,cin,B,T", GC.DISC_CASES)
def test_oracle_discriminator_matches_golden(name, cin, B, T):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    D = R.build_discriminator(cin)
    D.load_state_dict(GC.fill_state_dict(D.state_dict()))
    x, _, _ = GC.inputs(name, cin, cin, B, T, None)
    motion = R.calc_motion(x)
    assert rel_err(GC.grad_digest(motion), gold["motion_digest"]) < 1e-9
    D.eval()
    with torch.no_grad():
        assert rel_err(D(motion), gold["score_eval"]) < 1e-6


def test_oracle_rot6d_matches_golden():
    gold = np.load(os.path.join(GOLD, "rot6d.npz"))
    got = E.rot6d_to_mat(torch.from_numpy(gold["r6d"]))
    assert rel_err(got, gold["mat"]) < 1e-12


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,variant,rf,cin,cout,B,T", GC.CASES)
def test_cuda_generator_matches_golden(name, variant, rf, cin, cout, B, T):
    from b2h_b200 import _lib as L
    from b2h_b200 import nets
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    kind = None if not rf else ("image" if variant == "b2h" else "text")
    x, y, f = GC.inputs(name, cin, cout, B, T, kind)
    ref_sd = GC.fill_state_dict(R.build_generator(variant, cin, cout, rf).state_dict())
    for train in (False, True):
        spec = nets.generator_spec(variant, cin, cout, rf, train=train)
        store = nets.ParamStore(spec, "cuda", seed=0)
        store.load_state_dict({k: v.cuda() for k, v in ref_sd.items()})
        plan = nets.NetPlan(spec, store, B, T, L.F32, "cuda", train=train, drop_mode="mask" if train else "none")
        plan.x.copy_(x)
        if f is not None:
            plan.feats.copy_(f)
        if train:
            plan.set_masks({site: GC.mask_for(name, site, tuple(m.shape) if m.dim() == 3 else tuple(m.shape))
                            for site, m in _ref_shaped_masks(name, variant, cin, cout, rf, x, f).items()})
        plan.forward()
        torch.cuda.synchronize()
        assert rel_err(plan.out, gold["out_train" if train else "out_eval"]) < 1e-5, ("train" if train else "eval")
        if train:
            for k, _ in store.buffer_shapes:
                if k.rsplit(".", 1)[0] in {l.bnkey for l in spec.layers}:
                    assert rel_err(store.b(k), gold["buf:" + k]) < 1e-5, k


def _ref_shaped_masks(name, variant, cin, cout, rf, x, f):
    G = R.build_generator(variant, cin, cout, rf)
    return _masks(name, G, x, f)


@pytest.mark.gpu
def test_cuda_rot6d_matches_golden():
    from b2h_b200 import _lib as L
    gold = np.load(os.path.join(GOLD, "rot6d.npz"))
    r6d = torch.from_numpy(gold["r6d"]).float().cuda()
    out = torch.empty(r6d.shape[0], 9, device="cuda")
    L.run_oneshot(L.Rot6d(r6d=r6d.data_ptr(), mat=out.data_ptr(), n=r6d.shape[0]), L.F32,
                  torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rel_err(out, gold["mat"]) < 1e-5


@pytest.mark.parametrize("loss", ["L1", "L2", "Huber1", "RobustLoss"])
def test_loss_table_matches_golden(loss):
    """tests/golden/losses.npz (tools/make_golden_losses.py: the reference's LOSSES table, "RobustLoss" through the real
    AdaptiveLossFunction) against the oracle's reg_criterion and against the restatement of the b2h_l1 op
    (value, gradient in the BLC layout, zero padding)."""
    gold = np.load(os.path.join(GOLD, "losses.npz"))
    out, gt = torch.from_numpy(gold["out"]), torch.from_numpy(gold["gt"])
    ref_loss, ref_grad = float(gold[loss + "_loss"]), torch.from_numpy(gold[loss + "_grad"])
    if loss == "RobustLoss":
        assert np.all(gold["robust_alpha"] == 2.0) and np.allclose(gold["robust_scale"], 0.5, atol=1e-7)
    o = out.clone().requires_grad_(True)
    val = R.reg_criterion(loss, o, gt)
    grad, = torch.autograd.grad(val, o)
    assert abs(float(val) - ref_loss) <= 1e-6 * abs(ref_loss)
    assert float((grad - ref_grad).abs().max()) <= 1e-6 * float(ref_grad.abs().max())
    B, C, T = out.shape
    Cp = 256
    f = dict(out=out.contiguous(), gt=gt.contiguous(), dout=torch.full((B, T, Cp), 7.0), loss=torch.zeros(1), B=B, C=C, L=T,
             ld=Cp, Cfill=Cp, gscale=1.0, kind={"L1": 0, "L2": 1, "Huber1": 2, "RobustLoss": 3}[loss])
    E.l1(f)
    assert abs(float(f["loss"][0]) - ref_loss) <= 1e-6 * abs(ref_loss)
    assert float((f["dout"][:, :, :C].permute(0, 2, 1) - ref_grad).abs().max()) <= 1e-6 * float(ref_grad.abs().max())
    assert float(f["dout"][:, :, C:].abs().max()) == 0.0
