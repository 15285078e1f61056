"""Host-logic check without a GPU: interpret the recorded programs (graph wiring, residual / up /
pool routing, backward formulas, weight packing) with the CPU op restatements of
oracle/ops_emul.py and compare against the oracle models (oracle/ref_models.py)."""
import pytest
import torch

import b2h_b200  # noqa: F401
from b2h_b200 import _lib as L
from b2h_b200 import nets
from oracle import ops_emul as E
from oracle import ref_models as R


def rel_err(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def randomize_bn(model, seed=5):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 1.5 + 0.25)
                m.weight.mul_(torch.where(torch.rand(m.weight.shape, generator=g) < 0.2, -1.0, 1.0))
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.3)
                m.running_mean.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)
                m.running_var.copy_(torch.rand(m.bias.shape, generator=g) + 0.5)


def feats_for(variant, rf, B, T, g):
    if not rf:
        return None
    return torch.randn(B, T, 2000, generator=g) if variant == "b2h" else torch.randn(B, 512, generator=g)


GEN_CASES = [("v1", False, 36, 252, 16), ("v1", True, 36, 252, 16), ("b2h", True, 36, 252, 8),
             ("v2", True, 36, 48, 16), ("v4", True, 42, 246, 16), ("v4_deeper", True, 36, 252, 16),
             ("v1", False, 162, 126, 24), ("v1", False, 36, 252, 14)]


@pytest.mark.parametrize("variant,rf,cin,cout,T", GEN_CASES)
def test_generator_eval_forward(variant, rf, cin, cout, T):
    torch.manual_seed(0)
    G = R.build_generator(variant, cin, cout, rf)
    randomize_bn(G)
    G.eval()
    g = torch.Generator().manual_seed(1)
    B = 3
    x = torch.randn(B, cin, T, generator=g)
    f = feats_for(variant, rf, B, T, g)
    with torch.no_grad():
        ref = G(x, feats_=f)
    spec = nets.generator_spec(variant, cin, cout, rf, train=False)
    store = nets.ParamStore(spec, "cpu", seed=0)
    store.load_state_dict(G.state_dict())
    plan = nets.NetPlan(spec, store, B, T, L.F32, "cpu", train=False)
    plan.x.copy_(x)
    if f is not None:
        plan.feats.copy_(f)
    E.run_records(plan.prog.recs, *plan.prog.segments["pack"])
    E.run_records(plan.prog.recs, *plan.prog.segments["fwd"])
    assert rel_err(plan.out, ref) < 2e-5


@pytest.mark.parametrize("variant,rf,cin,cout,T", GEN_CASES)
def test_generator_train_forward_backward(variant, rf, cin, cout, T):
    torch.manual_seed(0)
    G = R.build_generator(variant, cin, cout, rf)
    randomize_bn(G)
    g = torch.Generator().manual_seed(1)
    B = 4
    x = torch.randn(B, cin, T, generator=g)
    y = torch.randn(B, cout, T, generator=g)
    f = feats_for(variant, rf, B, T, g)
    masks = R.make_masks(G, x, seed=3, feats=f)
    sd0 = {k: v.clone() for k, v in G.state_dict().items()}
    G.train()
    G.set_masks(masks)
    out = G(x, feats_=f)
    loss = torch.nn.functional.l1_loss(out, y)
    loss.backward()

    spec = nets.generator_spec(variant, cin, cout, rf, train=True)
    store = nets.ParamStore(spec, "cpu", seed=0)
    assert [k for k, _ in store.param_shapes] == [k for k, _ in G.named_parameters()]
    store.load_state_dict(sd0)
    plan = nets.NetPlan(spec, store, B, T, L.F32, "cpu", train=True, drop_mode="mask")
    plan.set_masks(masks)
    plan.x.copy_(x)
    if f is not None:
        plan.feats.copy_(f)
    recs = plan.prog.recs
    E.run_records(recs, *plan.prog.segments["pack"])
    E.run_records(recs, *plan.prog.segments["fwd"])
    assert rel_err(plan.out, out.detach()) < 2e-5
    # BN running statistics of every live layer
    live_bn = {l.bnkey for l in spec.layers if l.bn}
    for k, v in G.state_dict().items():
        base = k.rsplit(".", 1)[0]
        if base in live_bn and k.endswith(("running_mean", "running_var")):
            assert rel_err(store.b(k), v) < 2e-5, k
        if base in live_bn and k.endswith("num_batches_tracked"):
            assert int(store.nbt_view(k)[0]) == int(v), k
    # loss gradient -> backward program
    olb = plan.bufs[plan.out_layer.name]
    loss_buf = torch.zeros(1)
    E.l1(dict(out=plan.out, gt=y, dout=olb.dpre, loss=loss_buf, B=B, C=cout, L=T, ld=olb.Cp, Cfill=olb.Cp,
              gscale=1.0))
    assert abs(float(loss_buf[0]) - float(loss)) < 1e-5 * abs(float(loss))
    E.run_records(recs, *plan.prog.segments["bwd"])
    dead = {l.wkey for l in spec.dead} | {l.bnkey for l in spec.dead}
    for k, p in G.named_parameters():
        base = k.rsplit(".", 1)[0]
        if base in dead:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            assert float(store.g(k).abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        assert rel_err(store.g(k), p.grad) < 5e-5, k


@pytest.mark.parametrize("variant,rf", [("v1", False), ("v4", True), ("v2", True)])
def test_skip_connection_gradients_summed_in_the_dgrad_epilogue(monkeypatch, variant, rf):
    """bf16 train plans: the consumer of a skip connection that comes last in the backward pass adds the other
    consumer's input gradient in its dgrad epilogue (b2h_gemm_t.grad_add), the producer's bn_bwd reads one source and
    takes its first-pass sums from that dgrad.  Same gradients as the two-source plan up to one bf16 rounding of the
    summed gradient."""
    B, T, cin, cout = 4, 16, 36, 252
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cin, T, generator=g)
    y = torch.randn(B, cout, T, generator=g)
    f = feats_for(variant, rf, B, T, g)
    G = R.build_generator(variant, cin, cout, rf)
    randomize_bn(G)
    masks = R.make_masks(G, x, seed=3, feats=f)
    grads = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("B2H_NO_GRAD_ADD", "1")
        spec = nets.generator_spec(variant, cin, cout, rf, train=True)
        store = nets.ParamStore(spec, "cpu", seed=0)
        store.load_state_dict(G.state_dict())
        plan = nets.NetPlan(spec, store, B, T, L.BF16, "cpu", train=True, drop_mode="mask")
        bwd = plan.prog.recs[slice(*plan.prog.segments["bwd"])]
        n_add = sum(1 for r in bwd if r.kind == L.OP_GEMM and r.f.get("grad_add") is not None)
        two_src = [r.tag for r in bwd if r.kind == L.OP_BN_BWD and r.f["ngsrc"] == 2]
        if off:
            assert n_add == 0 and two_src
        else:
            assert n_add >= 1 and n_add == len(plan.grad_summed)
            for pname, c_last in plan.grad_summed.items():
                rec = next(r for r in bwd if r.tag == f"bn_bwd.{pname}")
                assert rec.f["ngsrc"] == 1 and rec.f["gsrc"][0]["g"] is plan.bufs[c_last.name].g
                if variant == "v1":     # conv5 / conv6: the last consumer's dgrad also carries the first-pass sums
                    assert rec.f["accum"] is not None
                    dg = next(r for r in bwd if r.tag == f"dgrad.{c_last.name}")
                    assert dg.f["bwd_sums"]["accum"] is rec.f["accum"]
        plan.set_masks(masks)
        plan.x.copy_(x)
        if f is not None:
            plan.feats.copy_(f)
        recs = plan.prog.recs
        E.run_records(recs, *plan.prog.segments["pack"])
        E.run_records(recs, *plan.prog.segments["fwd"])
        olb = plan.bufs[plan.out_layer.name]
        E.l1(dict(out=plan.out, gt=y, dout=olb.dpre, loss=torch.zeros(1), B=B, C=cout, L=T, ld=olb.Cp, Cfill=olb.Cp,
                  gscale=1.0))
        E.run_records(recs, *plan.prog.segments["bwd"])
        grads.append(store.grad.clone())
    assert rel_err(grads[0], grads[1]) < 2e-2
    assert float((grads[0] - grads[1]).abs().max()) > 0.0     # (the plans really differ: one more rounding)


@pytest.mark.parametrize("T,groups", [(16, 1), (16, 2), (64, 2), (192, 2), (21, 2)])
def test_discriminator(T, groups):
    torch.manual_seed(0)
    D = R.build_discriminator(252)
    randomize_bn(D)
    g = torch.Generator().manual_seed(1)
    Bg = 24   # small batches make the deep L=1 BatchNorm layers ill-conditioned (n = Bg samples)
    srcs = [torch.randn(Bg, 252, T, generator=g) for _ in range(groups)]
    sd0 = {k: v.clone() for k, v in D.state_dict().items()}
    spec = nets.discriminator_spec(252)
    # eval forward
    D.eval()
    store = nets.ParamStore(spec, "cpu", seed=0)
    store.load_state_dict(sd0)
    plan = nets.NetPlan(spec, store, Bg, T, L.F32, "cpu", train=False)
    plan.motion_src[0].copy_(srcs[0])
    E.run_records(plan.prog.recs, *plan.prog.segments["pack"])
    E.run_records(plan.prog.recs, *plan.prog.segments["fwd"])
    with torch.no_grad():
        ref = D(R.calc_motion(srcs[0]))
    Ld = ref.shape[2]
    assert rel_err(plan.out_blc[:, :, 0], ref[:, 0, :]) < 2e-5
    # train forward + backward on `groups` independent batches sharing the weights
    D.train()
    masks = [R.make_masks(D, R.calc_motion(s), seed=10 + i) for i, s in enumerate(srcs)]
    targets = [0.0, 1.0][:groups]
    scores = []
    loss = 0
    for i in range(groups):
        D.set_masks(masks[i])
        sc = D(R.calc_motion(srcs[i]))
        scores.append(sc)
        loss = loss + torch.nn.functional.mse_loss(sc, torch.full_like(sc, targets[i]))
    loss.backward()
    store = nets.ParamStore(spec, "cpu", seed=0)
    store.load_state_dict(sd0)
    plan = nets.NetPlan(spec, store, Bg * groups, T, L.F32, "cpu", train=True, groups=groups, drop_mode="mask")
    for i in range(groups):
        plan.set_masks(masks[i], group=i)
        plan.motion_src[i].copy_(srcs[i])
    recs = plan.prog.recs
    E.run_records(recs, *plan.prog.segments["pack"])
    E.run_records(recs, *plan.prog.segments["fwd"])
    ref_scores = torch.cat(scores, 0)
    assert rel_err(plan.out_blc[:, :, 0], ref_scores[:, 0, :].detach()) < 2e-5
    for k, v in D.state_dict().items():
        if k.endswith(("running_mean", "running_var")):
            assert rel_err(store.b(k), v) < 2e-5, k
        if k.endswith("num_batches_tracked"):
            assert int(store.nbt_view(k)[0]) == int(v), k
    # MSE loss + its gradient, then the backward program
    olb = plan.bufs[plan.out_layer.name]
    ld = plan.out_blc.shape[-1]
    dscore = torch.zeros_like(plan.out_blc)
    lbuf = torch.zeros(1)
    E.mse(dict(score=plan.out_blc, dscore=dscore, loss=lbuf, add=None, total=None, groups=groups, n=Bg * Ld, ld=ld,
               target=targets + [0.0] * (2 - groups)))
    assert abs(float(lbuf[0]) - float(loss)) < 1e-5 * abs(float(loss))
    E.prep(dict(src=dscore, out=olb.dpre, kind=E.SRC_ROWS, B=Bg * groups, L=Ld, C=1, ld=olb.Cp, Cfill=olb.Cp,
                src_ld=ld, drop=None, out_f32=0))
    E.run_records(recs, *plan.prog.segments["bwd"])
    for k, p in D.named_parameters():
        assert rel_err(store.g(k), p.grad) < 5e-5, k


@pytest.mark.parametrize("B,T", [(10, 38), (5, 10), (24, 64), (3, 21)])
def test_wgrad_workspace_covers_the_library_bound(B, T):
    """The split-K workspace a plan allocates (and declares in partial_bytes) is the library's own bound for every
    weight-gradient op, ragged tilings included (odd lengths make more k-blocks than rows / 64)."""
    import ctypes as C
    from b2h_b200.program import _fill_struct
    lib = L.load()
    for spec, groups in ((nets.discriminator_spec(252), 2), (nets.generator_spec("v1", 36, 252, False, train=True), 1)):
        for dtype in (L.F32, L.BF16):
            plan = nets.NetPlan(spec, nets.ParamStore(spec, "cpu", seed=0), B * groups, T, dtype, "cpu", train=True,
                                groups=groups, drop_mode="none")
            have = plan.wg_partial.numel() * 4
            for rec in plan.prog.recs:
                if rec.kind != L.OP_WGRAD:
                    continue
                assert rec.f["partial_bytes"] == have
                desc = _fill_struct(L.Wgrad(), {k: v for k, v in rec.f.items()
                                                if not k.startswith("_") and k not in ("P", "Q", "dW", "partial")})
                assert lib.b2h_wgrad_workspace_bytes(C.byref(desc), dtype) <= have, rec.tag


def test_batched_inference_plan_writes_ncl_from_the_output_layer(monkeypatch):
    """From NCL_DIRECT_MIN_ROWS frames per forward on, a bf16 eval plan has no to_ncl pass: the output layer's GEMM record
    carries out_f32 = 2 and the (B, C, T) fp32 tensor itself (include/b2h_abi.h).  The threshold is lowered so the
    interpreted plan stays small; v2's 48-channel output (Npad = 64) must keep the BLC + to_ncl form."""
    monkeypatch.setattr(nets, "NCL_DIRECT_MIN_ROWS", 32)
    for variant, rf, cout, direct in (("v1", False, 252, True), ("v2", True, 48, False)):
        torch.manual_seed(0)
        G = R.build_generator(variant, 36, cout, rf)
        randomize_bn(G)
        G.eval()
        g = torch.Generator().manual_seed(1)
        B, T = 3, 16
        x = torch.randn(B, 36, T, generator=g)
        f = feats_for(variant, rf, B, T, g)
        with torch.no_grad():
            ref = G(x, feats_=f)
        spec = nets.generator_spec(variant, 36, cout, rf, train=False)
        store = nets.ParamStore(spec, "cpu", seed=0)
        store.load_state_dict(G.state_dict())
        plan = nets.NetPlan(spec, store, B, T, L.BF16, "cpu", train=False)
        assert plan.ncl_direct is direct
        # the additions of the skip connections ride in the later producer's epilogue (b2h_gemm_t.resid), once with the
        # x2 up-sampling of that producer's rows, and the max-pooling in front of conv5 in the encoder's
        # (b2h_gemm_t.out_pool2): no bn_apply pass is left in the forward
        assert sorted(u for _, _, u in plan.eval_resid.values()) == [False, True] and list(plan.eval_pool) == ["encoder"]
        s0, e0 = plan.prog.segments["fwd"]
        assert not [r.tag for r in plan.prog.recs[s0:e0] if r.kind == L.OP_BN_APPLY]
        kinds = [r.kind for r in plan.prog.recs]
        assert (L.OP_TO_NCL in kinds) is (not direct)
        out_rec = [r for r in plan.prog.recs if r.kind == L.OP_GEMM and r.f["out_f32"]][-1]
        assert out_rec.f["out_f32"] == (2 if direct else 1)
        if direct:
            assert out_rec.f["out"] is plan.out and tuple(plan.out.shape) == (B, cout, T)
        plan.x.copy_(x)
        if f is not None:
            plan.feats.copy_(f)
        E.run_records(plan.prog.recs, *plan.prog.segments["pack"])
        E.run_records(plan.prog.recs, *plan.prog.segments["fwd"])
        assert rel_err(plan.out, ref) < 2e-2      # bf16 mode bar
    # train plans and fp32 plans never use it
    spec = nets.generator_spec("v1", 36, 252, False, train=False)
    store = nets.ParamStore(spec, "cpu", seed=0)
    assert not nets.NetPlan(spec, store, 3, 16, L.F32, "cpu", train=False).ncl_direct
