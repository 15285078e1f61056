"""Whole GAN step (generator step, then discriminator step) interpreted on CPU through the op
restatements, against the oracle's train_gan restatement with torch.optim.Adam."""
import pytest
import torch

import b2h_b200  # noqa: F401
from b2h_b200 import _lib as L
from b2h_b200.trainer import GanTrainer
from oracle import ops_emul as E
from oracle import ref_models as R
from tests.test_plan_emulated import feats_for, randomize_bn, rel_err


def grads_close(ours, ref, tol):
    """Max-norm agreement, except for the rare activation-kink flips: a pre-activation within fp32
    noise of 0 (ReLU / LeakyReLU / L1 sign) changes ONE mask element, which moves a few entries of a
    few gradients by percents.  Those must stay isolated (< 0.2 % of entries, small in L2)."""
    if rel_err(ours, ref) < tol:
        return True
    d = (ours - ref).abs()
    frac_bad = float((d > tol * ref.abs().max()).float().mean())
    rel_l2 = float(d.norm() / (ref.norm() + 1e-30))
    return frac_bad < 2e-3 and rel_l2 < 5e-2


def activation_hooks(model):
    """Record the output of every LeakyReLU / ReLU of an oracle model (keyed by module name)."""
    acts = {}
    hs = []
    for name, m in model.named_modules():
        if isinstance(m, (torch.nn.LeakyReLU, torch.nn.ReLU)):
            hs.append(m.register_forward_hook(lambda mod, i, o, name=name: acts.__setitem__(name, o.detach().clone())))
    return acts, hs


def count_kink_flips(plan, acts, groups_slice=None):
    """Number of activation-sign disagreements between a plan's stored z tensors and the oracle."""
    flips = 0
    for l in plan.spec.layers:
        key = f"{l.seq}.{l.w_idx + 1}"
        if key not in acts or not l.bn:
            continue
        z = plan.bufs[l.name].z[:, :, :l.cout].float().cpu()
        ref = acts[key]
        ref = ref.permute(0, 2, 1) if ref.dim() == 3 else ref.reshape(z.shape[0], -1, l.cout)
        if groups_slice is not None:
            z = z[groups_slice]
        flips += int(((z > 0) != (ref > 0)).sum())
    return flips


def check_adam_params(ours, ref_param, lr, what, tight=True):
    """Adam divides by sqrt(v) + 1e-8: where the gradient is ~0 the update amplifies fp32 noise, so
    only elements with a clear gradient are held to a tight bound; all are bounded by the step size."""
    diff = (ours - ref_param.detach()).abs()
    assert float(diff.max()) <= 2.05 * lr, what
    if tight and ref_param.grad is not None:
        clear = ref_param.grad.abs() > max(1e-6, 1e-2 * float(ref_param.grad.abs().max()))
        if clear.any():
            assert float(diff[clear].max()) < 5e-3 * lr, what


def emul(prog, seg):
    E.run_records(prog.recs, *prog.segments[seg])


def emul_g_forward_and_losses(tr):
    """Generator forward, regression loss (it writes G_train.out from the output layer's BLC tile), scoring pass of
    the discriminator on that output, adversarial term -- the order of GanTrainer._g_ops."""
    # (the eval plans read the packed weights of their train twins: pack those first)
    for p, s in ((tr.G_train.prog, "pack"), (tr.D_train.prog, "pack"), (tr.D_eval.prog, "pack"), (tr.G_train.prog, "fwd")):
        emul(p, s)
    ls, le = tr.g_loss_prog.segments["loss"]      # [l1, adv]
    E.run_records(tr.g_loss_prog.recs, ls, ls + 1)
    emul(tr.D_eval.prog, "fwd")
    E.run_records(tr.g_loss_prog.recs, ls + 1, le)


def emul_g_step(tr):
    emul_g_forward_and_losses(tr)
    for p, s in ((tr.G_train.prog, "bwd"), (tr.g_loss_prog, "opt")):
        emul(p, s)


def emul_d_step(tr):
    tr._sync_d_batch()   # the discriminator step reads its own copy of the batch (xd / yd)
    for p, s in ((tr.G_train.prog, "pack"), (tr.G_eval.prog, "pack"), (tr.D_train.prog, "pack"),
                 (tr.G_eval.prog, "fwd"), (tr.D_train.prog, "fwd"),
                 (tr.d_loss_prog, "loss"), (tr.D_train.prog, "bwd"), (tr.d_loss_prog, "opt")):
        emul(p, s)


@pytest.mark.parametrize("variant,rf,label_smooth", [("v1", False, False), ("v1", True, True), ("b2h", True, False)])
def test_gan_steps_match_oracle(variant, rf, label_smooth):
    torch.manual_seed(0)
    B, T, cin, cout, lr = 16, 16, 36, 252, 1e-3
    G = R.build_generator(variant, cin, cout, rf)
    D = R.build_discriminator(cout)
    randomize_bn(G, 5)
    randomize_bn(D, 6)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cin, T, generator=g)
    y = torch.randn(B, cout, T, generator=g)
    f = feats_for(variant, rf, B, T, g)
    tr = GanTrainer(variant, cin, cout, rf, B, T, precision="fp32", device="cpu", lr=lr, drop_mode="mask",
                    label_smooth=label_smooth)
    tr.g_store.load_state_dict(G.state_dict())
    tr.d_store.load_state_dict(D.state_dict())
    tr.x.copy_(x)
    tr.y.copy_(y)
    if f is not None:
        tr.feats.copy_(f)
    g_opt = torch.optim.Adam(G.parameters(), lr=lr)
    d_opt = torch.optim.Adam(D.parameters(), lr=lr)
    g_acts, _ = activation_hooks(G)
    any_flips = 0
    for it in range(2):
        if it > 0:
            # Adam amplifies fp32 noise where gradients vanish (see check_adam_params): restart every
            # iteration from the oracle's exact state, through the checkpoint-format loaders
            tr.g_store.load_state_dict(G.state_dict())
            tr.d_store.load_state_dict(D.state_dict())
            tr.g_opt.load_state_dict(g_opt.state_dict())
            tr.d_opt.load_state_dict(d_opt.state_dict())
        # ---- generator step
        g_masks = R.make_masks(G, x, seed=100 + it, feats=f)
        tr.G_train.set_masks(g_masks)
        g_loss, l1, adv, out = R.generator_step(G, D, g_opt, x, y, f, g_masks)
        emul_g_step(tr)
        assert rel_err(tr.G_train.out, out) < 3e-5
        assert abs(float(tr.losses[0]) - float(l1)) < 2e-5 * abs(float(l1))
        assert abs(float(tr.losses[1]) - float(adv)) < 2e-4 * abs(float(adv)) + 1e-6
        assert abs(float(tr.losses[2]) - float(g_loss)) < 1e-4 * abs(float(g_loss))
        # a pre-activation within fp32 noise of 0 makes the gradient discontinuous (different ReLU /
        # L1-sign branch on the two sides): such an iteration can only be compared loosely
        flips = count_kink_flips(tr.G_train, g_acts)
        flips += int((torch.sign(tr.G_train.out - y) != torch.sign(out - y)).sum())
        gtol = 5e-5 if flips == 0 else 0.2
        any_flips += flips
        for k, p in G.named_parameters():
            if p.grad is not None:
                assert grads_close(tr.g_store.g(k), p.grad, gtol), (it, k, flips)
            # Adam divides by |g| + 1e-8: elements whose gradient is ~0 amplify fp32 noise, so the
            # parameters are compared against the step size (exact Adam parity: test_gpu_replay / ops)
            check_adam_params(tr.g_store.p(k), p, lr, (it, k), tight=flips == 0)
        # ---- discriminator step
        D.train()
        with torch.no_grad():
            G.eval()
            fake = G(x, feats_=f)
        mf = R.make_masks(D, R.calc_motion(fake), seed=200 + it)
        mr = R.make_masks(D, R.calc_motion(y), seed=300 + it)
        tr.D_train.set_masks(mf, group=0)
        tr.D_train.set_masks(mr, group=1)
        tr.g_store.load_state_dict(G.state_dict())
        d_loss, fs, rs = R.discriminator_step(G, D, d_opt, x, y, f, mf, mr, label_smooth)
        emul_d_step(tr)
        assert abs(float(tr.losses[3]) - float(d_loss)) < 1e-4 * abs(float(d_loss))
        for k, p in D.named_parameters():
            assert grads_close(tr.d_store.g(k), p.grad, 5e-4), (it, k)  # L=1 BN layers over 16 samples: ill-conditioned
            check_adam_params(tr.d_store.p(k), p, lr, (it, k))
        for k, v in D.state_dict().items():
            if k.endswith(("running_mean", "running_var")):
                assert rel_err(tr.d_store.b(k), v) < 5e-5, (it, k)
    # optimizer state in torch.optim.Adam's format
    sd = tr.g_opt.state_dict()
    ref_sd = g_opt.state_dict()
    assert sd["param_groups"][0]["params"] == ref_sd["param_groups"][0]["params"]
    for i, s in ref_sd["state"].items():
        stol = 1e-4 if any_flips == 0 else 0.1
        assert rel_err(sd["state"][i]["exp_avg"], s["exp_avg"]) < stol
        assert rel_err(sd["state"][i]["exp_avg_sq"], s["exp_avg_sq"]) < stol
        assert float(sd["state"][i]["step"]) == float(s["step"]) == 2.0


@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_gan_steps_match_oracle_with_deferred_bn_backward(monkeypatch, mode):
    """The opt-in backward variants (b2h_bn_bwd_t.defer = 1 / 2 + b2h_colsum_t.bn_accum, b2h_bn_bwd_t.first_pass_only
    for the skip connections) compute the same step."""
    monkeypatch.setenv("B2H_DEFER_BN", mode)
    monkeypatch.setenv("B2H_BWD_HELPERS", "1")
    tr = GanTrainer("v1", 36, 252, False, 4, 16, precision="fp32", device="cpu", drop_mode="mask")
    tags = [r.tag for r in tr.G_train.prog.recs]
    assert "bwd_sums1.conv5.skip5" in tags and "bwd_sums1.conv6.skip4" in tags
    rec = {r.tag: r for r in tr.G_train.prog.recs}
    assert rec["bn_bwd.conv5"].f["defer"] == int(mode) and rec["bn_bwd.conv5"].f["_wait_tags"] == ["bwd_sums1.conv5.skip5"]
    assert rec["bn_bwd.encoder"].f["defer"] == int(mode)  # (pooled gradient source: sums from conv5's dgrad epilogue)
    assert ("bn_fin.conv5" in rec) == (mode != "0") and (mode == "0" or (rec["bn_fin.conv5"].f["src"] is None) == (mode == "2"))
    assert rec["dgrad.conv5"].f["bwd_sums"]["rowmap"] == L.ROW_POOL2
    test_gan_steps_match_oracle("v1", False, False)


def test_bucketed_optimizer_programs_equal_the_whole_update():
    """The optimizer step split along the gradient buckets (Adam phase 1 once, phase 2 per flat range, repack per
    bucket) must equal the single Adam op + the single repack: ranges tile the flat buffer, bias corrections are
    advanced exactly once, every packed operand is rewritten by exactly one bucket."""
    torch.manual_seed(3)
    B, T = 4, 16
    outs = []
    for bucketed in (False, True):
        tr = GanTrainer("v1", 36, 252, False, B, T, precision="fp32", device="cpu", lr=1e-3, drop_mode="none",
                        n_buckets=3)
        g = torch.Generator().manual_seed(5)
        for key, store, opt, plan, loss_prog in (("g", tr.g_store, tr.g_opt, tr.G_train, tr.g_loss_prog),
                                                 ("d", tr.d_store, tr.d_opt, tr.D_train, tr.d_loss_prog)):
            store.grad.copy_(torch.randn(store.n, generator=g) * 1e-2)
            opt.m.copy_(torch.randn(store.n, generator=g) * 1e-3)
            opt.v.copy_(torch.rand(store.n, generator=g) * 1e-4)
            opt.step.fill_(7)
            bp, P, packs = tr._buckets[key]
            # the buckets tile [0, n) exactly once
            assert bp[0][3] == store.n and bp[-1][2] == 0 and all(a[2] == b[3] for a, b in zip(bp, bp[1:]))
            if bucketed:
                E.run_records(P.recs, *P.segments["step"])
                for i in range(len(bp)):
                    E.run_records(P.recs, *P.segments[f"b{i}"])
                    E.run_records(plan.prog.recs, *plan.prog.segments[packs[i]])
            else:
                E.run_records(loss_prog.recs, *loss_prog.segments["opt"])
                E.run_records(plan.prog.recs, *plan.prog.segments["pack"])
            assert int(opt.step) == 8
        outs.append({"g": tr.g_store.flat.clone(), "d": tr.d_store.flat.clone(),
                     "gm": tr.g_opt.m.clone(), "gv": tr.g_opt.v.clone(),
                     "wf": {n: b.wf.clone() for n, b in tr.G_train.bufs.items() if b.wf is not None},
                     "wb": {n: b.wb.clone() for n, b in tr.D_train.bufs.items() if b.wb is not None},
                     "bias": {n: b.bias.clone() for n, b in tr.D_train.bufs.items() if b.bias is not None}})
    a, b = outs
    for k in ("g", "d", "gm", "gv"):
        assert torch.equal(a[k], b[k]), k
    for k in ("wf", "wb", "bias"):
        assert a[k].keys() == b[k].keys() and len(a[k]) > 0
        for n in a[k]:
            assert torch.equal(a[k][n], b[k][n]), (k, n)


@pytest.mark.parametrize("loss", ["L2", "Huber1", "RobustLoss"])
def test_generator_step_with_other_regression_losses(loss):
    """--loss {L2, Huber1, RobustLoss} (utils/constants.py:53-58): the b2h_l1 op's other kinds, through one emulated
    generator step against the oracle (loss value, every gradient, Adam update)."""
    torch.manual_seed(0)
    B, T, cin, cout, lr = 8, 16, 36, 252, 1e-3
    G = R.build_generator("v1", cin, cout, False)
    D = R.build_discriminator(cout)
    randomize_bn(G, 5)
    randomize_bn(D, 6)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, cin, T, generator=g)
    y = torch.randn(B, cout, T, generator=g) * 1.5   # |out - y| on both sides of Huber's delta = 1
    tr = GanTrainer("v1", cin, cout, False, B, T, precision="fp32", device="cpu", lr=lr, drop_mode="mask", loss=loss)
    tr.g_store.load_state_dict(G.state_dict())
    tr.d_store.load_state_dict(D.state_dict())
    tr.x.copy_(x)
    tr.y.copy_(y)
    g_opt = torch.optim.Adam(G.parameters(), lr=lr)
    g_acts, _ = activation_hooks(G)
    g_masks = R.make_masks(G, x, seed=100)
    tr.G_train.set_masks(g_masks)
    g_loss, reg, adv, out = R.generator_step(G, D, g_opt, x, y, None, g_masks, loss=loss)
    emul_g_step(tr)
    assert ((out - y).abs() > 1).any() and ((out - y).abs() < 1).any()
    assert abs(float(tr.losses[0]) - float(reg)) < 2e-5 * abs(float(reg))
    assert abs(float(tr.losses[2]) - float(g_loss)) < 1e-4 * abs(float(g_loss))
    flips = count_kink_flips(tr.G_train, g_acts)
    gtol = 5e-5 if flips == 0 else 0.2
    for k, p in G.named_parameters():
        if p.grad is not None:
            assert grads_close(tr.g_store.g(k), p.grad, gtol), (k, flips)
        check_adam_params(tr.g_store.p(k), p, lr, k, tight=flips == 0)


def test_unknown_loss_is_rejected():
    with pytest.raises(KeyError):
        GanTrainer("v1", 36, 252, False, 4, 16, precision="fp32", device="cpu", loss="L3")


def test_odd_window_length_is_rejected_like_the_reference():
    """modelZoo's generator returns 2 * floor(T / 2) frames: the reference's L1Loss(output, outputGT) fails for odd T
    (shape mismatch); the fused trainer says so at construction."""
    G = R.build_generator("v1", 36, 252)
    x, y = torch.randn(2, 36, 7), torch.randn(2, 252, 7)
    G.eval()
    with torch.no_grad():
        assert G(x).shape[-1] == 6
    with pytest.raises(RuntimeError):
        torch.nn.L1Loss()(G(x), y)
    with pytest.raises(ValueError):
        GanTrainer("v1", 36, 252, False, 2, 7, precision="fp32", device="cpu")


@pytest.mark.parametrize("variant,rf,precision", [("v4", True, "fp32"), ("v1", True, "fp32"), ("b2h", True, "fp32"),
                                                  ("v4_deeper", True, "fp32"), ("v4", True, "bf16"), ("v1", False, "bf16")])
def test_bucket_by_bucket_update_order_matches_the_oracle(variant, rf, precision):
    """The optimizer step in the order GanTrainer._bwd_update issues it — backward ops of bucket i, then the Adam range
    and the repack of bucket i, while later buckets' gradients do not exist yet — against the oracle's step.  The flat
    parameter order is the reference's module registration order, not the backward order (v4 registers the text branch
    first and uses it at the bottleneck): a bucket may only update parameters whose gradients are complete, and must
    repack exactly the layers it updated (regression: v4 + text updated conv5 / encoder one bucket too early)."""
    torch.manual_seed(0)
    B, T, cin, cout, lr = 16, 16, 36, 252, 1e-3
    G = R.build_generator(variant, cin, cout, rf)
    D = R.build_discriminator(cout)
    randomize_bn(G, 5)
    randomize_bn(D, 6)
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(B, cin, T, generator=g), torch.randn(B, cout, T, generator=g)
    f = feats_for(variant, rf, B, T, g)
    tr = GanTrainer(variant, cin, cout, rf, B, T, precision=precision, device="cpu", lr=lr, drop_mode="mask", n_buckets=3)
    tr.g_store.load_state_dict(G.state_dict())
    tr.d_store.load_state_dict(D.state_dict())
    tr.x.copy_(x)
    tr.y.copy_(y)
    if f is not None:
        tr.feats.copy_(f)
    g_opt = torch.optim.Adam(G.parameters(), lr=lr)
    masks = R.make_masks(G, x, seed=100, feats=f)
    tr.G_train.set_masks(masks)
    bp, P, packs = tr._buckets["g"]        # (bf16: the plans with fused statistics / backward sums / eval-BN epilogues)
    st = tr.g_store
    # ranges tile the buffer; every layer with parameters is repacked by exactly the bucket that updates it
    assert bp[0][3] == st.n and bp[-1][2] == 0 and all(a[2] == b[3] for a, b in zip(bp, bp[1:]))
    all_updated = [n for *_, names in bp for n in names]
    assert len(all_updated) == len(set(all_updated)) == len({l.name for l in st.spec.all_layers()})
    for steps in range(2):
        R.generator_step(G, D, g_opt, x, y, f, masks)
        st.grad.fill_(float("nan"))                  # a gradient consumed before it is produced poisons the update
        live = {n for n, _ in tr.G_train.bwd_marks}
        for l in st.spec.all_layers():               # (dead branches never write theirs: they stay zero, as on the device)
            if l.name not in live:
                for key in [l.wkey + ".weight", l.wkey + ".bias"] + ([l.bnkey + ".weight", l.bnkey + ".bias"] if l.bn else []):
                    st.g(key).zero_()
        emul_g_forward_and_losses(tr)                # (the loss op produces the output layer's bias gradient)
        E.run_records(P.recs, *P.segments["step"])
        for i, (s, e, lo, hi, _) in enumerate(bp):
            E.run_records(tr.G_train.prog.recs, s, e)
            assert torch.isfinite(st.grad[lo:hi]).all(), (i, lo, hi)
            E.run_records(P.recs, *P.segments[f"b{i}"])
            E.run_records(tr.G_train.prog.recs, *tr.G_train.prog.segments[packs[i]])
        assert torch.isfinite(st.flat).all()
        for k, p in G.named_parameters():
            check_adam_params(st.p(k), p, lr, (steps, k), tight=False)
        # the packed operands equal a full repack of the updated parameters
        snap = {n: b.wf.clone() for n, b in tr.G_train.bufs.items() if b.wf is not None}
        emul(tr.G_train.prog, "pack")
        for n, w in snap.items():
            assert torch.equal(w, tr.G_train.bufs[n].wf), n
        tr.g_store.load_state_dict(G.state_dict())   # restart the next iteration from the oracle's exact state
        tr.g_opt.load_state_dict(g_opt.state_dict())


def test_optimizer_checkpoint_hyperparameters_reach_the_recorded_ops():
    """torch.optim.Adam.load_state_dict applies the checkpoint's lr / betas / eps (train_gan.py:72,91); the recorded
    Adam ops carry them in their descriptors, so FlatAdam.load_state_dict must patch every record."""
    torch.manual_seed(0)
    B, T = 4, 16
    tr = GanTrainer("v1", 36, 252, False, B, T, precision="fp32", device="cpu", lr=1e-4, drop_mode="none")
    ref = torch.optim.Adam([torch.nn.Parameter(torch.zeros(s)) for _, s in tr.g_store.param_shapes], lr=3e-3,
                           betas=(0.8, 0.99), eps=1e-6)
    sd = ref.state_dict()
    dropped = []
    tr.g_opt.on_hparams_changed.append(lambda: dropped.append(1))
    tr.g_opt.load_state_dict(sd)
    assert (tr.g_opt.lr, tr.g_opt.betas, tr.g_opt.eps) == (3e-3, (0.8, 0.99), 1e-6) and dropped
    recs = [prog.recs[i].f for prog, i in tr.g_opt._recorded]
    assert len(recs) >= 2 + tr.n_buckets
    for f in recs:
        assert f["lr"] == 3e-3 and (f["beta1"], f["beta2"], f["eps"]) == (0.8, 0.99, 1e-6)
    # and the interpreted update really uses them: one step from zero moments moves every parameter by ~lr
    tr.g_store.grad.fill_(1.0)
    p0 = tr.g_store.flat.clone()
    emul(tr.g_loss_prog, "opt")
    assert torch.allclose(p0 - tr.g_store.flat, torch.full_like(p0, 3e-3), rtol=1e-3)
    # the discriminator's optimizer kept its constructor values
    assert all(prog.recs[i].f["lr"] == 1e-4 for prog, i in tr.d_opt._recorded)
