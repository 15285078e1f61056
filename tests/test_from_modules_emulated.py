"""The fused trainer over the parameters of drop-in modelZoo modules (GanTrainer.from_modules — what the kept
train_gan.py builds), on CPU: the modules' parameters / BatchNorm buffers ARE the trainer's flat buffers, so one emulated
generator step must show up in `generator.state_dict()`, match the oracle's step, and parameters loaded through the
modules' torch API (`load_state_dict`, as `--use_checkpoint` does) must reach the trainer and trigger a repack.
(The modules are bound to a CPU store through the test hook `_allow_cpu_store`; executing a program natively on CPU
still raises.)"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import modelZoo  # noqa: E402  (the drop-in at the repo root)
from b2h_b200 import _lib as L  # noqa: E402
from b2h_b200.trainer import GanTrainer  # noqa: E402
from oracle import ref_models as R  # noqa: E402
from tests.test_plan_emulated import feats_for, randomize_bn, rel_err  # noqa: E402
from tests.test_trainer_emulated import check_adam_params, emul_g_step, grads_close  # noqa: E402


CLASSES = {"v1": "regressor_fcn_bn_32", "b2h": "regressor_fcn_bn_32_b2h", "v2": "regressor_fcn_bn_32_v2",
           "v4": "regressor_fcn_bn_32_v4", "v4_deeper": "regressor_fcn_bn_32_v4_deeper"}


def _modules(cin=36, cout=252, variant="v1", rf=False):
    g = getattr(modelZoo, CLASSES[variant])()
    if variant == "b2h":
        g.build_net(cin, cout, require_image=rf)
    else:
        g.build_net(cin, cout, require_text=rf)
    d = modelZoo.regressor_fcn_bn_discriminator()
    d.build_net(cout)
    for m in (g, d):
        m._allow_cpu_store = True
    return g, d


@pytest.mark.parametrize("variant,rf", [("v1", False), ("v1", True), ("b2h", True), ("v2", True), ("v4", True),
                                        ("v4_deeper", False)])
def test_fused_step_updates_the_modules_and_matches_the_oracle(variant, rf):
    torch.manual_seed(0)
    B, T, cin, cout, lr = 8, 16, 36, 252, 1e-3
    Gm, Dm = _modules(cin, cout, variant, rf)
    G, D = R.build_generator(variant, cin, cout, rf), R.build_discriminator(cout)
    randomize_bn(G, 5)
    randomize_bn(D, 6)
    tr = GanTrainer.from_modules(Gm, Dm, batch_size=B, T=T, precision="fp32", lr=lr, drop_mode="mask")
    # the module tensors alias the flat buffers
    k = "decoder.9.weight"
    assert dict(Gm.named_parameters())[k].data_ptr() == tr.g_store.p(k).data_ptr()
    assert dict(Gm.named_buffers())["encoder.3.running_mean"].data_ptr() == tr.g_store.b("encoder.3.running_mean").data_ptr()
    # weights arrive through the torch API of the modules, as train_gan.py --use_checkpoint loads them
    v0 = (tr.g_store.version, tr.d_store.version)
    Gm.load_state_dict(G.state_dict(), strict=False)
    Dm.load_state_dict(D.state_dict(), strict=False)
    tr._sync_module_versions()
    assert tr.g_store.version > v0[0] and tr.d_store.version > v0[1]
    v1 = tr.g_store.version
    tr._sync_module_versions()
    assert tr.g_store.version == v1                   # nothing changed since: no further repack
    assert rel_err(tr.g_store.p(k), G.state_dict()[k]) == 0.0
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(B, cin, T, generator=g), torch.randn(B, cout, T, generator=g)
    f = feats_for(variant, rf, B, T, g)
    tr.load_batch(x, y, f)
    masks = R.make_masks(G, x, seed=100, feats=f)
    tr.G_train.set_masks(masks)
    g_opt = torch.optim.Adam(G.parameters(), lr=lr)
    g_loss, l1, adv, out = R.generator_step(G, D, g_opt, x, y, f, masks)
    emul_g_step(tr)
    assert rel_err(tr.G_train.out, out) < 3e-5
    assert abs(float(tr.losses[2]) - float(g_loss)) < 1e-4 * abs(float(g_loss))
    sd = Gm.state_dict()                                # what train_gan.py checkpoints
    for name, p in G.named_parameters():
        if p.grad is not None:
            assert grads_close(tr.g_store.g(name), p.grad, 0.2), name
        check_adam_params(sd[name], p, lr, name, tight=False)
        assert float((sd[name] - tr.g_store.p(name)).abs().max()) == 0.0
    for name, b in G.state_dict().items():
        if name.endswith(("running_mean", "running_var")) and name.rsplit(".", 1)[0] in {l.bnkey for l in tr.g_spec.layers}:
            assert rel_err(sd[name], b) < 5e-5, name
    assert int(sd["encoder.3.num_batches_tracked"]) == int(G.state_dict()["encoder.3.num_batches_tracked"])
    # the optimizer object train_gan.py checkpoints has torch.optim.Adam's layout
    osd = tr.g_opt.state_dict()
    assert osd["param_groups"][0]["params"] == g_opt.state_dict()["param_groups"][0]["params"]
    assert float(osd["state"][0]["step"]) == 1.0


def test_programs_still_refuse_to_run_on_cpu():
    Gm, _ = _modules()
    Gm.eval()
    with pytest.raises(L.B2HError):
        Gm(torch.randn(2, 36, 16))


def test_train_generator_loop_over_an_emulated_trainer(capsys):
    """The kept train_gan.train_generator / train_discriminator loops driving a module-backed trainer whose steps are
    interpreted on CPU (dropout off): three batches of each, against the oracle's step bodies on the same batches."""
    import argparse
    os.environ.setdefault("WANDB_MODE", "disabled")
    import train_gan
    from tests.test_trainer_emulated import emul_d_step
    torch.manual_seed(0)
    B, T, cin, cout, lr = 8, 16, 36, 252, 1e-3
    Gm, Dm = _modules(cin, cout)
    G, D = R.build_generator("v1", cin, cout), R.build_discriminator(cout)
    randomize_bn(G, 5)
    randomize_bn(D, 6)
    for m in (G, D):
        for mod in m.modules():
            if isinstance(mod, R.ReplayDropout):
                mod.p = 0.0
    tr = GanTrainer.from_modules(Gm, Dm, batch_size=B, T=T, precision="fp32", lr=lr, drop_mode="none")
    Gm.load_state_dict(G.state_dict(), strict=False)
    Dm.load_state_dict(D.state_dict(), strict=False)
    tr.generator_step = lambda graph=False: emul_g_step(tr)
    tr.discriminator_step = lambda graph=False: emul_d_step(tr)
    g = torch.Generator().manual_seed(1)
    X = torch.randn(3 * B + 3, cin, T, generator=g).numpy()
    Y = torch.randn(3 * B + 3, cout, T, generator=g).numpy()
    args = argparse.Namespace(batch_size=B, num_epochs=2, log_step=1, disc_label_smooth=False)
    g_opt, d_opt = torch.optim.Adam(G.parameters(), lr=lr), torch.optim.Adam(D.parameters(), lr=lr)
    train_gan.train_generator(args, Gm, Dm, None, None, tr.g_opt, X, Y, 0, trainer=tr)
    out = capsys.readouterr().out
    ref_losses = []
    for i in range(3):
        xs, ys = torch.from_numpy(X[i * B:(i + 1) * B]), torch.from_numpy(Y[i * B:(i + 1) * B])
        ref_losses.append(float(R.generator_step(G, D, g_opt, xs, ys)[0]))
    m = sum(l * B for l in ref_losses) / (3 * B)
    assert "Epoch [0/1], Tr. Loss: {:.4f}".format(m) in out
    assert float(tr.g_opt.step) == 3.0
    sd = Gm.state_dict()
    for name, p in G.named_parameters():
        assert float((sd[name] - p.detach()).abs().max()) <= 3 * 2.05 * lr, name     # three Adam steps
    # (Adam divides by sqrt(v) + eps: where a gradient is ~0, fp32 noise decides the direction of a step — hence the
    # step-size bound above; the typical element agrees far better)
    typical = float((sd["decoder.9.weight"] - G.state_dict()["decoder.9.weight"]).abs().median())
    assert typical < 0.02 * lr, typical
    train_gan.train_discriminator(args, Gm, Dm, None, tr.d_opt, X, Y, 1, trainer=tr)
    out = capsys.readouterr().out
    G.load_state_dict({k: v for k, v in sd.items() if k in G.state_dict()})       # same generator on both sides
    ref_d = [float(R.discriminator_step(G, D, d_opt, torch.from_numpy(X[i * B:(i + 1) * B]),
                                        torch.from_numpy(Y[i * B:(i + 1) * B]))[0]) for i in range(3)]
    shown = float(out.split("Tr. Disc. Loss:")[1].split()[0])
    assert abs(shown - sum(ref_d) / 3) < 2e-2 * abs(shown)
    assert float(tr.d_opt.step) == 3.0


def test_checkpoint_round_trip_through_the_modules(tmp_path):
    """What train_gan.py writes on a validation improvement (train_gan.py:353-370) and reads back under
    --use_checkpoint (:70-73): generator.state_dict() + the optimizer's torch-format state, through torch.save /
    torch.load, into fresh modules and a fresh trainer — parameters, BatchNorm buffers, Adam moments and step."""
    torch.manual_seed(0)
    B, T, cin, cout, lr = 8, 16, 36, 252, 1e-3
    Gm, Dm = _modules(cin, cout)
    tr = GanTrainer.from_modules(Gm, Dm, batch_size=B, T=T, precision="fp32", lr=lr, drop_mode="none")
    g = torch.Generator().manual_seed(1)
    tr.load_batch(torch.randn(B, cin, T, generator=g), torch.randn(B, cout, T, generator=g))
    for _ in range(2):
        emul_g_step(tr)
    path = tmp_path / "experiment_checkpoint.pth"
    torch.save({"epoch": 7, "state_dict": Gm.state_dict(), "g_optimizer": tr.g_opt.state_dict()}, path)
    G2, D2 = _modules(cin, cout)
    tr2 = GanTrainer.from_modules(G2, D2, batch_size=B, T=T, precision="fp32", lr=lr, drop_mode="none")
    assert not torch.equal(tr2.g_store.flat, tr.g_store.flat)
    ck = torch.load(path, map_location="cpu")
    G2.load_state_dict(ck["state_dict"], strict=False)
    tr2.g_opt.load_state_dict(ck["g_optimizer"])
    assert ck["epoch"] == 7
    assert torch.equal(tr2.g_store.flat, tr.g_store.flat)
    assert torch.equal(tr2.g_store.bufs, tr.g_store.bufs) and torch.equal(tr2.g_store.nbt, tr.g_store.nbt)
    assert torch.equal(tr2.g_opt.m, tr.g_opt.m) and torch.equal(tr2.g_opt.v, tr.g_opt.v)
    assert int(tr2.g_opt.step) == int(tr.g_opt.step) == 2
    # the reference's own optimizer class accepts the same dictionary
    Gr = R.build_generator("v1", cin, cout)
    ref_opt = torch.optim.Adam(Gr.parameters(), lr=lr)
    ref_opt.load_state_dict(ck["g_optimizer"])
    assert float(ref_opt.state_dict()["state"][0]["step"]) == 2.0
    # and the next step of the restored trainer equals the next step of the original
    tr2.load_batch(tr.x, tr.y)
    emul_g_step(tr)
    emul_g_step(tr2)
    assert torch.equal(tr2.g_store.flat, tr.g_store.flat)
