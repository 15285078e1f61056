"""GPU parity, kernel by kernel: every op of the recorded generator / discriminator programs is run
through the C ABI on the B200 and compared with its CPU restatement on identical inputs."""
import os

import pytest
import torch

import b2h_b200  # noqa: F401
from b2h_b200 import _lib as L
from b2h_b200 import nets
from oracle import ref_models as R
from tests.replay_util import format_report, replay_pair, sync_inputs
from tests.test_plan_emulated import feats_for, randomize_bn

pytestmark = pytest.mark.gpu

# fp32 mode: FFMA accumulation order differs from the CPU restatement only; bf16 mode: both sides
# round stored activations to bf16, accumulation-order differences flip single bf16 ulps (2^-8)
TOL = {L.F32: 3e-5, L.BF16: 1.2e-2}


def _log(name, text):
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "replay_report.txt"), "a") as fh:
        fh.write(f"== {name}\n{text}\n")


def _check(name, report, tol):
    txt = format_report(report)
    _log(name, txt)
    bad = [r for r in report if not (r[3] <= tol)]
    assert not bad, f"{name}: {len(bad)} op outputs above tol {tol}:\n" + format_report(bad)


GEN = [("v1", False, 36, 252, 4, 64), ("v1", True, 36, 252, 4, 32), ("b2h", True, 36, 252, 2, 16),
       ("v4", True, 42, 246, 4, 32), ("v1", False, 162, 126, 3, 24), ("v1", False, 36, 252, 5, 14),
       ("v1", False, 36, 252, 2, 192)]


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
@pytest.mark.parametrize("variant,rf,cin,cout,B,T", GEN)
def test_generator_train_replay(variant, rf, cin, cout, B, T, dtype):
    torch.manual_seed(0)
    G = R.build_generator(variant, cin, cout, rf)
    randomize_bn(G)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cin, T, generator=g)
    y = torch.randn(B, cout, T, generator=g)
    f = feats_for(variant, rf, B, T, g)
    masks = R.make_masks(G, x, seed=3, feats=f)
    spec_c = nets.generator_spec(variant, cin, cout, rf, train=True)
    spec_g = nets.generator_spec(variant, cin, cout, rf, train=True)
    st_c = nets.ParamStore(spec_c, "cpu", seed=0)
    st_g = nets.ParamStore(spec_g, "cuda", seed=0)
    st_c.load_state_dict(G.state_dict())
    pc = nets.NetPlan(spec_c, st_c, B, T, dtype, "cpu", train=True, drop_mode="mask", wgrad_direct=True)
    pg = nets.NetPlan(spec_g, st_g, B, T, dtype, "cuda", train=True, drop_mode="mask", wgrad_direct=True)
    pc.set_masks(masks)
    pc.x.copy_(x)
    if f is not None:
        pc.feats.copy_(f)
    olb = pc.bufs[pc.out_layer.name]
    olb.dpre[:, :, :cout] = (torch.sign(torch.randn(B, T, cout, generator=g)) / (B * T * cout)).to(olb.dpre.dtype)
    sync_inputs(pg.prog, pc.prog)
    rep = replay_pair(pg.prog, pc.prog, ["pack", "fwd", "bwd"])
    _check(f"gen-train {variant} feats={rf} {cin}->{cout} B={B} T={T} dtype={dtype}", rep, TOL[dtype])


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
@pytest.mark.parametrize("variant,rf,cin,cout,B,T", GEN[:4])
def test_generator_eval_replay(variant, rf, cin, cout, B, T, dtype):
    torch.manual_seed(0)
    G = R.build_generator(variant, cin, cout, rf)
    randomize_bn(G)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cin, T, generator=g)
    f = feats_for(variant, rf, B, T, g)
    spec_c = nets.generator_spec(variant, cin, cout, rf, train=False)
    spec_g = nets.generator_spec(variant, cin, cout, rf, train=False)
    st_c = nets.ParamStore(spec_c, "cpu", seed=0)
    st_g = nets.ParamStore(spec_g, "cuda", seed=0)
    st_c.load_state_dict(G.state_dict())
    pc = nets.NetPlan(spec_c, st_c, B, T, dtype, "cpu", train=False)
    pg = nets.NetPlan(spec_g, st_g, B, T, dtype, "cuda", train=False)
    pc.x.copy_(x)
    if f is not None:
        pc.feats.copy_(f)
    sync_inputs(pg.prog, pc.prog)
    rep = replay_pair(pg.prog, pc.prog, ["pack", "fwd"])
    _check(f"gen-eval {variant} feats={rf} B={B} T={T} dtype={dtype}", rep, TOL[dtype])


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
@pytest.mark.parametrize("Bg,T,groups", [(24, 64, 2), (8, 192, 2), (16, 21, 1)])
def test_discriminator_train_replay(Bg, T, groups, dtype):
    torch.manual_seed(0)
    D = R.build_discriminator(252)
    randomize_bn(D)
    g = torch.Generator().manual_seed(1)
    srcs = [torch.randn(Bg, 252, T, generator=g) for _ in range(groups)]
    masks = [R.make_masks(D, R.calc_motion(s), seed=10 + i) for i, s in enumerate(srcs)]
    spec_c, spec_g = nets.discriminator_spec(252), nets.discriminator_spec(252)
    st_c = nets.ParamStore(spec_c, "cpu", seed=0)
    st_g = nets.ParamStore(spec_g, "cuda", seed=0)
    st_c.load_state_dict(D.state_dict())
    pc = nets.NetPlan(spec_c, st_c, Bg * groups, T, dtype, "cpu", train=True, groups=groups, drop_mode="mask")
    pg = nets.NetPlan(spec_g, st_g, Bg * groups, T, dtype, "cuda", train=True, groups=groups, drop_mode="mask")
    for i in range(groups):
        pc.set_masks(masks[i], group=i)
        pc.motion_src[i].copy_(srcs[i])
    olb = pc.bufs[pc.out_layer.name]
    olb.dpre[:, :, :1] = (torch.randn(olb.dpre.shape[0], olb.dpre.shape[1], 1, generator=g) * 0.01).to(olb.dpre.dtype)
    sync_inputs(pg.prog, pc.prog)
    rep = replay_pair(pg.prog, pc.prog, ["pack", "fwd", "bwd"])
    _check(f"disc-train Bg={Bg} T={T} groups={groups} dtype={dtype}", rep, TOL[dtype])
