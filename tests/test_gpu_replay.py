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

# fp32 mode: the north_star bar (1e-5) per op on identical inputs -- only the accumulation order differs from the CPU
# restatement (measured worst case over all ops of all programs in round 1: 3e-6); bf16 mode: both sides round stored
# activations to bf16, accumulation-order differences flip single bf16 ulps (2^-8)
TOL = {L.F32: 1e-5, L.BF16: 1.2e-2}


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


def tile_classes(prog):
    """The kernel template instances a program launches, as a set of labels (b2h_program_op_plan)."""
    out = set()
    for r in prog.tile_report():
        if not r["tensor_core"]:
            continue
        if r["kind"] == L.OP_GEMM:
            out.add(f"gemm:BN{r['tile_n']}")
            out.add(f"gemm:BN{r['tile_n']}:" + ("merged" if r["merged"] else "unmerged"))
            if r["fuse_stats"]:
                out.add(f"gemm:BN{r['tile_n']}:stats")
            if r["fuse_bwd"]:
                out.add(f"gemm:BN{r['tile_n']}:bwdsum")
        else:
            out.add(f"wgrad:WN{r['tile_n']}")
            out.add(f"wgrad:WN{r['tile_n']}:" + ("direct" if r["splits"] == 1 else "splitk"))
    return out


# BASELINE config 2 / 3 / 4 at the benchmarked per-GPU batch (256 clips x 64 frames)
GEN_BENCH = [("v1", False, 36, 252, 256, 64), ("v1", True, 36, 252, 256, 64), ("b2h", True, 36, 252, 256, 64)]
# the template instances of the benchmarked bf16 step (profiles/launches_r01_final.md): every one must be replayed
BENCH_CLASSES = {"gemm:BN256", "gemm:BN128", "gemm:BN64", "gemm:BN256:stats", "gemm:BN256:bwdsum", "gemm:BN128:stats",
                 "gemm:BN128:bwdsum", "gemm:BN256:merged", "wgrad:WN256", "wgrad:WN128", "wgrad:WN256:direct",
                 "wgrad:WN128:direct"}


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
@pytest.mark.parametrize("variant,rf,cin,cout,B,T", GEN + GEN_BENCH)
def test_generator_train_replay(variant, rf, cin, cout, B, T, dtype):
    gen_train_replay(variant, rf, cin, cout, B, T, dtype)


def gen_train_replay(variant, rf, cin, cout, B, T, dtype):
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
    name = f"gen-train {variant} feats={rf} {cin}->{cout} B={B} T={T} dtype={dtype}"
    _check(name, rep, TOL[dtype])
    if dtype == L.BF16:
        classes = tile_classes(pg.prog)
        _log(name + " tile instances", "  " + " ".join(sorted(classes)))
        if (variant, rf, B, T) == ("v1", False, 256, 64):
            # the wide-tile / fused-epilogue instances the benchmark runs are the ones replayed here
            missing = BENCH_CLASSES - classes - {"wgrad:WN128", "wgrad:WN128:direct", "gemm:BN64"}
            assert not missing, f"benched generator plan no longer launches {sorted(missing)}: {sorted(classes)}"
        return classes
    return set()


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
@pytest.mark.parametrize("variant,rf,cin,cout,B,T", GEN[:4])
def test_generator_eval_replay(variant, rf, cin, cout, B, T, dtype):
    gen_eval_replay(variant, rf, cin, cout, B, T, dtype)


def gen_eval_replay(variant, rf, cin, cout, B, T, dtype):
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
    return tile_classes(pg.prog) if dtype == L.BF16 else set()


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
@pytest.mark.parametrize("variant,rf,cin,cout,B,T", [GEN_BENCH[0], ("v1", False, 36, 252, 64, 1024)])
def test_generator_eval_replay_benched_shapes(variant, rf, cin, cout, B, T, dtype):
    """Eval forward at the benchmarked training batch (the discriminator step's generator pass) and at the long end
    of the inference sweep (BASELINE config 5: T = 1024)."""
    gen_eval_replay(variant, rf, cin, cout, B, T, dtype)


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
@pytest.mark.parametrize("B,T", [(256, 64), (8, 192), (5, 21)])
def test_discriminator_eval_replay(B, T, dtype):
    """The scoring pass of the generator step (eval-mode BatchNorm folded into the GEMM epilogues)."""
    disc_eval_replay(B, T, dtype)


def disc_eval_replay(B, T, dtype):
    torch.manual_seed(0)
    D = R.build_discriminator(252)
    randomize_bn(D)
    g = torch.Generator().manual_seed(1)
    src = torch.randn(B, 252, T, generator=g)
    spec_c, spec_g = nets.discriminator_spec(252), nets.discriminator_spec(252)
    st_c = nets.ParamStore(spec_c, "cpu", seed=0)
    st_g = nets.ParamStore(spec_g, "cuda", seed=0)
    st_c.load_state_dict(D.state_dict())
    pc = nets.NetPlan(spec_c, st_c, B, T, dtype, "cpu", train=False)
    pg = nets.NetPlan(spec_g, st_g, B, T, dtype, "cuda", train=False)
    pc.motion_src[0].copy_(src)
    sync_inputs(pg.prog, pc.prog)
    rep = replay_pair(pg.prog, pc.prog, ["pack", "fwd"])
    _check(f"disc-eval B={B} T={T} dtype={dtype}", rep, TOL[dtype])
    return tile_classes(pg.prog) if dtype == L.BF16 else set()


@pytest.mark.parametrize("dtype", [L.F32, L.BF16])
@pytest.mark.parametrize("Bg,T,groups", [(24, 64, 2), (8, 192, 2), (16, 21, 1), (256, 64, 2)])
def test_discriminator_train_replay(Bg, T, groups, dtype):
    disc_train_replay(Bg, T, groups, dtype)


def disc_train_replay(Bg, T, groups, dtype):
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
    pc = nets.NetPlan(spec_c, st_c, Bg * groups, T, dtype, "cpu", train=True, groups=groups, drop_mode="mask",
                      wgrad_direct=True)
    pg = nets.NetPlan(spec_g, st_g, Bg * groups, T, dtype, "cuda", train=True, groups=groups, drop_mode="mask",
                      wgrad_direct=True)   # as GanTrainer builds it
    for i in range(groups):
        pc.set_masks(masks[i], group=i)
        pc.motion_src[i].copy_(srcs[i])
    olb = pc.bufs[pc.out_layer.name]
    olb.dpre[:, :, :1] = (torch.randn(olb.dpre.shape[0], olb.dpre.shape[1], 1, generator=g) * 0.01).to(olb.dpre.dtype)
    sync_inputs(pg.prog, pc.prog)
    rep = replay_pair(pg.prog, pc.prog, ["pack", "fwd", "bwd"])
    name = f"disc-train Bg={Bg} T={T} groups={groups} dtype={dtype}"
    _check(name, rep, TOL[dtype])
    if dtype == L.BF16:
        classes = tile_classes(pg.prog)
        _log(name + " tile instances", "  " + " ".join(sorted(classes)))
        return classes
    return set()


def test_benched_step_tile_instances_are_all_replayed():
    """Union of the kernel template instances of the four plans of the benchmarked GAN step (v1 body, 256 x 64, bf16:
    G train, G eval, grouped D train, D eval) == what the replay cases at that shape cover; and it contains every
    wide-tile / split-K / fused-epilogue instance (those were never compared with the oracle)."""
    from b2h_b200.trainer import GanTrainer
    tr = GanTrainer("v1", 36, 252, False, 256, 64, precision="bf16", device="cuda", drop_mode="mask")
    step = set()
    for plan in (tr.G_train, tr.G_eval, tr.D_train, tr.D_eval):
        step |= tile_classes(plan.prog)
    _log("benched GAN step tile instances", "  " + " ".join(sorted(step)))
    assert BENCH_CLASSES <= step, sorted(BENCH_CLASSES - step)
    assert any(c.endswith(":splitk") for c in step)
    replayed = gen_train_replay("v1", False, 36, 252, 256, 64, L.BF16)
    replayed |= disc_train_replay(256, 64, 2, L.BF16)
    replayed |= gen_eval_replay("v1", False, 36, 252, 256, 64, L.BF16)
    replayed |= disc_eval_replay(256, 64, L.BF16)
    assert step <= replayed, sorted(step - replayed)


@pytest.mark.parametrize("bn,wn", [(256, 256), (128, 128), (64, 64)])
def test_forced_tile_widths_replay(monkeypatch, bn, wn):
    """Every tile width of the tensor-core kernels at a SMALL problem (B2H_FORCE_BN / B2H_FORCE_WN override the
    choosers), kernel by kernel against the restatements: ragged / partially filled wide tiles, fused statistics and
    backward sums at every width, split-K and split-free weight gradients at every width."""
    monkeypatch.setenv("B2H_FORCE_BN", str(bn))
    monkeypatch.setenv("B2H_FORCE_WN", str(wn))
    g = gen_train_replay("v1", False, 36, 252, 8, 64, L.BF16)
    d = disc_train_replay(16, 64, 2, L.BF16)
    t = gen_train_replay("v1", True, 36, 252, 8, 32, L.BF16)
    assert {f"gemm:BN{bn}", f"gemm:BN{bn}:stats", f"gemm:BN{bn}:bwdsum", f"wgrad:WN{wn}"} <= g | t, sorted(g | t)
    assert f"wgrad:WN{wn}:direct" in g and f"wgrad:WN{wn}:splitk" in g, sorted(g)
    assert d


def test_fp32_mode_contractions_run_on_the_tensor_cores():
    """fp32 mode = 3xTF32 on tcgen05 (k_gemm_tf32.cu): every GEMM / weight-gradient op of the benchmarked step carries a
    tensor-core plan (B2H_FP32_SIMT=1 selects the FFMA kernels instead), weight gradients are split-K slices of at most
    24 k-blocks (the bound that keeps the accumulator's truncation below the 1e-5 bar)."""
    from b2h_b200.trainer import GanTrainer
    tr = GanTrainer("v1", 36, 252, False, 256, 64, precision="fp32", device="cuda", drop_mode="mask")
    n = fused_stats = fused_bwd = 0
    for plan in (tr.G_train, tr.G_eval, tr.D_train, tr.D_eval):
        for r in plan.prog.tile_report():
            if r["kind"] in (L.OP_GEMM, L.OP_WGRAD):
                assert r["tensor_core"] and r["tile_n"] in (64, 128), r
                n += 1
                fused_stats += int(r["fuse_stats"])
                fused_bwd += int(r["fuse_bwd"])
                if r["kind"] == L.OP_WGRAD:
                    assert r["splits"] >= 1
    assert n > 60
    # BatchNorm statistics / first backward pass ride in the 3xTF32 epilogues as they do in the bf16 ones
    assert fused_stats >= 14 and fused_bwd >= 10, (fused_stats, fused_bwd)
