"""Fused data-parallel optimizer step (b2h_dp_adam: reduce-scatter of the gradients over peer memory + Adam on the
owned slice + all-gather of the parameters, one kernel) — host logic and arithmetic on CPU with both "ranks" inside one
process (PeerBuffers.in_process), against the oracle that averages the gradients of two reference replicas
(SURVEY.md 8e), and against the NCCL-shaped path (all-reduce + Adam) it replaces.  The CUDA kernel itself needs two
GPUs: see test_dp_adam_kernel_two_devices (skipped on a single-GPU box)."""
import ctypes as C

import pytest
import torch

import b2h_b200  # noqa: F401
from b2h_b200 import _lib as L
from b2h_b200.program import _fill_struct
from b2h_b200.trainer import GanTrainer, PeerBuffers
from oracle import ops_emul as E
from oracle import ref_models as R
from tests.test_plan_emulated import randomize_bn
from tests.test_trainer_emulated import check_adam_params, grads_close

WORLD, B, T, CIN, COUT, LR = 2, 8, 16, 36, 252, 1e-3


def test_slices_tile_the_buffer():
    for n in (4, 8, 12, 1000, 2240864, 121684):
        for world in (1, 2, 3, 4, 8, 16):
            edges = [E.dp_slice(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert all(lo % 4 == 0 and hi % 4 == 0 and hi >= lo for lo, hi in edges)


def _trainers(n_buckets):
    torch.manual_seed(0)
    G = R.build_generator("v1", CIN, COUT)
    D = R.build_discriminator(COUT)
    randomize_bn(G, 5)
    randomize_bn(D, 6)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(WORLD * B, CIN, T, generator=g)
    y = torch.randn(WORLD * B, COUT, T, generator=g)
    probe = GanTrainer("v1", CIN, COUT, False, B, T, precision="fp32", device="cpu", lr=LR, drop_mode="mask")
    peers = {"g": PeerBuffers.in_process(probe.g_store.n, n_buckets, ["cpu"] * WORLD),
             "d": PeerBuffers.in_process(probe.d_store.n, n_buckets, ["cpu"] * WORLD)}
    trs = []
    for r in range(WORLD):
        tr = GanTrainer("v1", CIN, COUT, False, B, T, precision="fp32", device="cpu", lr=LR, drop_mode="mask",
                        world_size=WORLD, n_buckets=n_buckets, fused_dp=True, rank=r,
                        peer_buffers={k: v[r] for k, v in peers.items()})
        tr.g_store.load_state_dict(G.state_dict())
        tr.d_store.load_state_dict(D.state_dict())
        tr.x.copy_(x[r * B:(r + 1) * B])
        tr.y.copy_(y[r * B:(r + 1) * B])
        tr.G_train.set_masks(R.make_masks(G, x[r * B:(r + 1) * B], seed=100 + r))
        trs.append(tr)
    return G, D, x, y, trs


@pytest.mark.parametrize("n_buckets", [1, 3])
def test_two_rank_fused_generator_step_equals_averaged_replicas(n_buckets):
    G, D, x, y, trs = _trainers(n_buckets)
    run = lambda prog, seg: E.run_records(prog.recs, *prog.segments[seg])  # noqa: E731
    # the parameters / gradients of a rank ARE its peer buffers
    for r, tr in enumerate(trs):
        assert tr.g_store.flat.data_ptr() == tr._peer["g"].p_ptrs[r] and tr.g_store.grad.data_ptr() == tr._peer["g"].g_ptrs[r]
    for tr in trs:
        for prog, seg in ((tr.G_train.prog, "pack"), (tr.D_train.prog, "pack"), (tr.D_eval.prog, "pack"),
                          (tr.G_train.prog, "fwd"), (tr.g_loss_prog, "loss"), (tr.D_eval.prog, "fwd")):
            run(prog, seg)     # (l1 writes G_train.out; the adversarial VALUE is not checked here)
    # backward bucket by bucket; the optimizer program of a bucket is the fused op (no all-reduce anywhere)
    bp = trs[0]._buckets["g"][0]
    assert len(bp) == n_buckets
    for tr in trs:
        run(tr._buckets["g"][1], "step")
    sums = None
    for i, (s, e, lo, hi, _names) in enumerate(bp):
        for tr in trs:
            E.run_records(tr.G_train.prog.recs, s, e)
        if sums is None:
            sums = torch.zeros_like(trs[0].g_store.grad)
        sums[lo:hi] = sum(tr.g_store.grad[lo:hi] for tr in trs)
        for tr in reversed(trs):       # any order of the ranks is the same collective
            P = tr._buckets["g"][1]
            first, end = P.segments[f"b{i}"]
            assert [rec.kind for rec in P.recs[first:end]] == [L.OP_DP_ADAM]
            run(P, f"b{i}")
            run(tr.G_train.prog, tr._buckets["g"][2][i])
    # every rank holds identical parameters, and each rank's moments are live exactly on the slices it owns
    assert torch.equal(trs[0].g_store.flat, trs[1].g_store.flat)
    for r, tr in enumerate(trs):
        own = torch.zeros(tr.g_store.n, dtype=torch.bool)
        for (_, _, lo, hi, _) in bp:
            a, b = E.dp_slice(hi - lo, WORLD, r)
            own[lo + a:lo + b] = True
        assert float(tr.g_opt.m[~own].abs().max()) == 0.0 and float(tr.g_opt.m[own].abs().max()) > 0.0
    # single-process oracle: two reference replicas on the two shards, gradients averaged, one Adam step
    replicas = []
    for r in range(WORLD):
        Gr = R.build_generator("v1", CIN, COUT)
        Gr.load_state_dict(G.state_dict())
        xs, ys = x[r * B:(r + 1) * B], y[r * B:(r + 1) * B]
        Gr.train()
        Gr.set_masks(R.make_masks(Gr, xs, seed=100 + r))
        torch.nn.functional.l1_loss(Gr(xs), ys).backward()
        replicas.append(Gr)
    opt = torch.optim.Adam(G.parameters(), lr=LR)
    for (k, p), *rs in zip(G.named_parameters(), *[rep.named_parameters() for rep in replicas]):
        p.grad = sum(q.grad for _, q in rs) / WORLD
    opt.step()
    st = trs[0].g_store
    for k, p in G.named_parameters():
        assert grads_close(st._view(sums, k, st.param_shapes, st.offsets) / WORLD, p.grad, 5e-5), k
        check_adam_params(st.p(k), p, LR, k)


def test_fused_dp_is_rejected_where_it_cannot_work():
    with pytest.raises(ValueError):
        GanTrainer("v1", CIN, COUT, False, 4, 16, precision="fp32", device="cpu", world_size=1, fused_dp=True)


def test_descriptor_lowering():
    """The recorded op lowers to b2h_dp_adam_t: peer pointer tables offset to the bucket, per-site signal pads."""
    *_, trs = _trainers(3)
    tr = trs[1]
    bp, P, _ = tr._buckets["g"]
    for i, (_, _, lo, hi, _) in enumerate(bp):
        rec = P.recs[P.segments[f"b{i}"][0]]
        st = _fill_struct(L.DpAdam(), rec.f)
        pb = tr._peer["g"]
        assert st.rank == 1 and st.world == WORLD and st.n == hi - lo
        assert [st.p[q] for q in range(WORLD)] == [a + 4 * lo for a in pb.p_ptrs]
        assert [st.g[q] for q in range(WORLD)] == [a + 4 * lo for a in pb.g_ptrs]
        assert [st.signal[q] for q in range(WORLD)] == [a + 4 * i * PeerBuffers.SITE_WORDS for a in pb.sig_ptrs]
        assert st.p[WORLD] is None and not st.g_mc and not st.p_mc
        assert st.m == tr.g_opt.m[lo:hi].data_ptr() and st.scalars == tr.g_opt.scalars.data_ptr()
    assert C.sizeof(L.DpAdam) == L.load().b2h_desc_size(L.OP_DP_ADAM)


@pytest.mark.gpu
def test_dp_adam_kernel_two_devices():
    """The CUDA kernel with one rank per device inside one process (peer access on plain cudaMalloc memory): both
    launches are asynchronous, the CTAs of the two devices meet in the kernel's barriers.  Needs >= 2 GPUs."""
    import os
    if not os.environ.get("B2H_TEST_MULTI_GPU"):
        pytest.skip("opt-in (B2H_TEST_MULTI_GPU=1): first hardware run of the kernel is the next round's, under a timeout")
    if torch.cuda.device_count() < 2 or not torch.cuda.can_device_access_peer(0, 1):
        pytest.skip("needs two peer-accessible GPUs")
    world, n = 2, 4 * 50021
    devs = [torch.device("cuda", r) for r in range(world)]
    for a in devs:                       # a copy in each direction makes torch enable peer access both ways
        for b in devs:
            if a != b:
                torch.zeros(4, device=a).copy_(torch.ones(4, device=b))
    pbs = PeerBuffers.in_process(n, 1, devs)
    gen = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=gen)
    grads = [torch.randn(n, generator=gen) * 1e-2 for _ in range(world)]
    m0, v0 = torch.randn(n, generator=gen) * 1e-3, torch.rand(n, generator=gen) * 1e-4
    lr, b1, b2, eps, t = 1e-3, 0.9, 0.999, 1e-8, 5
    ms, vs, scal = [], [], []
    for r, pb in enumerate(pbs):
        pb.flat.copy_(p0)
        pb.grad.copy_(grads[r])
        ms.append(m0.to(devs[r]))
        vs.append(v0.to(devs[r]))
        scal.append(torch.tensor([-(lr / (1 - b1 ** t)), (1 - b2 ** t) ** 0.5], dtype=torch.float32, device=devs[r]))
    for d in devs:
        torch.cuda.synchronize(d)
    for rep in range(2):                 # twice: the signal pads reset themselves
        for r, pb in enumerate(pbs):
            with torch.cuda.device(devs[r]):
                desc = _fill_struct(L.DpAdam(), dict(p=pb.p_ptrs, g=pb.g_ptrs, signal=pb.sig_ptrs, m=ms[r], v=vs[r], n=n,
                                                     rank=r, world=world, beta1=b1, beta2=b2, eps=eps, gscale=1.0 / world,
                                                     scalars=scal[r], timeout_ms=5000))
                L.run_oneshot(desc, L.F32, C.c_void_p(torch.cuda.current_stream(devs[r]).cuda_stream))
        for d in devs:
            torch.cuda.synchronize(d)
    # reference: two plain Adam steps (same bias-correction scalars) on the averaged gradient
    g = sum(grads) / world
    p, m, v = p0.clone(), m0.clone(), v0.clone()
    for rep in range(2):
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        p.addcdiv_(m, (v.sqrt() / (1 - b2 ** t) ** 0.5).add_(eps), value=-(lr / (1 - b1 ** t)))
    for r, pb in enumerate(pbs):
        assert float((pb.flat.cpu() - p).abs().max()) <= 1e-6 * float(p.abs().max())
        assert int(pb.signals.abs().max()) == 0
    assert torch.equal(pbs[0].flat.cpu(), pbs[1].flat.cpu())


def test_signal_protocol_model():
    """Randomised interleaving model of the kernel's barrier protocol (k_dp.cu: thread q of a CTA puts = CAS 0->1 on peer
    q's slot [rank], then waits = CAS 1->0 on the own slot [q]; bar.sync around; the same slots serve the entry barrier,
    the exit barrier and the next call): no deadlock, the pads end reset, no rank passes the entry barrier of call k
    before every rank has finished call k-1's data phase and entered call k, nobody passes the exit barrier of call k
    before every rank has finished its data phase — over several back-to-back calls (pads never re-initialised)."""
    import random

    def signal_thread(r, q, pads):
        while pads[q][r] != 0:           # put: CAS(peer q's slot for me, 0 -> 1), spin while the last one is unconsumed
            yield
        pads[q][r] = 1
        yield
        while pads[r][q] != 1:           # wait: CAS(my slot for peer q, 1 -> 0)
            yield
        pads[r][q] = 0

    for world in (2, 3, 8):
        for seed in range(25):
            rnd = random.Random(seed * 100 + world)
            pads = [[0] * world for _ in range(world)]
            calls, log = 4, []
            phase = [0] * world          # per rank: 2 * call + (0 entry barrier | 1 exit barrier)
            live = {r: [signal_thread(r, q, pads) for q in range(world)] for r in range(world)}
            log += [("entered", r, 0) for r in range(world)]
            steps = 0
            while live:
                r = rnd.choice(list(live))
                for _ in range(rnd.choice((1, 1, 1, 7, 60))):      # bursts let one rank run far ahead
                    if not live[r]:
                        break
                    t = rnd.choice(live[r])
                    try:
                        next(t)
                    except StopIteration:
                        live[r].remove(t)
                if not live[r]:                                     # bar.sync: every signalling thread is through
                    k, which = divmod(phase[r], 2)
                    log.append(("passed_exit" if which else "passed_entry", r, k))
                    phase[r] += 1
                    if which == 0:
                        log.append(("data_done", r, k))             # (the data phase itself has no synchronisation)
                    elif k + 1 < calls:
                        log.append(("entered", r, k + 1))
                    if phase[r] < 2 * calls:
                        live[r] = [signal_thread(r, q, pads) for q in range(world)]
                    else:
                        del live[r]
                steps += 1
                assert steps < 3_000_000, "deadlock / livelock in the signal protocol"
            assert all(v == 0 for row in pads for v in row)
            pos = {e: i for i, e in enumerate(log)}
            for k in range(calls):
                first_entry = min(pos[("passed_entry", r, k)] for r in range(world))
                assert all(pos[("entered", r, k)] < first_entry for r in range(world))
                first_exit = min(pos[("passed_exit", r, k)] for r in range(world))
                assert all(pos[("data_done", r, k)] < first_exit for r in range(world))


def test_two_rank_fused_discriminator_step_equals_averaged_replicas():
    """The discriminator's optimizer programs under fused_dp (grouped fake / real batch, one bucket): both ranks'
    emulated steps against two reference replicas with averaged gradients + one torch.optim.Adam step."""
    import copy
    G, D, x, y, trs = _trainers(1)
    run = lambda prog, seg: E.run_records(prog.recs, *prog.segments[seg])  # noqa: E731
    G.eval()
    masks = []
    for r, tr in enumerate(trs):
        xs, ys = x[r * B:(r + 1) * B], y[r * B:(r + 1) * B]
        with torch.no_grad():
            fake = G(xs)
        mf = R.make_masks(D, R.calc_motion(fake), seed=200 + r)
        mr = R.make_masks(D, R.calc_motion(ys), seed=300 + r)
        masks.append((mf, mr))
        tr.D_train.set_masks(mf, group=0)
        tr.D_train.set_masks(mr, group=1)
        tr._sync_d_batch()
        for prog, seg in ((tr.G_train.prog, "pack"), (tr.G_eval.prog, "pack"), (tr.D_train.prog, "pack"),
                          (tr.G_eval.prog, "fwd"), (tr.D_train.prog, "fwd"), (tr.d_loss_prog, "loss")):
            run(prog, seg)
    (s, e, lo, hi, _), = trs[0]._buckets["d"][0]
    assert (lo, hi) == (0, trs[0].d_store.n)
    for tr in trs:
        run(tr._buckets["d"][1], "step")
        E.run_records(tr.D_train.prog.recs, s, e)
    sums = sum(tr.d_store.grad.clone() for tr in trs)
    for tr in trs:
        run(tr._buckets["d"][1], "b0")
        run(tr.D_train.prog, tr._buckets["d"][2][0])
    assert torch.equal(trs[0].d_store.flat, trs[1].d_store.flat)
    # oracle
    replicas = []
    for r in range(WORLD):
        Dr = copy.deepcopy(D)
        xs, ys = x[r * B:(r + 1) * B], y[r * B:(r + 1) * B]
        R.discriminator_step(G, Dr, torch.optim.SGD(Dr.parameters(), lr=0.0), xs, ys, None, masks[r][0], masks[r][1], False)
        replicas.append(Dr)
    opt = torch.optim.Adam(D.parameters(), lr=LR)
    for (k, p), *rs in zip(D.named_parameters(), *[rep.named_parameters() for rep in replicas]):
        p.grad = sum(q.grad for _, q in rs) / WORLD
    opt.step()
    st = trs[0].d_store
    for k, p in D.named_parameters():
        assert grads_close(st._view(sums, k, st.param_shapes, st.offsets) / WORLD, p.grad, 5e-4), k
        check_adam_params(st.p(k), p, LR, k)
