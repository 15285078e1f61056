"""Data-parallel host logic on CPU: two gloo ranks interpret their shard of the step with the op
restatements, all-reduce the flat gradient bucket by bucket exactly as GanTrainer does, and must land on
the parameters of a single-process oracle that averages the gradients of two reference replicas
(per-process BatchNorm statistics, summed gradients scaled by 1/world inside Adam)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import b2h_b200  # noqa: F401
from b2h_b200.trainer import GanTrainer
from oracle import ops_emul as E
from oracle import ref_models as R
from tests.test_plan_emulated import randomize_bn, rel_err
from tests.test_trainer_emulated import check_adam_params, grads_close

WORLD, B, T, CIN, COUT, LR = 2, 8, 16, 36, 252, 1e-3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(seed=0):
    torch.manual_seed(seed)
    G = R.build_generator("v1", CIN, COUT)
    D = R.build_discriminator(COUT)
    randomize_bn(G, 5)
    randomize_bn(D, 6)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(WORLD * B, CIN, T, generator=g)
    y = torch.randn(WORLD * B, COUT, T, generator=g)
    return G, D, x, y


def _worker(rank, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    torch.set_num_threads(2)
    G, D, x, y = _setup()
    xs, ys = x[rank * B:(rank + 1) * B], y[rank * B:(rank + 1) * B]
    tr = GanTrainer("v1", CIN, COUT, False, B, T, precision="fp32", device="cpu", lr=LR, drop_mode="mask",
                    world_size=WORLD, process_group=dist.group.WORLD, n_buckets=3)
    tr.g_store.load_state_dict(G.state_dict())
    tr.d_store.load_state_dict(D.state_dict())
    tr.x.copy_(xs)
    tr.y.copy_(ys)
    masks = R.make_masks(G, xs, seed=100 + rank)
    tr.G_train.set_masks(masks)
    run = lambda prog, seg: E.run_records(prog.recs, *prog.segments[seg])  # noqa: E731
    for prog, seg in ((tr.G_train.prog, "pack"), (tr.D_train.prog, "pack"), (tr.D_eval.prog, "pack"),
                      (tr.G_train.prog, "fwd"),
                      (tr.g_loss_prog, "loss"),       # (l1 writes G_train.out; the adversarial VALUE is not checked here)
                      (tr.D_eval.prog, "fwd")):
        run(prog, seg)
    buckets = tr.bucket_plan(tr.G_train)
    # the buckets tile the flat gradient exactly once, from the end of the buffer towards its start
    assert buckets[0][3] == tr.g_store.n and buckets[-1][2] == 0
    assert all(a[2] == b[3] for a, b in zip(buckets, buckets[1:]))
    first, end = tr.G_train.prog.segments["bwd"]
    assert buckets[0][0] == first and buckets[-1][1] == end and all(a[1] == b[0] for a, b in zip(buckets, buckets[1:]))
    for (s, e, lo, hi, _names) in buckets:
        E.run_records(tr.G_train.prog.recs, s, e)
        if hi > lo:
            dist.all_reduce(tr.g_store.grad[lo:hi], op=dist.ReduceOp.SUM)
    run(tr.g_loss_prog, "opt")      # Adam with gscale = 1 / world
    if rank == 0:
        ret["flat"] = tr.g_store.flat.clone()
        ret["grad"] = tr.g_store.grad.clone()
        ret["offsets"] = dict(tr.g_store.offsets)
        ret["shapes"] = dict(tr.g_store.param_shapes)
    # every rank must hold identical parameters after the step
    other = tr.g_store.flat.clone()
    dist.broadcast(other, 0)
    assert torch.equal(other, tr.g_store.flat)
    dist.destroy_process_group()


def test_two_rank_generator_step_equals_averaged_replicas():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(port, ret), nprocs=WORLD, join=True)
    # single-process oracle: two reference replicas on the two shards, gradients averaged, one Adam step
    G, D, x, y = _setup()
    replicas = []
    for r in range(WORLD):
        Gr = R.build_generator("v1", CIN, COUT)
        Gr.load_state_dict(G.state_dict())
        xs, ys = x[r * B:(r + 1) * B], y[r * B:(r + 1) * B]
        masks = R.make_masks(Gr, xs, seed=100 + r)
        Gr.train()
        Gr.set_masks(masks)
        out = Gr(xs)
        torch.nn.functional.l1_loss(out, ys).backward()
        replicas.append(Gr)
    opt = torch.optim.Adam(G.parameters(), lr=LR)
    for (k, p), *rs in zip(G.named_parameters(), *[rep.named_parameters() for rep in replicas]):
        p.grad = sum(q.grad for _, q in rs) / WORLD
    opt.step()
    flat, grad, offsets, shapes = ret["flat"], ret["grad"], ret["offsets"], ret["shapes"]
    for k, p in G.named_parameters():
        n = p.numel()
        ours_g = grad[offsets[k]:offsets[k] + n].view(shapes[k]) / WORLD
        assert grads_close(ours_g, p.grad, 5e-5), k
        ours_p = flat[offsets[k]:offsets[k] + n].view(shapes[k])
        check_adam_params(ours_p, p, LR, k)
