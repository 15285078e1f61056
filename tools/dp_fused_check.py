#!/usr/bin/env python
"""Multi-GPU check of the fused data-parallel optimizer step (b2h_dp_adam) against the ncclAllReduce + b2h_adam path.

    gpurun --gpus 2 --timeout 300 -- 'timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29511 tools/dp_fused_check.py > gpurun_out/dp_fused.log 2>&1'

Per rank: two trainers with identical weights and batches, one per exchange; K pipelined GAN steps each; the
parameters must agree to fp32 summation-order noise, every rank must hold the same parameters, and the device time
per step of both is printed (CUDA events, max over ranks).  B2H_DP_NO_MULTICAST=1 forces the peer load / store
path instead of multimem (NVLS); B2H_DP_PEER=ipc maps the peers through CUDA IPC handles instead of torch's symmetric
memory.  The kernel traps after B2H_DP_TIMEOUT_MS (default 10 s) if a peer never arrives.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2h_b200  # noqa: E402,F401
from b2h_b200.trainer import GanTrainer  # noqa: E402


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, T, steps = int(os.environ.get("B", 256)), int(os.environ.get("T", 64)), int(os.environ.get("STEPS", 20))
    precision = os.environ.get("PRECISION", "bf16")
    trs = {}
    for name, fused in (("nccl", False), ("fused", True)):
        tr = GanTrainer("v1", 36, 252, False, B, T, precision=precision, device=dev, lr=1e-4, seed=23456,
                        drop_mode="none", world_size=world, process_group=dist.group.WORLD, fused_dp=fused)
        for st in (tr.g_store, tr.d_store):
            dist.broadcast(st.flat, 0)
            dist.broadcast(st.bufs, 0)
        g = torch.Generator().manual_seed(100 + rank)
        tr.load_batch(torch.randn(B, 36, T, generator=g).to(dev), torch.randn(B, 252, T, generator=g).to(dev))
        trs[name] = tr
    if rank == 0:
        pb = trs["fused"]._peer["g"]
        print(f"world {world}: multicast {'on' if pb.g_mc else 'off'}, generator {trs['fused'].g_store.n} params, "
              f"{trs['fused'].n_buckets} bucket(s)", flush=True)
    res = {}
    for name, tr in trs.items():
        tr.generator_step(graph=True)      # pipeline prologue, as bench.py: G0, then every step is [D_k || G_k+1]
        tr._sync_d_batch()
        for _ in range(3):
            tr.gan_step(graph=True)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            tr.gan_step(graph=True)
        e1.record()
        tr.flush_adv()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[name] = float(ms)
    ok = True
    for key in ("g_store", "d_store"):
        a, b = getattr(trs["nccl"], key).flat, getattr(trs["fused"], key).flat
        # (Adam divides by sqrt(v) + eps: where a gradient is ~0 the two summation orders may step in opposite
        # directions, so the bound is the step size times the number of steps; typical elements agree to ~1e-6)
        err = float((a - b).abs().max() / a.abs().max())
        bound = 2.05 * 1e-4 * (steps + 4) / float(a.abs().max())
        lo, hi = b.double().sum().reshape(1).clone(), b.double().sum().reshape(1).clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = float(hi - lo) == 0.0
        ok = ok and err <= bound and same and bool(torch.isfinite(b).all())
        if rank == 0:
            print(f"{key}: fused vs nccl max rel diff {err:.3e} after {steps + 4} steps; identical on all ranks: {same}",
                  flush=True)
    if rank == 0:
        print(f"ms per gan_step: nccl {res['nccl']:.4f}  fused {res['fused']:.4f}", flush=True)
        print("dp_fused_check", "OK" if ok else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
