#!/usr/bin/env python
"""Per-kernel device time of the data-parallel GAN step on rank 0 (CUPTI through torch.profiler; nsys is not in the
image and ncu must not wrap a multi-rank command): how long the exchange kernels (ncclDevKernel_AllReduce* or
dp_adam_kernel) run per step next to the compute kernels.

    gpurun --gpus 8 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        tools/dp_profile.py > gpurun_out/dp_profile_n8.txt'

B2H_FUSED_DP=1 profiles the fused exchange.  Numbers under the profiler are for the SHARE of the step only.
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
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    B, T, steps = int(os.environ.get("B", 256)), int(os.environ.get("T", 64)), int(os.environ.get("STEPS", 20))
    variant, feats = os.environ.get("VARIANT", "v1"), os.environ.get("FEATS", "0") == "1"
    kw = {"n_buckets": int(os.environ["B2H_BUCKETS"])} if os.environ.get("B2H_BUCKETS") else {}
    tr = GanTrainer(variant, 36, 252, feats, B, T, precision=os.environ.get("PRECISION", "bf16"), device=dev, lr=1e-4,
                    seed=23456 + rank, drop_mode="philox", world_size=world, process_group=pg, **kw)
    if world > 1:
        for st in (tr.g_store, tr.d_store):
            dist.broadcast(st.flat, 0)
            dist.broadcast(st.bufs, 0)
    g = torch.Generator().manual_seed(100 + rank)
    f = None
    if feats:
        f = (torch.randn(B, T, 2000, generator=g) if variant == "b2h" else torch.randn(B, 512, generator=g)).to(dev)
    tr.load_batch(torch.randn(B, 36, T, generator=g).to(dev), torch.randn(B, 252, T, generator=g).to(dev), f)
    tr.generator_step(graph=True)
    tr._sync_d_batch()
    for _ in range(5):
        tr.gan_step(graph=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            tr.gan_step(graph=True)
        torch.cuda.synchronize()
    tr.flush_adv()
    if rank == 0:
        rows = {}
        t0, t1 = None, None
        for ev in prof.events():
            if ev.device_type != torch.autograd.DeviceType.CUDA:
                continue
            name = ev.name.split("<")[0].split("(")[0]
            us = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            r = rows.setdefault(name, [0, 0.0])
            r[0] += 1
            r[1] += us
            s, e = ev.time_range.start, ev.time_range.end
            t0 = s if t0 is None else min(t0, s)
            t1 = e if t1 is None else max(t1, e)
        span = (t1 - t0) / steps if t0 is not None else float("nan")
        print(f"world {world}, {variant} feats={feats} {B}x{T}, {steps} gan_steps under torch.profiler (CUPTI): "
              f"{span:.1f} us per step wall span on rank 0; exchange = {'fused dp_adam' if tr.fused_dp else 'NCCL'}, {tr.n_buckets} bucket(s) per network")
        print(f"{'kernel':60s} {'launches/step':>14s} {'us/step':>10s} {'us/launch':>10s}")
        for name, (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:25]:
            print(f"{name[:60]:60s} {n / steps:14.1f} {us / steps:10.1f} {us / n:10.2f}")
    if world > 1:
        torch.cuda.synchronize()
        tr.release_graphs()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
