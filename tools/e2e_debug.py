#!/usr/bin/env python
"""Where does the end-to-end step time go at N ranks?  (host-side timing of the phases of bench.py's e2e loop)"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2h_b200  # noqa: F401,E402
from b2h_b200.trainer import GanTrainer  # noqa: E402
from bench import synth_batch  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
pg = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
tr = GanTrainer("v1", 36, 252, False, 256, 64, precision="bf16", device=dev, lr=1e-4, seed=1 + rank, drop_mode="philox",
                world_size=world, process_group=pg)
x, y, _ = synth_batch(256, 64, 36, 252, None, seed=rank)
hx, hy = x.pin_memory(), y.pin_memory()
tr.x.copy_(hx)
tr.y.copy_(hy)
for _ in range(5):
    tr.generator_step(graph=True)
    tr.discriminator_step(graph=True)
torch.cuda.synchronize()
h_loss = torch.empty(8).pin_memory()
for mode in ("step+sync", "copy+step+sync", "prefetch+step+sync", "prefetch+step, sync every 10"):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if mode.startswith("prefetch"):
        tr.prefetch_batch(hx, hy)
    t0 = time.perf_counter()
    n = 40
    for i in range(n):
        if mode.startswith("copy"):
            tr.load_batch(hx, hy)
        elif mode.startswith("prefetch"):
            tr.swap_batch()
            tr.prefetch_batch(hx, hy)
        tr.generator_step(graph=True)
        tr.discriminator_step(graph=True)
        h_loss.copy_(tr.losses, non_blocking=True)
        if not mode.endswith("10") or i % 10 == 9:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    if rank == 0:
        print(f"{mode:32s} {dt * 1e3:7.3f} ms/step", flush=True)
if world > 1:
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)
