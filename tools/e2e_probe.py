#!/usr/bin/env python
"""Which part of bench.py's end-to-end loop costs what: the pipelined gan_step alone, + the batch hand-over copies
(swap_batch from device staging), + the pinned-host -> device prefetch, + the lagged loss read.  Host wall clock per
step, median of 3 runs of 50 steps each."""
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2h_b200  # noqa: F401,E402
from b2h_b200.trainer import GanTrainer  # noqa: E402
from bench import synth_batch  # noqa: E402

dev = torch.device("cuda", 0)
tr = GanTrainer("v1", 36, 252, False, 256, 64, precision="bf16", device=dev, lr=1e-4, seed=1, drop_mode="philox")
x, y, _ = synth_batch(256, 64, 36, 252, None, seed=0)
hx, hy = x.pin_memory(), y.pin_memory()
tr.load_batch(hx, hy)
tr.generator_step(graph=True)
for _ in range(5):
    tr.prefetch_batch(hx, hy)
    tr.swap_batch(pipelined=True)
    tr.gan_step(graph=True)
torch.cuda.synchronize()
h_loss = [torch.empty(8).pin_memory() for _ in range(2)]
N = 50


def run(mode):
    ev = [None, None]
    torch.cuda.synchronize()
    if "h2d" in mode:
        tr.prefetch_batch(hx, hy)
    t0 = time.perf_counter()
    for k in range(N):
        if "swap" in mode:
            tr.swap_batch(pipelined=True)
        if "h2d" in mode:
            tr.prefetch_batch(hx, hy)
        tr.gan_step(graph=True)
        if "loss" in mode:
            h_loss[k & 1].copy_(tr.losses, non_blocking=True)
            ev[k & 1] = torch.cuda.Event()
            ev[k & 1].record()
            if k > 0:
                ev[(k - 1) & 1].synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / N * 1e3


# a device-only H2D timing: how long does the 18.9 MB copy take on this box?
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.empty_like(tr.x), torch.empty_like(tr.y)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    st[0].copy_(hx, non_blocking=True)
    st[1].copy_(hy, non_blocking=True)
e1.record()
e1.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"H2D of one batch (18.9 MB, pinned): {ms:.3f} ms = {(hx.numel() + hy.numel()) * 4 / ms / 1e6:.1f} GB/s", flush=True)
for mode in ("step", "step+swap", "step+swap+h2d", "step+swap+h2d+loss", "step+loss"):
    if "h2d" not in mode and "swap" in mode:
        tr.prefetch_batch(hx, hy)     # a valid prefetch event for swap_batch to wait on
        torch.cuda.synchronize()
    run(mode)
    vals = [run(mode) for _ in range(3)]
    print(f"{mode:24s} {statistics.median(vals):7.3f} ms/step   runs {[round(v, 3) for v in vals]}", flush=True)
