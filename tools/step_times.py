#!/usr/bin/env python
"""Device time of the generator step, the discriminator step and the pipelined gan_step, each replayed alone."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2h_b200  # noqa: F401,E402
from b2h_b200.trainer import GanTrainer  # noqa: E402

tr = GanTrainer("v1", 36, 252, False, 256, 64, precision=os.environ.get("PREC", "bf16"), device="cuda:0")
tr.x.normal_()
tr.y.normal_()
tr._sync_d_batch()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(name, fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) * 1e3 / n:8.1f} us")


timeit("generator_step", lambda: tr.generator_step(graph=True))
timeit("discriminator_step", lambda: tr.discriminator_step(graph=True))
timeit("gan_step(lag_adv=False)", lambda: tr.gan_step(graph=True, lag_adv=False))
timeit("gan_step(lag_adv=True)", lambda: tr.gan_step(graph=True, lag_adv=True))

# experiment: the two sequential-step graphs replayed concurrently on two streams (ignores the cross-step
# dependencies; timing only) — an upper bound of what branch overlap can give without intra-graph events
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gg, gd = tr._graphs["g"], tr._graphs["d"]


def both():
    ev = torch.cuda.Event()
    ev.record()
    s1.wait_event(ev)
    s2.wait_event(ev)
    with torch.cuda.stream(s1):
        gg.replay()
        e1_ = torch.cuda.Event()
        e1_.record()
    with torch.cuda.stream(s2):
        gd.replay()
        e2_ = torch.cuda.Event()
        e2_.record()
    torch.cuda.current_stream().wait_event(e1_)
    torch.cuda.current_stream().wait_event(e2_)


timeit("two graphs on two streams", both)
