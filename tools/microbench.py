#!/usr/bin/env python
"""True device time of single ops: the op is captured N times in one CUDA graph (no host gaps) and the replay
is timed with CUDA events.  usage: python tools/microbench.py [substring-of-op-tag ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2h_b200  # noqa: F401,E402
from b2h_b200.trainer import GanTrainer  # noqa: E402

pats = sys.argv[1:] or ["stats.conv5", "apply.conv6", "bn_bwd.conv6", "gemm.conv5", "gemm.decoder.9", "wgrad.conv5",
                        "dgrad.conv6", "stats.convs.25", "bn_bwd.convs.25", "apply.convs.29", "gemm.convs.25",
                        "prep.encoder", "l1", "adam", "pack_multi", "out"]
N = 40
tr = GanTrainer("v1", 36, 252, False, 256, 64, precision=os.environ.get("PREC", "bf16"), device="cuda:0")
tr.x.normal_()
tr.y.normal_()
for _ in range(2):
    tr.generator_step()
    tr.discriminator_step()
torch.cuda.synchronize()
progs = [("G_train", tr.G_train.prog), ("D_eval", tr.D_eval.prog), ("g_loss", tr.g_loss_prog), ("G_eval", tr.G_eval.prog),
         ("D_train", tr.D_train.prog), ("d_loss", tr.d_loss_prog)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
from b2h_b200 import _lib as L  # noqa: E402
by_kind = {}
STEP_SEGS = {"G_train": ("pack", "fwd", "bwd"), "D_eval": ("pack", "fwd"), "g_loss": ("loss", "opt"), "G_eval": ("pack", "fwd"),
             "D_train": ("pack", "fwd", "bwd"), "d_loss": ("loss", "opt")}
for pname, prog in progs:
    for i, rec in enumerate(prog.recs):
        tag = f"{pname}.{rec.tag}"
        if not any(p in tag for p in pats):
            continue
        prog.run_range(i, i + 1)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(N):
                prog.run_range(i, i + 1)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        e1.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (5 * N)
        in_step = any(prog.segments[sg][0] <= i < prog.segments[sg][1] for sg in STEP_SEGS[pname] if sg in prog.segments)
        if in_step:
            k = L.OP_STRUCT[rec.kind].__name__
            by_kind.setdefault(k, [0, 0.0])
            by_kind[k][0] += 1
            by_kind[k][1] += us
        if os.environ.get("QUIET") is None:
            print(f"{us:8.2f} us  {tag}")
tot = sum(v[1] for v in by_kind.values())
print(f"== ops of one GAN step: {sum(v[0] for v in by_kind.values())} ops, {tot:.1f} us back-to-back device time")
for k, (n, us) in sorted(by_kind.items(), key=lambda kv: -kv[1][1]):
    print(f"   {k:10s} {n:3d} ops {us:8.1f} us {100 * us / tot:5.1f}%  avg {us / n:6.2f}")
