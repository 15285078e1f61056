#!/usr/bin/env python
"""Steady-state device time of every op of the GAN step: each op is launched `reps` times back to back
(warm caches, no host gaps) between two CUDA events.  Prints a table and the sum next to the time of the
CUDA-graph replay of the whole step."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2h_b200  # noqa: F401,E402
from b2h_b200 import _lib as L  # noqa: E402
from b2h_b200.trainer import GanTrainer  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
reps = 20
tr = GanTrainer("v1", 36, 252, False, 256, 64, precision=precision, device="cuda:0")
tr.x.normal_()
tr.y.normal_()
for _ in range(3):
    tr.generator_step(graph=True)
    tr.discriminator_step(graph=True)
torch.cuda.synchronize()
# graph replay time of the whole step (warm)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    tr.generator_step(graph=True)
    tr.discriminator_step(graph=True)
e1.record()
e1.synchronize()
step_ms = e0.elapsed_time(e1) / 20
e0.record()
for _ in range(20):
    tr.generator_step(graph=True)
e1.record()
e1.synchronize()
g_ms = e0.elapsed_time(e1) / 20
rows = []
seq = [("G_train", tr.G_train.prog, ("pack", "fwd", "bwd")), ("D_eval", tr.D_eval.prog, ("pack", "fwd")),
       ("g_loss", tr.g_loss_prog, ("loss", "opt")), ("G_eval", tr.G_eval.prog, ("pack", "fwd")),
       ("D_train", tr.D_train.prog, ("pack", "fwd", "bwd")), ("d_loss", tr.d_loss_prog, ("loss", "opt"))]
for pname, prog, segs in seq:
    for seg in segs:
        s, e = prog.segments[seg]
        for i in range(s, e):
            rec = prog.recs[i]
            prog.run_range(i, i + 1)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                prog.run_range(i, i + 1)
            e1.record()
            e1.synchronize()
            rows.append((e0.elapsed_time(e1) / reps * 1e3, f"{pname}.{seg}.{rec.tag}", L.OP_STRUCT[rec.kind].__name__))
tot = sum(r[0] for r in rows)
print(f"precision={precision}  graph step (G+D) = {step_ms*1e3:.1f} us, G step alone = {g_ms*1e3:.1f} us, "
      f"sum of per-op steady-state times = {tot:.1f} us over {len(rows)} ops")
by_kind = {}
for us, tag, kind in rows:
    by_kind.setdefault(kind, [0, 0.0])
    by_kind[kind][0] += 1
    by_kind[kind][1] += us
for k, (n, us) in sorted(by_kind.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:10s} {n:4d} ops {us:9.1f} us  {100*us/tot:5.1f}%  avg {us/n:7.2f}")
print("top ops:")
for us, tag, kind in sorted(rows, reverse=True)[:40]:
    print(f"  {us:8.2f} us  {tag}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"step_us": step_ms * 1e3, "g_step_us": g_ms * 1e3, "ops": rows}, open(f"gpurun_out/op_times_{precision}.json", "w"))
