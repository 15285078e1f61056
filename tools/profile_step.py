#!/usr/bin/env python
"""Run one eager generator step + discriminator step inside an NVTX range for `ncu --nvtx`.
   ncu --set full --nvtx --nvtx-include "prof/" -k regex:<kernels> -c N python tools/profile_step.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2h_b200  # noqa: F401,E402
from b2h_b200.trainer import GanTrainer  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
tr = GanTrainer("v1", 36, 252, False, 256, 64, precision=precision, device="cuda:0")
tr.x.normal_()
tr.y.normal_()
for _ in range(2):
    tr.generator_step()
    tr.discriminator_step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("prof")
tr.generator_step()
tr.discriminator_step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done", tr.losses.cpu().tolist()[:4])
