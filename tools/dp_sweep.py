#!/usr/bin/env python
"""Data-parallel exchange variants of the GAN training step inside ONE torchrun job (process start-up and NCCL
initialisation are paid once; multi-GPU minutes are the scarce resource):

    gpurun --gpus 8 -- 'timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
        --master-addr 127.0.0.1 tools/dp_sweep.py > gpurun_out/dp_sweep_n8.txt'

For each (configuration, exchange, buckets): device time per pipelined gan_step (one interval of STEPS steps, CUDA events,
barrier + synchronize on both sides, max over ranks — bench.py's own method) and whether all ranks hold identical
parameters afterwards.  exchange: nccl = ncclAllReduce per gradient bucket + b2h_adam; fused = b2h_dp_adam (reduce-scatter
+ Adam + all-gather in one kernel over NVLink peer memory, multimem when available); fused-p2p = the same without
multicast."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    steps = int(os.environ.get("STEPS", 40))
    cases = os.environ.get("CASES", "v1:0,v1:1,b2h:1").split(",")
    variants = os.environ.get("EXCHANGES", "nccl:1,nccl:2,nccl:3,fused:1,fused:2,fused:3,fused-p2p:1").split(",")
    rows = []
    for case in cases:
        variant, feats = case.split(":")
        feats = feats == "1"
        for v in variants:
            exch, nb = v.split(":")
            if world == 1 and (exch != "nccl" or nb != "1"):
                continue
            for k in ("B2H_FUSED_DP", "B2H_DP_NO_MULTICAST", "B2H_BUCKETS"):
                os.environ.pop(k, None)
            os.environ["B2H_FUSED_DP"] = "1" if exch.startswith("fused") else "0"
            if exch == "fused-p2p":
                os.environ["B2H_DP_NO_MULTICAST"] = "1"
            if world > 1 or nb != "1":
                os.environ["B2H_BUCKETS"] = nb
            t0 = time.time()
            try:
                res, (tr, *_rest) = bench.run_train_config(variant, feats, "bf16", 256, 64, 36, 252, dev, world, rank, pg,
                                                           steps, 5, pipelined=True, keep=True)
                sync = bench.check_ranks_in_sync(tr, world) if world > 1 else None
                tr.release_graphs()
                del tr, _rest
                row = {"case": case, "exchange": exch, "buckets": int(nb), "n_gpus": world,
                       "ms_per_step": round(res["ms_per_step"], 4), "frames_per_s": round(res["value"]),
                       "ranks_in_sync": sync, "wall_s": round(time.time() - t0, 1)}
            except Exception as e:  # noqa: BLE001
                row = {"case": case, "exchange": exch, "buckets": int(nb), "error": repr(e)[:300]}
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
