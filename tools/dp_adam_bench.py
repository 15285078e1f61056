#!/usr/bin/env python
"""Kernel-level timing of the fused data-parallel optimizer step (b2h_dp_adam) against ncclAllReduce + b2h_adam on the
same flat buffer, one process per GPU:

    gpurun --gpus 8 --timeout 300 -- 'timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
        --master-addr 127.0.0.1 --master-port 29512 tools/dp_adam_bench.py > gpurun_out/dp_adam_bench.log 2>&1'

Prints, per size (generator 2 240 864 and discriminator 121 684 parameters by default): device time per call (CUDA
events on the launch stream, max over ranks, after warm-up), the NVLink bytes a rank moves per call
(2 * n * 4 * (world - 1) / world) and the rate that makes, for peer loads / stores and — when the switch has a multicast
object — for the multimem path.
"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b2h_b200  # noqa: E402,F401
from b2h_b200 import _lib as L  # noqa: E402
from b2h_b200.program import _fill_struct  # noqa: E402
from b2h_b200.trainer import PeerBuffers  # noqa: E402


def timed(fn, iters, dev):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)      # us
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L.load()
    iters = int(os.environ.get("ITERS", 200))
    stream = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)  # noqa: E731
    for n in [int(v) for v in os.environ.get("SIZES", "2240864,121684").split(",")]:
        make = PeerBuffers.ipc if os.environ.get("B2H_DP_PEER") == "ipc" else PeerBuffers.symmetric
        pb = make(n, 1, dev, dist.group.WORLD)
        pb.flat.normal_()
        pb.grad.normal_(std=1e-2)
        m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        scal = torch.tensor([-1e-4, 1.0], dtype=torch.float32, device=dev)
        step = torch.zeros(1, dtype=torch.int64, device=dev)
        common = dict(m=m, v=v, n=n, beta1=0.9, beta2=0.999, eps=1e-8, gscale=1.0 / world)

        def fused(mc):
            d = _fill_struct(L.DpAdam(), dict(p=pb.p_ptrs, g=pb.g_ptrs, signal=pb.sig_ptrs, rank=pb.rank, world=world,
                                              g_mc=pb.g_mc if mc else None, p_mc=pb.p_mc if mc else None,
                                              scalars=scal, timeout_ms=5000, **common))
            return lambda: L.run_oneshot(d, L.F32, stream())

        adam = _fill_struct(L.Adam(), dict(p=pb.flat, g=pb.grad, lr=1e-4, step=step, scalars=scal, phase=2, **common))

        def nccl():
            dist.all_reduce(pb.grad, op=dist.ReduceOp.SUM)
            L.run_oneshot(adam, L.F32, stream())

        wire = 2 * n * 4 * (world - 1) / world
        rows = [("ncclAllReduce + b2h_adam", timed(nccl, iters, dev)),
                ("b2h_dp_adam, peer loads / stores", timed(fused(False), iters, dev))]
        if pb.g_mc:
            rows.append(("b2h_dp_adam, multimem (NVLS)", timed(fused(True), iters, dev)))
        if rank == 0:
            print(f"n = {n} fp32 parameters, {world} GPUs, {wire / 1e6:.2f} MB over NVLink per rank and call", flush=True)
            for name, us in rows:
                print(f"  {name:36s} {us:8.1f} us   {wire / us / 1e3:7.1f} GB/s per rank", flush=True)
        del pb
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
