#!/usr/bin/env python
"""Benchmark of the hot path: the Body2Hands-style GAN training step (train_gan.py:215-299).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" = one generator step + one discriminator step on one synthetic How2Sign-shaped batch
(BASELINE.json configs[1]: v1 body-only regressor + discriminator, batch 256 x 64 frames per GPU).
metric = training frames/s = B*T*N / step time.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--variant", default="v1")
    ap.add_argument("--feats", action="store_true", help="text (v1/v2/v4) or image (b2h) conditioning")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): --batch clips per GPU; strong: --batch clips in total, split over the GPUs "
                         "(SURVEY 8d C3: strong scaling at global B = 256)")
    ap.add_argument("--pipeline", default="arm2wh", help="FEATURE_MAP row (utils/constants.py:11-27): arm2wh, "
                    "arm_wh2wh, wh2wh, arm_wh2finger1..12 -- the incremental-fingers models of BASELINE config 5")
    ap.add_argument("--schedule", default="pipelined", choices=["pipelined", "sequential"],
                    help="pipelined: each discriminator step overlaps the next generator step (GanTrainer.gan_step); "
                         "sequential: generator_step then discriminator_step on the same batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-breakdown", action="store_true")
    return ap.parse_args()


def workload_name(a):
    cond = ""
    if a.feats:
        cond = "+image" if a.variant == "b2h" else "+text"
    what = "GAN training step (1 generator step + 1 discriminator step)" if a.mode == "train" else "eval forward"
    cin, cout = pipeline_dims(a)
    return f"{a.variant}{cond} {a.pipeline} {cin}->{cout}, {what}, batch {a.batch} x {a.frames} frames per GPU"


def pipeline_dims(a):
    from b2h_b200.data import FEATURE_MAP
    return FEATURE_MAP[a.pipeline]


def synth_batch(B, T, cin, cout, feats_kind, seed=23456):
    """How2Sign-shaped synthetic batch: standardised 6-D rotations with a temporal random walk."""
    rng = np.random.RandomState(seed)
    base = rng.randn(B, cin + cout, 1).astype(np.float32)
    walk = np.cumsum(rng.randn(B, cin + cout, T).astype(np.float32) * 0.05, axis=2)
    data = base + walk
    data = (data - data.mean(axis=(0, 2), keepdims=True)) / (data.std(axis=(0, 2), keepdims=True) + 1e-6)
    x = torch.from_numpy(np.ascontiguousarray(data[:, :cin]))
    y = torch.from_numpy(np.ascontiguousarray(data[:, cin:]))
    f = None
    if feats_kind == "text":
        f = rng.randn(B, 512).astype(np.float32)
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        f = torch.from_numpy(f)
    elif feats_kind == "image":
        f = torch.from_numpy((rng.randn(B, T, 2000) * 2).astype(np.float32))
    return x, y, f


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle restatement of train_gan's step bodies on host cores
# ------------------------------------------------------------------------------------------------
def run_cpu_reference(a, steps, warmup, device="cpu", autocast=False):
    """The oracle restatement of the reference step through stock PyTorch: on the host cores (the baseline), or —
    device="cuda" — through PyTorch eager / cuDNN on this GPU ("the honest bar to beat")."""
    from oracle import ref_models as R
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(23456)
    cin, cout = pipeline_dims(a)
    feats_kind = None if not a.feats else ("image" if a.variant == "b2h" else "text")
    G = R.build_generator(a.variant, cin, cout, a.feats).to(device)
    D = R.build_discriminator(cout).to(device)
    x, y, f = synth_batch(a.batch, a.frames, cin, cout, feats_kind)
    x, y = x.to(device), y.to(device)
    f = f.to(device) if f is not None else None
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    g_opt = torch.optim.Adam(G.parameters(), lr=1e-4)
    d_opt = torch.optim.Adam(D.parameters(), lr=1e-4)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            if a.mode == "train":
                R.generator_step(G, D, g_opt, x, y, f)
                R.discriminator_step(G, D, d_opt, x, y, f)
            else:
                G.eval()
                with torch.no_grad():
                    G(x, feats_=f)

    for _ in range(warmup):
        step()
    sync()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        sync()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return a.batch * a.frames / med, med, os.cpu_count() or 1


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(a.steps, 8)
    warmup = min(a.warmup, 2)
    fps, med, cores = run_cpu_reference(a, steps, warmup)
    sample = f"{steps} steps (+{warmup} warm-up) of the full {a.batch}x{a.frames} batch, torch CPU, {cores} threads"
    line = {
        "impl": "reference", "metric": "training frames/sec" if a.mode == "train" else "inference frames/sec",
        "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a)},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.thread = None
        self.samples = []

    def _nvml_loop(self, h, nv):
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        # NVML in a thread of this process (4 ms period: the timed region is ~100 ms); nvidia-smi as a fallback
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML indexes physical devices: map through CUDA_VISIBLE_DEVICES via the PCI bus id
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(nv.nvmlDeviceGetCount()):
                    hi = nv.nvmlDeviceGetHandleByIndex(i)
                    if int(nv.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                        h = hi
                        break
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self._stop = False
            self.thread = threading.Thread(target=self._nvml_loop, args=(h, nv), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            sm = [s for s, _ in self.samples]
            reasons = sorted({r for _, rs in self.samples for r in rs})
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 4 ms period over the timed + e2e regions"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def kernel_breakdown(tr, flush, reps=5):
    """CUDA-event time of every GEMM-class op of the step (L2 flushed before each launch) and its
    algorithmic FLOP/s; returns (list, dominant entry)."""
    from b2h_b200 import _lib as L
    rows = []
    plans = [("G_train", tr.G_train), ("D_eval", tr.D_eval), ("G_eval", tr.G_eval), ("D_train", tr.D_train)]
    for pname, plan in plans:
        for idx, macs in sorted(plan.op_macs.items()):
            rec = plan.prog.recs[idx]
            seg_ok = any(s <= idx < e for n, (s, e) in plan.prog.segments.items() if n in ("fwd", "bwd"))
            if not seg_ok:
                continue
            ts = []
            for _ in range(reps):
                flush()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                plan.prog.run_range(idx, idx + 1)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = statistics.median(ts)
            row = {"op": f"{pname}.{rec.tag}", "ms": ms, "gflop": 2 * macs / 1e9,
                   "tflops": 2 * macs / (ms * 1e-3) / 1e12 if ms > 0 else 0.0}
            if rec.kind == L.OP_WGRAD and plan.wgrad_direct and rec.f.get("splits") == 1:
                # split-free weight gradient: a dozen CTAs by design (it runs beside the backward chain on a side
                # stream); its duration alone says nothing about the GPU's tensor throughput
                row["side_stream_few_ctas"] = True
            rows.append(row)
    return rows


def main():
    a = parse()
    if a.impl == "reference":
        return reference_main(a)
    import torch.distributed as dist
    import b2h_b200  # noqa: F401
    from b2h_b200.trainer import GanTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == a.gpus or world == 1, f"--gpus {a.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    cin, cout = pipeline_dims(a)
    feats_kind = None if not a.feats else ("image" if a.variant == "b2h" else "text")
    if a.scaling == "strong":
        assert a.batch % world == 0, "--scaling strong: --batch must be divisible by the number of GPUs"
        a.batch //= world                     # from here on a.batch is the per-GPU batch
    B, T = a.batch, a.frames
    kw = {}
    if os.environ.get("B2H_BUCKETS"):
        kw["n_buckets"] = int(os.environ["B2H_BUCKETS"])
    tr = GanTrainer(a.variant, cin, cout, a.feats, B, T, precision=a.precision, device=dev, lr=1e-4, seed=23456 + rank,
                    drop_mode="philox", world_size=world, process_group=pg, **kw)
    if world > 1:   # identical initial weights on every rank (DDP convention)
        for st in (tr.g_store, tr.d_store):
            dist.broadcast(st.flat, 0)
            dist.broadcast(st.bufs, 0)
    x, y, f = synth_batch(B, T, cin, cout, feats_kind, seed=23456 + rank)
    hx, hy = x.pin_memory(), y.pin_memory()
    hf = f.pin_memory() if f is not None else None
    tr.x.copy_(hx)
    tr.y.copy_(hy)
    if hf is not None:
        tr.feats.copy_(hf)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def flush():
        flush_buf.fill_(1)

    use_graph = True

    pipelined = a.mode == "train" and a.schedule == "pipelined"

    def step():
        if pipelined:
            tr.gan_step(graph=use_graph)      # D step on the previous batch || G step on the current one
        elif a.mode == "train":
            tr.generator_step(graph=use_graph)
            tr.discriminator_step(graph=use_graph)
        else:
            tr.infer()

    if pipelined:                             # pipeline prologue: G0; every timed step is then [D_k || G_k+1]
        tr.generator_step(graph=use_graph)
        tr._sync_d_batch()
    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    torch.cuda.synchronize()
    ev = []
    t_wall0 = time.perf_counter()
    for _ in range(a.steps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        ev.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    # ---- end to end: pinned host buffers -> H2D -> step -> D2H of the losses, every step
    h_loss = torch.empty(8, dtype=torch.float32).pin_memory()
    for _ in range(2):                     # untimed: creates the copy stream and the staging buffers
        tr.prefetch_batch(hx, hy, hf)
        tr.swap_batch(pipelined=pipelined)
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    if a.mode != "train":
        # inference: every step's prediction goes back to pinned host memory; the D2H of step k runs on its own
        # stream from a device staging copy while step k+1 computes (the timed region ends when all have landed)
        h_out = torch.empty(tr.G_eval.out.shape, dtype=torch.float32).pin_memory()
        stage_out = torch.empty_like(tr.G_eval.out)
        d2h = torch.cuda.Stream(dev)
        d2h_done = None
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.prefetch_batch(hx, hy, hf)          # batch 0; every later batch is copied while the previous one trains
    for _ in range(a.steps):
        tr.swap_batch(pipelined=pipelined)
        tr.prefetch_batch(hx, hy, hf)      # H2D of the next step's inputs, pinned host -> staging, copy stream
        step()
        if a.mode == "train":
            h_loss.copy_(tr.losses, non_blocking=True)
            torch.cuda.synchronize()
        else:
            cur = torch.cuda.current_stream(dev)
            if d2h_done is not None:
                cur.wait_event(d2h_done)             # the staging copy is free again
            stage_out.copy_(tr.G_eval.out, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cur)
            d2h.wait_event(ready)
            with torch.cuda.stream(d2h):
                h_out.copy_(stage_out, non_blocking=True)
            d2h_done = torch.cuda.Event()
            d2h_done.record(d2h)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    ranks_in_sync = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # data-parallel sanity: identical initial weights + summed gradients -> identical weights on every rank
        chk = torch.stack([tr.g_store.flat.double().sum(), tr.g_store.flat.double().abs().sum(),
                           tr.d_store.flat.double().sum(), tr.d_store.flat.double().abs().sum()])
        hi, lo = chk.clone(), chk.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        ranks_in_sync = bool(((hi - lo).abs() <= 1e-9 * hi.abs()).all()) and bool(torch.isfinite(chk).all())
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    frames = B * T * world * a.steps
    value = frames / (dev_ms * 1e-3)
    e2e_value = frames / (e2e_ms * 1e-3)
    # whole-job bytes per step (every rank copies its own batch)
    h2d = (hx.numel() * 4 + hy.numel() * 4 + (hf.numel() * 4 if hf is not None else 0)) * world
    d2h = (32 if a.mode == "train" else tr.G_eval.out.numel() * 4) * world
    launches = tr.launches_per_gan_step() if a.mode == "train" else tr.G_eval.prog.segment_launches.get("fwd", 0)
    line = {
        "metric": "training frames/sec" if a.mode == "train" else "inference frames/sec",
        "value": value, "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": a.precision if a.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": B * world, "frames": T,
                   "parallelism": f"dp{world}" if world > 1 else "single",
                   **({"dp_exchange": "b2h_dp_adam: reduce-scatter + Adam + all-gather in one kernel over peer memory"
                       if getattr(tr, "fused_dp", False) else "ncclAllReduce of the flat gradient, then b2h_adam"}
                      if world > 1 else {}),
                   "timing": "CUDA events per step on the launch stream, 256 MiB L2 flush before every timed step, "
                             "CUDA-graph replay, dropout = Philox",
                   "schedule": ("pipelined: every timed step = discriminator step k overlapped with generator step "
                                "k+1 (independent work: same results as the alternating order)") if pipelined else
                               "sequential: generator step then discriminator step"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches) * a.steps,
        "ranks_in_sync": ranks_in_sync,
        "wall_s": t_wall,
    }
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        if not a.no_kernel_breakdown and a.mode == "train":
            rows = kernel_breakdown(tr, flush)
            rows.sort(key=lambda r: -r["ms"])
            dom = [r for r in rows if not r.get("side_stream_few_ctas")][0]   # dominant full-grid GEMM launch
            peak = peaks.get("bf16_tflops", 1590.0)
            which = "measured burst (MEASURED_PEAKS.json bf16_tflops)" if "bf16_tflops" in peaks else "fallback 1.59 PF"
            traffic = None
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(dom["op"])
            except Exception:
                pass
            line["roofline"] = {"bound": "tensor", "kernel": dom["op"], "achieved": dom["tflops"], "peak": peak,
                                "unit": "TFLOP/s", "frac": dom["tflops"] / peak, "traffic": traffic,
                                "peak_source": which, "launch_ms": dom["ms"], "algorithmic_gflop": dom["gflop"],
                                "how": "CUDA events around the launch alone, L2 flushed before it (cold operands; in "
                                       "the step they are L2-resident), median of 5; the split-free side-stream "
                                       "wgrads (a dozen CTAs by design) are excluded from the choice"}
            line["kernel_breakdown"] = [{k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items()}
                                        for r in rows[:12]]
            line["gemm_ms_sum"] = sum(r["ms"] for r in rows)
        if not a.no_cpu_baseline:
            fps, med, cores = run_cpu_reference(a, steps=5, warmup=2)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"5 steps (+2 warm-up) of the full {B}x{T} batch through the oracle "
                                              f"restatement of train_gan (torch CPU fp32, {cores} threads)"}
            try:   # the same port through stock PyTorch eager (cuDNN) on this GPU: the honest bar
                e32, _, _ = run_cpu_reference(a, steps=20, warmup=5, device="cuda")
                e16, _, _ = run_cpu_reference(a, steps=20, warmup=5, device="cuda", autocast=True)
                line["cpu_baseline"]["same_port_torch_eager_on_this_gpu"] = {
                    "fp32_frames_per_s": e32, "bf16_autocast_frames_per_s": e16,
                    "note": "oracle restatement, torch eager + cuDNN, host-timed with a sync per step, 20 steps"}
            except Exception as ex:   # never let the extra baseline break the bench line
                line["cpu_baseline"]["same_port_torch_eager_on_this_gpu"] = {"error": str(ex)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL kernels keep the communicator busy at interpreter teardown: leave
        # without destroying the process group (every rank has finished its work and its output by now)
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
