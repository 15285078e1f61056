#!/usr/bin/env python
"""Benchmark of the hot path: the Body2Hands-style GAN training step (train_gan.py:215-299).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" = one generator step + one discriminator step on one synthetic How2Sign-shaped batch
(BASELINE.json configs[1]: v1 body-only regressor + discriminator, batch 256 x 64 frames per GPU).
metric = training frames/s = B*T*N / step time.  One JSON line on stdout (rank 0).

Besides the headline workload the same line carries (key "configs") short runs of the other BASELINE configs --
the fp32 mode of the same step, the text- and image-conditioned steps, the inference sweep points -- so one
driver run records all of them (`--no-extra-configs` to skip).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

POOL = 8   # resident input batches the timed steps rotate over (8 x 18.9 MB = 151 MB > the 126 MB L2)


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
    ap.add_argument("--no-extra-configs", action="store_true")
    ap.add_argument("--extra-steps", type=int, default=20)
    return ap.parse_args()


def workload_name(a):
    cond = ""
    if a.feats:
        cond = "+image" if a.variant == "b2h" else "+text"
    what = "GAN training step (1 generator step + 1 discriminator step)" if a.mode == "train" else "eval forward"
    cin, cout = pipeline_dims(a)
    return f"{a.variant}{cond} {a.pipeline} {cin}->{cout}, {what}, batch {a.batch} x {a.frames} frames per GPU"


def config_of(a, world):
    """The `config` object: identical in both arms (the reference arm times this arm's workload)."""
    return {"workload": workload_name(a), "global_batch": a.batch * world, "frames": a.frames,
            "parallelism": f"dp{world}" if world > 1 else "single",
            "l2": f"no flush kernels in the timed interval: the steps rotate over {POOL} resident input batches "
                  f"({POOL} x per-step inputs > the 126 MB L2), so no step finds its inputs cached"}


def pipeline_dims(a):
    from b2h_b200.data import FEATURE_MAP
    return FEATURE_MAP[a.pipeline]


def feats_kind_of(variant, feats):
    return None if not feats else ("image" if variant == "b2h" else "text")


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
    feats_kind = feats_kind_of(a.variant, a.feats)
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = a.steps, a.warmup
    fps, med, cores = run_cpu_reference(a, steps, warmup)
    sample = (f"{steps} steps (+{warmup} warm-up), each the full {a.batch}x{a.frames} per-GPU batch of the workload, through "
              f"the oracle restatement of train_gan.py's step bodies on torch CPU fp32, {cores} threads")
    line = {
        "impl": "reference", "metric": "training frames/sec" if a.mode == "train" else "inference frames/sec",
        "value": fps, "unit": "frames/s", "n_gpus": a.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config_of(a, world),
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
def _time_op(prog, idx, flush, reps=5):
    ts = []
    for _ in range(reps):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prog.run_range(idx, idx + 1)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def kernel_breakdown(tr, flush):
    """CUDA-event time of every GEMM-class op of the step, launched alone (L2 flushed before each launch), and its
    algorithmic FLOP/s.  Returns the rows, slowest first (no op is excluded)."""
    from b2h_b200 import _lib as L
    rows = []
    plans = [("G_train", tr.G_train), ("D_eval", tr.D_eval), ("G_eval", tr.G_eval), ("D_train", tr.D_train)]
    for pname, plan in plans:
        for idx, macs in sorted(plan.op_macs.items()):
            rec = plan.prog.recs[idx]
            if not any(s <= idx < e for n, (s, e) in plan.prog.segments.items() if n in ("fwd", "bwd")):
                continue
            ms = _time_op(plan.prog, idx, flush)
            row = {"op": f"{pname}.{rec.tag}", "ms": ms, "gflop": 2 * macs / 1e9,
                   "tflops": 2 * macs / (ms * 1e-3) / 1e12 if ms > 0 else 0.0}
            try:
                pl = plan.prog.op_plan(idx)
                row["tile_n"], row["ctas"] = pl["tile_n"], pl["grid"][0] * pl["grid"][1] * pl["grid"][2]
                if rec.kind == L.OP_WGRAD:
                    row["splits"] = pl["splits"]
            except Exception:
                pass
            rows.append(row)
    rows.sort(key=lambda r: -r["ms"])
    return rows


def hbm_records(tr, flush, peak_gbs):
    """The bandwidth-bound kernels of the step launched alone (L2 flushed before): achieved GB/s of their ALGORITHMIC
    bytes (L1 12 B/element in fp32 mode, Adam 28 B/parameter, BN-apply read z + write a + keep flags)."""
    from b2h_b200 import _lib as L
    esz = 2 if tr.dtype == L.BF16 else 4
    out = []

    def add(name, prog, idx, nbytes, what):
        ms = _time_op(prog, idx, flush)
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "launch_ms": round(ms, 5), "algorithmic_bytes": int(nbytes),
                    "achieved": round(gbs, 1), "peak": peak_gbs, "unit": "GB/s", "frac": round(gbs / peak_gbs, 4),
                    "bytes": what})

    for i, rec in enumerate(tr.g_loss_prog.recs):
        if rec.kind == L.OP_L1:
            n = rec.f["B"] * rec.f["C"] * rec.f["L"]
            if rec.f.get("out_blc") is not None:
                add("l1_kernel (" + rec.tag + ", reads the output layer's BLC tile, writes NCL out)", tr.g_loss_prog, i,
                    n * (4 + 4 + 4 + esz), f"read out_blc fp32 + read gt fp32 + write out fp32 + write dout {esz} B, per element")
            else:
                add("l1_kernel (" + rec.tag + ")", tr.g_loss_prog, i, n * (4 + 4 + esz),
                    f"read out fp32 + read gt fp32 + write dout {esz} B, per element")
        if rec.kind == L.OP_ADAM and rec.f.get("phase", 0) == 0:
            add("adam_kernel (generator, whole flat buffer)", tr.g_loss_prog, i, rec.f["n"] * 28,
                "28 B/parameter: read p, g, m, v; write p, m, v")
    big = None
    for i, rec in enumerate(tr.G_train.prog.recs):
        if rec.kind == L.OP_BN_APPLY:
            n = rec.f["B"] * rec.f["L"] * rec.f["C"] * rec.f["nsrc"]
            if big is None or n > big[1]:
                big = (i, n, rec)
    if big is not None:
        i, _, rec = big
        elems = rec.f["B"] * rec.f["L"] * rec.f["C"]
        nbytes = elems * (esz * rec.f["nsrc"] + esz + (1 if rec.f["drop"].get("save") is not None else 0))
        add(f"bn_apply_kernel ({rec.tag})", tr.G_train.prog, i, nbytes,
            f"read {rec.f['nsrc']} x z + write a ({esz} B each) + 1 B keep flag, per element")
    return out


def make_pool(B, T, cin, cout, feats_kind, dev, seed):
    xs, ys, fs = [], [], []
    for k in range(POOL):
        x, y, f = synth_batch(B, T, cin, cout, feats_kind, seed=seed + 1000 * k)
        xs.append(x.to(dev))
        ys.append(y.to(dev))
        fs.append(f.to(dev) if f is not None else None)
    return xs, ys, fs


def interval_ms(fn_step, steps, world):
    """K steps as ONE device-timed interval (CUDA events on the launch stream, barrier + synchronize on both sides)."""
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        fn_step(k)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return e0.elapsed_time(e1)


def check_ranks_in_sync(tr, world):
    """Data-parallel sanity: identical initial weights + summed gradients -> identical weights on every rank."""
    import torch.distributed as dist
    chk = torch.stack([tr.g_store.flat.double().sum(), tr.g_store.flat.double().abs().sum(),
                       tr.d_store.flat.double().sum(), tr.d_store.flat.double().abs().sum()])
    hi, lo = chk.clone(), chk.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return bool(((hi - lo).abs() <= 1e-9 * hi.abs()).all()) and bool(torch.isfinite(chk).all())


def max_over_ranks(v, dev, world):
    if world <= 1:
        return float(v)
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def run_train_config(variant, feats, precision, B, T, cin, cout, dev, world, rank, pg, steps, warmup, pipelined=True,
                     keep=False):
    """GAN training step of one configuration: builds the trainer, warms up, times `steps` steps as one interval.
    Returns (result dict, trainer or None)."""
    import torch.distributed as dist
    from b2h_b200.trainer import GanTrainer
    kw = {}
    if os.environ.get("B2H_BUCKETS"):
        kw["n_buckets"] = int(os.environ["B2H_BUCKETS"])
    tr = GanTrainer(variant, cin, cout, feats, B, T, precision=precision, device=dev, lr=1e-4, seed=23456 + rank,
                    drop_mode="philox", world_size=world, process_group=pg, **kw)
    if world > 1:   # identical initial weights on every rank (DDP convention)
        for st in (tr.g_store, tr.d_store):
            dist.broadcast(st.flat, 0)
            dist.broadcast(st.bufs, 0)
    fk = feats_kind_of(variant, feats)
    xs, ys, fs = make_pool(B, T, cin, cout, fk, dev, seed=23456 + rank)
    tr.load_batch(xs[0], ys[0], fs[0])

    def step(k):
        j = (k + 1) % POOL
        if pipelined:
            tr.advance_batch(xs[j], ys[j], fs[j])   # the batch G just saw goes to D, the next one to G
            tr.gan_step(graph=True)                 # D step on the previous batch || G step on the current one
        else:
            tr.load_batch(xs[j], ys[j], fs[j])
            tr.generator_step(graph=True)
            tr.discriminator_step(graph=True)

    if pipelined:                             # pipeline prologue: G0; every timed step is then [D_k || G_k+1]
        tr.generator_step(graph=True)
    for k in range(max(warmup, 3)):
        step(k)
    ms = max_over_ranks(interval_ms(step, steps, world), dev, world)
    res = {"ms_per_step": ms / steps, "value": B * T * world * steps / (ms * 1e-3), "steps": steps,
           "dtype": "bf16" if precision == "bf16" else "f32", "n_gpus": world}
    if not keep:
        tr.release_graphs()
        tr = None
    return res, (tr, xs, ys, fs)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def run_infer_config(variant, feats, precision, B, T, cin, cout, dev, world, steps, warmup):
    """Batched eval forward (inference.py:96-121): inputs resident, rotating over POOL batches."""
    from b2h_b200 import _lib as L
    from b2h_b200 import nets
    spec = nets.generator_spec(variant, cin, cout, feats, train=False)
    store = nets.ParamStore(spec, dev, seed=23456)
    plan = nets.NetPlan(spec, store, B, T, L.BF16 if precision == "bf16" else L.F32, dev, train=False)
    fk = feats_kind_of(variant, feats)
    g = torch.Generator(device=dev).manual_seed(7)
    xs = [torch.randn(B, cin, T, device=dev, generator=g) for _ in range(POOL)]
    fs = [None] * POOL
    if fk == "text":
        fs = [torch.nn.functional.normalize(torch.randn(B, 512, device=dev, generator=g), dim=1) for _ in range(POOL)]
    elif fk == "image":
        fs = [torch.randn(B, T, 2000, device=dev, generator=g) * 2 for _ in range(POOL)]

    def step(k):
        j = k % POOL
        plan.x.copy_(xs[j], non_blocking=True)
        if fs[j] is not None:
            plan.feats.copy_(fs[j], non_blocking=True)
        plan.forward()

    for k in range(max(warmup, 3)):
        step(k)
    ms = max_over_ranks(interval_ms(step, steps, world), dev, world)
    launches = plan.prog.segment_launches.get("fwd", 0)
    gflop = 2 * sum(plan.op_macs.values()) / 1e9          # algorithmic FLOP of one forward (no padding)
    tflops = gflop / (ms / steps)                          # per GPU
    peak = load_peaks().get("bf16_tflops", 1590.0) / (1.0 if precision == "bf16" else 6.0)   # fp32 mode: 3 x TF32
    return {"ms_per_step": ms / steps, "value": B * T * world * steps / (ms * 1e-3), "steps": steps,
            "dtype": "bf16" if precision == "bf16" else "f32", "n_gpus": world, "gpu_launches_per_step": launches,
            "algorithmic_gflop": round(gflop, 3), "tflops_per_gpu": round(tflops, 1),
            "frac_of_tensor_peak": round(tflops / peak, 4)}


def algorithmic_gflop_per_step(tr):
    """generator step 3 x forward MACs minus the first layer's dgrad plus the discriminator scoring pass;
    discriminator step = generator forward + 2 x (D forward + backward minus the first layer's dgrad)."""
    g_fwd = sum(tr.G_train._layer_macs(l) for l in tr.G_train.spec.layers)
    g_first = sum(tr.G_train._layer_macs(l) for l in tr.G_train.spec.layers if not tr.G_train._needs_dgrad(l))
    d_fwd_1 = sum(tr.D_eval._layer_macs(l) for l in tr.D_eval.spec.layers)             # B clips
    d_first_1 = sum(tr.D_eval._layer_macs(l) for l in tr.D_eval.spec.layers if not tr.D_eval._needs_dgrad(l))
    g_step = 3 * g_fwd - g_first + d_fwd_1
    d_step = g_fwd + 2 * d_fwd_1 + 2 * (2 * d_fwd_1 - d_first_1)
    return 2 * (g_step + d_step) / 1e9


def bind_to_gpu_numa_node(index):
    """Run this process on the CPUs next to its GPU (sysfs local_cpulist of the device's PCI function): on a two-socket
    host a process that lands on the far socket pays for every launch and every pinned-host copy across the socket
    link (observed: the end-to-end loop at 1.2 instead of 0.76 ms per step in about one run of six).  Best effort: a
    container without the sysfs information keeps its affinity.  Returns what it did."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return f"bound to {len(use)} of {len(allowed)} CPUs next to GPU {index} ({bdf})"
        return f"all {len(allowed)} allowed CPUs are local to GPU {index}" if use else "no local CPU in the allowed set"
    except Exception as ex:   # noqa: BLE001
        return f"not bound ({type(ex).__name__})"


def main():
    a = parse()
    if a.impl == "reference":
        return reference_main(a)
    import torch.distributed as dist
    import b2h_b200  # noqa: F401

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == a.gpus or world == 1, f"--gpus {a.gpus} but WORLD_SIZE={world}"
    affinity = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    cin, cout = pipeline_dims(a)
    feats_kind = feats_kind_of(a.variant, a.feats)
    cfg = config_of(a, world)   # (before a.batch becomes the per-GPU batch under strong scaling)
    if a.scaling == "strong":
        assert a.batch % world == 0, "--scaling strong: --batch must be divisible by the number of GPUs"
        a.batch //= world                     # from here on a.batch is the per-GPU batch
        cfg["global_batch"] = a.batch * world
    B, T = a.batch, a.frames
    pipelined = a.mode == "train" and a.schedule == "pipelined"
    sampler = ClockSampler(local)
    sampler.start()
    t_wall0 = time.perf_counter()
    tr = None
    if a.mode == "train":
        res, (tr, xs, ys, fs) = run_train_config(a.variant, a.feats, a.precision, B, T, cin, cout, dev, world, rank, pg,
                                                 a.steps, a.warmup, pipelined=pipelined, keep=True)
    else:
        res = run_infer_config(a.variant, a.feats, a.precision, B, T, cin, cout, dev, world, a.steps, a.warmup)
    t_wall = time.perf_counter() - t_wall0
    # ---- end to end: pinned host buffers -> H2D -> step -> D2H of the losses, every step
    e2e_value = h2d = d2h = None
    if a.mode == "train":
        x, y, f = synth_batch(B, T, cin, cout, feats_kind, seed=99 + rank)
        hx, hy = x.pin_memory(), y.pin_memory()
        hf = f.pin_memory() if f is not None else None
        # the losses of every step are read back into a ring of two pinned buffers; the host waits for the read of
        # step k-1 while step k is already enqueued (a training loop that logs its losses never needs step k's value
        # before it has launched step k+1), so launch latency and the copy's completion are off the device's chain
        h_loss = [torch.empty(8, dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_read = [None, None]

        def step():
            if pipelined:
                tr.gan_step(graph=True)
            else:
                tr.generator_step(graph=True)
                tr.discriminator_step(graph=True)

        for _ in range(5):                     # untimed: creates the copy stream and the staging buffers
            tr.prefetch_batch(hx, hy, hf)
            tr.swap_batch(pipelined=pipelined)
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e2e_runs = []
        for _rep in range(7):                  # host wall clock jitters on a shared box (the first K-step runs of a
            # process are slower in about one process of three -- 19-22 ms against 15.3 ms for 20 steps, decaying: host
            # clocks, page pinning, copy engines ramp up): two untimed K-step rehearsals, then the median of five runs
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            tr.prefetch_batch(hx, hy, hf)      # batch 0; every later batch is copied while the previous one trains
            loss_sum = 0.0
            for k in range(a.steps):
                tr.swap_batch(pipelined=pipelined)
                tr.prefetch_batch(hx, hy, hf)  # H2D of the next step's inputs, pinned host -> staging, copy stream
                step()
                h_loss[k & 1].copy_(tr.losses, non_blocking=True)
                loss_read[k & 1] = torch.cuda.Event()
                loss_read[k & 1].record()
                if k > 0:                      # step k-1's losses are on the host now: use them
                    loss_read[(k - 1) & 1].synchronize()
                    loss_sum += float(h_loss[(k - 1) & 1][2])
            loss_read[(a.steps - 1) & 1].synchronize()
            loss_sum += float(h_loss[(a.steps - 1) & 1][2])
            torch.cuda.synchronize()
            if _rep >= 2:
                e2e_runs.append(max_over_ranks((time.perf_counter() - t0) * 1e3, dev, world))
        e2e_ms = statistics.median(e2e_runs)
        e2e_value = B * T * world * a.steps / (e2e_ms * 1e-3)
        h2d = (hx.numel() * 4 + hy.numel() * 4 + (hf.numel() * 4 if hf is not None else 0)) * world
        d2h = 32 * world
    clocks = sampler.stop()
    ranks_in_sync = check_ranks_in_sync(tr, world) if (world > 1 and tr is not None) else None
    launches = tr.launches_per_gan_step() if tr is not None else res.get("gpu_launches_per_step", 0)
    line = {
        "metric": "training frames/sec" if a.mode == "train" else "inference frames/sec",
        "value": res["value"], "unit": "frames/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": a.precision if a.precision == "bf16" else "f32", "data": "synthetic",
        "config": cfg,
        "method": {"timing": "the K steps are ONE interval between two CUDA events on the launch stream (barrier + "
                             "synchronize on both sides, max over ranks); CUDA-graph replay; dropout = Philox",
                   "schedule": ("pipelined: every timed step = discriminator step k overlapped with generator step "
                                "k+1 (independent work: same results as the alternating order)") if pipelined else
                               "sequential: generator step then discriminator step",
                   **({"dp_exchange": "b2h_dp_adam: reduce-scatter + Adam + all-gather in one kernel over peer memory"
                       if getattr(tr, "fused_dp", False) else "ncclAllReduce of the flat gradient, then b2h_adam"}
                      if world > 1 and tr is not None else {})},
        "clocks": clocks,
        "gpu_launches": int(launches) * a.steps,
        "ranks_in_sync": ranks_in_sync,
        "wall_s": t_wall,
    }
    if e2e_value is not None:
        line["e2e"] = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "how": "host wall clock over K calls of the public step API; every step: pinned-host inputs -> "
                              "device (copy stream, under the previous step), step, losses -> pinned host; the host "
                              "consumes the losses of step k-1 after enqueuing step k (all K read inside the interval); "
                              "median of five K-step runs after two untimed K-step rehearsals",
                       "runs_ms": [round(v, 3) for v in e2e_runs],
                       "mean_g_loss_read_on_host": loss_sum / a.steps, "host_affinity": affinity}
    # ---- the other BASELINE configs, short runs, every rank takes part (data parallel where they train)
    if not a.no_extra_configs and a.mode == "train":
        extras = {}
        ks, kw_ = a.extra_steps, 3
        todo = [("config2_fp32_mode", "train", "v1", False, "fp32", 256, 64),
                ("config3_text_bf16", "train", "v1", True, "bf16", 256, 64),
                ("config4_image_bf16", "train", "b2h", True, "bf16", 256, 64),
                ("config1_eval_32x64_fp32", "infer", "v1", False, "fp32", 32, 64),
                ("config5_infer_4096x64_bf16", "infer", "v1", False, "bf16", 4096, 64),
                ("config5_infer_64x1024_bf16", "infer", "v1", False, "bf16", 64, 1024),
                ("config5_infer_v2text_4096x64_bf16", "infer", "v2", True, "bf16", 4096, 64),
                ("config5_infer_4096x64_fp32", "infer", "v1", False, "fp32", 4096, 64)]
        for name, mode, variant, feats, prec, eb, et in todo:
            try:
                if mode == "train":
                    r, _ = run_train_config(variant, feats, prec, eb, et, 36, 252, dev, world, rank, pg, ks, kw_)
                else:
                    r = run_infer_config(variant, feats, prec, eb, et, 36, 252, dev, world, ks, kw_)
                cond = "" if not feats else ("+image" if variant == "b2h" else "+text")
                r["workload"] = (f"{variant}{cond} arm2wh 36->252, " + ("GAN training step" if mode == "train" else
                                 "eval forward") + f", batch {eb} x {et} frames per GPU")
                r["unit"] = "frames/s"
                extras[name] = {k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items()}
            except Exception as ex:   # an extra must never take the headline line down
                extras[name] = {"error": str(ex)[:300]}
            torch.cuda.empty_cache()
        line["configs"] = extras
    if rank == 0:
        peaks = load_peaks()
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

        def flush():
            flush_buf.fill_(1)

        if not a.no_kernel_breakdown and tr is not None:
            rows = kernel_breakdown(tr, flush)
            dom = rows[0]                                 # the time-dominant launch, whatever it is
            peak = peaks.get("bf16_tflops", 1590.0)
            which = "measured burst (MEASURED_PEAKS.json bf16_tflops)" if "bf16_tflops" in peaks else "fallback 1.59 PF"
            if a.precision != "bf16":
                # fp32 mode = 3 x TF32 passes on the tensor cores: the ceiling is a third of the TF32 rate (half of bf16)
                peak, which = peak / 6.0, which + " / 6 (TF32 = half the bf16 rate, three passes per product)"
            traffic = None
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(dom["op"])
            except Exception:
                pass
            gflop_step = algorithmic_gflop_per_step(tr)
            step_tflops = gflop_step / res["ms_per_step"]          # GFLOP / ms = TFLOP/s
            line["roofline"] = {"bound": "tensor", "kernel": dom["op"], "achieved": dom["tflops"], "peak": peak,
                                "unit": "TFLOP/s", "frac": dom["tflops"] / peak, "traffic": traffic,
                                "peak_source": which, "launch_ms": dom["ms"], "algorithmic_gflop": dom["gflop"],
                                "step_frac": step_tflops / peak, "step_tflops": step_tflops,
                                "step_algorithmic_gflop": gflop_step,
                                "how": "kernel = the slowest GEMM-class launch of the step, nothing excluded; CUDA events "
                                       "around the launch alone, L2 flushed before it (cold operands; in the step they "
                                       "are L2-resident), median of 5.  step_frac = algorithmic FLOP of the whole GAN "
                                       "step / ms_per_step / peak",
                                "hbm": hbm_records(tr, flush, peaks.get("hbm_gbs", 6549.4))}
            line["kernel_breakdown"] = [{k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items()}
                                        for r in rows[:14]]
            line["gemm_ms_sum"] = sum(r["ms"] for r in rows)
        if not a.no_cpu_baseline and world == 1:   # (N > 1: the other ranks' host threads would share the cores)
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
        # teardown: the captured graphs hold NCCL kernels -- drop them before the communicator goes away.  A watchdog
        # ends the process if the teardown itself wedges (the result line is already out).
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        dist.barrier()           # rank 0 has finished its single-rank measurements (breakdown, CPU baseline)
        threading.Thread(target=lambda: (time.sleep(45), sys.stderr.write("bench: teardown timed out\n"), os._exit(0)),
                         daemon=True).start()
        if tr is not None:
            tr.release_graphs()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
