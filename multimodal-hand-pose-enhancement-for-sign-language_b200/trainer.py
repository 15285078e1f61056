"""GAN training step (train_gan.py:215-299) and batched inference (inference.py:96-121) as recorded
programs of libb2h ops, replayed on one stream and captured in CUDA graphs.

  generator step   G fwd (train) -> calc_motion -> D fwd (eval, no grad) -> L1(out, gt) + MSE(score, 1)
                   -> G bwd -> [bucketed NCCL all-reduce of the flat gradient, overlapped] -> fused Adam
  discriminator    G fwd (eval, no grad) -> calc_motion(fake), calc_motion(gt) -> ONE grouped D fwd
  step             (train; separate BN statistics / dropout per group, train_gan.py:240-241) -> MSE+MSE
                   -> D bwd -> all-reduce -> fused Adam

Reference quirks kept: calc_motion is frame0 - frames[0..T-2]; the adversarial term of
the generator loss adds a value but no gradient; there is no velocity term.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import nets
from .program import Program


def dtype_of(precision: str) -> int:
    if precision not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    return L.BF16 if precision == "bf16" else L.F32


class FlatAdam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) over a ParamStore's flat buffers.
    `state_dict()` / `load_state_dict()` speak torch.optim.Adam's format (train_gan.py:356-370)."""

    def __init__(self, store: nets.ParamStore, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8):
        self.store, self.lr, self.betas, self.eps = store, float(lr), (float(betas[0]), float(betas[1])), float(eps)
        dev = store.device
        self.m = torch.zeros_like(store.flat)
        self.v = torch.zeros_like(store.flat)
        self.step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.scalars = torch.zeros(2, dtype=torch.float32, device=dev)
        self._recorded = []     # (program, op index) of every optimizer op that carries the hyper-parameters
        self.on_hparams_changed = []   # callbacks (the trainer drops its captured graphs)

    def record(self, prog: Program, gscale: float = 1.0, lo: int = 0, hi: Optional[int] = None, phase: int = 0,
               tag: str = "adam"):
        """phase 0: the whole update.  phase 1: only advance the step / bias corrections.  phase 2: update the
        flat range [lo, hi) with the current bias corrections (gradient buckets)."""
        hi = self.store.n if hi is None else hi
        i = prog.add(L.OP_ADAM, tag, p=self.store.flat[lo:hi], g=self.store.grad[lo:hi], m=self.m[lo:hi],
                     v=self.v[lo:hi], n=max(hi - lo, 0 if phase == 1 else 1), lr=self.lr, beta1=self.betas[0],
                     beta2=self.betas[1], eps=self.eps, gscale=float(gscale), step=self.step, scalars=self.scalars,
                     phase=phase)
        self._recorded.append((prog, i))
        return i

    def set_hparams(self, lr: float, betas, eps: float):
        """torch.optim.Adam.load_state_dict applies the checkpoint's lr / betas / eps (train_gan.py:72,91).  The
        recorded optimizer ops carry them in their descriptors: patch the records, have the programs lowered again
        on their next run and the captured graphs (which froze the old kernel arguments) dropped."""
        lr, betas, eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        if (lr, betas, eps) == (self.lr, self.betas, self.eps):
            return
        self.lr, self.betas, self.eps = lr, betas, eps
        for prog, i in self._recorded:
            f = prog.recs[i].f
            if "lr" in f:
                f["lr"] = lr
            if "_lr" in f:
                f["_lr"] = lr
            f["beta1"], f["beta2"], f["eps"] = betas[0], betas[1], eps
            prog.invalidate()
        for cb in self.on_hparams_changed:
            cb()

    def record_dp(self, prog: Program, pb: "PeerBuffers", site: int, gscale: float, lo: int, hi: int,
                  tag: str = "dp_adam"):
        """The phase-2 update of the flat range [lo, hi) fused with its collective: gradients summed over the ranks
        through peer memory, Adam on the slice this rank owns, parameters stored to every rank (b2h_dp_adam)."""
        assert self.store.flat.data_ptr() == pb.flat.data_ptr() and self.store.grad.data_ptr() == pb.grad.data_ptr()
        assert lo % 4 == 0 and (hi - lo) % 4 == 0 and hi > lo
        sig_off = 4 * site * PeerBuffers.SITE_WORDS
        i = prog.add(L.OP_DP_ADAM, tag, p=[a + 4 * lo for a in pb.p_ptrs], g=[a + 4 * lo for a in pb.g_ptrs],
                     signal=[a + sig_off for a in pb.sig_ptrs],
                     g_mc=pb.g_mc + 4 * lo if pb.g_mc else None, p_mc=pb.p_mc + 4 * lo if pb.p_mc else None,
                     m=self.m[lo:hi], v=self.v[lo:hi], n=hi - lo, rank=pb.rank, world=pb.world,
                     beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, gscale=float(gscale),
                     scalars=self.scalars, timeout_ms=int(os.environ.get("B2H_DP_TIMEOUT_MS", "0")),
                     _p=[t[lo:hi] for t in pb.p_tensors] if pb.p_tensors else None,
                     _g=[t[lo:hi] for t in pb.g_tensors] if pb.g_tensors else None,
                     _step=self.step, _lr=self.lr)
        self._recorded.append((prog, i))
        return i

    def state_dict(self):
        st = self.store
        state = {}
        step = float(self.step.item())
        for i, (k, shp) in enumerate(st.param_shapes):
            o, n = st.offsets[k], math.prod(shp)
            state[i] = {"step": torch.tensor(step), "exp_avg": self.m[o:o + n].view(shp).clone(),
                        "exp_avg_sq": self.v[o:o + n].view(shp).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(st.param_shapes)))}
        return {"state": state if step > 0 else {}, "param_groups": [group]}

    def load_state_dict(self, sd):
        st = self.store
        steps = []
        for i, (k, shp) in enumerate(st.param_shapes):
            s = sd["state"].get(i)
            if s is None:
                continue   # parameter never had a gradient in the reference (dead branch): state stays zero
            o, n = st.offsets[k], math.prod(shp)
            self.m[o:o + n].copy_(s["exp_avg"].reshape(-1))
            self.v[o:o + n].copy_(s["exp_avg_sq"].reshape(-1))
            steps.append(int(float(s["step"])))
        if steps:
            self.step.fill_(max(steps))
        g = sd["param_groups"][0]
        self.set_hparams(g["lr"], g["betas"], g["eps"])


class PeerBuffers:
    """Peer-mapped (symmetric) flat parameter / gradient buffers and signal pads of ONE network for the fused
    data-parallel optimizer step (b2h_dp_adam, include/b2h_abi.h): every rank can load the other ranks' gradients
    and store into their parameters over NVLink, so reduce-scatter + Adam + all-gather is one kernel."""
    SITE_WORDS = L.DP_MAX_BLOCKS * L.DP_MAX_PEERS      # uint32 per call site (gradient bucket)

    def __init__(self, rank, world, flat, grad, signals, p_ptrs, g_ptrs, sig_ptrs, p_mc=0, g_mc=0,
                 p_tensors=None, g_tensors=None, handles=()):
        assert 1 <= world <= L.DP_MAX_PEERS and len(p_ptrs) == len(g_ptrs) == len(sig_ptrs) == world
        self.rank, self.world = rank, world
        self.flat, self.grad, self.signals = flat, grad, signals
        self.p_ptrs, self.g_ptrs, self.sig_ptrs = list(p_ptrs), list(g_ptrs), list(sig_ptrs)
        self.p_mc, self.g_mc = int(p_mc or 0), int(g_mc or 0)
        self.p_tensors, self.g_tensors = p_tensors, g_tensors   # every rank's buffers as tensors (in-process only)
        self._handles = handles                                 # keeps the symmetric-memory handles alive

    @classmethod
    def symmetric(cls, n: int, n_sites: int, device, group, multicast: bool = True):
        """One process per GPU: allocate through torch.distributed._symmetric_memory (cuMem + fabric / fd handles,
        multicast object when the NVSwitch supports it) and exchange the peer mappings — a collective call."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        bufs, hdls = [], []
        for numel, dt in ((n, torch.float32), (n, torch.float32), (n_sites * cls.SITE_WORDS, torch.int32)):
            t = symm.empty(numel, dtype=dt, device=device)
            t.zero_()
            bufs.append(t)
            hdls.append(symm.rendezvous(t, group))
        torch.cuda.synchronize(device)
        dist.barrier(group=group)          # every pad is zero before anybody signals
        # a tensor may sit at an offset inside its symmetric block: peer addresses = block bases + that offset
        offs = [int(getattr(h, "offset", 0) or 0) for h in hdls]
        ptrs = [[int(a) + o for a in h.buffer_ptrs] for h, o in zip(hdls, offs)]
        rank = hdls[0].rank
        for t, pp in zip(bufs, ptrs):
            if pp[rank] != t.data_ptr():
                raise L.B2HError("symmetric memory: this rank's peer address does not match its local tensor "
                                 f"({pp[rank]:#x} vs {t.data_ptr():#x})")
        mc = [int(h.multicast_ptr or 0) + o if multicast and h.multicast_ptr else 0
              for h, o in zip(hdls[:2], offs[:2])]                                  # 0: no NVLS multicast object
        if not all(mc):
            mc = [0, 0]
        return cls(rank, hdls[0].world_size, bufs[0], bufs[1], bufs[2], ptrs[0], ptrs[1], ptrs[2],
                   p_mc=mc[0], g_mc=mc[1], handles=tuple(hdls))

    @classmethod
    def ipc(cls, n: int, n_sites: int, device, group):
        """One process per GPU, without torch's symmetric memory (`B2H_DP_PEER=ipc`): plain caching-allocator
        tensors exported through CUDA IPC handles (torch.multiprocessing.reductions), opened by every peer, peer
        access switched on by one tiny copy in each direction.  No multicast address in this mode.  Collective."""
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor
        device = torch.device(device)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        own = [torch.zeros(n, dtype=torch.float32, device=device), torch.zeros(n, dtype=torch.float32, device=device),
               torch.zeros(n_sites * cls.SITE_WORDS, dtype=torch.int32, device=device)]
        torch.cuda.synchronize(device)
        exported = [None] * world
        dist.all_gather_object(exported, [reduce_tensor(t) for t in own], group=group)
        opened, ptrs = [], [[], [], []]
        for q in range(world):
            for j in range(3):
                if q == rank:
                    t = own[j]
                else:
                    fn, args = exported[q][j]
                    t = fn(*args)                    # the peer's tensor, mapped into this process
                    probe = torch.zeros(1, dtype=t.dtype, device=device)
                    probe.copy_(t[:1])               # either direction once: torch enables peer access for the pair
                    t[:1].copy_(probe.zero_())       # (the buffers are still all zero)
                opened.append(t)
                ptrs[j].append(t.data_ptr())
        torch.cuda.synchronize(device)
        dist.barrier(group=group)
        return cls(rank, world, own[0], own[1], own[2], ptrs[0], ptrs[1], ptrs[2], handles=tuple(opened))

    @classmethod
    def in_process(cls, n: int, n_sites: int, devices):
        """All "ranks" inside one process (tests): CPU tensors for the emulated programs, or one CUDA device per
        rank with peer access enabled (plain cudaMalloc memory is peer-addressable once access is on)."""
        world = len(devices)
        flats = [torch.zeros(n, dtype=torch.float32, device=d) for d in devices]
        grads = [torch.zeros(n, dtype=torch.float32, device=d) for d in devices]
        sigs = [torch.zeros(n_sites * cls.SITE_WORDS, dtype=torch.int32, device=d) for d in devices]
        return [cls(r, world, flats[r], grads[r], sigs[r], [t.data_ptr() for t in flats],
                    [t.data_ptr() for t in grads], [t.data_ptr() for t in sigs], p_tensors=flats, g_tensors=grads)
                for r in range(world)]


FUSED_DP_MIN_WORLD = 4   # data parallel: the fused exchange (b2h_dp_adam) is the default from this many ranks on


class GanTrainer:
    """One process = one GPU.  Static device buffers `x`, `y`, `feats` hold the current batch."""

    def __init__(self, variant: str = "v1", in_dim: int = 36, out_dim: int = 252, require_feats: bool = False,
                 batch_size: int = 256, T: int = 64, precision: str = "bf16", device="cuda", lr: float = 1e-4,
                 seed: int = 23456, drop_mode: str = "philox", label_smooth: bool = False,
                 world_size: int = 1, process_group=None, n_buckets: Optional[int] = None, stores=None,
                 loss: str = "L1", fused_dp: Optional[bool] = None, peer_buffers=None, rank: Optional[int] = None):
        self.device = torch.device(device)
        if loss not in L.LOSS_KINDS:   # --loss of train_gan.py (LOSSES, utils/constants.py:53-58)
            raise KeyError(f"loss must be one of {sorted(L.LOSS_KINDS)}, got {loss!r}")
        self.loss = loss
        if T < 2 or T % 2:
            # MaxPool1d(2) + ConvTranspose1d(stride 2) return 2 * floor(T / 2) frames (modelZoo.py:197,262): for odd T
            # the reference's L1Loss(output, outputGT) (train_gan.py:292) fails on mismatched lengths as well
            raise ValueError(f"the training step needs an even number of frames >= 2 per clip, got T = {T} "
                             "(eval forwards of odd lengths go through the modelZoo modules)")
        self.B, self.T, self.precision = batch_size, T, precision
        self.dtype = dtype_of(precision)
        self.variant, self.require_feats = variant, require_feats
        # gradient / optimizer buckets per network: 3 on one GPU (the update of the late layers runs beside the
        # backward of the early ones); 1 under data parallelism, where every bucket is a collective and a
        # collective costs more than the overlap buys (8 GPUs: 1.01 ms per step with 1 bucket, 1.16 ms with 3)
        if n_buckets is None:
            n_buckets = 3 if world_size == 1 else 1
        self.world_size, self.pg, self.n_buckets = world_size, process_group, max(1, n_buckets)
        B = batch_size
        dev = self.device
        self.g_spec = nets.generator_spec(variant, in_dim, out_dim, require_feats, train=True)
        self.g_spec_eval = nets.generator_spec(variant, in_dim, out_dim, require_feats, train=False)
        self.d_spec = nets.discriminator_spec(out_dim)
        if stores is not None:      # share the flat buffers of existing modelZoo modules
            self.g_store, self.d_store = stores
        else:
            self.g_store = nets.ParamStore(self.g_spec, dev, seed=seed)
            self.d_store = nets.ParamStore(self.d_spec, dev, seed=seed + 1)
        # data parallel, optional (B2H_JOINT_ALLREDUCE=1): both networks' gradients in ONE buffer, so the pipelined
        # gan_step needs a single collective per step.  Measured: 2 GPUs 0.80 ms vs 0.82 ms per step with one
        # collective per network, but 8 GPUs 1.11 ms vs 1.01 ms (single samples) -> off by default
        self._joint_grad = None
        if world_size > 1 and stores is None and os.environ.get("B2H_JOINT_ALLREDUCE"):
            gn = (self.g_store.n + 63) // 64 * 64
            self._joint_grad = torch.zeros(gn + self.d_store.n, dtype=torch.float32, device=dev)
            self.g_store.grad = self._joint_grad[:self.g_store.n]
            self.d_store.grad = self._joint_grad[gn:gn + self.d_store.n]
        # data parallel, optional (fused_dp=True or B2H_FUSED_DP=1): parameters and gradients live in peer-mapped
        # memory and the optimizer step of a bucket is ONE kernel that also does the exchange over NVLink
        # (reduce-scatter of the gradients, Adam on the owned slice, all-gather of the parameters) instead of
        # ncclAllReduce + Adam.  Measured (profiles/dp_sweep_n8_r02.jsonl, 8 B200, v1 body 256 x 64 per GPU, one bucket):
        # 0.820 ms per step against 0.899 ms with NCCL (1 GPU: 0.743 ms); text 1.325 / 1.439, image 1.554 / 1.665 ->
        # the default where the trainer owns its stores, from FUSED_DP_MIN_WORLD ranks on (2 ranks: NCCL 0.837 ms,
        # fused 0.857 ms).  B2H_FUSED_DP=1 / 0 forces it on / off.
        if fused_dp is None:
            env = os.environ.get("B2H_FUSED_DP")
            able = world_size > 1 and stores is None and self._joint_grad is None and self.device.type == "cuda"
            fused_dp = able and (env not in (None, "", "0") if env is not None else world_size >= FUSED_DP_MIN_WORLD)
        self.fused_dp = bool(fused_dp)
        self._peer = {}
        self._watch = []          # (module, store) pairs whose torch-side versions are checked before a step
        self._dp_order = None     # event after the last fused exchange kernel enqueued in the current step
        if self.fused_dp:
            if world_size <= 1 or stores is not None or self._joint_grad is not None:
                raise ValueError("fused_dp needs world_size > 1, trainer-owned parameter stores and no joint all-reduce")
            for key, store in (("g", self.g_store), ("d", self.d_store)):
                if peer_buffers is not None:
                    pb = peer_buffers[key]
                elif os.environ.get("B2H_DP_PEER") == "ipc":
                    pb = PeerBuffers.ipc(store.n, self.n_buckets, dev, process_group)
                else:
                    pb = PeerBuffers.symmetric(store.n, self.n_buckets, dev, process_group,
                                               multicast=not os.environ.get("B2H_DP_NO_MULTICAST"))
                assert pb.world == world_size and pb.flat.numel() == store.n and (rank is None or pb.rank == rank)
                pb.flat.copy_(store.flat)
                store.flat, store.grad = pb.flat, pb.grad
                self._peer[key] = pb
        self.g_opt = FlatAdam(self.g_store, lr)
        self.d_opt = FlatAdam(self.d_store, lr)
        for opt in (self.g_opt, self.d_opt):
            opt.on_hparams_changed.append(self.release_graphs)
        # Philox (seed, step): the generator steps use the even steps 0, 2, 4, ..., the discriminator steps the
        # odd ones — the same numbers as one counter bumped after every step, but each network owns its
        # counter, so a discriminator step and the next generator step can run side by side (gan_step)
        self.drop_state = torch.zeros(2, dtype=torch.int64, device=dev)
        self.drop_state[0] = seed
        self.drop_state_d = torch.zeros(2, dtype=torch.int64, device=dev)
        self.drop_state_d[0] = seed
        self.drop_state_d[1] = 1
        kw = dict(drop_mode=drop_mode, drop_state=self.drop_state)
        kw_d = dict(drop_mode=drop_mode, drop_state=self.drop_state_d)
        # generator: train plan (G step) and eval plan (D step / inference)
        direct = os.environ.get("B2H_NO_WGRAD_DIRECT") is None
        # (the regression loss reads the output layer's BLC tile and writes the NCL `out` itself: no to_ncl pass)
        self.l1_reads_blc = os.environ.get("B2H_NO_L1_FUSE") is None
        # (opt-in, B2H_L1_DBIAS=1: its 16-byte form can also emit the bias gradient of the output layer -- column sums
        # by warp shuffle + one fp64 atomic per channel and CTA; measured: the loss kernel, which is on the chain,
        # grows from 12.0 to 14.9 us while the column-sum launch it replaces runs beside the chain: 0.656 vs 0.660 ms)
        self.l1_dbias = self.l1_reads_blc and T % 4 == 0 and bool(os.environ.get("B2H_L1_DBIAS"))
        self.G_train = nets.NetPlan(self.g_spec, self.g_store, B, T, self.dtype, dev, train=True, site_base=0,
                                    wgrad_direct=direct, out_by_loss=self.l1_reads_blc,
                                    out_dbias_external=self.l1_dbias, **kw)
        self.G_eval = nets.NetPlan(self.g_spec_eval, self.g_store, B, T, self.dtype, dev, train=False,
                                   weights_from=self.G_train)
        self.y = torch.zeros(B, out_dim, T, dtype=torch.float32, device=dev)
        # x / y / feats: the batch of the generator step.  xd / yd / featsd: the batch of the discriminator step
        # (and of infer()); load_batch() fills both with the same batch, advance_batch() shifts x -> xd.
        self.x = self.G_train.x
        self.feats = self.G_train.feats
        self.xd = self.G_eval.x
        self.featsd = self.G_eval.feats
        self.yd = torch.zeros(B, out_dim, T, dtype=torch.float32, device=dev)
        # discriminator: eval plan scoring calc_motion(G_train.out); grouped train plan on (fake, real)
        self.D_train = nets.NetPlan(self.d_spec, self.d_store, 2 * B, T, self.dtype, dev, train=True, groups=2,
                                    motion_src=[self.G_eval.out, self.yd], site_base=100, out_dbias_external=True,
                                    wgrad_direct=direct, **kw_d)
        self.D_eval = nets.NetPlan(self.d_spec, self.d_store, B, T, self.dtype, dev, train=False,
                                   motion_src=[self.G_train.out], weights_from=self.D_train)
        self.losses = torch.zeros(8, dtype=torch.float32, device=dev)  # [l1, adv, g_total, d_loss]
        self._build_loss_programs(label_smooth)
        self._build_bucket_programs()
        self._graphs: Dict[str, torch.cuda.CUDAGraph] = {}
        self._comm_stream = None
        self._copy_stream = None
        self._wgrad_streams = {}
        self._wgrad_rr = {}
        self._helper_events = {}       # first-pass helper launches of the running backward: tag -> event
        self._adv_stream = None
        self._d_stream = None
        self._g_stream = None
        self._opt_streams = {}
        self.bucketed_opt = os.environ.get("B2H_NO_BUCKETED_OPT") is None
        self.overlap_adv = os.environ.get("B2H_NO_ADV_OVERLAP") is None
        self.overlap_wgrad = os.environ.get("B2H_NO_WGRAD_OVERLAP") is None

    @classmethod
    def from_modules(cls, generator, discriminator, batch_size: int, T: int, precision: str = "fp32", **kw):
        """Fused trainer over the parameters of drop-in modelZoo modules: the modules' parameters and BN buffers
        ARE the trainer's flat buffers, so state_dict() / checkpoints / eval forwards see every update."""
        dev = next(generator.parameters()).device
        g_store = generator._materialize(dev)
        d_store = discriminator._materialize(dev)
        v, cin, cout, rf, D = generator._spec_args
        assert D == 256, "the fused trainer is built for default_size=256"
        tr = cls(v, cin, cout, rf, batch_size, T, precision=precision, device=dev, stores=(g_store, d_store), **kw)
        tr._watch = [(generator, g_store), (discriminator, d_store)]
        return tr

    def _sync_module_versions(self):
        """from_modules: parameters changed through the modules' torch API since the last look (load_state_dict,
        an in-place edit, a torch optimizer) must trigger a repack of the GEMM operand copies — the same bookkeeping
        as the modules' own forward (modelzoo._B2HModule.forward)."""
        for mod, store in self._watch:
            v = mod._param_versions()
            if v != mod._seen_versions:
                store.version += 1
                mod._seen_versions = v

    def load_batch(self, x, y, feats=None):
        """Copy one batch (device or pinned-host tensors in the reference layouts) into the static buffers: it is
        the batch of the next generator_step(), discriminator_step() and infer()."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if self.feats is not None:
            self.feats.copy_(feats.reshape(self.feats.shape), non_blocking=True)

    def _sync_d_batch(self):
        self.xd.copy_(self.x, non_blocking=True)
        self.yd.copy_(self.y, non_blocking=True)
        if self.feats is not None:
            self.featsd.copy_(self.feats, non_blocking=True)

    def advance_batch(self, x, y, feats=None):
        """Pipelined feeding for gan_step(): the current batch becomes the discriminator's batch, (x, y, feats)
        the generator's."""
        self._sync_d_batch()
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        if self.feats is not None:
            self.feats.copy_(feats.reshape(self.feats.shape), non_blocking=True)

    def prefetch_batch(self, x, y, feats=None):
        """Start the host->device copy of the NEXT batch (pinned host tensors) on a copy stream, into staging
        buffers: it overlaps the steps running on the current batch.  swap_batch() makes it current."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._stage = [torch.empty_like(self.x), torch.empty_like(self.y),
                           torch.empty_like(self.feats) if self.feats is not None else None]
            self._swap_done = None
        cs = self._copy_stream
        if self._swap_done is not None:
            cs.wait_event(self._swap_done)     # the previous swap has finished reading the staging buffers
        with torch.cuda.stream(cs):
            self._stage[0].copy_(x, non_blocking=True)
            self._stage[1].copy_(y, non_blocking=True)
            if self.feats is not None:
                self._stage[2].copy_(feats.reshape(self.feats.shape), non_blocking=True)
        self._prefetch_done = torch.cuda.Event()
        self._prefetch_done.record(cs)

    def swap_batch(self, pipelined: bool = False):
        """Make the prefetched batch the current one (device-to-device copies into the static step inputs).
        pipelined=True (gan_step feeding): the previous batch moves to the discriminator's inputs first."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._prefetch_done)
        if pipelined:
            self._sync_d_batch()          # the batch the generator just trained on goes to the discriminator
        self.x.copy_(self._stage[0], non_blocking=True)
        self.y.copy_(self._stage[1], non_blocking=True)
        if self.feats is not None:
            self.feats.copy_(self._stage[2], non_blocking=True)
        self._swap_done = torch.cuda.Event()
        self._swap_done.record(cur)

    def _build_loss_programs(self, label_smooth: bool):
        B, T, dev = self.B, self.T, self.device
        Gt, De, Dt = self.G_train, self.D_eval, self.D_train
        out_dim = self.g_spec.out_dim
        olb = Gt.bufs[Gt.out_layer.name]
        self.ticket = torch.zeros(8, dtype=torch.int32, device=dev)
        # ---- generator loss: L1(out, y) (+ gradient into the G backward) and MSE(D(motion(out)), 1)
        P = self.g_loss_prog = Program(self.dtype, dev)
        nblk = ((T + 31) // 32) * ((olb.Cp + 31) // 32) * B
        self.l1_partial = torch.zeros(nblk, dtype=torch.float32, device=dev)
        self.l1_dbias_accum = torch.zeros(16, out_dim, dtype=torch.float64, device=dev)
        Ld = De.bufs[De.out_layer.name].Lz
        with P.segment("loss"):
            P.add(L.OP_L1, "l1", out=Gt.out, gt=self.y, dout=olb.dpre, loss=self.losses[0:1], partial=self.l1_partial,
                  ticket=self.ticket[0:1], B=B, C=out_dim, L=T, ld=olb.Cp, Cfill=olb.Cp, gscale=1.0,
                  kind=L.LOSS_KINDS[self.loss],
                  dbias=self.g_store.g(Gt.out_layer.wkey + ".bias") if self.l1_dbias else None,
                  dbias_accum=self.l1_dbias_accum if self.l1_dbias else None,
                  out_blc=Gt.out_blc if self.l1_reads_blc else None,
                  out_blc_ld=Gt.out_blc.shape[-1] if self.l1_reads_blc else 0)
            # (the scalar form of the kernel -- T % 4 != 0 -- pays far more for its in-kernel column sums than the
            # separate 8 us colsum launch costs: 29.9 vs 13.4 + 7.7 us)
            P.add(L.OP_MSE, "adv", score=De.out_blc, dscore=None, loss=self.losses[1:2], add=self.losses[0:1],
                  total=self.losses[2:3], groups=1, n=B * Ld, ld=De.out_blc.shape[-1], target=[1.0, 0.0])
        with P.segment("opt"):
            self.g_opt.record(P, gscale=1.0 / self.world_size)
        # ---- discriminator loss: MSE(fake, t_fake) + MSE(real, t_real) and its gradient
        P = self.d_loss_prog = Program(self.dtype, dev)
        tf, tr = (0.1, 0.9) if label_smooth else (0.0, 1.0)   # train_gan.py:244-245
        dlb = Dt.bufs[Dt.out_layer.name]
        with P.segment("loss"):
            # the loss writes its gradient straight into the score layer's dpre rows (column 0; the padding stays
            # zero) and that layer's bias gradient (their sum): no separate row copy / column sum
            P.add(L.OP_MSE, "d_mse", score=Dt.out_blc, dscore=None, loss=self.losses[3:4], add=None, total=None,
                  groups=2, n=B * Ld, ld=Dt.out_blc.shape[-1], target=[tf, tr], dpre=dlb.dpre, dpre_ld=dlb.Cp,
                  dpre_bf16=1 if self.dtype == L.BF16 else 0, dbias=self.d_store.g(Dt.out_layer.wkey + ".bias"))
        with P.segment("opt"):
            self.d_opt.record(P, gscale=1.0 / self.world_size)

    def _build_bucket_programs(self):
        """Per network: the optimizer step split along the gradient buckets of the backward (Adam over the flat
        range of a bucket + the repack of its layers), so that the update of the late layers runs beside the
        backward of the early ones; only the last bucket's update stays on the dependency chain."""
        self._buckets = {}
        for key, plan, opt in (("g", self.G_train, self.g_opt), ("d", self.D_train, self.d_opt)):
            bp = self.bucket_plan(plan)
            P = Program(self.dtype, self.device)
            with P.segment("step"):
                opt.record(P, phase=1, tag="adam_step")
            for i, (_, _, lo, hi, _) in enumerate(bp):
                with P.segment(f"b{i}"):
                    if hi > lo and self.fused_dp:
                        opt.record_dp(P, self._peer[key], i, 1.0 / self.world_size, lo, hi, tag=f"dp_adam_b{i}")
                    elif hi > lo:
                        opt.record(P, gscale=1.0 / self.world_size, lo=lo, hi=hi, phase=2, tag=f"adam_b{i}")
            packs = plan.add_pack_buckets([names for *_, names in bp])
            self._buckets[key] = (bp, P, packs)

    def _bwd_update(self, key: str, pack_after=None, opt_after=None, joint=None):
        """Backward + optimizer step + weight repack of one network, bucket by bucket (see
        _build_bucket_programs).  pack_after / opt_after: events the repack / the parameter update must wait for
        (a concurrent reader of the packed weights / parameters in gan_step).  joint (a list): backward only —
        the events that complete this network's gradients are appended and the caller issues one collective
        for both networks, then _apply_update()."""
        plan, loss_prog = (self.G_train, self.g_loss_prog) if key == "g" else (self.D_train, self.d_loss_prog)
        cur = torch.cuda.current_stream(self.device)
        if joint is not None:
            for (s, e, _, _, _) in self._buckets[key][0]:
                for side in self._run_bwd_ops(plan, s, e, cur):
                    ev = torch.cuda.Event()
                    ev.record(side)
                    joint.append(ev)
            ev = torch.cuda.Event()
            ev.record(cur)
            joint.append(ev)
            return
        extra = [ev for ev in (opt_after, pack_after) if ev is not None]
        assert self.bucketed_opt or not self.fused_dp, "fused_dp runs through the bucketed optimizer programs"
        if not self.bucketed_opt:
            self._bwd_bucketed(plan)
            for ev in extra:
                cur.wait_event(ev)
            loss_prog.run("opt")
            plan.prog.run("pack")
            return
        bp, P, packs = self._buckets[key]
        if key not in self._opt_streams:
            self._opt_streams[key] = torch.cuda.Stream(self.device)
        opt_stream = self._opt_streams[key]
        P.run("step")                                   # advance the Adam step / bias corrections once
        used_opt = False
        carried = []     # side-stream gradients of earlier buckets that a LATER bucket's update consumes
        g0 = plan.store.grad.data_ptr()
        for i, (s, e, lo, hi, _) in enumerate(bp):
            side_ops = []
            used_sides = self._run_bwd_ops(plan, s, e, cur, side_ops)
            last = i == len(bp) - 1
            target = cur if last else opt_stream
            deps = []
            if not last or self.world_size > 1:
                ev = torch.cuda.Event()
                ev.record(cur)
                deps.append(ev)
            side_events = {}
            for side in used_sides:
                evw = torch.cuda.Event()
                evw.record(side)
                deps.append(evw)
                side_events[side] = evw
            deps += carried
            # a layer stored below this bucket's range (parameter order = the reference's registration order, not
            # the backward order: v4's text branch) is updated by a later bucket, which must see its gradient
            for side, rec in side_ops:
                out = rec.f["dW"] if rec.kind == L.OP_WGRAD else rec.f["out"]
                if (out.data_ptr() - g0) // 4 < lo and side_events[side] not in carried:
                    carried.append(side_events[side])
            if self.world_size > 1 and hi > lo and not self.fused_dp:   # (fused: the exchange is inside b{i})
                import torch.distributed as dist
                if self._comm_stream is None:
                    self._comm_stream = torch.cuda.Stream(self.device)
                for d in deps:
                    self._comm_stream.wait_event(d)
                with torch.cuda.stream(self._comm_stream):
                    dist.all_reduce(plan.store.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
                done = torch.cuda.Event()
                done.record(self._comm_stream)
                deps = [done]
            for d in deps + extra:
                target.wait_event(d)
            fused = self.fused_dp and hi > lo
            if fused and self._dp_order is not None:
                # the fused exchange kernels of a step wait on their peers: chain them in enqueue order (the same
                # on every rank), so that no rank can run two of them in the opposite order of another rank
                target.wait_event(self._dp_order)
            P.run(f"b{i}", target.cuda_stream)
            if fused:
                self._dp_order = torch.cuda.Event()
                self._dp_order.record(target)
            plan.prog.run(packs[i], target.cuda_stream)
            used_opt = used_opt or not last
        if used_opt:
            ev = torch.cuda.Event()
            ev.record(opt_stream)
            cur.wait_event(ev)

    def _apply_update(self, key: str):
        """Adam over the whole flat buffer + repack, on the current stream (after the caller's collective)."""
        plan = self.G_train if key == "g" else self.D_train
        bp, P, packs = self._buckets[key]
        P.run("step")
        for i in range(len(bp)):
            P.run(f"b{i}")
            plan.prog.run(packs[i])

    # ---- gradient all-reduce (data parallel) ---------------------------------------------------
    def bucket_plan(self, plan: nets.NetPlan):
        """Split the backward segment into `n_buckets` contiguous op ranges.  Parameters are laid out in forward
        order and the backward runs in reverse, so the gradients finished after bucket i form a suffix
        [lo_i, hi_i) of the flat buffer; the ranges tile [0, n) exactly once.
        Returns [(op_first, op_end, lo, hi, names of the layers stored in [lo, hi))]."""
        st = plan.store
        first, end = plan.prog.segments["bwd"]
        marks = plan.bwd_marks
        nb = min(self.n_buckets, len(marks))
        per = math.ceil(len(marks) / nb)
        # every layer that owns parameters, by the start of its block in the flat buffer (conv + BN parameters of a
        # layer are contiguous); layers without backward ops (dead branches) keep zero gradients: always complete
        start = {l.name: min(st.offsets[l.wkey + ".weight"], st.offsets[l.wkey + ".bias"]) for l in st.spec.all_layers()}
        by_offset = sorted(start, key=start.get, reverse=True)
        live = {n for n, _ in marks}
        out, done = [], set()
        hi = st.n
        for bi in range(nb):
            names = marks[bi * per:(bi + 1) * per]
            if not names:
                break
            s = names[0][1]
            last = (bi + 1) * per >= len(marks)
            e = end if last else marks[(bi + 1) * per][1]
            layer_names = {n for n, _ in names}
            done |= layer_names
            # the longest suffix [lo, n) of the flat buffer whose gradients are ALL complete once this bucket's ops
            # have run.  Usually that is "from this bucket's first layer on", but the parameter order is the
            # reference's module registration order, not the order of the backward pass: a text branch registered
            # early and used at the bottleneck (v4) finishes before the layers stored behind it
            lo = 0 if last else st.n
            if not last:
                for name in by_offset:
                    if name in done or name not in live:
                        lo = start[name]
                    else:
                        break
            lo = min(lo, hi)
            # the layers this bucket's optimizer step updates — whose GEMM operand copies it therefore repacks: those
            # stored in [lo, hi), not necessarily the ones whose backward ops it runs
            updated = {name for name, off in start.items() if lo <= off < hi}
            out.append((s, e, lo, hi, updated))
            hi = lo
        return out

    def _run_bwd_ops(self, plan: nets.NetPlan, s: int, e: int, cur, side_ops: Optional[list] = None) -> list:
        """Ops [s, e) of a backward segment.  The weight-gradient GEMMs (+ their split-K reduce) only feed the
        optimizer, so they go to side streams and overlap the bn_bwd -> dgrad chain of the following layers.
        Returns the side streams that were used; side_ops (a list) collects (stream, record) of every op sent there."""
        if not self.overlap_wgrad:
            plan.prog.run_range(s, e, cur.cuda_stream)
            return []
        key = "d" if plan is self.D_train else "g"
        sides, recs, used, i = self._side_streams_for(plan), plan.prog.recs, [], s
        helper_ev = self._helper_events.setdefault(key, {})

        def off_chain(rec):
            # weight gradients, bias-gradient column sums (also the finishing op of a deferred BatchNorm backward) and
            # the first-pass shares of skip-connection consumers (b2h_bn_bwd_t.first_pass_only): nothing on the
            # bn_bwd -> dgrad chain reads them before the optimizer / the producer's own bn_bwd
            return rec.kind in (L.OP_WGRAD, L.OP_COLSUM) or (rec.kind == L.OP_BN_BWD and rec.f.get("first_pass_only"))

        while i < e:
            is_w = off_chain(recs[i])
            j = i
            while j < e and off_chain(recs[j]) == is_w:
                j += 1
            if is_w:
                ev = torch.cuda.Event()
                ev.record(cur)             # dpre / the input gradient of this layer is complete
                for k in range(i, j):
                    # split-K wgrads share the plan's partial-plane workspace: they all go to stream 0, in order;
                    # the split-free ones (and the column sums / first-pass shares) own their outputs and rotate over
                    # the rest
                    shared_ws = recs[k].kind == L.OP_WGRAD and not (plan.wgrad_direct and recs[k].f["splits"] == 1)
                    if shared_ws or len(sides) == 1:
                        side = sides[0]
                    else:
                        side = sides[1 + self._wgrad_rr[key] % (len(sides) - 1)]
                        self._wgrad_rr[key] += 1
                    side.wait_event(ev)
                    plan.prog.run_range(k, k + 1, side.cuda_stream)
                    if recs[k].kind == L.OP_BN_BWD:     # a later bn_bwd on the chain needs these sums
                        hev = torch.cuda.Event()
                        hev.record(side)
                        helper_ev[recs[k].tag] = hev
                    elif side_ops is not None:
                        side_ops.append((side, recs[k]))
                    if side not in used:
                        used.append(side)
            else:
                k0 = i
                for k in range(i, j):
                    tags = recs[k].f.get("_wait_tags")
                    if tags:                # bn_bwd of a skip-connection producer: its helpers' sums must be in
                        if k > k0:
                            plan.prog.run_range(k0, k, cur.cuda_stream)
                        for t in tags:
                            cur.wait_event(helper_ev.pop(t))
                        k0 = k
                plan.prog.run_range(k0, j, cur.cuda_stream)
            i = j
        return used

    def _side_streams_for(self, plan: nets.NetPlan):
        """The wgrad streams of a train plan: the weight gradients of successive layers go round-robin to a few
        streams (a split-free wgrad keeps only a dozen SMs busy for tens of microseconds; several run side by
        side) — one set per network, the two backward passes of gan_step overlap."""
        key = "d" if plan is self.D_train else "g"
        if key not in self._wgrad_streams:
            k = max(1, int(os.environ.get("B2H_WGRAD_STREAMS", "4")))
            self._wgrad_streams[key] = [torch.cuda.Stream(self.device) for _ in range(k)]
            self._wgrad_rr[key] = 0
        return self._wgrad_streams[key]

    def _bwd_bucketed(self, plan: nets.NetPlan):
        """Backward in buckets; each bucket's flat-gradient range is all-reduced over NCCL on a side stream while
        the next bucket computes (the variant without the bucketed optimizer step)."""
        cur = torch.cuda.current_stream(self.device)
        st = plan.store
        pending = []
        for (s, e, lo, hi, _) in (self.bucket_plan(plan) if self.world_size > 1 else
                                  [plan.prog.segments["bwd"] + (0, 0, None)]):
            used = self._run_bwd_ops(plan, s, e, cur)
            evs = []
            for side in used:
                ev = torch.cuda.Event()
                ev.record(side)
                evs.append(ev)
            if self.world_size > 1 and hi > lo:
                import torch.distributed as dist
                if self._comm_stream is None:
                    self._comm_stream = torch.cuda.Stream(self.device)
                ev = torch.cuda.Event()
                ev.record(cur)
                for d in evs + [ev]:
                    self._comm_stream.wait_event(d)
                with torch.cuda.stream(self._comm_stream):
                    dist.all_reduce(st.grad[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
                done = torch.cuda.Event()
                done.record(self._comm_stream)
                pending.append(done)
            else:
                pending += evs
        for ev in pending:
            cur.wait_event(ev)

    # ---- steps ---------------------------------------------------------------------------------
    # The packed (GEMM-layout) weights of a network are refreshed right after its Adam update, once per
    # step, and shared by its train and eval plans; the eval plan's only per-step preparation is folding the
    # running BN statistics (one launch).
    def _g_ops(self, pack_after=None, adv_after=None, deferred_adv=None, joint=None):
        """pack_after / adv_after: events of a concurrent discriminator step (gan_step) that the weight repack
        (it overwrites what the D step's eval generator reads) and the D scoring branch (it wants the updated
        discriminator) have to wait for.  deferred_adv = (prep_done, adv_done): the scoring of THIS step is left
        to the next gan_step; the scoring of the previous step is in flight on the adv stream and must have read
        G_train.out (prep_done) before this forward overwrites it and losses[0] (adv_done) before L1 does."""
        if not self.overlap_adv:
            assert pack_after is None and adv_after is None and deferred_adv is None
            self.D_eval.prog.run("pack")      # fold D's running statistics
            self.G_train.prog.run("fwd")
            ls, le = self.g_loss_prog.segments["loss"]             # [l1, adv]
            self.g_loss_prog.run_range(ls, ls + 1)                 # (writes G_train.out when it reads the BLC tile)
            self.D_eval.prog.run("fwd")
            self.g_loss_prog.run_range(ls + 1, le)
            self._bwd_update("g")             # backward, Adam, repack of the updated generator weights
            return
        # The adversarial term of the generator loss carries no gradient (train_gan.py:285-287 detaches the
        # score): scoring the fake with D only produces a reported VALUE.  It runs as a parallel
        # branch (side stream) next to L1 -> backward -> Adam instead of in front of them.
        cur = torch.cuda.current_stream(self.device)
        ls, le = self.g_loss_prog.segments["loss"]             # [l1, adv]
        if deferred_adv is not None:
            prep_done, adv_done = deferred_adv
            fs, fe = self.G_train.prog.segments["fwd"]
            if self.l1_reads_blc:        # the loss kernel writes G_train.out and losses[0]
                self.G_train.prog.run_range(fs, fe, cur.cuda_stream)
                cur.wait_event(prep_done)
            else:
                assert self.G_train.prog.recs[fe - 1].kind == L.OP_TO_NCL
                self.G_train.prog.run_range(fs, fe - 1, cur.cuda_stream)
                cur.wait_event(prep_done)
                self.G_train.prog.run_range(fe - 1, fe, cur.cuda_stream)   # writes G_train.out
            cur.wait_event(adv_done)
            self.g_loss_prog.run_range(ls, ls + 1, cur.cuda_stream)
            self._bwd_update("g", pack_after=pack_after, joint=joint)
            return
        assert joint is None
        adv = self._get_adv_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        adv.wait_event(fork)
        if adv_after is not None:
            adv.wait_event(adv_after)
        self.D_eval.prog.run("pack", adv.cuda_stream)        # fold D's running statistics
        self.G_train.prog.run("fwd")
        if self.l1_reads_blc:
            # G_train.out (what the scoring branch reads) is written by the loss kernel
            self.g_loss_prog.run_range(ls, ls + 1, cur.cuda_stream)
            l1_done = torch.cuda.Event()
            l1_done.record(cur)
            adv.wait_event(l1_done)
            self.D_eval.prog.run("fwd", adv.cuda_stream)
        else:
            fwd_done = torch.cuda.Event()
            fwd_done.record(cur)
            adv.wait_event(fwd_done)
            self.D_eval.prog.run("fwd", adv.cuda_stream)
            self.g_loss_prog.run_range(ls, ls + 1, cur.cuda_stream)
            l1_done = torch.cuda.Event()
            l1_done.record(cur)
            adv.wait_event(l1_done)                          # total = l1 + adv
        self.g_loss_prog.run_range(ls + 1, le, adv.cuda_stream)
        self._bwd_update("g", pack_after=pack_after)   # backward, Adam, repack of the updated weights
        join = torch.cuda.Event()
        join.record(adv)
        cur.wait_event(join)

    def _get_adv_stream(self):
        if self._adv_stream is None:
            self._adv_stream = torch.cuda.Stream(self.device)
        return self._adv_stream

    def _score_previous_g_step(self, adv):
        """On stream `adv`: the adversarial value of the generator step whose output is in G_train.out, with the
        discriminator as it is now.  Returns (folded, prep_done, done) events."""
        self.D_eval.prog.run("pack", adv.cuda_stream)
        folded = torch.cuda.Event()
        folded.record(adv)
        s, e = self.D_eval.prog.segments["fwd"]
        assert self.D_eval.prog.recs[s].kind == L.OP_PREP
        self.D_eval.prog.run_range(s, s + 1, adv.cuda_stream)      # reads G_train.out
        prep_done = torch.cuda.Event()
        prep_done.record(adv)
        self.D_eval.prog.run_range(s + 1, e, adv.cuda_stream)
        ls, le = self.g_loss_prog.segments["loss"]
        self.g_loss_prog.run_range(ls + 1, le, adv.cuda_stream)    # adv, total = losses[0] + adv
        done = torch.cuda.Event()
        done.record(adv)
        return folded, prep_done, done

    def _d_ops(self):
        self.G_eval.prog.run("pack")      # fold G's running statistics
        self.G_eval.prog.run("fwd")
        self.D_train.prog.run("fwd")
        self.d_loss_prog.run("loss")
        self._bwd_update("d")

    def _gan_ops(self, lag_adv: bool):
        """[discriminator step on xd / yd] side by side with [generator step on x / y], see gan_step()."""
        assert self.overlap_adv, "gan_step needs the adversarial scoring branch on its own stream"
        self._dp_order = None
        cur = torch.cuda.current_stream(self.device)
        if self._d_stream is None:
            # (raising the priority of the two dependency chains over the wgrad / scoring branches was measured
            # slower: 0.87 vs 0.77 ms per step; B2H_STREAM_PRIORITY=1 to retry)
            hi = -1 if os.environ.get("B2H_STREAM_PRIORITY") else 0
            self._d_stream = torch.cuda.Stream(self.device, priority=hi)
            self._g_stream = torch.cuda.Stream(self.device, priority=hi)
        sd, sg = self._d_stream, self._g_stream
        outer = cur
        fork = torch.cuda.Event()
        fork.record(outer)
        sd.wait_event(fork)
        sg.wait_event(fork)
        adv_folded = adv_prep = adv_done = None
        if lag_adv:
            # third branch: the adversarial VALUE of the previous generator step (D is exactly what that step's
            # alternating-order scoring would see: the discriminator step before this one has completed)
            adv = self._get_adv_stream()
            adv.wait_event(fork)
            adv_folded, adv_prep, adv_done = self._score_previous_g_step(adv)
        with torch.cuda.stream(sd):
            self.G_eval.prog.run("pack")      # fold G's running statistics (before the G branch updates them)
            folded = torch.cuda.Event()
            folded.record(sd)
            self.G_eval.prog.run("fwd")
            geval_done = torch.cuda.Event()   # the packed generator weights have been read
            geval_done.record(sd)
            if lag_adv:
                sd.wait_event(adv_folded)     # D_train's forward updates the running statistics the fold reads
            self.D_train.prog.run("fwd")
            self.d_loss_prog.run("loss")
            # (Adam / repack change what the scoring branch reads: they wait for it)
            joint = [] if (self._joint_grad is not None and lag_adv) else None
            self._bwd_update("d", opt_after=adv_done if lag_adv else None, joint=joint)
            d_done = torch.cuda.Event()
            if joint is None:
                d_done.record(sd)
        with torch.cuda.stream(sg):
            sg.wait_event(folded)
            if lag_adv:
                self._g_ops(pack_after=geval_done, deferred_adv=(adv_prep, adv_done), joint=joint)
            else:
                self._g_ops(pack_after=geval_done, adv_after=d_done)
            g_done = torch.cuda.Event()
            if joint is None:
                g_done.record(sg)
        if joint is not None:
            # one collective over the joint gradient buffer of both networks, then the two updates side by side
            import torch.distributed as dist
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(self.device)
            comm = self._comm_stream
            for ev in joint:
                comm.wait_event(ev)
            with torch.cuda.stream(comm):
                dist.all_reduce(self._joint_grad, op=dist.ReduceOp.SUM, group=self.pg)
            reduced = torch.cuda.Event()
            reduced.record(comm)
            with torch.cuda.stream(sd):
                sd.wait_event(reduced)
                sd.wait_event(adv_done)
                self._apply_update("d")
                d_done.record(sd)
            with torch.cuda.stream(sg):
                sg.wait_event(reduced)
                sg.wait_event(geval_done)
                self._apply_update("g")
                g_done.record(sg)
        outer.wait_event(d_done)
        outer.wait_event(g_done)
        if lag_adv:
            outer.wait_event(adv_done)

    def flush_adv(self):
        """gan_step(lag_adv=True) leaves the adversarial value of its generator step to the next call; this
        computes it now (losses[1], losses[2]) — call after the last gan_step of a sequence."""
        cur = torch.cuda.current_stream(self.device)
        self._score_previous_g_step(cur)

    def gan_step(self, graph: bool = False, lag_adv: bool = True):
        """One discriminator step on the batch in xd / yd / featsd — with the generator as it is now — side by
        side with one generator step on the batch in x / y / feats.  The two are independent (the adversarial
        term of the generator loss has no gradient; only its reported value waits for the new discriminator), so
        generator_step(); [advance_batch(); gan_step()] * n  computes exactly the alternating schedule
        G0, D0, G1, D1, ... of n+1 generator and n discriminator steps, each D_k on the batch of G_k with the
        generator after G_k — with D_k overlapped with G_k+1.

        lag_adv (default): the reported VALUE of the adversarial term of a generator step (losses[1], and
        losses[2] = l1 + adv) needs the discriminator step that runs beside it, so it is computed at the start of
        the NEXT gan_step (or by flush_adv()) instead of at the end of this one: after the call losses[0] / [3]
        belong to this call's steps, losses[1] / [2] to the previous generator step.  lag_adv=False keeps all
        four current at the price of a serial tail."""
        self._ensure_packed()
        gkey = "gan_lag" if lag_adv else "gan"
        g = self._graphs.get(gkey) if graph else None
        if g is not None:
            g.replay()
        else:
            self._gan_ops(lag_adv)
            if graph:   # the eager pass above was the warm-up; capture for the next calls
                for key in ("d", "g"):
                    self._mark_stepped(key)
                    self._bump_step(key)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._gan_ops(lag_adv)
                    self._bump_step("d")
                    self._bump_step("g")
                self._graphs[gkey] = g
                return
            self._bump_step("d")
            self._bump_step("g")
        for key in ("d", "g"):
            self._mark_stepped(key)

    def release_graphs(self):
        """Drop the captured CUDA graphs (they are re-captured on the next graph=True step): after a change of what
        they froze (optimizer hyper-parameters), and before the process group goes away -- a graph that captured
        NCCL kernels keeps the communicator busy."""
        if self._graphs:
            torch.cuda.synchronize(self.device)
        self._graphs.clear()

    def _ensure_packed(self):
        """Weights changed from outside (load_state_dict, user edits): repack before the step."""
        self._sync_module_versions()
        self.G_train.ensure_packed()
        self.D_train.ensure_packed()

    def _mark_stepped(self, key):
        if key == "g":
            self.g_store.version += 1
            self.G_train._packed_version = self.g_store.version
        else:
            self.d_store.version += 1
            self.D_train._packed_version = self.d_store.version

    def _g_step_body(self):
        self._dp_order = None
        self._g_ops()

    def _d_step_body(self):
        self._dp_order = None
        self._d_ops()

    def _bump_step(self, key):
        (self.drop_state if key == "g" else self.drop_state_d)[1] += 2

    def generator_step(self, graph: bool = False):
        """One train_generator iteration on the batch currently in x / y / feats (train_gan.py:266-299)."""
        self._run("g", self._g_step_body, graph)

    def discriminator_step(self, graph: bool = False):
        """One train_discriminator iteration (train_gan.py:221-251) on the batch currently in x / y / feats."""
        self._sync_d_batch()
        self._run("d", self._d_step_body, graph)

    def _run(self, key, body, graph):
        self._ensure_packed()
        if not graph:
            body()
            self._mark_stepped(key)
            self._bump_step(key)
            return
        g = self._graphs.get(key)
        if g is None:
            body()   # warm up eagerly (finalises programs, sets kernel attributes)
            self._mark_stepped(key)
            self._bump_step(key)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
                self._bump_step(key)
            self._graphs[key] = g
            return
        g.replay()
        self._mark_stepped(key)

    # ---- inference -----------------------------------------------------------------------------
    def infer(self):
        """Eval forward of the generator on the batch in x / feats; result in G_eval.out (inference.py:115)."""
        self.xd.copy_(self.x, non_blocking=True)
        if self.feats is not None:
            self.featsd.copy_(self.feats, non_blocking=True)
        self.G_eval.forward()
        return self.G_eval.out

    def launches_per_gan_step(self) -> int:
        """Kernel launches of one generator step + one discriminator step: the library's launch counter around
        one EAGER pass of both steps (the graphs replay exactly this sequence).  Trains one step."""
        lib = L.load()
        n0 = int(lib.b2h_launch_count())
        self.generator_step(graph=False)
        self.discriminator_step(graph=False)
        torch.cuda.synchronize(self.device)
        return int(lib.b2h_launch_count()) - n0
