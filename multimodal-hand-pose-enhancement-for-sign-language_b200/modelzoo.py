"""Drop-in mirror of the reference's `modelZoo` classes (modelZoo.py:6-817): same class names, zero-argument
constructors, `build_net(...)` / `forward(input_, audio_=None, percent_rand_=0.7, feats_=None)` signatures and
`state_dict` keys -- backed by libb2h.so programs instead of torch.nn kernels.

The torch.nn containers created by `build_net` (Sequential of Dropout / Conv1d / LeakyReLU / BatchNorm1d ... in
the reference's order) are PARAMETER HOLDERS only: they give identical state_dict keys, identical default
initialisation (same RNG consumption under torch.manual_seed) and the usual `.to()/.train()/.eval()/
.parameters()` behaviour; their `forward` is never called.  On first use on a CUDA device the parameters and
BN buffers are re-pointed into one flat fp32 buffer (`ParamStore`) that the recorded programs read.

Precision: module attribute `precision` in {"fp32", "bf16"} (default from env B2H_PRECISION, else "fp32":
the reference's arithmetic).  There is no CPU path: calling forward on CPU tensors raises.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import _lib as L
from . import nets


def _default_precision() -> str:
    return os.environ.get("B2H_PRECISION", "fp32")


def _holder_block(l: nets.Layer):
    if l.kind == "conv":
        core = nn.Conv1d(l.cin, l.cout, l.k, stride=l.stride, padding=l.pad)
    elif l.kind == "convT":
        core = nn.ConvTranspose1d(l.cin, l.cout, l.k, stride=l.stride, padding=l.pad, output_padding=1)
    else:
        core = nn.Linear(l.cin, l.cout)
    mods = [nn.Dropout(0.5), core]
    if l.act == L.ACT_LEAKY:
        mods.append(nn.LeakyReLU(0.2, True))
    elif l.act == L.ACT_RELU:
        mods.append(nn.ReLU(True))
    if l.bn:
        mods.append(nn.BatchNorm1d(l.cout, momentum=l.momentum))
    return mods


class _B2HModule(nn.Module):
    """Shared machinery: holder containers, flat store aliasing, plan cache, autograd bridge."""

    def __init__(self):
        super().__init__()
        self.precision = _default_precision()
        self._spec_args: Optional[tuple] = None
        self._store: Optional[nets.ParamStore] = None
        self._plans: Dict[Tuple, list] = {}
        self._seen_versions = -1
        self._drop_state: Optional[torch.Tensor] = None
        self.seed = 23456
        self.drop_mode = "philox"   # "none": no dropout in train mode (deterministic comparisons in tests)

    # ---- construction -------------------------------------------------------------------------
    def _make_spec(self, train: bool) -> nets.NetSpec:
        raise NotImplementedError

    def _build_holders(self):
        spec = self._make_spec(True)
        by_seq: Dict[str, list] = {}
        for l in spec.all_layers():
            by_seq.setdefault(l.seq, []).append(l)
        for seq in spec.module_order:
            mods = []
            for l in sorted(by_seq[seq], key=lambda x: x.w_idx):
                mods += _holder_block(l)
            if seq == "encoder":
                mods.append(nn.MaxPool1d(kernel_size=2, stride=2))
            setattr(self, seq, nn.Sequential(*mods))
            if seq == "text_embeds_postprocess" and getattr(self, "_text_reduce", False):
                self.text_reduce = nn.Sequential(nn.MaxPool1d(kernel_size=2, stride=2))
            if seq == "image_resnet_postprocess":
                self.image_reduce = nn.Sequential(nn.MaxPool1d(kernel_size=2, stride=2))
        self._store = None
        self._plans.clear()

    # ---- flat store aliasing ----------------------------------------------------------------------
    def _named_tensors(self):
        out = dict(self.named_parameters())
        out.update(dict(self.named_buffers()))
        return out

    def _materialize(self, device: torch.device):
        st = self._store
        tensors = self._named_tensors()
        if st is not None and st.device == device:
            k0 = st.param_shapes[0][0]
            k1 = st.param_shapes[-1][0]
            if tensors[k0].data_ptr() == st.p(k0).data_ptr() and tensors[k1].data_ptr() == st.p(k1).data_ptr():
                return st
        if device.type != "cuda" and not getattr(self, "_allow_cpu_store", False):
            # (_allow_cpu_store: tests bind the parameters to a CPU store to check the aliasing and the recorded
            # programs with the op restatements; executing a program still needs the device and raises)
            raise L.B2HError("b2h_b200 modules run on CUDA (B200) devices only: move the module and its inputs "
                             "to a cuda device (there is no CPU fallback)")
        st = nets.ParamStore(self._make_spec(True), device, seed=0)
        with torch.no_grad():
            for k, _ in st.param_shapes:
                st.p(k).copy_(tensors[k].detach())
                tensors[k].data = st.p(k)
            for k, _ in st.buffer_shapes:
                st.b(k).copy_(tensors[k])
                tensors[k].data = st.b(k)
            for k in st.nbt_index:
                st.nbt_view(k).fill_(int(tensors[k]))
                tensors[k].data = st.nbt_view(k)[0]
        self._store = st
        self._plans.clear()
        self._drop_state = torch.zeros(2, dtype=torch.int64, device=device)
        self._drop_state[0] = self.seed
        return st

    def _param_versions(self) -> int:
        return sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers())

    MAX_LIVE_PLANS = 4   # grad-enabled forwards of one shape whose backward has not run yet

    def _plan(self, B: int, T: int, train: bool, lease: bool = False) -> nets.NetPlan:
        """The recorded plan of (B, T, mode, precision).  A plan owns the activations, batch statistics and dropout
        masks its backward reads, so a grad-enabled forward takes a LEASE on it (`lease=True`) that its backward (or
        the death of its autograd node) returns: a second forward of the same shape before that backward -- the
        reference's discriminator step scores fake and real before one `d_loss.backward()`, train_gan.py:240-249 --
        gets its own plan instead of overwriting the first one's saved state."""
        key = (B, T, train, self.precision, self.drop_mode)
        pool = self._plans.get(key)
        if pool is None:
            if len(self._plans) >= 8:
                self._plans.pop(next(iter(self._plans)))
            pool = self._plans[key] = []
        plan = next((p for p in pool if not p._leased), None) if lease else (pool[0] if pool else None)
        if plan is None:
            if len(pool) >= self.MAX_LIVE_PLANS:
                raise RuntimeError(f"{len(pool)} train-mode forwards of shape ({B}, {T}) are waiting for their backward: "
                                   "call backward() (or drop the outputs) before running more of them")
            dtype = L.BF16 if self.precision == "bf16" else L.F32
            plan = nets.NetPlan(self._make_spec(train), self._store, B, T, dtype, self._store.device, train=train,
                                drop_mode=self.drop_mode, drop_state=self._drop_state,
                                weights_from=None)
            plan._leased = False
            if train:
                olb = plan.bufs[plan.out_layer.name]
                plan.gout = torch.zeros_like(plan.out)
                with plan.prog.segment("gout"):
                    plan.prog.add(L.OP_PREP, "gout", src=plan.gout, out=olb.dpre, kind=L.SRC_NCL, B=B, L=olb.Lz,
                                  C=plan.out_layer.cout, ld=olb.Cp, Cfill=olb.Cp, src_ld=0, drop=None, out_f32=0)
            pool.append(plan)
        return plan

    def _run(self, x: torch.Tensor, feats: Optional[torch.Tensor]):
        if x.dim() != 3:
            raise ValueError(f"expected input of shape (B, C, T), got {tuple(x.shape)}")
        st = self._materialize(x.device)
        B, C, T = x.shape
        spec_in = st.spec.in_dim
        if C != spec_in:
            raise RuntimeError(f"expected {spec_in} input channels, got {C}")
        train = self.training
        needs_grad = train and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        plan = self._plan(B, T, train, lease=needs_grad)
        v = self._param_versions()
        if v != self._seen_versions:
            st.version += 1
            self._seen_versions = v
        plan.x.copy_(x.detach().to(torch.float32))
        if plan.feats is not None:
            if feats is None:
                raise RuntimeError("this model was built with require_text/require_image: feats_ is required")
            plan.feats.copy_(feats.detach().to(torch.float32).reshape(plan.feats.shape))
        if needs_grad:
            params = [p for _, p in self.named_parameters()]
            out = _NetFn.apply(self, plan, *params)
        else:
            plan.forward()
            out = plan.out.clone()
        if train:
            self._drop_state[1] += 1
            # BN buffers were updated in place by the kernels: keep version bookkeeping consistent
            self._seen_versions = self._param_versions()
        return out


class _Lease:
    """Held by the autograd node of one grad-enabled forward: the plan's saved state (activations, statistics, masks)
    belongs to that node until its backward has run or the node dies (output dropped / graph freed)."""

    def __init__(self, plan):
        self.plan = plan
        plan._leased = True

    def release(self):
        if self.plan is not None:
            self.plan._leased = False
            self.plan = None

    def __del__(self):
        self.release()


class _NetFn(torch.autograd.Function):
    """Bridges the recorded forward / backward programs into torch.autograd (g_loss.backward(), train_gan.py:294)."""

    @staticmethod
    def forward(ctx, module, plan, *params):
        ctx.lease = _Lease(plan)
        plan.forward()
        ctx.module, ctx.plan = module, plan
        return plan.out.clone()

    @staticmethod
    def backward(ctx, gout):
        plan, st = ctx.plan, ctx.module._store
        if ctx.lease.plan is None:
            raise RuntimeError("backward through a b2h_b200 module a second time: its saved activations were released "
                               "after the first backward (retain_graph is not supported)")
        plan.gout.copy_(gout.contiguous())
        plan.prog.run("gout")
        plan.backward()
        live = {l.wkey for l in plan.spec.layers} | {l.bnkey for l in plan.spec.layers if l.bn}
        grads = []
        for k, _ in st.param_shapes:
            base = k.rsplit(".", 1)[0]
            grads.append(st.g(k).clone() if base in live else None)
        ctx.lease.release()
        return (None, None, *grads)


class _Generator(_B2HModule):
    _variant = "v1"

    def build_net(self, feature_in_dim, feature_out_dim, require_text=None, default_size=256):
        self._build(feature_in_dim, feature_out_dim, require_text, default_size, None)

    def _build(self, feature_in_dim, feature_out_dim, require_text, default_size, require_image):
        rf = bool(require_image) if self._variant == "b2h" else bool(require_text)
        self.require_text = require_text
        self.require_image = bool(require_image)
        self.default_size = default_size
        self._spec_args = (self._variant, feature_in_dim, feature_out_dim, rf, default_size)
        self._text_reduce = rf and self._variant == "v1"
        self._build_holders()

    def _make_spec(self, train):
        v, cin, cout, rf, D = self._spec_args
        return nets.generator_spec(v, cin, cout, rf, D, train=train)

    def upsample(self, tensor, shape):  # modelZoo.py:295-296, kept for API parity
        return tensor.repeat_interleave(2, dim=2)[:, :, :shape[2]]

    def forward(self, input_, audio_=None, percent_rand_=0.7, feats_=None):
        return self._run(input_, feats_)


class regressor_fcn_bn_32(_Generator):            # modelZoo.py:169
    _variant = "v1"


class regressor_fcn_bn_32_b2h(_Generator):        # modelZoo.py:6
    _variant = "b2h"

    def build_net(self, feature_in_dim, feature_out_dim, require_image=False, default_size=256):
        self._build(feature_in_dim, feature_out_dim, None, default_size, require_image)
        self.use_resnet = True


class regressor_fcn_bn_32_v2(_Generator):         # modelZoo.py:331
    _variant = "v2"


class regressor_fcn_bn_32_v4(_Generator):         # modelZoo.py:443
    _variant = "v4"


class regressor_fcn_bn_32_v4_deeper(_Generator):  # modelZoo.py:557
    _variant = "v4_deeper"


class regressor_fcn_bn_discriminator(_B2HModule):  # modelZoo.py:763
    def build_net(self, feature_in_dim):
        self._spec_args = (feature_in_dim,)
        self._build_holders()

    def _make_spec(self, train):
        return nets.discriminator_spec(self._spec_args[0], motion_input=False)

    def forward(self, input_):
        return self._run(input_, None)
