"""Network specs (the five generator variants + the discriminator of modelZoo.py) as small DAGs of
Dropout->Conv/Linear->Act->BN blocks, the flat parameter store with the reference's state_dict
names, and `NetPlan`, which lowers a spec at a fixed (B, T) to recorded programs of libb2h ops:
  pack      master fp32 weights -> GEMM operand layouts (activation dtype)
  fwd       prep / bn_apply(+residual, up/pool, dropout) -> tap-GEMM(bias, act) -> bn_stats
  bwd       bn_bwd -> wgrad -> dgrad(+dropout mask) per block, in reverse order

Reference citations: modelZoo.py:6-166 (b2h), :169-328 (v1), :331-440 (v2), :443-554 (v4),
:557-710 (v4_deeper), :763-817 (discriminator).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import _lib as L
from .program import Program, no_drop

BN_EPS = 1e-5


def ceil64(c: int) -> int:
    return (c + 63) // 64 * 64


# ------------------------------------------------------------------------------------------------
# spec
# ------------------------------------------------------------------------------------------------
@dataclass
class Feed:
    src: Union["Layer", str]  # a Layer, or 'x' / 'feats' / 'motion'
    rowmap: int = L.ROW_IDENT
    dst_coff: int = 0


@dataclass
class Layer:
    name: str
    seq: str           # nn.Sequential attribute holding the block
    w_idx: int         # index of the Conv1d / ConvTranspose1d / Linear inside it
    kind: str          # 'conv' | 'convT' | 'linear'
    cin: int
    cout: int
    k: int = 1
    stride: int = 1
    pad: int = 0
    act: int = L.ACT_LEAKY
    bn: bool = True
    momentum: float = 0.1
    feeds: List[Feed] = field(default_factory=list)
    rows_like: Optional["Layer"] = None   # 'feats' linear layers: rows per sample follow this layer's output
    # filled by NetPlan
    La: int = 0
    Lo: int = 0

    @property
    def drop_site(self):
        return f"{self.seq}.{self.w_idx - 1}"

    @property
    def wkey(self):
        return f"{self.seq}.{self.w_idx}"

    @property
    def bnkey(self):
        return f"{self.seq}.{self.w_idx + 2}"


@dataclass
class NetSpec:
    name: str
    layers: List[Layer]            # topological order, live layers only
    module_order: List[str]        # registration order of the Sequentials (parameter order)
    in_dim: int
    out_dim: int
    feats: Optional[str] = None    # None | 'text' | 'image'
    dead: List[Layer] = field(default_factory=list)  # parameters that exist but never reach the output
    input_kind: str = "x"          # 'x' (generator) | 'motion' (discriminator)

    def all_layers(self):
        return self.layers + self.dead


def _blk(name, seq, w_idx, cin, cout, k, stride=1, pad=None, **kw):
    return Layer(name, seq, w_idx, "conv", cin, cout, k, stride, k // 2 if pad is None else pad, **kw)


def generator_spec(variant: str, in_dim: int, out_dim: int, require_feats: bool = False,
                   default_size: int = 256, train: bool = True) -> NetSpec:
    """DAG of one generator variant.  `train` only matters for text branches whose rows are identical
    over time in eval mode (no dropout): they are then computed once per clip and broadcast."""
    D = default_size
    rf = bool(require_feats)
    E = D + (D if rf else 0)
    layers: List[Layer] = []
    dead: List[Layer] = []
    feats_kind = None
    if variant in ("v1", "b2h"):
        enc_out = D
    else:
        enc_out = E
    enc = _blk("encoder", "encoder", 1, in_dim, enc_out, 3, feeds=[Feed("x")])
    layers.append(enc)
    side = None
    order: List[str] = []
    if rf and variant == "v1":
        feats_kind = "text"
        side = Layer("text", "text_embeds_postprocess", 1, "linear", 512, D, momentum=0.01, feeds=[Feed("feats")])
        layers.append(side)
        order += ["text_embeds_postprocess"]
    elif rf and variant == "b2h":
        feats_kind = "image"
        side = Layer("image", "image_resnet_postprocess", 1, "linear", 2000, D, momentum=0.01, feeds=[Feed("feats")])
        layers.append(side)
        order += ["image_resnet_postprocess"]
    elif rf:
        feats_kind = "text"
        order += ["text_embeds_postprocess"]
    order += ["encoder", "conv5", "conv6", "conv7"]
    f5 = [Feed(enc, L.ROW_POOL2, 0)]
    if side is not None:
        text_eval_bcast = (feats_kind == "text" and not train)
        f5.append(Feed(side, L.ROW_BCAST if text_eval_bcast else L.ROW_POOL2, D))
    conv5 = _blk("conv5", "conv5", 1, E, E, 3, feeds=f5)
    conv6 = _blk("conv6", "conv6", 1, E, E, 3, feeds=[Feed(conv5)])
    narrow = E // 2 if (rf and variant in ("v4", "v4_deeper")) else E
    c7out = narrow if variant == "v4" else E
    conv7 = _blk("conv7", "conv7", 1, E, c7out, 5, stride=2, pad=2, feeds=[Feed(conv6)])
    layers += [conv5, conv6, conv7]
    up_feeds = [Feed(conv7, L.ROW_UP2, 0)]
    if rf and variant == "v4":
        text = Layer("text", "text_embeds_postprocess", 1, "linear", 512, E // 2, momentum=0.01,
                     feeds=[Feed("feats")], rows_like=conv7)
        layers.append(text)
        up_feeds.append(Feed(text, L.ROW_UP2 if train else L.ROW_BCAST, c7out))
    if rf and variant == "v2":
        # modelZoo.py:429-433: the text row is appended as an extra TIME step and cropped away by
        # upsample(); it never reaches the output -> parameters exist, gradients are 0
        dead.append(Layer("text", "text_embeds_postprocess", 1, "linear", 512, E, momentum=0.01))
    if variant == "v4_deeper":
        sa, sb = "skip3", "skip4"
        order += ["conv8", "conv9", "conv10", "skip1", "skip2", "skip3", "skip4"]
        # modelZoo.py:689-704: conv8..conv10, skip1, skip2 and the text branch never reach the output (S7)
        dead += [_blk("conv8", "conv8", 1, E, E, 3), _blk("conv9", "conv9", 1, E, narrow, 3),
                 _blk("conv10", "conv10", 1, narrow, narrow, 3), _blk("skip1", "skip1", 1, E, E, 3),
                 _blk("skip2", "skip2", 1, E, E, 3)]
        if rf:
            dead.append(Layer("text", "text_embeds_postprocess", 1, "linear", 512, E // 2, momentum=0.01))
    else:
        sa, sb = "skip4", "skip5"
        order += ["skip4", "skip5"]
    order += ["decoder"]
    skipa = _blk(sa, sa, 1, E, E, 3, feeds=up_feeds + [Feed(conv6)])
    skipb = _blk(sb, sb, 1, E, E, 3, feeds=[Feed(skipa), Feed(conv5)])
    dec1 = _blk("decoder.1", "decoder", 1, E, E, 3, feeds=[Feed(skipb)])
    dec5 = Layer("decoder.5", "decoder", 5, "convT", E, out_dim, 7, 2, 3, act=L.ACT_RELU, feeds=[Feed(dec1)])
    dec9 = _blk("decoder.9", "decoder", 9, out_dim, out_dim, 7, act=L.ACT_NONE, bn=False, feeds=[Feed(dec5)])
    layers += [skipa, skipb, dec1, dec5, dec9]
    return NetSpec(f"{variant}{'+' + feats_kind if feats_kind else ''}", layers, order, in_dim, out_dim, feats_kind,
                   dead)


def discriminator_spec(in_dim: int, motion_input: bool = True) -> NetSpec:
    """`motion_input`: the plan computes calc_motion of an NCL tensor itself (trainer); otherwise the
    input already is the motion tensor (drop-in module, modelZoo.py:815-817)."""
    chans = [in_dim, 64, 64, 32, 32, 16, 16, 8]
    layers = []
    prev: Union[Layer, str] = "motion" if motion_input else "x"
    for i in range(7):
        lay = _blk(f"convs.{4 * i + 1}", "convs", 4 * i + 1, chans[i], chans[i + 1], 5, stride=2, pad=2,
                   feeds=[Feed(prev)])
        layers.append(lay)
        prev = lay
    layers.append(_blk("convs.29", "convs", 29, 8, 1, 3, act=L.ACT_NONE, bn=False, feeds=[Feed(prev)]))
    return NetSpec("discriminator", layers, ["convs"], in_dim, 1, None, [],
                   input_kind="motion" if motion_input else "x")


# ------------------------------------------------------------------------------------------------
# parameters: flat fp32 buffers exposed under the reference's state_dict names
# ------------------------------------------------------------------------------------------------
def _wshape(l: Layer):
    if l.kind == "conv":
        return (l.cout, l.cin, l.k)
    if l.kind == "convT":
        return (l.cin, l.cout, l.k)
    return (l.cout, l.cin)


class ParamStore:
    """Flat parameter / gradient / Adam-state buffers + BN buffers, with named views.

    Parameter order == the reference's `named_parameters()` order (module registration order, then
    weight/bias of each child in index order) so optimizer state indices line up."""

    def __init__(self, spec: NetSpec, device, seed: Optional[int] = None):
        self.spec = spec
        self.device = torch.device(device)
        by_seq: Dict[str, List[Layer]] = {}
        for l in spec.all_layers():
            by_seq.setdefault(l.seq, []).append(l)
        self.param_shapes: List[Tuple[str, Tuple[int, ...]]] = []
        self.buffer_shapes: List[Tuple[str, Tuple[int, ...]]] = []
        for seq in spec.module_order:
            for l in sorted(by_seq[seq], key=lambda x: x.w_idx):
                self.param_shapes += [(l.wkey + ".weight", _wshape(l)), (l.wkey + ".bias", (l.cout,))]
                if l.bn:
                    self.param_shapes += [(l.bnkey + ".weight", (l.cout,)), (l.bnkey + ".bias", (l.cout,))]
                    self.buffer_shapes += [(l.bnkey + ".running_mean", (l.cout,)), (l.bnkey + ".running_var", (l.cout,))]
        self.offsets: Dict[str, int] = {}
        off = 0
        for k, shp in self.param_shapes:
            self.offsets[k] = off
            off += (math.prod(shp) + 3) // 4 * 4
        self.n = off
        boff = 0
        self.boffsets: Dict[str, int] = {}
        for k, shp in self.buffer_shapes:
            self.boffsets[k] = boff
            boff += (math.prod(shp) + 3) // 4 * 4
        self.flat = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.grad = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.bufs = torch.zeros(max(boff, 4), dtype=torch.float32, device=self.device)
        bn_layers = [l for l in spec.all_layers() if l.bn]
        self.nbt_index = {l.bnkey + ".num_batches_tracked": i for i, l in enumerate(bn_layers)}
        self.nbt = torch.zeros(max(len(bn_layers), 1), dtype=torch.int64, device=self.device)
        self.version = 0  # bumped by whoever changes `flat` (re-pack trigger)
        self.reset_parameters(seed)

    # views -----------------------------------------------------------------------------------
    def _view(self, flat, k, shapes, offsets):
        shp = dict(shapes)[k]
        o = offsets[k]
        return flat[o:o + math.prod(shp)].view(shp)

    def p(self, k):
        return self._view(self.flat, k, self.param_shapes, self.offsets)

    def g(self, k):
        return self._view(self.grad, k, self.param_shapes, self.offsets)

    def b(self, k):
        return self._view(self.bufs, k, self.buffer_shapes, self.boffsets)

    def nbt_view(self, k):
        i = self.nbt_index[k]
        return self.nbt[i:i + 1]

    def reset_parameters(self, seed: Optional[int] = None):
        """PyTorch default init (kaiming_uniform(a=sqrt(5)) weights, U(+-1/sqrt(fan_in)) biases, BN 1/0)."""
        gen = torch.Generator(device="cpu")
        if seed is not None:
            gen.manual_seed(seed)
        else:
            gen.seed()
        with torch.no_grad():
            for k, shp in self.param_shapes:
                v = self.p(k)
                if k.endswith(".weight") and len(shp) >= 2:
                    fan_in = shp[1] * (shp[2] if len(shp) == 3 else 1)
                    bound = 1.0 / math.sqrt(fan_in)
                    v.copy_((torch.rand(shp, generator=gen) * 2 - 1) * bound)
                    bk = k[:-len("weight")] + "bias"
                    self.p(bk).copy_((torch.rand(dict(self.param_shapes)[bk], generator=gen) * 2 - 1) * bound)
                elif k.endswith(".weight"):
                    v.fill_(1.0)   # BN gamma
                elif len(shp) == 1 and (k[:-len("bias")] + "weight") in self.offsets and \
                        len(dict(self.param_shapes)[k[:-len("bias")] + "weight"]) == 1:
                    v.zero_()      # BN beta
            for k, _ in self.buffer_shapes:
                self.b(k).fill_(1.0 if k.endswith("running_var") else 0.0)
            self.nbt.zero_()
        self.version += 1

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Reference-ordered state_dict (clones)."""
        out = {}
        bn_keys = {k for k, _ in self.buffer_shapes}
        for k, _ in self.param_shapes:
            out[k] = self.p(k).detach().clone()
            if k.endswith(".bias"):
                base = k[:-len(".bias")]
                if base + ".running_mean" in bn_keys:
                    out[base + ".running_mean"] = self.b(base + ".running_mean").clone()
                    out[base + ".running_var"] = self.b(base + ".running_var").clone()
                    out[base + ".num_batches_tracked"] = self.nbt_view(base + ".num_batches_tracked")[0].clone()
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        missing = []
        with torch.no_grad():
            for k, _ in self.param_shapes:
                if k in sd:
                    self.p(k).copy_(sd[k])
                else:
                    missing.append(k)
            for k, _ in self.buffer_shapes:
                if k in sd:
                    self.b(k).copy_(sd[k])
                else:
                    missing.append(k)
            for k in self.nbt_index:
                if k in sd:
                    self.nbt_view(k).fill_(int(sd[k]))
        if strict and missing:
            raise KeyError(f"missing keys: {missing}")
        self.version += 1
        return missing


# ------------------------------------------------------------------------------------------------
# plan
# ------------------------------------------------------------------------------------------------
def _conv_out_len(La, k, stride, pad):
    return (La + 2 * pad - k) // stride + 1


@dataclass
class LayerBufs:
    a: torch.Tensor = None        # (B, La, Kc)    conv input after dropout
    z: torch.Tensor = None        # (B, Lo_act, Cp) post-activation, pre-BN
    mean: torch.Tensor = None     # (groups, Cp) batch statistics and the folded affine y = z*scale + shift
    invstd: torch.Tensor = None
    scale: torch.Tensor = None
    shift: torch.Tensor = None
    dpre: torch.Tensor = None     # (B, Lo_act, Cp)
    g: torch.Tensor = None        # (B, La, Kc) gradient w.r.t. the pre-dropout input (mask applied)
    wf: torch.Tensor = None       # packed forward weights
    wb: torch.Tensor = None       # packed dgrad weights
    bias: torch.Tensor = None     # packed bias
    bwd_accum: torch.Tensor = None  # fp64 (copies, groups, C, 2): first-pass sums of this layer's BN backward
    Kc: int = 0
    Cp: int = 0
    Lz: int = 0                   # rows per sample of z (2*La for convT)
    ain_C: int = 0                # valid channels of `a`


# frames per eval forward from which the output layer writes NCL itself and the skip additions / the pooling ride in
# the GEMM epilogues (see NetPlan); B2H_NCL_DIRECT_MIN_ROWS overrides it (tuning aid)
NCL_DIRECT_MIN_ROWS = int(os.environ.get("B2H_NCL_DIRECT_MIN_ROWS", "16384"))


class NetPlan:
    """One network at a fixed batch / length / precision / mode, lowered to a recorded program."""

    def __init__(self, spec: NetSpec, store: ParamStore, B: int, T: int, dtype: int, device,
                 train: bool, groups: int = 1, drop_mode: str = "philox",
                 drop_state: Optional[torch.Tensor] = None, site_base: int = 0,
                 motion_src: Optional[Sequence[torch.Tensor]] = None,
                 weights_from: Optional["NetPlan"] = None, out_dbias_external: bool = False,
                 wgrad_direct: bool = False, out_by_loss: bool = False):
        """`weights_from`: another plan of the SAME store and dtype whose packed forward weights / biases this
        plan reads instead of packing its own (the eval twin of a train plan: one repack per optimizer step
        serves both)."""
        assert drop_mode in ("none", "mask", "philox")
        assert weights_from is None or (weights_from.store is store and weights_from.dtype == dtype and not train)
        self.weights_from = weights_from
        # the op that writes the output layer's dpre (the loss) also produces that layer's bias gradient
        self.out_dbias_external = out_dbias_external
        # bf16: weight gradients without split-K (one CTA per output tile and tap walks all rows and writes dW
        # itself: no partial planes, no reduce launch, a fraction of the SM time) for every layer whose wgrad has
        # slack behind it, i.e. all but the first two layers of the network (the last two of the backward); meant
        # for callers that run the wgrads on several side streams (GanTrainer)
        self.wgrad_direct = wgrad_direct and dtype == L.BF16
        # train plans of a GanTrainer: the forward stops at the output layer's fp32 BLC tile; the loss kernel reads it
        # and writes the NCL `out` itself (b2h_l1_t.out_blc) -- no to_ncl launch, no re-read of its result
        self.out_by_loss = bool(out_by_loss) and train
        self.spec, self.store, self.B, self.T, self.dtype = spec, store, B, T, dtype
        self.device = torch.device(device)
        self.train, self.groups = train, groups
        self.drop_mode = drop_mode if train else "none"
        self.act_dtype = torch.bfloat16 if dtype == L.BF16 else torch.float32
        self.prog = Program(dtype, self.device)
        self.bufs: Dict[str, LayerBufs] = {}
        self.masks: Dict[str, torch.Tensor] = {}
        self.saved_masks: Dict[str, torch.Tensor] = {}
        self.site_ids: Dict[str, int] = {}
        self.drop_state = drop_state
        if self.drop_mode == "philox" and drop_state is None:
            self.drop_state = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._tickets: List[torch.Tensor] = []
        self._partial_need = 4
        self._wg_need = 16
        self._pending_partial: List[Tuple[int, str]] = []
        self._site_base = site_base
        self._motion_src_arg = list(motion_src) if motion_src is not None else None
        self.bwd_marks: List[Tuple[str, int]] = []   # (layer name, first op index) inside the bwd segment
        self.op_macs: Dict[int, int] = {}            # op index -> algorithmic MACs (padding not counted)
        self._build()

    # ---- helpers -----------------------------------------------------------------------------
    def _zeros(self, *shape, dtype=None):
        return torch.zeros(*shape, dtype=dtype or self.act_dtype, device=self.device)

    def _ticket(self):
        if not self._tickets or self._tickets[-1][1] >= 256:
            self._tickets.append([torch.zeros(256, dtype=torch.int32, device=self.device), 0])
        t, n = self._tickets[-1]
        self._tickets[-1][1] = n + 1
        return t[n:n + 1]

    def _drop(self, site: str, shape: Tuple[int, int, int], backward: bool = False, save: bool = False):
        """Dropout descriptor of a site whose tensor is (B, L, C) in BLC order.  In Philox mode the forward
        op of a site whose gradient is needed also stores the keep flags it drew (`save`), and the
        backward replays them in MASK mode (cheaper than regenerating Philox in the GEMM epilogue)."""
        if self.drop_mode == "none":
            return no_drop()
        sid = self.site_ids.setdefault(site, self._site_base + len(self.site_ids) + 1)
        if self.drop_mode == "mask":
            if site not in self.masks:
                self.masks[site] = torch.ones(shape, dtype=torch.uint8, device=self.device)
            assert tuple(self.masks[site].shape) == tuple(shape), (site, self.masks[site].shape, shape)
            return {"mode": L.DROP_MASK, "site": sid, "mask": self.masks[site], "state": None, "save": None}
        if backward:
            return {"mode": L.DROP_MASK, "site": sid, "mask": self.saved_masks[site], "state": None, "save": None}
        buf = None
        if save:
            buf = self.saved_masks.setdefault(site, torch.zeros(shape, dtype=torch.uint8, device=self.device))
        return {"mode": L.DROP_PHILOX, "site": sid, "mask": None, "state": self.drop_state, "save": buf}

    def set_masks(self, masks_ncl: Dict[str, torch.Tensor], group: Optional[int] = None):
        """Install keep-masks given in the reference's layout: (B, C, L) per conv site, (rows, C) per
        Linear site.  `group` selects the batch slice of a grouped (fake / real) discriminator batch."""
        assert self.drop_mode == "mask"
        for site, dst in self.masks.items():
            m = masks_ncl[site]
            if m.dim() == 3:
                m = m.permute(0, 2, 1)
            if group is not None:
                Bg = self.B // self.groups
                dst = dst[group * Bg:(group + 1) * Bg]
            dst.copy_(m.reshape(dst.shape).to(dst.device, torch.uint8))

    def _bn_src(self, l: Layer, rowmap=L.ROW_IDENT, coff=0):
        lb = self.bufs[l.name]
        scale, shift = lb.scale, lb.shift
        if l.name in getattr(self, "eval_y", {}):   # lb.z already holds BN(z): the consumers apply the identity
            scale, shift = self._ident_affine(lb.Cp)
        return {"z": lb.z, "ld": lb.Cp, "coff": coff, "rowmap": rowmap, "L_src": lb.Lz if rowmap != L.ROW_BCAST else 1,
                "Cs": lb.Cp, "scale": scale, "shift": shift,
                "mean": lb.mean if self.train else None, "invstd": lb.invstd if self.train else None}

    def _ident_affine(self, C: int):
        """(ones, zeros) per-channel arrays [1][C] for sources that are stored normalised (eval plans, eval_y)."""
        if getattr(self, "_ident", None) is None or self._ident[0].shape[-1] < C:
            n = max(C, 1024)
            self._ident = (torch.ones(1, n, dtype=torch.float32, device=self.device),
                           torch.zeros(1, n, dtype=torch.float32, device=self.device))
        return self._ident

    # ---- build -------------------------------------------------------------------------------
    def _build(self):
        spec, B, T = self.spec, self.B, self.T
        P = self.prog
        self.consumers: Dict[str, List[Tuple[Layer, Feed]]] = {l.name: [] for l in spec.layers}
        # shapes
        for l in spec.layers:
            f0 = l.feeds[0]
            if f0.src == "x":
                l.La = T
            elif f0.src == "motion":
                l.La = T - 1
            elif f0.src == "feats":
                if not (self.train or spec.feats == "image"):
                    l.La = 1          # eval-mode text rows are identical over time: one row per clip
                else:
                    l.La = self.bufs[l.rows_like.name].Lz if l.rows_like is not None else T
            else:
                srcl = f0.src
                Lz = self.bufs[srcl.name].Lz
                l.La = {L.ROW_IDENT: Lz, L.ROW_UP2: None, L.ROW_POOL2: Lz // 2, L.ROW_BCAST: None}[f0.rowmap]
                if l.La is None:  # up-sampled / broadcast feeds take the length of the identity feed
                    ident = [f for f in l.feeds if f.rowmap == L.ROW_IDENT]
                    l.La = self.bufs[ident[0].src.name].Lz
            for f in l.feeds:
                if isinstance(f.src, Layer):
                    self.consumers[f.src.name].append((l, f))
            if l.kind == "convT":
                l.Lo = l.La
                Lz = 2 * l.La
            else:
                l.Lo = _conv_out_len(l.La, l.k, l.stride, l.pad)
                Lz = l.Lo
            assert l.Lo >= 1, f"{l.name}: sequence too short (T={T})"
            lb = LayerBufs(Kc=ceil64(l.cin), Cp=ceil64(l.cout), Lz=Lz, ain_C=l.cin)
            lb.a = self._zeros(B, l.La, lb.Kc)
            if l.bn:
                lb.z = self._zeros(B, Lz, lb.Cp)
                # per-channel arrays padded to Cp (zeros beyond the layer's channels)
                lb.mean = self._zeros(self.groups, lb.Cp, dtype=torch.float32)
                lb.invstd = self._zeros(self.groups, lb.Cp, dtype=torch.float32)
                lb.scale = self._zeros(self.groups, lb.Cp, dtype=torch.float32)
                lb.shift = self._zeros(self.groups, lb.Cp, dtype=torch.float32)
            self.bufs[l.name] = lb
        self.out_layer = spec.layers[-1]
        ol = self.out_layer
        # static inputs / outputs (reference layouts)
        if spec.input_kind == "x":
            self.x = self._zeros(B, spec.in_dim, T, dtype=torch.float32)
        else:
            # discriminator: `groups` NCL tensors of (B/groups, C, T) whose calc_motion it scores
            if self._motion_src_arg is not None:
                self.motion_src = self._motion_src_arg
                assert len(self.motion_src) == self.groups
                for t in self.motion_src:
                    assert tuple(t.shape) == (B // self.groups, spec.in_dim, T) and t.dtype == torch.float32
            else:
                self.motion_src = [self._zeros(B // self.groups, spec.in_dim, T, dtype=torch.float32)
                                   for _ in range(self.groups)]
        if spec.feats == "text":
            self.feats = self._zeros(B, 512, dtype=torch.float32)
        elif spec.feats == "image":
            self.feats = self._zeros(B, T, 2000, dtype=torch.float32)
        else:
            self.feats = None
        olb = self.bufs[ol.name]
        self.out_blc = self._zeros(B, olb.Lz, max(olb.Cp, 4), dtype=torch.float32)
        self.out = self._zeros(B, ol.cout, olb.Lz, dtype=torch.float32)   # (B, C_out, T) NCL
        # batched inference (bf16, >= NCL_DIRECT_MIN_ROWS frames per forward): the output layer's GEMM writes the NCL
        # fp32 result itself (b2h_gemm_t.out_f32 = 2: persistent kernel, transposing staging buffer, TMA stores) -- no
        # fp32 BLC round trip through HBM and no to_ncl pass (at 4096 x 64: 264 MB written + read + written again)
        import os as _os0
        self.ncl_direct = (not self.train and self.dtype == L.BF16 and spec.input_kind == "x" and ol.kind != "convT" and
                           ol.stride == 1 and olb.Cp % 256 == 0 and olb.Lz % 4 == 0 and
                           B * olb.Lz >= NCL_DIRECT_MIN_ROWS and not _os0.environ.get("B2H_NO_NCL_DIRECT"))

        with P.segment("pack"):
            self._pack_items: List[dict] = []
            for l in spec.layers:
                self._emit_pack(l)
            if self._pack_items:
                self._emit_pack_table()
            if not self.train:
                self._emit_fold_table()
        # eval mode: a BN layer whose only consumer reads BN(z) as it is (one identity feed over all its columns) gets
        # its folded BatchNorm applied in the GEMM epilogue, which then writes the consumer's input directly
        self.eval_fused: Dict[str, Layer] = {}
        import os as _os
        if not self.train and not _os.environ.get("B2H_NO_EVAL_FUSE"):
            for pl in spec.layers:
                cons = self.consumers[pl.name]
                if pl.bn and len(cons) == 1:
                    c, f = cons[0]
                    if (len(c.feeds) == 1 and f.rowmap == L.ROW_IDENT and f.dst_coff == 0 and pl.cout == c.cin and
                            c.La == self.bufs[pl.name].Lz and self.groups == 1):
                        self.eval_fused[pl.name] = c
        # eval mode, BN layers with SEVERAL consumers (skip connections): the GEMM epilogue stores BN(z) in the layer's
        # own buffer (eval_y) -- every consumer then reads it with the identity affine, and a consumer fed by this
        # layer alone (identity rows, all columns) takes the buffer as its GEMM operand without a bn_apply pass
        # (eval_alias).  Only additions / up-sampling / pooling of normalised tensors keep their pass.
        self.eval_y: Dict[str, bool] = {}
        self.eval_alias: Dict[str, Layer] = {}
        if not self.train and self.groups == 1 and not _os.environ.get("B2H_NO_EVAL_FUSE"):
            for pl in spec.layers:
                cons = self.consumers[pl.name]
                if pl.bn and len(cons) >= 2 and pl.name not in self.eval_fused:
                    self.eval_y[pl.name] = True
                    pb = self.bufs[pl.name]
                    for c, f in cons:
                        if (len(c.feeds) == 1 and f.rowmap == L.ROW_IDENT and f.dst_coff == 0 and pl.cout == c.cin and
                                c.La == pb.Lz and self.bufs[c.name].Kc == pb.Cp):
                            self.eval_alias[c.name] = pl
        # batched inference (bf16, >= NCL_DIRECT_MIN_ROWS frames): the additions of the skip connections move into the
        # epilogue of the LATER producer (b2h_gemm_t.resid): a consumer fed by two BN layers over all its columns --
        # the later one (its only consumer; identity rows or x2 up-sampling) and an earlier one (identity rows, stored
        # normalised: eval_y) -- gets its input written by the later producer's GEMM: BN(z_later) [up-sampled] + y_earlier.
        self.eval_pool: Dict[str, Layer] = {}       # producer name -> its only consumer, fed through MaxPool1d(2)
        self.eval_resid: Dict[str, tuple] = {}      # later producer name -> (consumer, earlier producer, up2)
        self.eval_resid_consumers = set()
        big = (not self.train and self.groups == 1 and self.dtype == L.BF16 and spec.input_kind == "x" and
               B * self.bufs[ol.name].Lz >= NCL_DIRECT_MIN_ROWS and not _os.environ.get("B2H_NO_EVAL_RESID") and
               not _os.environ.get("B2H_NO_EVAL_FUSE"))
        if big:
            order = {l.name: i for i, l in enumerate(spec.layers)}
            for c in spec.layers:
                if len(c.feeds) != 2 or not all(isinstance(f.src, Layer) and f.src.bn and f.dst_coff == 0 and
                                                f.src.cout == c.cin for f in c.feeds):
                    continue
                fl, fe = sorted(c.feeds, key=lambda f: -order[f.src.name])      # later, earlier producer
                pl, pe = fl.src, fe.src
                pb, eb, cb = self.bufs[pl.name], self.bufs[pe.name], self.bufs[c.name]
                up2 = fl.rowmap == L.ROW_UP2
                ok = (pl is not pe and len(self.consumers[pl.name]) == 1 and pl.kind != "convT" and
                      pl.name not in self.eval_fused and pl.name not in self.eval_y and
                      (fl.rowmap == L.ROW_IDENT or up2) and fe.rowmap == L.ROW_IDENT and
                      pb.Lz * (2 if up2 else 1) == c.La and eb.Lz == c.La and
                      pb.Cp == eb.Cp == cb.Kc and pb.Cp % 256 == 0 and pe.name not in self.eval_fused)
                if not ok:
                    continue
                self.eval_resid[pl.name] = (c, pe, up2)
                self.eval_resid_consumers.add(c.name)
                self.eval_y[pe.name] = True      # (a single-consumer earlier producer stores BN(z) as well)
            # MaxPool1d(2) between a BN layer and its only consumer (encoder -> conv5): pooled in the producer's epilogue
            # (b2h_gemm_t.out_pool2), which writes the consumer's half-length input directly
            for pl in spec.layers:
                cons = self.consumers[pl.name]
                if not (pl.bn and len(cons) == 1 and pl.kind != "convT" and pl.name not in self.eval_fused and
                        pl.name not in self.eval_y and pl.name not in self.eval_resid):
                    continue
                c, f = cons[0]
                pb, cb = self.bufs[pl.name], self.bufs[c.name]
                if (len(c.feeds) == 1 and f.rowmap == L.ROW_POOL2 and f.dst_coff == 0 and pl.cout == c.cin and
                        pb.Lz % 2 == 0 and pb.Lz // 2 == c.La and pb.Cp == cb.Kc and pb.Cp % 256 == 0 and pb.Lz >= 16):
                    self.eval_pool[pl.name] = c
                    self.eval_resid_consumers.add(c.name)
        with P.segment("fwd"):
            for l in spec.layers:
                self._emit_input(l)
                self._emit_fwd_gemm(l)
            if spec.input_kind == "x" and not self.ncl_direct and not self.out_by_loss:
                P.add(L.OP_TO_NCL, "out", src=self.out_blc, dst=self.out, B=B, L=olb.Lz, C=ol.cout,
                      ld=self.out_blc.shape[-1], src_f32=1)
        if self.train:
            # gradient of the loss w.r.t. the output layer's pre-activation, BLC act dtype
            olb.dpre = self._zeros(B, olb.Lz, olb.Cp)
            self._plan_bwd_fusion()
            with P.segment("bwd"):
                for l in reversed(spec.layers):
                    self.bwd_marks.append((l.name, len(P.recs)))
                    self._emit_bwd(l)
        # shared workspaces, patched into the records
        self.partial = self._zeros(self._partial_need, dtype=torch.float32)
        self.wg_partial = self._zeros(self._wg_need // 4 + 4, dtype=torch.float32)
        for i, fld in self._pending_partial:
            f = self.prog.recs[i].f
            if isinstance(fld, tuple):   # nested descriptor, e.g. the statistics of a GEMM output
                f, fld = f[fld[0]], fld[1]
            f[fld] = self.wg_partial if self.prog.recs[i].kind == L.OP_WGRAD else self.partial
            if self.prog.recs[i].kind == L.OP_WGRAD:
                f["partial_bytes"] = self.wg_partial.numel() * 4   # checked by the library against its plan

    def _need_partial(self, idx: int, floats: int, fld: str = "partial"):
        self._partial_need = max(self._partial_need, int(floats))
        self._pending_partial.append((idx, fld))

    # ---- pack --------------------------------------------------------------------------------
    def _add_pack(self, tag: str, **fields):
        fields["_tag"] = tag
        self._pack_items.append(fields)

    def _emit_pack_table(self):
        """All weight repacks of the network as ONE launch: a device array of b2h_pack_t descriptors."""
        import ctypes as C
        from .program import _fill_struct
        n = len(self._pack_items)
        raw = bytearray()
        max_elems = 0
        for it in self._pack_items:
            st = _fill_struct(L.Pack(), {k: v for k, v in it.items() if not k.startswith("_")})
            raw += bytes(st)
            max_elems = max(max_elems, it["nphase"] * it["Opad"] * it["ntaps"] * it["Ipad"])
        self.pack_table = torch.frombuffer(raw, dtype=torch.uint8).clone().to(self.device)
        self.prog.add(L.OP_PACK_MULTI, "pack_multi", descs=self.pack_table, n=n, max_elems=max_elems,
                      _items=self._pack_items)

    def _emit_fold_table(self):
        """Eval-mode BN of every layer folded to scale/shift in ONE launch (device array of b2h_bn_fold_t)."""
        from .program import _fill_struct
        st, items = self.store, []
        for l in self.spec.layers:
            if l.bn:
                lb = self.bufs[l.name]
                items.append(dict(_tag=f"fold.{l.name}", gamma=st.p(l.bnkey + ".weight"), beta=st.p(l.bnkey + ".bias"),
                                  running_mean=st.b(l.bnkey + ".running_mean"),
                                  running_var=st.b(l.bnkey + ".running_var"), scale=lb.scale, shift=lb.shift,
                                  C=l.cout, Cpad=lb.Cp, eps=BN_EPS))
        if not items:
            return
        raw = bytearray()
        for it in items:
            raw += bytes(_fill_struct(L.BnFold(), {k: v for k, v in it.items() if not k.startswith("_")}))
        self.fold_table = torch.frombuffer(raw, dtype=torch.uint8).clone().to(self.device)
        self.prog.add(L.OP_BN_FOLD_MULTI, "fold_multi", descs=self.fold_table, n=len(items),
                      max_cpad=max(it["Cpad"] for it in items), _items=items)

    def add_pack_buckets(self, layer_sets) -> List[str]:
        """One extra repack op per set of layer names (segments 'pack_b0', 'pack_b1', ...): the buckets of a
        bucketed optimizer step repack their own layers as soon as their parameters are updated."""
        from .program import _fill_struct
        segs = []
        for bi, names in enumerate(layer_sets):
            items = [it for it in self._pack_items if it["_tag"].split(".", 1)[1].rsplit(".", 1)[0] in names]
            seg = f"pack_b{bi}"
            with self.prog.segment(seg):
                if items:
                    raw = bytearray()
                    for it in items:
                        raw += bytes(_fill_struct(L.Pack(), {k: v for k, v in it.items() if not k.startswith("_")}))
                    table = torch.frombuffer(raw, dtype=torch.uint8).clone().to(self.device)
                    self._bucket_tables = getattr(self, "_bucket_tables", []) + [table]
                    self.prog.add(L.OP_PACK_MULTI, seg, descs=table, n=len(items),
                                  max_elems=max(it["nphase"] * it["Opad"] * it["ntaps"] * it["Ipad"] for it in items),
                                  _items=items)
            segs.append(seg)
        return segs

    def _emit_pack(self, l: Layer):
        P, st, lb = self.prog, self.store, self.bufs[l.name]
        src = self.weights_from.bufs.get(l.name) if self.weights_from is not None else None
        if src is not None and src.wf is not None and (src.Kc, src.Cp) == (lb.Kc, lb.Cp):
            lb.wf, lb.bias, lb.fwd_taps = src.wf, src.bias, src.fwd_taps
            return
        W = st.p(l.wkey + ".weight")
        bias = st.p(l.wkey + ".bias")
        k = l.k
        lb.bias = self._zeros(lb.Cp, dtype=torch.float32)
        if l.kind in ("conv", "linear"):
            lb.fwd_taps = [t - l.pad for t in range(k)]
            lb.wf = self._zeros(lb.Cp, k, lb.Kc)
            self._add_pack(f"pack.{l.name}.fwd", W=W, out=lb.wf, O=l.cout, I=l.cin, Opad=lb.Cp, Ipad=lb.Kc,
                  ntaps=k, nphase=1, o_stride=l.cin * k, i_stride=k, k_stride=1,
                  tapmap=[list(range(k)) + [-1] * (L.MAX_TAPS - k), [-1] * L.MAX_TAPS], bias=bias, out_bias=lb.bias)
        else:  # convT k7 s2 p3 op1 as a 2-phase sub-pixel conv over taps {-1, 0, 1, 2}
            assert (l.k, l.stride, l.pad) == (7, 2, 3)
            lb.fwd_taps = [-1, 0, 1, 2]
            lb.wf = self._zeros(2 * lb.Cp, 4, lb.Kc)
            tm = [[3 - 2 * o for o in lb.fwd_taps], [4 - 2 * o for o in lb.fwd_taps]]
            tm = [[kk if 0 <= kk < k else -1 for kk in row] + [-1] * 4 for row in tm]
            self._add_pack(f"pack.{l.name}.fwd", W=W, out=lb.wf, O=l.cout, I=l.cin, Opad=lb.Cp, Ipad=lb.Kc,
                  ntaps=4, nphase=2, o_stride=k, i_stride=l.cout * k, k_stride=1, tapmap=tm, bias=bias,
                  out_bias=lb.bias)
        if not (self.train and self._needs_dgrad(l)):
            return
        if l.kind == "convT":
            # dA[b,i,c] = sum_k sum_n dpre[b, 2i + k - 3, n] W[c,n,k]: a stride-2 conv over dpre
            lb.bwd_taps = [kk - l.pad for kk in range(k)]
            lb.bwd_stride, lb.bwd_nphase = 2, 1
            lb.wb = self._zeros(lb.Kc, k, lb.Cp)
            self._add_pack(f"pack.{l.name}.bwd", W=W, out=lb.wb, O=l.cin, I=l.cout, Opad=lb.Kc, Ipad=lb.Cp,
                  ntaps=k, nphase=1, o_stride=l.cout * k, i_stride=k, k_stride=1,
                  tapmap=[list(range(k)) + [-1] * (L.MAX_TAPS - k), [-1] * L.MAX_TAPS], bias=None, out_bias=None)
        elif l.stride == 1:
            # dA[b,i,c] = sum_k sum_n dpre[b, i + p - k, n] W[n,c,k]
            offs = [t - (k - 1 - l.pad) for t in range(k)]
            lb.bwd_taps, lb.bwd_stride, lb.bwd_nphase = offs, 1, 1
            lb.wb = self._zeros(lb.Kc, k, lb.Cp)
            self._add_pack(f"pack.{l.name}.bwd", W=W, out=lb.wb, O=l.cin, I=l.cout, Opad=lb.Kc, Ipad=lb.Cp,
                  ntaps=k, nphase=1, o_stride=k, i_stride=l.cin * k, k_stride=1,
                  tapmap=[[l.pad - o for o in offs] + [-1] * (L.MAX_TAPS - k), [-1] * L.MAX_TAPS], bias=None,
                  out_bias=None)
        else:
            # stride 2: input row i = 2m + ph gets dpre[b, m + off, n] W[n,c,k] with k = ph + p - 2 off
            assert l.stride == 2
            cand = sorted({(ph + l.pad - kk) // 2 for ph in (0, 1) for kk in range(k) if (ph + l.pad - kk) % 2 == 0})
            lb.bwd_taps, lb.bwd_stride, lb.bwd_nphase = cand, 1, 2
            tm = [[(ph + l.pad - 2 * o) for o in cand] for ph in (0, 1)]
            tm = [[kk if 0 <= kk < k else -1 for kk in row] + [-1] * (L.MAX_TAPS - len(cand)) for row in tm]
            lb.wb = self._zeros(2 * lb.Kc, len(cand), lb.Cp)
            self._add_pack(f"pack.{l.name}.bwd", W=W, out=lb.wb, O=l.cin, I=l.cout, Opad=lb.Kc, Ipad=lb.Cp,
                  ntaps=len(cand), nphase=2, o_stride=k, i_stride=l.cin * k, k_stride=1, tapmap=tm, bias=None,
                  out_bias=None)

    def _needs_dgrad(self, l: Layer) -> bool:
        return any(isinstance(f.src, Layer) for f in l.feeds)

    # ---- forward -----------------------------------------------------------------------------
    def _emit_input(self, l: Layer):
        P, B, lb = self.prog, self.B, self.bufs[l.name]
        f0 = l.feeds[0]
        site_shape = (B, l.La, l.cin)
        if f0.src == "x":
            P.add(L.OP_PREP, f"prep.{l.name}", src=self.x, out=lb.a, kind=L.SRC_NCL, B=B, L=l.La, C=l.cin, ld=lb.Kc,
                  Cfill=lb.Kc, src_ld=0, drop=self._drop(l.drop_site, site_shape), out_f32=0)
            return
        if f0.src == "motion":
            Bg = B // self.groups
            drop = self._drop(l.drop_site, site_shape)
            for g in range(self.groups):
                d = dict(drop)
                a_g = lb.a[g * Bg:(g + 1) * Bg]
                if d["mode"] == L.DROP_MASK:
                    d["mask"] = self.masks[l.drop_site][g * Bg:(g + 1) * Bg]
                elif d["mode"] == L.DROP_PHILOX:
                    d["site"] = d["site"] + 1000 * g   # independent streams for the fake / real halves
                P.add(L.OP_PREP, f"motion.{l.name}.{g}", src=self.motion_src[g], out=a_g, kind=L.SRC_MOTION, B=Bg,
                      L=l.La, C=l.cin, ld=lb.Kc, Cfill=lb.Kc, src_ld=0, drop=d, out_f32=0)
            return
        if f0.src == "feats":
            if self.spec.feats == "image":
                P.add(L.OP_PREP, f"prep.{l.name}", src=self.feats, out=lb.a, kind=L.SRC_ROWS, B=B, L=l.La, C=l.cin,
                      ld=lb.Kc, Cfill=lb.Kc, src_ld=l.cin, drop=self._drop(l.drop_site, site_shape), out_f32=0)
            else:
                P.add(L.OP_PREP, f"prep.{l.name}", src=self.feats, out=lb.a, kind=L.SRC_BCAST, B=B, L=l.La, C=l.cin,
                      ld=lb.Kc, Cfill=lb.Kc, src_ld=l.cin, drop=self._drop(l.drop_site, site_shape), out_f32=0)
            return
        if isinstance(f0.src, Layer) and self.eval_fused.get(f0.src.name) is l:
            return   # the producer's GEMM epilogue wrote BN(z) into lb.a
        if l.name in self.eval_alias:
            return   # the producer's buffer holds BN(z) and IS this layer's GEMM operand
        if l.name in self.eval_resid_consumers:
            return   # the later producer's GEMM epilogue wrote BN(z_later) + y_earlier into lb.a
        # BN outputs of producer layers (+ residual / up-sampling / pooling), then this block's dropout
        cuts = sorted({f.dst_coff for f in l.feeds} | {f.dst_coff + f.src.cout for f in l.feeds})
        assert cuts[0] == 0 and cuts[-1] == l.cin, (l.name, cuts, l.cin)
        drop = self._drop(l.drop_site, site_shape, save=self.train and self._needs_dgrad(l))
        for s, e in zip(cuts[:-1], cuts[1:]):
            srcs = [f for f in l.feeds if f.dst_coff <= s and e <= f.dst_coff + f.src.cout]
            assert 1 <= len(srcs) <= 2, (l.name, s, e)
            srcs.sort(key=lambda f: f.rowmap == L.ROW_IDENT)   # non-identity source first (reference add order)
            src_d = [self._bn_src(f.src, f.rowmap, s - f.dst_coff) for f in srcs]
            while len(src_d) < 2:
                src_d.append({})
            last = e == l.cin
            P.add(L.OP_BN_APPLY, f"apply.{l.name}[{s}:{e}]", src=src_d, nsrc=len(srcs), out=lb.a, out_ld=lb.Kc,
                  out_coff=s, B=B, L=l.La, C=e - s, Cfill=(lb.Kc - s) if last else (e - s), groups=self.groups,
                  drop=drop, drop_C=l.cin, drop_coff=s)

    def _emit_fwd_gemm(self, l: Layer):
        P, B, lb = self.prog, self.B, self.bufs[l.name]
        is_out = l is self.out_layer
        out = self.out_blc if is_out else lb.z
        ldo = out.shape[-1]
        out_f32 = 1 if is_out else 0
        if is_out and self.ncl_direct:
            out, out_f32 = self.out, 2     # (B, C_out, T) fp32, written by the GEMM itself
        taps = lb.fwd_taps
        a_in = self.bufs[self.eval_alias[l.name].name].z if l.name in self.eval_alias else lb.a
        common = dict(A=a_in, W=lb.wf, bias=lb.bias, out=out, B=B, La=l.La, lda=lb.Kc, ldo=ldo, out_coff=0, Kc=lb.Kc,
                      Nvalid=l.cout, ntaps=len(taps), tap_off=taps + [0] * (L.MAX_TAPS - len(taps)), act=l.act,
                      post_scale=None, post_shift=None, out_f32=out_f32, drop=no_drop(), drop_C=0)
        if l.bn and self.train:
            # the GEMM also produces the batch statistics of its output (tensor-core path: in the epilogue)
            common["stats"] = self._bn_stats_desc(l)
        if l.name in self.eval_fused:
            cb = self.bufs[self.eval_fused[l.name].name]
            common.update(out=cb.a, ldo=cb.Kc, post_scale=lb.scale, post_shift=lb.shift)
        elif l.name in self.eval_resid:
            c, pe, up2 = self.eval_resid[l.name]
            cb, eb = self.bufs[c.name], self.bufs[pe.name]
            common.update(out=cb.a, ldo=cb.Kc, post_scale=lb.scale, post_shift=lb.shift, resid=eb.z, ld_resid=eb.Cp,
                          resid_up2=1 if up2 else 0)
        elif l.name in self.eval_pool:
            cb = self.bufs[self.eval_pool[l.name].name]
            common.update(out=cb.a, ldo=cb.Kc, post_scale=lb.scale, post_shift=lb.shift, out_pool2=1)
        elif l.name in self.eval_y:
            common.update(post_scale=lb.scale, post_shift=lb.shift)   # lb.z holds BN(z) from here on
        if l.kind == "convT":
            i = P.add(L.OP_GEMM, f"gemm.{l.name}", Lo=l.La, Npad=2 * lb.Cp, stride=1, nphase=2, Lo_actual=2 * l.La,
                      **common)
        else:
            i = P.add(L.OP_GEMM, f"gemm.{l.name}", Lo=l.Lo, Npad=lb.Cp, stride=l.stride, nphase=1, Lo_actual=l.Lo,
                      **common)
        if l.bn and self.train:
            self._need_partial(i, _bn_partial_floats(B * lb.Lz, l.cout, self.groups), fld=("stats", "partial"))
        self.op_macs[i] = self._layer_macs(l)

    def _layer_macs(self, l: Layer) -> int:
        """MACs of the layer's contraction as the reference counts them, whole batch."""
        rows = self.B * (l.La if l.kind == "convT" else l.Lo)
        return rows * l.cin * l.cout * l.k

    def _bn_stats_desc(self, l: Layer) -> dict:
        st, lb = self.store, self.bufs[l.name]
        rows = self.B * lb.Lz
        return dict(z=lb.z, ld=lb.Cp, C=l.cout, rows_per_group=rows // self.groups,
                    groups=self.groups, Cs=lb.Cp, mean=lb.mean, invstd=lb.invstd, scale=lb.scale, shift=lb.shift,
                    gamma=st.p(l.bnkey + ".weight"), beta=st.p(l.bnkey + ".bias"),
                    running_mean=st.b(l.bnkey + ".running_mean"),
                    running_var=st.b(l.bnkey + ".running_var"),
                    num_batches_tracked=st.nbt_view(l.bnkey + ".num_batches_tracked"), momentum=l.momentum,
                    eps=BN_EPS, partial=None, ticket=self._ticket(), update_all_groups=1 if self.groups > 1 else 0)

    # ---- backward ----------------------------------------------------------------------------
    def _plan_bwd_fusion(self):
        """Layers whose BN-backward first pass (sum dy, sum dy*zhat) is produced by the dgrad GEMMs of their
        consumers (b2h_gemm_t.bwd_sums).  A dgrad GEMM can serve one producer: its output must be exactly the
        gradient of that producer's BN output (one feed over all columns, IDENT or regular x2 up-sampling)."""
        self.dgrad_target: Dict[str, Tuple[Layer, Feed]] = {}
        self.bwd_fused = set()
        self.bwd_helpers = {}
        self.grad_add: Dict[str, Layer] = {}      # consumer (last in the backward) -> the other consumer of the tensor
        self.grad_summed: Dict[str, Layer] = {}   # producer -> the consumer whose input gradient holds the whole sum
        if os.environ.get("B2H_NO_FUSED_BWD"):
            return
        cands = [p for p in self.spec.layers if p.bn and p is not self.out_layer and self.consumers[p.name]]
        # Skip connections (bf16 plans): a BN layer whose output feeds two consumers row by row (conv5 -> conv6 and,
        # added to skip4's, -> skip5) receives the sum of two input gradients.  The consumer that comes last in the
        # backward pass adds the other one's gradient in its dgrad epilogue (b2h_gemm_t.grad_add): the producer's
        # bn_bwd then reads ONE source, and -- that dgrad being free to carry the sums -- runs one pass instead of two
        # over two sources (bn_bwd.conv5 / conv6: 18.4 -> 9.8 us each at 256 x 64).
        if self.dtype == L.BF16 and not os.environ.get("B2H_NO_GRAD_ADD"):
            order = {l.name: i for i, l in enumerate(self.spec.layers)}
            for p in cands:
                cons = self.consumers[p.name]
                pb = self.bufs[p.name]
                if len(cons) != 2 or cons[0][0] is cons[1][0]:
                    continue
                if not all(f.dst_coff == 0 and f.rowmap == L.ROW_IDENT and p.cout == c.cin and c.La == pb.Lz and
                           self.bufs[c.name].Kc == pb.Cp and self._needs_dgrad(c) for (c, f) in cons):
                    continue
                (c_last, _), (c_first, _) = sorted(cons, key=lambda cf: order[cf[0].name])
                # c_last's input gradient must belong to p alone: with a second feed (an addition in front of c_last)
                # it is also the gradient of that other tensor, which the added term would corrupt
                if c_last.name in self.grad_add or len(c_last.feeds) != 1:
                    continue
                self.grad_add[c_last.name] = c_first
                self.grad_summed[p.name] = c_last
        for p in sorted(cands, key=lambda q: len(self._grad_sources(q))):   # single-source layers first
            pb = self.bufs[p.name]
            ok = True
            for (c, f) in self._grad_sources(p):
                covers = f.dst_coff == 0 and p.cout == c.cin and self.bufs[c.name].Kc == pb.Cp
                simple = (f.rowmap == L.ROW_IDENT and c.La == pb.Lz) or (f.rowmap == L.ROW_UP2 and c.La == 2 * pb.Lz)
                # MaxPool1d(2) in between (encoder -> conv5): the dgrad's epilogue picks the pooled row of z itself
                # (b2h_bwd_sums_t.rowmap = POOL2; the only consumer, even length)
                pooled = (f.rowmap == L.ROW_POOL2 and pb.Lz == 2 * c.La and len(self._grad_sources(p)) == 1 and
                          not os.environ.get("B2H_NO_POOL_BWDSUM"))
                ok = ok and covers and (simple or pooled) and c.name not in self.dgrad_target and self._needs_dgrad(c)
            if not ok:
                continue
            for (c, f) in self._grad_sources(p):
                self.dgrad_target[c.name] = (p, f)
            self.bwd_fused.add(p.name)
            pb.bwd_accum = torch.zeros(L.BWD_COPIES, self.groups, p.cout, 2, dtype=torch.float64, device=self.device)
        # Skip connections: a layer with several consumers, some of whose dgrad GEMMs already serve another producer
        # (skip5's input is skip4 + conv5: its dgrad carries the sums of skip4).  The consumer that comes LAST in the
        # backward pass carries the sums in its dgrad; every other consumer's share is one first-pass launch over that
        # consumer's gradient alone (b2h_bn_bwd_t.first_pass_only), emitted right behind its dgrad and run beside the
        # chain (trainer: side stream) -- long before this layer's own bn_bwd is due, which then runs one pass only.
        # Opt-in (B2H_BWD_HELPERS=1): measured on the B200 it shortens the chain but not the step -- the step is bound
        # by the total work of its ~120 small launches, and a helper adds one (0.674 -> 0.690 ms, profiles/ab_r02_bn.log).
        self.bwd_helpers: Dict[str, List[Tuple[Layer, Feed]]] = {}     # consumer name -> [(producer, feed)]
        if not os.environ.get("B2H_BWD_HELPERS"):
            return
        order = {l.name: i for i, l in enumerate(self.spec.layers)}
        for p in cands:
            cons = self.consumers[p.name]
            if p.name in self.bwd_fused or len(cons) < 2 or p.name in self.grad_summed:
                continue
            pb = self.bufs[p.name]
            ok = all(f.dst_coff == 0 and p.cout == c.cin and self.bufs[c.name].Kc == pb.Cp and self._needs_dgrad(c) and
                     ((f.rowmap == L.ROW_IDENT and c.La == pb.Lz) or (f.rowmap == L.ROW_UP2 and c.La == 2 * pb.Lz))
                     for (c, f) in cons)
            last_c, last_f = min(cons, key=lambda cf: order[cf[0].name])   # first in the forward = last in the backward
            if not ok or last_c.name in self.dgrad_target:
                continue
            self.dgrad_target[last_c.name] = (p, last_f)
            for (c, f) in cons:
                if c is not last_c:
                    self.bwd_helpers.setdefault(c.name, []).append((p, f))
            self.bwd_fused.add(p.name)
            pb.bwd_accum = torch.zeros(L.BWD_COPIES, self.groups, p.cout, 2, dtype=torch.float64, device=self.device)

    def _grad_sources(self, p: Layer):
        """(consumer, feed) pairs whose input gradients the BN backward of `p` sums: all its consumers -- or, where one
        consumer's dgrad has added the other's gradient to its own (grad_add), that consumer alone."""
        cons = self.consumers[p.name]
        c_sum = getattr(self, "grad_summed", {}).get(p.name)
        return [(c, f) for (c, f) in cons if c is c_sum] if c_sum is not None else cons

    def _emit_bwd(self, l: Layer):
        P, st, B, lb = self.prog, self.store, self.B, self.bufs[l.name]
        rows = B * lb.Lz
        if l is self.out_layer and self.out_dbias_external:
            pass   # dpre AND the bias gradient come from the loss op
        elif l is self.out_layer:
            # dpre is provided by the loss (or by an external output gradient); bias grad = column sums
            # (own workspace: like the weight gradients this op only feeds the optimizer and may run beside the
            # BN kernels of the backward chain, which share `partial`)
            self.colsum_partial = self._zeros(128 * l.cout * 2, dtype=torch.float32)
            P.add(L.OP_COLSUM, f"dbias.{l.name}", src=lb.dpre, out=st.g(l.wkey + ".bias"),
                  partial=self.colsum_partial, ticket=self._ticket(), rows=rows, ld=lb.Cp, C=l.cout, f32=0)
        else:
            lb.dpre = self._zeros(B, lb.Lz, lb.Cp)
            gs = []
            for (c, f) in self._grad_sources(l):
                cb = self.bufs[c.name]
                gs.append({"g": cb.g, "ld": cb.Kc, "coff": f.dst_coff, "rowmap": f.rowmap, "L_src": c.La, "f32": 0})
            assert 1 <= len(gs) <= 2, (l.name, len(gs))
            ngs = len(gs)
            while len(gs) < 2:
                gs.append({})
            lb.sums = self._zeros(self.groups, l.cout, 2, dtype=torch.float32)
            fused = l.name in self.bwd_fused
            # First-pass sums complete in their own accumulator: the chain only needs dpre from this op, its tail
            # (threadfence, ticket, the last CTA summing the accumulator copies into dgamma / dbeta / dbias and re-zeroing
            # them: ~2.4 us of every launch) only feeds the optimizer.  B2H_DEFER_BN selects who runs it:
            #   2 (default)  bn_bwd still accumulates the sums of dpre -- in its OWN workspace, the finishing launch runs
            #                while the chain's next bn_bwd does -- and skips ticket + last-CTA pass; a one-CTA launch
            #                beside the chain (b2h_colsum, src = NULL) finishes: 0.663 -> 0.648 ms per GAN step
            #   1            bn_bwd writes dpre only (12.5 -> 6.5 us), a full column-sum launch over dpre finishes:
            #                measured slower (0.674 -> 0.679 ms; the re-read costs the step more than the chain gains)
            #   0            the op's own tail
            # (profiles/ab_r02_bn.log, profiles/ab_r02_defer2.log)
            defer = int(os.environ.get("B2H_DEFER_BN", "2") or 0) if fused else 0
            assert defer in (0, 1, 2), defer
            waits = [f"bwd_sums1.{l.name}.{c.name}" for (c, _) in self.consumers[l.name]
                     if any(p is l for (p, _) in self.bwd_helpers.get(c.name, []))]
            i = P.add(L.OP_BN_BWD, f"bn_bwd.{l.name}", gsrc=gs, ngsrc=ngs,
                      bn=self._bn_src(l), dpre=lb.dpre, ld_dpre=lb.Cp, Cfill=lb.Cp, B=B, L=lb.Lz, C=l.cout,
                      groups=self.groups, act=l.act, dgamma=None if defer else st.g(l.bnkey + ".weight"),
                      dbeta=None if defer else st.g(l.bnkey + ".bias"),
                      dbias=None if defer else st.g(l.wkey + ".bias"), sums=None if defer else lb.sums, partial=None,
                      ticket=self._ticket(), accum=lb.bwd_accum if fused else None, defer=defer,
                      _wait_tags=waits)
            if defer == 2:      # its own accumulators (the finishing launch runs while the chain's next bn_bwd does)
                lb.fin_partial = self._zeros(_bn_partial_floats(rows, l.cout, self.groups), dtype=torch.float32)
                P.recs[i].f["partial"] = lb.fin_partial
                P.add(L.OP_COLSUM, f"bn_fin.{l.name}", src=None, out=st.g(l.wkey + ".bias"), partial=lb.fin_partial,
                      ticket=None, rows=0, ld=lb.Cp, C=l.cout, f32=0, bn_accum=lb.bwd_accum,
                      dgamma=st.g(l.bnkey + ".weight"), dbeta=st.g(l.bnkey + ".bias"), bn_groups=self.groups)
            else:
                self._need_partial(i, _bn_partial_floats(rows, l.cout, self.groups))
            if defer == 1:
                lb.fin_partial = self._zeros(128 * l.cout * 2, dtype=torch.float32)
                P.add(L.OP_COLSUM, f"bn_fin.{l.name}", src=lb.dpre, out=st.g(l.wkey + ".bias"), partial=lb.fin_partial,
                      ticket=self._ticket(), rows=rows, ld=lb.Cp, C=l.cout, f32=0, bn_accum=lb.bwd_accum,
                      dgamma=st.g(l.bnkey + ".weight"), dbeta=st.g(l.bnkey + ".bias"), bn_groups=self.groups)
        # weight gradient
        k = l.k
        if l.kind == "convT":
            wg = dict(P=lb.a, Q=lb.dpre, Lp=l.La, Lq=lb.Lz, ldp=lb.Kc, ldq=lb.Cp, Mpad=lb.Kc, Npad=lb.Cp, Mvalid=l.cin,
                      Nvalid=l.cout, ntaps=k, stride=2, tap_off=[kk - l.pad for kk in range(k)] + [0] * (L.MAX_TAPS - k))
        else:
            wg = dict(P=lb.dpre, Q=lb.a, Lp=lb.Lz, Lq=l.La, ldp=lb.Cp, ldq=lb.Kc, Mpad=lb.Cp, Npad=lb.Kc, Mvalid=l.cout,
                      Nvalid=l.cin, ntaps=k, stride=l.stride,
                      tap_off=[t - l.pad for t in range(k)] + [0] * (L.MAX_TAPS - k))
        import os as _os
        direct = self.wgrad_direct and l not in self.spec.layers[:int(_os.environ.get("B2H_WGRAD_DIRECT_SKIP", "2"))]
        i = P.add(L.OP_WGRAD, f"wgrad.{l.name}", dW=st.g(l.wkey + ".weight"), partial=None, B=B,
                  splits=1 if direct else 0, **wg)
        self.op_macs[i] = self._layer_macs(l)
        self._wg_need = max(self._wg_need, _wgrad_ws_bytes(self.prog.recs[i], self.dtype))
        self._pending_partial.append((i, "partial"))
        # input gradient (with the dropout mask of this block's site)
        if not self._needs_dgrad(l):
            return
        lb.g = self._zeros(B, l.La, lb.Kc)
        taps = lb.bwd_taps
        drop = self._drop(l.drop_site, (B, l.La, l.cin), backward=True)
        common = dict(A=lb.dpre, W=lb.wb, bias=None, out=lb.g, B=B, La=lb.Lz, lda=lb.Cp, ldo=lb.Kc, out_coff=0,
                      Kc=lb.Cp, Nvalid=l.cin, ntaps=len(taps), tap_off=taps + [0] * (L.MAX_TAPS - len(taps)),
                      act=L.ACT_NONE, post_scale=None, post_shift=None, out_f32=0, drop=drop, drop_C=l.cin)
        if l.name in self.grad_add:        # skip connection: + the gradient the other consumer of the tensor wrote
            ob = self.bufs[self.grad_add[l.name].name]
            assert ob.g is not None and ob.g.shape == lb.g.shape, (l.name, self.grad_add[l.name].name)
            common["grad_add"] = ob.g
            common["ld_grad_add"] = ob.Kc
        if l.name in self.dgrad_target:
            p, f = self.dgrad_target[l.name]
            pb = self.bufs[p.name]
            common["bwd_sums"] = dict(z=pb.z, ld=pb.Cp, Lz=pb.Lz, rowmap=f.rowmap, C=p.cout, Cs=pb.Cp,
                                      groups=self.groups, mean=pb.mean, invstd=pb.invstd, accum=pb.bwd_accum,
                                      scale=pb.scale if f.rowmap == L.ROW_POOL2 else None,
                                      shift=pb.shift if f.rowmap == L.ROW_POOL2 else None)
        if lb.bwd_nphase == 2:
            i = P.add(L.OP_GEMM, f"dgrad.{l.name}", Lo=_ceil_div(l.La, 2), Npad=2 * lb.Kc, stride=1, nphase=2,
                      Lo_actual=l.La, **common)
        else:
            i = P.add(L.OP_GEMM, f"dgrad.{l.name}", Lo=l.La, Npad=lb.Kc, stride=lb.bwd_stride, nphase=1,
                      Lo_actual=l.La, **common)
        self.op_macs[i] = self._layer_macs(l)
        # this consumer's share of the first-pass sums of producers whose sums another dgrad carries (see
        # _plan_bwd_fusion): over lb.g alone, into the producer's accumulator
        for (p, f) in self.bwd_helpers.get(l.name, []):
            pb = self.bufs[p.name]
            gs = [{"g": lb.g, "ld": lb.Kc, "coff": f.dst_coff, "rowmap": f.rowmap, "L_src": l.La, "f32": 0}, {}]
            P.add(L.OP_BN_BWD, f"bwd_sums1.{p.name}.{l.name}", gsrc=gs, ngsrc=1, bn=self._bn_src(p), dpre=None,
                  ld_dpre=pb.Cp, Cfill=pb.Cp, B=B, L=pb.Lz, C=p.cout, groups=self.groups, act=p.act, dgamma=None,
                  dbeta=None, dbias=None, sums=None, partial=None, ticket=None, accum=pb.bwd_accum, defer=0,
                  first_pass_only=1)

    # ---- execution ---------------------------------------------------------------------------
    def pack(self):
        self.prog.run("pack")
        self._packed_version = self.store.version

    def ensure_packed(self):
        if self.weights_from is not None:
            self.weights_from.ensure_packed()
        if getattr(self, "_packed_version", None) != self.store.version:
            self.pack()

    def forward(self):
        self.ensure_packed()
        self.prog.run("fwd")

    def backward(self):
        self.prog.run("bwd")


def _ceil_div(a, b):
    return (a + b - 1) // b


def _bn_partial_floats(rows, C, groups):
    return 128 * groups * C * 2   # kMaxChunks partials of (a, b) per group and channel


def _wgrad_ws_bytes(rec, dtype) -> int:
    """Size of the wgrad split-K workspace: the library's own b2h_wgrad_workspace_bytes (an upper bound that does not
    need a GPU); the formula below only serves CPU-side plan emulation when the library cannot be loaded."""
    f = rec.f
    try:
        import ctypes as C
        from .program import _fill_struct
        lib = L.load()
        desc = _fill_struct(L.Wgrad(), {k: v for k, v in f.items()
                                        if not k.startswith("_") and k not in ("P", "Q", "dW", "partial")})
        need = int(lib.b2h_wgrad_workspace_bytes(C.byref(desc), dtype))
        if need > 0:
            return need
    except (OSError, L.B2HError):
        pass
    planes = f["ntaps"] * f["Mpad"] * f["Npad"] * 4
    # the split count is bounded by the number of 64-row k-blocks of the (tb clips x tl rows) tiling, which
    # exceeds rows // 64 when the tiles are ragged (odd lengths: tl = 1, tb = 64 -> one k-block per time step)
    Lp, B = f["Lp"], f["B"]
    tl, best_pad = 1, Lp
    t = 1
    while t <= 64:
        pad = -(-Lp // t) * t
        if pad <= best_pad:
            tl, best_pad = t, pad
        t *= 2
    tb = 64 // tl
    total_kb = -(-B // tb) * -(-Lp // tl)
    max_splits = min(64, max(1, total_kb, B * Lp // 64))
    return (max_splits + 1) * planes
