"""ORACLE (test infrastructure only) -- CPU restatement of the reference hot path.

This file restates, in plain torch.nn on CPU, the algorithm of the reference's
`modelZoo.py` generators / discriminator and of the `train_gan.py` step bodies.
It is the *checker* for the CUDA path: only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.  The
product package never does (it fails loudly if its CUDA library is missing).

Parity pinning: `tests/test_oracle_vs_reference.py` imports the real reference
from /root/reference (when present, i.e. in the authoring container) and checks
state_dict keys/shapes and bit-identical outputs/gradients against this
restatement; `tests/golden/*.npz` hold outputs of the real reference produced
by `tools/make_golden.py` so the pin travels to the GPU box.

Reference citations (file:line relative to /root/reference):
  * block order Dropout -> Conv1d -> LeakyReLU(0.2) -> BatchNorm1d ........ modelZoo.py:192-198
  * v1  `regressor_fcn_bn_32`            build_net :173-281, forward :299-328
  * b2h `regressor_fcn_bn_32_b2h`        build_net :10-118,  forward :137-166
  * v2  `regressor_fcn_bn_32_v2`         build_net :335-405, forward :422-440
  * v4  `regressor_fcn_bn_32_v4`         build_net :447-519, forward :534-554
  * v4_deeper                            build_net :561-667, forward :683-710
  * discriminator                        build_net :767-813, forward :815-817
  * calc_motion                          train_gan.py:209-211
  * generator / discriminator step       train_gan.py:266-299 / :221-251

Unlike the reference, every Dropout site here can *replay* an explicit keep-mask
(`masks[site_name]`, uint8/bool, same shape as the site's input), which is how
train-mode parity is defined: RNG streams can never match
between a CPU oracle and a CUDA Philox generator.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
from torch import nn


class ReplayDropout(nn.Module):
    """nn.Dropout(p) that replays an explicit keep-mask when one is installed.

    Semantics of nn.Dropout in train mode: y = x * keep / (1 - p)
    (modelZoo.py:193 etc. use p = 0.5 -> scale 2).  In eval mode: identity.
    """

    def __init__(self, p: float = 0.5):
        super().__init__()
        self.p = p
        self.site = ""          # filled by the owner: e.g. "encoder.0"
        self.store: Optional[Dict[str, torch.Tensor]] = None

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        if self.store is not None:
            if self.site not in self.store:
                raise KeyError(f"no dropout mask for site {self.site}")
            keep = self.store[self.site].to(x.dtype)
            assert keep.shape == x.shape, (self.site, keep.shape, x.shape)
            return x * keep * (1.0 / (1.0 - self.p))
        return nn.functional.dropout(x, self.p, True)


def _block(cin, cout, k, stride=1, pad=None, momentum=0.1):
    pad = (k // 2) if pad is None else pad
    return [ReplayDropout(0.5), nn.Conv1d(cin, cout, k, stride=stride, padding=pad),
            nn.LeakyReLU(0.2, True), nn.BatchNorm1d(cout, momentum=momentum)]


def _decoder(E, out_dim):
    # modelZoo.py:265-281 (identical in every generator variant)
    return nn.Sequential(
        *_block(E, E, 3),
        ReplayDropout(0.5),
        nn.ConvTranspose1d(E, out_dim, 7, stride=2, padding=3, output_padding=1),
        nn.ReLU(True),
        nn.BatchNorm1d(out_dim),
        ReplayDropout(0.5),
        nn.Conv1d(out_dim, out_dim, 7, padding=3),
    )


def _proj(fin, fout):
    # text / image post-process branch: Dropout, Linear, LeakyReLU, BN(momentum=0.01)
    # modelZoo.py:182-187 (text), :19-24 (image)
    return nn.Sequential(ReplayDropout(0.5), nn.Linear(fin, fout), nn.LeakyReLU(0.2, True),
                         nn.BatchNorm1d(fout, momentum=0.01))


VARIANTS = ("v1", "b2h", "v2", "v4", "v4_deeper")
REF_CLASS = {  # utils/constants.py:45-51
    "v1": "regressor_fcn_bn_32", "b2h": "regressor_fcn_bn_32_b2h", "v2": "regressor_fcn_bn_32_v2",
    "v4": "regressor_fcn_bn_32_v4", "v4_deeper": "regressor_fcn_bn_32_v4_deeper",
}


class RefGenerator(nn.Module):
    """Restatement of the five generator variants with identical state_dict keys."""

    def __init__(self, variant: str, feature_in_dim: int, feature_out_dim: int,
                 require_feats: bool = False, default_size: int = 256):
        super().__init__()
        assert variant in VARIANTS
        self.variant, self.require_feats, self.default_size = variant, bool(require_feats), default_size
        D = default_size
        E = D + (D if require_feats else 0)
        self.embed_size = E
        if variant in ("v1", "b2h"):
            enc = D
            if require_feats and variant == "v1":
                self.text_embeds_postprocess = _proj(512, D)
                self.text_reduce = nn.Sequential(nn.MaxPool1d(2, 2))
            if require_feats and variant == "b2h":
                self.image_resnet_postprocess = _proj(2000, D)
                self.image_reduce = nn.Sequential(nn.MaxPool1d(2, 2))
        else:
            enc = E
            if require_feats:
                self.text_embeds_postprocess = _proj(512, E if variant == "v2" else E // 2)
        self.encoder = nn.Sequential(*_block(feature_in_dim, enc, 3), nn.MaxPool1d(2, 2))
        self.conv5 = nn.Sequential(*_block(E, E, 3))
        self.conv6 = nn.Sequential(*_block(E, E, 3))
        narrow = E // 2 if (require_feats and variant in ("v4", "v4_deeper")) else E
        self.conv7 = nn.Sequential(*_block(E, narrow if variant == "v4" else E, 5, stride=2, pad=2))
        if variant == "v4_deeper":
            # dead branch: computed by the reference, never reaches the output
            self.conv8 = nn.Sequential(*_block(E, E, 3))
            self.conv9 = nn.Sequential(*_block(E, narrow, 3))
            self.conv10 = nn.Sequential(*_block(narrow, narrow, 3))
            self.skip1 = nn.Sequential(*_block(E, E, 3))
            self.skip2 = nn.Sequential(*_block(E, E, 3))
            self.skip3 = nn.Sequential(*_block(E, E, 3))
            self.skip4 = nn.Sequential(*_block(E, E, 3))
        else:
            self.skip4 = nn.Sequential(*_block(E, E, 3))
            self.skip5 = nn.Sequential(*_block(E, E, 3))
        self.decoder = _decoder(E, feature_out_dim)
        self.mask_store: Optional[Dict[str, torch.Tensor]] = None
        for name, m in self.named_modules():
            if isinstance(m, ReplayDropout):
                m.site = name

    # ---- dropout mask replay -------------------------------------------------
    def set_masks(self, masks: Optional[Dict[str, torch.Tensor]]):
        for m in self.modules():
            if isinstance(m, ReplayDropout):
                m.store = masks

    @staticmethod
    def upsample(t, shape):  # modelZoo.py:295-296
        return t.repeat_interleave(2, dim=2)[:, :, :shape[2]]

    def _proj_rows(self, seq, rows, B, T):
        feat = seq(rows)                      # (B*T, D)
        return feat.view(B, T, -1).permute(0, 2, 1).contiguous()

    def forward(self, input_, audio_=None, percent_rand_=0.7, feats_=None):
        B, T = input_.shape[0], input_.shape[2]
        v, rf = self.variant, self.require_feats
        fourth = self.encoder(input_)
        if rf and v == "v1":      # modelZoo.py:284-292,303-306
            rows = feats_.unsqueeze(1).repeat(1, T, 1).view(-1, feats_.shape[-1])
            feat = self.text_reduce(self._proj_rows(self.text_embeds_postprocess, rows, B, T))
            fourth = torch.cat((fourth, feat), dim=1)
        if rf and v == "b2h":     # modelZoo.py:122-129,141-144
            rows = feats_.reshape(-1, 2000)
            feat = self.image_reduce(self._proj_rows(self.image_resnet_postprocess, rows, B, T))
            fourth = torch.cat((fourth, feat), dim=1)
        fifth = self.conv5(fourth)
        sixth = self.conv6(fifth)
        seventh = self.conv7(sixth)
        if v == "v4_deeper":      # modelZoo.py:683-710 (dead branch evaluated for BN side effects)
            eighth = self.conv8(seventh)
            ninth = self.conv9(eighth)
            tenth = self.conv10(ninth)
            ninth = tenth + ninth
            if rf:
                T4 = ninth.shape[2]
                rows = feats_.unsqueeze(1).repeat(1, T4, 1).view(-1, feats_.shape[-1])
                ninth = torch.cat((ninth, self._proj_rows(self.text_embeds_postprocess, rows, B, T4)), dim=1)
            ninth = self.skip1(ninth)
            eighth = ninth + eighth
            eighth = self.skip2(eighth)
            sixth = self.upsample(seventh, sixth.shape) + sixth
            sixth = self.skip3(sixth)
            fifth = sixth + fifth
            fifth = self.skip4(fifth)
            return self.decoder(fifth)
        if rf and v == "v2":      # modelZoo.py:408-415,429-431: text appended as an extra TIME step
            feat = self._proj_rows(self.text_embeds_postprocess, feats_, B, 1)
            seventh = torch.cat((seventh, feat), dim=2)
        if rf and v == "v4":      # modelZoo.py:521-528,541-545
            T4 = seventh.shape[2]
            rows = feats_.unsqueeze(1).repeat(1, T4, 1).view(-1, feats_.shape[-1])
            seventh = torch.cat((seventh, self._proj_rows(self.text_embeds_postprocess, rows, B, T4)), dim=1)
        sixth = self.upsample(seventh, sixth.shape) + sixth
        sixth = self.skip4(sixth)
        fifth = sixth + fifth
        fifth = self.skip5(fifth)
        return self.decoder(fifth)


class RefDiscriminator(nn.Module):
    """modelZoo.py:763-817."""

    def __init__(self, feature_in_dim: int):
        super().__init__()
        chans = [feature_in_dim, 64, 64, 32, 32, 16, 16, 8]
        layers = []
        for i in range(7):
            layers += _block(chans[i], chans[i + 1], 5, stride=2, pad=2)
        layers += [ReplayDropout(0.5), nn.Conv1d(8, 1, 3, padding=1)]
        self.convs = nn.Sequential(*layers)
        for name, m in self.named_modules():
            if isinstance(m, ReplayDropout):
                m.site = name

    def set_masks(self, masks):
        for m in self.modules():
            if isinstance(m, ReplayDropout):
                m.store = masks

    def forward(self, input_):
        return self.convs(input_)


def build_generator(variant, feature_in_dim, feature_out_dim, require_feats=False, default_size=256):
    return RefGenerator(variant, feature_in_dim, feature_out_dim, require_feats, default_size)


def build_discriminator(feature_in_dim):
    return RefDiscriminator(feature_in_dim)


def dropout_sites(model) -> Dict[str, None]:
    return {m.site: None for m in model.modules() if isinstance(m, ReplayDropout)}


def make_masks(model, example_inputs, seed: int, feats=None) -> Dict[str, torch.Tensor]:
    """Draw one Bernoulli(0.5) keep-mask per dropout site by tracing input shapes once."""
    shapes = {}
    hooks = []
    for m in model.modules():
        if isinstance(m, ReplayDropout):
            hooks.append(m.register_forward_pre_hook(
                lambda mod, inp: shapes.__setitem__(mod.site, tuple(inp[0].shape))))
    was_training = model.training
    model.eval()
    with torch.no_grad():
        if feats is not None:
            model(example_inputs, feats_=feats)
        else:
            model(example_inputs)
    model.train(was_training)
    for h in hooks:
        h.remove()
    g = torch.Generator().manual_seed(seed)
    return {k: (torch.rand(s, generator=g) < 0.5).to(torch.uint8) for k, s in sorted(shapes.items())}


# ---------------------------------------------------------------------------------------------
# step bodies
# ---------------------------------------------------------------------------------------------
def calc_motion(t):  # train_gan.py:209-211 (frame 0 minus frames 0..T-2; reproduced bug and all)
    return t[:, :, :1] - t[:, :, :-1]


def reg_criterion(loss: str, out, y):
    """LOSSES[--loss] as train_gan.py:74-77,286-292 evaluates it (utils/constants.py:53-58).  "RobustLoss": the
    AdaptiveLossFunction's latent alpha / scale never reach the optimiser (train_gan.py:69 is built from
    generator.parameters() only), so they keep their initial values alpha = 2, scale = 1/2
    (utils/robust_loss/adaptive.py:55-59) and general.lossfun reduces to (d / scale)^2 / 2, to which
    lossfun adds log(scale) + log Z(2) = log(1/2) + log sqrt(2 pi).  tests/test_oracle_vs_reference.py pins this
    against the real class."""
    if loss == "L1":
        return nn.functional.l1_loss(out, y)
    if loss == "L2":
        return nn.functional.mse_loss(out, y)
    if loss == "Huber1":
        return nn.functional.huber_loss(out, y, delta=1.0)
    if loss == "RobustLoss":
        d = (out - y).reshape(out.shape[0], -1)
        return torch.mean(0.5 * (d / 0.5) ** 2 + (math.log(0.5) + 0.5 * math.log(2.0 * math.pi)))
    raise KeyError(loss)


def generator_step(G, D, g_opt, x, y, feats=None, g_masks=None, loss="L1"):
    """train_gan.py:260-299 for one batch. Returns (g_loss, l1, adv, output); `l1` is the regression term."""
    D.eval()
    G.train()
    G.set_masks(g_masks)
    out = G(x, feats_=feats)
    fake_motion = calc_motion(out)
    with torch.no_grad():
        fake_score = D(fake_motion)
    fake_score = fake_score.detach()
    l1 = reg_criterion(loss, out, y)
    adv = nn.functional.mse_loss(fake_score, torch.ones_like(fake_score))
    g_loss = l1 + adv
    g_opt.zero_grad()
    g_loss.backward()
    g_opt.step()
    G.set_masks(None)
    return g_loss.detach(), l1.detach(), adv.detach(), out.detach()


def discriminator_step(G, D, d_opt, x, y, feats=None, d_masks_fake=None, d_masks_real=None,
                       label_smooth=False):
    """train_gan.py:216-251 for one batch. Returns (d_loss, fake_score, real_score)."""
    G.eval()
    D.train()
    with torch.no_grad():
        fake = G(x, feats_=feats).detach()
    fake_motion, real_motion = calc_motion(fake), calc_motion(y)
    D.set_masks(d_masks_fake)
    fake_score = D(fake_motion)
    D.set_masks(d_masks_real)
    real_score = D(real_motion)
    D.set_masks(None)
    tf, tr = (0.1, 0.9) if label_smooth else (0.0, 1.0)
    d_loss = nn.functional.mse_loss(fake_score, torch.full_like(fake_score, tf)) + \
        nn.functional.mse_loss(real_score, torch.full_like(real_score, tr))
    d_opt.zero_grad()
    d_loss.backward()
    d_opt.step()
    return d_loss.detach(), fake_score.detach(), real_score.detach()


def macs_per_clip(model, T: int, feats_kind: Optional[str] = None) -> int:
    """Hook-measured MAC count of one forward over a single T-frame clip."""
    total = [0]

    def hook(m, inp, out):
        if isinstance(m, nn.Conv1d):
            total[0] += out.shape[0] * out.shape[1] * out.shape[2] * m.in_channels * m.kernel_size[0]
        elif isinstance(m, nn.ConvTranspose1d):
            total[0] += inp[0].shape[0] * inp[0].shape[2] * m.in_channels * m.out_channels * m.kernel_size[0]
        elif isinstance(m, nn.Linear):
            total[0] += out.shape[0] * m.in_features * m.out_features

    hs = [m.register_forward_hook(hook) for m in model.modules()
          if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d, nn.Linear))]
    was = model.training
    model.eval()
    with torch.no_grad():
        cin = model.encoder[1].in_channels if hasattr(model, "encoder") else model.convs[1].in_channels
        x = torch.zeros(1, cin, T)
        if feats_kind == "text":
            model(x, feats_=torch.zeros(1, 512))
        elif feats_kind == "image":
            model(x, feats_=torch.zeros(1, T, 2000))
        else:
            model(x)
    model.train(was)
    for h in hs:
        h.remove()
    return total[0]
