"""ORACLE (test infrastructure only) -- CPU restatement of every libb2h op contract.

Each function interprets one op record (see b2h_b200/program.py) with plain torch on CPU, following the
contract written in include/b2h_abi.h.  Two uses, both in tests only:
  * GPU tests check each CUDA kernel against its restatement on random inputs;
  * CPU tests interpret whole recorded programs (graph wiring, backward formulas) and compare the
    result with oracle/ref_models.py, so the host logic is validated without a GPU.
The product never imports this module.

Arithmetic is fp32 with the results rounded to the activation dtype on store, like the kernels.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

# mirrored constants of include/b2h_abi.h
ACT_NONE, ACT_LEAKY, ACT_RELU = 0, 1, 2
ROW_IDENT, ROW_UP2, ROW_POOL2, ROW_BCAST = 0, 1, 2, 3
SRC_NCL, SRC_ROWS, SRC_BCAST, SRC_MOTION = 0, 1, 2, 3
DROP_NONE, DROP_MASK, DROP_PHILOX = 0, 1, 2
BWD_COPIES = 8   # B2H_BWD_COPIES: copies of the first-pass accumulators (region 1 of a bn_bwd workspace)
(OP_GEMM, OP_WGRAD, OP_BN_STATS, OP_BN_APPLY, OP_BN_BWD, OP_PREP, OP_TO_NCL, OP_L1, OP_MSE, OP_COLSUM,
 OP_ADAM, OP_PACK, OP_BN_FOLD, OP_ROT6D, OP_FILL, OP_PACK_MULTI, OP_BN_FOLD_MULTI, OP_FK, OP_DP_ADAM) = range(1, 20)


def _act(x, act):
    if act == ACT_LEAKY:
        return torch.where(x > 0, x, x * 0.2)
    if act == ACT_RELU:
        return torch.clamp_min(x, 0)
    return x


def _dact(out, act):
    if act == ACT_LEAKY:
        return torch.where(out > 0, torch.ones_like(out), torch.full_like(out, 0.2))
    if act == ACT_RELU:
        return (out > 0).to(out.dtype)
    return torch.ones_like(out)


def _drop_scale(drop, rows, C, row0=0):
    """(rows, C) multiplier of a dropout site restricted to rows [row0, row0+rows)."""
    if not drop or drop.get("mode", 0) == DROP_NONE:
        return None
    if drop["mode"] == DROP_MASK:
        m = drop["mask"].reshape(-1, C)[row0:row0 + rows]
        return m.to(torch.float32) * 2.0
    raise NotImplementedError("Philox dropout is not emulated on CPU (statistical test on GPU)")


def _rows2d(t):
    return t.reshape(-1, t.shape[-1])


# ---------------------------------------------------------------------------------------------
def gemm(f: Dict):
    A = f["A"].to(torch.float32)                      # (B, La, lda)
    B, La, Kc = f["B"], f["La"], f["Kc"]
    Lo, nt, stride = f["Lo"], f["ntaps"], f["stride"]
    W = f["W"].to(torch.float32).reshape(f["Npad"], nt, Kc)
    A = A.reshape(B, La, -1)[:, :, :Kc]
    acc = torch.zeros(B, Lo, f["Npad"])
    lo = torch.arange(Lo)
    for t in range(nt):
        li = lo * stride + f["tap_off"][t]
        ok = (li >= 0) & (li < La)
        if not ok.any():
            continue
        rows = torch.zeros(B, Lo, Kc)
        rows[:, ok] = A[:, li[ok]]
        acc += rows @ W[:, t, :].T
    nph = f["nphase"]
    half = f["Npad"] // nph
    Nv, Lact = f["Nvalid"], f["Lo_actual"]
    out = f["out"]
    ncl = f["out_f32"] == 2     # the op writes the (B, Nvalid, Lo_actual) fp32 NCL tensor itself (b2h_abi.h)
    if ncl:
        assert nph == 1 and f["out_coff"] == 0 and tuple(out.shape) == (B, Nv, Lact)
        out2 = torch.zeros(B, Lact, Nv)
    else:
        ldo = out.shape[-1]
        up2 = bool(f.get("resid_up2")) and f.get("resid") is not None
        out2 = None if f.get("out_pool2") else out.reshape(B, Lact * (2 if up2 else 1), ldo)
        assert ldo == f["ldo"]
    for ph in range(nph):
        v = acc[:, :, ph * half: ph * half + Nv]
        if f.get("bias") is not None:
            v = v + f["bias"][:Nv]
        v = _act(v, f["act"])
        if f.get("post_scale") is not None:
            v = v * f["post_scale"].reshape(-1)[:Nv] + f["post_shift"].reshape(-1)[:Nv]
        rows_act = lo * nph + ph
        ok = rows_act < Lact
        v = v[:, ok]
        ra = rows_act[ok]
        drop = f.get("drop")
        if drop and drop.get("mode", 0) != DROP_NONE:
            C = f["drop_C"]
            m = _drop_scale(drop, B * Lact, C).reshape(B, Lact, C)[:, ra, :]
            nn = min(Nv, C)
            v = v.clone()
            v[:, :, :nn] = v[:, :, :nn] * m[:, :, :nn]
        if f.get("grad_add") is not None:   # dgrad of a skip connection: + the other consumer's gradient (b2h_gemm_t.grad_add)
            assert f["out_coff"] == 0 and not ncl and f.get("resid") is None and not f.get("out_pool2")
            ga = f["grad_add"].to(torch.float32).reshape(B, Lact, f["ld_grad_add"])[:, :, :Nv]
            v = v + ga[:, ra]
        if f.get("out_pool2"):           # MaxPool1d(2) in the epilogue: out has Lo_actual // 2 rows per sample
            assert nph == 1 and f["out_coff"] == 0 and not ncl and f.get("resid") is None
            Lh = Lact // 2
            full = torch.zeros(B, Lact, Nv)
            full[:, ra] = v
            pooled = torch.maximum(full[:, 0:2 * Lh:2], full[:, 1:2 * Lh:2])
            out.reshape(B, Lh, ldo)[:, :, :Nv] = pooled.to(out.dtype)
            continue
        if f.get("resid") is not None:   # residual add in the epilogue (b2h_gemm_t.resid), optionally after x2 up-sampling
            assert nph == 1 and f["out_coff"] == 0 and not ncl
            rs = f["resid"].to(torch.float32).reshape(B, -1, f["ld_resid"])[:, :, :Nv]
            if f.get("resid_up2"):
                for k in range(2):
                    out2[:, 2 * ra + k, :Nv] = (v + rs[:, 2 * ra + k]).to(out.dtype)
            else:
                out2[:, ra, :Nv] = (v + rs[:, ra]).to(out.dtype)
            continue
        out2[:, ra, f["out_coff"]: f["out_coff"] + Nv] = v.to(out.dtype)
    if ncl:
        out.copy_(out2.permute(0, 2, 1))
    st = f.get("stats")
    if st and st.get("z") is not None:   # the op also produces the batch statistics of its output
        bn_stats(st)
    bs = f.get("bwd_sums")
    if bs and bs.get("z") is not None:   # first pass of the producer layer's BN backward over this op's output
        C, G, Lz = bs["C"], bs["groups"], bs["Lz"]
        assert f["out_coff"] == 0 and not f["out_f32"] and nph * 0 == 0
        g = out2[:, :, :C].to(torch.float32)                           # (B, Lo_actual, C) as stored
        z = bs["z"].to(torch.float32).reshape(B, Lz, -1)[:, :, :C]
        pooled = bs["rowmap"] == ROW_POOL2
        if bs["rowmap"] == ROW_UP2:
            assert Lact == 2 * Lz
            z = z.repeat_interleave(2, dim=1)
        elif pooled:
            assert Lz == 2 * Lact     # the row of each pair with the larger z*scale + shift (the first on ties)
        else:
            assert bs["rowmap"] == ROW_IDENT and Lact == Lz
        Bg = B // G
        acc = bs["accum"].reshape(-1, G, C, 2)
        for gi in range(G):
            sl = slice(gi * Bg, (gi + 1) * Bg)
            zg = z[sl]
            if pooled:
                y = zg * bs["scale"][gi, :C] + bs["shift"][gi, :C]
                zg = torch.where(y[:, 1::2] > y[:, 0::2], zg[:, 1::2], zg[:, 0::2])
            zh = (zg - bs["mean"][gi, :C]) * bs["invstd"][gi, :C]
            acc[0, gi, :, 0] += g[sl].sum((0, 1)).double()
            acc[0, gi, :, 1] += (g[sl] * zh).sum((0, 1)).double()


def wgrad(f: Dict):
    B, Lp, Lq = f["B"], f["Lp"], f["Lq"]
    P = f["P"].to(torch.float32).reshape(B, Lp, -1)[:, :, :f["Mvalid"]]
    Q = f["Q"].to(torch.float32).reshape(B, Lq, -1)[:, :, :f["Nvalid"]]
    dW = f["dW"].reshape(f["Mvalid"], f["Nvalid"], f["ntaps"])
    r = torch.arange(Lp)
    for t in range(f["ntaps"]):
        rq = r * f["stride"] + f["tap_off"][t]
        ok = (rq >= 0) & (rq < Lq)
        if not ok.any():
            dW[:, :, t] = 0
            continue
        Pm = P[:, ok].reshape(-1, f["Mvalid"])
        Qm = Q[:, rq[ok]].reshape(-1, f["Nvalid"])
        dW[:, :, t] = Pm.T @ Qm


def bn_stats(f: Dict):
    C, G, rpg = f["C"], f["groups"], f["rows_per_group"]
    z = _rows2d(f["z"]).to(torch.float32)[:, :C].reshape(G, rpg, C)
    gamma = f["gamma"][:C] if f.get("gamma") is not None else torch.ones(C)
    beta = f["beta"][:C] if f.get("beta") is not None else torch.zeros(C)
    for g in range(G):
        mean = z[g].mean(0)
        var_b = z[g].var(0, unbiased=False)
        invstd = 1.0 / torch.sqrt(var_b + f["eps"])
        f["mean"][g, :C] = mean
        f["invstd"][g, :C] = invstd
        f["scale"][g, :C] = invstd * gamma
        f["shift"][g, :C] = beta - mean * (invstd * gamma)
        if f.get("running_mean") is not None and (g == 0 or f["update_all_groups"]):
            var_u = z[g].var(0, unbiased=True) if rpg > 1 else var_b
            m = f["momentum"]
            f["running_mean"].mul_(1 - m).add_(m * mean)
            f["running_var"].mul_(1 - m).add_(m * var_u)
    if f.get("running_mean") is not None and f.get("num_batches_tracked") is not None:
        f["num_batches_tracked"] += G if f["update_all_groups"] else 1


def _bn_src_eval(src, B, L, C, groups):
    """(B, L, C) tensor of BN(src) = z*scale + shift under the source's row map."""
    z = src["z"].to(torch.float32)
    Ls = src["L_src"]
    o = src["coff"]
    z = z.reshape(B, Ls, -1)[:, :, o: o + C]
    Bg = B // groups
    y = torch.empty_like(z)
    for g in range(groups):
        s, t = src["scale"][g, o:o + C], src["shift"][g, o:o + C]
        y[g * Bg:(g + 1) * Bg] = z[g * Bg:(g + 1) * Bg] * s + t
    rm = src["rowmap"]
    if rm == ROW_IDENT:
        assert Ls == L
        return y
    if rm == ROW_UP2:
        return y.repeat_interleave(2, dim=1)[:, :L]
    if rm == ROW_POOL2:
        assert L == Ls // 2
        y0, y1 = y[:, 0:2 * L:2], y[:, 1:2 * L:2]
        return torch.where(y1 > y0, y1, y0)
    if rm == ROW_BCAST:
        return y[:, :1].expand(B, L, C)
    raise ValueError(rm)


def bn_apply(f: Dict):
    B, L, C, G = f["B"], f["L"], f["C"], f["groups"]
    y = torch.zeros(B, L, C)
    for i in range(f["nsrc"]):
        y = y + _bn_src_eval(f["src"][i], B, L, C, G)
    drop = f.get("drop")
    if drop and drop.get("mode", 0) != DROP_NONE:
        m = _drop_scale(drop, B * L, f["drop_C"]).reshape(B, L, f["drop_C"])
        y = y * m[:, :, f["drop_coff"]: f["drop_coff"] + C]
    out = f["out"].reshape(B, L, -1)
    o = f["out_coff"]
    out[:, :, o:o + C] = y.to(out.dtype)
    out[:, :, o + C:o + f["Cfill"]] = 0


def _grad_src(gs, B, L, C, bnf, aff):
    """dy contribution (B, L, C) of one gradient source."""
    g = gs["g"].to(torch.float32).reshape(B, gs["L_src"], -1)[:, :, gs["coff"]: gs["coff"] + C]
    rm = gs["rowmap"]
    if rm == ROW_IDENT:
        return g
    if rm == ROW_UP2:
        Ls = gs["L_src"]
        out = torch.zeros(B, L, C)
        n0 = (Ls + 1) // 2
        out[:, :n0] += g[:, 0::2]
        out[:, :Ls // 2] += g[:, 1::2]
        return out
    if rm == ROW_POOL2:
        z, s, t = aff
        y = z * s + t
        Lp = gs["L_src"]
        y0, y1 = y[:, 0:2 * Lp:2], y[:, 1:2 * Lp:2]
        sel1 = y1 > y0
        out = torch.zeros(B, L, C)
        out[:, 0:2 * Lp:2] = torch.where(sel1, torch.zeros_like(g), g)
        out[:, 1:2 * Lp:2] = torch.where(sel1, g, torch.zeros_like(g))
        return out
    raise ValueError(rm)


def bn_bwd(f: Dict):
    B, L, C, G = f["B"], f["L"], f["C"], f["groups"]
    bn = f["bn"]
    z = bn["z"].to(torch.float32).reshape(B, L, -1)[:, :, bn["coff"]: bn["coff"] + C]
    Bg = B // G
    dpre = f["dpre"].reshape(B, L, -1) if f.get("dpre") is not None else None
    dgamma, dbeta, dbias = torch.zeros(C), torch.zeros(C), torch.zeros(C)
    for g in range(G):
        sl = slice(g * Bg, (g + 1) * Bg)
        mean, invstd = bn["mean"][g, :C], bn["invstd"][g, :C]
        s, t = bn["scale"][g, :C], bn["shift"][g, :C]
        zg = z[sl]
        dy = torch.zeros(Bg, L, C)
        for i in range(f["ngsrc"]):
            gs = dict(f["gsrc"][i])
            gfull = gs["g"].reshape(B, gs["L_src"], -1)
            gs["g"] = gfull[sl]
            dy = dy + _grad_src(gs, Bg, L, C, bn, (zg, s, t))
        zh = (zg - mean) * invstd
        n = Bg * L
        sdy = dy.sum((0, 1))
        sdyz = (dy * zh).sum((0, 1))
        if f.get("first_pass_only"):     # b2h_bn_bwd_t.first_pass_only: this source's share of the sums, nothing else
            acc = f["accum"].reshape(-1, G, C, 2)
            acc[0, g, :, 0] += sdy.double()
            acc[0, g, :, 1] += sdyz.double()
            continue
        if f.get("accum") is not None:   # the first-pass sums come from the GEMMs that wrote the sources
            acc = f["accum"].reshape(-1, G, C, 2)
            sdy, sdyz = acc[:, g, :, 0].sum(0).float(), acc[:, g, :, 1].sum(0).float()
        dz = s * (dy - sdy / n - zh * (sdyz / n))
        dp = dz * _dact(zg, f["act"])
        dpre[sl, :, :C] = dp.to(dpre.dtype)
        dpre[sl, :, C:f["Cfill"]] = 0
        dgamma += sdyz
        dbeta += sdy
        dbias += dp.sum((0, 1))
        if f.get("defer") == 2:          # the sums of dpre go to the op's own accumulators (region 2 of `partial`)
            acc2 = f["partial"].view(torch.float64)[BWD_COPIES * G * C * 2:][: 16 * G * C].reshape(16, G, C)
            acc2[0, g] += dp.sum((0, 1)).double()
    if f.get("first_pass_only") or f.get("defer"):   # defer: dpre only; b2h_colsum(bn_accum) finishes the op
        return
    if f.get("accum") is not None:
        f["accum"].zero_()
    if f.get("dgamma") is not None:
        f["dgamma"].copy_(dgamma)
    if f.get("dbeta") is not None:
        f["dbeta"].copy_(dbeta)
    if f.get("dbias") is not None:
        f["dbias"].copy_(dbias)


def prep(f: Dict):
    B, L, C = f["B"], f["L"], f["C"]
    src = f["src"]
    k = f["kind"]
    if k == SRC_NCL:
        v = src.reshape(B, C, L).permute(0, 2, 1)
    elif k == SRC_MOTION:
        x = src.reshape(B, C, L + 1)
        v = (x[:, :, :1] - x[:, :, :-1]).permute(0, 2, 1)
    elif k == SRC_ROWS:
        v = src.reshape(B * L, -1)[:, :C].reshape(B, L, C)
    else:
        v = src.reshape(B, -1)[:, :C].unsqueeze(1).expand(B, L, C)
    v = v.to(torch.float32)
    m = _drop_scale(f.get("drop"), B * L, C)
    if m is not None:
        v = v * m.reshape(B, L, C)
    out = f["out"].reshape(B, L, -1)
    out[:, :, :C] = v.to(out.dtype)
    out[:, :, C:f["Cfill"]] = 0


def to_ncl(f: Dict):
    B, L, C = f["B"], f["L"], f["C"]
    f["dst"].reshape(B, C, L).copy_(f["src"].reshape(B, L, -1)[:, :, :C].permute(0, 2, 1).to(torch.float32))


def l1(f: Dict):
    B, C, L = f["B"], f["C"], f["L"]
    if f.get("out_blc") is not None:   # the prediction arrives BLC fp32; the op also writes the NCL tensor (b2h_abi.h)
        f["out"].reshape(B, C, L).copy_(f["out_blc"].reshape(B, L, -1)[:, :, :C].permute(0, 2, 1))
    o, g = f["out"].reshape(B, C, L), f["gt"].reshape(B, C, L)
    d = o - g
    kind = f.get("kind", 0)   # B2H_LOSS_* of include/b2h_abi.h: 0 L1, 1 L2, 2 Huber(delta 1), 3 frozen adaptive loss
    if kind == 0:
        val, grad = d.abs(), torch.sign(d)
    elif kind == 2:
        val, grad = torch.where(d.abs() < 1, 0.5 * d * d, d.abs() - 0.5), d.clamp(-1, 1)
    else:
        w = 2.0 if kind == 3 else 1.0
        val, grad = w * d * d, 2.0 * w * d
    const = 0.22579135264472743 if kind == 3 else 0.0   # log(1/2) + log sqrt(2 pi)
    f["loss"][0] = (val.double().mean() + const).float()
    if f.get("dout") is not None:
        gv = f["gscale"] / d.numel()
        dout = f["dout"].reshape(B, L, -1)
        dout[:, :, :C] = (grad * gv).permute(0, 2, 1).to(dout.dtype)
        dout[:, :, C:f["Cfill"]] = 0
        if f.get("dbias") is not None:   # bias gradient of the output layer: column sums of dout as stored
            f["dbias"].copy_(dout[:, :, :C].double().sum((0, 1)).float())


def mse(f: Dict):
    G, n, ld = f["groups"], f["n"], f["ld"]
    s = f["score"].reshape(-1)[: G * n * ld: ld].reshape(G, n)
    total = 0.0
    for g in range(G):
        diff = s[g] - f["target"][g]
        total += float((diff.double() ** 2).mean())
        if f.get("dscore") is not None:
            f["dscore"].reshape(-1)[g * n * ld:(g + 1) * n * ld: ld] = 2.0 * diff / n
        if f.get("dpre") is not None:
            pl = f["dpre_ld"]
            f["dpre"].reshape(-1)[g * n * pl:(g + 1) * n * pl: pl] = (2.0 * diff / n).to(f["dpre"].dtype)
    if f.get("dbias") is not None:
        pl = f["dpre_ld"]
        f["dbias"][0] = float(f["dpre"].reshape(-1)[: G * n * pl: pl].double().sum())
    f["loss"][0] = total
    if f.get("total") is not None:
        f["total"][0] = total + (float(f["add"][0]) if f.get("add") is not None else 0.0)


def colsum(f: Dict):
    if f.get("src") is None:            # b2h_colsum_t.src = NULL: the sums of dpre wait in the bn_bwd's own accumulators
        C, G = f["C"], f["bn_groups"]
        acc2 = f["partial"].view(torch.float64)[BWD_COPIES * G * C * 2:][: 16 * G * C].reshape(16, G, C)
        f["out"].copy_(acc2.sum(0).sum(0).float())
        acc2.zero_()
    else:
        src = _rows2d(f["src"]).to(torch.float32)[: f["rows"], : f["C"]]
        f["out"].copy_(src.double().sum(0).float())
    if f.get("bn_accum") is not None:   # finishes a deferred BatchNorm backward (b2h_colsum_t.bn_accum)
        C = f["C"]
        acc = f["bn_accum"].reshape(-1, f["bn_groups"], C, 2)
        tot = acc.sum(0).sum(0)         # copies, then groups
        if f.get("dbeta") is not None:
            f["dbeta"].copy_(tot[:, 0].float())
        if f.get("dgamma") is not None:
            f["dgamma"].copy_(tot[:, 1].float())
        f["bn_accum"].zero_()


def adam(f: Dict):
    phase = f.get("phase", 0)
    if phase != 2:
        f["step"] += 1
    if phase == 1:
        return
    t = int(f["step"][0])
    b1, b2, lr, eps = f["beta1"], f["beta2"], f["lr"], f["eps"]
    g = f["g"] * f["gscale"]
    f["m"].lerp_(g, 1 - b1)
    f["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    denom = (f["v"].sqrt() / math.sqrt(bc2)).add_(eps)
    f["p"].addcdiv_(f["m"], denom, value=-(lr / bc1))


def dp_slice(n: int, world: int, rank: int):
    """The element range of the flat buffer that `rank` owns in b2h_dp_adam: ceil(n/4 / world) float4s per rank."""
    n4 = n // 4
    chunk = -(-n4 // world)
    lo = min(chunk * rank, n4)
    return 4 * lo, 4 * min(lo + chunk, n4)


def dp_adam(f: Dict):
    """b2h_dp_adam_t as ONE rank executes it: gradient slice summed over the ranks in rank order, Adam phase 2 on the
    owned slice (local moments), the new parameters stored into every rank's buffer.  `_p` / `_g` hold every rank's
    range as tensors (the descriptor itself carries peer pointers), `_step` / `_lr` what b2h_adam phase 1 turned into
    `scalars`.  Interpreting the op of every rank, in any order, is the whole collective."""
    lo, hi = dp_slice(f["n"], f["world"], f["rank"])
    if hi <= lo:
        return
    g = torch.zeros(hi - lo)
    for q in range(f["world"]):
        g = g + f["_g"][q][lo:hi]
    g = g * f["gscale"]
    t = int(f["_step"][0])
    b1, b2, lr, eps = f["beta1"], f["beta2"], f["_lr"], f["eps"]
    m, v = f["m"][lo:hi], f["v"][lo:hi]
    m.lerp_(g, 1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    newp = f["_p"][f["rank"]][lo:hi].addcdiv(m, denom, value=-(lr / bc1))
    for q in range(f["world"]):
        f["_p"][q][lo:hi] = newp


def pack(f: Dict):
    W = f["W"].reshape(-1)
    nph, Op, nt, Ip = f["nphase"], f["Opad"], f["ntaps"], f["Ipad"]
    out = torch.zeros(nph, Op, nt, Ip)
    o = torch.arange(f["O"]).view(-1, 1)
    i = torch.arange(f["I"]).view(1, -1)
    for ph in range(nph):
        for t in range(nt):
            k = f["tapmap"][ph][t]
            if k < 0:
                continue
            idx = o * f["o_stride"] + i * f["i_stride"] + k * f["k_stride"]
            out[ph, : f["O"], t, : f["I"]] = W[idx]
    f["out"].reshape(nph, Op, nt, Ip).copy_(out.to(f["out"].dtype))
    if f.get("out_bias") is not None:
        f["out_bias"].zero_()
        if f.get("bias") is not None:
            f["out_bias"][: f["O"]] = f["bias"]


def pack_multi(f: Dict):
    for item in f["_items"]:
        pack(item)


def bn_fold(f: Dict):
    C = f["C"]
    invstd = 1.0 / torch.sqrt(f["running_var"][:C] + f["eps"])
    s = invstd * (f["gamma"][:C] if f.get("gamma") is not None else 1.0)
    scale, shift = f["scale"].reshape(-1), f["shift"].reshape(-1)
    scale[: f["Cpad"]] = 0
    shift[: f["Cpad"]] = 0
    scale[:C] = s
    shift[:C] = (f["beta"][:C] if f.get("beta") is not None else 0.0) - f["running_mean"][:C] * s


def bn_fold_multi(f: Dict):
    for item in f["_items"]:
        bn_fold(item)


def rot6d(f: Dict):
    f["mat"].reshape(-1, 9).copy_(rot6d_to_mat(f["r6d"].reshape(-1, 6)))


def rot6d_to_mat(r6d: torch.Tensor) -> torch.Tensor:
    """Row-wise restatement of np_rot6d_to_mat (utils/conversion_utils.py:86-107; SURVEY S10: the
    reference function is only correct one row at a time, which is how it is called at :36-37)."""
    x_raw, y_raw = r6d[:, 0:3], r6d[:, 3:6]
    x = x_raw / (x_raw.norm(dim=1, keepdim=True) + 1e-6)
    z = torch.linalg.cross(x, y_raw)
    z = z / (z.norm(dim=1, keepdim=True) + 1e-6)
    y = torch.linalg.cross(z, x)
    return torch.stack([x, y, z], dim=-1).reshape(-1, 9)


def fill(f: Dict):
    raise NotImplementedError


DISPATCH = {OP_GEMM: gemm, OP_WGRAD: wgrad, OP_BN_STATS: bn_stats, OP_BN_APPLY: bn_apply, OP_BN_BWD: bn_bwd,
            OP_PREP: prep, OP_TO_NCL: to_ncl, OP_L1: l1, OP_MSE: mse, OP_COLSUM: colsum, OP_ADAM: adam,
            OP_PACK: pack, OP_BN_FOLD: bn_fold, OP_ROT6D: rot6d, OP_FILL: fill,
            OP_PACK_MULTI: pack_multi, OP_BN_FOLD_MULTI: bn_fold_multi, OP_DP_ADAM: dp_adam}


def run_records(recs, first=0, end=None):
    """Interpret op records [first, end) on CPU."""
    for rec in recs[first: end if end is not None else len(recs)]:
        DISPATCH[rec.kind](rec.f)
