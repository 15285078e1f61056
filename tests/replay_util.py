"""Differential replay: run a recorded program op by op on the GPU (through libb2h.so) and on a CPU
twin (oracle/ops_emul.py), compare every output of every op, and "teacher-force" the GPU buffers with
the CPU results so each kernel is judged on identical inputs.  Pinpoints the first kernel that
disagrees with its contract."""
import torch

from b2h_b200 import _lib as L
from oracle import ops_emul as E

OUT_FIELDS = {
    L.OP_GEMM: ["out"], L.OP_WGRAD: ["dW"],
    L.OP_BN_STATS: ["mean", "invstd", "scale", "shift", "running_mean", "running_var", "num_batches_tracked"],
    L.OP_BN_APPLY: ["out"], L.OP_BN_BWD: ["dpre", "dgamma", "dbeta", "dbias"], L.OP_PREP: ["out"],
    L.OP_TO_NCL: ["dst"], L.OP_L1: ["loss", "dout", "dbias"], L.OP_MSE: ["loss", "dscore", "total", "dbias"], L.OP_COLSUM: ["out"],
    L.OP_ADAM: ["p", "m", "v", "step"], L.OP_PACK: ["out", "out_bias"], L.OP_BN_FOLD: ["scale", "shift"],
    L.OP_ROT6D: ["mat"], L.OP_PACK_MULTI: [], L.OP_BN_FOLD_MULTI: [],
}


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def sync_inputs(prog_gpu, prog_cpu):
    """Copy every tensor the CPU program references onto its GPU counterpart (weights, masks, inputs)."""
    seen = set()

    def walk(fg, fc):
        for k, vc in fc.items():
            if k == "descs":
                continue   # device table of pointers: each plan keeps its own
            vg = fg.get(k)
            if isinstance(vc, torch.Tensor):
                if vc.data_ptr() not in seen:
                    seen.add(vc.data_ptr())
                    vg.copy_(vc)
            elif isinstance(vc, dict):
                walk(vg, vc)
            elif isinstance(vc, list):
                for a, b in zip(vg, vc):
                    if isinstance(b, dict):
                        walk(a, b)

    for rg, rc in zip(prog_gpu.recs, prog_cpu.recs):
        walk(rg.f, rc.f)


def replay_pair(prog_gpu, prog_cpu, segments, teacher_force=True):
    """Returns a list of (op index, tag, field, rel_err)."""
    report = []
    assert len(prog_gpu.recs) == len(prog_cpu.recs)
    for seg in segments:
        first, end = prog_gpu.segments[seg]
        assert (first, end) == prog_cpu.segments[seg]
        for i in range(first, end):
            rg, rc = prog_gpu.recs[i], prog_cpu.recs[i]
            assert rg.kind == rc.kind and rg.tag == rc.tag
            prog_gpu.run_range(i, i + 1)
            torch.cuda.synchronize()
            E.DISPATCH[rc.kind](rc.f)
            fields = [(fld, rc.f, rg.f) for fld in OUT_FIELDS[rc.kind]]
            if rc.kind == L.OP_GEMM and (rc.f.get("stats") or {}).get("z") is not None:
                fields += [("stats." + fld, rc.f["stats"], rg.f["stats"]) for fld in OUT_FIELDS[L.OP_BN_STATS]]
            if rc.kind == L.OP_GEMM and (rc.f.get("bwd_sums") or {}).get("z") is not None:
                fields += [("bwd_sums.accum", rc.f["bwd_sums"], rg.f["bwd_sums"])]
            if rc.kind == L.OP_BN_BWD and rc.f.get("first_pass_only"):
                fields += [("accum", rc.f, rg.f)]   # one consumer's share of the first-pass sums
            if rc.kind == L.OP_COLSUM and rc.f.get("bn_accum") is not None:
                fields += [("dgamma", rc.f, rg.f), ("dbeta", rc.f, rg.f)]   # finishes a deferred BatchNorm backward
            if rc.kind == L.OP_L1 and rc.f.get("out_blc") is not None:
                fields += [("out", rc.f, rg.f)]     # the loss also writes the NCL prediction (b2h_l1_t.out_blc)
            for fld, fc, fg in fields:
                vc, vg = fc.get(fld.split(".")[-1]), fg.get(fld.split(".")[-1])
                if vc is None:
                    continue
                got = vg.detach().to("cpu")
                if fld.endswith("accum"):   # the GPU spreads its contributions over the copies: compare the sums
                    n_copies = vc.shape[0]
                    got = got.reshape(n_copies, -1).sum(0)
                    vc_cmp = vc.reshape(n_copies, -1).sum(0)
                    report.append((i, rc.tag, fld, rel_err(got.float(), vc_cmp.float())))
                    if teacher_force:
                        vg.copy_(vc)
                    continue
                if not torch.isfinite(got.float()).all():
                    report.append((i, rc.tag, fld, float("inf")))
                else:
                    report.append((i, rc.tag, fld, rel_err(got.float(), vc.float())))
                if teacher_force:
                    vg.copy_(vc)
    return report


def format_report(report, top=12):
    worst = sorted(report, key=lambda r: -r[3])[:top]
    return "\n".join(f"  op {i:4d} {tag:32s} {fld:12s} rel_err={e:.3e}" for i, tag, fld, e in worst)
