"""Static race check of the multi-stream step schedules (GanTrainer._g_ops / _d_ops / _gan_ops: dependency chain, wgrad
side streams, adversarial scoring branch, optimizer stream, D_k || G_k+1) without a GPU.

CUDA streams and events are replaced by fakes that carry vector clocks, Program.run / run_range log every op with the
clock of the stream it was enqueued on instead of launching it, and every pair of ops that touch overlapping bytes —
at least one of them writing (outputs, statistics, accumulators, tickets, workspaces) — must be ordered by
happens-before (same stream, or an event recorded after the first and waited for before the second).  This is the
property a CUDA graph captured from the same enqueue order inherits; an unordered pair is a data race that the
numerical tests on the device would only catch by luck."""
import contextlib
import itertools

import pytest
import torch

import b2h_b200  # noqa: F401
from b2h_b200 import _lib as L
from b2h_b200 import program as program_mod
from b2h_b200.trainer import GanTrainer

# fields an op writes (everything else it references is read); scratch that is reset in place counts as written
WRITES = {
    L.OP_GEMM: {"out", "stats.mean", "stats.invstd", "stats.scale", "stats.shift", "stats.running_mean",
                "stats.running_var", "stats.num_batches_tracked", "stats.partial", "stats.ticket", "stats.z",
                "bwd_sums.accum", "drop.save"},
    L.OP_WGRAD: {"dW", "partial"},
    L.OP_BN_STATS: {"mean", "invstd", "scale", "shift", "running_mean", "running_var", "num_batches_tracked", "partial",
                    "ticket"},
    L.OP_BN_APPLY: {"out", "drop.save"},
    L.OP_BN_BWD: {"dpre", "dgamma", "dbeta", "dbias", "accum", "partial", "ticket", "sums"},
    L.OP_PREP: {"out", "drop.save"},
    L.OP_TO_NCL: {"dst"},
    L.OP_L1: {"loss", "dout", "dbias", "partial", "ticket", "dbias_accum"},
    L.OP_MSE: {"loss", "total", "dscore", "dpre", "dbias"},
    L.OP_COLSUM: {"out", "partial", "ticket", "bn_accum", "dgamma", "dbeta"},
    L.OP_ADAM: {"p", "m", "v", "step", "scalars"},
    L.OP_PACK: {"out", "out_bias"},
    L.OP_PACK_MULTI: {"_items[].out", "_items[].out_bias"},
    L.OP_BN_FOLD: {"scale", "shift"},
    L.OP_BN_FOLD_MULTI: {"_items[].scale", "_items[].shift"},
    L.OP_DP_ADAM: {"m", "v", "_p[T]"},          # (every rank's parameter range; reads every rank's gradient range)
}


class FakeStream:
    _ids = itertools.count(1)
    registry = {}

    def __init__(self, device=None, priority=0):
        self.id = next(FakeStream._ids)
        self.clock = {self.id: 0}
        FakeStream.registry[self.id] = self

    @property
    def cuda_stream(self):
        return self.id

    def wait_event(self, ev):
        if ev.clock is not None:
            for k, v in ev.clock.items():
                self.clock[k] = max(self.clock.get(k, 0), v)

    def wait_stream(self, other):
        for k, v in other.clock.items():
            self.clock[k] = max(self.clock.get(k, 0), v)

    def tick(self):
        self.clock[self.id] += 1
        return dict(self.clock)


class FakeEvent:
    def __init__(self, *a, **k):
        self.clock = None

    def record(self, stream=None):
        self.clock = dict((stream or CURRENT[-1]).clock)


CURRENT = []


@contextlib.contextmanager
def fake_stream_ctx(s):
    CURRENT.append(s)
    try:
        yield
    finally:
        CURRENT.pop()


def accesses(rec):
    """[(start, end, is_write, path)] over every tensor the record references."""
    out = []
    writes = WRITES[rec.kind]

    def walk(f, pre):
        for k, v in f.items():
            path = pre + k
            if isinstance(v, torch.Tensor):
                if v.numel():
                    w = path in writes
                    if rec.kind == L.OP_ADAM and k in ("step", "scalars") and rec.f.get("phase", 0) == 2:
                        w = False
                    if rec.kind == L.OP_ADAM and rec.f.get("phase", 0) == 1 and k in ("p", "m", "v", "g"):
                        continue                       # phase 1 only advances the step
                    if rec.kind == L.OP_L1 and k == "out" and rec.f.get("out_blc") is not None:
                        w = True                       # the loss reads the BLC tile and writes the NCL `out` itself
                    if rec.kind == L.OP_GEMM and k == "out" and rec.f.get("out_f32") == 2:
                        w = True
                    if rec.kind == L.OP_WGRAD and k == "partial" and rec.f.get("splits") == 1:
                        continue                       # split-free form (bf16 plans only): no workspace traffic
                    if rec.kind == L.OP_BN_BWD and k == "accum" and rec.f.get("defer"):
                        w = False                      # deferred: reads the sums; b2h_colsum(bn_accum) re-zeroes them
                    if rec.kind == L.OP_BN_BWD and rec.f.get("first_pass_only") and k in ("partial", "ticket", "sums"):
                        continue
                    if ((rec.kind == L.OP_BN_BWD and k == "accum" and rec.f.get("first_pass_only")) or
                            path == "bwd_sums.accum"):
                        w = 2                          # fp64 atomic adds: commute with each other, conflict with the rest
                    out.append((v.data_ptr(), v.data_ptr() + v.numel() * v.element_size(), w, path))
            elif isinstance(v, dict):
                walk(v, path + ".")
            elif isinstance(v, list) and v and isinstance(v[0], dict):
                for item in v:
                    walk(item, path + "[].")
            elif isinstance(v, list) and v and isinstance(v[0], torch.Tensor):
                for t in v:
                    out.append((t.data_ptr(), t.data_ptr() + t.numel() * t.element_size(), path + "[T]" in writes,
                                path + "[T]"))
    walk(rec.f, "")
    return out


def find_races(log):
    """log: [(op name, stream id, clock, accesses)] in enqueue order -> unordered conflicting pairs."""
    items = []
    for i, (_, _, _, acc) in enumerate(log):
        for (a, b, w, path) in acc:
            items.append((a, b, w, path, i))
    items.sort()
    races, active = [], []
    for a, b, w, path, i in items:
        active = [t for t in active if t[1] > a]
        for (a2, b2, w2, path2, j) in active:
            if i == j or not (w or w2) or (w == 2 and w2 == 2):
                continue
            first, second = (j, i) if j < i else (i, j)
            s1, c1 = log[first][1], log[first][2]
            c2 = log[second][2]
            if c2.get(s1, 0) < c1[s1]:
                races.append((log[first][0], path2 if first == j else path, log[second][0], path if first == j else path2))
        active.append((a, b, w, path, i))
    return sorted(set(races))


class _Log(list):
    pass


@pytest.fixture
def schedule_log(monkeypatch):
    log = _Log()
    base = FakeStream()
    CURRENT.clear()
    CURRENT.append(base)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: CURRENT[-1])
    monkeypatch.setattr(torch.cuda, "stream", fake_stream_ctx)

    def enqueue(self, first, end, stream):
        s = FakeStream.registry[stream] if stream is not None else CURRENT[-1]
        for rec in self.recs[first:end]:
            log.append((f"{rec.tag or rec.kind}@s{s.id}", s.id, s.tick(), accesses(rec)))

    def run(self, segment=None, stream=None):
        first, end = self.segments[segment] if segment is not None else (0, len(self.recs))
        enqueue(self, first, end, stream)

    def run_range(self, first, end, stream=None):
        enqueue(self, first, end, stream)

    def all_reduce(t, op=None, group=None, async_op=False):      # in place on the stream it is enqueued on
        cur = CURRENT[-1]
        log.append((f"all_reduce@s{cur.id}", cur.id, cur.tick(),
                    [(t.data_ptr(), t.data_ptr() + t.numel() * t.element_size(), True, "tensor")]))

    real_copy = torch.Tensor.copy_

    def copy_(dst, src, non_blocking=False):                     # batch feeding: tensor copies on the current stream
        cur = CURRENT[-1]
        acc = [(dst.data_ptr(), dst.data_ptr() + dst.numel() * dst.element_size(), True, "dst")]
        if isinstance(src, torch.Tensor) and src.numel():
            acc.append((src.data_ptr(), src.data_ptr() + src.numel() * src.element_size(), False, "src"))
        log.append((f"copy@s{cur.id}", cur.id, cur.tick(), acc))
        return real_copy(dst, src, non_blocking)

    log_copies = []
    monkeypatch.setattr(torch.Tensor, "copy_", lambda d, s_, non_blocking=False:
                        copy_(d, s_, non_blocking) if log_copies else real_copy(d, s_, non_blocking))
    log.log_copies = log_copies
    import torch.distributed as dist
    monkeypatch.setattr(dist, "all_reduce", all_reduce)
    monkeypatch.setattr(program_mod.Program, "run", run)
    monkeypatch.setattr(program_mod.Program, "run_range", run_range)
    yield log
    CURRENT.clear()


def _trainer(variant="v1", rf=False, precision="bf16", **kw):
    return GanTrainer(variant, 36, 252, rf, 8, 16, precision=precision, device="cpu", drop_mode="philox", **kw)


def test_detector_finds_a_planted_race(schedule_log):
    tr = _trainer()
    side = torch.cuda.Stream()
    tr.G_train.prog.run("fwd")                                  # base stream
    tr.G_train.prog.run("fwd", side.cuda_stream)                # the same buffers from an unordered stream
    assert find_races(schedule_log)
    del schedule_log[:]
    ev = torch.cuda.Event()
    tr.G_train.prog.run("fwd")
    ev.record(torch.cuda.current_stream())
    side.wait_event(ev)
    tr.G_train.prog.run("fwd", side.cuda_stream)
    assert find_races(schedule_log) == []


@pytest.mark.parametrize("variant,rf,precision", [("v1", False, "bf16"), ("v1", True, "bf16"), ("b2h", True, "bf16"),
                                                  ("v4", True, "bf16"), ("v2", True, "fp32"), ("v1", False, "fp32")])
def test_sequential_steps_have_no_unordered_conflicts(schedule_log, variant, rf, precision):
    tr = _trainer(variant, rf, precision)
    for _ in range(2):
        tr.G_train.pack()
        tr.D_train.pack()
        tr._g_step_body()
        tr._d_step_body()
    assert len(schedule_log) > 200
    assert len({e[1] for e in schedule_log}) >= 4               # dependency chain + wgrad / scoring / optimizer streams
    races = find_races(schedule_log)
    assert races == [], "\n".join(map(str, races[:20]))


@pytest.mark.parametrize("mode", ["0", "1", "2"])
@pytest.mark.parametrize("variant,rf", [("v1", False), ("v4", True)])
def test_deferred_bn_backward_and_first_pass_helpers_have_no_unordered_conflicts(schedule_log, monkeypatch, variant, rf,
                                                                                 mode):
    """Opt-in backward variants: the finishing column sums and the first-pass shares of the skip connections run on the
    weight-gradient streams; the producer's bn_bwd waits for its helpers' events."""
    monkeypatch.setenv("B2H_DEFER_BN", mode)
    monkeypatch.setenv("B2H_BWD_HELPERS", "1")
    monkeypatch.setenv("B2H_NO_GRAD_ADD", "1")     # (the default bf16 plans sum skip-connection gradients in the dgrad)
    tr = _trainer(variant, rf)
    assert any(r.f.get("first_pass_only") for r in tr.G_train.prog.recs if r.kind == L.OP_BN_BWD)
    assert any(r.f.get("bn_accum") is not None for r in tr.G_train.prog.recs if r.kind == L.OP_COLSUM) == (mode != "0")
    tr.G_train.pack()
    tr.D_train.pack()
    tr._g_step_body()
    for _ in range(2):
        tr._gan_ops(True)
    tr.flush_adv()
    races = find_races(schedule_log)
    assert races == [], "\n".join(map(str, races[:20]))


@pytest.mark.parametrize("lag_adv", [True, False])
@pytest.mark.parametrize("variant,rf", [("v1", False), ("v1", True), ("b2h", True), ("v2", True), ("v4", True),
                                        ("v4_deeper", True)])
def test_pipelined_gan_step_has_no_unordered_conflicts(schedule_log, variant, rf, lag_adv):
    tr = _trainer(variant, rf)
    tr.G_train.pack()
    tr.D_train.pack()
    tr._g_step_body()                                           # pipeline prologue: G0
    for _ in range(3):                                          # D_k || G_k+1, back to back
        tr._gan_ops(lag_adv)
    if lag_adv:
        tr.flush_adv()
    assert len({e[1] for e in schedule_log}) >= 6
    races = find_races(schedule_log)
    assert races == [], "\n".join(map(str, races[:20]))


@pytest.mark.parametrize("n_buckets", [1, 3])
@pytest.mark.parametrize("variant,rf", [("v1", False), ("v4", True)])
def test_data_parallel_steps_have_no_unordered_conflicts(schedule_log, variant, rf, n_buckets):
    """world_size 2: the gradient all-reduce of every bucket (comm stream) against the weight-gradient side streams
    that feed it and the optimizer ranges that consume it, sequential and pipelined schedules."""
    tr = _trainer(variant, rf, world_size=2, n_buckets=n_buckets)
    tr.G_train.pack()
    tr.D_train.pack()
    tr._g_step_body()
    tr._d_step_body()
    for _ in range(2):
        tr._gan_ops(True)
    tr.flush_adv()
    assert sum(1 for e in schedule_log if e[0].startswith("all_reduce")) == n_buckets * 2 * 3
    races = find_races(schedule_log)
    assert races == [], "\n".join(map(str, races[:20]))


@pytest.mark.parametrize("n_buckets", [1, 3])
def test_fused_exchange_schedule_has_no_unordered_conflicts(schedule_log, n_buckets):
    """fused_dp (b2h_dp_adam instead of all-reduce + Adam): one rank's schedule — the exchange kernel reads the local
    gradients that the side-stream weight gradients write and stores parameters that the repack reads."""
    from b2h_b200.trainer import PeerBuffers
    probe = _trainer()
    peers = {"g": PeerBuffers.in_process(probe.g_store.n, n_buckets, ["cpu"] * 2),
             "d": PeerBuffers.in_process(probe.d_store.n, n_buckets, ["cpu"] * 2)}
    del schedule_log[:]
    tr = _trainer(world_size=2, n_buckets=n_buckets, fused_dp=True, rank=0, peer_buffers={k: v[0] for k, v in peers.items()})
    tr.G_train.pack()
    tr.D_train.pack()
    tr._g_step_body()
    tr._d_step_body()
    for _ in range(2):
        tr._gan_ops(True)
    tr.flush_adv()
    assert sum(1 for e in schedule_log if e[0].startswith("dp_adam")) == n_buckets * 2 * 3
    assert not any(e[0].startswith("all_reduce") for e in schedule_log)
    # the exchange kernels of a step are chained in enqueue order (they spin on their peers)
    dp = [e for e in schedule_log if e[0].startswith("dp_adam")]
    for a, b in zip(dp, dp[1:]):                           # total order: no rank can run two of them the other way round
        assert b[2].get(a[1], 0) >= a[2][a[1]], (a[0], b[0])
    races = find_races(schedule_log)
    assert races == [], "\n".join(map(str, races[:20]))


@pytest.mark.parametrize("variant,rf", [("v1", False), ("b2h", True)])
def test_batch_feeding_has_no_unordered_conflicts(schedule_log, variant, rf):
    """bench.py's end-to-end loop: prefetch of batch k+1 (copy stream, staging buffers) under step k, swap into the
    static step inputs, the previous batch moving to the discriminator's inputs — copies included in the check."""
    tr = _trainer(variant, rf)
    tr.G_train.pack()
    tr.D_train.pack()
    hx, hy = torch.randn_like(tr.x), torch.randn_like(tr.y)
    hf = torch.randn_like(tr.feats) if tr.feats is not None else None
    schedule_log.log_copies.append(True)
    tr.load_batch(hx, hy, hf)
    tr._g_step_body()
    tr._sync_d_batch()
    tr.prefetch_batch(hx, hy, hf)
    for _ in range(3):
        tr.swap_batch(pipelined=True)
        tr.prefetch_batch(hx, hy, hf)
        tr._gan_ops(True)
    tr.flush_adv()
    assert sum(1 for e in schedule_log if e[0].startswith("copy")) >= 3 * 4
    races = find_races(schedule_log)
    assert races == [], "\n".join(map(str, races[:20]))
