"""Host logic of the kept train_gan.py loops (train_gan.py:215-308 of the reference) on CPU with a stub trainer: batch
slicing (incomplete last batch dropped), device-resident tensors or numpy arrays, the loss average the reference
prints (sum of loss * batch_size over steps / (steps * batch_size)) without a per-step host read, and the epoch
shuffle with the reference's RandomState order."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("WANDB_MODE", "disabled")
import train_gan  # noqa: E402


class StubTrainer:
    def __init__(self):
        self.losses = torch.zeros(8)
        self.batches = []

    def load_batch(self, x, y, f=None):
        self.batches.append((x.clone(), y.clone(), None if f is None else f.clone()))

    def generator_step(self, graph=False):
        x, y, _ = self.batches[-1]
        self.losses[2] = (x.mean() - y.mean()).abs() + 1.0

    def discriminator_step(self, graph=False):
        x, y, _ = self.batches[-1]
        self.losses[3] = (x.sum() * 1e-3).abs() + 0.5


def _args():
    return argparse.Namespace(batch_size=4, num_epochs=3, log_step=2, disc_label_smooth=False)


def _data(n=14):
    rng = np.random.RandomState(0)
    return (rng.randn(n, 6, 5).astype(np.float32), rng.randn(n, 7, 5).astype(np.float32),
            rng.randn(n, 3).astype(np.float32))


def test_generator_and_discriminator_loops(capsys):
    X, Y, F = _data()
    mod = torch.nn.Identity()
    for resident in (False, True):
        tx, ty, tf = (torch.from_numpy(a) for a in (X, Y, F)) if resident else (X, Y, F)
        tr = StubTrainer()
        train_gan.train_generator(_args(), mod, mod, None, None, None, tx, ty, 1, train_feats=tf, trainer=tr)
        out = capsys.readouterr().out
        assert len(tr.batches) == 3                                   # 14 // 4: the incomplete batch is dropped
        expect = 0.0
        for i, (x, y, f) in enumerate(tr.batches):
            np.testing.assert_array_equal(x.numpy(), X[4 * i:4 * i + 4])
            np.testing.assert_array_equal(y.numpy(), Y[4 * i:4 * i + 4])
            np.testing.assert_array_equal(f.numpy(), F[4 * i:4 * i + 4])
            expect += float(abs(X[4 * i:4 * i + 4].mean() - Y[4 * i:4 * i + 4].mean()) + 1.0) * 4
        expect /= 3 * 4
        assert "Epoch [1/2], Tr. Loss: {:.4f}".format(expect) in out
        assert "Step [1/3]" in out and "Step [3/3]" in out and "Step [2/3]" not in out    # log_step = 2
        tr = StubTrainer()
        train_gan.train_discriminator(_args(), mod, mod, None, None, tx, ty, 2, train_feats=tf, trainer=tr)
        out = capsys.readouterr().out
        expect = sum(float(abs(X[4 * i:4 * i + 4].sum() * 1e-3) + 0.5) * 4 for i in range(3)) / 12
        assert f"Tr. Disc. Loss: {expect}"[:30] in out


def test_loss_meter_matches_item_accumulation():
    vals = torch.rand(50) * 3
    m = train_gan.LossMeter(torch.device("cpu"))
    ref = 0.0
    for v in vals:
        m.add(v, 128)
        ref += v.item() * 128
    assert abs(m.value() - ref) <= 1e-12 * ref


def test_epoch_shuffle_order_is_the_references():
    X, _, _ = _data(11)
    rng_a, rng_b = np.random.RandomState(23456), np.random.RandomState(23456)
    a, b = X.copy(), torch.from_numpy(X.copy())
    for _ in range(3):
        I = np.arange(len(a))
        rng_a.shuffle(I)
        a = a[I]
        J = np.arange(len(b))
        rng_b.shuffle(J)
        b = b[torch.from_numpy(J)]
        np.testing.assert_array_equal(a, b.numpy())


class _Gen(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Conv1d(6, 7, 1)

    def forward(self, x, audio_=None, percent_rand_=0.7, feats_=None):
        return self.lin(x) + (0.0 if feats_ is None else feats_.mean())


def test_val_generator_loss_and_checkpoints(tmp_path):
    """train_gan.py:312-372: batches of batch_size // 2 (incomplete one dropped), mean of loss * vbs, checkpoint files
    and their dictionary keys on improvement only."""
    X, Y, F = _data(11)
    torch.manual_seed(0)
    gen, disc = _Gen(), torch.nn.Conv1d(7, 1, 1)
    g_opt, d_opt = torch.optim.Adam(gen.parameters()), torch.optim.Adam(disc.parameters())
    args = argparse.Namespace(batch_size=6, num_epochs=5, model_path=str(tmp_path / "m"), exp_name="e9")
    crit = torch.nn.L1Loss()
    best, saved = train_gan.val_generator(args, gen, disc, crit, g_opt, d_opt, X, Y, 1e9, 0, 3, val_feats=F)
    with torch.no_grad():
        expect = sum(crit(gen(torch.from_numpy(X[3 * i:3 * i + 3]), feats_=torch.from_numpy(F[3 * i:3 * i + 3])),
                          torch.from_numpy(Y[3 * i:3 * i + 3])).item() * 3 for i in range(3)) / 9
    assert abs(best - expect) < 1e-7 and saved == 3
    ck = torch.load(tmp_path / "m" / "e9_checkpoint.pth")
    assert set(ck) == {"epoch", "state_dict", "g_optimizer"} and ck["epoch"] == 3
    assert set(ck["state_dict"]) == set(gen.state_dict())
    dk = torch.load(tmp_path / "m" / "discriminator_e9.pth")
    assert set(dk) == {"epoch", "state_dict", "d_optimizer"}
    assert train_gan.lastCheckpoint == str(tmp_path / "m" / "e9_checkpoint.pth")
    # no improvement: nothing is rewritten, the best loss and its epoch stay
    os.remove(tmp_path / "m" / "e9_checkpoint.pth")
    best2, saved2 = train_gan.val_generator(args, gen, disc, crit, g_opt, d_opt, X, Y, best - 1e-3, 3, 4, val_feats=F)
    assert best2 == best - 1e-3 and saved2 == 3 and not os.path.exists(tmp_path / "m" / "e9_checkpoint.pth")


def test_frozen_adaptive_loss_is_the_oracle_formula():
    from oracle import ref_models as R
    g = torch.Generator().manual_seed(0)
    o, t = torch.randn(3, 7, 5, generator=g), torch.randn(3, 7, 5, generator=g)
    assert abs(float(train_gan.LOSSES["RobustLoss"](o, t)) - float(R.reg_criterion("RobustLoss", o, t))) < 1e-6
    for k in ("L1", "L2", "Huber1"):
        assert abs(float(train_gan.LOSSES[k](o, t)) - float(R.reg_criterion(k, o, t))) < 1e-7


def _val_worker(rank, port, model_path, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2", RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=2)
    train_gan.device = torch.device("cpu")
    X, Y, F = _data(11)
    torch.manual_seed(0)
    gen, disc = _Gen(), torch.nn.Conv1d(7, 1, 1)
    with torch.no_grad():
        gen.lin.bias += 0.05 * rank          # ranks see different validation losses (their own BN statistics)
    g_opt, d_opt = torch.optim.Adam(gen.parameters()), torch.optim.Adam(disc.parameters())
    args = argparse.Namespace(batch_size=6, num_epochs=300, model_path=model_path, exp_name=f"r{rank}")
    crit = torch.nn.L1Loss()
    best, saved, hist = 1e9, 0, []
    for epoch in (0, 1, 2):
        with torch.no_grad():
            gen.lin.weight *= 0.5 if epoch < 2 else 4.0    # improves twice, then gets worse
        best, saved = train_gan.val_generator(args, gen, disc, crit, g_opt, d_opt, X, Y, best, saved, epoch, val_feats=F)
        hist.append((best, saved))
    ret[rank] = hist
    dist.barrier()
    dist.destroy_process_group()


def test_val_generator_decides_once_for_all_ranks(tmp_path):
    """Data parallel: the best loss and the epoch of the last improvement -- the inputs of the early-stopping
    test train_gan.py:105-107 -- must be identical on every rank (rank 0's validation loss decides), or one rank leaves
    the epoch loop alone and the others hang in the next gradient all-reduce; only rank 0 writes checkpoints."""
    import torch.multiprocessing as mp
    from tests.test_ddp_gloo import _free_port
    ret = mp.Manager().dict()
    mp.spawn(_val_worker, args=(_free_port(), str(tmp_path), ret), nprocs=2, join=True)
    assert ret[0] == ret[1]
    assert [s for _, s in ret[0]] == [0, 1, 1]
    assert os.path.exists(tmp_path / "r0_checkpoint.pth") and not os.path.exists(tmp_path / "r1_checkpoint.pth")
