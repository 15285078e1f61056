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
