#!/usr/bin/env python
"""Kept entry point of the reference's `train_gan.py` (same flags, same epoch schedule, same checkpoint files)
running on the B200 path: `modelZoo` classes backed by libb2h.so.

Two execution modes for the per-batch step bodies (train_gan.py:215-299):
  * default ("fused"): the `GanTrainer` programs -- forward, losses, backward and Adam are all libb2h kernels
    replayed from CUDA graphs; the modules' parameters ARE the trainer's flat buffers, so checkpoints,
    `state_dict()` and the validation loop see the updates;
  * `--autograd`: the reference's literal flow -- module forward through torch.autograd, torch L1/MSE,
    `torch.optim.Adam` -- useful to cross-check the fused path.

Data: the How2Sign pickles the reference loads (utils/load_save_utils.py:37-58) when `--base_path/--data_dir`
hold them, or `--synthetic N` How2Sign-shaped clips when they do not.
Launch one process per GPU with torchrun for data parallel training (gradients all-reduced over NCCL).
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import modelZoo  # noqa: E402
from b2h_b200 import data as b2h_data  # noqa: E402
from b2h_b200.trainer import FlatAdam, GanTrainer  # noqa: E402

# utils/constants.py:11-27,45-51
FEATURE_MAP = b2h_data.FEATURE_MAP
MODELS = b2h_data.MODELS
device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
lastCheckpoint = ""

try:  # W&B needs the network; everything works without it
    import wandb
    os.environ.setdefault("WANDB_MODE", "disabled")
except Exception:  # pragma: no cover
    wandb = None


def _log(d):
    if wandb is not None and wandb.run is not None:
        wandb.log(d)


class FrozenAdaptiveLoss(nn.Module):
    """`--loss RobustLoss` as the reference evaluates it (train_gan.py:74-77,286-290,336-339): the
    AdaptiveLossFunction's alpha / scale are never given to an optimiser, so they stay at their initial values
    alpha = 2, scale = 1/2 (utils/robust_loss/adaptive.py:55-59) and the loss is mean(2 d^2) + log(1/2) + log sqrt(2 pi)
    (the fused path computes exactly this in b2h_l1, B2H_LOSS_ROBUST)."""

    def forward(self, out, gt):
        return 2.0 * torch.mean((out - gt) ** 2) + (np.log(0.5) + 0.5 * np.log(2.0 * np.pi))


# utils/constants.py:53-58
LOSSES = {"L1": nn.L1Loss(), "L2": nn.MSELoss(), "Huber1": nn.HuberLoss(delta=1.0), "RobustLoss": FrozenAdaptiveLoss()}


def calc_motion(tensor):  # train_gan.py:209-211
    return tensor[:, :, :1] - tensor[:, :, :-1]


def load_data(args, rng, data_dir="video_data"):
    """train_gan.py:129-205: load, standardise (stats saved to *_preprocess_core.npz), shuffle."""
    return b2h_data.load_train_val(args, rng, data_dir)


def _batches(n, bs):
    return np.arange(n // bs)   # integer division: the last incomplete batch is dropped (train_gan.py:218)


def _to_dev(a):
    if torch.is_tensor(a):          # device-resident dataset (fused path): a slice is already where it is needed
        return a.to(device)
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


class LossMeter:
    """`avgLoss += loss.item() * batch_size` (train_gan.py:251,299) without a host synchronisation per step: the
    running sum lives on the device in float64 and is read only where the reference prints it."""

    def __init__(self, dev):
        self.acc = torch.zeros((), dtype=torch.float64, device=dev)

    def add(self, loss, weight):
        self.acc += loss.detach().double().reshape(()) * weight

    def value(self):
        return float(self.acc)


def train_discriminator(args, generator, discriminator, gan_criterion, d_optimizer, train_X, train_Y, epoch,
                        train_feats=None, trainer: GanTrainer = None):
    generator.eval()
    discriminator.train()
    batchinds = _batches(train_X.shape[0], args.batch_size)
    meter = LossMeter(device)
    for bi in batchinds:
        s = bi * args.batch_size
        x, y = _to_dev(train_X[s:s + args.batch_size]), _to_dev(train_Y[s:s + args.batch_size])
        f = _to_dev(train_feats[s:s + args.batch_size]) if train_feats is not None else None
        if trainer is not None:
            trainer.load_batch(x, y, f)
            trainer.discriminator_step(graph=True)
            d_loss = trainer.losses[3]
        else:
            with torch.no_grad():
                fake = generator(x, feats_=f).detach()
            fake_score = discriminator(calc_motion(fake))
            real_score = discriminator(calc_motion(y))
            tf, tr = (0.1, 0.9) if args.disc_label_smooth else (0.0, 1.0)
            loss = gan_criterion(fake_score, torch.full_like(fake_score, tf)) + \
                gan_criterion(real_score, torch.full_like(real_score, tr))
            d_optimizer.zero_grad()
            loss.backward()
            d_optimizer.step()
            d_loss = loss
        meter.add(d_loss, args.batch_size)
    avg = meter.value()
    n = max(len(batchinds) * args.batch_size, 1)
    print(f"Epoch [{epoch}/{args.num_epochs - 1}], Tr. Disc. Loss: {avg / n}", flush=True)
    _log({"epoch": epoch, "loss_train_disc": avg / n})


def train_generator(args, generator, discriminator, reg_criterion, gan_criterion, g_optimizer, train_X, train_Y, epoch,
                    clip_grad=False, train_feats=None, trainer: GanTrainer = None):
    discriminator.eval()
    generator.train()
    batchinds = _batches(train_X.shape[0], args.batch_size)
    total = len(batchinds)
    meter = LossMeter(device)
    for bii, bi in enumerate(batchinds):
        s = bi * args.batch_size
        x, y = _to_dev(train_X[s:s + args.batch_size]), _to_dev(train_Y[s:s + args.batch_size])
        f = _to_dev(train_feats[s:s + args.batch_size]) if train_feats is not None else None
        if trainer is not None:
            trainer.load_batch(x, y, f)
            trainer.generator_step(graph=True)
            g_loss = trainer.losses[2]
        else:
            out = generator(x, feats_=f)
            with torch.no_grad():
                fake_score = discriminator(calc_motion(out))
            loss = reg_criterion(out, y) + gan_criterion(fake_score.detach(), torch.ones_like(fake_score))
            g_optimizer.zero_grad()
            loss.backward()
            if clip_grad:
                torch.nn.utils.clip_grad_norm_(generator.parameters(), 1)
            g_optimizer.step()
            g_loss = loss
        meter.add(g_loss, args.batch_size)
        if bii % args.log_step == 0:
            m = meter.value() / (total * args.batch_size)
            print("Epoch [{}/{}], Step [{}/{}], Tr. Loss: {:.4f}, Tr. Perplexity: {:5.4f}".format(
                epoch, args.num_epochs - 1, bii + 1, total, m, np.exp(m)), flush=True)
    m = meter.value() / max(total * args.batch_size, 1)
    print("Epoch [{}/{}], Tr. Loss: {:.4f}, Tr. Perplexity: {:5.4f}".format(epoch, args.num_epochs - 1, m, np.exp(m)),
          flush=True)
    _log({"epoch": epoch, "loss_train_gen": m})


def val_generator(args, generator, discriminator, reg_criterion, g_optimizer, d_optimizer, val_X, val_Y, currBestLoss,
                  prev_save_epoch, epoch, val_feats=None):
    """train_gan.py:312-372: eval-mode L1 on the validation set (batch = batch_size // 2), checkpoint on improvement."""
    global lastCheckpoint
    generator.eval()
    discriminator.eval()
    vbs = max(args.batch_size // 2, 1)
    batchinds = _batches(val_X.shape[0], vbs)
    test_loss = 0.0
    with torch.no_grad():
        for bi in batchinds:
            s = bi * vbs
            x, y = _to_dev(val_X[s:s + vbs]), _to_dev(val_Y[s:s + vbs])
            f = _to_dev(val_feats[s:s + vbs]) if val_feats is not None else None
            test_loss += reg_criterion(generator(x, feats_=f), y).item() * vbs
    test_loss /= max(len(batchinds) * vbs, 1)
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:
        # one decision for all ranks: every rank holds the same weights but its own BN running statistics (local
        # batch statistics), so the validation losses differ slightly -- rank 0's value decides the
        # checkpoint AND the early stop everywhere (a rank leaving the epoch loop alone would hang the others in the
        # next gradient all-reduce)
        import torch.distributed as dist
        t = torch.tensor([test_loss], dtype=torch.float64, device=device)
        dist.broadcast(t, 0)
        test_loss = float(t.item())
    _log({"loss_val_gen": test_loss})
    print("Epoch [{}/{}], Val. Loss: {:.4f}, Val. Perplexity: {:5.4f}".format(epoch, args.num_epochs - 1, test_loss,
                                                                               np.exp(test_loss)), flush=True)
    if test_loss < currBestLoss:
        prev_save_epoch = epoch                         # on every rank (train_gan.py:350-352)
        currBestLoss = test_loss
        if rank == 0:                                   # only the files are rank 0's business
            os.makedirs(args.model_path, exist_ok=True)
            fileName = os.path.join(args.model_path, f"{args.exp_name}_checkpoint.pth")
            torch.save({"epoch": epoch, "state_dict": generator.state_dict(), "g_optimizer": g_optimizer.state_dict()},
                       fileName)
            lastCheckpoint = fileName
            torch.save({"epoch": epoch, "state_dict": discriminator.state_dict(),
                        "d_optimizer": d_optimizer.state_dict()},
                       os.path.join(args.model_path, f"discriminator_{args.exp_name}.pth"))
    return currBestLoss, prev_save_epoch


def main(args):
    if not torch.cuda.is_available():
        raise SystemExit("train_gan.py (B200 build) needs a CUDA device: there is no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    global device
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        pg = dist.group.WORLD
    if wandb is not None:
        wandb.init(project="B2H-H2S", name=args.exp_name, mode=os.environ.get("WANDB_MODE", "disabled"))
    feature_in_dim, feature_out_dim = FEATURE_MAP[args.pipeline]
    rng = np.random.RandomState(23456)
    torch.manual_seed(23456)
    torch.cuda.manual_seed(23456)
    data = load_data(args, rng, args.data_dir)
    train_X, train_Y, val_X, val_Y, train_feats, val_feats = data
    if world > 1:   # shard clips across ranks (one process per GPU); every rank gets the same number of clips, so
        # every rank runs the same number of steps (each step holds a collective)
        sl = slice(rank, (train_X.shape[0] // world) * world, world)
        train_X, train_Y = train_X[sl], train_Y[sl]
        train_feats = train_feats[sl] if train_feats is not None else None
    mod = MODELS[args.model]
    generator = getattr(modelZoo, mod)()
    if mod == "regressor_fcn_bn_32_b2h":
        generator.build_net(feature_in_dim, feature_out_dim, require_image=args.require_image)
    else:
        generator.build_net(feature_in_dim, feature_out_dim, require_text=args.require_text)
    generator.precision = args.precision
    generator.to(device)
    discriminator = modelZoo.regressor_fcn_bn_discriminator()
    discriminator.build_net(feature_out_dim)
    discriminator.precision = args.precision
    discriminator.to(device)
    if args.loss not in LOSSES:
        raise SystemExit(f"--loss must be one of {sorted(LOSSES)}")
    reg_criterion, gan_criterion = LOSSES[args.loss], nn.MSELoss()
    trainer = None
    if args.autograd and world > 1:
        raise SystemExit("--autograd is the single-process cross-check of the reference's literal flow: it has no "
                         "gradient exchange.  Data-parallel training runs through the fused trainer (drop --autograd)")
    if args.autograd:
        g_optimizer = torch.optim.Adam(generator.parameters(), lr=args.learning_rate, weight_decay=0)
        d_optimizer = torch.optim.Adam(discriminator.parameters(), lr=args.learning_rate, weight_decay=0)
    else:
        T = train_X.shape[2]
        trainer = GanTrainer.from_modules(generator, discriminator, batch_size=args.batch_size, T=T,
                                          precision=args.precision, lr=args.learning_rate,
                                          label_smooth=args.disc_label_smooth, world_size=world, process_group=pg,
                                          loss=args.loss, seed=23456 + rank)   # own dropout stream per rank
        g_optimizer, d_optimizer = trainer.g_opt, trainer.d_opt
    if args.use_checkpoint:
        st = torch.load(os.path.join(args.model_path, f"lastCheckpoint_{args.exp_name}.pth"), map_location="cpu")
        generator.load_state_dict(st["state_dict"], strict=False)
        g_optimizer.load_state_dict(st["g_optimizer"])
        st = torch.load(os.path.join(args.model_path, f"discriminator_{args.exp_name}.pth"), map_location="cpu")
        discriminator.load_state_dict(st["state_dict"], strict=False)
        d_optimizer.load_state_dict(st["d_optimizer"])
    if trainer is not None and not args.host_data:
        # the training set stays on the GPU (How2Sign r6d at T = 192 is ~7 GB of fp32); a batch is
        # a device-to-device copy into the trainer's static buffers and the epoch shuffle is a device gather with
        # the reference's permutation
        train_X, train_Y = torch.from_numpy(train_X).to(device), torch.from_numpy(train_Y).to(device)
        train_feats = torch.from_numpy(train_feats).to(device) if train_feats is not None else None
    currBestLoss, prev_save_epoch = 1e9, 0
    for epoch in range(args.num_epochs):
        args.epoch = epoch
        if epoch > 100 and (epoch - prev_save_epoch) > args.patience:
            print("early stopping at:", epoch - 1, flush=True)
            break
        if epoch > 0 and (args.epochs_train_disc == 0 or epoch % args.epochs_train_disc == 0):   # train_gan.py:108
            train_discriminator(args, generator, discriminator, gan_criterion, d_optimizer, train_X, train_Y, epoch,
                                train_feats=train_feats, trainer=trainer)
        else:
            train_generator(args, generator, discriminator, reg_criterion, gan_criterion, g_optimizer, train_X,
                            train_Y, epoch, train_feats=train_feats, trainer=trainer)
            currBestLoss, prev_save_epoch = val_generator(args, generator, discriminator, reg_criterion, g_optimizer,
                                                          d_optimizer, val_X, val_Y, currBestLoss, prev_save_epoch,
                                                          epoch, val_feats=val_feats)
        I = np.arange(len(train_X))
        rng.shuffle(I)                                        # train_gan.py:113-118, same generator, same order
        if torch.is_tensor(train_X):
            I = torch.from_numpy(I).to(train_X.device)
        train_X, train_Y = train_X[I], train_Y[I]
        if train_feats is not None:
            train_feats = train_feats[I]
    if lastCheckpoint and rank == 0:
        shutil.copyfile(lastCheckpoint, os.path.join(args.model_path, f"lastCheckpoint_{args.exp_name}.pth"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.synchronize()
        if trainer is not None:
            trainer.release_graphs()      # captured NCCL kernels hold the communicator: drop them first
        dist.barrier()
        dist.destroy_process_group()


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--base_path", type=str, default="./")
    p.add_argument("--pipeline", type=str, default="arm2wh")
    p.add_argument("--num_epochs", type=int, default=200)
    p.add_argument("--batch_size", type=int, default=128)
    p.add_argument("--learning_rate", type=float, default=1e-4)
    p.add_argument("--require_text", action="store_true")
    p.add_argument("--require_image", action="store_true")
    p.add_argument("--embeds_type", type=str, default="normal")
    p.add_argument("--model_path", type=str, default="models/")
    p.add_argument("--log_step", type=int, default=25)
    p.add_argument("--tag", type=str, default="")
    p.add_argument("--exp_name", type=str, default="experiment")
    p.add_argument("--patience", type=int, default=100)
    p.add_argument("--use_checkpoint", action="store_true")
    p.add_argument("--epochs_train_disc", type=int, default=3)
    p.add_argument("--model", type=str, default="v1")
    p.add_argument("--disc_label_smooth", action="store_true")
    p.add_argument("--data_dir", type=str, default="video_data")
    p.add_argument("--loss", type=str, default="L1")
    # additions (defaults keep the reference's behaviour)
    p.add_argument("--precision", type=str, default="fp32", choices=["fp32", "bf16"])
    p.add_argument("--autograd", action="store_true", help="literal reference flow: torch autograd + torch.optim.Adam")
    p.add_argument("--host_data", action="store_true", help="keep the training set in host memory (default: on the GPU)")
    p.add_argument("--synthetic", type=int, default=0, help="use N synthetic How2Sign-shaped clips instead of the pickles")
    p.add_argument("--frames", type=int, default=192, help="window length of the synthetic clips (reference: 192)")
    return p


if __name__ == "__main__":
    args = build_parser().parse_args()
    print(args, flush=True)
    main(args)
