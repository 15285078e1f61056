"""The kept entry points end to end on the device: `train_gan.py` on synthetic How2Sign-shaped clips (generator epochs
with validation + checkpoints, a discriminator epoch, resume with --use_checkpoint) and `inference.py` on its
checkpoint (train_gan.py:27-121 / inference.py:24-153 of the reference, end to end)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,extra", [("bf16", []), ("fp32", ["--require_text"]), ("bf16", ["--loss", "Huber1"])])
def test_train_then_infer(tmp_path, capsys, precision, extra):
    os.environ.setdefault("WANDB_MODE", "disabled")
    import inference
    import train_gan
    models = str(tmp_path / "models") + "/"
    common = ["--synthetic", "256", "--frames", "64", "--batch_size", "32", "--model_path", models, "--exp_name", "t1",
              "--precision", precision, "--epochs_train_disc", "2", "--log_step", "4", "--learning_rate", "1e-3"] + extra
    # (synthetic hands follow the arms, data.synthetic_r6d; the same 3 generator epochs through the CPU oracle take the
    # validation L1 from 5.98 to 5.13 at this learning rate)
    train_gan.main(train_gan.build_parser().parse_args(common + ["--num_epochs", "4"]))
    out = capsys.readouterr().out
    vals = [float(line.split("Val. Loss:")[1].split(",")[0]) for line in out.splitlines() if "Val. Loss:" in line]
    assert len(vals) == 3 and all(np.isfinite(vals)) and vals[-1] < vals[0]       # epochs 0, 1, 3 train the generator
    assert "Tr. Disc. Loss" in out
    for name in ("t1_checkpoint.pth", "discriminator_t1.pth", "lastCheckpoint_t1.pth", "t1arm2wh_preprocess_core.npz"):
        assert os.path.exists(os.path.join(models, name)), name
    ck = torch.load(os.path.join(models, "lastCheckpoint_t1.pth"), map_location="cpu")
    assert set(ck) == {"epoch", "state_dict", "g_optimizer"} and "decoder.9.weight" in ck["state_dict"]
    # resume
    train_gan.main(train_gan.build_parser().parse_args(common + ["--num_epochs", "2", "--use_checkpoint"]))
    out = capsys.readouterr().out
    resumed = [float(line.split("Val. Loss:")[1].split(",")[0]) for line in out.splitlines() if "Val. Loss:" in line]
    assert resumed and resumed[0] < vals[0]                                        # starts from the trained weights
    # inference on the checkpoint
    res = str(tmp_path / "results")
    inference.main(inference.build_parser().parse_args(
        ["--checkpoint", os.path.join(models, "lastCheckpoint_t1.pth"), "--model_path", models, "--exp_name", "t1",
         "--synthetic", "16", "--frames", "64", "--batch_size", "8", "--results_dir", res, "--precision", precision]
        + [e for e in extra if e.startswith("--require")]))
    r6d = np.load(os.path.join(res, "t1_r6d.npy"))
    assert r6d.shape == (16, 64, 252) and np.isfinite(r6d).all()
    assert np.isfinite(np.load(os.path.join(res, "t1_xyz.npy"))).all()


def test_inference_refuses_a_missing_checkpoint(tmp_path):
    import inference
    with pytest.raises(SystemExit):
        inference.main(inference.build_parser().parse_args(
            ["--checkpoint", str(tmp_path / "nope.pth"), "--synthetic", "8", "--frames", "64", "--results_dir",
             str(tmp_path)]))


def test_autograd_flow_matches_reference_step_order(tmp_path, capsys):
    """`--autograd`: the reference's literal flow through the drop-in modules (module forward under torch.autograd,
    torch losses, torch.optim.Adam), including the discriminator epoch with its two train-mode forwards before one
    backward."""
    os.environ.setdefault("WANDB_MODE", "disabled")
    import train_gan
    models = str(tmp_path / "models") + "/"
    train_gan.main(train_gan.build_parser().parse_args(
        ["--synthetic", "64", "--frames", "64", "--batch_size", "16", "--model_path", models, "--exp_name", "a1",
         "--precision", "fp32", "--epochs_train_disc", "2", "--num_epochs", "3", "--autograd"]))
    out = capsys.readouterr().out
    assert "Tr. Disc. Loss" in out and out.count("Val. Loss:") == 2
    d = float(out.split("Tr. Disc. Loss:")[1].split()[0])
    assert np.isfinite(d)
