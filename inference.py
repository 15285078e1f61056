#!/usr/bin/env python
"""Kept entry point of the reference's `inference.py` (same flags) on the B200 path: load a checkpoint, run the
batched eval forward (inference.py:96-121), de-standardise (inference.py:134) and save the predicted 6-D
rotations plus their rotation matrices (libb2h `rot6d_to_mat`, utils/conversion_utils.py:86-107).

One process per GPU (`torchrun --nproc-per-node N inference.py ...` shards the clips; no collective is needed),
replacing the reference's nn.DataParallel wrap (inference.py:45-47).  The axis-angle / xyz forward kinematics of
`save_results` (utils/utils.py:388-427) run on the GPU too (libb2h `b2h_fk`, one launch); GIF rendering is out of
scope.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import modelZoo  # noqa: E402
from b2h_b200 import _lib as L  # noqa: E402
from b2h_b200 import data as b2h_data  # noqa: E402


def rot6d_to_mat(r6d: torch.Tensor) -> torch.Tensor:
    """(..., 6) fp32 CUDA tensor -> (..., 9) rotation matrices through the C ABI."""
    flat = r6d.reshape(-1, 6).contiguous().float()
    out = torch.empty(flat.shape[0], 9, device=flat.device, dtype=torch.float32)
    L.run_oneshot(L.Rot6d(r6d=flat.data_ptr(), mat=out.data_ptr(), n=flat.shape[0]), L.F32,
                  torch.cuda.current_stream(flat.device).cuda_stream)
    return out.reshape(*r6d.shape[:-1], 9)


def prepare_inputs(args, rank: int = 0, world: int = 1):
    """inference.py:50-87 of the reference: load the windows of the inference split (+ text / video features), drop
    clips with NaNs, standardise with the statistics train_gan.py saved next to the checkpoint.  Returns
    (X, Yn, feats, input_feats, (mX, sX, mY, sY)) with X / Yn standardised (N, C, T) float32, this rank's clips only."""
    kind = "text" if args.require_text else ("image" if args.require_image else None)
    if args.synthetic:
        X, Y = b2h_data.split_pipeline(b2h_data.synthetic_r6d(args.synthetic, args.frames, seed=99), args.pipeline)
        feats = b2h_data.synthetic_feats(kind, args.synthetic, args.frames, seed=98)
    else:                                                            # inference.py:52-60
        text_path, image_path = b2h_data.feature_paths(args.data_dir, args.infer_set, args.embeds_type)
        X, Y, feats = b2h_data.load_windows(f"{args.data_dir}/r6d_{args.infer_set}.pkl", args.pipeline,
                                            kind == "text", text_path, kind == "image", image_path)
    X, Y, feats = b2h_data.rmv_clips_nan(X, Y, feats)
    input_feats = X                       # arms + hands as loaded: the FK post-processing needs all 48 bones
    if args.pipeline == "wh2wh":          # inference.py:72-73
        X = X[:, :, 6 * 6:]
    X, Y = np.swapaxes(X, 1, 2).astype(np.float32), np.swapaxes(Y, 1, 2).astype(np.float32)
    feats = feats.astype(np.float32) if feats is not None else None
    # the statistics train_gan.py saved next to the checkpoint (inference.py:79-87)
    stats_name = f"{args.exp_name}{args.pipeline}_preprocess_core.npz"
    stats_path = next((p for p in (os.path.join(os.path.split(args.checkpoint)[0], stats_name),
                                   os.path.join(args.model_path, stats_name)) if os.path.exists(p)), None)
    if stats_path is not None:
        c = np.load(stats_path)
        mX, sX, mY, sY = c["body_mean_X"], c["body_std_X"], c["body_mean_Y"], c["body_std_Y"]
    elif args.synthetic:
        mX, sX, mY, sY = b2h_data.calc_standard(X, Y, args.pipeline)
    else:
        raise SystemExit(f"{stats_name} not found next to the checkpoint or under --model_path")
    X = ((X - mX) / sX).astype(np.float32)
    Yn = ((Y - mY) / sY).astype(np.float32)
    X, Yn, input_feats = X[rank::world], Yn[rank::world], input_feats[rank::world]
    feats = feats[rank::world] if feats is not None else None
    return X, Yn, feats, input_feats, (mX, sX, mY, sY)


def main(args):
    if not torch.cuda.is_available():
        raise SystemExit("inference.py (B200 build) needs a CUDA device: there is no CPU fallback")
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(device)
    cin, cout = b2h_data.FEATURE_MAP[args.pipeline]
    mod = b2h_data.MODELS[args.model]
    model = getattr(modelZoo, mod)()
    if mod == "regressor_fcn_bn_32_b2h":
        model.build_net(cin, cout, require_image=args.require_image)
    else:
        model.build_net(cin, cout, require_text=args.require_text)
    model.precision = args.precision
    if args.random_weights:
        print("WARNING: --random_weights: no checkpoint is loaded, the predictions are meaningless", flush=True)
    else:
        if not os.path.exists(args.checkpoint):   # the reference's torch.load fails on a missing file as well
            raise SystemExit(f"checkpoint {args.checkpoint!r} not found (pass --checkpoint, or --random_weights to "
                             "run the pipeline on an untrained model)")
        st = torch.load(args.checkpoint, map_location="cpu")
        model.load_state_dict(st["state_dict"], strict=False)      # inference.py:41-43
    model.to(device).eval()
    X, Yn, feats, input_feats, (mX, sX, mY, sY) = prepare_inputs(args, rank, world)
    outs, err, steps = [], 0.0, 0
    bs = args.batch_size
    with torch.no_grad():
        for s in range(0, min(X.shape[0], args.num_samples), bs):    # inference.py:96-105 (last batch may be short)
            x = torch.from_numpy(X[s:s + bs]).to(device)
            f = torch.from_numpy(feats[s:s + bs]).to(device) if feats is not None else None
            out = model(x, feats_=f)
            err += torch.nn.functional.l1_loss(out, torch.from_numpy(Yn[s:s + bs]).to(device)).item() * bs
            steps += 1
            outs.append(out)
    out = torch.cat(outs, 0)
    print(f">>> TOTAL ERROR: {err / max(steps * bs, 1)}", flush=True)
    sY_d, mY_d = (torch.from_numpy(np.asarray(a, dtype=np.float32)).to(device) for a in (sY, mY))
    pred = out * sY_d + mY_d                                                         # inference.py:134
    r6d = pred.permute(0, 2, 1).contiguous()                                         # (N, T, 6*J)
    mats = rot6d_to_mat(r6d.reshape(r6d.shape[0], r6d.shape[1], -1, 6))
    os.makedirs(args.results_dir, exist_ok=True)
    tag = f"{args.exp_name}_rank{rank}" if world > 1 else args.exp_name
    np.save(os.path.join(args.results_dir, f"{tag}_r6d.npy"), r6d.cpu().numpy())
    np.save(os.path.join(args.results_dir, f"{tag}_rotmat.npy"), mats.cpu().numpy())
    print(f"saved {tuple(r6d.shape)} r6d and {tuple(mats.shape)} rotation matrices to {args.results_dir}", flush=True)
    # save_results (utils/utils.py:388-427): input + prediction -> axis-angle -> xyz over the 49-bone skeleton, with
    # the root bone / bone lengths the reference pickles next to the data (utils/utils.py:412-419)
    inp = input_feats[:out.shape[0]]
    if args.pipeline in ("arm_wh2wh", "wh2wh"):
        inp = inp[:, :, :6 * 6]                                      # keep arms (utils/utils.py:396-397)
    if inp.shape[2] + cout == 48 * 6:
        from b2h_b200 import postprocess as PP
        root_p, bone_p = os.path.join(args.base_path, "root.pkl"), os.path.join(args.base_path, "bone_len.pkl")
        if os.path.exists(root_p) and os.path.exists(bone_p):
            root, bone = b2h_data._load_pickle(root_p), b2h_data._load_pickle(bone_p)
        elif args.synthetic:
            bone = np.array([250, 190, 300, 260, 190, 300, 260] + 2 * ([35] + 5 * [40, 30, 25, 20]), dtype=np.float32)
            root = np.array([0, 0, 0, 0, bone[0], 0], dtype=np.float32)
        else:
            root = None
        if root is not None:
            n = out.shape[0]
            frames = torch.cat([torch.from_numpy(np.ascontiguousarray(inp, dtype=np.float32)).to(device), r6d],
                               dim=2).reshape(-1, 288).contiguous()   # np.concatenate((input, output), axis=2)
            xyz = PP.r6d_to_xyz(frames, root, bone).reshape(n, -1, 150)
            np.save(os.path.join(args.results_dir, f"{tag}_xyz.npy"), xyz.cpu().numpy())
            print(f"saved {tuple(xyz.shape)} joint positions (b2h_fk)", flush=True)


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--checkpoint", type=str, default="models/lastCheckpoint.pth")   # inference.py:157
    p.add_argument("--base_path", type=str, default="./")
    p.add_argument("--data_dir", type=str, default="video_data")
    p.add_argument("--pipeline", type=str, default="arm2wh")
    p.add_argument("--require_text", action="store_true")
    p.add_argument("--require_image", action="store_true")
    p.add_argument("--embeds_type", type=str, default="normal")
    p.add_argument("--tag", type=str, default="")
    p.add_argument("--exp_name", type=str, default="experiment")
    p.add_argument("--model_path", type=str, default="models/")
    p.add_argument("--model", type=str, default="v1")
    p.add_argument("--infer_set", type=str, default="test")
    p.add_argument("--batch_size", type=int, default=128)
    p.add_argument("--seqs_to_viz", type=int, default=2, help="accepted for compatibility: GIF rendering is out of scope")
    p.add_argument("--num_samples", type=int, default=3000)
    p.add_argument("--results_dir", type=str, default="results/")
    p.add_argument("--precision", type=str, default="fp32", choices=["fp32", "bf16"])
    p.add_argument("--synthetic", type=int, default=0)
    p.add_argument("--frames", type=int, default=192)
    p.add_argument("--random_weights", action="store_true", help="smoke runs only: skip the checkpoint")
    return p


if __name__ == "__main__":
    main(build_parser().parse_args())
