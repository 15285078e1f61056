"""Host-side data path of the kept entry points: the reference's pipeline tables, window preparation and
standardisation (restated from utils/constants.py, utils/load_save_utils.py, utils/postprocess_utils.py,
utils/standardization_utils.py) plus the How2Sign-shaped synthetic generator."""
from __future__ import annotations

import os
import pickle
from typing import Optional

import numpy as np

# utils/constants.py:11-27
FEATURE_MAP = {"arm2wh": (6 * 6, 42 * 6), "arm_wh2wh": ((6 + 42) * 6, 42 * 6), "wh2wh": (42 * 6, 42 * 6)}
for _i, (_a, _b) in enumerate([(38, 4), (34, 8), (30, 12), (26, 16), (22, 20), (21, 21), (17, 25), (13, 29), (9, 33),
                               (5, 37), (1, 41), (0, 42)], start=1):
    FEATURE_MAP[f"arm_wh2finger{_i}"] = ((6 + _a) * 6, _b * 6)
# utils/constants.py:45-51
MODELS = {"v1": "regressor_fcn_bn_32", "b2h": "regressor_fcn_bn_32_b2h", "v2": "regressor_fcn_bn_32_v2",
          "v4": "regressor_fcn_bn_32_v4", "v4_deeper": "regressor_fcn_bn_32_v4_deeper"}
DATA_PATHS_r6d = {"train": "r6d_train.pkl", "val": "r6d_val.pkl", "test": "r6d_test.pkl"}
EPSILON = 1e-10


def make_equal_len(clips, maxpad: int = 192):
    """utils/postprocess_utils.py:50-51 ("cutting+reflect"): cut to maxpad frames or reflect-pad up to it."""
    return np.array([c[:maxpad] if c.shape[0] >= maxpad else np.pad(c, ((0, maxpad - c.shape[0]), (0, 0)), "reflect")
                     for c in clips])


def rmv_clips_nan(X, Y=None, F=None):
    """utils/postprocess_utils.py:5-28: drop every clip with a NaN in X, Y or its features."""
    bad = np.isnan(X).any(axis=(1, 2))
    if Y is not None:
        bad |= np.isnan(Y).any(axis=(1, 2))
    if F is not None:
        bad |= np.isnan(F).reshape(F.shape[0], -1).any(axis=1)
    keep = ~bad
    return X[keep], (Y[keep] if Y is not None else None), (F[keep] if F is not None else None)


def split_pipeline(data, pipeline):
    """utils/load_save_utils.py:44-50."""
    p0, p1 = FEATURE_MAP[pipeline]
    if pipeline in ("arm_wh2wh", "wh2wh"):
        return data, data[:, :, 6 * 6:]
    return data[:, :, :p0], data[:, :, p0:p0 + p1]


def mean_std(feat, data):
    """utils/standardization_utils.py:51-59 on (N, C, T) arrays (quirks kept: 'wh' std = std over clips of the
    per-clip temporal std; otherwise ONE global scalar std)."""
    mean = data.mean(axis=2).mean(axis=0)[np.newaxis, :, np.newaxis]
    if feat == "wh":
        std = data.std(axis=2).std(axis=0)[np.newaxis, :, np.newaxis] + EPSILON
    else:
        std = np.array([[[data.std()]]]).repeat(data.shape[1], axis=1)
    return mean, std


def calc_standard(train_X, train_Y, pipeline):
    """utils/standardization_utils.py:37-47."""
    in_feat, out_feat = pipeline.split("2")[0], pipeline.split("2")[1]
    mX, sX = mean_std(in_feat, train_X)
    if in_feat == out_feat:
        return mX, sX, mX, sX
    mY, sY = mean_std(out_feat, train_Y)
    return mX, sX, mY, sY


def synthetic_r6d(n_clips: int, T: int, seed: int = 23456):
    """(N, T, 288) 6-D rotations of 6 arm + 42 hand bones: random axis-angle per joint + temporal random walk,
    converted to the first two columns of the rotation matrix (np_mat_to_rot6d, conversion_utils.py:26)."""
    rng = np.random.RandomState(seed)
    aa = rng.randn(n_clips, 1, 48, 3) * 0.5 + np.cumsum(rng.randn(n_clips, T, 48, 3) * 0.05, axis=1)
    # the 42 hand bones follow the 6 arm bones through a fixed smooth map (+ 20 % of their own motion), so that
    # arm -> hand regression has something to learn (the entry-point tests check that training makes progress)
    mix = np.random.RandomState(4242).randn(18, 126) / np.sqrt(18.0)
    arm = aa[:, :, :6].reshape(n_clips, T, 18)
    aa[:, :, 6:] = 0.2 * aa[:, :, 6:] + np.tanh(arm @ mix).reshape(n_clips, T, 42, 3)
    th = np.linalg.norm(aa, axis=-1, keepdims=True) + 1e-12
    k = aa / th
    K = np.zeros(aa.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = np.sin(th)[..., None], np.cos(th)[..., None]
    R = np.eye(3) + s * K + (1 - c) * (K @ K)      # Rodrigues
    r6d = np.concatenate([R[..., 0], R[..., 1]], axis=-1)   # first two columns
    return r6d.reshape(n_clips, T, 288).astype(np.float32)


def synthetic_feats(kind: Optional[str], n: int, T: int, seed: int = 1):
    rng = np.random.RandomState(seed)
    if kind == "text":
        f = rng.randn(n, 512).astype(np.float32)
        return f / np.linalg.norm(f, axis=1, keepdims=True)
    if kind == "image":
        return (rng.randn(n, T, 2000) * 2).astype(np.float32)
    return None


def _load_pickle(path):
    with open(path, "rb") as fh:
        return pickle.load(fh)


def load_windows(data_path, pipeline, require_text=False, text_path=None, require_image=False, image_path=None):
    """utils/load_save_utils.py:37-58: the r6d pickle cut / reflect-padded to 192 frames and split into the pipeline's
    input / output joints, plus the sentence embeddings (N, 512) or the per-frame video features (N, T, 2000).
    Returns (X, Y, feats) with feats None for body-only models (the reference returns `(X, feats)` in X's place)."""
    X, Y = split_pipeline(make_equal_len(_load_pickle(data_path)), pipeline)
    feats = None
    if require_text and not require_image:
        feats = np.asarray(_load_pickle(text_path))
    elif require_image and not require_text:
        feats = make_equal_len(_load_pickle(image_path))
    return X, Y, feats


def feature_paths(data_dir, split, embeds_type="normal"):
    """train_gan.py:140-160 / inference.py:54-59: (text embeddings, video features) pickles of a split."""
    pre = "" if embeds_type == "normal" else "average_"
    return f"{data_dir}/{pre}{split}_sentence_embeddings.pkl", f"{data_dir}/{split}_vid_feats.pkl"


def load_train_val(args, rng, data_dir):
    """train_gan.py:129-205.  Returns (train_X, train_Y, val_X, val_Y, train_feats, val_feats), X/Y as
    standardised (N, C, T) float32; the statistics are written to {exp}{pipeline}_preprocess_core.npz."""
    kind = "text" if args.require_text else ("image" if args.require_image else None)
    sets = {}
    for name in ("train", "val"):
        if getattr(args, "synthetic", 0):
            n = args.synthetic if name == "train" else max(args.synthetic // 8, args.batch_size)
            X, Y = split_pipeline(synthetic_r6d(n, args.frames, seed=23456 + (name == "val")), args.pipeline)
            feats = synthetic_feats(kind, n, args.frames, seed=7 + (name == "val"))
        else:
            text_path, image_path = feature_paths(data_dir, name, args.embeds_type)
            X, Y, feats = load_windows(os.path.join(args.base_path, data_dir, DATA_PATHS_r6d[name]), args.pipeline,
                                       kind == "text", text_path, kind == "image", image_path)
        if args.pipeline == "wh2wh":
            X = X[:, :, 6 * 6:]
        X, Y, feats = rmv_clips_nan(X, Y, feats)
        sets[name] = (np.swapaxes(X, 1, 2).astype(np.float32), np.swapaxes(Y, 1, 2).astype(np.float32),
                      feats.astype(np.float32) if feats is not None else None)
    tX, tY, tF = sets["train"]
    vX, vY, vF = sets["val"]
    mX, sX, mY, sY = calc_standard(tX, tY, args.pipeline)
    if int(os.environ.get("RANK", "0")) == 0:      # data parallel: every rank computes the same statistics, one writes
        os.makedirs(args.model_path, exist_ok=True)
        np.savez_compressed(os.path.join(args.model_path, f"{args.exp_name}{args.pipeline}_preprocess_core.npz"),
                            body_mean_X=mX, body_std_X=sX, body_mean_Y=mY, body_std_Y=sY)
    tX, vX = ((tX - mX) / sX).astype(np.float32), ((vX - mX) / sX).astype(np.float32)
    tY, vY = ((tY - mY) / sY).astype(np.float32), ((vY - mY) / sY).astype(np.float32)
    I = np.arange(len(tX))
    rng.shuffle(I)
    return tX[I], tY[I], vX, vY, (tF[I] if tF is not None else None), vF
