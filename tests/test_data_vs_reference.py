"""Host data path (b2h_b200/data.py) against the reference's utils (where /root/reference is mounted)."""
import sys

import numpy as np
import pytest

import b2h_b200  # noqa: F401
from b2h_b200 import data

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref_utils():
    sys.path.insert(0, "/root/reference/utils")
    try:
        import constants
        import postprocess_utils
        import standardization_utils
    finally:
        sys.path.pop(0)
    return constants, postprocess_utils, standardization_utils


def test_tables(ref_utils):
    constants = ref_utils[0]
    assert data.FEATURE_MAP == {k: tuple(v) for k, v in constants.FEATURE_MAP.items()}
    assert data.MODELS == constants.MODELS
    assert data.DATA_PATHS_r6d == constants.DATA_PATHS_r6d


@pytest.mark.parametrize("pipeline", ["arm2wh", "wh2wh", "arm_wh2finger3"])
def test_calc_standard(ref_utils, pipeline):
    std = ref_utils[2]
    rng = np.random.RandomState(0)
    cin, cout = data.FEATURE_MAP[pipeline]
    X = rng.randn(7, cin, 20).astype(np.float32) * 3 + 1
    Y = rng.randn(7, cout, 20).astype(np.float32) * 0.5 - 2
    for a, b in zip(std.calc_standard(X, Y, pipeline), data.calc_standard(X, Y, pipeline)):
        np.testing.assert_array_equal(a, b)


def test_windows_and_nan_removal(ref_utils):
    pp = ref_utils[1]
    rng = np.random.RandomState(1)
    clips = [rng.randn(n, 12) for n in (250, 100, 192, 30)]
    np.testing.assert_array_equal(pp.make_equal_len(clips, method="cutting+reflect"), data.make_equal_len(clips))
    X = rng.randn(6, 10, 4)
    Y = rng.randn(6, 10, 5)
    F = rng.randn(6, 8)
    X[1, 2, 3] = np.nan
    Y[4, 0, 0] = np.nan
    F[5, 7] = np.nan
    for a, b in zip(pp.rmv_clips_nan(X.copy(), Y.copy(), F.copy()), data.rmv_clips_nan(X, Y, F)):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("pipeline,kind", [("arm2wh", None), ("arm_wh2wh", "text"), ("wh2wh", "image"),
                                           ("arm_wh2finger7", None)])
def test_load_windows(ref_utils, tmp_path, pipeline, kind):
    """data.load_windows against utils/load_save_utils.py:37-58 on pickles of ragged clips (+ embeddings / video features)."""
    import pickle
    sys.path.insert(0, "/root/reference/utils")
    try:
        import load_save_utils
    finally:
        sys.path.pop(0)
    rng = np.random.RandomState(2)
    lens = (250, 100, 192, 40)
    r6d, txt, img = tmp_path / "r6d_test.pkl", tmp_path / "txt.pkl", tmp_path / "img.pkl"
    pickle.dump([rng.randn(n, 288).astype(np.float32) for n in lens], open(r6d, "wb"))
    pickle.dump(rng.randn(len(lens), 512).astype(np.float32), open(txt, "wb"))
    pickle.dump([rng.randn(n, 2000).astype(np.float32) for n in lens], open(img, "wb"))
    ref = load_save_utils.load_windows(str(r6d), pipeline, require_text=kind == "text", text_path=str(txt),
                                       require_image=kind == "image", image_path=str(img))
    X, Y, F = data.load_windows(str(r6d), pipeline, kind == "text", str(txt), kind == "image", str(img))
    ref_X, ref_F = ref[0] if kind else (ref[0], None)
    np.testing.assert_array_equal(ref_X, X)
    np.testing.assert_array_equal(ref[1], Y)
    if kind:
        np.testing.assert_array_equal(ref_F, F)
    else:
        assert F is None
    assert X.shape == (len(lens), 192, 288 if pipeline in ("arm_wh2wh", "wh2wh") else data.FEATURE_MAP[pipeline][0])


@pytest.mark.parametrize("pipeline,kind", [("arm2wh", None), ("wh2wh", "text"), ("arm_wh2finger9", "image")])
def test_inference_prepare_inputs(ref_utils, tmp_path, pipeline, kind):
    """inference.prepare_inputs (the kept inference.py, lines 50-87 of the reference) on real-format files: pickles of
    ragged clips with a NaN clip, embeddings / video features, the *_preprocess_core.npz next to the checkpoint —
    against the reference's own load_windows / rmv_clips_nan / standardisation lines; and the rank sharding."""
    import os
    import pickle
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import inference
    sys.path.insert(0, "/root/reference/utils")
    try:
        import load_save_utils
    finally:
        sys.path.pop(0)
    pp = ref_utils[1]
    rng = np.random.RandomState(3)
    lens = (250, 100, 192, 40, 300)
    clips = [rng.randn(n, 288).astype(np.float32) for n in lens]
    clips[1][5, 7] = np.nan
    data_dir, ckpt_dir = tmp_path / "video_data", tmp_path / "models"
    data_dir.mkdir()
    ckpt_dir.mkdir()
    pickle.dump(clips, open(data_dir / "r6d_val.pkl", "wb"))
    pickle.dump(rng.randn(len(lens), 512).astype(np.float32), open(data_dir / "val_sentence_embeddings.pkl", "wb"))
    pickle.dump([rng.randn(n, 2000).astype(np.float32) for n in lens], open(data_dir / "val_vid_feats.pkl", "wb"))
    cin, cout = data.FEATURE_MAP[pipeline]
    stats = dict(body_mean_X=rng.randn(1, cin, 1).astype(np.float32), body_std_X=rng.rand(1, cin, 1).astype(np.float32) + 0.5,
                 body_mean_Y=rng.randn(1, cout, 1).astype(np.float32), body_std_Y=rng.rand(1, cout, 1).astype(np.float32) + 0.5)
    np.savez_compressed(ckpt_dir / f"exp7{pipeline}_preprocess_core.npz", **stats)
    argv = ["--checkpoint", str(ckpt_dir / "lastCheckpoint_exp7.pth"), "--data_dir", str(data_dir), "--pipeline", pipeline,
            "--exp_name", "exp7", "--infer_set", "val", "--model_path", str(tmp_path / "nowhere")]
    argv += {"text": ["--require_text"], "image": ["--require_image"], None: []}[kind]
    args = inference.build_parser().parse_args(argv)
    X, Yn, F, input_feats, st = inference.prepare_inputs(args)
    # the reference's lines
    tX, tY = load_save_utils.load_windows(str(data_dir / "r6d_val.pkl"), pipeline, require_text=kind == "text",
                                          text_path=str(data_dir / "val_sentence_embeddings.pkl"),
                                          require_image=kind == "image", image_path=str(data_dir / "val_vid_feats.pkl"))
    tF = None
    if kind:
        tX, tF = tX
    tX, tY, tF = pp.rmv_clips_nan(tX, tY, tF)
    ref_input_feats = tX.copy()
    if pipeline == "wh2wh":
        tX = tX[:, :, 6 * 6:]
    tX = np.swapaxes(tX, 1, 2).astype(np.float32)
    tY = np.swapaxes(tY, 1, 2).astype(np.float32)
    tX = (tX - stats["body_mean_X"]) / stats["body_std_X"]
    tY = (tY - stats["body_mean_Y"]) / stats["body_std_Y"]
    assert X.shape == (4, cin, 192)
    np.testing.assert_array_equal(X, tX.astype(np.float32))
    np.testing.assert_array_equal(Yn, tY.astype(np.float32))
    np.testing.assert_array_equal(input_feats, ref_input_feats)
    if kind:
        np.testing.assert_array_equal(F, np.asarray(tF, dtype=np.float32))
    else:
        assert F is None
    # two ranks partition the clips
    parts = [inference.prepare_inputs(args, r, 2) for r in range(2)]
    assert sorted(np.concatenate([p[0][:, 0, 0] for p in parts]).tolist()) == sorted(X[:, 0, 0].tolist())
    assert all(p[0].shape[0] == p[3].shape[0] == 2 for p in parts)
    # statistics are required for real data
    args.checkpoint = str(tmp_path / "elsewhere" / "x.pth")
    with pytest.raises(SystemExit):
        inference.prepare_inputs(args)


@pytest.mark.parametrize("pipeline,kind", [("arm2wh", None), ("wh2wh", None), ("arm2wh", "text"), ("arm_wh2finger4", "image")])
def test_load_train_val_matches_reference_load_data(tmp_path, pipeline, kind):
    """data.load_train_val (behind the kept train_gan.load_data) against the REAL train_gan.load_data
    (train_gan.py:127-205) on real-format pickles: standardised train / val arrays, the shuffle drawn from the same
    RandomState, the feature arrays and the saved *_preprocess_core.npz."""
    import argparse
    import importlib.util
    import os
    import pickle
    os.environ.setdefault("WANDB_MODE", "disabled")
    saved = list(sys.path)
    sys.path[:0] = ["/root/reference", "/root/reference/utils"]
    try:
        spec = importlib.util.spec_from_file_location("_ref_train_gan", "/root/reference/train_gan.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        sys.path[:] = saved
    rng = np.random.RandomState(4)
    data_dir = tmp_path / "video_data"
    data_dir.mkdir()
    for split, lens in (("train", (250, 100, 192, 40, 300, 192, 191)), ("val", (200, 64, 192))):
        clips = [rng.randn(n, 288).astype(np.float32) * 0.7 + 0.1 for n in lens]
        if split == "train":
            clips[3][2, 100] = np.nan
        pickle.dump(clips, open(data_dir / f"r6d_{split}.pkl", "wb"))
        pickle.dump(rng.randn(len(lens), 512).astype(np.float32), open(data_dir / f"{split}_sentence_embeddings.pkl", "wb"))
        pickle.dump([rng.randn(n, 2000).astype(np.float32) for n in lens], open(data_dir / f"{split}_vid_feats.pkl", "wb"))
    outs = []
    for which in ("ref", "ours"):
        args = argparse.Namespace(base_path="", pipeline=pipeline, require_text=kind == "text", require_image=kind == "image",
                                  embeds_type="normal", model_path=str(tmp_path / f"models_{which}") + "/", exp_name="e1",
                                  synthetic=0, batch_size=2, frames=192)
        r = np.random.RandomState(23456)
        res = ref.load_data(args, r, str(data_dir)) if which == "ref" else data.load_train_val(args, r, str(data_dir))
        outs.append((res, dict(np.load(os.path.join(args.model_path, f"e1{pipeline}_preprocess_core.npz"))), r.rand()))
    (a, sa, ra), (b, sb, rb) = outs
    assert ra == rb                                   # the generators were advanced identically
    for k in sa:
        np.testing.assert_array_equal(sa[k], sb[k])
    for i in range(4):
        assert a[i].dtype == b[i].dtype == np.float32
        np.testing.assert_array_equal(a[i], b[i])
    if kind:
        np.testing.assert_array_equal(np.asarray(a[4], dtype=np.float32), b[4])
        np.testing.assert_array_equal(np.asarray(a[5], dtype=np.float32), b[5])
    else:
        assert len(a) == 4 and b[4] is None and b[5] is None
