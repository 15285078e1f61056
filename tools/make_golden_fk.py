#!/usr/bin/env python
"""Generate tests/golden/fk.npz by running the REAL reference (/root/reference, read-only): clip_rot6d_to_aa
(utils/conversion_utils.py:44-48) and aa_to_xyz (:117-137) over getSkeletalModelStructure()
(3DposeEstimator/skeletalModel.py) on procedural 6-D rotations.

Run in the authoring container only:  python tools/make_golden_fk.py
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fk  # noqa: E402  (only for the fixed synthetic skeleton: root + bone lengths)

REF = "/root/reference"


def load_ref(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    conv = load_ref("utils/conversion_utils.py", "_ref_conv")
    skel = load_ref("3DposeEstimator/skeletalModel.py", "_ref_skel")
    structure = skel.getSkeletalModelStructure()
    rng = np.random.RandomState(11)
    # two clips of 8 frames: unit-ish 6-D rotations with a temporal drift, 48 joints
    base = rng.randn(2, 1, 288)
    r6d = (base + 0.1 * np.cumsum(rng.randn(2, 8, 288), axis=1)).astype(np.float64)
    root, bone_len = fk.synthetic_skeleton_mm()
    aa = [conv.clip_rot6d_to_aa(c) for c in r6d]
    xyz = conv.aa_to_xyz(np.array(aa), root, bone_len, structure)
    out = os.path.join(ROOT, "tests", "golden", "fk.npz")
    np.savez_compressed(out, r6d=r6d, aa=np.array(aa), xyz=np.array(xyz), root=root, bone_len=bone_len,
                        J=np.array([t[0] for t in structure]), B=np.array([t[3] for t in structure]))
    print("wrote", out, np.array(xyz).shape)


if __name__ == "__main__":
    main()
