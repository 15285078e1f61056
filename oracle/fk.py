"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's pose post-processing, the path behind
`save_results` (utils/utils.py:388-427): 6-D rotations -> rotation matrix -> axis-angle -> forward kinematics over
the 49-bone skeleton -> joint positions, and the MPJPE between two predictions (SURVEY 8c "MPJPE").

Follows, function by function:
  rot6d_to_mat   utils/conversion_utils.py:86-107  (np_rot6d_to_mat, applied per row as the reference calls it, S10)
  rot6d_to_aa    utils/conversion_utils.py:33-41   (_rot6d_to_aa: scipy Rotation.from_matrix(...).as_rotvec())
  aa_to_xyz      utils/conversion_utils.py:111-137 (_retrieve_axis_angle + Rodrigues step per bone)
  SKELETON       3DposeEstimator/skeletalModel.py:42-118 (bone i: joint J[i] -> joint i+1, reference joint B[i])

Pinned against the reference's own functions in tests/test_fk_vs_reference.py (run where /root/reference exists)
and against tests/golden/fk_golden.npz (made by tools/make_golden.py from the reference) everywhere else.
Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation

# (J, B) per bone; the end joint of bone i is i + 1.  Bone 0 is the root bone (head), given by `root`.
SKEL_J = [0, 1, 2, 3, 1, 5, 6, 4] + [j for f in range(5) for j in (8, 9 + 4 * f, 10 + 4 * f, 11 + 4 * f)] + \
         [7] + [j for f in range(5) for j in (29, 30 + 4 * f, 31 + 4 * f, 32 + 4 * f)]
SKEL_B = [-1, 0, 1, 2, 0, 1, 5, 3] + [j for f in range(5) for j in (4, 8, 9 + 4 * f, 10 + 4 * f)] + \
         [6] + [j for f in range(5) for j in (7, 29, 30 + 4 * f, 31 + 4 * f)]
assert len(SKEL_J) == len(SKEL_B) == 49
N_BONES = 49
HAND_JOINTS = list(range(8, 50))   # the 42 hand joints (end joints of bones 7..48)


def rot6d_to_mat(r6d: np.ndarray) -> np.ndarray:
    """(n, 6) -> (n, 9), row-major [x y z] as COLUMNS; the +1e-6 regularisers of the reference kept."""
    r6d = np.asarray(r6d, dtype=np.float64).reshape(-1, 6)
    x_raw, y_raw = r6d[:, 0:3], r6d[:, 3:6]
    x = x_raw / (np.linalg.norm(x_raw, axis=1, keepdims=True) + 1e-6)
    z = np.cross(x, y_raw)
    z = z / (np.linalg.norm(z, axis=1, keepdims=True) + 1e-6)
    y = np.cross(z, x)
    return np.stack([x, y, z], axis=-1).reshape(-1, 9)


def rot6d_to_aa(r6d: np.ndarray) -> np.ndarray:
    """(n, 6) -> (n, 3) rotation vectors (scipy's log map, as the reference)."""
    return Rotation.from_matrix(rot6d_to_mat(r6d).reshape(-1, 3, 3)).as_rotvec()


def frames_to_aa(r6d_frames: np.ndarray) -> np.ndarray:
    """(n, 6*J) -> (n, 3*J), joint by joint (clip_rot6d_to_aa, conversion_utils.py:44-48)."""
    n, c = r6d_frames.shape
    return rot6d_to_aa(r6d_frames.reshape(n * (c // 6), 6)).reshape(n, (c // 6) * 3)


def aa_to_xyz(aa: np.ndarray, root: np.ndarray, bone_len: np.ndarray) -> np.ndarray:
    """(n, 3*48) axis-angles -> (n, 150) joint positions.  Zero-angle joints leave the direction unchanged (the
    reference divides by the zero norm there; no finite input of the path reaches that case)."""
    aa = np.asarray(aa, dtype=np.float64)
    n = aa.shape[0]
    xyz = np.zeros((n, (N_BONES + 1) * 3))
    xyz[:, 0:6] = np.asarray(root, dtype=np.float64).reshape(1, 6)
    for i in range(1, N_BONES):
        pj, pb = xyz[:, SKEL_J[i] * 3: SKEL_J[i] * 3 + 3], xyz[:, SKEL_B[i] * 3: SKEL_B[i] * 3 + 3]
        u = pj - pb
        u = u / np.linalg.norm(u, axis=1, keepdims=True)
        v_aa = aa[:, (i - 1) * 3: (i - 1) * 3 + 3]
        th = np.linalg.norm(v_aa, axis=1, keepdims=True)
        a = np.divide(v_aa, th, out=np.zeros_like(v_aa), where=th > 0)
        v = u * np.cos(th) + np.cross(a, u) * np.sin(th) + a * (a * u).sum(1, keepdims=True) * (1 - np.cos(th))
        xyz[:, (i + 1) * 3: (i + 1) * 3 + 3] = pj + bone_len[i] * v
    return xyz


def r6d_to_xyz(r6d_frames: np.ndarray, root: np.ndarray, bone_len: np.ndarray) -> np.ndarray:
    """(n, 288) de-standardised 6-D rotations of the 48 joints (6 arm + 42 hand) -> (n, 150)."""
    return aa_to_xyz(frames_to_aa(r6d_frames), root, bone_len)


def synthetic_skeleton_mm(seed: int = 7):
    """Fixed root bone and anthropometric bone lengths in millimetres (the reference's xyz units are
    sigma-normalised image units, so the MPJPE bar of the task is defined on this fixed synthetic skeleton)."""
    rng = np.random.RandomState(seed)
    bone = np.zeros(N_BONES)
    bone[0:7] = [250, 190, 300, 260, 190, 300, 260]       # head/neck, shoulders, upper arms, forearms
    for base in (7, 28):                                   # right / left hand
        bone[base] = 35                                    # wrist stub
        for f in range(5):
            bone[base + 1 + 4 * f: base + 5 + 4 * f] = np.array([[40, 32, 28, 22], [68, 40, 25, 20], [65, 45, 28, 21],
                                                                  [60, 41, 27, 21], [55, 32, 20, 18]][f])
    bone = bone * (1.0 + 0.02 * rng.randn(N_BONES))
    root = np.array([0.0, 0.0, 0.0, 0.0, bone[0], 0.0])   # joint 0 at the origin, joint 1 one head-bone below
    return root, bone


def mpjpe(xyz_a: np.ndarray, xyz_b: np.ndarray, joints=HAND_JOINTS) -> float:
    """Mean per-joint position error over the given joints (same units as the skeleton)."""
    a = np.asarray(xyz_a, dtype=np.float64).reshape(xyz_a.shape[0], -1, 3)[:, joints]
    b = np.asarray(xyz_b, dtype=np.float64).reshape(xyz_b.shape[0], -1, 3)[:, joints]
    return float(np.linalg.norm(a - b, axis=2).mean())
