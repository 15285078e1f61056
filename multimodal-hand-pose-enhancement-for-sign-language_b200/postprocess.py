"""Inference post-processing on the GPU: the path behind the reference's `save_results`
(utils/utils.py:388-427): de-standardise -> rot6d -> axis-angle (`rot6d_to_aa`, utils/conversion_utils.py:33-56)
-> forward kinematics (`aa_to_xyz`, utils/conversion_utils.py:117-137), one launch of libb2h's `b2h_fk` instead of a
`Pool(24)` of per-row SciPy calls.  No CPU fallback: raises if the library or a B200 is missing."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from .program import _fill_struct

# 3DposeEstimator/skeletalModel.py:42-118 as (J, B) per bone; bone i ends at joint i + 1, bone 0 is the root bone
SKEL_J = [0, 1, 2, 3, 1, 5, 6, 4] + [j for f in range(5) for j in (8, 9 + 4 * f, 10 + 4 * f, 11 + 4 * f)] + \
         [7] + [j for f in range(5) for j in (29, 30 + 4 * f, 31 + 4 * f, 32 + 4 * f)]
SKEL_B = [-1, 0, 1, 2, 0, 1, 5, 3] + [j for f in range(5) for j in (4, 8, 9 + 4 * f, 10 + 4 * f)] + \
         [6] + [j for f in range(5) for j in (7, 29, 30 + 4 * f, 31 + 4 * f)]


def structure_arrays(structure=None) -> Tuple[list, list]:
    """(J, B) lists from the reference's `getSkeletalModelStructure()` tuples (J, E, L, B), or the built-in ones."""
    if structure is None:
        return list(SKEL_J), list(SKEL_B)
    assert all(t[1] == i + 1 for i, t in enumerate(structure)), "bone i must end at joint i + 1"
    return [int(t[0]) for t in structure], [int(t[3]) for t in structure]


def r6d_to_xyz(r6d, root, bone_len, structure=None, mean=None, std=None, return_aa: bool = False):
    """r6d: (n, 6*(nbones-1)) tensor/array of per-frame 6-D rotations (the reference's (T, C) clip layout, clips
    stacked along n).  Returns xyz (n, 3*(nbones+1)) [and the axis-angles (n, 3*(nbones-1))] as CUDA tensors."""
    L.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    J, Bf = structure_arrays(structure)
    nb = len(J)
    t = r6d if torch.is_tensor(r6d) else torch.from_numpy(np.ascontiguousarray(np.asarray(r6d, dtype=np.float32)))
    t = t.to(dev, torch.float32).contiguous()
    assert t.dim() == 2 and t.shape[1] >= (nb - 1) * 6, t.shape
    n = t.shape[0]
    xyz = torch.empty(n, (nb + 1) * 3, dtype=torch.float32, device=dev)
    aa = torch.empty(n, (nb - 1) * 3, dtype=torch.float32, device=dev) if return_aa else None
    m = s = None
    if mean is not None:
        m = torch.from_numpy(np.ascontiguousarray(np.asarray(mean, dtype=np.float32).reshape(-1))).to(dev)
        s = torch.from_numpy(np.ascontiguousarray(np.asarray(std, dtype=np.float32).reshape(-1))).to(dev)
        assert m.numel() >= (nb - 1) * 6 and s.numel() >= (nb - 1) * 6
    bl = [float(x) for x in np.asarray(bone_len).reshape(-1)]
    assert len(bl) == nb
    pad = L.FK_MAX_BONES - nb
    desc = _fill_struct(L.Fk(), dict(r6d=t, ld=t.shape[1], mean=m, std=s, aa=aa, xyz=xyz, n=n, nbones=nb,
                                     joint=J + [0] * pad, before=[max(b, 0) for b in Bf] + [0] * pad,
                                     bone_len=bl + [0.0] * pad,
                                     root=[float(x) for x in np.asarray(root).reshape(-1)[:6]]))
    L.run_oneshot(desc, L.F32, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    return (xyz, aa) if return_aa else xyz


def rot6d_to_aa(r6d_clips: Sequence) -> list:
    """Drop-in for utils/conversion_utils.py:51-56: list of (T, 6*J) clips -> list of (T, 3*J) numpy arrays."""
    out = []
    for clip in r6d_clips:
        c = np.asarray(clip, dtype=np.float32)
        nj = c.shape[1] // 6
        if nj + 1 > L.FK_MAX_BONES or nj < 1:
            raise ValueError(f"rot6d_to_aa: {nj} joints per frame is outside [1, {L.FK_MAX_BONES - 1}]")
        # a chain skeleton: only the axis-angles are wanted here
        st = [(i, i + 1, i, i - 1) for i in range(nj + 1)]
        _, aa = r6d_to_xyz(c, np.array([0, 0, 0, 0, 1, 0], dtype=np.float32), np.ones(nj + 1), structure=st,
                           return_aa=True)
        out.append(aa.cpu().numpy().astype(np.float64))
    return out


def mpjpe(xyz_a: torch.Tensor, xyz_b: torch.Tensor, joints: Optional[Sequence[int]] = None) -> float:
    """Mean per-joint position error between two (n, 3*J) predictions (over `joints`, default the 42 hand joints)."""
    joints = list(range(8, 50)) if joints is None else list(joints)
    a = xyz_a.reshape(xyz_a.shape[0], -1, 3)[:, joints].double()
    b = xyz_b.reshape(xyz_b.shape[0], -1, 3)[:, joints].double()
    return float((a - b).norm(dim=2).mean())
