"""Icosahedral rotation group tables (|G| = 60).

The reference ships three tables in utils/group_related/ and loads them from
cfg.SO3_related_files (test/estimator.py:78, network/group_feat.py:12-14,
network/rot_coh_match.py:127):
    Rotation.npy                       [60,3,3] f64   the rotation matrices, R[0] = I
    60_60.npy                          [60,60]        P[a][b] = index of (R_b . R_a)
    Nei_Index_in_SO3_ordered_13.npy    [60,13]        N[g][k] = P[g][N[0][k]]
`load()` reads that directory when given (drop-in use) and otherwise the packed copy
in roreg_b200/data/group/icosa60.npz (stand-alone tests / bench on a box without the reference).
"""
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class GroupTables:
    def __init__(self, rot, perm, nei):
        self.rot = np.ascontiguousarray(rot, dtype=np.float64)        # [60,3,3]
        self.perm = np.ascontiguousarray(perm, dtype=np.int32)        # [60,60]
        self.nei = np.ascontiguousarray(nei, dtype=np.int32)          # [60,13]
        assert self.rot.shape == (60, 3, 3) and self.perm.shape == (60, 60) and self.nei.shape == (60, 13)
        # inverse element: P[a][inv[a]] == 0  (R_inv[a] . R_a = I)
        self.inv = np.array([int(np.where(self.perm[a] == 0)[0][0]) for a in range(60)], dtype=np.int32)


def load(so3_dir=None):
    if so3_dir is not None and os.path.exists(os.path.join(so3_dir, "60_60.npy")):
        rot = np.load(os.path.join(so3_dir, "Rotation.npy"))
        perm = np.load(os.path.join(so3_dir, "60_60.npy"))
        nei = np.load(os.path.join(so3_dir, "Nei_Index_in_SO3_ordered_13.npy"))
        return GroupTables(rot, perm, nei)
    z = np.load(os.path.join(_HERE, "data", "group", "icosa60.npz"))
    return GroupTables(z["rot"], z["perm"], z["nei"])
