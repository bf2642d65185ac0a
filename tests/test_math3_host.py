"""Host build of roreg_b200/csrc/math3.cuh (the exact source the RANSAC / refine kernels compile)
against NumPy/LAPACK: polar factor U V^T, proper 3-point Kabsch, quaternion x anchor product."""
import ctypes
import os
import subprocess
import numpy as np
from oracle import roreg_oracle as O

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
SRC = os.path.join(REPO, "roreg_b200", "csrc", "math3_host.cpp")
SO = os.path.join(REPO, "roreg_b200", "csrc", "libmath3_host.so")
dp = ctypes.POINTER(ctypes.c_double); fp = ctypes.POINTER(ctypes.c_float)


def _lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC])
    return ctypes.CDLL(SO)


def test_polar_factor_matches_lapack():
    L = _lib(); rng = np.random.default_rng(0)
    for i in range(500):
        H = rng.standard_normal((3, 3)) * 10 ** rng.uniform(-3, 3)
        U, S, VT = np.linalg.svd(H)
        R = np.zeros((3, 3)); L.rr_host_polar_uvt(H.ctypes.data_as(dp), R.ctypes.data_as(dp))
        if S[2] / S[0] > 1e-8:
            assert np.abs(R - U @ VT).max() < 1e-7 / (S[2] / S[0]) * 1e-6 + 1e-12


def test_three_point_transform_is_proper_branch_of_reference():
    L = _lib(); rng = np.random.default_rng(1)
    for i in range(500):
        k1 = rng.random((3, 3)) * 3
        Q = np.linalg.qr(rng.standard_normal((3, 3)))[0]
        if np.linalg.det(Q) < 0: Q[:, 0] *= -1
        k0 = np.ascontiguousarray(k1 @ Q.T + rng.standard_normal(3) + rng.standard_normal((3, 3)) * 0.01)
        T = np.zeros((3, 4)); L.rr_host_three_point_transform(k0.ctypes.data_as(dp), k1.ctypes.data_as(dp), T.ctypes.data_as(dp))
        Tr = O.threepps2tran(k0, k1)
        assert abs(np.linalg.det(T[:, :3]) - 1) < 1e-9
        if np.linalg.det(Tr[:, :3]) > 0:
            assert np.abs(T - Tr).max() < 1e-9
        assert np.abs(k1 @ T[:, :3].T + T[:, 3] - (k1 @ Tr[:, :3].T + Tr[:, 3])).max() < 1e-8


def test_three_point_transform_arbitrary_triplets():
    """Non-rigid triplets (wrong correspondences, repeated points): sigma_3 of the cross-covariance is pure
    rounding noise of any relative size; the proper-rotation branch must still equal the reference's result
    whenever LAPACK's coin flip lands on det = +1."""
    L = _lib(); rng = np.random.default_rng(4)
    nproper = 0
    for i in range(3000):
        k0 = np.ascontiguousarray(rng.random((3, 3)) * 3 + rng.uniform(-5, 5, 3))
        k1 = np.ascontiguousarray(rng.random((3, 3)) * 3)
        T = np.zeros((3, 4)); L.rr_host_three_point_transform(k0.ctypes.data_as(dp), k1.ctypes.data_as(dp), T.ctypes.data_as(dp))
        Tr = O.threepps2tran(k0, k1)
        assert abs(np.linalg.det(T[:, :3]) - 1) < 1e-9 and np.abs(T[:, :3] @ T[:, :3].T - np.eye(3)).max() < 1e-9
        if np.linalg.det(Tr[:, :3]) > 0:
            nproper += 1
            assert np.abs(T - Tr).max() < 1e-8
    assert nproper > 1000


def test_three_point_transform_repeated_points_is_still_a_rotation():
    """Triplets drawn WITH replacement (test/estimator.py:228) repeat a match: the cross-covariance has rank 1 (two equal
    points) or 0 (three equal points).  LAPACK returns an orthogonal matrix for the reference; the device arithmetic must also
    return a proper rotation (round 1 returned a rank-1 projector / the zero matrix) that maps the distinct points correctly."""
    L = _lib(); rng = np.random.default_rng(9)
    for i in range(300):
        k1 = rng.random((3, 3)) * 3; k0 = rng.random((3, 3)) * 3 + rng.uniform(-2, 2, 3)
        kind = i % 3
        if kind == 0: k1[2] = k1[1]; k0[2] = k0[1]                     # rank 1
        elif kind == 1: k1[1] = k1[0]; k1[2] = k1[0]; k0[1] = k0[0]; k0[2] = k0[0]      # rank 0
        else: k1[2] = k1[0]; k0[2] = k0[0]
        k0 = np.ascontiguousarray(k0); k1 = np.ascontiguousarray(k1)
        T = np.zeros((3, 4)); L.rr_host_three_point_transform(k0.ctypes.data_as(dp), k1.ctypes.data_as(dp), T.ctypes.data_as(dp))
        R = T[:, :3]
        assert np.isfinite(T).all() and np.abs(R @ R.T - np.eye(3)).max() < 1e-9 and abs(np.linalg.det(R) - 1) < 1e-9
        # centroids map onto each other, and for rank 1 the direction between the two distinct points is preserved
        assert np.abs(k1.mean(0) @ R.T + T[:, 3] - k0.mean(0)).max() < 1e-9
        if kind != 1:
            d1 = k1[1] - k1[0]; d0 = k0[1] - k0[0]
            assert np.abs(R @ (d1 / np.linalg.norm(d1)) - d0 / np.linalg.norm(d0)).max() < 1e-7


def test_quat_times_anchor_bit_exact(tables):
    L = _lib(); rng = np.random.default_rng(2)
    for i in range(300):
        q = rng.standard_normal(4).astype(np.float32); q /= np.linalg.norm(q)
        Rg = np.ascontiguousarray(tables.rot[int(rng.integers(60))].astype(np.float32))
        R = np.zeros((3, 3)); L.rr_host_quat_times_anchor(q.ctypes.data_as(fp), Rg.ctypes.data_as(fp), R.ctypes.data_as(dp))
        assert np.array_equal(R, O.matrix_from_quaternion(q) @ Rg)
