"""roreg_b200/dataio.py: the open3d-free reader of the reference's origin-data directories (dataops/dataset.py:41-129)."""
import os
import struct
import numpy as np
import pytest
from roreg_b200 import dataio

DEMO = "/root/reference/data/origin_data/demo/kitchen"


def _write_ply(path, xyz, fmt, extra=True):
    n = xyz.shape[0]
    props = "property float x\nproperty float y\nproperty float z\n" + ("property float nx\nproperty uchar red\n" if extra else "")
    head = f"ply\nformat {fmt} 1.0\ncomment made by a test\nelement vertex {n}\n{props}element face 0\nproperty list uchar int vertex_indices\nend_header\n"
    with open(path, "wb") as f:
        f.write(head.encode())
        for p in xyz.astype(np.float32):
            if fmt == "ascii":
                f.write((" ".join(repr(float(v)) for v in p) + (" 0.5 7" if extra else "") + " \n").encode())
            else:
                e = "<" if fmt == "binary_little_endian" else ">"
                f.write(struct.pack(e + "fff", *p) + (struct.pack(e + "fB", 0.5, 7) if extra else b""))


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
@pytest.mark.parametrize("extra", [True, False])
def test_ply_reader_formats(tmp_path, fmt, extra):
    xyz = np.random.RandomState(0).randn(257, 3) * 3
    p = str(tmp_path / "c.ply"); _write_ply(p, xyz, fmt, extra)
    got = dataio.read_ply_xyz(p)
    assert got.dtype == np.float64 and np.array_equal(got, xyz.astype(np.float32).astype(np.float64))


def test_scene_files_duck_type(tmp_path):
    root = str(tmp_path / "scene"); os.makedirs(f"{root}/PointCloud"); os.makedirs(f"{root}/Keypoints")
    rng = np.random.RandomState(1)
    clouds = [rng.rand(300, 3), rng.rand(280, 3), rng.rand(310, 3)]
    for k, c in enumerate(clouds):
        _write_ply(f"{root}/PointCloud/cloud_bin_{k}.ply", c, "ascii" if k else "binary_little_endian")
    idx = rng.permutation(300)[:50]
    np.savetxt(f"{root}/Keypoints/cloud_bin_0Keypoints.txt", idx)               # float text, as np.savetxt in the reference
    with open(f"{root}/PointCloud/gt.log", "w") as f:
        f.write("0\t 1\t 3\t\n0.0\t-1.0\t0.0\t0.5\n1.0\t0.0\t0.0\t-0.25\n0.0\t0.0\t1.0\t2.0\n0.000\t0.000\t0.000\t1.000\n")
        f.write("1 2 3\n1 0 0 0\n0 1 0 0\n0 0 1 0\n0 0 0 1\n")
    ds = dataio.SceneFiles(root, 3, "demo/scene", n_keypoints=40)
    assert ds.name == "demo/scene" and ds.pc_ids == ["0", "1", "2"] and ds.pair_ids == [("0", "1"), ("1", "2")]
    T = ds.get_transform("0", "1")
    assert T.dtype == np.float32 and np.array_equal(T, np.array([[0, -1, 0, 0.5], [1, 0, 0, -0.25], [0, 0, 1, 2.0]], np.float32))
    k0 = ds.get_kps("0")
    assert k0.dtype == np.float64 and np.array_equal(k0, clouds[0].astype(np.float32).astype(np.float64)[idx])
    # no index file: 40 random points from the global RNG (shuffle of arange), indices stored for the next run
    np.random.seed(9)
    k1 = ds.get_kps("1")
    np.random.seed(9); perm = np.arange(280); np.random.shuffle(perm)
    assert np.array_equal(k1, clouds[1].astype(np.float32).astype(np.float64)[perm[:40]])
    assert np.array_equal(np.loadtxt(f"{root}/Keypoints/cloud_bin_1Keypoints.txt").astype(np.int64), perm[:40])
    assert np.array_equal(dataio.SceneFiles(root, 3, "demo/scene", n_keypoints=40).get_kps(1), k1)      # second object reads the stored indices


@pytest.mark.skipif(not os.path.exists(DEMO), reason="the reference's demo scene lives only in the build container")
def test_reference_demo_scene():
    """The shipped demo pair (data/origin_data/demo/kitchen): 5000 keypoints per cloud; the ground-truth pose of gt.log maps cloud 1
    onto cloud 0 (R @ pts1 + t = pts0): most transformed keypoints of cloud 1 land within 5 cm of cloud 0's scan."""
    ds = dataio.SceneFiles(DEMO, 2, "demo/kitchen")
    assert ds.pair_ids == [("0", "1")]
    k0, k1 = ds.get_kps("0"), ds.get_kps("1")
    assert k0.shape == (5000, 3) and k1.shape == (5000, 3) and k0.dtype == np.float64
    pc0 = ds.get_pc("0")
    assert pc0.shape == (258342, 3)
    T = ds.get_transform("0", "1").astype(np.float64)
    moved = k1 @ T[:, :3].T + T[:, 3]
    sub = pc0[::8].astype(np.float32)
    d = np.array([np.sqrt(((sub - p.astype(np.float32)) ** 2).sum(1).min()) for p in moved[::25]])
    assert np.mean(d < 0.05) > 0.5, np.mean(d < 0.05)


def test_pair_file_writer_matches_numpy(tmp_path):
    """roreg_write_pair_files (plain C, used by the scene driver's writer threads): the .npy files are byte-identical to np.save
    of the same arrays, the .npz is read back by np.load with the keys / dtypes / shapes np.savez(trans=..., recalltime=...)
    produces (test/matcher.py:108-109, test/estimator.py:111,242).  No GPU involved."""
    from roreg_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for K in (0, 1, 7, 3400):
        m = rng.integers(0, 5000, (K, 2)).astype(np.int64); dr = rng.integers(0, 60, K).astype(np.int64); T = rng.random((4, 4))
        f = {k: str(tmp_path / f"{k}{K}") for k in ("m.npy", "s.npy", "d.npy", "r.npz")}
        rc = lib.roreg_write_pair_files(f["m.npy"].encode(), f["s.npy"].encode(), f["d.npy"].encode(), f["r.npz"].encode(),
                                        m.ctypes.data, dr.ctypes.data, K, T.ctypes.data, 50000)
        assert rc == 0
        for name, ref in (("m.npy", m), ("s.npy", np.ones(K)), ("d.npy", dr)):
            np.save(str(tmp_path / "ref.npy"), ref)
            assert open(f[name], "rb").read() == open(str(tmp_path / "ref.npy"), "rb").read(), (name, K)
        z = np.load(f["r.npz"], allow_pickle=True)
        np.savez(str(tmp_path / "ref.npz"), trans=T, recalltime=50000)
        zr = np.load(str(tmp_path / "ref.npz"))
        assert sorted(z.files) == sorted(zr.files) == ["recalltime", "trans"]
        for k in z.files:
            assert z[k].dtype == zr[k].dtype and z[k].shape == zr[k].shape and np.array_equal(z[k], zr[k])
    assert lib.roreg_write_pair_files(str(tmp_path / "no_such_dir" / "x.npy").encode(), None, None, None, m.ctypes.data, dr.ctypes.data, K, T.ctypes.data, 1) == -5
