"""Drop-in parity at the plugin boundary: the mirrors of test/{matcher,estimator}.py are run through
their reference signatures against a cache directory and the files they write are compared with the
golden fixtures recorded from the UNMODIFIED reference (tests/golden/make_golden.py)."""
import os
import types
import numpy as np
import pytest
from conftest import load_golden
from roreg_b200 import synth

pytestmark = pytest.mark.gpu


def _cfg(cache, **kw):
    c = types.SimpleNamespace(output_cache_fn=cache, model_fn="", SO3_related_files=None, backbone="FCGF",
                              bs_GF=1250, bs_ET=1000, RD=False, RM=False, match_n=0.5, ransac_ird=0.1,
                              keynum=5000, max_iter=1000)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


@pytest.mark.parametrize("name", ["s256", "s700"])
def test_plugins_reproduce_reference_files(name, tmp_path):
    import roreg_b200.test as rt
    z, n, keynum, max_iter, seeds = load_golden(name)
    ds = synth.SynthDataset(seeds, n=n, name=f"synth/{name}", max_res_deg=2.0)
    cache = str(tmp_path / "cache")
    ds.write_cache(cache)
    cfg = _cfg(cache)
    base = f"{cache}/{ds.name}/match_{keynum}"
    # registries and signatures are the reference's (test/__init__.py:6-22)
    assert set(rt.name2matcher) == {"matmul", "yoho_mat"} and set(rt.name2estimator) == {"yohoc", "yohoo"}
    # ---- matcher: same global-RNG consumption as the reference run that produced the fixture
    np.random.seed(1234)
    rt.name2matcher["matmul"](cfg).run(ds, keynum)
    for (id0, id1) in ds.pair_ids:
        m = np.load(f"{base}/{id0}-{id1}.npy"); s = np.load(f"{base}/scores/{id0}-{id1}.npy")
        assert m.dtype == np.int64 and np.array_equal(m, z[f"match_{id0}-{id1}"])
        assert s.dtype == np.float64 and np.array_equal(s, z[f"scores_{id0}-{id1}"])
    # ---- coarse rotation index
    rt.extractor_dr_index(cfg).Rindex(ds, keynum)
    for (id0, id1) in ds.pair_ids:
        d = np.load(f"{base}/DR_index/{id0}-{id1}.npy")
        assert d.dtype == np.int64 and np.array_equal(d, z[f"dr_index_{id0}-{id1}"])
    # ---- one-shot RANSAC on the reference's own Trans_pre (the ET network is outside this build)
    os.makedirs(f"{base}/Trans_pre", exist_ok=True)
    for (id0, id1) in ds.pair_ids:
        np.save(f"{base}/Trans_pre/{id0}-{id1}.npy", z[f"trans_pre_{id0}-{id1}"])
    np.random.seed(4321)
    rt.yohoo_ransac(cfg).ransac(ds, keynum, max_iter)
    for (id0, id1) in ds.pair_ids:
        r = np.load(f"{base}/yohoo/{max_iter}iters/{id0}-{id1}.npz")
        assert int(r["recalltime"]) == int(z[f"yohoo_recall_{id0}-{id1}"])
        assert np.abs(r["trans"] - z[f"yohoo_trans_{id0}-{id1}"]).max() < 1e-9      # stated tolerance on poses: 1e-4
    got = open(f"{base}/yohoo/{max_iter}iters/pre.log", "rb").read()
    ref = bytes(z["pre_log_yohoo"])
    assert len(got.splitlines()) == len(ref.splitlines())
    for a, b in zip(got.decode().split(), ref.decode().split()):
        assert abs(float(a) - float(b)) < 1e-9
    # ---- yohoc, parity mode (host RNG + host LAPACK for the rank-2 Kabsch, device scoring/refine)
    os.makedirs(f"{base}/yohoc/{max_iter}iters", exist_ok=True)
    yc = rt.yohoc_ransac(cfg)
    for pi, pair in enumerate(ds.pair_ids):
        np.random.seed(777 + pi)
        yc.ransac_once(ds, keynum, max_iter, pair)
        id0, id1 = pair
        r = np.load(f"{base}/yohoc/{max_iter}iters/{id0}-{id1}.npz")
        assert int(r["recalltime"]) == int(z[f"yohoc_recall_{id0}-{id1}"])
        assert np.abs(r["trans"] - z[f"yohoc_trans_{id0}-{id1}"]).max() < 1e-9


def test_yohoc_run_device_mode(tmp_path):
    """yohoc.run end to end (Rindex + ransac + pre.log) with device-side draws: poses agree with ground truth."""
    import roreg_b200.test as rt
    ds = synth.SynthDataset([61, 62], n=500, name="synth/dev")
    cache = str(tmp_path / "cache"); ds.write_cache(cache)
    cfg = _cfg(cache, yohoc_mode="device")
    np.random.seed(9)
    rt.mutual(cfg).run(ds, 500)
    rt.yohoc(cfg).run(ds, 500, 400)
    for pi, (id0, id1) in enumerate(ds.pair_ids):
        r = np.load(f"{cache}/{ds.name}/match_500/yohoc/400iters/{id0}-{id1}.npz")
        assert np.abs(r["trans"][:3] - ds.pairs[pi]["gt"]).max() < 5e-3
    assert os.path.exists(f"{cache}/{ds.name}/match_500/yohoc/400iters/pre.log")


def test_missing_checkpoint_raises_like_the_reference():
    """`raise ValueError("No model exists")` when a checkpoint is absent (test/matcher.py:129, test/detector.py:24)."""
    import roreg_b200.test as rt
    cfg = _cfg("/tmp", model_fn="/nonexistent")
    with pytest.raises(ValueError):
        rt.yoho_mat(cfg)
    with pytest.raises(ValueError):
        rt.yoho_det(cfg)
    with pytest.raises(ValueError):
        rt.yoho_des(cfg).run(types.SimpleNamespace(name="x", pc_ids=[], pair_ids=[]))


def test_rd_rm_paths_reproduce_reference_files(tmp_path):
    """--RD (NMS sampling on detector scores, device 5-NN) and --RM (top-`match_n` selection by score, float32 match scores)
    through the real context, against tests/golden/s400rdrm.npz (tests/golden/make_golden_rd_rm.py, unmodified reference).
    CPU twin: tests/test_plugins_host.py::test_rd_rm_paths_reproduce_reference_files_on_host."""
    import roreg_b200.test as rt
    z, n, keynum, max_iter, seeds = load_golden("s400rdrm")
    ds = synth.SynthDataset(seeds, n=n, name="synth/s400rdrm", max_res_deg=2.0)
    cache = str(tmp_path / "cache")
    ds.write_cache(cache)
    os.makedirs(f"{cache}/{ds.name}/det_score", exist_ok=True)
    for cid in ds.pc_ids:
        np.save(f"{cache}/{ds.name}/det_score/{cid}.npy", z[f"det_score_{cid}"])
    cfg = _cfg(cache, RD=True, RM=True, match_n=0.5)
    base = f"{cache}/{ds.name}/match_{keynum}"
    keys = ds.get_kps(ds.pc_ids[0]); sc0 = z[f"det_score_{ds.pc_ids[0]}"]
    for num in (40, 300, 380, 500):
        assert np.array_equal(rt.NMS_sample(num, 5, cfg).sample(keys, sc0), z[f"nms_{num}"])
    for num in (40, 300):
        assert np.array_equal(rt.NMS_sample(num, 5, cfg).sample(keys, z["nms_flat_scores"]), z[f"nms_flat_{num}"])
    np.random.seed(2468)
    rt.mutual(cfg).run(ds, keynum)
    for (id0, id1) in ds.pair_ids:
        assert np.array_equal(np.load(f"{base}/{id0}-{id1}.npy"), z[f"match_{id0}-{id1}"])
        np.save(f"{base}/scores/{id0}-{id1}.npy", z[f"scores_{id0}-{id1}"])
    rt.extractor_dr_index(cfg).Rindex(ds, keynum)
    os.makedirs(f"{base}/Trans_pre", exist_ok=True)
    for (id0, id1) in ds.pair_ids:
        assert np.array_equal(np.load(f"{base}/DR_index/{id0}-{id1}.npy"), z[f"dr_index_{id0}-{id1}"])
        np.save(f"{base}/Trans_pre/{id0}-{id1}.npy", z[f"trans_pre_{id0}-{id1}"])
    np.random.seed(1357)
    rt.yohoo_ransac(cfg).ransac(ds, keynum, max_iter)
    yc = rt.yohoc_ransac(cfg)
    os.makedirs(f"{base}/yohoc/{max_iter}iters", exist_ok=True)
    for pi, pair in enumerate(ds.pair_ids):
        np.random.seed(555 + pi)
        yc.ransac_once(ds, keynum, max_iter, pair)
    for (id0, id1) in ds.pair_ids:
        for est in ("yohoo", "yohoc"):
            r = np.load(f"{base}/{est}/{max_iter}iters/{id0}-{id1}.npz")
            assert int(r["recalltime"]) == int(z[f"{est}_recall_{id0}-{id1}"]), est
            assert np.abs(r["trans"] - z[f"{est}_trans_{id0}-{id1}"]).max() < 2e-6, est   # float32 weights, normalised in float32 by the reference


@pytest.mark.parametrize("nn_mode,corr_mode", [(0, 0), (4, 3)])
def test_scene_driver_equals_the_plugins(tmp_path, nn_mode, corr_mode):
    """roreg_b200.scene.register_scene (clouds uploaded once, batched engine, background writer): same samples and - with the
    reference-arithmetic kernels - the same match / DR_index files as the mutual + Rindex plugin passes; file contract and
    poses in both kernel configurations.  CPU twin: tests/test_scene_host.py."""
    import roreg_b200.test as rt
    from roreg_b200 import scene
    ds = synth.SynthDataset([81, 82, 83], n=600, name="synth/scene3", with_fcgf=False)
    keynum, max_iter = 500, 300
    a = str(tmp_path / "a"); b = str(tmp_path / "b")
    ds.write_cache(a); ds.write_cache(b)
    np.random.seed(77); rt.mutual(_cfg(a, corr_mode=0)).run(ds, keynum); rt.extractor_dr_index(_cfg(a, corr_mode=0)).Rindex(ds, keynum)
    np.random.seed(77)
    res = scene.register_scene(_cfg(b, corr_mode=corr_mode), ds, keynum=keynum, max_iter=max_iter, batch_pairs=2, nn_mode=nn_mode)
    from roreg_b200.test._common import context
    context(_cfg(b, corr_mode=0))                                   # leave the shared context in the reference-arithmetic mode
    assert (res["lo"], res["hi"]) == (0, 3)
    base_a = f"{a}/{ds.name}/match_{keynum}"; base_b = f"{b}/{ds.name}/match_{keynum}"
    for pi, (id0, id1) in enumerate(ds.pair_ids):
        ma = np.load(f"{base_a}/{id0}-{id1}.npy"); mb = np.load(f"{base_b}/{id0}-{id1}.npy")
        da = np.load(f"{base_a}/DR_index/{id0}-{id1}.npy"); db = np.load(f"{base_b}/DR_index/{id0}-{id1}.npy")
        assert mb.dtype == np.int64 and db.dtype == np.int64
        if nn_mode == 0:
            assert np.array_equal(ma, mb) and np.array_equal(da, db)
        else:
            assert len({tuple(r) for r in ma.tolist()} ^ {tuple(r) for r in mb.tolist()}) <= 4      # near ties of the two NN arithmetics
        s = np.load(f"{base_b}/scores/{id0}-{id1}.npy")
        assert s.dtype == np.float64 and np.array_equal(s, np.ones(mb.shape[0]))
        r = np.load(f"{base_b}/yohoc/{max_iter}iters/{id0}-{id1}.npz")
        assert 1 <= int(r["recalltime"]) <= max_iter and np.abs(r["trans"][:3] - ds.pairs[pi]["gt"]).max() < 1e-2
    assert len(open(f"{base_b}/yohoc/{max_iter}iters/pre.log").read().splitlines()) == 5 * len(ds.pair_ids)
