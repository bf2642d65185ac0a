"""Host logic of the plugin mirrors (roreg_b200/test) on the CPU: the same drop-in scenario as tests/test_gpu_dropin.py - run
the plugins through their reference signatures against a cache directory, compare the files they write with the fixtures
recorded from the UNMODIFIED reference (tests/golden/make_golden.py) - but with the device context replaced by the oracle
(tests/_host_ctx.py).  What is pinned here is everything around the kernels: keypoint sampling and the consumption order of
the global NumPy RNG, index bookkeeping, dtypes and layout of every file, the yohoc draw loop + host Kabsch, the
--RM top-score selection, pre.log, and the error behaviour."""
import os
import types
import numpy as np
import pytest
from conftest import load_golden
from roreg_b200 import synth
import _host_ctx


def _cfg(cache, **kw):
    c = types.SimpleNamespace(output_cache_fn=cache, model_fn="", SO3_related_files=None, backbone="FCGF",
                              bs_GF=1250, bs_ET=1000, RD=False, RM=False, match_n=0.5, ransac_ird=0.1,
                              keynum=5000, max_iter=1000)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


@pytest.mark.parametrize("name", ["s256", "s700"])
def test_plugins_reproduce_reference_files_on_host(name, tmp_path, monkeypatch):
    _host_ctx.install(monkeypatch)
    import roreg_b200.test as rt
    z, n, keynum, max_iter, seeds = load_golden(name)
    ds = synth.SynthDataset(seeds, n=n, name=f"synth/{name}", max_res_deg=2.0)
    cache = str(tmp_path / "cache")
    ds.write_cache(cache)
    cfg = _cfg(cache)
    base = f"{cache}/{ds.name}/match_{keynum}"
    assert set(rt.name2matcher) == {"matmul", "yoho_mat"} and set(rt.name2estimator) == {"yohoc", "yohoo"}
    assert set(rt.name2extractor) == {"yoho_des"} and set(rt.name2detector) == {"yoho_det"}
    np.random.seed(1234)
    rt.name2matcher["matmul"](cfg).run(ds, keynum)
    for (id0, id1) in ds.pair_ids:
        m = np.load(f"{base}/{id0}-{id1}.npy"); s = np.load(f"{base}/scores/{id0}-{id1}.npy")
        assert m.dtype == np.int64 and np.array_equal(m, z[f"match_{id0}-{id1}"])
        assert s.dtype == np.float64 and np.array_equal(s, z[f"scores_{id0}-{id1}"])
    rt.extractor_dr_index(cfg).Rindex(ds, keynum)
    for (id0, id1) in ds.pair_ids:
        d = np.load(f"{base}/DR_index/{id0}-{id1}.npy")
        assert d.dtype == np.int64 and np.array_equal(d, z[f"dr_index_{id0}-{id1}"])
    os.makedirs(f"{base}/Trans_pre", exist_ok=True)
    for (id0, id1) in ds.pair_ids:
        np.save(f"{base}/Trans_pre/{id0}-{id1}.npy", z[f"trans_pre_{id0}-{id1}"])
    np.random.seed(4321)
    rt.yohoo_ransac(cfg).ransac(ds, keynum, max_iter)
    for (id0, id1) in ds.pair_ids:
        r = np.load(f"{base}/yohoo/{max_iter}iters/{id0}-{id1}.npz")
        assert int(r["recalltime"]) == int(z[f"yohoo_recall_{id0}-{id1}"])
        assert r["trans"].shape == (4, 4) and np.abs(r["trans"] - z[f"yohoo_trans_{id0}-{id1}"]).max() < 1e-12
    got = open(f"{base}/yohoo/{max_iter}iters/pre.log", "rb").read().decode().splitlines()
    ref = bytes(z["pre_log_yohoo"]).decode().splitlines()
    assert len(got) == len(ref) == 5 * len(ds.pair_ids)
    for i, (a, b) in enumerate(zip(got, ref)):
        if i % 5 in (0, 4):
            assert a == b                                           # 'id0\tid1\tn_clouds' and the constant last row: byte for byte
        else:
            fa, fb = a.split("\t"), b.split("\t")
            assert len(fa) == len(fb) == 4 and max(abs(float(x) - float(y)) for x, y in zip(fa, fb)) < 1e-12
    # number formatting: exactly what the reference's f-strings print for the poses stored in the .npz files
    want = ""
    for (id0, id1) in ds.pair_ids:
        T = np.load(f"{base}/yohoo/{max_iter}iters/{id0}-{id1}.npz")["trans"]
        want += f"{int(id0)}\t{int(id1)}\t{len(ds.pc_ids)}\n"
        for r in range(3):
            want += f"{T[r][0]}\t{T[r][1]}\t{T[r][2]}\t{T[r][3]}\n"
        want += f"{0.0}\t{0.0}\t{0.0}\t{1.0}\n"
    assert open(f"{base}/yohoo/{max_iter}iters/pre.log").read() == want
    os.makedirs(f"{base}/yohoc/{max_iter}iters", exist_ok=True)
    yc = rt.yohoc_ransac(cfg)
    for pi, pair in enumerate(ds.pair_ids):
        np.random.seed(777 + pi)
        yc.ransac_once(ds, keynum, max_iter, pair)
        id0, id1 = pair
        r = np.load(f"{base}/yohoc/{max_iter}iters/{id0}-{id1}.npz")
        assert int(r["recalltime"]) == int(z[f"yohoc_recall_{id0}-{id1}"])
        assert np.abs(r["trans"] - z[f"yohoc_trans_{id0}-{id1}"]).max() < 1e-12


def test_rd_rm_paths_reproduce_reference_files_on_host(tmp_path, monkeypatch):
    """--RD (NMS sampling on detector scores) and --RM (top-`match_n` selection by score) against tests/golden/s400rdrm.npz
    (tests/golden/make_golden_rd_rm.py, unmodified reference)."""
    _host_ctx.install(monkeypatch)
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
    # the sampler alone: trim branch, top-up branch, fewer points than requested, heavy score ties
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
            assert np.abs(r["trans"] - z[f"{est}_trans_{id0}-{id1}"]).max() < 1e-10, est     # float32 weights: the reference normalises them in float32
    # host helpers of yohoc_ransac (test/estimator.py:119-147)
    dr = z[f"dr_index_{ds.pair_ids[0][0]}-{ds.pair_ids[0][1]}"]
    stat, prob = yc.DR_statictic(dr)
    assert np.array_equal(prob, z["drstat_prob"])
    assert np.array_equal(np.array([len(stat[i]) for i in range(60)]), z["drstat_counts"])
    assert np.array_equal(np.concatenate([np.asarray(stat[i], np.int64) for i in range(60)]), z["drstat_members"])
    assert yc.DR_statictic(np.arange(60))[0] is None and np.array_equal(yc.DR_statictic(np.arange(60))[1], np.zeros(60))
    k0 = ds.get_kps(ds.pair_ids[0][0]); k1 = ds.get_kps(ds.pair_ids[0][1])
    for t, T in zip(z["kabsch_triplets"], z["kabsch_T"]):
        assert np.array_equal(yc.Threepps2Tran(k0[t], k1[t]), T)


def test_plugin_error_behaviour_on_host(tmp_path, monkeypatch):
    """No mutual match -> the ValueError np.concatenate([]) raises in the reference (test/matcher.py:106); a pair whose coarse
    rotations never repeat -> yohoc writes a random 4x4 with recalltime 50000 and returns 0 (test/estimator.py:214-218)."""
    _host_ctx.install(monkeypatch)
    import roreg_b200.test as rt
    ds = synth.SynthDataset([71], n=64, name="synth/err")
    cache = str(tmp_path / "cache"); ds.write_cache(cache)
    cfg = _cfg(cache)
    base = f"{cache}/{ds.name}/match_64"
    os.makedirs(f"{base}/scores"); os.makedirs(f"{base}/DR_index"); os.makedirs(f"{base}/yohoc/50iters")
    id0, id1 = ds.pair_ids[0]
    np.save(f"{base}/{id0}-{id1}.npy", np.stack([np.arange(60), np.arange(60)], 1))
    np.save(f"{base}/scores/{id0}-{id1}.npy", np.ones(60))
    np.save(f"{base}/DR_index/{id0}-{id1}.npy", np.arange(60))           # every coarse rotation exactly once
    np.random.seed(5)
    assert rt.yohoc_ransac(cfg).ransac_once(ds, 64, 50, (id0, id1)) == 0
    r = np.load(f"{base}/yohoc/50iters/{id0}-{id1}.npz")
    np.random.seed(5)
    assert int(r["recalltime"]) == 50000 and np.array_equal(r["trans"], np.random.rand(4, 4)) and r["center"].shape == (6, 3)


def test_per_cloud_plugins_on_host(tmp_path, monkeypatch):
    """yoho_des.run / yoho_det.run (test/extractor.py:33-60, test/detector.py:26-47): files, dtypes, skip-if-cached, clouds addressed
    by position, rank normalisation - with the networks replaced by the oracle's forward passes; outputs against the
    reference-written fixture (random weights from the same seeds as tests/golden/make_golden.py)."""
    import torch
    from oracle import roreg_oracle as O
    _host_ctx.install(monkeypatch); _host_ctx.install_nets(monkeypatch)
    import roreg_b200.test as rt
    z, n, keynum, max_iter, seeds = load_golden("s256")
    ds = synth.SynthDataset(seeds[:1], n=n, name="synth/s256", max_res_deg=2.0)          # clouds 0, 1
    cache = str(tmp_path / "cache"); ds.write_cache(cache, yoho=False)
    model_fn = str(tmp_path / "ckpt")
    for kind, seed in (("GF", 101), ("RD", 103)):
        os.makedirs(f"{model_fn}/{kind}")
        torch.save({"best_para": 0, "network_state_dict": {k: torch.from_numpy(v) for k, v in O.random_state_dict(kind, seed).items()}},
                   f"{model_fn}/{kind}/model_best.pth")
    cfg = _cfg(cache, model_fn=model_fn)
    out_dir = f"{cache}/{ds.name}/YOHO_Output_Group_feature"
    os.makedirs(out_dir)
    sentinel = np.full((3, 32, 60), 7, np.float32)
    np.save(f"{out_dir}/1.npy", sentinel)                                                  # cloud 1 is "already cached"
    rt.yoho_des(cfg).run(ds)
    got = np.load(f"{out_dir}/0.npy")
    assert got.dtype == np.float32 and got.shape == (n, 32, 60) and np.abs(got[:40] - z["gf_eqv_0"]).max() < 5e-6
    assert np.array_equal(np.load(f"{out_dir}/1.npy"), sentinel)
    det_dir = f"{cache}/{ds.name}/det_score"
    os.makedirs(det_dir)
    np.save(f"{det_dir}/1.npy", np.zeros(3))
    rt.yoho_det(cfg).run(ds)
    s = np.load(f"{det_dir}/0.npy")
    assert s.shape == (n,) and np.array_equal(np.sort(s), (np.arange(n) / n).astype(s.dtype))    # a permutation of rank / N
    assert np.mean(np.abs(s - z["det_score_0"]) * n <= 1.0) > 0.98                          # ranks agree up to float32 near-ties
    assert np.array_equal(np.load(f"{det_dir}/1.npy"), np.zeros(3))
    with pytest.raises(ValueError):
        rt.yoho_des(_cfg(cache, model_fn=str(tmp_path / "none"))).run(ds)


def test_yohoo_run_end_to_end_on_host(tmp_path, monkeypatch):
    """yohoo.run = Rindex + Rt_pre (ET network, side swap of test/estimator.py:293-306, pose arithmetic :349-366) + one-shot
    RANSAC, with the ET forward pass from the oracle: Trans_pre and the final files against the reference-written fixture."""
    import torch
    from oracle import roreg_oracle as O
    _host_ctx.install(monkeypatch); _host_ctx.install_nets(monkeypatch)
    import roreg_b200.test as rt
    z, n, keynum, max_iter, seeds = load_golden("s256")
    ds = synth.SynthDataset(seeds, n=n, name="synth/s256", max_res_deg=2.0)
    cache = str(tmp_path / "cache"); ds.write_cache(cache)
    model_fn = str(tmp_path / "ckpt"); os.makedirs(f"{model_fn}/ET")
    torch.save({"best_para": 0, "network_state_dict": {k: torch.from_numpy(v) for k, v in O.random_state_dict("ET", 102).items()}},
               f"{model_fn}/ET/model_best.pth")
    cfg = _cfg(cache, model_fn=model_fn)
    base = f"{cache}/{ds.name}/match_{keynum}"
    os.makedirs(f"{base}/scores")
    for (id0, id1) in ds.pair_ids:                                   # matcher output of the reference run
        np.save(f"{base}/{id0}-{id1}.npy", z[f"match_{id0}-{id1}"]); np.save(f"{base}/scores/{id0}-{id1}.npy", z[f"scores_{id0}-{id1}"])
    np.random.seed(4321)
    rt.name2estimator["yohoo"](cfg).run(ds, keynum, max_iter)
    for (id0, id1) in ds.pair_ids:
        tr = np.load(f"{base}/Trans_pre/{id0}-{id1}.npy")
        assert tr.dtype == np.float64 and tr.shape == z[f"trans_pre_{id0}-{id1}"].shape
        assert np.abs(tr - z[f"trans_pre_{id0}-{id1}"]).max() < 2e-5          # float32 network, tolerance of tests/test_oracle_golden.py
        r = np.load(f"{base}/yohoo/{max_iter}iters/{id0}-{id1}.npz")
        assert int(r["recalltime"]) == int(z[f"yohoo_recall_{id0}-{id1}"])
        assert np.abs(r["trans"] - z[f"yohoo_trans_{id0}-{id1}"]).max() < 1e-4
    assert os.path.exists(f"{base}/yohoo/{max_iter}iters/pre.log")


def test_yoho_mat_plugin_on_host(tmp_path, monkeypatch):
    """yoho_mat.run (test/matcher.py:152-210): sampling, the source / target swap, index mapping back through the samples, float32
    scores - with Match_ot's forward pass from the oracle; files against tests/golden/rm300.npz (unmodified reference)."""
    import torch
    from oracle import roreg_oracle as O
    _host_ctx.install(monkeypatch); _host_ctx.install_nets(monkeypatch)
    import roreg_b200.test as rt
    z, n, keynum, _, seeds = load_golden("rm300")
    ds = synth.SynthDataset(seeds, n=n, name="synth/rm", with_fcgf=False)
    cache = str(tmp_path / "cache"); ds.write_cache(cache)
    model_fn = str(tmp_path / "ckpt"); os.makedirs(f"{model_fn}/RM")
    torch.save({"best_para": 0, "network_state_dict": {k: torch.from_numpy(np.asarray(v)) for k, v in O.random_state_dict("RM", 104).items()}},
               f"{model_fn}/RM/model_best.pth")
    cfg = _cfg(cache, model_fn=model_fn, RM=True)
    np.random.seed(2468)
    rt.name2matcher["yoho_mat"](cfg).run(ds, keynum)
    for (id0, id1) in ds.pair_ids:
        m = np.load(f"{cache}/synth/rm/match_{keynum}/{id0}-{id1}.npy"); s = np.load(f"{cache}/synth/rm/match_{keynum}/scores/{id0}-{id1}.npy")
        ref = z[f"match_{id0}-{id1}"]
        assert m.dtype == ref.dtype and s.dtype == z[f"scores_{id0}-{id1}"].dtype == np.float32
        a = {tuple(r) for r in m.tolist()}; b = {tuple(r) for r in ref.tolist()}
        assert len(a ^ b) <= max(2, len(b) // 50), (len(a), len(b), len(a ^ b))      # float32 top-k / argmax near ties (as the GPU test)
        if np.array_equal(m, ref):
            assert np.abs(s - z[f"scores_{id0}-{id1}"]).max() < 1e-3


def test_host_rules_equal_the_oracle_on_random_inputs():
    """The plugin-side host rules (roreg_b200/test/_hostlogic.py) against the oracle's independent restatements on seeded random
    inputs beyond the fixtures: heavy score ties, tiny buckets, --RM thresholds around the `max(.., 10)` floor, degenerate triplets."""
    from oracle import roreg_oracle as O
    from roreg_b200.test import _hostlogic as host
    rng = np.random.RandomState(42)
    for trial in range(40):
        n = int(rng.randint(20, 400)); k = 5
        keys = rng.rand(n, 3)
        scores = rng.rand(n) if trial % 3 else np.round(rng.rand(n) * 6) / 6          # every third trial: many equal scores
        _, nn = O.knn(keys.astype(np.float32), keys.astype(np.float32), k)
        for num in (1, n // 7 + 1, n // 2, n - 1, n):
            assert np.array_equal(host.nms_select(scores, nn, num), O.nms_sample(keys, scores, num, k)), (trial, num)
    for trial in range(40):
        K = int(rng.randint(1, 300))
        dr = rng.randint(0, 60 if trial % 2 else 7, K)
        members, prob = host.rotation_buckets(dr)
        ref_members, ref_prob = O.dr_statistic(dr)
        assert np.array_equal(prob, ref_prob)
        assert (members is None) == (ref_members is None)
        if members is not None:
            assert all(np.array_equal(members[r], np.asarray(ref_members[r], np.int64)) for r in range(60))
    for K in (3, 9, 10, 19, 20, 21, 57, 500):
        s = rng.rand(K).astype(np.float32)
        for match_n in (0.05, 0.5, 0.998, 0.9995, 5, 40):
            assert np.array_equal(host.top_scored(s, match_n), O.select_hypotheses(s, K, True, match_n)), (K, match_n)
    for trial in range(50):
        p0 = rng.rand(3, 3) * 3; p1 = rng.rand(3, 3) * 3
        if trial % 10 == 0:
            p1[2] = p1[1]                                                             # repeated point: rank-1 cross-covariance
        assert np.array_equal(host.kabsch_3pt(p0, p1), O.threepps2tran(p0, p1))
    # the guided draw loop consumes the global RNG exactly as the oracle's restatement of test/estimator.py:221-228
    dr = rng.randint(0, 12, 150)
    members, prob = host.rotation_buckets(dr)
    np.random.seed(31); mine = host.draw_guided_triplets(members, prob, 64)
    np.random.seed(31); ref, _, _ = O.yohoc_draws(dr, 64)
    assert mine.shape == (64, 3) and np.array_equal(mine, np.stack([d[1] for d in ref]))
    np.random.seed(31); host.draw_guided_triplets(members, prob, 64); after_mine = np.random.rand()
    np.random.seed(31); O.yohoc_draws(dr, 64); after_ref = np.random.rand()
    assert after_mine == after_ref                                                    # same RNG state afterwards


def test_yohoc_run_device_mode_on_host(tmp_path, monkeypatch):
    """yohoc.run end to end (Rindex + ransac + pre.log) with cfg.yohoc_mode = 'device' (private Generator draws + kabsch3 entry point):
    CPU twin of tests/test_gpu_dropin.py::test_yohoc_run_device_mode."""
    _host_ctx.install(monkeypatch)
    import roreg_b200.test as rt
    ds = synth.SynthDataset([61, 62], n=300, name="synth/dev")
    cache = str(tmp_path / "cache"); ds.write_cache(cache)
    cfg = _cfg(cache, yohoc_mode="device")
    np.random.seed(9)
    rt.mutual(cfg).run(ds, 300)
    rt.yohoc(cfg).run(ds, 300, 200)
    for pi, (id0, id1) in enumerate(ds.pair_ids):
        r = np.load(f"{cache}/{ds.name}/match_300/yohoc/200iters/{id0}-{id1}.npz")
        assert 1 <= int(r["recalltime"]) <= 200 and np.abs(r["trans"][:3] - ds.pairs[pi]["gt"]).max() < 1e-2
    assert os.path.exists(f"{cache}/{ds.name}/match_300/yohoc/200iters/pre.log")
