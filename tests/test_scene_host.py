"""roreg_b200/scene.py (whole-dataset driver: clouds uploaded once, pairs through the batched engine, file contract emitted by a
background writer) on the CPU with the oracle-backed context: same sampling order / match files as the `mutual` plugin, the
reference's file contract, sharding over two gloo ranks."""
import os
import types
import numpy as np
import torch.multiprocessing as mp
from roreg_b200 import synth
import _host_ctx


def _cfg(cache, **kw):
    c = types.SimpleNamespace(output_cache_fn=cache, model_fn="", SO3_related_files=None, backbone="FCGF", bs_GF=1250, bs_ET=1000,
                              RD=False, RM=False, match_n=0.5, ransac_ird=0.1)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def _check_contract(cache, ds, keynum, max_iter):
    base = f"{cache}/{ds.name}/match_{keynum}"
    for pi, (id0, id1) in enumerate(ds.pair_ids):
        m = np.load(f"{base}/{id0}-{id1}.npy"); s = np.load(f"{base}/scores/{id0}-{id1}.npy"); d = np.load(f"{base}/DR_index/{id0}-{id1}.npy")
        assert m.dtype == np.int64 and m.ndim == 2 and m.shape[1] == 2 and s.dtype == np.float64 and np.array_equal(s, np.ones(m.shape[0]))
        assert d.dtype == np.int64 and d.shape == (m.shape[0],) and d.min() >= 0 and d.max() < 60
        r = np.load(f"{base}/yohoc/{max_iter}iters/{id0}-{id1}.npz")
        assert r["trans"].shape == (4, 4) and 1 <= int(r["recalltime"]) <= max_iter
        assert np.abs(r["trans"][:3] - ds.pairs[pi]["gt"]).max() < 1e-2
    log = open(f"{base}/yohoc/{max_iter}iters/pre.log").read().splitlines()
    assert len(log) == 5 * len(ds.pair_ids) and log[0] == f"{int(ds.pair_ids[0][0])}\t{int(ds.pair_ids[0][1])}\t{len(ds.pc_ids)}"


def test_scene_driver_equals_the_plugins_on_host(tmp_path, monkeypatch):
    ctx = _host_ctx.install(monkeypatch)
    import roreg_b200.test as rt
    from roreg_b200 import scene
    ds = synth.SynthDataset([81, 82, 83], n=300, name="synth/scene3", with_fcgf=False)
    keynum, max_iter = 200, 150
    a = str(tmp_path / "a"); b = str(tmp_path / "b")
    ds.write_cache(a); ds.write_cache(b)
    np.random.seed(77); rt.mutual(_cfg(a)).run(ds, keynum); rt.extractor_dr_index(_cfg(a)).Rindex(ds, keynum)
    np.random.seed(77)
    res = scene.register_scene(_cfg(b), ds, keynum=keynum, max_iter=max_iter, batch_pairs=2, nn_mode=0, ctx=ctx)
    assert (res["lo"], res["hi"]) == (0, 3) and res["poses"].shape == (3, 4, 4)
    for (id0, id1) in ds.pair_ids:                                # same RNG order, same samples, same match rows as the plugin pass
        for sub in ("", "scores/", "DR_index/"):
            assert np.array_equal(np.load(f"{a}/{ds.name}/match_{keynum}/{sub}{id0}-{id1}.npy"), np.load(f"{b}/{ds.name}/match_{keynum}/{sub}{id0}-{id1}.npy"))
    _check_contract(b, ds, keynum, max_iter)
    assert np.array_equal(res["n_matches"], [np.load(f"{b}/{ds.name}/match_{keynum}/{i}-{j}.npy").shape[0] for i, j in ds.pair_ids])


def test_scene_driver_rd_sampling_on_host(tmp_path, monkeypatch):
    ctx = _host_ctx.install(monkeypatch)
    import roreg_b200.test as rt
    from roreg_b200 import scene
    ds = synth.SynthDataset([84, 85], n=300, name="synth/scene2", with_fcgf=False)
    keynum, max_iter = 220, 150
    rng = np.random.RandomState(3)
    caches = [str(tmp_path / "a"), str(tmp_path / "b")]
    dets = {cid: rng.permutation(300) / 300 for cid in ds.pc_ids}
    for c in caches:
        ds.write_cache(c); os.makedirs(f"{c}/{ds.name}/det_score")
        for cid in ds.pc_ids:
            np.save(f"{c}/{ds.name}/det_score/{cid}.npy", dets[cid])
    rt.mutual(_cfg(caches[0], RD=True)).run(ds, keynum)
    scene.register_scene(_cfg(caches[1], RD=True), ds, keynum=keynum, max_iter=max_iter, batch_pairs=64, nn_mode=0, ctx=ctx)
    for (id0, id1) in ds.pair_ids:
        assert np.array_equal(np.load(f"{caches[0]}/{ds.name}/match_{keynum}/{id0}-{id1}.npy"), np.load(f"{caches[1]}/{ds.name}/match_{keynum}/{id0}-{id1}.npy"))
    _check_contract(caches[1], ds, keynum, max_iter)


def _worker(rank, world, port, cache, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import traceback
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    try:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import _host_ctx as hc
        from roreg_b200 import scene, synth as sy
        ds = sy.SynthDataset([86, 87, 88], n=200, name="synth/sharded", with_fcgf=False)
        np.random.seed(5)
        res = scene.register_scene(_cfg(cache), ds, keynum=200, max_iter=100, batch_pairs=1, nn_mode=0, ctx=hc.HostContext())
        q.put((rank, res["lo"], res["hi"], res["poses"]))
        dist.barrier()
        dist.destroy_process_group()
    except BaseException:
        q.put((rank, "error", traceback.format_exc(), None))       # a failing rank is reported at once, not by the queue's timeout
        raise


def test_scene_driver_world2_gloo(tmp_path):
    ds = synth.SynthDataset([86, 87, 88], n=200, name="synth/sharded", with_fcgf=False)
    cache = str(tmp_path / "c"); ds.write_cache(cache)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cache, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs: p.join(timeout=120)
    assert not [r for r in res if r[1] == "error"], "\n".join(r[2] for r in res if r[1] == "error")
    res = sorted(res, key=lambda r: r[0])
    assert [(r[1], r[2]) for r in res] == [(0, 2), (2, 3)]                 # 3 pairs over 2 ranks
    _check_contract(cache, ds, 200, 100)                                   # every pair's files + rank 0's pre.log
    for rank, lo, hi, poses in res:
        for i in range(lo, hi):
            id0, id1 = ds.pair_ids[i]
            assert np.array_equal(poses[i - lo], np.load(f"{cache}/{ds.name}/match_200/yohoc/100iters/{id0}-{id1}.npz")["trans"])


def test_scene_loader_chunked_reads_and_error_paths(tmp_path, monkeypatch):
    """SceneLoader reads every cloud file in READ_CHUNK pieces through its reader pool: the arena must equal np.load of the files
    for chunk sizes that do not divide the payload, and a truncated or mis-shaped file must surface as a ValueError naming the
    cloud in wait() (the reference's np.load would raise there too) - not hang the loader thread."""
    import pytest
    from roreg_b200 import scene
    from roreg_b200.test._common import CacheLayout
    ctx = _host_ctx.install(monkeypatch)
    ds = synth.SynthDataset([91, 92], n=150, name="synth/loader", with_fcgf=False)
    cache = str(tmp_path / "c"); ds.write_cache(cache)
    lay = CacheLayout(_cfg(cache), ds, 150)
    for chunk in (1 << 20, 100_003, 4096):                                   # one piece; ragged pieces; many small pieces
        monkeypatch.setattr(scene, "READ_CHUNK", chunk)
        desc, keys, slot = scene.load_scene(ctx, lay, ds, readers=3, ring=2)
        for cid in ds.pc_ids:
            assert np.array_equal(desc[slot[cid]].numpy(), np.load(lay.yoho_desc(cid)))
            assert np.array_equal(keys[slot[cid]].numpy(), ds.get_kps(cid).astype(np.float64))
    victim = lay.yoho_desc(ds.pc_ids[2])
    whole = open(victim, "rb").read()
    open(victim, "wb").write(whole[:len(whole) // 2])                       # truncated payload
    with pytest.raises(ValueError, match="truncated"):
        scene.load_scene(ctx, lay, ds, readers=3, ring=2)
    np.save(victim, np.zeros((149, 32, 60), np.float32))                      # wrong keypoint count
    with pytest.raises(ValueError, match=f"cloud {ds.pc_ids[2]}"):
        scene.load_scene(ctx, lay, ds, readers=3, ring=2)
    open(victim, "wb").write(whole)
    desc, _, slot = scene.load_scene(ctx, lay, ds, readers=1, ring=1)      # and the loader works again afterwards
    assert np.array_equal(desc[slot[ds.pc_ids[2]]].numpy(), np.load(victim))
