"""Generate tests/golden/s400rdrm.npz: the UNMODIFIED reference (imported from /root/reference on CPU through
oracle/ref_shim.py) run with --RD (NMS keypoint sampling on detector scores, test/matcher.py:11-42,76-80) and --RM
(estimators keep the top `match_n` fraction of the matches by score, test/estimator.py:195-203,415-421).

Run in the build container:   python tests/golden/make_golden_rd_rm.py
Inputs are regenerated from the seeds by roreg_b200.synth; the detector scores and the per-match scores (which --RM runs get
from the rotation-coherence matcher) are seeded random arrays stored in the fixture.
"""
import os
import sys
import shutil
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, REPO)
from oracle import ref_shim  # noqa: E402

NAME, SEEDS, N, KEYNUM, MAX_ITER = "s400rdrm", [31, 32], 400, 300, 200


def main():
    ref_shim.install()
    from roreg_b200 import synth
    from test.matcher import mutual, NMS_sample
    from test.estimator import yohoo_ransac, yohoc_ransac, extractor_dr_index
    tmp = tempfile.mkdtemp(prefix="roreg_golden_rdrm_")
    out = {}
    ds = synth.SynthDataset(SEEDS, n=N, name=f"synth/{NAME}", max_res_deg=2.0)
    cache = f"{tmp}/cache"
    ds.write_cache(cache)
    rng = np.random.RandomState(99)
    os.makedirs(f"{cache}/{ds.name}/det_score", exist_ok=True)
    for cid in ds.pc_ids:                      # rank-normalised saliency as test/detector.py:44-46 writes it
        sc = rng.permutation(N) / N
        np.save(f"{cache}/{ds.name}/det_score/{cid}.npy", sc)
        out[f"det_score_{cid}"] = sc
    # ---- the sampler alone, on the three branches of :27-39 (too many maxima / top-up / fewer points than requested)
    keys = ds.get_kps(ds.pc_ids[0]); sc0 = out[f"det_score_{ds.pc_ids[0]}"]
    for num in (40, 300, 380, 500):
        out[f"nms_{num}"] = NMS_sample(num, 5).sample(keys, sc0)
    flat = np.round(sc0 * 8) / 8               # heavy ties in the scores
    out["nms_flat_scores"] = flat
    for num in (40, 300):
        out[f"nms_flat_{num}"] = NMS_sample(num, 5).sample(keys, flat)
    # ---- matcher with --RD
    cfg = ref_shim.cfg(output_cache_fn=cache, RD=True, RM=True, match_n=0.5)
    np.random.seed(2468)
    mutual(cfg).run(ds, KEYNUM)
    base = f"{cache}/{ds.name}/match_{KEYNUM}"
    for (id0, id1) in ds.pair_ids:             # float32 scores as yoho_mat writes them (test/matcher.py:210)
        m = np.load(f"{base}/{id0}-{id1}.npy")
        s = rng.rand(m.shape[0]).astype(np.float32)
        np.save(f"{base}/scores/{id0}-{id1}.npy", s)
        out[f"match_{id0}-{id1}"] = m; out[f"scores_{id0}-{id1}"] = s
    extractor_dr_index(cfg).Rindex(ds, KEYNUM)
    # ---- yohoo one-shot RANSAC with --RM on seeded hypotheses (the ET network is pinned by make_golden.py)
    os.makedirs(f"{base}/Trans_pre", exist_ok=True)
    for pi, (id0, id1) in enumerate(ds.pair_ids):
        m = out[f"match_{id0}-{id1}"]
        gt = ds.get_transform(id0, id1).astype(np.float64)
        Tr = np.tile(gt[None], (m.shape[0], 1, 1))
        Tr[:, :, 3] += rng.normal(0, 0.05, (m.shape[0], 3))
        Tr[::3, :, :3] = np.linalg.qr(rng.normal(size=(len(Tr[::3]), 3, 3)))[0]
        np.save(f"{base}/Trans_pre/{id0}-{id1}.npy", Tr)
        out[f"trans_pre_{id0}-{id1}"] = Tr
    np.random.seed(1357)
    yohoo_ransac(cfg).ransac(ds, KEYNUM, MAX_ITER)
    # ---- yohoc with --RM, in-process (see make_golden.py)
    yc = yohoc_ransac(cfg)
    os.makedirs(f"{base}/yohoc/{MAX_ITER}iters", exist_ok=True)
    for pi, pair in enumerate(ds.pair_ids):
        np.random.seed(555 + pi)
        yc.ransac_once(ds, KEYNUM, MAX_ITER, pair)
    for (id0, id1) in ds.pair_ids:
        out[f"dr_index_{id0}-{id1}"] = np.load(f"{base}/DR_index/{id0}-{id1}.npy")
        z = np.load(f"{base}/yohoo/{MAX_ITER}iters/{id0}-{id1}.npz")
        out[f"yohoo_trans_{id0}-{id1}"] = z["trans"]; out[f"yohoo_recall_{id0}-{id1}"] = z["recalltime"]
        z = np.load(f"{base}/yohoc/{MAX_ITER}iters/{id0}-{id1}.npz")
        out[f"yohoc_trans_{id0}-{id1}"] = z["trans"]; out[f"yohoc_recall_{id0}-{id1}"] = z["recalltime"]
    # ---- host helpers of yohoc_ransac on their own (test/estimator.py:119-147)
    dr = out[f"dr_index_{ds.pair_ids[0][0]}-{ds.pair_ids[0][1]}"]
    stat, prob = yc.DR_statictic(dr)
    out["drstat_prob"] = prob
    out["drstat_members"] = np.concatenate([np.array(stat[i], np.int64) for i in range(60)])
    out["drstat_counts"] = np.array([len(stat[i]) for i in range(60)])
    k0 = ds.get_kps(ds.pair_ids[0][0]); k1 = ds.get_kps(ds.pair_ids[0][1])
    trip = rng.randint(0, N, (16, 3))
    out["kabsch_triplets"] = trip
    out["kabsch_T"] = np.stack([yc.Threepps2Tran(k0[t], k1[t]) for t in trip])
    out["meta"] = np.array([N, KEYNUM, MAX_ITER] + SEEDS)
    np.savez_compressed(f"{HERE}/{NAME}.npz", **out)
    print(NAME, os.path.getsize(f"{HERE}/{NAME}.npz"), "bytes", {k: v.shape for k, v in out.items() if k.startswith(("nms", "match"))})
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
