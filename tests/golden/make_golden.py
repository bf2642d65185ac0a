"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference
on CPU through oracle/ref_shim.py) on seeded synthetic scenes.

Run in the build container:   python tests/golden/make_golden.py
The reference's own classes are driven through their public run() methods against a temp cache
directory, so the fixtures pin the complete file-level contract:
  mutual.run          test/matcher.py:49-109
  yohoo.run           test/estimator.py:445-454 (Rindex, Rt_pre, ransac)
  yohoc_ransac.ransac_once   test/estimator.py:163-242  (called in-process: the Pool at :258 forks one
                      worker per pair with a copy of the parent's RNG state, which is not reproducible)
  yoho_des.run / yoho_det.run   test/extractor.py:33-60, test/detector.py:26-47
Networks use random weights written as temporary checkpoints (oracle.random_state_dict) because the
shipped checkpoints are too large to commit and absent on the GPU box; the reference code that
consumes them is unchanged.  Inputs are regenerated from the seeds by roreg_b200.synth.
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
from oracle import roreg_oracle as O  # noqa: E402

SCENES = {
    # name: (pair seeds, n keypoints, keynum, max_iter)
    "s256": ([11, 12], 256, 256, 1000),        # BASELINE config 1: 256 keypoints
    "s700": ([21], 700, 500, 300),             # keynum < N  (shuffle + truncate sampling), max_iter < K
}


def write_ckpt(path, sd):
    import torch
    os.makedirs(os.path.dirname(path), exist_ok=True)
    full = {}
    for k, v in sd.items():
        full[k] = torch.from_numpy(v)
        if k.endswith("running_var"):
            full[k.replace("running_var", "num_batches_tracked")] = torch.tensor(0)
    torch.save({"best_para": 0, "network_state_dict": full}, path)


def main():
    ref_shim.install()                       # chdir(/root/reference), sys.path, shims
    from roreg_b200 import synth, group
    from test.matcher import mutual
    from test.estimator import yohoo, yohoc_ransac, extractor_dr_index
    from test.extractor import yoho_des
    from test.detector import yoho_det

    tb = group.load()
    tmp = tempfile.mkdtemp(prefix="roreg_golden_")
    model_fn = f"{tmp}/ckpt"
    for kind, seed in (("GF", 101), ("ET", 102), ("RD", 103)):
        write_ckpt(f"{model_fn}/{kind}/model_best.pth", O.random_state_dict(kind, seed))
    for name, (seeds, n, keynum, max_iter) in SCENES.items():
        out = {}
        ds = synth.SynthDataset(seeds, n=n, name=f"synth/{name}", max_res_deg=2.0)
        cache = f"{tmp}/cache_{name}"
        ds.write_cache(cache)
        cfg = ref_shim.cfg(output_cache_fn=cache, model_fn=model_fn)
        base = f"{cache}/{ds.name}"
        # ---- per-cloud stages on cloud 0 only (kept small): extractor + detector with random weights
        if name == "s256":
            cfg_pc = ref_shim.cfg(output_cache_fn=f"{tmp}/cache_pc_{name}", model_fn=model_fn)
            ds.write_cache(cfg_pc.output_cache_fn, yoho=False)
            ds_pc = synth.SynthDataset(seeds[:1], n=n, name=f"synth/{name}", max_res_deg=2.0)     # clouds 0,1
            yoho_des(cfg_pc).run(ds_pc)
            yoho_det(cfg_pc).run(ds_pc)
            for cid in ds_pc.pc_ids:
                # first 40 keypoints only: keeps the committed fixture small
                out[f"gf_eqv_{cid}"] = np.load(f"{cfg_pc.output_cache_fn}/{ds.name}/YOHO_Output_Group_feature/{cid}.npy")[:40]
                out[f"det_score_{cid}"] = np.load(f"{cfg_pc.output_cache_fn}/{ds.name}/det_score/{cid}.npy")
        # ---- matcher (global NumPy RNG consumed in the reference's order: H4)
        np.random.seed(1234)
        mutual(cfg).run(ds, keynum)
        # ---- yohoo
        np.random.seed(4321)
        yohoo(cfg).run(ds, keynum, max_iter)
        # ---- yohoc, in-process
        rind = extractor_dr_index(cfg)      # DR_index already on disk from yohoo.run
        yc = yohoc_ransac(cfg)
        os.makedirs(f"{base}/match_{keynum}/yohoc/{max_iter}iters", exist_ok=True)
        for pi, pair in enumerate(ds.pair_ids):
            np.random.seed(777 + pi)
            yc.ransac_once(ds, keynum, max_iter, pair)
        for (id0, id1) in ds.pair_ids:
            m = f"{base}/match_{keynum}"
            out[f"match_{id0}-{id1}"] = np.load(f"{m}/{id0}-{id1}.npy")
            out[f"scores_{id0}-{id1}"] = np.load(f"{m}/scores/{id0}-{id1}.npy")
            out[f"dr_index_{id0}-{id1}"] = np.load(f"{m}/DR_index/{id0}-{id1}.npy")
            out[f"trans_pre_{id0}-{id1}"] = np.load(f"{m}/Trans_pre/{id0}-{id1}.npy")
            z = np.load(f"{m}/yohoo/{max_iter}iters/{id0}-{id1}.npz")
            out[f"yohoo_trans_{id0}-{id1}"] = z["trans"]; out[f"yohoo_recall_{id0}-{id1}"] = z["recalltime"]
            z = np.load(f"{m}/yohoc/{max_iter}iters/{id0}-{id1}.npz")
            out[f"yohoc_trans_{id0}-{id1}"] = z["trans"]; out[f"yohoc_recall_{id0}-{id1}"] = z["recalltime"]
        out["pre_log_yohoo"] = np.frombuffer(open(f"{base}/match_{keynum}/yohoo/{max_iter}iters/pre.log", "rb").read(), dtype=np.uint8)
        out["meta"] = np.array([n, keynum, max_iter] + list(seeds))
        np.savez_compressed(f"{HERE}/{name}.npz", **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith(("match", "yohoo_recall", "yohoc_recall"))},
              os.path.getsize(f"{HERE}/{name}.npz"), "bytes")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
