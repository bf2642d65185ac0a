"""Subprocess driver of tests/test_evaluator_dropin.py (TEST INFRASTRUCTURE; needs /root/reference, build container only).

Runs the UNMODIFIED `yoho_evaluator` of the reference (test/evaluator.py:13-101: __init__, process_scene, fmr_ir_scene,
rr_scene) on a seeded synthetic scene, imported from the tree named by ROREG_REFERENCE_ROOT:
  --arm ref    the reference itself (its own plugins on the CPU through the shims of oracle/ref_shim.py);
  --arm b200   an overlay of the reference in which ONLY test/__init__.py is replaced by the stub of INTEGRATION.md, so that
               `from test import name2extractor, ...` (test/evaluator.py:11) resolves to roreg_b200.test.  There is no GPU in the
               build container, so the plugins' device context is the oracle-backed one of tests/_host_ctx.py: what this pins is
               the drop-in claim itself - registries, constructor / run() signatures, cfg fields, the dataset duck type, the file
               contract the evaluator's metric code reads, the global-RNG order - not the kernels (tests/test_gpu_*.py do that).
Writes the metrics and every per-pair file's content to --out (npz)."""
import argparse
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)


class _Patch:
    """monkeypatch.setattr stand-in."""

    def setattr(self, obj, name, value, raising=True):
        if raising and not hasattr(obj, name):
            raise AttributeError(name)
        setattr(obj, name, value)


def _install_nibabel_stub():
    """`nibabel.quaternions.mat2quat` (used by utils/RR_cal.py:61 only, inside RR_cal.benchmark, which needs the datasets' gt.info
    files and is not called here): the published algorithm (Bar-Itzhack 2000 - largest eigenvector of the symmetric K matrix,
    w >= 0), so the import at the top of the reference's evaluator resolves to something correct."""
    import numpy as np

    def mat2quat(M):
        Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = np.asarray(M, float).flat
        K = np.array([[Qxx - Qyy - Qzz, 0, 0, 0], [Qyx + Qxy, Qyy - Qxx - Qzz, 0, 0], [Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy, 0],
                      [Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz]]) / 3.0
        vals, vecs = np.linalg.eigh(K)
        q = vecs[[3, 0, 1, 2], np.argmax(vals)]
        return -q if q[0] < 0 else q
    nib = types.ModuleType("nibabel"); nq = types.ModuleType("nibabel.quaternions")
    nq.mat2quat = mat2quat; nib.quaternions = nq
    sys.modules.setdefault("nibabel", nib); sys.modules.setdefault("nibabel.quaternions", nq)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", required=True, choices=["ref", "b200"])
    ap.add_argument("--cache", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--n", type=int, default=400)
    ap.add_argument("--keynum", type=int, default=300)
    ap.add_argument("--max-iter", type=int, default=300)
    ap.add_argument("--rd", type=int, default=0)
    ap.add_argument("--seeds", default="31,32,33")
    a = ap.parse_args()
    import numpy as np
    from roreg_b200 import synth
    from oracle import ref_shim
    from oracle import roreg_oracle as O
    seeds = [int(s) for s in a.seeds.split(",")]
    ds = synth.SynthDataset(seeds, n=a.n, name="synth/eval", max_res_deg=2.0, with_fcgf=False)
    ds.write_cache(a.cache)
    if a.rd:                                       # --RD: detector scores are cached per cloud (test/detector.py:36 skips existing files)
        os.makedirs(f"{a.cache}/{ds.name}/det_score", exist_ok=True)
        for cid in ds.pc_ids:
            np.save(f"{a.cache}/{ds.name}/det_score/{cid}.npy", np.random.default_rng(100 + int(cid)).random(a.n))
    # checkpoints the constructors insist on (random weights; the descriptors / scores they would produce are cached already)
    import torch
    model_fn = f"{a.cache}/ckpt"
    for kind, seed in (("GF", 101), ("RD", 103)):
        sd = O.random_state_dict(kind, seed)
        full = {}
        for k, v in sd.items():
            full[k] = torch.from_numpy(v)
            if k.endswith("running_var"):
                full[k.replace("running_var", "num_batches_tracked")] = torch.tensor(0)
        os.makedirs(f"{model_fn}/{kind}", exist_ok=True)
        torch.save({"best_para": 0, "network_state_dict": full}, f"{model_fn}/{kind}/model_best.pth")
    ref_shim.install()                             # chdir + sys.path of ROREG_REFERENCE_ROOT, np.int / .cuda() / open3d shims
    _install_nibabel_stub()                        # utils/RR_cal.py:10 imports it at module level (absent from this image)
    _orig_array = np.array                         # utils/r_eval.py:42 passes copy=False with NumPy-1.x meaning ("copy only if needed")

    def _array(*args, **kw):
        if kw.get("copy") is False:
            kw["copy"] = None
        return _orig_array(*args, **kw)
    np.array = _array
    if a.arm == "b200":
        import _host_ctx
        _host_ctx.install(_Patch()); _host_ctx.install_nets(_Patch())
    from test.evaluator import yoho_evaluator      # the reference's file, unmodified, in both arms
    import test as test_pkg
    cfg = ref_shim.cfg(output_cache_fn=a.cache, model_fn=model_fn, keynum=a.keynum, max_iter=a.max_iter, ET="yohoc", RD=bool(a.rd),
                       RM=False, SO3_related_files=f"{ref_shim.REF_ROOT}/utils/group_related")
    np.random.seed(20240)
    ev = yoho_evaluator(cfg)
    ev.process_scene(ds)
    fmr, ir = ev.fmr_ir_scene(ds)
    rr, rre, rte = ev.rr_scene(ds)
    out = dict(fmr=fmr, ir=ir, rr=rr, rre=rre, rte=rte, plugin_module=np.array(type(ev.matcher).__module__),
               registry_module=np.array(test_pkg.name2matcher["matmul"].__module__), rng_after=np.random.get_state()[1][:8])
    base = f"{a.cache}/{ds.name}/match_{a.keynum}"
    for id0, id1 in ds.pair_ids:
        out[f"match_{id0}-{id1}"] = np.load(f"{base}/{id0}-{id1}.npy")
        out[f"scores_{id0}-{id1}"] = np.load(f"{base}/scores/{id0}-{id1}.npy")
        out[f"dr_{id0}-{id1}"] = np.load(f"{base}/DR_index/{id0}-{id1}.npy")
        z = np.load(f"{base}/yohoc/{a.max_iter}iters/{id0}-{id1}.npz")
        out[f"trans_{id0}-{id1}"] = z["trans"]; out[f"recall_{id0}-{id1}"] = z["recalltime"]
        out[f"gt_{id0}-{id1}"] = ds.get_transform(id0, id1)
    out["pre_log"] = np.frombuffer(open(f"{base}/yohoc/{a.max_iter}iters/pre.log", "rb").read(), dtype=np.uint8)
    np.savez(a.out, **out)


if __name__ == "__main__":
    main()
