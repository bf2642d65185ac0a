"""Golden fixture for the rotation-coherence matcher: run the UNMODIFIED reference `yoho_mat.run`
(test/matcher.py:111-210, network/rot_coh_match.py) on CPU on a seeded synthetic scene with seeded random
RM weights (oracle.random_state_dict('RM', 104), written as a temporary checkpoint) and record the files it
writes.   python tests/golden/make_golden_rm.py   (build container only; needs /root/reference)"""
import os
import shutil
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, REPO); sys.path.insert(0, HERE)
from oracle import ref_shim  # noqa: E402
from oracle import roreg_oracle as O  # noqa: E402
from make_golden import write_ckpt  # noqa: E402


def main():
    ref_shim.install()
    from roreg_b200 import synth
    from test.matcher import yoho_mat
    tmp = tempfile.mkdtemp(prefix="roreg_golden_rm_")
    write_ckpt(f"{tmp}/ckpt/RM/model_best.pth", O.random_state_dict("RM", 104))
    seeds, n, keynum = [31, 32], 300, 256
    ds = synth.SynthDataset(seeds, n=n, name="synth/rm", with_fcgf=False)
    cache = f"{tmp}/cache"; ds.write_cache(cache)
    cfg = ref_shim.cfg(output_cache_fn=cache, model_fn=f"{tmp}/ckpt", RM=True)
    np.random.seed(2468)
    yoho_mat(cfg).run(ds, keynum)
    out = {"meta": np.array([n, keynum, 0] + seeds)}
    for (id0, id1) in ds.pair_ids:
        out[f"match_{id0}-{id1}"] = np.load(f"{cache}/synth/rm/match_{keynum}/{id0}-{id1}.npy")
        out[f"scores_{id0}-{id1}"] = np.load(f"{cache}/synth/rm/match_{keynum}/scores/{id0}-{id1}.npy")
        print(id0, id1, out[f"match_{id0}-{id1}"].shape, out[f"scores_{id0}-{id1}"].dtype)
    np.savez_compressed(f"{HERE}/rm300.npz", **out)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
