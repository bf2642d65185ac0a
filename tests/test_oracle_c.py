"""The C/pthreads restatement (oracle/oracle_c.c, the timed CPU baseline) against the NumPy oracle that is
pinned to the reference's outputs."""
import numpy as np
import pytest
from oracle import roreg_oracle as O
from roreg_b200 import synth


@pytest.fixture(scope="module")
def oc():
    import __graft_entry__
    __graft_entry__.build()
    from oracle import oracle_c
    assert oracle_c.available()
    return oracle_c


def test_c_hot_loops_match_numpy_oracle(oc, tables):
    pr = synth.make_pair(9, n=500)
    f0 = O.inv_pool(pr["feats0"]); f1 = O.inv_pool(pr["feats1"])
    assert np.abs(oc.inv_pool(pr["feats0"]) - f0).max() < 2e-7
    d, i = oc.nn(f1, f0); dr_, ir = O.knn(f1, f0, 1)
    assert np.array_equal(i, ir[:, 0]) and np.abs(d - dr_[:, 0]).max() < 1e-6
    pps, sc = O.mutual_run(pr["feats0"], pr["feats1"])
    dr = oc.des2r(pr["feats1"], pr["feats0"], pps[:, 1], pps[:, 0], tables.perm)
    assert np.array_equal(dr, O.rindex(pr["feats0"], pr["feats1"], pps, tables.perm))
    k0 = pr["keys0"][pps[:, 0]]; k1 = pr["keys1"][pps[:, 1]]
    rng = np.random.default_rng(0)
    H = np.concatenate([np.linalg.qr(rng.standard_normal((40, 3, 3)))[0], rng.standard_normal((40, 3, 1))], 2); H[7] = pr["gt"]
    _, _, ovs = O.oneshot_ransac(k0, k1, sc, H, 0.1)
    assert np.array_equal(oc.score(k0, k1, sc, np.ascontiguousarray(H), 0.1), ovs)
    T, pps2, dr2 = oc.register_pair(pr, tables, 100, 0.1, seed=1)
    assert np.array_equal(pps2, pps) and np.abs(T[:3] - pr["gt"]).max() < 5e-3
