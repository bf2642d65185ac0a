"""TEST / MEASUREMENT INFRASTRUCTURE ONLY - never imported by the product.

The reference's per-pair CPU path restated WITH THE REFERENCE'S OWN TENSOR OPERATIONS (torch on CPU, all intra-op
threads), for the `cpu_baseline` / `--impl reference` figure of bench.py: the C/pthreads port (oracle_c.c) is an optimised
re-implementation and therefore a *stronger* baseline than what a RoReg user actually runs (reported beside it as
`optimised_c_port`); this module times the operations the reference executes when CUDA is absent:

  * matcher  test/matcher.py:69-72,94-106 + utils/knn_search.py:17-66,138-162: mean over the group axis, L2 normalise,
    `pdist` = sqrt(sum((A[:,None]-B[None])**2, 2) + 1e-7) in 500-row chunks with `.min(dim=1)` per chunk, Python mutual loop;
  * Des2R    test/estimator.py:85-89,105-110: advanced-index gather `X[:, :, P.reshape(-1)]` to [K,32,60,60] then
    `einsum('bfag,bfg->ba')` and argmax;
  * yohoc RANSAC + refinement: oracle/roreg_oracle.py (NumPy, as the reference's estimator is NumPy).

Results are checked against oracle/roreg_oracle.py in tests/test_oracle_golden.py::test_torch_mirror_equals_numpy_oracle.
"""
import numpy as np
import torch

from . import roreg_oracle as O


def inv_pool(feats):
    """test/matcher.py:69-72 (np.mean over the last axis, x / (norm + 1e-5))."""
    f = np.mean(feats, axis=-1)
    return (f / (np.sqrt(np.sum(np.square(f), axis=1, keepdims=True)) + 1e-5)).astype(np.float32)


def find_nn(source, target, nn_max_n=500):
    """utils/knn_search.py:26-66 (dist_type 'L2' as modified_knn_matcher.__call__ passes it, :141,147-151)."""
    F0 = torch.from_numpy(source); F1 = torch.from_numpy(target)
    N = F0.shape[0]
    inds = []
    for i in range(int(np.ceil(N / nn_max_n))):
        A = F0[i * nn_max_n:(i + 1) * nn_max_n]
        D2 = torch.sum((A.unsqueeze(1) - F1.unsqueeze(0)).pow(2), 2)
        dist = torch.sqrt(D2 + 1e-7)
        _, ind = dist.min(dim=1)
        inds.append(ind)
    return torch.cat(inds).numpy()


def mutual_run(feats0, feats1):
    """mutual.run test/matcher.py:66-109 with identity sampling: match_pps [K,2], scores = ones."""
    f0 = inv_pool(feats0); f1 = inv_pool(feats1)
    idx01 = find_nn(f0, f1)                       # KNN(feats1, feats0): NN of every cloud-0 row in cloud 1 (:94-95)
    idx10 = find_nn(f1, f0)                       # (:96-97)
    pps = []
    for i in range(idx01.shape[0]):               # (:98-105)
        if idx10[idx01[i]] == i:
            pps.append([i, idx01[i]])
    pps = np.array(pps, dtype=np.int64).reshape(-1, 2)
    return pps, np.ones(pps.shape[0])


def rindex(feats0, feats1, match_pps, perm, chunk=1000):
    """extractor_dr_index.Rindex / Batch_Des2R_torch test/estimator.py:85-89,105-110 (X = cloud id1, Y = cloud id0).
    The reference materialises [K,32,60,60] in one go (2.3 GB at K = 5000); chunked here over K, same operations."""
    nei = torch.from_numpy(perm.reshape(-1).astype(np.int64))
    X = torch.from_numpy(feats1[match_pps[:, 1]].astype(np.float32))
    Y = torch.from_numpy(feats0[match_pps[:, 0]].astype(np.float32))
    out = []
    for s in range(0, X.shape[0], chunk):
        x = X[s:s + chunk]; y = Y[s:s + chunk]
        B, F, G = x.shape
        xg = x[:, :, nei].reshape([B, F, 60, 60])
        cor = torch.einsum('bfag,bfg->ba', xg, y)
        out.append(torch.argmax(cor, dim=1))
    return torch.cat(out).numpy() if out else np.zeros((0,), np.int64)


def register_pair(pr, perm, max_iter, ird, seed):
    """One pair through the reference's default CLI pipeline (mutual + yohoc), reference tensor operations."""
    pps, sc = mutual_run(pr["feats0"], pr["feats1"])
    dr = rindex(pr["feats0"], pr["feats1"], pps, perm)
    k0 = pr["keys0"][pps[:, 0]]; k1 = pr["keys1"][pps[:, 1]]
    T, recall, _ = O.yohoc_ransac(k0, k1, sc, dr, ird, max_iter, rng=np.random.RandomState(seed))
    return T, pps, dr
