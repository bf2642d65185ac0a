"""CPU restatement (NumPy) of RoReg's per-pair registration hot path.

TEST INFRASTRUCTURE ONLY.  This module is the parity checker for the CUDA path in
roreg_b200/csrc; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.  The product (roreg_b200/*) never does, and fails
loudly when the CUDA library is missing.

Every function cites the reference file:line it restates (paths relative to the
reference root).  PINNING: the reference has no tests or golden vectors of its own
(SURVEY.md section 4), so this restatement is pinned the other way the task allows -
against outputs of the UNMODIFIED reference imported on CPU in the build container
(oracle/ref_shim.py) - see tests/golden/make_golden.py, whose fixtures are committed and
re-checked by tests/test_oracle_golden.py, and tests/test_oracle_vs_reference.py which
compares live whenever /root/reference is present.

dtype policy mirrors the reference: descriptors and matcher arithmetic float32,
RANSAC / Kabsch geometry float64.
"""
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------
# a13  mutual matcher front end: invariant pooling + L2 normalisation
# ----------------------------------------------------------------------------------------
def inv_pool(feats, normalise=True):
    """test/matcher.py:69-72 - mean over the 60 group elements, then x/(||x||+1e-5).
    With normalise=False it is the plain mean of network/rot_coh_match.py:346-347."""
    f = np.mean(feats, axis=-1).astype(F32)
    if normalise:
        f = f / (np.sqrt(np.sum(np.square(f), axis=1, keepdims=True)) + F32(1e-5))
    return f.astype(F32)


# ----------------------------------------------------------------------------------------
# a15  brute-force (k-)NN, utils/knn_search.py
# ----------------------------------------------------------------------------------------
def pdist_l2(A, B):
    """utils/knn_search.py:17-21 - difference form, sqrt(D2 + 1e-7), float32."""
    D2 = np.sum(np.square(A[:, None, :] - B[None, :, :]), axis=2, dtype=F32)
    return np.sqrt(D2 + F32(1e-7))


def knn(target, source, k=1, chunk=500):
    """modified_knn_matcher.__call__ utils/knn_search.py:138-162 with target [n,f], source [m,f]
    (already transposed).  For each SOURCE row the k nearest TARGET rows.
    k=1 -> min (first minimal index, torch.min semantics, :42); k>1 -> topk(-d) (:85).
    Returns d [m,k] float32, idx [m,k] int64."""
    m = source.shape[0]
    ds, ids = [], []
    for s in range(0, m, chunk):
        d = pdist_l2(source[s:s + chunk].astype(F32), target.astype(F32))
        if k == 1:
            i = np.argmin(d, axis=1)
            ds.append(d[np.arange(d.shape[0]), i][:, None]); ids.append(i[:, None])
        else:
            # torch.topk(-d): largest -d first; ties are broken towards the lower index here
            i = np.argsort(d, axis=1, kind="stable")[:, :k]
            ds.append(np.take_along_axis(d, i, 1)); ids.append(i)
    return np.concatenate(ds, 0), np.concatenate(ids, 0).astype(np.int64)


def knn_f64(target, source):
    """float64 adjudicator for near ties: best, second-best squared distance
    and the argmin, all in float64.  Not a reference function."""
    t = target.astype(np.float64); s = source.astype(np.float64)
    D2 = (s * s).sum(1)[:, None] + (t * t).sum(1)[None, :] - 2.0 * s @ t.T
    part = np.partition(D2, 1, axis=1)
    return np.argmin(D2, axis=1), part[:, 0], part[:, 1]


def mutual_matches(f0, f1):
    """test/matcher.py:94-106 - 1-NN both ways on the (sampled) invariant features and the
    Python mutual-check loop; rows come out in increasing index of cloud 0."""
    _, nn01 = knn(f1, f0, 1)          # KNN(feats1, feats0): for each row of cloud0 its NN in cloud1
    _, nn10 = knn(f0, f1, 1)
    nn01 = nn01[:, 0]; nn10 = nn10[:, 0]
    i = np.arange(f0.shape[0])
    keep = nn10[nn01] == i
    return np.stack([i[keep], nn01[keep]], 1).astype(np.int64), nn01, nn10


def mutual_run(feats0, feats1, sample0=None, sample1=None):
    """mutual.run test/matcher.py:66-109 for one pair, given the sample index arrays
    (identity when None).  Returns match_pps [K,2] int64 in ORIGINAL keypoint indices
    (col0 -> cloud id0, col1 -> cloud id1) and scores = ones (float64, :109)."""
    f0 = inv_pool(feats0); f1 = inv_pool(feats1)
    if sample0 is None: sample0 = np.arange(f0.shape[0])
    if sample1 is None: sample1 = np.arange(f1.shape[0])
    pps, _, _ = mutual_matches(f0[sample0], f1[sample1])
    out = np.stack([sample0[pps[:, 0]], sample1[pps[:, 1]]], 1).astype(np.int64)
    return out, np.ones(out.shape[0])


# ----------------------------------------------------------------------------------------
# a14  NMS keypoint sampler
# ----------------------------------------------------------------------------------------
def nms_sample(keys, scores, num, k=5):
    """NMS_sample.sample test/matcher.py:18-42."""
    if keys.shape[0] < num:
        return np.arange(keys.shape[0])
    kf = keys.astype(F32)
    _, nn = knn(kf, kf, k)
    nei_max = np.max(scores[nn.reshape(-1)].reshape(-1, k), axis=-1)
    sam = np.where(scores >= nei_max)[0]
    if sam.shape[0] > num:
        ss = scores[sam]; ss = ss / np.sum(ss)
        sam = sam[np.argsort(ss)[-num:]]
    if sam.shape[0] < num:
        left = num - sam.shape[0]
        idx_left = np.where(scores < nei_max)[0]
        sam = np.concatenate([sam, idx_left[np.argsort(scores[idx_left])[-left:]]], 0)
    return sam


# ----------------------------------------------------------------------------------------
# a4 / a5  equivariant correlation
# ----------------------------------------------------------------------------------------
def group_corr_v1(X, Y, perm, dtype=F32):
    """Batch_Des2R_torch test/estimator.py:85-89 (== GF_train.Des2DR network/group_feat.py:55-58):
        cor[b,a] = sum_{f,g} X[b,f,P[a,g]] * Y[b,f,g]
    evaluated as the 60x60 Gram G[h,g] = sum_f X[f,h] Y[f,g] followed by the generalised
    diagonal sums cor[a] = sum_g G[P[a,g], g] (same terms, fewer temporaries)."""
    G = np.einsum("bfh,bfg->bhg", X.astype(dtype), Y.astype(dtype))
    g = np.arange(60)
    return np.stack([G[:, perm[a], g].sum(1) for a in range(60)], 1)


def group_corr_v2(S, T, perm, dtype=F32):
    """R-indicator network/rot_coh_match.py:158-163, index convention P[g,h] (summed index first):
        Rind[b,h] = sum_{f,g} S[b,f,P[g,h]] * T[b,f,g]
    s2t (:159-160): S = source descriptor, T = descriptor of its 1-NN in the target cloud.
    t2s (:162-163): S = descriptor of the 1-NN (called target_eqv there), T = the point's own."""
    G = np.einsum("bfh,bfg->bhg", S.astype(dtype), T.astype(dtype))
    g = np.arange(60)
    return np.stack([G[:, perm[:, h], g].sum(1) for h in range(60)], 1)


def des2r(X, Y, perm):
    """argmax_a of group_corr_v1 (first maximal index, torch.argmax)."""
    return np.argmax(group_corr_v1(X, Y, perm), axis=1).astype(np.int64)


def rindex(feats0, feats1, match_pps, perm):
    """extractor_dr_index.Rindex test/estimator.py:105-111: X = feats of cloud id1, Y = id0."""
    return des2r(feats1[match_pps[:, 1]], feats0[match_pps[:, 0]], perm)


# ----------------------------------------------------------------------------------------
# a17  per-match SE(3) hypotheses from (quaternion, coarse rotation)
# ----------------------------------------------------------------------------------------
def matrix_from_quaternion(q):
    """utils/r_eval.py:90-106, (w,x,y,z) order, no normalisation, float64 output."""
    q = np.asarray(q)
    w, x, y, z = q[0], q[1], q[2], q[3]       # products are formed in q's dtype (float32 in the reference)
    m = np.eye(3)
    m[0, 0] = 1 - 2 * y * y - 2 * z * z; m[0, 1] = 2 * x * y - 2 * z * w; m[0, 2] = 2 * x * z + 2 * y * w
    m[1, 0] = 2 * x * y + 2 * z * w; m[1, 1] = 1 - 2 * x * x - 2 * z * z; m[1, 2] = 2 * y * z - 2 * x * w
    m[2, 0] = 2 * x * z - 2 * y * w; m[2, 1] = 2 * y * z + 2 * x * w; m[2, 2] = 1 - 2 * x * x - 2 * y * y
    return m


def hypotheses_from_quat(quat, pre_idx, keys0_m, keys1_m, rot):
    """extractor_localtrans.Rt_pre test/estimator.py:349-366.
    quat [K,4] float32 (unit), pre_idx [K], keys*_m [K,3] float64, rot = Rotation.npy.
    Rgroup is cast to float32 (:285) and the product R_residual(float64 holding float32-formed
    values) @ R_anchor(float32) is float64; t = key0 - key1 @ R.T.  Returns Trans [K,3,4] float64."""
    rot32 = rot.astype(F32)
    out = np.empty((quat.shape[0], 3, 4))
    for i in range(quat.shape[0]):
        R = matrix_from_quaternion(quat[i]) @ rot32[int(pre_idx[i])]
        out[i, :, :3] = R
        out[i, :, 3] = keys0_m[i] - keys1_m[i] @ R.T
    return out


# ----------------------------------------------------------------------------------------
# a18 / a19  one-shot RANSAC scoring and the weighted-Kabsch refiner
# ----------------------------------------------------------------------------------------
def transform_points(pts, T):
    """utils/utils.py:38-46."""
    T = np.asarray(T)
    if T.shape == (3, 3): return pts @ T.T
    if T.shape == (3, 4): return pts @ T[:, :3].T + T[:, 3:].T
    h = np.concatenate([pts, np.ones((pts.shape[0], 1))], 1) @ T.T
    return h[:, :-1] / h[:, -1:]


def inlier_mask(k0, k1, T, ird):
    d = np.sum(np.square(k0 - transform_points(k1, T)), axis=-1)
    return d < ird * ird


def overlap_cal(k0, k1, T, scores, ird):
    """yohoo_ransac.overlap_cal test/estimator.py:377-382 (== yohoc_ransac.overlap_cal :149-154)."""
    ov = np.where(inlier_mask(k0, k1, T, ird))[0]
    return np.sum(scores[ov]) / scores.shape[0]


def oneshot_ransac(k0, k1, scores, trans_ordered, ird):
    """yohoo_ransac.ransac test/estimator.py:426-436: score every hypothesis on ALL matches,
    strict '>' keeps the first best.  Returns (best index or -1, best overlap, overlaps[H])."""
    best, best_ov = -1, 0
    ovs = np.zeros(trans_ordered.shape[0])
    for t in range(trans_ordered.shape[0]):
        ov = overlap_cal(k0, k1, trans_ordered[t], scores, ird)
        ovs[t] = ov
        if ov > best_ov:
            best_ov, best = ov, t
    return best, best_ov, ovs


def refine_once(k0, k1, T, scores, radius):
    """refiner.Refine_trans test/estimator.py:60-72 with SVDR_w :39-43 (R = U @ VT, no
    reflection fix), weights = scores/sum(scores) of the inliers (:54)."""
    m = inlier_mask(k0, k1, T, radius)
    s = scores[m]                                   # keeps the scores' dtype (float32 from yoho_mat, float64 ones from mutual)
    a0 = k0[m]; a1 = k1[m]
    w = s / np.sum(s)
    c0 = np.sum(a0 * w[:, None], axis=0); c1 = np.sum(a1 * w[:, None], axis=0)
    H = (a0 - c0[None]).T @ ((a1 - c1[None]) * w[:, None])     # == afterrot.T @ diag(w) @ beforerot
    U, _, VT = np.linalg.svd(H)
    R = U @ VT
    Tn = np.eye(4); Tn[:3, :3] = R; Tn[:3, 3] = c0 - c1 @ R.T
    return Tn


def refine(k0, k1, T, scores, ird):
    """test/estimator.py:438-439 / :240-241: radius 2*ird, then ird."""
    T = refine_once(k0, k1, T, scores, ird * 2.0)
    return refine_once(k0, k1, T, scores, ird)


def select_hypotheses(scores, n_hyp, RM, match_n):
    """test/estimator.py:415-421 - with --RM keep the hypotheses of the top `match_n`
    fraction by score (ascending argsort tail); otherwise all."""
    if RM:
        num = max(scores.shape[0] * match_n, 10) if match_n < 0.999 else match_n
        return np.argsort(scores)[-int(num):]
    return np.arange(n_hyp)


def yohoo_ransac(k0_m, k1_m, scores, trans, ird, max_iter, RM=False, match_n=0.5, rng=np.random):
    """yohoo_ransac.ransac test/estimator.py:404-441 for one pair.  Consumes rng exactly as the
    reference does (one shuffle of arange(H), :424)."""
    T = trans
    if RM:
        T = T[select_hypotheses(scores, T.shape[0], True, match_n)]
    index = np.arange(T.shape[0]); rng.shuffle(index)
    Tr = T[index[0:max_iter]]
    best, best_ov, ovs = oneshot_ransac(k0_m, k1_m, scores, Tr, ird)
    Tb = Tr[best] if best >= 0 else 0
    if best < 0:
        raise ValueError("no hypothesis scored > 0 (the reference then fails in transform_points)")
    return refine(k0_m, k1_m, Tb, scores, ird), best, dict(order=index[0:max_iter], overlaps=ovs, best_overlap=best_ov)


# ----------------------------------------------------------------------------------------
# a20  coarse-rotation-guided RANSAC (yohoc)
# ----------------------------------------------------------------------------------------
def dr_statistic(dr_index):
    """yohoc_ransac.DR_statictic test/estimator.py:119-137."""
    buckets = {i: [] for i in range(60)}
    for t in range(dr_index.shape[0]):
        buckets[int(dr_index[t])].append(t)
    prob = []
    for i in range(60):
        if len(buckets[i]) < 2:
            prob.append(0)
        else:
            num = float(len(buckets[i])) / 100.0
            prob.append(num * (num - 0.01) * (num - 0.02))
    prob = np.array(prob)
    if np.sum(prob) == 0:
        return None, np.zeros(60)
    return buckets, prob / np.sum(prob)


def threepps2tran(kps0, kps1):
    """yohoc_ransac.Threepps2Tran test/estimator.py:139-147 (rotation = VT.T @ U.T, no det fix).
    NOTE the 3-point cross-covariance has rank 2, so the sign of the third singular pair - and
    with it whether `rotation` is proper or a reflection - is decided by LAPACK rounding noise
    (measured: 50.4 % reflections on random triplets).  Parity for this function therefore means
    'given the same LAPACK'; the CUDA path's parity mode takes these 3x4 from the host for that reason."""
    c0 = np.mean(kps0, 0, keepdims=True); c1 = np.mean(kps1, 0, keepdims=True)
    m = (kps1 - c1).T @ (kps0 - c0)
    U, S, VT = np.linalg.svd(m)
    R = VT.T @ U.T
    return np.concatenate([R, (c0 - c1 @ R.T).T], 1)


def yohoc_draws(dr_index_sel, max_iter, rng=np.random, max_time=50000):
    """The RNG-consuming part of yohoc_ransac.ransac_once test/estimator.py:221-228, factored out:
    returns the list of (rotation id, idxs_init[3]) in iteration order."""
    buckets, prob = dr_statistic(dr_index_sel)
    draws = []
    if np.sum(prob) < 1e-5:
        return draws, buckets, prob
    it = 0; ex = 0
    while it < max_iter:
        if ex > max_time: break
        ex += 1
        r = rng.choice(range(60), p=prob)
        if len(buckets[r]) < 2:
            continue
        it += 1
        draws.append((int(r), rng.choice(np.array(buckets[r]), 3)))
    return draws, buckets, prob


def yohoc_ransac(k0_init, k1_init, scores, dr_index, ird, max_iter, RM=False, match_n=0.5, rng=np.random):
    """yohoc_ransac.ransac_once test/estimator.py:181-242 for one pair (file I/O stripped).
    Triplets are drawn from the (optionally top-`match_n`) subset, overlap is scored on ALL matches."""
    sel = select_hypotheses(scores, k0_init.shape[0], RM, match_n)
    k0 = k0_init[sel]; k1 = k1_init[sel]
    draws, _, prob = yohoc_draws(dr_index[sel], max_iter, rng)
    if np.sum(prob) < 1e-5:
        return None, 50000, dict(degenerate=True)          # reference writes np.random.rand(4,4) (:216-218)
    best_ov, best_T, recall = 0, np.ones(4), 0
    hyps = []
    for it, (r, idx) in enumerate(draws, start=1):
        T = threepps2tran(k0[idx], k1[idx])
        hyps.append(T)
        ov = overlap_cal(k0_init, k1_init, T, scores, ird)
        if ov > best_ov:
            best_ov, best_T, recall = ov, T, it
    return refine(k0_init, k1_init, best_T, scores, ird), recall, dict(hyps=np.array(hyps), best_overlap=best_ov, draws=draws)


# ----------------------------------------------------------------------------------------
# a1 / a2 / a3 / a22  group convolution networks (GF, ET, RD) - inference only
# ----------------------------------------------------------------------------------------
def bn_eval(x, bn):
    """nn.BatchNorm2d in eval mode on [B,C,G]: (x-mean)/sqrt(var+eps)*w+b, eps=1e-5."""
    w, b, mean, var = bn
    s = (w / np.sqrt(var + F32(1e-5))).astype(F32)
    return x * s[None, :, None] + (b - mean * s).astype(F32)[None, :, None]


def group_conv(x, W, bias, nei, bn=None, relu=False):
    """data_process + (BN, ReLU,) Conv2d(Cin,Cout,(1,13)) - network/group_feat.py:20-24,
    network/ops.py:11-20:   out[b,o,g] = bias[o] + sum_{c,k} W[o,c,0,k] * act(x)[b,c,N[g,k]].
    x [B,Cin,60] float32, W [Cout,Cin,1,13]."""
    if bn is not None: x = bn_eval(x, bn)
    if relu: x = np.maximum(x, 0)
    D = x[:, :, nei]                                   # [B,C,60,13]
    B, C = x.shape[:2]
    A = D.transpose(0, 2, 1, 3).reshape(B * 60, C * 13)    # rows (b,g), cols (c,k)
    out = A @ W.reshape(W.shape[0], C * 13).T + bias[None]
    return out.reshape(B, 60, -1).transpose(0, 2, 1).astype(F32)


def conv1x1(x, W, bias):
    return (np.einsum("oc,bcg->bog", W.reshape(W.shape[0], -1), x) + bias[None, :, None]).astype(F32)


def _bn(sd, p):
    return tuple(sd[f"{p}.{k}"] for k in ("weight", "bias", "running_mean", "running_var"))


def residual_comb_conv(x, sd, p, nei):
    """Residual_Comb_Conv.forward network/ops.py:53-63."""
    h = group_conv(x, sd[f"{p}.comb_layer_in.2.weight"], sd[f"{p}.comb_layer_in.2.bias"], nei, _bn(sd, f"{p}.comb_layer_in.0"), True)
    h = group_conv(h, sd[f"{p}.comb_layer_out.2.weight"], sd[f"{p}.comb_layer_out.2.bias"], nei, _bn(sd, f"{p}.comb_layer_out.0"), True)
    if f"{p}.short_cut_layer.2.weight" in sd:
        sc = group_conv(x, sd[f"{p}.short_cut_layer.2.weight"], sd[f"{p}.short_cut_layer.2.bias"], nei, _bn(sd, f"{p}.short_cut_layer.0"), True)
    else:
        sc = x
    return h + sc


def gf_forward(x, sd, nei, prefix="PartI_net."):
    """Group_feat_network.forward network/group_feat.py:26-45.  Returns (eqv [B,32,60], inv [B,32])."""
    p = prefix
    h = group_conv(x, sd[p + "Conv_in.0.weight"], sd[p + "Conv_in.0.bias"], nei)
    h = residual_comb_conv(h, sd, p + "SO3_Conv_layers.0", nei)
    h = group_conv(h, sd[p + "Conv_out.comb_layer.2.weight"], sd[p + "Conv_out.comb_layer.2.bias"], nei, _bn(sd, p + "Conv_out.comb_layer.0"), True)
    eqv = h + x
    inv = np.mean(eqv, axis=-1)
    eqv = eqv / np.maximum(np.linalg.norm(eqv, axis=1, keepdims=True), F32(1e-4))
    inv = inv / np.maximum(np.linalg.norm(inv, axis=1, keepdims=True), F32(1e-4))
    return eqv.astype(F32), inv.astype(F32)


def et_forward(before0, before1, after0, after1, pre_idx, sd, nei, perm):
    """ET_test.forward network/eqv_trans.py:119-138.  Side-0 tensors are permuted by P[pre_idx]
    (:126-128); only group element 0 of the FC head is kept (:136).  Returns unit quaternions [B,4]."""
    pi = perm[pre_idx]                                          # [B,60]
    b0 = np.take_along_axis(before0, pi[:, None, :], 2); a0 = np.take_along_axis(after0, pi[:, None, :], 2)
    x = np.concatenate([b0, before1, a0, after1], 1).astype(F32)
    h = group_conv(x, sd["Conv_init.comb_layer.2.weight"], sd["Conv_init.comb_layer.2.bias"], nei, _bn(sd, "Conv_init.comb_layer.0"), True)
    h = residual_comb_conv(h, sd, "PartII_SO3_Conv_layers.0", nei)
    h = conv1x1(h, sd["PartII_To_R_FC.0.weight"], sd["PartII_To_R_FC.0.bias"])
    h = np.maximum(bn_eval(h, _bn(sd, "PartII_To_R_FC.1")), 0)
    h = conv1x1(h, sd["PartII_To_R_FC.3.weight"], sd["PartII_To_R_FC.3.bias"])
    h = np.maximum(bn_eval(h, _bn(sd, "PartII_To_R_FC.4")), 0)
    q = conv1x1(h, sd["PartII_To_R_FC.6.weight"], sd["PartII_To_R_FC.6.bias"])[:, :, 0]
    return (q / np.linalg.norm(q, axis=1)[:, None]).astype(F32)


def rd_forward(x, sd, nei, perm):
    """detector_eqv_test.forward network/rot_detect.py:43-55: Residual_Comb_Conv(32,64,16) ->
    L2 norm over channels -> autocorrelation V1 (X = Y) -> unbiased std over the 60 values."""
    h = residual_comb_conv(x, sd, "eqv_encoder.0", nei)
    h = h / np.linalg.norm(h, axis=1, keepdims=True)
    cor = group_corr_v1(h, h, perm)
    return np.std(cor, axis=1, ddof=1).astype(F32)


def rank_normalise(scores):
    """test/detector.py:44-46."""
    s = scores.copy(); a = np.argsort(s)
    s[a] = np.arange(s.shape[0]) / s.shape[0]
    return s


def random_state_dict(kind, seed):
    """Random-init weights with the parameter names/shapes of the shipped checkpoints
    (checkpoints/FCGF/{GF,ET,RD}/model_best.pth) - the GPU box has no checkpoints."""
    rng = np.random.default_rng(seed)
    sd = {}

    def conv(name, cout, cin, k):
        fan = cin * k
        sd[name + ".weight"] = (rng.standard_normal((cout, cin, 1, k)) / np.sqrt(fan)).astype(F32)
        sd[name + ".bias"] = (0.1 * rng.standard_normal(cout)).astype(F32)

    def bn(name, c):
        sd[name + ".weight"] = (1 + 0.1 * rng.standard_normal(c)).astype(F32)
        sd[name + ".bias"] = (0.1 * rng.standard_normal(c)).astype(F32)
        sd[name + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(F32)
        sd[name + ".running_var"] = (1 + 0.2 * rng.random(c)).astype(F32)

    def rcc(p, cin, mid, cout):
        bn(p + ".comb_layer_in.0", cin); conv(p + ".comb_layer_in.2", mid, cin, 13)
        bn(p + ".comb_layer_out.0", mid); conv(p + ".comb_layer_out.2", cout, mid, 13)
        if cin != cout:
            bn(p + ".short_cut_layer.0", cin); conv(p + ".short_cut_layer.2", cout, cin, 13)

    if kind == "GF":
        conv("PartI_net.Conv_in.0", 256, 32, 13)
        rcc("PartI_net.SO3_Conv_layers.0", 256, 512, 256)
        bn("PartI_net.Conv_out.comb_layer.0", 256); conv("PartI_net.Conv_out.comb_layer.2", 32, 256, 13)
    elif kind == "ET":
        bn("Conv_init.comb_layer.0", 128); conv("Conv_init.comb_layer.2", 256, 128, 13)
        rcc("PartII_SO3_Conv_layers.0", 256, 512, 256)
        conv("PartII_To_R_FC.0", 512, 256, 1); bn("PartII_To_R_FC.1", 512)
        conv("PartII_To_R_FC.3", 128, 512, 1); bn("PartII_To_R_FC.4", 128)
        conv("PartII_To_R_FC.6", 4, 128, 1)
        # bias the head towards the identity quaternion so that the hypotheses R_res @ Rgroup[a] are
        # near the planted pose (random heads give hypotheses with <= 3 inliers and the refiner's SVD
        # then sits on a rank-deficient matrix whose sign is LAPACK rounding noise)
        sd["PartII_To_R_FC.6.weight"] *= F32(0.05)
        sd["PartII_To_R_FC.6.bias"] = (np.array([3.0, 0, 0, 0]) + 0.02 * rng.standard_normal(4)).astype(F32)
    elif kind == "RD":
        rcc("eqv_encoder.0", 32, 64, 16)
    elif kind == "RM":
        def c1(name, cout, cin):
            sd[name + ".weight"] = (rng.standard_normal((cout, cin, 1, 1)) / np.sqrt(cin)).astype(F32)
            sd[name + ".bias"] = (0.1 * rng.standard_normal(cout)).astype(F32)

        def mlp(p, cin, mid, cout):
            c1(p + ".net.0", mid, cin); c1(p + ".net.3", cout, mid)
            if cin != cout: c1(p + ".res", cout, cin)

        def mhattn(p):
            c1(p + ".merge", 32, 32)
            for i in range(3): c1(p + f".proj.{i}", 32, 32)
        for b in range(2):
            for cg in ("cross_graph_s2t", "cross_graph_t2s"):
                p = f"Graph.merge_blocks.{b}.{cg}"
                mhattn(p + ".cross_attn"); mlp(p + ".merge", 96, 64, 32)
            for sg in ("self_graph_s", "self_graph_t"):
                p = f"Graph.merge_blocks.{b}.{sg}"
                mhattn(p + ".self_attn"); mlp(p + ".pos_en", 3, 64, 32); mlp(p + ".ambiguity", 120, 128, 32)
                mlp(p + ".val_en", 96, 64, 32); mlp(p + ".merge", 96, 64, 32)
        mlp("final_mlp", 64, 64, 32)
        sd["ot_layer.bin_score"] = np.array(1.0, F32)
    else:
        raise KeyError(kind)
    return sd


# ----------------------------------------------------------------------------------------
# a6-a11  rotation-coherence matcher Match_ot  (network/rot_coh_match.py), inference only
# ----------------------------------------------------------------------------------------
def _conv1x1(x, sd, p):
    """nn.Conv2d(cin, cout, 1) on [C, P] (P = flattened positions)."""
    W = sd[p + ".weight"].reshape(sd[p + ".weight"].shape[0], -1)
    return (W @ x + sd[p + ".bias"][:, None]).astype(F32)


def _instnorm(x):
    """nn.InstanceNorm2d(affine=False, eps=1e-5) on [C, P]: per-channel statistics over ALL positions (biased var)."""
    m = x.mean(axis=1, keepdims=True, dtype=np.float64); v = x.var(axis=1, keepdims=True, dtype=np.float64)
    return ((x - m) / np.sqrt(v + 1e-5)).astype(F32)


def mlp_2layer(x, sd, p):
    """mlp_2layer / Contextnorm forward (rot_coh_match.py:14-32, :63-81) on [Cin, P]."""
    h = _conv1x1(x, sd, p + ".net.0")
    h = np.maximum(_instnorm(h), 0)
    out = _conv1x1(h, sd, p + ".net.3")
    if p + ".res.weight" in sd:
        out = out + _conv1x1(x, sd, p + ".res")
    return out.astype(F32)


def knn_index_desc(score, k):
    """Knn_index_extract(axis=2) rot_coh_match.py:34-45: first k columns of a descending argsort per row."""
    return np.argsort(-score, axis=1, kind="stable")[:, :k]


def multi_head_attention(q, kf, vf, sd, p):
    """MultiHeadedAttention.forward + attention (rot_coh_match.py:84-119): q [32,m], kf / vf [32,m,k];
    4 heads, channel c = d*4 + h (view(b, dim, heads, -1))."""
    m, k = kf.shape[1], kf.shape[2]
    Q = _conv1x1(q, sd, p + ".proj.0").reshape(8, 4, m)
    K = _conv1x1(kf.reshape(32, m * k), sd, p + ".proj.1").reshape(8, 4, m, k)
    V = _conv1x1(vf.reshape(32, m * k), sd, p + ".proj.2").reshape(8, 4, m, k)
    sc = np.einsum("fhm,fhmk->hmk", Q, K) / np.float32(8 ** .5)
    sc = sc - sc.max(axis=-1, keepdims=True)
    pr = np.exp(sc); pr = pr / pr.sum(axis=-1, keepdims=True)
    x = np.einsum("hmk,dhmk->dhm", pr.astype(F32), V).reshape(32, m)
    return _conv1x1(x.astype(F32), sd, p + ".merge")


def _knn_hook(score, k, p, forced, trace):
    """Test hook (not a reference function): record the block's score matrix and own top-k in `trace`, and - for the
    teacher-forced parity test of the CUDA matcher - continue with the neighbour lists given in `forced`."""
    knn = knn_index_desc(score, k)
    if trace is not None:
        trace[p] = dict(score=score, knn=knn)
    if forced is not None and p in forced:
        knn = np.asarray(forced[p]).astype(np.int64)
    return knn


def cross_attention_block(src, tgt, src_eqv, tgt_eqv, featinv, k, s2t, sd, p, perm, forced=None, trace=None):
    """Cross_attention_block.forward rot_coh_match.py:132-165.  src/tgt/featinv [32,m]/[32,n]; *_eqv [m,32,60]."""
    score = (src.T @ tgt).astype(F32)                      # score_mat
    knn = _knn_hook(score, k, p, forced, trace); nn = knn[:, 0]          # :145 argsort(1) = first column of the stable argsort
    knn_fea = tgt[:, knn]                                  # [32,m,k]
    feat = multi_head_attention(src, knn_fea, knn_fea, sd, p + ".cross_attn")
    feat = mlp_2layer(np.concatenate([featinv, src, feat], 0), sd, p + ".merge")
    t_nn = tgt_eqv[nn]                                     # eqv descriptor of each point's 1-NN
    if s2t:
        rind = group_corr_v2(src_eqv, t_nn, perm)          # sum_{f,g} S[f,P[g,h]] T_nn[f,g]
    else:
        rind = group_corr_v2(t_nn, src_eqv, perm)          # sum_{f,g} T_nn[f,P[g,h]] S[f,g]
    return feat, rind.T.astype(F32)                        # [32,m], [60,m]


def self_attention_block(feat, coor, rind, featinv, k, sd, p, forced=None, trace=None):
    """Self_attention_block.forward rot_coh_match.py:187-210.  feat/featinv [32,m], coor [3,m], rind [60,m]."""
    m = feat.shape[1]
    score = (feat.T @ feat).astype(F32)
    knn = _knn_hook(score, k, p, forced, trace)
    knn_fea = feat[:, knn]                                 # [32,m,k]
    knn_coor = coor[:, knn] - coor[:, :, None]             # [3,m,k]
    pe = mlp_2layer(knn_coor.reshape(3, m * k).astype(F32), sd, p + ".pos_en").reshape(32, m, k)
    r2 = np.concatenate([rind, np.repeat(rind.max(axis=1, keepdims=True), m, 1)], 0)      # [120,m]
    conf = mlp_2layer(r2.astype(F32), sd, p + ".ambiguity")                                # [32,m]
    pe = pe / np.linalg.norm(pe, axis=0, keepdims=True)
    kf = knn_fea / np.linalg.norm(knn_fea, axis=0, keepdims=True)
    conf = conf / np.linalg.norm(conf, axis=0, keepdims=True)
    val_in = np.concatenate([pe, kf, np.repeat(conf[:, :, None], k, 2)], 0).reshape(96, m * k).astype(F32)
    value = mlp_2layer(val_in, sd, p + ".val_en").reshape(32, m, k)
    out = multi_head_attention(feat, kf.astype(F32), value, sd, p + ".self_attn")
    return mlp_2layer(np.concatenate([featinv, feat, out], 0), sd, p + ".merge")


def log_sinkhorn(scores, alpha, iters=100):
    """sinkhorn_ot.log_optimal_transport rot_coh_match.py:284-313 (float32, scipy-free logsumexp)."""
    m, n = scores.shape
    Z = np.full((m + 1, n + 1), alpha, F32); Z[:m, :n] = scores
    norm = F32(-np.log(F32(m + n)))
    log_mu = np.concatenate([np.full(m, norm, F32), [F32(np.log(F32(n))) + norm]]).astype(F32)
    log_nu = np.concatenate([np.full(n, norm, F32), [F32(np.log(F32(m))) + norm]]).astype(F32)

    def lse(a, axis):
        mx = a.max(axis=axis, keepdims=True)
        return (mx + np.log(np.exp(a - mx).sum(axis=axis, keepdims=True, dtype=F32))).squeeze(axis).astype(F32)
    u = np.zeros(m + 1, F32); v = np.zeros(n + 1, F32)
    for _ in range(iters):
        u = log_mu - lse(Z + v[None, :], 1)
        v = log_nu - lse(Z + u[:, None], 0)
    return (Z + u[:, None] + v[None, :] - norm).astype(F32)


def match_ot_forward(feats0, feats1, keys0, keys1, sd, perm, iters=100, forced=None, trace=None):
    """Match_ot.forward rot_coh_match.py:339-390 for one pair.  feats0/keys0 = the batch's 'source' side
    (test/matcher.py:192-197 feeds cloud id1 there).  Returns matches0 [m] (-1 = none), matching_scores0 [m],
    matches1 [n], matching_scores1 [n], and the OT matrix."""
    src_eqv = feats0.astype(F32); tgt_eqv = feats1.astype(F32)             # [m,32,60]
    sc = (keys0.astype(F32).T / F32(0.025)); tc = (keys1.astype(F32).T / F32(0.025))
    s_inv = src_eqv.mean(axis=-1).T.astype(F32); t_inv = tgt_eqv.mean(axis=-1).T.astype(F32)   # [32,m]
    src, tgt = s_inv, t_inv
    for li, k in enumerate((16, 8)):
        p = f"Graph.merge_blocks.{li}"
        s2t, r_s = cross_attention_block(src, tgt, src_eqv, tgt_eqv, s_inv, k, True, sd, p + ".cross_graph_s2t", perm, forced, trace)
        eh_s = self_attention_block(s2t, sc, r_s, s_inv, k, sd, p + ".self_graph_s", forced, trace)
        t2s, r_t = cross_attention_block(tgt, src, tgt_eqv, src_eqv, t_inv, k, False, sd, p + ".cross_graph_t2s", perm, forced, trace)
        eh_t = self_attention_block(t2s, tc, r_t, t_inv, k, sd, p + ".self_graph_t", forced, trace)
        src, tgt = eh_s, eh_t
    s_fin = mlp_2layer(np.concatenate([s_inv, src], 0), sd, "final_mlp")
    t_fin = mlp_2layer(np.concatenate([t_inv, tgt], 0), sd, "final_mlp")
    score = (s_fin.T @ t_fin).astype(F32)
    if trace is not None:
        trace["final_score"] = score
    Z = log_sinkhorn(score, F32(sd["ot_layer.bin_score"]), iters)
    inner = Z[:-1, :-1]
    i0 = inner.argmax(1); i1 = inner.argmax(0)
    m0 = np.arange(inner.shape[0]) == i1[i0]; m1 = np.arange(inner.shape[1]) == i0[i1]
    ms0 = np.where(m0, np.exp(inner.max(1)), 0).astype(F32)
    ms1 = np.where(m1, ms0[i1], 0).astype(F32)
    return np.where(m0, i0, -1), ms0, np.where(m1 & m0[i1], i1, -1), ms1, Z
