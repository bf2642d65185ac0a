"""Host-side rules of the reference's plugins that stay on the host in this build (they consume the global NumPy RNG, decide
with float64 NumPy semantics, or write text) - restated in vectorised form.  Each function names the reference lines whose
observable behaviour (values, dtypes, RNG consumption order) it must reproduce; tests/test_plugins_host.py checks them
against files written by the unmodified reference."""
import numpy as np


def nms_select(scores, nn_idx, num):
    """Keypoint choice of NMS_sample.sample (test/matcher.py:23-40) given each point's k nearest neighbours nn_idx [n,k]
    (self included): local maxima of `scores` first - the `num` best of them when there are too many - topped up with the
    best non-maxima.  Ties and ordering follow np.argsort on the same arrays the reference sorts."""
    ridge = scores[nn_idx].max(axis=1)                       # neighbourhood maximum
    chosen = np.flatnonzero(scores >= ridge)
    if chosen.size > num:
        w = scores[chosen]
        w = w / np.sum(w)                                    # the reference sorts the NORMALISED scores (same float values -> same order)
        chosen = chosen[np.argsort(w)[-num:]]
    if chosen.size < num:
        rest = np.flatnonzero(scores < ridge)
        best_rest = rest[np.argsort(scores[rest])[chosen.size - num:]]
        chosen = np.concatenate([chosen, best_rest], axis=0)
    return chosen


def top_scored(scores, match_n):
    """--RM: indices of the matches the estimators keep (test/estimator.py:195-201, :415-420): the top `match_n` fraction by score
    (at least 10), or the top `match_n` matches when match_n >= 0.999; ascending-score order (tail of np.argsort)."""
    keep = max(scores.shape[0] * match_n, 10) if match_n < 0.999 else match_n
    return np.argsort(scores)[-int(keep):]


def rotation_buckets(dr_index):
    """yohoc_ransac.DR_statictic (test/estimator.py:119-137): members[r] = ascending match indices whose coarse rotation is r;
    prob[r] proportional to c (c - 0.01) (c - 0.02) with c = |members[r]| / 100 for buckets of >= 2 matches, else 0.
    Returns (None, zeros) when no bucket qualifies."""
    dr_index = np.asarray(dr_index)
    members = [np.flatnonzero(dr_index == r) for r in range(60)]
    count = np.array([m.size for m in members])
    c = count.astype(np.float64) / 100.0
    weight = np.where(count < 2, 0.0, c * (c - 0.01) * (c - 0.02))
    total = np.sum(weight)
    if total == 0:
        return None, np.zeros(60)
    return members, weight / total


def kabsch_3pt(p0, p1):
    """yohoc_ransac.Threepps2Tran (test/estimator.py:139-147): [R|t] with R p1 + t ~ p0 from the SVD of the centred
    cross-covariance, R = V U^T WITHOUT a determinant fix (rank-2 input, the sign is LAPACK's)."""
    mu0 = np.mean(p0, 0, keepdims=True)
    mu1 = np.mean(p1, 0, keepdims=True)
    U, _, Vt = np.linalg.svd((p1 - mu1).T @ (p0 - mu0))
    R = Vt.T @ U.T
    t = mu0 - (mu1 @ R.T)
    return np.hstack([R, t.T])


def kabsch_3pt_batch(p0, p1):
    """kabsch_3pt for a stack of triplets p0, p1 [H,3,3] -> [H,3,4].  NumPy's stacked mean / matmul / svd run the same
    per-matrix routines (LAPACK gesdd included) as the reference's per-hypothesis calls, so every entry is bit-identical to
    kabsch_3pt on that triplet - reflections included (tests/test_plugins_host.py checks the equality)."""
    mu0 = np.mean(p0, 1, keepdims=True)
    mu1 = np.mean(p1, 1, keepdims=True)
    U, _, Vt = np.linalg.svd(np.matmul((p1 - mu1).transpose(0, 2, 1), p0 - mu0))
    R = np.matmul(Vt.transpose(0, 2, 1), U.transpose(0, 2, 1))
    t = mu0 - np.matmul(mu1, R.transpose(0, 2, 1))
    return np.concatenate([R, t.transpose(0, 2, 1)], axis=2)


def draw_guided_triplets(members, prob, max_iter, max_draws=50000):
    """The RNG-consuming loop of yohoc_ransac.ransac_once (test/estimator.py:221-228) on the GLOBAL NumPy RNG: per iteration one
    categorical draw of a coarse rotation (rejected, without counting, when its bucket has < 2 matches) and one draw of three
    matches, with replacement, from that bucket.  At most max_draws + 1 rotation draws.  Returns int64 [iters,3]."""
    out = []
    draws = 0
    while len(out) < max_iter and draws <= max_draws:
        draws += 1
        r = np.random.choice(60, p=prob)
        if members[r].size < 2:
            continue
        out.append(np.random.choice(members[r], 3))
    return np.asarray(out, dtype=np.int64).reshape(-1, 3)


def trajectory_block(id0, id1, n_clouds, T):
    """The five pre.log lines of one pair (R_pre_log, test/estimator.py:14-26): 'id0<TAB>id1<TAB>n_clouds', the three pose rows in
    str() form, the constant last row."""
    rows = ['\t'.join(str(T[r][c]) for c in range(4)) + '\n' for r in range(3)]
    return f'{int(id0)}\t{int(id1)}\t{n_clouds}\n' + ''.join(rows) + '0.0\t0.0\t0.0\t1.0\n'


def write_trajectory(dataset, result_dir, poses=None, blocks=None):
    """R_pre_log (test/estimator.py:14-26): `pre.log` in the 3DMatch trajectory format read by utils/RR_cal.py:339 - per pair
    'id0<TAB>id1<TAB>n_clouds' then the 4 rows of the pose, last row constant; numbers in str() form.  `poses` (optional,
    [n_pairs,4,4] in pair order) = the matrices just stored in the .npz files, saving their re-read; `blocks` (optional, one
    trajectory_block string per pair, in pair order) = the text already formatted while the pairs were being registered."""
    n_clouds = len(dataset.pc_ids)
    with open(f'{result_dir}/pre.log', 'w') as f:
        if blocks is not None:
            f.write(''.join(blocks))
            return
        for i, (id0, id1) in enumerate(dataset.pair_ids):
            T = poses[i] if poses is not None else np.load(f'{result_dir}/{id0}-{id1}.npz', allow_pickle=True)['trans']
            f.write(trajectory_block(id0, id1, n_clouds, T))
