"""Mirror of test/matcher.py: NMS_sample (:11-42), mutual (:44-109), yoho_mat (:111-210).  Same class names, constructor /
run() signatures and files; sampling rules and the global-RNG order are the reference's (checked against files it wrote,
tests/test_plugins_host.py), the arithmetic runs in libroreg_b200.so."""
import numpy as np
import torch
import tqdm
from ._common import context, make_non_exists_dir, CloudCache, CacheLayout
from . import _hostlogic as host


class NMS_sample():
    """test/matcher.py:11-42.  The 5-NN on xyz runs on the device (roreg_knn); the selection rule (_hostlogic.nms_select)
    works on the host with the reference's NumPy semantics (float64 scores, np.argsort order)."""

    def __init__(self, num, k, cfg=None):
        self.k = k
        self.num = num
        self.ctx = context(cfg)

    def sample(self, keys, scores):
        n = keys.shape[0]
        if n < self.num:                                   # fewer keypoints than requested: all of them (:19-20)
            return np.arange(n)
        xyz = self.ctx.dev(keys.astype(np.float32))
        _, nn_idx = self.ctx.knn(xyz, xyz, self.k)         # [n,k], the point itself included
        return host.nms_select(scores, nn_idx.cpu().numpy().astype(np.int64), self.num)


def _sample_pair(cfg, lay, dataset, sampler, id0, id1, n0, n1, keynum):
    """Keypoint sampling shared by both matchers (test/matcher.py:76-88): NMS on detector scores with
    --RD, otherwise two shuffles of the GLOBAL NumPy RNG in the reference's order (cloud id0 first)."""
    if cfg.RD:
        return (sampler.sample(dataset.get_kps(id0), np.load(lay.det_score(id0))),
                sampler.sample(dataset.get_kps(id1), np.load(lay.det_score(id1))))
    picks = []
    for n in (n0, n1):
        perm = np.arange(n)
        np.random.shuffle(perm)
        picks.append(perm)
    return picks[0][0:keynum], picks[1][0:keynum]


class mutual():
    """test/matcher.py:44-109 - mutual nearest neighbours on the 32-d invariant descriptor."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.ctx = context(cfg)
        self.nn_mode = int(getattr(cfg, "nn_mode", 0))

    def run(self, dataset, keynum=5000):
        self.sampler = NMS_sample(keynum, 5, self.cfg)
        print(f'Matching the keypoints with mutual on {dataset.name}')
        lay = CacheLayout(self.cfg, dataset, keynum)
        make_non_exists_dir(lay.match_dir)
        make_non_exists_dir(lay.scores_dir)
        clouds = CloudCache(self.ctx)
        for id0, id1 in tqdm.tqdm(dataset.pair_ids):
            eqv0 = clouds.get(lay.yoho_desc(id0))              # [n,32,60] on the device
            eqv1 = clouds.get(lay.yoho_desc(id1))
            sample0, sample1 = _sample_pair(self.cfg, lay, dataset, self.sampler, id0, id1, eqv0.shape[0], eqv1.shape[0], keynum)
            inv0 = self.ctx.inv_pool(eqv0, self.ctx.dev(sample0.astype(np.int32)), normalise=True)
            inv1 = self.ctx.inv_pool(eqv1, self.ctx.dev(sample1.astype(np.int32)), normalise=True)
            pairs, count, _, _ = self.ctx.mutual_match(inv0, inv1, self.nn_mode)
            k = int(count.item())
            if k == 0:
                raise ValueError("need at least one array to concatenate")     # np.concatenate([]) in the reference (:106)
            pairs = pairs[:k].cpu().numpy().astype(np.int64)
            # back to indices into the full clouds: column 0 -> cloud id0, column 1 -> cloud id1
            np.save(lay.matches(id0, id1), np.stack([sample0[pairs[:, 0]], sample1[pairs[:, 1]]], axis=1).astype(np.int64))
            np.save(lay.scores(id0, id1), np.ones(k))


class yoho_mat():
    """test/matcher.py:111-210 - rotation-coherence matcher (Match_ot, --RM)."""

    def __init__(self, cfg):
        from .extractor import load_state_dict
        from .. import matchot
        self.cfg = cfg
        self.ctx = context(cfg)
        self.best_model_fn = f'{self.cfg.model_fn}/RM/model_best.pth'
        self.npass = int(getattr(cfg, "net_passes", 3))
        self.network = matchot.MatchOT(self.ctx, load_state_dict(self.best_model_fn), npass=self.npass)

    def get_ot_match(self, feats_src, feats_tgt, keys_src, keys_tgt):
        """test/matcher.py:131-150 on device tensors: pairs [K,2] = (index in 'source', its match), scores [K];
        pairs is None below three matches (the reference then fails on np.ones(1,2), see run)."""
        assigned, confidence = self.network.forward(feats_src, feats_tgt, keys_src, keys_tgt)
        assigned = assigned.cpu().numpy(); confidence = confidence.cpu().numpy()
        src = np.flatnonzero(assigned != -1)
        if src.size < 3:
            return None, confidence[src]
        return np.stack([src, assigned[src]], axis=1).astype(np.int64), confidence[src]

    def run(self, dataset, keynum=2500):
        self.sampler = NMS_sample(keynum, 5, self.cfg)
        lay = CacheLayout(self.cfg, dataset, keynum)
        make_non_exists_dir(lay.match_dir)
        make_non_exists_dir(lay.scores_dir)
        print(f'Matching the keypoints with rotation coherence matcher on {dataset.name}')
        clouds = CloudCache(self.ctx)
        for id0, id1 in tqdm.tqdm(dataset.pair_ids):
            eqv0 = clouds.get(lay.yoho_desc(id0))
            eqv1 = clouds.get(lay.yoho_desc(id1))
            sample0, sample1 = _sample_pair(self.cfg, lay, dataset, self.sampler, id0, id1, eqv0.shape[0], eqv1.shape[0], keynum)
            rows0 = self.ctx.dev(sample0.astype(np.int64)); rows1 = self.ctx.dev(sample1.astype(np.int64))
            f0 = eqv0[rows0].contiguous(); f1 = eqv1[rows1].contiguous()            # plumbing: rows of the sampled keypoints
            xyz0 = self.ctx.dev(dataset.get_kps(id0)[sample0].astype(np.float32))
            xyz1 = self.ctx.dev(dataset.get_kps(id1)[sample1].astype(np.float32))
            # NOTE THE SWAP (test/matcher.py:192-197): the network's "source" side is cloud id1
            pairs, scores = self.get_ot_match(f1, f0, xyz1, xyz0)
            if pairs is None:
                raise TypeError("ones() takes 1 positional argument")                 # np.ones(1,2) in the reference (:201)
            # pairs[:,0] indexes cloud id1's sample, pairs[:,1] cloud id0's: store (id0 index, id1 index)
            np.save(lay.matches(id0, id1), np.stack([sample0[pairs[:, 1]], sample1[pairs[:, 0]]], axis=1))
            np.save(lay.scores(id0, id1), scores)
