"""Mirror of test/matcher.py: NMS_sample (:11-42), mutual (:44-109), yoho_mat (:111-210)."""
import numpy as np
import torch
import tqdm
from ._common import context, make_non_exists_dir, feature_dataset_name, CloudCache


class NMS_sample():
    """test/matcher.py:11-42.  The 5-NN on xyz runs on the device (roreg_knn); the selection rule
    operates on the host exactly as the reference does (np.where / np.argsort on float64 scores)."""

    def __init__(self, num, k, cfg=None):
        self.k = k
        self.num = num
        self.ctx = context(cfg)

    def sample(self, keys, scores):
        if keys.shape[0] < self.num:
            return np.arange(keys.shape[0])
        kf = self.ctx.dev(keys.astype(np.float32))
        _, argmin = self.ctx.knn(kf, kf, self.k)
        argmin = argmin.cpu().numpy().astype(np.int64)            # [n,k]
        scores_nei = scores[argmin.reshape(-1)].reshape(-1, self.k)
        nei_max = np.max(scores_nei, axis=-1)
        sam_indexs = np.where(scores >= nei_max)[0]
        if sam_indexs.shape[0] > self.num:
            sam_scores = scores[sam_indexs]
            sam_scores = sam_scores / np.sum(sam_scores)
            resam_indexs = np.argsort(sam_scores)[-self.num:]
            sam_indexs = sam_indexs[resam_indexs]
        if sam_indexs.shape[0] < self.num:
            left = self.num - sam_indexs.shape[0]
            index_left = np.where(scores < nei_max)[0]
            scores_left = scores[index_left]
            left_index = np.argsort(scores_left)[-left:]
            left_index = index_left[left_index]
            sam_indexs = np.concatenate([sam_indexs, left_index], axis=0)
        return sam_indexs


def _sample_pair(cfg, dataset, datasetname, sampler, id0, id1, n0, n1, keynum):
    """Keypoint sampling shared by both matchers (test/matcher.py:76-88): NMS on detector scores with
    --RD, otherwise two shuffles of the GLOBAL NumPy RNG in the reference's order."""
    if cfg.RD:
        det_scores0 = np.load(f'{cfg.output_cache_fn}/{datasetname}/det_score/{id0}.npy')
        det_scores1 = np.load(f'{cfg.output_cache_fn}/{datasetname}/det_score/{id1}.npy')
        sample0 = sampler.sample(dataset.get_kps(id0), det_scores0)
        sample1 = sampler.sample(dataset.get_kps(id1), det_scores1)
    else:
        sample0 = np.arange(n0)
        sample1 = np.arange(n1)
        np.random.shuffle(sample0)
        np.random.shuffle(sample1)
        sample0 = sample0[0:keynum]
        sample1 = sample1[0:keynum]
    return sample0, sample1


class mutual():
    """test/matcher.py:44-109 - mutual nearest neighbours on the 32-d invariant descriptor."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.ctx = context(cfg)
        self.nn_mode = int(getattr(cfg, "nn_mode", 0))

    def run(self, dataset, keynum=5000):
        self.sampler = NMS_sample(keynum, 5, self.cfg)
        print(f'Matching the keypoints with mutual on {dataset.name}')
        Save_dir = f'{self.cfg.output_cache_fn}/{dataset.name}/match_{keynum}'
        make_non_exists_dir(Save_dir)
        Save_score_dir = f'{Save_dir}/scores'
        make_non_exists_dir(Save_score_dir)
        datasetname = feature_dataset_name(dataset)
        Feature_dir = f'{self.cfg.output_cache_fn}/{datasetname}/YOHO_Output_Group_feature'
        cache = CloudCache(self.ctx)
        for pair in tqdm.tqdm(dataset.pair_ids):
            id0, id1 = pair
            feats0 = cache.get(f'{Feature_dir}/{id0}.npy')      # [n,32,60] on the device
            feats1 = cache.get(f'{Feature_dir}/{id1}.npy')
            sample0, sample1 = _sample_pair(self.cfg, dataset, datasetname, self.sampler, id0, id1,
                                            feats0.shape[0], feats1.shape[0], keynum)
            s0 = self.ctx.dev(sample0.astype(np.int32)); s1 = self.ctx.dev(sample1.astype(np.int32))
            f0 = self.ctx.inv_pool(feats0, s0, normalise=True)
            f1 = self.ctx.inv_pool(feats1, s1, normalise=True)
            matches, cnt, _, _ = self.ctx.mutual_match(f0, f1, self.nn_mode)
            k = int(cnt.item())
            if k == 0:
                raise ValueError("need at least one array to concatenate")     # np.concatenate([]) in the reference (:106)
            m = matches[:k].cpu().numpy().astype(np.int64)
            match_pps = np.stack([sample0[m[:, 0]], sample1[m[:, 1]]], axis=1).astype(np.int64)
            np.save(f'{Save_dir}/{id0}-{id1}.npy', match_pps)
            np.save(f'{Save_score_dir}/{id0}-{id1}.npy', np.ones(match_pps.shape[0]))


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
        """test/matcher.py:131-150 on device tensors: pairs [K,2] = (index in 'source', its match), scores [K]."""
        matches0, scores = self.network.forward(feats_src, feats_tgt, keys_src, keys_tgt)
        matches0 = matches0.cpu().numpy(); scores = scores.cpu().numpy()
        sel = np.where(matches0 != -1)[0]
        if sel.shape[0] < 3:
            return None, scores[sel]
        return np.stack([sel, matches0[sel]], axis=1).astype(np.int64), scores[sel]

    def run(self, dataset, keynum=2500):
        self.sampler = NMS_sample(keynum, 5, self.cfg)
        Save_dir = f'{self.cfg.output_cache_fn}/{dataset.name}/match_{keynum}'
        make_non_exists_dir(Save_dir)
        Save_score_dir = f'{Save_dir}/scores'
        make_non_exists_dir(Save_score_dir)
        datasetname = feature_dataset_name(dataset)
        Feature_dir = f'{self.cfg.output_cache_fn}/{datasetname}/YOHO_Output_Group_feature'
        print(f'Matching the keypoints with rotation coherence matcher on {dataset.name}')
        cache = CloudCache(self.ctx)
        for pair in tqdm.tqdm(dataset.pair_ids):
            id0, id1 = pair
            feats0 = cache.get(f'{Feature_dir}/{id0}.npy')
            feats1 = cache.get(f'{Feature_dir}/{id1}.npy')
            sample0, sample1 = _sample_pair(self.cfg, dataset, datasetname, self.sampler, id0, id1,
                                            feats0.shape[0], feats1.shape[0], keynum)
            s0 = self.ctx.dev(sample0.astype(np.int64)); s1 = self.ctx.dev(sample1.astype(np.int64))
            f0 = feats0[s0].contiguous(); f1 = feats1[s1].contiguous()                # plumbing: row selection of the sampled keypoints
            keys0 = self.ctx.dev(dataset.get_kps(id0)[sample0].astype(np.float32))
            keys1 = self.ctx.dev(dataset.get_kps(id1)[sample1].astype(np.float32))
            # NOTE THE SWAP (test/matcher.py:192-197): the network's "source" (feats0/keys0) is cloud id1
            matches, scores = self.get_ot_match(f1, f0, keys1, keys0)
            if matches is None:
                raise TypeError("ones() takes 1 positional argument")                 # np.ones(1,2) in the reference (:201)
            matches_in_former = np.concatenate([sample0[matches[:, 1]][:, None], sample1[matches[:, 0]][:, None]], axis=1)
            np.save(f'{Save_dir}/{id0}-{id1}.npy', matches_in_former)
            np.save(f'{Save_score_dir}/{id0}-{id1}.npy', scores)
