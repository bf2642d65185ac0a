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
    """test/matcher.py:111-210 - rotation-coherence matcher (Match_ot).  SURVEY.md section 8(f) rank 2:
    the graph blocks + Sinkhorn kernels are not in this build; constructing the plugin fails loudly
    rather than running the PyTorch network (no fallback paths in the product)."""

    def __init__(self, cfg):
        raise NotImplementedError("yoho_mat (--RM, Match_ot) kernels are not part of this build; "
                                  "use the mutual matcher ('matmul')")

    def run(self, dataset, keynum=2500):
        raise NotImplementedError
