"""Mirror of test/estimator.py: R_pre_log (:14-26), refiner (:28-72), extractor_dr_index (:75-111),
yohoc_ransac (:113-264), yohoc (:266-272), extractor_localtrans (:275-367), yohoo_ransac (:369-443),
yohoo (:445-454).  Same signatures, same files written, same consumption order of the global NumPy
RNG; the arithmetic runs in libroreg_b200.so."""
import os
import numpy as np
import torch
from tqdm import tqdm
from ._common import context, make_non_exists_dir, feature_dataset_name, CloudCache


def R_pre_log(dataset, save_dir):
    """test/estimator.py:14-26 - 3DMatch trajectory text consumed by utils/RR_cal.py:339."""
    writer = open(f'{save_dir}/pre.log', 'w')
    pair_num = int(len(dataset.pc_ids))
    for pair in dataset.pair_ids:
        pc0, pc1 = pair
        ransac_result = np.load(f'{save_dir}/{pc0}-{pc1}.npz', allow_pickle=True)
        transform_pr = ransac_result['trans']
        writer.write(f'{int(pc0)}\t{int(pc1)}\t{pair_num}\n')
        writer.write(f'{transform_pr[0][0]}\t{transform_pr[0][1]}\t{transform_pr[0][2]}\t{transform_pr[0][3]}\n')
        writer.write(f'{transform_pr[1][0]}\t{transform_pr[1][1]}\t{transform_pr[1][2]}\t{transform_pr[1][3]}\n')
        writer.write(f'{transform_pr[2][0]}\t{transform_pr[2][1]}\t{transform_pr[2][2]}\t{transform_pr[2][3]}\n')
        writer.write(f'{0.0}\t{0.0}\t{0.0}\t{1.0}\n')
    writer.close()


def _scores_dev(ctx, scores):
    if scores.dtype == np.float64:
        return ctx.dev(scores, torch.float64)
    return ctx.dev(scores.astype(np.float32), torch.float32)


class refiner:
    """test/estimator.py:28-72.  Refine_trans = one weighted-Kabsch round on the inliers of T."""

    def __init__(self, cfg=None):
        self.ctx = context(cfg)

    def Refine_trans(self, key_m0, key_m1, T, scores, inlinerdist=None):
        ctx = self.ctx
        k0 = ctx.dev(key_m0, torch.float64); k1 = ctx.dev(key_m1, torch.float64)
        Tn, _ = ctx.refine_once(k0, k1, _scores_dev(ctx, scores), ctx.dev(np.asarray(T)[:3], torch.float64), inlinerdist)
        return Tn.cpu().numpy()

    def refine_twice(self, key_m0, key_m1, T, scores, ird):
        ctx = self.ctx
        k0 = ctx.dev(key_m0, torch.float64); k1 = ctx.dev(key_m1, torch.float64)
        Tn, _ = ctx.refine(k0, k1, _scores_dev(ctx, scores), ctx.dev(np.asarray(T)[:3], torch.float64), ird)
        return Tn.cpu().numpy()


# yohoc
class extractor_dr_index:
    """test/estimator.py:75-111 - coarse rotation index of every match (Des2R)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.ctx = context(cfg)

    def Batch_Des2R_torch(self, des1_eqv, des2_eqv):  # beforerot afterrot
        """test/estimator.py:85-89 on device tensors [B,32,60]."""
        _, am = self.ctx.group_corr(des1_eqv.contiguous(), des2_eqv.contiguous(), variant=1, want_cor=False)
        return am.to(torch.int64)

    def Des2R_torch(self, des1_eqv, des2_eqv):
        return self.Batch_Des2R_torch(des1_eqv[None], des2_eqv[None])[0]

    def Rindex(self, dataset, keynum):
        match_dir = f'{self.cfg.output_cache_fn}/{dataset.name}/match_{keynum}'
        Save_dir = f'{match_dir}/DR_index'
        make_non_exists_dir(Save_dir)
        datasetname = feature_dataset_name(dataset)
        Feature_dir = f'{self.cfg.output_cache_fn}/{datasetname}/YOHO_Output_Group_feature'
        print(f'extract the drindex of the matches on {dataset.name}')
        cache = CloudCache(self.ctx)
        for pair in tqdm(dataset.pair_ids):
            id0, id1 = pair
            match_pps = np.load(f'{match_dir}/{id0}-{id1}.npy')
            feats0 = cache.get(f'{Feature_dir}/{id0}.npy')
            feats1 = cache.get(f'{Feature_dir}/{id1}.npy')
            i0 = self.ctx.dev(match_pps[:, 0].astype(np.int32)); i1 = self.ctx.dev(match_pps[:, 1].astype(np.int32))
            # Batch_Des2R_torch(feats1, feats0): X = cloud id1, Y = cloud id0   (:110)
            _, am = self.ctx.group_corr(feats1, feats0, i1, i0, variant=1, want_cor=False)
            np.save(f'{Save_dir}/{id0}-{id1}.npy', am.cpu().numpy().astype(np.int64))


class yohoc_ransac:
    """test/estimator.py:113-264 - coarse-rotation-guided RANSAC.

    cfg.yohoc_mode (extension, default 'parity'):
      'parity' - the triplet draws consume the global NumPy RNG exactly as :224-228 and the 3-point
                 Kabsch runs through np.linalg.svd on the host, because the rank-2 SVD's sign (rotation
                 vs reflection) is LAPACK rounding noise; scoring, selection and refinement
                 run on the device.
      'device' - draws (counter-based RNG) and the proper-rotation Kabsch also run on the device."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.inliner_dist = cfg.ransac_ird
        self.ctx = context(cfg)
        self.mode = getattr(cfg, "yohoc_mode", "parity")

    def DR_statictic(self, DR_indexs):
        R_index_pre_statistic = {}
        for i in range(60):
            R_index_pre_statistic[i] = []
        for t in range(DR_indexs.shape[0]):
            R_index_pre_statistic[DR_indexs[t]].append(t)
        R_index_pre_probability = []
        for i in range(60):
            if len(R_index_pre_statistic[i]) < 2:
                R_index_pre_probability.append(0)
            else:
                num = float(len(R_index_pre_statistic[i])) / 100.0
                R_index_pre_probability.append(num * (num - 0.01) * (num - 0.02))
        R_index_pre_probability = np.array(R_index_pre_probability)
        if np.sum(R_index_pre_probability) == 0:
            return None, np.zeros(60)
        R_index_pre_probability = R_index_pre_probability / np.sum(R_index_pre_probability)
        return R_index_pre_statistic, R_index_pre_probability

    def Threepps2Tran(self, kps0_init, kps1_init):
        center0 = np.mean(kps0_init, 0, keepdims=True)
        center1 = np.mean(kps1_init, 0, keepdims=True)
        m = (kps1_init - center1).T @ (kps0_init - center0)
        U, S, VT = np.linalg.svd(m)
        rotation = VT.T @ U.T
        offset = center0 - (center1 @ rotation.T)
        return np.concatenate([rotation, offset.T], 1)

    def ransac_once(self, dataset, keynum, max_iter, pair):
        match_dir = f'{self.cfg.output_cache_fn}/{dataset.name}/match_{keynum}'
        Index_dir = f'{match_dir}/DR_index'
        Save_dir = f'{match_dir}/yohoc/{max_iter}iters'
        id0, id1 = pair
        Keys0 = dataset.get_kps(id0)
        Keys1 = dataset.get_kps(id1)
        scores = np.load(f'{match_dir}/scores/{id0}-{id1}.npy')
        pps = np.load(f'{match_dir}/{id0}-{id1}.npy')
        Keys_m0_init = Keys0[pps[:, 0]]
        Keys_m1_init = Keys1[pps[:, 1]]
        sample_index = np.arange(pps.shape[0])
        if self.cfg.RM:
            if self.cfg.match_n < 0.999:
                num = max(scores.shape[0] * self.cfg.match_n, 10)
            else:
                num = self.cfg.match_n
            sample_index = np.argsort(scores)[-int(num):]
        Keys_m0 = Keys_m0_init[sample_index]
        Keys_m1 = Keys_m1_init[sample_index]
        Index = np.load(f'{Index_dir}/{id0}-{id1}.npy')[sample_index]
        R_index_pre_statistic, R_index_pre_probability = self.DR_statictic(Index)
        best_3p_in_0 = np.ones([3, 3]); best_3p_in_1 = np.ones([3, 3])
        if np.sum(R_index_pre_probability) < 1e-5:
            np.savez(f'{Save_dir}/{id0}-{id1}.npz', trans=np.random.rand(4, 4),
                     center=np.concatenate([best_3p_in_0, best_3p_in_1], axis=0), recalltime=50000)
            return 0
        ctx = self.ctx
        k0 = ctx.dev(Keys_m0_init, torch.float64); k1 = ctx.dev(Keys_m1_init, torch.float64)
        sc = _scores_dev(ctx, scores)
        if self.mode == 'device':
            seed = int(np.random.randint(0, 2 ** 31 - 1))
            trip = self._device_triplets(Index, max_iter, seed)
            hyps = ctx.kabsch3(ctx.dev(Keys_m0, torch.float64), ctx.dev(Keys_m1, torch.float64), ctx.dev(trip))
        else:
            iter_ransac, exec_time, max_time = 0, 0, 50000
            hyps = []
            while iter_ransac < max_iter:
                if exec_time > max_time: break
                exec_time += 1
                R_index = np.random.choice(range(60), p=R_index_pre_probability)
                if (len(R_index_pre_statistic[R_index]) < 2):
                    continue
                iter_ransac += 1
                idxs_init = np.random.choice(np.array(R_index_pre_statistic[R_index]), 3)
                hyps.append(self.Threepps2Tran(Keys_m0[idxs_init], Keys_m1[idxs_init]))
            hyps = ctx.dev(np.stack(hyps, 0), torch.float64)
        best, bov, _ = ctx.ransac_oneshot(k0, k1, sc, hyps, None, self.inliner_dist)
        b = int(best.item())
        if b < 0:
            raise ValueError("no 3-point hypothesis has a positive overlap (the reference fails in transform_points here)")
        T, _ = ctx.refine(k0, k1, sc, hyps, self.inliner_dist, order=None, T_index=best)
        np.savez(f'{Save_dir}/{id0}-{id1}.npz', trans=T.cpu().numpy(), recalltime=b + 1)

    def _device_triplets(self, Index, max_iter, seed):
        # host-side draw with a private Generator (the 'device' mode of the single-pair path keeps the
        # kernel inputs explicit; the batched engine draws inside coarse_hyp_kernel)
        rng = np.random.default_rng(seed)
        stat, prob = self.DR_statictic(Index)
        rs = rng.choice(60, size=max_iter, p=prob)
        trip = np.empty((max_iter, 3), np.int32)
        for i, r in enumerate(rs):
            trip[i] = rng.choice(np.array(stat[r]), 3)
        return trip

    def ransac(self, dataset, keynum, max_iter=1000):
        match_dir = f'{self.cfg.output_cache_fn}/{dataset.name}/match_{keynum}'
        Save_dir = f'{match_dir}/yohoc/{max_iter}iters'
        make_non_exists_dir(Save_dir)
        print(f'Ransac with YOHO-C on {dataset.name}:')
        # the reference forks Pool(len(pair_ids)) (:258); pairs are independent, so the single device
        # context processes them in order instead
        for pair in tqdm(dataset.pair_ids):
            self.ransac_once(dataset, keynum, max_iter, pair)
        R_pre_log(dataset, Save_dir)
        print('Done')


class yohoc:
    def __init__(self, cfg):
        self.rind_extractor = extractor_dr_index(cfg)
        self.ransacer = yohoc_ransac(cfg)

    def run(self, dataset, keynum, max_iter):
        self.rind_extractor.Rindex(dataset, keynum)
        self.ransacer.ransac(dataset, keynum, max_iter)


# yohoo
class extractor_localtrans():
    """test/estimator.py:275-367: ET network -> residual quaternion per match -> R = quat2mat(q) @ Rgroup[pre_idx],
    t = key0 - key1 @ R.T -> Trans_pre/{id0}-{id1}.npy float64 [K,3,4].  Network and pose arithmetic on the device."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.ctx = context(cfg)
        self.best_model_fn = f'{self.cfg.model_fn}/ET/model_best.pth'
        self.test_batch_size = self.cfg.bs_ET
        self.npass = int(getattr(cfg, "net_passes", 3))
        self.net = None

    def _load_model(self):
        from .extractor import load_state_dict
        from .. import nets
        # strict=False in the reference (:289): the checkpoint also carries an unused PartI_net.* copy
        self.net = nets.ETNet(self.ctx, load_state_dict(self.best_model_fn), npass=self.npass, chunk=int(self.test_batch_size))

    def hypotheses(self, quat, pre_idx, Keys0_m, Keys1_m):
        ctx = self.ctx
        tr = ctx.hypotheses_from_quat(ctx.dev(quat, torch.float32), ctx.dev(np.asarray(pre_idx).astype(np.int32)),
                                      ctx.dev(Keys0_m, torch.float64), ctx.dev(Keys1_m, torch.float64))
        return tr.cpu().numpy()

    def Rt_pre(self, dataset, keynum):
        self._load_model()
        ctx = self.ctx
        match_dir = f'{self.cfg.output_cache_fn}/{dataset.name}/match_{keynum}'
        DRindex_dir = f'{match_dir}/DR_index'
        Save_dir = f'{match_dir}/Trans_pre'
        make_non_exists_dir(Save_dir)
        datasetname = feature_dataset_name(dataset)
        FCGF_dir = f'{self.cfg.output_cache_fn}/{datasetname}/{self.cfg.backbone}_Input_Group_feature'
        YOMO_dir = f'{self.cfg.output_cache_fn}/{datasetname}/YOHO_Output_Group_feature'
        print(f'Extracting the local transformation on each correspondence of {dataset.name}')
        cache = CloudCache(ctx)
        for pair in tqdm(dataset.pair_ids):
            id0, id1 = pair
            pps = np.load(f'{match_dir}/{id0}-{id1}.npy')
            Index_pre = np.load(f'{DRindex_dir}/{id0}-{id1}.npy')
            i0 = ctx.dev(pps[:, 0].astype(np.int32)); i1 = ctx.dev(pps[:, 1].astype(np.int32))
            pre = ctx.dev(Index_pre.astype(np.int32))
            # batch_create (:293-306): side 0 of the network = cloud id1 ("exchanged")
            quat = self.net.forward(cache.get(f'{FCGF_dir}/{id1}.npy'), i1, cache.get(f'{FCGF_dir}/{id0}.npy'), i0,
                                    cache.get(f'{YOMO_dir}/{id1}.npy'), i1, cache.get(f'{YOMO_dir}/{id0}.npy'), i0, pre)
            Keys0 = dataset.get_kps(id0)[pps[:, 0], :]
            Keys1 = dataset.get_kps(id1)[pps[:, 1], :]
            Trans = ctx.hypotheses_from_quat(quat, pre, ctx.dev(Keys0, torch.float64), ctx.dev(Keys1, torch.float64))
            np.save(f'{Save_dir}/{id0}-{id1}.npy', Trans.cpu().numpy())


class yohoo_ransac:
    """test/estimator.py:369-443 - one-shot RANSAC over the per-match hypotheses + two refinements."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.inliner_dist = cfg.ransac_ird
        self.ctx = context(cfg)

    def ransac(self, dataset, keynum, max_iter=1000):
        match_dir = f'{self.cfg.output_cache_fn}/{dataset.name}/match_{keynum}'
        Trans_dir = f'{match_dir}/Trans_pre'
        Save_dir = f'{match_dir}/yohoo/{max_iter}iters'
        make_non_exists_dir(Save_dir)
        ctx = self.ctx
        print(f'Ransac with YOHO-O on {dataset.name}:')
        for pair in tqdm(dataset.pair_ids):
            id0, id1 = pair
            Keys0 = dataset.get_kps(id0)
            Keys1 = dataset.get_kps(id1)
            scores = np.load(f'{match_dir}/scores/{id0}-{id1}.npy')
            pps = np.load(f'{match_dir}/{id0}-{id1}.npy')
            Keys_m0 = Keys0[pps[:, 0]]
            Keys_m1 = Keys1[pps[:, 1]]
            Trans = np.load(f'{Trans_dir}/{id0}-{id1}.npy')
            if self.cfg.RM:
                if self.cfg.match_n < 0.999:
                    num = max(scores.shape[0] * self.cfg.match_n, 10)
                else:
                    num = self.cfg.match_n
                sample_index = np.argsort(scores)[-int(num):]
                Trans = Trans[sample_index]
            index = np.arange(Trans.shape[0])
            np.random.shuffle(index)
            order = ctx.dev(index[0:max_iter].astype(np.int32))
            k0 = ctx.dev(Keys_m0, torch.float64); k1 = ctx.dev(Keys_m1, torch.float64)
            sc = _scores_dev(ctx, scores)
            tr = ctx.dev(Trans, torch.float64)
            best, _, _ = ctx.ransac_oneshot(k0, k1, sc, tr, order, self.inliner_dist)
            b = int(best.item())
            if b < 0:
                raise ValueError("no hypothesis has a positive overlap (the reference fails in transform_points here)")
            T, _ = ctx.refine(k0, k1, sc, tr, self.inliner_dist, order=order, T_index=best)
            np.savez(f'{Save_dir}/{id0}-{id1}.npz', trans=T.cpu().numpy(), recalltime=b)
        R_pre_log(dataset, Save_dir)


class yohoo:
    def __init__(self, cfg):
        self.cfg = cfg
        self.rind_extractor = extractor_dr_index(cfg)
        self.localT_extractor = extractor_localtrans(cfg)
        self.ransacer = yohoo_ransac(cfg)

    def run(self, dataset, keynum, max_iter):
        self.rind_extractor.Rindex(dataset, keynum)
        self.localT_extractor.Rt_pre(dataset, keynum)
        self.ransacer.ransac(dataset, keynum, max_iter)
