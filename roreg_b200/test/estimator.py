"""Mirror of test/estimator.py: R_pre_log (:14-26), refiner (:28-72), extractor_dr_index (:75-111),
yohoc_ransac (:113-264), yohoc (:266-272), extractor_localtrans (:275-367), yohoo_ransac (:369-443),
yohoo (:445-454).  Same class / method names and signatures, same files written, same consumption order of the global
NumPy RNG - with one caveat: the reference's yohoc forks multiprocessing.Pool(len(pair_ids)) (:258), so every
pair's draws start from a COPY of the parent's RNG state and the parent's state is not advanced; yohoc_ransac.ransac reproduces
exactly that (state restored before every pair and after the loop).  The arithmetic runs in libroreg_b200.so, the host-side
rules live in _hostlogic.py and the file layout in _common.CacheLayout."""
import numpy as np
import torch
from tqdm import tqdm
from ._common import context, make_non_exists_dir, CloudCache, CacheLayout
from . import _hostlogic as host


def R_pre_log(dataset, save_dir):
    """test/estimator.py:14-26 - 3DMatch trajectory text consumed by utils/RR_cal.py:339."""
    host.write_trajectory(dataset, save_dir)


def _scores_dev(ctx, scores):
    if scores.dtype == np.float64:
        return ctx.dev(scores, torch.float64)
    return ctx.dev(scores.astype(np.float32), torch.float32)


class _PairInputs:
    """What both estimators read for one pair (test/estimator.py:186-194, :404-411): keypoints of the matches, their scores,
    and - with --RM - the indices of the top-scored matches the hypotheses are restricted to."""

    def __init__(self, cfg, lay, dataset, id0, id1):
        self.scores = np.load(lay.scores(id0, id1))
        self.pps = np.load(lay.matches(id0, id1))
        self.k0 = dataset.get_kps(id0)[self.pps[:, 0]]
        self.k1 = dataset.get_kps(id1)[self.pps[:, 1]]
        self.kept = host.top_scored(self.scores, cfg.match_n) if cfg.RM else None

    def upload(self, ctx):
        return ctx.dev(self.k0, torch.float64), ctx.dev(self.k1, torch.float64), _scores_dev(ctx, self.scores)


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

    def Batch_Des2R_torch(self, des1_eqv, des2_eqv):
        """test/estimator.py:85-89 on device tensors [B,32,60] (des1 = before, des2 = after the rotation)."""
        _, am = self.ctx.group_corr(des1_eqv.contiguous(), des2_eqv.contiguous(), variant=1, want_cor=False)
        return am.to(torch.int64)

    def Des2R_torch(self, des1_eqv, des2_eqv):
        return self.Batch_Des2R_torch(des1_eqv[None], des2_eqv[None])[0]

    def Rindex(self, dataset, keynum):
        lay = CacheLayout(self.cfg, dataset, keynum)
        make_non_exists_dir(lay.dr_index_dir)
        print(f'extract the drindex of the matches on {dataset.name}')
        clouds = CloudCache(self.ctx)
        for id0, id1 in tqdm(dataset.pair_ids):
            pps = np.load(lay.matches(id0, id1))
            rows0 = self.ctx.dev(pps[:, 0].astype(np.int32)); rows1 = self.ctx.dev(pps[:, 1].astype(np.int32))
            # the reference calls Batch_Des2R_torch(feats1, feats0) (:110): X = cloud id1, Y = cloud id0
            _, am = self.ctx.group_corr(clouds.get(lay.yoho_desc(id1)), clouds.get(lay.yoho_desc(id0)), rows1, rows0,
                                        variant=1, want_cor=False)
            np.save(lay.dr_index(id0, id1), am.cpu().numpy().astype(np.int64))


class yohoc_ransac:
    """test/estimator.py:113-264 - coarse-rotation-guided RANSAC.

    cfg.yohoc_mode (extension, default 'parity'):
      'parity' - the triplet draws consume the global NumPy RNG exactly as :224-228 and the 3-point
                 Kabsch runs through ONE stacked np.linalg.svd on the host (the same LAPACK routine per
                 matrix as the reference's loop), because the rank-2 SVD's sign (rotation vs reflection)
                 is LAPACK rounding noise; scoring, selection and refinement run on the device.
      'device' - draws (counter-based RNG) and the proper-rotation Kabsch also run on the device."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.inliner_dist = cfg.ransac_ird
        self.ctx = context(cfg)
        self.mode = getattr(cfg, "yohoc_mode", "parity")

    def DR_statictic(self, DR_indexs):
        """:119-137 -> (matches per coarse rotation, sampling probability per coarse rotation)."""
        return host.rotation_buckets(DR_indexs)

    def Threepps2Tran(self, kps0_init, kps1_init):
        """:139-147 -> [3,4]."""
        return host.kabsch_3pt(kps0_init, kps1_init)

    def ransac_once(self, dataset, keynum, max_iter, pair):
        lay = CacheLayout(self.cfg, dataset, keynum)
        id0, id1 = pair
        out_file = lay.result('yohoc', max_iter, id0, id1)
        pin = _PairInputs(self.cfg, lay, dataset, id0, id1)
        kept = pin.kept if pin.kept is not None else np.arange(pin.pps.shape[0])
        k0_kept, k1_kept = pin.k0[kept], pin.k1[kept]            # triplets come from the kept matches, overlap is scored on all
        members, prob = self.DR_statictic(np.load(lay.dr_index(id0, id1))[kept])
        if np.sum(prob) < 1e-5:
            # no coarse rotation is shared by two matches: the reference stores a random pose and the sentinel 50000 (:214-218)
            np.savez(out_file, trans=np.random.rand(4, 4), center=np.ones([6, 3]), recalltime=50000)
            return 0
        ctx = self.ctx
        k0, k1, sc = pin.upload(ctx)
        if self.mode == 'device':
            seed = int(np.random.randint(0, 2 ** 31 - 1))
            trip = self._device_triplets(members, prob, max_iter, seed)
            hyps = ctx.kabsch3(ctx.dev(k0_kept, torch.float64), ctx.dev(k1_kept, torch.float64), ctx.dev(trip))
        else:
            trip = host.draw_guided_triplets(members, prob, max_iter)
            hyps = ctx.dev(host.kabsch_3pt_batch(k0_kept[trip], k1_kept[trip]), torch.float64)   # == Threepps2Tran per triplet, bit for bit
        best, _, _ = ctx.ransac_oneshot(k0, k1, sc, hyps, None, self.inliner_dist)
        b = int(best.item())
        if b < 0:
            raise ValueError("no 3-point hypothesis has a positive overlap (the reference fails in transform_points here)")
        T, _ = ctx.refine(k0, k1, sc, hyps, self.inliner_dist, order=None, T_index=best)
        np.savez(out_file, trans=T.cpu().numpy(), recalltime=b + 1)      # the reference counts iterations from 1 (:226,:236)

    def _device_triplets(self, members, prob, max_iter, seed):
        # host-side draw with a private Generator (the 'device' mode of the single-pair path keeps the
        # kernel inputs explicit; the batched engine draws inside coarse_hyp_kernel)
        rng = np.random.default_rng(seed)
        trip = np.empty((max_iter, 3), np.int32)
        for i, r in enumerate(rng.choice(60, size=max_iter, p=prob)):
            trip[i] = rng.choice(members[r], 3)
        return trip

    def ransac(self, dataset, keynum, max_iter=1000):
        lay = CacheLayout(self.cfg, dataset, keynum)
        make_non_exists_dir(lay.result_dir('yohoc', max_iter))
        print(f'Ransac with YOHO-C on {dataset.name}:')
        # the reference forks Pool(len(pair_ids)) (:258): one worker per pair, each starting from a copy of the parent's global
        # NumPy RNG state, the parent's own state untouched.  Pairs are independent, so the single device context processes them
        # in order - from the same starting state each, which is restored afterwards.
        state = np.random.get_state()
        for pair in tqdm(dataset.pair_ids):
            np.random.set_state(state)
            self.ransac_once(dataset, keynum, max_iter, pair)
        np.random.set_state(state)
        R_pre_log(dataset, lay.result_dir('yohoc', max_iter))
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
        self.net = nets.ETNet(self.ctx, load_state_dict(self.best_model_fn), npass=self.npass, chunk=max(16000, int(self.test_batch_size)))   # bs_ET is the reference's memory knob; the output does not depend on the chunking (eval-mode BN)

    def hypotheses(self, quat, pre_idx, Keys0_m, Keys1_m):
        ctx = self.ctx
        tr = ctx.hypotheses_from_quat(ctx.dev(quat, torch.float32), ctx.dev(np.asarray(pre_idx).astype(np.int32)),
                                      ctx.dev(Keys0_m, torch.float64), ctx.dev(Keys1_m, torch.float64))
        return tr.cpu().numpy()

    def Rt_pre(self, dataset, keynum):
        self._load_model()
        ctx = self.ctx
        lay = CacheLayout(self.cfg, dataset, keynum)
        make_non_exists_dir(lay.trans_pre_dir)
        print(f'Extracting the local transformation on each correspondence of {dataset.name}')
        clouds = CloudCache(ctx)
        for id0, id1 in tqdm(dataset.pair_ids):
            pps = np.load(lay.matches(id0, id1))
            rows0 = ctx.dev(pps[:, 0].astype(np.int32)); rows1 = ctx.dev(pps[:, 1].astype(np.int32))
            coarse = ctx.dev(np.load(lay.dr_index(id0, id1)).astype(np.int32))
            # batch_create (:293-306): side 0 of the network = cloud id1 ("exchanged")
            quat = self.net.forward(clouds.get(lay.fcgf_desc(id1)), rows1, clouds.get(lay.fcgf_desc(id0)), rows0,
                                    clouds.get(lay.yoho_desc(id1)), rows1, clouds.get(lay.yoho_desc(id0)), rows0, coarse)
            k0 = ctx.dev(dataset.get_kps(id0)[pps[:, 0], :], torch.float64)
            k1 = ctx.dev(dataset.get_kps(id1)[pps[:, 1], :], torch.float64)
            np.save(lay.trans_pre(id0, id1), ctx.hypotheses_from_quat(quat, coarse, k0, k1).cpu().numpy())


class yohoo_ransac:
    """test/estimator.py:369-443 - one-shot RANSAC over the per-match hypotheses + two refinements."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.inliner_dist = cfg.ransac_ird
        self.ctx = context(cfg)

    def ransac(self, dataset, keynum, max_iter=1000):
        lay = CacheLayout(self.cfg, dataset, keynum)
        out_dir = lay.result_dir('yohoo', max_iter)
        make_non_exists_dir(out_dir)
        ctx = self.ctx
        print(f'Ransac with YOHO-O on {dataset.name}:')
        for id0, id1 in tqdm(dataset.pair_ids):
            pin = _PairInputs(self.cfg, lay, dataset, id0, id1)
            hyps = np.load(lay.trans_pre(id0, id1))
            if pin.kept is not None:
                hyps = hyps[pin.kept]
            visit = np.arange(hyps.shape[0])
            np.random.shuffle(visit)                                     # the pair's only RNG use (:423-424)
            order = ctx.dev(visit[0:max_iter].astype(np.int32))
            k0, k1, sc = pin.upload(ctx)
            tr = ctx.dev(hyps, torch.float64)
            best, _, _ = ctx.ransac_oneshot(k0, k1, sc, tr, order, self.inliner_dist)
            b = int(best.item())
            if b < 0:
                raise ValueError("no hypothesis has a positive overlap (the reference fails in transform_points here)")
            T, _ = ctx.refine(k0, k1, sc, tr, self.inliner_dist, order=order, T_index=best)
            np.savez(lay.result('yohoo', max_iter, id0, id1), trans=T.cpu().numpy(), recalltime=b)
        R_pre_log(dataset, out_dir)


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
