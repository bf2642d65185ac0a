"""Whole-dataset driver on the batched engine - SURVEY.md section 8(f) rank 3 (stage I/O).

The reference registers a dataset as two plugin passes, `mutual.run` then `yohoc.run` (test/evaluator.py:23-48), and every pass
re-reads both clouds' 38 MB descriptor files for every PAIR (test/matcher.py:66-67, test/estimator.py:106-107).  `register_scene`
does the same work in one pass: every cloud is read and uploaded ONCE into a device arena, the pairs go through
`roreg_register_batch` B at a time (sharded over ranks when torch.distributed is initialised: pairs are independent, no
data-path collective), and the reference's on-disk contract - match / scores / DR_index files, `{id0}-{id1}.npz`, `pre.log`
 - is emitted by a background writer so that file I/O overlaps the next batch.

Same algorithm and files as `mutual` + `yohoc` of roreg_b200/test with `yohoc_mode='device'`: keypoint sampling consumes the
global NumPy RNG in the reference's order (one rank) or comes from the NMS sampler with --RD; the RANSAC triplets are drawn on
the device (counter-based RNG), so poses agree with the host-RNG plugins statistically, not bit for bit.  --RM (Match_ot
matcher) and the yohoo estimator are per-pair network paths and stay with their plugins.
"""
import concurrent.futures as cf
import numpy as np
import torch
import torch.distributed as dist
from .shard import shard_pairs
from .test._common import context, make_non_exists_dir, CacheLayout
from .test import _hostlogic as host


class AsyncWriter:
    """A small thread pool for the per-pair files; close() waits for every write and re-raises the first failure."""

    def __init__(self, threads=2):
        self.pool = cf.ThreadPoolExecutor(max_workers=max(1, threads))
        self.pending = []

    def submit(self, fn, *args):
        self.pending.append(self.pool.submit(fn, *args))

    def close(self):
        self.pool.shutdown(wait=True)
        for f in self.pending:
            f.result()
        self.pending = []


def load_scene(ctx, lay, dataset):
    """Every cloud of the dataset once: descriptors [C,n,32,60] float32 and keypoints [C,n,3] float64 on the device, plus the
    cloud id -> arena slot map.  Clouds are uploaded one by one (host peak = one cloud)."""
    ids = list(dataset.pc_ids)
    slot = {pc: i for i, pc in enumerate(ids)}
    desc = keys = None
    for i, pc in enumerate(ids):
        d = np.load(lay.yoho_desc(pc)); k = dataset.get_kps(pc)
        if desc is None:
            n = d.shape[0]
            desc = torch.empty((len(ids), n, 32, 60), dtype=torch.float32, device=ctx.device)
            keys = torch.empty((len(ids), n, 3), dtype=torch.float64, device=ctx.device)
        if d.shape != tuple(desc.shape[1:]) or k.shape != (desc.shape[1], 3):
            raise ValueError(f"cloud {pc}: {d.shape} descriptors / {k.shape} keypoints, the arena holds clouds of {desc.shape[1]} "
                             "keypoints (the reference's caches are 5000 per cloud)")
        desc[i].copy_(ctx.dev(d.astype(np.float32))); keys[i].copy_(ctx.dev(k, torch.float64))
    return desc, keys, slot


def draw_samples(cfg, lay, dataset, pairs, n, keynum):
    """[P,2,keynum] int32 keypoint samples in the reference's way (test/matcher.py:76-88): with --RD the NMS sampler on the
    detector scores (a function of the cloud alone: computed once per cloud), otherwise two shuffles of the GLOBAL NumPy RNG
    per pair, cloud id0 first, pairs in list order."""
    out = np.empty((len(pairs), 2, keynum), np.int32)
    if cfg.RD:
        from .test.matcher import NMS_sample
        sampler, per_cloud = NMS_sample(keynum, 5, cfg), {}
        for p, pair in enumerate(pairs):
            for side, pc in enumerate(pair):
                if pc not in per_cloud:
                    per_cloud[pc] = sampler.sample(dataset.get_kps(pc), np.load(lay.det_score(pc)))
                out[p, side] = per_cloud[pc]
        return out
    for p in range(len(pairs)):
        for side in range(2):
            perm = np.arange(n)
            np.random.shuffle(perm)
            out[p, side] = perm[0:keynum]
    return out


def _write_pair(lay, max_iter, id0, id1, matches, dr_index, pose, recall):
    np.save(lay.matches(id0, id1), matches)
    np.save(lay.scores(id0, id1), np.ones(matches.shape[0]))
    np.save(lay.dr_index(id0, id1), dr_index)
    np.savez(lay.result('yohoc', max_iter, id0, id1), trans=pose, recalltime=recall)


def register_scene(cfg, dataset, keynum=5000, max_iter=1000, batch_pairs=64, nn_mode=4, seed=0, writer_threads=2, ctx=None):
    """mutual.run + yohoc.run of the reference for a whole dataset on the batched engine.  Returns (on every rank) a dict with this
    rank's slice: pair indices `lo, hi`, `poses` [hi-lo,4,4] float64, `recall` [hi-lo], `n_matches` [hi-lo] (NumPy).  Files are
    written for the rank's own pairs; rank 0 writes pre.log after a barrier."""
    if getattr(cfg, "RM", False):
        raise NotImplementedError("--RM uses the Match_ot matcher: run the yoho_mat / yohoo plugins (roreg_b200.test)")
    ctx = ctx or context(cfg)
    lay = CacheLayout(cfg, dataset, keynum)
    out_dir = lay.result_dir('yohoc', max_iter)
    for d in (lay.match_dir, lay.scores_dir, lay.dr_index_dir, out_dir):
        make_non_exists_dir(d)
    dist_on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if dist_on else 0
    world = dist.get_world_size() if dist_on else 1
    pairs_all = list(dataset.pair_ids)
    desc, keys, slot = load_scene(ctx, lay, dataset)
    n = desc.shape[1]
    if keynum > n:
        raise ValueError(f"keynum {keynum} exceeds the {n} keypoints per cloud")
    samples_all = draw_samples(cfg, lay, dataset, pairs_all, n, keynum)        # all pairs on every rank: one RNG order for any world size
    lo, hi = shard_pairs(len(pairs_all), rank, world)
    poses = np.zeros((hi - lo, 4, 4)); recall = np.zeros(hi - lo, np.int64); counts = np.zeros(hi - lo, np.int64)
    writer = AsyncWriter(writer_threads)
    try:
        for s in range(lo, hi, batch_pairs):
            e = min(hi, s + batch_pairs)
            pc = ctx.dev(np.array([[slot[a], slot[b]] for a, b in pairs_all[s:e]], np.int32))
            o = ctx.register_batch(desc, keys, pc, keynum=keynum, sample=ctx.dev(samples_all[s:e]), nn_mode=nn_mode, estimator=0,
                                   max_iter=max_iter, ird=cfg.ransac_ird, seed=(int(seed) * 1000003 + s) & 0x7fffffffffffffff)
            h = {k: v.cpu().numpy() for k, v in o.items()}                        # one synchronisation per batch
            for j, (id0, id1) in enumerate(pairs_all[s:e]):
                k = int(h["n_matches"][j])
                if k == 0:
                    raise ValueError(f"pair {id0}-{id1}: need at least one array to concatenate")      # test/matcher.py:106
                m = h["matches"][j, :k].astype(np.int64); dr = h["dr_index"][j, :k].astype(np.int64)
                T = h["poses"][j].copy(); r = int(h["recall"][j])
                if r < 0:
                    if host.rotation_buckets(dr)[0] is None:                   # no coarse rotation shared by two matches (test/estimator.py:214-218)
                        T, r = np.random.rand(4, 4), 50000
                    else:
                        raise ValueError(f"pair {id0}-{id1}: no 3-point hypothesis has a positive overlap")
                else:
                    r += 1                                                      # engine: 0-based id of the winner; reference: 1-based iteration (test/estimator.py:226,236)
                poses[s - lo + j] = T; recall[s - lo + j] = r; counts[s - lo + j] = k
                writer.submit(_write_pair, lay, max_iter, id0, id1, m, dr, T, r)
    finally:
        writer.close()
    if dist_on:
        dist.barrier()
    if rank == 0:
        host.write_trajectory(dataset, out_dir)
    return dict(lo=lo, hi=hi, poses=poses, recall=recall, n_matches=counts)
