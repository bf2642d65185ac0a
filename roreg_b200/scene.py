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


def _read_npy_into(path, dst):
    """Read a C-ordered .npy file straight into `dst` (a NumPy view of pinned host memory): header parsed with numpy.lib.format,
    payload read with readinto - one copy (page cache -> pinned buffer), the GIL released while it runs."""
    with open(path, "rb") as f:
        major, _ = np.lib.format.read_magic(f)
        shape, fortran, dtype = (np.lib.format.read_array_header_1_0 if major == 1 else np.lib.format.read_array_header_2_0)(f)
        if fortran or dtype != dst.dtype or tuple(shape) != tuple(dst.shape):
            raise ValueError(f"{path}: {dtype} {shape} (fortran={fortran}), expected C-ordered {dst.dtype} {dst.shape}")
        buf = dst.reshape(-1).view(np.uint8)
        got = f.readinto(memoryview(buf))
        if got != buf.nbytes:
            raise ValueError(f"{path}: truncated ({got} of {buf.nbytes} bytes)")


def load_scene(ctx, lay, dataset, readers=8, ring=6):
    """Every cloud of the dataset once: descriptors [C,n,32,60] float32 and keypoints [C,n,3] float64 on the device, plus the
    cloud id -> arena slot map.  `readers` threads read the descriptor files into a ring of pinned host buffers (host peak =
    `ring` clouds); each buffer is copied host->device asynchronously on a side stream as soon as it is full and recycled when
    that copy has finished, so file reads, PCIe copies and (for the caller) the first batches overlap."""
    ids = list(dataset.pc_ids)
    slot = {pc: i for i, pc in enumerate(ids)}
    first = np.load(lay.yoho_desc(ids[0]), mmap_mode="r")
    n = first.shape[0]
    if first.shape != (n, 32, 60) or first.dtype != np.float32:
        raise ValueError(f"cloud {ids[0]}: {first.dtype} {first.shape}, expected float32 [n,32,60]")
    del first
    desc = torch.empty((len(ids), n, 32, 60), dtype=torch.float32, device=ctx.device)
    keys = torch.empty((len(ids), n, 3), dtype=torch.float64, device=ctx.device)
    ring = max(1, min(ring, len(ids)))
    on_gpu = torch.device(ctx.device).type == "cuda"            # the oracle-backed context of the CPU test suite has no streams to overlap
    cache = getattr(ctx, "_scene_ring", None)                    # pinning ~40 MB buffers costs tens of ms each: keep the ring with the context
    if cache is not None and cache[0] == (n, ring):
        pinned = cache[1]
    else:
        pinned = [torch.empty((n, 32, 60), dtype=torch.float32) for _ in range(ring)]
        if on_gpu:
            pinned = [p.pin_memory() for p in pinned]
        try:
            ctx._scene_ring = ((n, ring), pinned)
        except AttributeError:
            pass
    views = [p.numpy() for p in pinned]
    free = [torch.cuda.Event() for _ in range(ring)] if on_gpu else None
    side = torch.cuda.Stream(device=ctx.device) if on_gpu else None

    def read(i, b):
        try:
            _read_npy_into(lay.yoho_desc(ids[i]), views[b])
        except ValueError as e:
            raise ValueError(f"cloud {ids[i]}: {e} (the arena holds clouds of {n} keypoints; the reference's caches are 5000 per cloud)")
        return i, b

    with cf.ThreadPoolExecutor(max_workers=max(1, readers)) as pool:
        pending = {}
        nxt = 0
        for b in range(ring):                                   # prime the ring
            pending[b] = pool.submit(read, nxt, b); nxt += 1
        done_clouds = 0
        while done_clouds < len(ids):
            b = done_clouds % ring                              # buffers complete in submission order
            i, _ = pending.pop(b).result()
            if on_gpu:
                with torch.cuda.stream(side):
                    desc[i].copy_(pinned[b], non_blocking=True)
                    free[b].record(side)
            else:
                desc[i].copy_(pinned[b])
            k = dataset.get_kps(ids[i])
            if k.shape != (n, 3):
                raise ValueError(f"cloud {ids[i]}: {k.shape} keypoints, the arena holds clouds of {n} keypoints")
            keys[i].copy_(torch.from_numpy(np.ascontiguousarray(k, np.float64)))
            done_clouds += 1
            if nxt < len(ids):
                if on_gpu:
                    free[b].synchronize()                       # the copy out of this buffer has finished: refill it
                pending[b] = pool.submit(read, nxt, b); nxt += 1
    if on_gpu:
        torch.cuda.current_stream(ctx.device).wait_stream(side)
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


def register_scene(cfg, dataset, keynum=5000, max_iter=1000, batch_pairs=64, nn_mode=4, seed=0, writer_threads=4, ctx=None, shard=True,
                   readers=8):
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
    dist_on = shard and dist.is_available() and dist.is_initialized()      # shard=False: this rank registers the WHOLE dataset itself
    rank = dist.get_rank() if dist_on else 0
    world = dist.get_world_size() if dist_on else 1
    pairs_all = list(dataset.pair_ids)
    desc, keys, slot = load_scene(ctx, lay, dataset, readers=readers)
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
        host.write_trajectory(dataset, out_dir, poses if (lo, hi) == (0, len(pairs_all)) else None)
    return dict(lo=lo, hi=hi, poses=poses, recall=recall, n_matches=counts)
