"""Whole-dataset driver on the batched engine - SURVEY.md section 8(f) rank 3 (stage I/O).

The reference registers a dataset as two plugin passes, `mutual.run` then `yohoc.run` (test/evaluator.py:23-48), and every pass
re-reads both clouds' 38 MB descriptor files for every PAIR (test/matcher.py:66-67, test/estimator.py:106-107).  `register_scene`
does the same work in one pass: every cloud is read and uploaded ONCE into a device arena (SceneLoader: reader threads -> ring
of pinned buffers -> side stream), the pairs go through `roreg_register_batch` B at a time in the order their clouds arrive
(sharded over ranks when torch.distributed is initialised: pairs are independent, no data-path collective), and the reference's
on-disk contract - match / scores / DR_index files, `{id0}-{id1}.npz`, `pre.log` - is emitted by background writer
threads through a plain-C writer (roreg_write_pair_files), so that file reads, registration and file writes overlap.

Same algorithm and files as `mutual` + `yohoc` of roreg_b200/test with `yohoc_mode='device'`: keypoint sampling consumes the
global NumPy RNG in the reference's order (one rank) or comes from the NMS sampler with --RD; the RANSAC triplets are drawn on
the device (counter-based RNG), so poses agree with the host-RNG plugins statistically, not bit for bit.  --RM (Match_ot
matcher) and the yohoo estimator are per-pair network paths and stay with their plugins.
"""
import concurrent.futures as cf
import os
import threading
import numpy as np
import torch
import torch.distributed as dist
from . import _lib
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


def _open_npy(path, dst):
    """Open a C-ordered .npy file whose array must match `dst` (dtype, shape); returns (fd, payload offset).  Header parsed with
    numpy.lib.format."""
    with open(path, "rb") as f:
        major, _ = np.lib.format.read_magic(f)
        shape, fortran, dtype = (np.lib.format.read_array_header_1_0 if major == 1 else np.lib.format.read_array_header_2_0)(f)
        if fortran or dtype != dst.dtype or tuple(shape) != tuple(dst.shape):
            raise ValueError(f"{path}: {dtype} {shape} (fortran={fortran}), expected C-ordered {dst.dtype} {dst.shape}")
        return os.open(path, os.O_RDONLY), f.tell()


def _pread_into(fd, buf, off, path):
    """pread the whole of `buf` (a uint8 NumPy view) from file offset `off`; the GIL is released while the kernel copies."""
    mv, done = memoryview(buf), 0
    while done < len(mv):
        got = os.preadv(fd, [mv[done:]], off + done)
        if got <= 0:
            raise ValueError(f"{path}: truncated ({off + done} of {off + len(mv)} bytes)")
        done += got


READ_CHUNK = 8 << 20      # bytes per reader task: one ~40 MB cloud is split over five threads, so clouds arrive in order and early


def _read_npy_into(path, dst):
    """Read a C-ordered .npy file straight into `dst` (a NumPy view of pinned host memory): one copy (page cache -> pinned
    buffer).  Single-threaded form of what SceneLoader does with its reader pool."""
    fd, off = _open_npy(path, dst)
    try:
        _pread_into(fd, dst.reshape(-1).view(np.uint8), off, path)
    finally:
        os.close(fd)


class SceneLoader(threading.Thread):
    """Every cloud of the dataset once, in `pc_ids` order: descriptors [C,n,32,60] float32 and keypoints [C,n,3] float64 on the
    device.  `readers` threads read the descriptor files into a ring of pinned host buffers (host peak = `ring` clouds); this
    thread copies each buffer host->device on a side stream as soon as it is full, records a per-cloud event and recycles the
    buffer when its copy has finished.  `wait(k)` returns when clouds [0, k) are uploaded or enqueued for upload and makes the
    caller's current stream wait for them - so the first batches are registered (and their files written) while later clouds are
    still being read."""

    def __init__(self, ctx, lay, dataset, readers=8, ring=6):
        super().__init__(daemon=True)
        self.ctx, self.lay, self.dataset, self.readers = ctx, lay, dataset, max(1, readers)
        self.ids = list(dataset.pc_ids)
        self.slot = {pc: i for i, pc in enumerate(self.ids)}
        first = np.load(lay.yoho_desc(self.ids[0]), mmap_mode="r")
        self.n = n = first.shape[0]
        if first.shape != (n, 32, 60) or first.dtype != np.float32:
            raise ValueError(f"cloud {self.ids[0]}: {first.dtype} {first.shape}, expected float32 [n,32,60]")
        del first
        self.desc = torch.empty((len(self.ids), n, 32, 60), dtype=torch.float32, device=ctx.device)
        self.keys = torch.empty((len(self.ids), n, 3), dtype=torch.float64, device=ctx.device)
        ring = max(1, min(ring, len(self.ids)))
        self.on_gpu = torch.device(ctx.device).type == "cuda"   # the oracle-backed context of the CPU test suite has no streams to overlap
        cache = getattr(ctx, "_scene_ring", None)                # pinning ~40 MB buffers costs tens of ms each: keep the ring with the context
        if cache is not None and cache[0] == (n, ring):
            self.pinned = cache[1]
        else:
            self.pinned = [torch.empty((n, 32, 60), dtype=torch.float32) for _ in range(ring)]
            if self.on_gpu:
                self.pinned = [p.pin_memory() for p in self.pinned]
            try:
                ctx._scene_ring = ((n, ring), self.pinned)
            except AttributeError:
                pass
        self.side = torch.cuda.Stream(device=ctx.device) if self.on_gpu else None
        self.uploaded = [torch.cuda.Event() for _ in self.ids] if self.on_gpu else None
        self.cv = threading.Condition()
        self.count = 0
        self.error = None

    def run(self):
        try:
            self._load()
        except BaseException as e:                              # surfaces in wait()
            with self.cv:
                self.error = e
                self.cv.notify_all()

    def _load(self):
        ids, n, ring = self.ids, self.n, len(self.pinned)
        views = [p.numpy() for p in self.pinned]
        free = [torch.cuda.Event() for _ in range(ring)] if self.on_gpu else None
        if self.on_gpu:
            torch.cuda.set_device(self.ctx.device)

        def submit(pool, i, b):
            """Split cloud i's file into READ_CHUNK tasks filling pinned buffer b; returns (fd, futures)."""
            path = self.lay.yoho_desc(ids[i])
            try:
                fd, off = _open_npy(path, views[b])
            except ValueError as e:
                raise ValueError(f"cloud {ids[i]}: {e} (the arena holds clouds of {n} keypoints; the reference's caches are 5000 per cloud)")
            flat = views[b].reshape(-1).view(np.uint8)
            return fd, [pool.submit(_pread_into, fd, flat[o:o + READ_CHUNK], off + o, path) for o in range(0, flat.nbytes, READ_CHUNK)]

        with cf.ThreadPoolExecutor(max_workers=self.readers) as pool:
            pending = {}
            nxt = 0
            try:
                for b in range(ring):                               # prime the ring (tasks run in submission order: cloud 0 first)
                    pending[b] = submit(pool, nxt, b); nxt += 1
                for done in range(len(ids)):
                    b = done % ring                                 # buffers complete in submission order
                    fd, futs = pending.pop(b)
                    try:
                        for f in futs:
                            f.result()
                    finally:
                        os.close(fd)
                    i = done
                    k = self.dataset.get_kps(ids[i])
                    if k.shape != (n, 3):
                        raise ValueError(f"cloud {ids[i]}: {k.shape} keypoints, the arena holds clouds of {n} keypoints")
                    kt = torch.from_numpy(np.ascontiguousarray(k, np.float64))
                    if self.on_gpu:
                        with torch.cuda.stream(self.side):
                            self.desc[i].copy_(self.pinned[b], non_blocking=True)
                            self.keys[i].copy_(kt)
                            free[b].record(self.side)
                            self.uploaded[i].record(self.side)
                    else:
                        self.desc[i].copy_(self.pinned[b]); self.keys[i].copy_(kt)
                    with self.cv:
                        self.count = done + 1
                        self.cv.notify_all()
                    if nxt < len(ids):
                        if self.on_gpu:
                            free[b].synchronize()                   # the copy out of this buffer has finished: refill it
                        pending[b] = submit(pool, nxt, b); nxt += 1
            finally:
                for fd, futs in pending.values():                   # error path: let the outstanding reads finish, then close
                    cf.wait(futs)
                    os.close(fd)

    def wait(self, k):
        """Block until clouds [0, k) are enqueued for upload; the current stream then waits for the last of them."""
        k = min(k, len(self.ids))
        with self.cv:
            while self.count < k and self.error is None:
                self.cv.wait()
            if self.error is not None:
                raise self.error
        if self.on_gpu and k > 0:
            torch.cuda.current_stream(self.ctx.device).wait_event(self.uploaded[k - 1])     # copies are ordered on the side stream


def load_scene(ctx, lay, dataset, readers=8, ring=6):
    """All clouds at once: (desc, keys, cloud id -> arena slot)."""
    ld = SceneLoader(ctx, lay, dataset, readers, ring)
    ld.start(); ld.wait(len(ld.ids)); ld.join()
    return ld.desc, ld.keys, ld.slot


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
    """The pair's four files through roreg_write_pair_files (plain C: the ctypes call releases the GIL, so writer threads run in
    parallel; byte-identical .npy files to np.save, an .npz np.load reads back - tests/test_dataio.py)."""
    lib = _lib.load()
    matches = np.ascontiguousarray(matches, np.int64); dr_index = np.ascontiguousarray(dr_index, np.int64)
    pose = np.ascontiguousarray(pose, np.float64)
    rc = lib.roreg_write_pair_files(lay.matches(id0, id1).encode(), lay.scores(id0, id1).encode(), lay.dr_index(id0, id1).encode(),
                                    lay.result('yohoc', max_iter, id0, id1).encode(), matches.ctypes.data, dr_index.ctypes.data,
                                    int(matches.shape[0]), pose.ctypes.data, int(recall))
    if rc != 0:
        raise OSError(f"pair {id0}-{id1}: could not write the result files under {lay.match_dir} (status {rc})")


def register_scene(cfg, dataset, keynum=5000, max_iter=1000, batch_pairs=64, nn_mode=4, seed=0, writer_threads=2, ctx=None, shard=True,
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
    loader = SceneLoader(ctx, lay, dataset, readers=readers)
    desc, keys, slot, n = loader.desc, loader.keys, loader.slot, loader.n
    if keynum > n:
        raise ValueError(f"keynum {keynum} exceeds the {n} keypoints per cloud")
    loader.start()                                                              # file reads + uploads run from here on
    samples_all = draw_samples(cfg, lay, dataset, pairs_all, n, keynum)        # all pairs on every rank: one RNG order for any world size
    lo, hi = shard_pairs(len(pairs_all), rank, world)
    poses = np.zeros((hi - lo, 4, 4)); recall = np.zeros(hi - lo, np.int64); counts = np.zeros(hi - lo, np.int64)
    whole = (lo, hi) == (0, len(pairs_all))                                    # this rank sees every pose: pre.log text is built pair by pair
    blocks = [None] * (hi - lo) if whole else None
    n_clouds = len(dataset.pc_ids)
    # pairs are registered in the order their clouds arrive (a pair is ready when its later cloud is uploaded); results and files
    # are indexed by the pair's position in dataset.pair_ids, so the order is invisible outside
    order = sorted(range(lo, hi), key=lambda p: max(slot[pairs_all[p][0]], slot[pairs_all[p][1]]))
    writer = AsyncWriter(writer_threads)
    try:
        for s in range(0, len(order), batch_pairs):
            idx = order[s:s + batch_pairs]
            batch = [pairs_all[p] for p in idx]
            loader.wait(1 + max(max(slot[a], slot[b]) for a, b in batch))
            pc = ctx.dev(np.array([[slot[a], slot[b]] for a, b in batch], np.int32))
            o = ctx.register_batch(desc, keys, pc, keynum=keynum, sample=ctx.dev(samples_all[idx]), nn_mode=nn_mode, estimator=0,
                                   max_iter=max_iter, ird=cfg.ransac_ird, seed=(int(seed) * 1000003 + lo + s) & 0x7fffffffffffffff)
            h = {k: v.cpu().numpy() for k, v in o.items()}                        # one synchronisation per batch
            for j, (id0, id1) in enumerate(batch):
                p = idx[j]
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
                poses[p - lo] = T; recall[p - lo] = r; counts[p - lo] = k
                if whole:
                    blocks[p] = host.trajectory_block(id0, id1, n_clouds, T)
                writer.submit(_write_pair, lay, max_iter, id0, id1, m, dr, T, r)
        loader.wait(len(loader.ids))
        loader.join()
    finally:
        writer.close()
    if dist_on:
        dist.barrier()
    if rank == 0:
        host.write_trajectory(dataset, out_dir, blocks=blocks)
    return dict(lo=lo, hi=hi, poses=poses, recall=recall, n_matches=counts)
