"""world_size-2 gloo test (CPU) of the N>1 path's host logic: balanced contiguous sharding of the pair list
and the final gather of poses in global pair order."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from roreg_b200 import shard
    lo, hi = shard.shard_pairs(n_pairs, rank, world)
    poses = torch.zeros((hi - lo, 4, 4), dtype=torch.float64)
    for i in range(lo, hi):
        poses[i - lo] = torch.eye(4, dtype=torch.float64) * (i + 1)
    recall = torch.arange(lo, hi, dtype=torch.int32)
    p, r = shard.gather_poses(poses, recall, n_pairs)
    q.put((rank, p.numpy(), r.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_slices_cover_everything():
    from roreg_b200 import shard
    for n in (0, 1, 7, 8, 1623):
        for w in (1, 2, 4, 8):
            s = [shard.shard_pairs(n, r, w) for r in range(w)]
            assert s[0][0] == 0 and s[-1][1] == n
            assert all(s[i][1] == s[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in s]
            assert max(sizes) - min(sizes) <= 1


def test_gather_poses_world2_gloo():
    n_pairs = 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    for rank, poses, recall in res:
        assert poses.shape == (n_pairs, 4, 4)
        assert np.array_equal(recall, np.arange(n_pairs))
        for i in range(n_pairs):
            assert np.array_equal(poses[i], np.eye(4) * (i + 1))
