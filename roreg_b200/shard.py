"""Pair sharding across the GPUs of one box.

Pairs are independent units (the reference loops `for pair in dataset.pair_ids`, test/matcher.py:64,
test/estimator.py:102), so each rank registers a disjoint slice with NO data-path collective; the only
collective is the final gather of the [n_local,4,4] float64 poses (+ the winning-hypothesis index) -
NCCL over NVLink on GPUs, gloo in the CPU tests.  The reference has no multi-GPU path to mirror."""
import numpy as np
import torch
import torch.distributed as dist


def shard_pairs(n_pairs, rank, world):
    """Contiguous balanced slice [lo, hi) of the global pair list for this rank (sizes differ by <= 1)."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_poses(poses_local, recall_local, n_pairs):
    """all_gather of variable-length per-rank results, returned in global pair order on every rank.
    poses_local [n_local,4,4] float64, recall_local [n_local] int32 (tensors on the backend's device)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return poses_local, recall_local
    world = dist.get_world_size()
    cap = (n_pairs + world - 1) // world
    dev = poses_local.device
    pad_p = torch.zeros((cap, 4, 4), dtype=torch.float64, device=dev); pad_p[:poses_local.shape[0]] = poses_local
    pad_r = torch.full((cap,), -2, dtype=torch.int32, device=dev); pad_r[:recall_local.shape[0]] = recall_local
    all_p = torch.empty((world * cap, 4, 4), dtype=torch.float64, device=dev)
    all_r = torch.empty((world * cap,), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_p, pad_p) if dev.type == "cuda" else dist.all_gather(list(all_p.view(world, cap, 4, 4).unbind(0)), pad_p)
    dist.all_gather_into_tensor(all_r, pad_r) if dev.type == "cuda" else dist.all_gather(list(all_r.view(world, cap).unbind(0)), pad_r)
    keep = torch.cat([torch.arange(r * cap, r * cap + (shard_pairs(n_pairs, r, world)[1] - shard_pairs(n_pairs, r, world)[0]))
                      for r in range(world)]).to(dev)
    return all_p[keep], all_r[keep]


def register_pairs_sharded(ctx, desc, keys, pair_cloud_all, **kw):
    """Register this rank's slice of `pair_cloud_all` [P,2] (host int32 array) through the batched engine and
    gather all poses.  desc/keys are the rank-local device arenas holding (at least) the clouds its pairs touch."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_pairs(len(pair_cloud_all), rank, world)
    pc = ctx.dev(np.ascontiguousarray(pair_cloud_all[lo:hi], np.int32))
    out = ctx.register_batch(desc, keys, pc, **kw)
    return gather_poses(out["poses"], out["recall"], len(pair_cloud_all))
