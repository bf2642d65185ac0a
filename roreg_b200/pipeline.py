"""Throughput engine for the yohoo estimator (test/estimator.py:445-454): mutual matcher -> Des2R -> ET network on
the hypotheses that will actually be scored -> per-match SE(3) hypotheses -> one-shot RANSAC -> refine, for B pairs
per call with everything resident on the device.

The reference computes the ET network for all K matches and then scores only the first `max_iter` of a shuffled order
(test/estimator.py:424-425); here the shuffle comes first, so the network runs on min(K, max_iter) matches per pair -
the same hypotheses are scored, the unused ones are never computed.  torch is used for index bookkeeping only
(permutation keys, index gathers); the arithmetic is in libroreg_b200.so."""
import numpy as np
import torch

from . import nets


class YohooEngine:
    def __init__(self, ctx, et_state_dict, npass=3, max_iter=1000, ird=0.1, nn_mode=4, et_chunk=16000):
        self.ctx, self.max_iter, self.ird, self.nn_mode = ctx, max_iter, ird, nn_mode
        self.et = nets.ETNet(ctx, et_state_dict, npass=npass, chunk=et_chunk)

    def register(self, desc, fcgf, keys, pair_cloud, order=None, seed=0, events=None):
        """desc / fcgf [n_clouds,n,32,60] f32 (GF-out / FCGF-in descriptors), keys [n_clouds,n,3] f64, pair_cloud [B,2] int32.
        order (optional) [B,H] int64: hypothesis j of pair b is match order[b,j] (parity tests); default = device shuffle.
        events (optional list): receives (name, torch.cuda.Event) marks on the current stream after each stage (bench.py).
        Returns the register_batch output dict (poses, recall = index in the scored order, ...)."""
        def mark(name):
            if events is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); events.append((name, e))
        mark("start")
        ctx = self.ctx
        n_clouds, n = desc.shape[0], desc.shape[1]
        B = pair_cloud.shape[0]; H = self.max_iter
        out = ctx.register_batch(desc, keys, pair_cloud, max_iter=H, ird=self.ird, nn_mode=self.nn_mode, estimator=2)
        mark("matcher+des2r")
        K = out["n_matches"].to(torch.int64)                                     # [B]
        S = out["matches"].shape[1]
        if order is None:
            g = torch.Generator(device=ctx.device); g.manual_seed(int(seed))
            keys_r = torch.rand((B, S), device=ctx.device, generator=g)
            keys_r = torch.where(torch.arange(S, device=ctx.device)[None] < K[:, None], keys_r, torch.full_like(keys_r, 2.0))
            order = torch.argsort(keys_r, dim=1)[:, :H]                          # valid matches first, in random order
        n_hyp = torch.clamp(K, max=H).to(torch.int32)
        valid = torch.arange(order.shape[1], device=ctx.device)[None] < n_hyp[:, None].to(torch.int64)
        order = torch.where(valid, order, torch.zeros_like(order))
        if order.shape[1] < H:
            order = torch.cat([order, torch.zeros((B, H - order.shape[1]), dtype=order.dtype, device=ctx.device)], 1)
        m = torch.gather(out["matches"].to(torch.int64), 1, order[:, :, None].expand(B, H, 2))        # [B,H,2] keypoint ids (id0, id1)
        pc = pair_cloud.to(torch.int64)
        row0 = (pc[:, 0:1] * n + m[:, :, 0]).reshape(-1).to(torch.int32).contiguous()                   # rows of cloud id0 in the arenas
        row1 = (pc[:, 1:2] * n + m[:, :, 1]).reshape(-1).to(torch.int32).contiguous()
        pre = torch.gather(out["dr_index"].to(torch.int64), 1, order).reshape(-1).to(torch.int32).contiguous()
        d2 = desc.reshape(n_clouds * n, 32, 60); f2 = fcgf.reshape(n_clouds * n, 32, 60)
        # side 0 of the network = cloud id1 (test/estimator.py:293-306)
        mark("select")
        quat = self.et.forward(f2, row1, f2, row0, d2, row1, d2, row0, pre)
        mark("et_network")
        k2 = keys.reshape(n_clouds * n, 3)
        k0m = k2.index_select(0, row0.to(torch.int64)).contiguous(); k1m = k2.index_select(0, row1.to(torch.int64)).contiguous()
        hyps = ctx.hypotheses_from_quat(quat, pre, k0m, k1m).reshape(B, H, 3, 4)
        ctx.estimate_batch(out, hyps, n_hyp)
        mark("hypotheses+ransac+refine")
        out["order"] = order; out["n_hyp"] = n_hyp; out["hyps"] = hyps
        return out
