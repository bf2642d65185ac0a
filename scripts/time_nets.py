"""GPU-box timing of the group-convolution networks (CUDA events, resident inputs, random weights of the checkpoint shapes):
GF on one 5000-keypoint cloud (network/group_feat.py:26-45), ET on 3400 matches (network/eqv_trans.py:119-138), RD on 5000 keypoints,
Match_ot forward on 5000 x 5000, all-pairs 60-rotation correlation.   python scripts/time_nets.py > gpurun_out/nets_timing.txt"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from roreg_b200 import ops, nets, synth, matchot

ctx = ops.Context(0)
rng = np.random.default_rng(0)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    l0 = ctx.launches
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (ctx.launches - l0) / reps


x = rng.standard_normal((5000, 32, 60)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True)
xd = ctx.dev(x)
for npass in (1, 3):
    gf = nets.GFNet(ctx, synth.random_weights("GF", 101), npass=npass)
    ms, ln = timed(lambda: gf.forward(xd))
    print(f"GF 5000 keypoints, npass {npass}: {ms:.2f} ms, {ln:.0f} launches, {2.17e12 / (ms * 1e-3) / 1e12:.0f} TFLOP/s algorithmic (434 MFLOP/keypoint)")
    K = 3400
    rows = ctx.dev(rng.integers(0, 5000, K).astype(np.int32)); pre = ctx.dev(rng.integers(0, 60, K).astype(np.int32))
    et = nets.ETNet(ctx, synth.random_weights("ET", 102), npass=npass, chunk=4000)
    ms, ln = timed(lambda: et.forward(xd, rows, xd, rows, xd, rows, xd, rows, pre))
    print(f"ET {K} matches, npass {npass}: {ms:.2f} ms, {ln:.0f} launches, {K * 99.2e6 / (ms * 1e-3) / 1e12:.0f} TFLOP/s algorithmic (99.2 MFLOP/match pruned)")
    rd = nets.RDNet(ctx, synth.random_weights("RD", 103), npass=npass)
    ms, ln = timed(lambda: rd.forward(xd))
    print(f"RD 5000 keypoints, npass {npass}: {ms:.2f} ms, {ln:.0f} launches")
    pr = synth.make_pair(2, n=5000)
    f0 = ctx.dev(pr["feats0"]); f1 = ctx.dev(pr["feats1"]); k0 = ctx.dev(pr["keys0"].astype(np.float32)); k1 = ctx.dev(pr["keys1"].astype(np.float32))
    mo = matchot.MatchOT(ctx, synth.random_weights("RM", 104), npass=npass)
    ms, ln = timed(lambda: mo.forward(f1, f0, k1, k0), reps=3, warm=1)
    print(f"Match_ot 5000 x 5000, npass {npass}: {ms:.2f} ms, {ln:.0f} launches")
    g = nets.GroupNets(ctx, npass)
    n_all = 5000
    Xp = g.pack([f1[:n_all].contiguous()], [None], [0], None, n_all); Yp = g.pack([f0[:n_all].contiguous()], [None], [0], None, n_all)
    best = torch.empty((n_all, n_all), dtype=torch.float32, device=ctx.device); ba = torch.empty((n_all, n_all), dtype=torch.uint8, device=ctx.device)

    def allpairs():
        rc = ctx.lib.roreg_group_corr_allpairs(ctx.h, ops._ptr(Xp[0]), ops._ptr(Xp[1]), n_all, ops._ptr(Yp[0]), ops._ptr(Yp[1]), n_all, npass,
                                               ops._ptr(best), ops._ptr(ba), None, None, None, ops._stream())
        assert rc == 0
    ms, ln = timed(allpairs, reps=2, warm=1)
    print(f"all-pairs 60-rotation correlation 5000 x 5000, npass {npass}: {ms:.2f} ms, {ln:.0f} launches, {5.76e12 / (ms * 1e-3) / 1e12:.0f} TFLOP/s algorithmic")
