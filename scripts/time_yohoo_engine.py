"""Throughput of the batched yohoo engine (mutual -> Des2R -> ET on <= 1000 scored hypotheses per pair -> one-shot RANSAC -> refine)."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, pipeline, synth
from oracle import roreg_oracle as O
ctx = ops.Context(0); ctx.set_corr_mode(1)
B, n = 8, 5000
prs = [synth.make_pair(300 + i, n=n, with_fcgf=True, max_res_deg=2.0) for i in range(B)]
desc = ctx.dev(np.stack([x for pr in prs for x in (pr["feats0"], pr["feats1"])]))
fcgf = ctx.dev(np.stack([x for pr in prs for x in (pr["fcgf0"], pr["fcgf1"])]))
keys = ctx.dev(np.stack([x for pr in prs for x in (pr["keys0"], pr["keys1"])]), torch.float64)
pc = ctx.dev(np.array([[2 * i, 2 * i + 1] for i in range(B)], np.int32))
for npass in (3, 1):
    eng = pipeline.YohooEngine(ctx, O.random_state_dict("ET", 102), npass=npass, max_iter=1000, ird=0.1, nn_mode=2)
    eng.register(desc, fcgf, keys, pc, seed=1); torch.cuda.synchronize()
    t = time.time(); reps = 5
    for r in range(reps): out = eng.register(desc, fcgf, keys, pc, seed=2 + r)
    torch.cuda.synchronize(); dt = (time.time() - t) / reps
    err = max(np.abs(out["poses"][i].cpu().numpy()[:3] - prs[i]["gt"]).max() for i in range(B))
    print(f"yohoo engine npass {npass}: {dt*1e3:.1f} ms per {B} pairs = {B/dt:.0f} pairs/s, max |pose-gt| {err:.2e}")
