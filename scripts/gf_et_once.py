"""One GF forward (5000 keypoints) and one ET forward (3400 matches), 1 pass, after a warm-up - for ncu launch lists (GPU box only)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, nets, synth
ctx = ops.Context(0); rng = np.random.default_rng(0)
npass = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = rng.standard_normal((5000, 32, 60)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True); xd = ctx.dev(x)
gf = nets.GFNet(ctx, synth.random_weights("GF", 101), npass=npass)
K = 3400
rows = ctx.dev(rng.integers(0, 5000, K).astype(np.int32)); pre = ctx.dev(rng.integers(0, 60, K).astype(np.int32))
et = nets.ETNet(ctx, synth.random_weights("ET", 102), npass=npass, chunk=4000)
for _ in range(2):
    gf.forward(xd); et.forward(xd, rows, xd, rows, xd, rows, xd, rows, pre)
torch.cuda.synchronize()
