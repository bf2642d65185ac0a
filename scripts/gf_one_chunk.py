"""One GF chunk (500 keypoints) for ncu captures of the implicit group-convolution GEMM layers (GPU box only)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, nets, synth
ctx = ops.Context(0)
npass = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rng = np.random.default_rng(0)
x = rng.standard_normal((500, 32, 60)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True)
gf = nets.GFNet(ctx, synth.random_weights("GF", 101), npass=npass, chunk=500)
for _ in range(2):
    gf.forward(ctx.dev(x))
torch.cuda.synchronize()
