import numpy as np, torch, sys
sys.path.insert(0, '.')
from roreg_b200 import ops, synth
c = ops.Context(0); c.set_corr_mode(2)
pr = synth.make_pair(5, n=1200)
rng = np.random.default_rng(1)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 500
ix = rng.integers(0, 1200, K).astype(np.int32); iy = rng.integers(0, 1200, K).astype(np.int32)
cor, am = c.group_corr(c.dev(pr["feats1"]), c.dev(pr["feats0"]), c.dev(ix), c.dev(iy), 1)
torch.cuda.synchronize()
print("ok", am[:8].tolist())
