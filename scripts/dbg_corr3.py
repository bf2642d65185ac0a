import numpy as np, torch, sys, time
sys.path.insert(0, '.')
from roreg_b200 import ops, synth
B = int(sys.argv[1]); n = int(sys.argv[2]); nn_mode = int(sys.argv[3]) if len(sys.argv) > 3 else 2
c = ops.Context(0); c.set_corr_mode(3)
prs = [synth.make_pair(100 + i, n=n) for i in range(min(B, 4))]
prs = [prs[i % len(prs)] for i in range(B)]
desc = c.dev(np.stack([x for pr in prs for x in (pr["feats0"], pr["feats1"])]))
keys = c.dev(np.stack([x for pr in prs for x in (pr["keys0"], pr["keys1"])]), torch.float64)
pc = c.dev(np.array([[2 * i, 2 * i + 1] for i in range(B)], np.int32))
for rep in range(3):
    o = c.register_batch(desc, keys, pc, max_iter=300, seed=3, nn_mode=nn_mode)
    torch.cuda.synchronize()
print("ok B", B, "n", n, "matches", o["n_matches"].tolist()[:4])
