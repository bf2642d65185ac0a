"""Wall-clock of roreg_group_corr_allpairs at N = M = 5000 (60 K-permuted tcgen05 GEMMs, 5.76 TFLOP algorithmic)."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from roreg_b200 import ops, nets, synth, _lib
from roreg_b200.ops import _ptr, _stream
ctx = ops.Context(0)
pr = synth.make_pair(2, n=5000)
N = M = 5000
for npass in (1, 3):
    g = nets.GroupNets(ctx, npass)
    xh, xl = g.pack([ctx.dev(pr["feats1"])], [None], [0], None, N); yh, yl = g.pack([ctx.dev(pr["feats0"])], [None], [0], None, M)
    best = torch.empty((N, M), dtype=torch.float32, device=ctx.device); ba = torch.empty((N, M), dtype=torch.uint8, device=ctx.device)
    nn = torch.empty(N, dtype=torch.int32, device=ctx.device); nna = torch.empty(N, dtype=torch.int32, device=ctx.device); nd = torch.empty(N, dtype=torch.float32, device=ctx.device)
    for rep in range(2):
        torch.cuda.synchronize(); t = time.time()
        rc = ctx.lib.roreg_group_corr_allpairs(ctx.h, _ptr(xh), _ptr(xl), N, _ptr(yh), _ptr(yl), M, npass, _ptr(best), _ptr(ba), _ptr(nn), _ptr(nna), _ptr(nd), _stream())
        _lib.check(ctx.h, rc, "allpairs"); torch.cuda.synchronize(); dt = time.time() - t
    flops = 2.0 * N * M * 1920 * 60
    ok = (pr["corr0"][nn.cpu().numpy()[pr["corr0"][:0].shape[0]:]] is not None)
    corr = pr["corr0"]; sel = np.where(corr >= 0)[0]
    # rows of cloud1 (X) whose partner in cloud0 is known: cloud0 row i pairs with cloud1 row corr0[i]
    inv = np.full(N, -1); inv[corr[sel]] = sel
    have = inv >= 0
    acc = (nn.cpu().numpy()[have] == inv[have]).mean(); rot = (nna.cpu().numpy()[have] == pr["a"]).mean()
    print(f"allpairs npass {npass}: {dt*1e3:.1f} ms  {flops/dt/1e12:.0f} TFLOP/s algorithmic  NN accuracy {acc:.3f}  rotation accuracy {rot:.3f}")
