import sys, time, numpy as np, torch
sys.path.insert(0,".")
from roreg_b200 import ops, matchot, nets, synth
from oracle import roreg_oracle as O
ctx=ops.Context(0)
pr=synth.make_pair(2,n=5000,with_fcgf=True)
f0=ctx.dev(pr["feats0"]); f1=ctx.dev(pr["feats1"]); k0=ctx.dev(pr["keys0"].astype(np.float32)); k1=ctx.dev(pr["keys1"].astype(np.float32))
for npass in (3,1):
    mo=matchot.MatchOT(ctx,O.random_state_dict("RM",104),npass=npass)
    mo.forward(f1,f0,k1,k0); torch.cuda.synchronize()
    l0=ctx.launches; t=time.time(); m0,s0=mo.forward(f1,f0,k1,k0); torch.cuda.synchronize(); dt=time.time()-t
    print("Match_ot 5000x5000 npass",npass,"ms",dt*1e3,"launches",ctx.launches-l0,"matches",int((m0>=0).sum()))
