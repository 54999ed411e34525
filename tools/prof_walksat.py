import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from pdp_solver_b200 import cnfgen
from pdp_solver_b200.engine import Context
dev = torch.device("cuda:0")
gm, bvm, bfm, ef = cnfgen.random_batch(4, 1000000, 3, 4.2, 1234)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
ctx = Context(t(gm), t(bvm), t(bfm), t(ef), batch_size=4)
ctx.simplify()
n_act = ctx.count_active_variables()
ctx.random_fill(torch.rand(max(n_act, 1), device=dev))
pred, it = ctx.walksat(6, 0.5, None, None, seed=7, sync=True)
print("done", it)
