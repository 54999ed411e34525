#!/usr/bin/env python3
"""Graph ingest only (pdp_create + simplify + constant state load) at the bench size, for a launch list:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/create.csv python tools/prof_create.py"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdp_solver_b200 import cnfgen  # noqa: E402
from pdp_solver_b200.engine import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--problems", type=int, default=8)
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
gm, bvm, bfm, ef = [torch.from_numpy(x).to(dev) for x in cnfgen.random_batch(a.problems, a.n, 3, 4.2, 1)]
for rep in range(a.reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record()
    ctx = Context(gm, bvm, bfm, ef, batch_size=a.problems)
    e1.record()
    ctx.simplify()
    ctx.load_state_const(1 / 3, 1 / 3, 1 / 3, 0.5, 0.0)
    e2.record()
    torch.cuda.synchronize()
    print("rep %d: create %.2f ms, simplify+load %.2f ms, wall %.2f ms" % (rep, e0.elapsed_time(e1), e1.elapsed_time(e2),
                                                                           (time.perf_counter() - t0) * 1e3))
    del ctx
