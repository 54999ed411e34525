#!/usr/bin/env python3
"""Profiling driver: one batch of random k-SAT resident on cuda:0, T iterations of the persistent SP kernel.
    ncu --set full --import-source on -k regex:k_sp_run -c 1 -o gpurun_out/prof python tools/prof_sweep.py
Prints the CUDA-event time of the launch (meaningless under ncu)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdp_solver_b200 import cnfgen  # noqa: E402
from pdp_solver_b200.engine import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--problems", type=int, default=2)
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--alpha", type=float, default=4.2)
ap.add_argument("--iterations", type=int, default=10)
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--generic", action="store_true")
a = ap.parse_args()

dev = torch.device("cuda:0")
gm, bvm, bfm, ef = cnfgen.random_batch(a.problems, a.n, a.k, a.alpha, 1234)
E = gm.shape[1]
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
ctx = Context(t(gm), t(bvm), t(bfm), t(ef), batch_size=a.problems)
errs, info = ctx.check_layout()
print("layout", info, "errs", errs)
ctx.simplify()
q3 = torch.full((E, 3), 1.0 / 3.0, device=dev)
fs2 = torch.zeros((E, 2), device=dev)
fs2[:, 0] = 0.5
ctx.load_state((q3, fs2), (q3, fs2))
for rep in range(a.repeat):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    it = ctx.sp_run(a.iterations, 0.02, 100, True, generic=a.generic)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    done = int(it.item())
    print("E=%d iterations=%d  %.3f ms  %.3f ms/iter  %.2f G edge-updates/s  %.1f GB/s algorithmic" % (
        E, done, ms, ms / max(done, 1), E * done / ms / 1e6, 20.0 * E * done / ms / 1e6))
