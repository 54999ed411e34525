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
if os.environ.get("PDP_PHASE_TIMING"):
    ctx.enable_trace(256)
for rep in range(a.repeat):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    it = ctx.sp_run(a.iterations, 0.02, 100, True, generic=a.generic)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    done = int(it.item())
    if os.environ.get("PDP_PHASE_TIMING"):
        # profiling builds (-DPDP_PHASE_TIMING): thread 0 of every CTA accumulates clock64 cycles >> 10 per phase
        nct = 148 * int(os.environ.get("PDP_PHASE_TIMING"))
        trl = ctx._trace.reshape(-1)[:32].cpu().numpy().astype(np.float64) * 1024.0 / nct / max(done, 1)
        print("cycles/iteration/CTA: clause pass %.0f, barrier %.0f, variable pass %.0f, barrier %.0f, decide+barrier %.0f, "
              "local decimation %.0f, barrier %.0f, grid decimation + termination %.0f" % tuple(trl[16:24]))
        print("  phases (cycles/iteration/CTA): clause load %.0f, node %.0f, write-out %.0f | variable load %.0f, node %.0f, "
              "write-out %.0f" % tuple(trl[0:6]))
        print("  grid decimation (cycles/iteration/CTA): score %.0f, arg-max + fix %.0f, closure %.0f | CNF count %.0f, termination %.0f"
              % tuple(trl[24:29]))
        ctx._trace.zero_()
    print("E=%d iterations=%d  %.3f ms  %.3f ms/iter  %.2f G edge-updates/s  %.1f GB/s algorithmic" % (
        E, done, ms, ms / max(done, 1), E * done / ms / 1e6, 20.0 * E * done / ms / 1e6))

if os.environ.get("PDP_PROF_WALKSAT"):
    n_act = ctx.count_active_variables()
    ctx.random_fill(torch.rand(max(n_act, 1), device=dev))
    for W in (0, 1, 10, 100):
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pred, it = ctx.walksat(W, 0.5, None, None, seed=7)
            e1.record()
            torch.cuda.synchronize()
        print("walksat W=%d: %.3f ms (iterations done %d)" % (W, e0.elapsed_time(e1), int(it.item())))
        if os.environ.get("PDP_PHASE_TIMING") and ctx._trace is not None:
            tr = ctx._trace.reshape(-1)[:8].cpu().numpy().astype(np.float64) * 16.0 / 2 / max(W, 1)
            print("  CTA0 cycles per iteration: loop-top %.0f, select %.0f, finish %.0f, sync1 %.0f, check+sync2 %.0f, flip %.0f, sync3 %.0f" % tuple(tr[:7]))
            ctx._trace.zero_()
