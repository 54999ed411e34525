#!/usr/bin/env python3
"""Stage-by-stage CUDA-event timing of one forward() of the p-d-p solver (what bench.py's `other` is made of)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdp_solver_b200 import cnfgen  # noqa: E402
from pdp_solver_b200.engine import Context  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--problems", type=int, default=8)
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--alpha", type=float, default=4.2)
ap.add_argument("--iterations", type=int, default=100)
ap.add_argument("--module-only", action="store_true", help="skip the raw Context stages (launch lists of the module path)")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda:0")
gm, bvm, bfm, ef = [torch.from_numpy(x).to(dev) for x in cnfgen.random_batch(a.problems, a.n, a.k, a.alpha, 1)]
E = gm.shape[1]


class Stages(object):
    def __init__(self):
        self.ev, self.names, self.t0 = [], [], time.perf_counter()

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.ev.append(e)
        self.names.append(name)

    def report(self):
        torch.cuda.synchronize()
        wall = time.perf_counter() - self.t0
        for i in range(1, len(self.ev)):
            print("  %-28s %8.3f ms" % (self.names[i], self.ev[i - 1].elapsed_time(self.ev[i])))
        print("  %-28s %8.3f ms (wall %.3f ms)" % ("total", self.ev[0].elapsed_time(self.ev[-1]), wall * 1e3))


for rep in range(0 if a.module_only else a.reps):
    torch.cuda.synchronize()
    st = Stages()
    st.mark("start")
    ctx = Context(gm, bvm, bfm, ef, batch_size=a.problems)
    st.mark("Context (pdp_create)")
    ctx.simplify()
    st.mark("simplify")
    q3 = torch.full((E, 3), 1.0 / 3.0, device=dev)
    fs2 = torch.zeros((E, 2), device=dev)
    fs2[:, 0] = 0.5
    st.mark("init state tensors")
    ctx.load_state((q3, fs2), (q3, fs2))
    st.mark("load_state")
    ctx.sp_run(a.iterations, 0.02, 100, True)
    st.mark("sp_run")
    ctx.store_state()
    st.mark("store_state")
    n_act = ctx.count_active_variables()
    st.mark("count_active (sync)")
    ctx.random_fill(torch.rand(max(n_act, 1), device=dev))
    st.mark("random_fill")
    pred, it = ctx.walksat(100, 0.5, None, None, seed=3)
    st.mark("walksat")
    ctx.cnf_eval(pred)
    st.mark("cnf_eval")
    print("rep", rep)
    st.report()
    del ctx

# ---- the same through the reference-interface module, as bench.py's one_step does
from pdp_solver_b200.nn import solver as pdp_solver  # noqa: E402
model = pdp_solver.SurveyPropagatorSolver(dev, "p-d-p", tolerance=0.02, t_max=100, local_search_iterations=100, epsilon=0.5)


def termination(active, prediction, sat_problem):
    raise RuntimeError("unreachable")


termination._pdp_standard_termination = True
for rep in range(a.reps):
    torch.cuda.synchronize()
    st = Stages()
    st.mark("start")
    torch.manual_seed(1)
    init = model.get_init_state(gm, bvm, bfm, ef, None, randomized=False, batch_replication=1)
    st.mark("get_init_state")
    (pred, _), _ = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm, edge_feature=ef,
                         meta_data=None, is_training=False, iteration_num=a.iterations, check_termination=termination,
                         batch_replication=1)
    st.mark("model.forward")
    ctx = model.last_problem._ctx
    solved, _ = ctx.cnf_eval(pred)
    st.mark("cnf_eval")
    _, _, freeze = ctx.problem_flags()
    iters = int(model.last_iterations.item())
    st.mark("flags + item")
    print("module rep", rep)
    st.report()
