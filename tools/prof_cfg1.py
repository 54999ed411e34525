#!/usr/bin/env python3
"""BASELINE.json configs[0] on the B200: p-d-p on 5000 x random 3-SAT n=100, alpha=4.2 (+ WalkSAT): time and solved count."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdp_solver_b200 import cnfgen
from pdp_solver_b200.nn import solver as S

ap = argparse.ArgumentParser()
ap.add_argument("--problems", type=int, default=5000)
ap.add_argument("--n", type=int, default=100)
ap.add_argument("--alpha", type=float, default=4.2)
ap.add_argument("--iterations", type=int, default=1000)
ap.add_argument("--walksat", type=int, default=100)
a = ap.parse_args()
dev = torch.device("cuda:0")
gm, bvm, bfm, ef = [torch.from_numpy(x).to(dev) for x in cnfgen.random_batch(a.problems, a.n, 3, a.alpha, 1000)]
model = S.SurveyPropagatorSolver(dev, "p-d-p", tolerance=0.02, t_max=100, local_search_iterations=a.walksat, epsilon=0.5)
def term(*x): raise RuntimeError
term._pdp_standard_termination = True
os.environ["PDP_B200_TIMING"] = "1"
for rep in range(2):
    torch.manual_seed(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    init = model.get_init_state(gm, bvm, bfm, ef, None, randomized=False, batch_replication=1)
    (pred, _), _ = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm, edge_feature=ef,
                         meta_data=None, is_training=False, iteration_num=a.iterations, check_termination=term, batch_replication=1)
    ctx = model.last_problem._ctx
    solved, nun = ctx.cnf_eval(pred)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if os.environ.get("PDP_PHASE_TIMING") and ctx._trace is not None:
        tr = ctx._trace.reshape(-1)[:32].cpu().numpy().astype(np.float64) * 1024.0 / 148 / max(int(model.last_iterations.item()), 1)
        print("  cycles/iter/CTA: clause pass %.0f, sync %.0f, var pass %.0f, sync %.0f, decide+sync %.0f, local decimation %.0f, sync %.0f, rest %.0f" % tuple(tr[16:24]))
    flags, counters, freeze = ctx.problem_flags()
    print("rep %d: %.1f ms, iterations %d, solved %d / %d, phases %s, flags: trivial %d solved-in-loop %d contradiction %d" % (
        rep, dt * 1e3, int(model.last_iterations.item()), int(solved.sum().item()), a.problems, ctx.timing,
        int((flags & 1).ne(0).sum()), int((flags & 2).ne(0).sum()), int((flags & 4).ne(0).sum())))
