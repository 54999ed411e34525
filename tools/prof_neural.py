#!/usr/bin/env python3
"""BASELINE.json configs[1] / configs[2] on the B200: p-nd-np on random 3-SAT n = 10 000 (alpha 4.2) and np-nd-np on random
4-SAT n = 1 000 (alpha 9.0), random-init weights under torch.manual_seed(1), dims of config/Predict/PDP-np-nd-np-*.yaml
(hidden 150 / mem 100 / agg 100 / mem_agg 50 / classifier 50).  Prints seconds per iteration and edge-updates/s."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdp_solver_b200 import cnfgen
from pdp_solver_b200.nn import solver as S, util as U

ap = argparse.ArgumentParser()
ap.add_argument("--iterations", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
H, MH, AH, MAH, CH = 150, 100, 100, 50, 50


def term(*x):
    raise RuntimeError
term._pdp_standard_termination = True
for name, mt, B, n, k, alpha in (("configs[1] p-nd-np", "p-nd-np", 8, 10000, 3, 4.2), ("configs[2] np-nd-np", "np-nd-np", 32, 1000, 4, 9.0)):
    torch.manual_seed(1)
    clf = U.Perceptron(H, CH, 1)
    if mt == "p-nd-np":
        model = S.NeuralSurveyPropagatorSolver(dev, "m", 1, 0, H, MH, AH, MAH, 1, variable_classifier=clf, local_search_iterations=100, epsilon=0.5)
    else:
        model = S.NeuralPropagatorDecimatorSolver(dev, "m", 1, 0, H, H, MH, AH, MAH, 1, variable_classifier=clf, local_search_iterations=100, epsilon=0.5)
    model = model.to(dev).eval()
    gm, bvm, bfm, ef = [torch.from_numpy(x).to(dev) for x in cnfgen.random_batch(B, n, k, alpha, 2000)]
    E = gm.shape[1]
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.no_grad():
            init = model.get_init_state(gm, bvm, bfm, ef, None, randomized=False, batch_replication=1)
            (pred, _), _ = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm, edge_feature=ef,
                                 meta_data=None, is_training=False, iteration_num=a.iterations, check_termination=term, batch_replication=1)
        solved, _ = model.last_problem._ctx.cnf_eval(pred)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    it = int(model.last_iterations.item())
    print("%s: B=%d n=%d k=%d E=%d, %d iterations: %.1f ms total, %.2f ms/iteration, %.1f M edge-updates/s, solved %d/%d, params %d" % (
        name, B, n, k, E, it, dt * 1e3, dt * 1e3 / max(it, 1), E * it / dt / 1e6, int(solved.sum().item()), B, model.parameter_count()))
