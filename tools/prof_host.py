import cProfile, pstats, os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from pdp_solver_b200 import cnfgen
from pdp_solver_b200.nn import solver as S
dev = torch.device("cuda:0")
gm, bvm, bfm, ef = [torch.from_numpy(x).to(dev) for x in cnfgen.random_batch(5000, 100, 3, 4.2, 1000)]
model = S.SurveyPropagatorSolver(dev, "p-d-p", tolerance=0.02, t_max=100, local_search_iterations=100, epsilon=0.5)
def term(*x): raise RuntimeError
term._pdp_standard_termination = True
def step():
    torch.manual_seed(1)
    init = model.get_init_state(gm, bvm, bfm, ef, None, randomized=False, batch_replication=1)
    (pred, _), _ = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm, edge_feature=ef,
                         meta_data=None, is_training=False, iteration_num=20, check_termination=term, batch_replication=1)
    solved, _ = model.last_problem._ctx.cnf_eval(pred)
    torch.cuda.synchronize()
for i in range(3): step()
for i in range(3):
    t0 = time.perf_counter(); step(); print("step %.1f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable()
for i in range(5): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
