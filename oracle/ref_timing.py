"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never imported by the product path.

Times the UNMODIFIED reference (microsoft/PDP-Solver, loaded through oracle/compat.py) on its own stock code
path: `model.get_init_state(...)` + `model(init_state=..., is_training=False, iteration_num=T,
check_termination=trainer._check_recurrence_termination, ...)`, exactly the call of
`FactorGraphTrainerBase._predict_batch` (reference src/pdp/factorgraph/base.py:280-305), with
`use_cuda=False` (== `satyr.py --cpu_mode`, base.py:32,40) and `torch.set_num_threads(cpu_count)` as the
reference sets it (base.py:43,50).  Used by `bench.py --impl reference` and bench.py's `cpu_baseline` leg.

The forward is split with wall-clock probes around the reference's own methods (attribute wrappers on the live
object, nothing is re-implemented): `_forward_core` = the propagate / decimate / predict loop,
`_local_search` = WalkSAT, the rest = set-up (SATProblem construction, simplify, init state, merge).
"""
import os
import time

import numpy as np
import torch

from . import compat


def available():
    return compat.reference_available()


def pin_threads():
    """the reference's own thread setting (base.py:43,50), immune to OMP_NUM_THREADS=1 exported by torchrun"""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def build_model(model_type, device, W, epsilon, tol=0.02, t_max=100):
    solver = compat.load_reference()[0]
    if model_type == "p-d-p":
        return solver.SurveyPropagatorSolver(device, "sp", tolerance=tol, t_max=t_max, local_search_iterations=W, epsilon=epsilon)
    if model_type == "walk-sat":
        return solver.WalkSATSolver(device, "ws", iteration_num=W, epsilon=epsilon)
    if model_type in ("p-nd-np", "np-nd-np"):     # dims of config/Predict/PDP-np-nd-np-*.yaml, random-init weights
        trainer = compat.load_reference()[5]
        H, MH, AH, MAH, CH = 150, 100, 100, 50, 50
        torch.manual_seed(1)
        clf = trainer.Perceptron(H, CH, 1)          # reference trainer.py:20-29
        if model_type == "p-nd-np":
            m = solver.NeuralSurveyPropagatorSolver(device, "m", 1, 0, H, MH, AH, MAH, 1, variable_classifier=clf,
                                                    local_search_iterations=W, epsilon=epsilon)
        else:
            m = solver.NeuralPropagatorDecimatorSolver(device, "m", 1, 0, H, H, MH, AH, MAH, 1, variable_classifier=clf,
                                                       local_search_iterations=W, epsilon=epsilon)
        return m.to(device).eval()
    raise ValueError(model_type)


def timed_forward(model, batch, T, device, seed=1, batch_replication=1):
    """one forward of the reference on `device`.  batch = numpy (graph_map, bvm, bfm, edge_feature[E]).
    Returns dict(total_s, loop_s, walksat_s, setup_s, iterations, solved, edges)."""
    gm, bvm, bfm, ef = batch
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(device)
    gm_t, bvm_t, bfm_t = t(gm), t(bvm), t(bfm)
    ef_t = t(np.asarray(ef, dtype=np.float32).reshape(-1, 1))
    cb = compat.make_termination_callback(device)
    probes = {"loop": 0.0, "ws": 0.0, "iters": 0}
    sync = (lambda: torch.cuda.synchronize(device)) if device.type == "cuda" else (lambda: None)

    core, search, prop = model._forward_core, model._local_search, model._propagator

    def core_w(*a, **k):
        sync(); t0 = time.perf_counter()
        out = core(*a, **k)
        sync(); probes["loop"] += time.perf_counter() - t0
        return out

    def search_w(*a, **k):
        sync(); t0 = time.perf_counter()
        out = search(*a, **k)
        sync(); probes["ws"] += time.perf_counter() - t0
        return out

    if prop is not None:
        prop_fwd = prop.forward

        def prop_w(*a, **k):
            probes["iters"] += 1
            return prop_fwd(*a, **k)
        prop.forward = prop_w
    model._forward_core, model._local_search = core_w, search_w
    try:
        torch.manual_seed(seed)
        with torch.no_grad():
            sync(); t0 = time.perf_counter()
            init = model.get_init_state(gm_t, bvm_t, bfm_t, ef_t, None, randomized=False, batch_replication=batch_replication)
            (pred, _), _ = model(init_state=init, graph_map=gm_t, batch_variable_map=bvm_t, batch_function_map=bfm_t,
                                 edge_feature=ef_t, meta_data=None, is_training=False, iteration_num=T,
                                 check_termination=cb, batch_replication=batch_replication)
            sync(); total = time.perf_counter() - t0
            util = compat.load_reference()[4]
            solved, _ = util.SatCNFEvaluator(device)(pred, gm_t, bvm_t, bfm_t, ef_t, None)
    finally:
        model._forward_core, model._local_search = core, search
        if prop is not None:
            prop.forward = prop_fwd
    return {"total_s": total, "loop_s": probes["loop"], "walksat_s": probes["ws"],
            "setup_s": total - probes["loop"] - probes["ws"], "iterations": probes["iters"],
            "solved": int((solved > 0.5).sum().item()), "edges": int(gm.shape[1])}
