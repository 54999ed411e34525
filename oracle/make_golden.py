"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, through oracle/compat.py shims) on CPU in the authoring container.

    python oracle/make_golden.py            # rewrites every fixture

The reference is Python and cannot travel to the GPU box, so its outputs are committed as small
fixtures; `tests/` checks both the C oracle and the CUDA path against them.  Every array is produced
by reference code; this script only builds inputs, installs recording hooks and saves.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import compat  # noqa: E402
from pdp_solver_b200 import cnfgen  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DEV = torch.device("cpu")


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def tensors(batch):
    gm, bvm, bfm, ef = batch
    return T(gm), T(bvm), T(bfm), T(ef)


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))


# ----------------------------------------------------------------------------------------------
# batches
# ----------------------------------------------------------------------------------------------
def ragged_batch(seed):
    """Ragged clause lengths (1..5), repeated variables inside a clause, a degree-0 variable, an
    empty-clause-free problem without clauses, unit clauses."""
    rng = np.random.Generator(np.random.PCG64(seed))
    probs = []
    for n, m in [(12, 30), (7, 0), (9, 25), (15, 50)]:
        clauses = []
        for _ in range(m):
            k = int(rng.integers(1, 6))
            vs = rng.integers(1, n, size=k)            # variable n never appears -> degree 0
            sg = rng.integers(0, 2, size=k) * 2 - 1
            clauses.append([int(v * s) for v, s in zip(vs, sg)])
        probs.append((n, clauses))
    return cnfgen.from_clauses(probs)


def crafted_simplify_batch():
    """Unit chains, a one-variable conflict, a two-variable conflict (solver.py:257 `== 1` quirk),
    pure literals, degree-0 variables, a duplicate literal, a tautology."""
    probs = [
        (5, [[1], [-1, 2], [-2, 3], [-3, 4, 5], [4, -5]]),                 # unit chain
        (3, [[1], [-1], [2, 3]]),                                           # single conflict -> wipe
        (4, [[1], [-1], [2], [-2], [3, 4]]),                                # two conflicts -> quirk
        (6, [[1, 2], [1, 3], [-2, -3], [4, -5], [5, -4], [6, 6], [2, -2]]),  # pure literal 1, tautology
        (4, [[1, 2, 3], [-1, -2, -3], [1, -2, 3], [-1, 2, -3]]),            # nothing to do, var 4 degree 0
        (3, [[1, 2], [-1], [-2], [3]]),                                     # UP empties a clause
    ]
    return cnfgen.from_clauses(probs)


# ----------------------------------------------------------------------------------------------
# operator-level fixtures
# ----------------------------------------------------------------------------------------------
def gen_ops(name, batch, seed, adversarial):
    solver, prop, dec, pred, util, trainer = compat.load_reference()
    gm, bvm, bfm, ef = tensors(batch)
    E, V, F = gm.shape[1], bvm.shape[0], bfm.shape[0]
    B = int(bvm.max()) + 1
    g = torch.Generator().manual_seed(seed)
    sp = solver.SATProblem((gm, bvm, bfm, ef, None, None), DEV, 1)

    dq = torch.rand(E, 3, generator=g)
    df = torch.rand(E, 2, generator=g)
    df[:, 1] = 0
    pq = torch.rand(E, 3, generator=g)
    pf = torch.rand(E, 2, generator=g)
    pf[:, 1] = 0
    if adversarial:
        idx = torch.randperm(E, generator=g)
        n4 = max(E // 8, 1)
        df[idx[:n4], 0] = 1.0            # eta = 1  -> log(0) clamp at 1e-40
        df[idx[n4:2 * n4], 0] = 0.0
        dq[idx[2 * n4:3 * n4], 0] = 0.0  # q_u = 0  -> clamp
        dq[idx[3 * n4:4 * n4], 0] = 1e-42  # subnormal below eps
    av = (torch.rand(V, 1, generator=g) > 0.25).float()
    af = (torch.rand(F, 1, generator=g) > 0.25).float()
    em = torch.mm(sp._graph_mask_tuple[1], av) * torch.mm(sp._graph_mask_tuple[3], af)
    active = (torch.rand(B, 1, generator=g) > 0.3).to(torch.uint8)

    P = prop.SurveyPropagator(DEV, decimator_dimension=1, include_adaptors=False)
    out = {}
    with torch.no_grad():
        # (1) no edge mask, all active
        q1, f1 = P((pq, pf), (dq, df), sp, False, None)
        # (2) edge mask + frozen problems
        q2, f2 = P((pq, pf), (dq, df, em), sp, False, active)
        S = pred.SurveyScorer(DEV, message_dimension=1, include_adaptors=False)
        sp._active_functions = af.clone()
        score, _ = S((dq, df), sp)
        sp._active_functions = torch.ones(F, 1)
        score_all, _ = S((dq, df), sp)

        ev = util.SatCNFEvaluator(DEV)
        vp = torch.rand(V, 1, generator=g)
        vp[torch.rand(V, generator=g) < 0.2, 0] = 0.5
        vp[torch.rand(V, generator=g) < 0.2, 0] = 1.0
        vp[torch.rand(V, generator=g) < 0.2, 0] = 0.0
        solved, nun = ev(vp, gm, bvm, bfm, ef, None)

        base = solver.PropagatorDecimatorSolverBase(DEV, "x", None, None, None)
        asg = (torch.randint(0, 2, (V, 1), generator=g).float() * 2 - 1)
        sp._active_variables = av.clone()
        sp._active_functions = af.clone()
        sp._edge_mask = em
        energy, unsat_fn = base._compute_energy(asg.clone(), sp)
        delta = base._compute_energy_diff(asg.clone(), sp)

        sm = util.sparse_smooth_max(df[:, 0].unsqueeze(1), sp._graph_mask_tuple[0], DEV)
        smax = util.sparse_max((sm * av).squeeze(1), sp._batch_mask_tuple[0], DEV)
        amax = util.sparse_argmax((score.abs() * av).squeeze(1), sp._batch_mask_tuple[0], DEV)

    save(name, graph_map=gm.numpy(), bvm=bvm.numpy(), bfm=bfm.numpy(), ef=ef.numpy(),
         dq=dq.numpy(), df=df.numpy(), pq=pq.numpy(), pf=pf.numpy(), av=av.numpy(), af=af.numpy(),
         em=em.numpy(), active=active.numpy(),
         sp1_q=q1.numpy(), sp1_f=f1.numpy(), sp2_q=q2.numpy(), sp2_f=f2.numpy(),
         score=score.numpy(), score_all=score_all.numpy(),
         vp=vp.numpy(), solved=solved.numpy(), n_unsat=nun.numpy(),
         asg=asg.numpy(), energy=energy.numpy(), unsat_fn=unsat_fn.numpy(), delta=delta.numpy(),
         smooth_max=sm.numpy(), sparse_max=smax.numpy(), sparse_argmax=amax.numpy())


def gen_simplify(name, batch):
    solver, *_ = compat.load_reference()
    gm, bvm, bfm, ef = tensors(batch)
    sp = solver.SATProblem((gm, bvm, bfm, ef, None, None), DEV, 1)
    sp.simplify()
    out = dict(graph_map=gm.numpy(), bvm=bvm.numpy(), bfm=bfm.numpy(), ef=ef.numpy(),
               av=sp._active_variables.numpy().copy(), af=sp._active_functions.numpy().copy(),
               sol=sp._solution.numpy().copy(), is_sat=sp._is_sat.numpy().copy())
    # then fix two variables and simplify again (set_variables, solver.py:275-279)
    V = bvm.shape[0]
    g = torch.Generator().manual_seed(5)
    asg = torch.zeros(V, 1)
    pick = torch.randperm(V, generator=g)[: max(V // 6, 1)]
    asg[pick, 0] = (torch.randint(0, 2, (pick.shape[0],), generator=g).float() * 2 - 1)
    out["asg"] = asg.numpy().copy()
    sp.set_variables(asg)
    out.update(av2=sp._active_variables.numpy().copy(), af2=sp._active_functions.numpy().copy(),
               sol2=sp._solution.numpy().copy(), is_sat2=sp._is_sat.numpy().copy())
    save(name, **out)


# ----------------------------------------------------------------------------------------------
# trajectory fixtures: the whole forward() of SurveyPropagatorSolver with recording hooks
# ----------------------------------------------------------------------------------------------
def run_reference_forward(batch, T_iters, W, epsilon, randomized, seed, tol=0.02, t_max=100, b=1,
                          model_type="p-d-p", record_iters=True):
    solver, prop, dec, pred, util, trainer = compat.load_reference()
    gm, bvm, bfm, ef = tensors(batch)
    torch.manual_seed(seed)
    if model_type == "p-d-p":
        model = solver.SurveyPropagatorSolver(DEV, "sp", tolerance=tol, t_max=t_max,
                                              local_search_iterations=W, epsilon=epsilon)
    else:
        model = solver.WalkSATSolver(DEV, "ws", iteration_num=W, epsilon=epsilon)
    cb0 = compat.make_termination_callback(DEV)
    init = model.get_init_state(gm, bvm, bfm, ef, None, randomized=randomized, batch_replication=b)

    rec = dict(eta=[], qu=[], av=[], af=[], sol=[], active=[], counters=[], events=[], draws=[])
    state = dict(it=0, problem=None)

    def cb(active, prediction, sat_problem):
        cb0(active, prediction, sat_problem)
        if record_iters:
            rec["av"].append(sat_problem._active_variables[:, 0].numpy().copy())
            rec["af"].append(sat_problem._active_functions[:, 0].numpy().copy())
            rec["sol"].append(sat_problem._solution.numpy().copy())
            rec["active"].append(active[:, 0].numpy().copy())
            rec["counters"].append(model._decimator._counters[:, 0].numpy().copy())

    if model._decimator is not None:
        orig_dec = model._decimator.forward

        def dec_hook(init_state, message_state, sat_problem, is_training, active_mask=None):
            state["it"] += 1
            state["problem"] = sat_problem
            out = orig_dec(init_state, message_state, sat_problem, is_training, active_mask)
            if record_iters:
                rec["eta"].append(message_state[1][:, 0].numpy().copy())
                rec["qu"].append(message_state[0][:, 0].numpy().copy())
            return out

        model._decimator.forward = dec_hook

    orig_set = solver.SATProblem.set_variables

    def set_hook(self, assignment):
        nz = torch.nonzero(assignment[:, 0]).flatten()
        for i in nz.tolist():
            rec["events"].append((state["it"], i, int(assignment[i, 0].item())))
        return orig_set(self, assignment)

    solver.SATProblem.set_variables = set_hook
    orig_rand = torch.rand

    def rand_hook(*a, **k):
        r = orig_rand(*a, **k)
        rec["draws"].append(r.numpy().copy())
        return r

    try:
        with torch.no_grad():
            init_np = [[x.numpy().copy() for x in st] for st in init] if init[0] is not None else None
            torch.rand = rand_hook
            (p, _), states = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm,
                                   edge_feature=ef, meta_data=None, is_training=False, iteration_num=T_iters,
                                   check_termination=cb, batch_replication=b)
    finally:
        torch.rand = orig_rand
        solver.SATProblem.set_variables = orig_set
    ev = util.SatCNFEvaluator(DEV)
    solved, nun = ev(p, gm, bvm, bfm, ef, None)
    return dict(init=init_np, pred=p.numpy().copy(), rec=rec, solved=solved.numpy(), n_unsat=nun.numpy(),
                final_states=states)


def gen_traj(name, batch, T_iters, W, epsilon, randomized, seed, **kw):
    r = run_reference_forward(batch, T_iters, W, epsilon, randomized, seed, **kw)
    gm, bvm, bfm, ef = batch
    rec = r["rec"]
    V = bvm.shape[0]
    B = int(bvm.max()) + 1
    draws = rec["draws"]
    # draw order: rand(n_active) once [1-D, only if n_active > 0], then per
    # WalkSAT iteration rand([V,1]) [2-D] followed by rand(B) [1-D]
    fill = np.zeros(0, np.float32)
    rvl, rcl = [], []
    for j, d in enumerate(draws):
        if d.ndim == 2:
            rvl.append(d.reshape(-1))
        elif j > 0 and draws[j - 1].ndim == 2:
            rcl.append(d.reshape(-1))
        else:
            assert j == 0
            fill = d.reshape(-1)
    assert len(rvl) == len(rcl)
    rv = np.stack(rvl) if rvl else np.zeros((0, V), np.float32)
    rc = np.stack(rcl) if rcl else np.zeros((0, B), np.float32)
    ev = np.array(rec["events"], dtype=np.int64).reshape(-1, 3)
    arrays = dict(graph_map=gm, bvm=bvm, bfm=bfm, ef=ef, T=T_iters, W=W, epsilon=epsilon,
                  tol=kw.get("tol", 0.02), t_max=kw.get("t_max", 100),
                  pred=r["pred"], solved=r["solved"], n_unsat=r["n_unsat"], events=ev,
                  fill=fill, rand_var=rv, rand_coin=rc)
    if r["init"] is not None:
        arrays.update(init_pq=r["init"][0][0], init_pf=r["init"][0][1], init_dq=r["init"][1][0], init_df=r["init"][1][1])
    if rec["eta"]:
        arrays.update(eta=np.stack(rec["eta"]).astype(np.float32), qu=np.stack(rec["qu"]).astype(np.float32),
                      av=np.stack(rec["av"]).astype(np.uint8), af=np.stack(rec["af"]).astype(np.uint8),
                      sol=np.stack(rec["sol"]).astype(np.float32), active=np.stack(rec["active"]).astype(np.uint8),
                      counters=np.stack(rec["counters"]).astype(np.float32))
        fs = r["final_states"]
        arrays.update(final_q3=fs[0][0].numpy(), final_fs2=fs[0][1].numpy())
    save(name, **arrays)


# ----------------------------------------------------------------------------------------------
# neural model types (p-nd-np, np-nd-np): weights + injected initial states + what the reference returns
# ----------------------------------------------------------------------------------------------
def gen_neural(name, model_type, batch, T_iters, seed, dims, only=None):
    """dims = (hidden, mem_hidden, agg_hidden, mem_agg_hidden, classifier).  The reference model is built under
    torch.manual_seed(seed); its state_dict is stored so that the B200 modules can load it (same keys)."""
    solver, prop, dec, pred, util, trainer = compat.load_reference()
    gm, bvm, bfm, ef = tensors(batch)
    H, MH, AH, MAH, CH = dims
    torch.manual_seed(seed)
    clf = trainer.Perceptron(H, CH, 1)
    if model_type == "p-nd-np":
        model = solver.NeuralSurveyPropagatorSolver(DEV, "m", edge_dimension=1, meta_data_dimension=0, decimator_dimension=H,
                                                    mem_hidden_dimension=MH, agg_hidden_dimension=AH, mem_agg_hidden_dimension=MAH,
                                                    prediction_dimension=1, variable_classifier=clf, function_classifier=None,
                                                    dropout=0, local_search_iterations=0, epsilon=0.5)
    else:
        model = solver.NeuralPropagatorDecimatorSolver(DEV, "m", edge_dimension=1, meta_data_dimension=0, propagator_dimension=H,
                                                       decimator_dimension=H, mem_hidden_dimension=MH, agg_hidden_dimension=AH,
                                                       mem_agg_hidden_dimension=MAH, prediction_dimension=1, variable_classifier=clf,
                                                       function_classifier=None, dropout=0, local_search_iterations=0, epsilon=0.5)
    model.eval()
    cb = compat.make_termination_callback(DEV)
    preds = []
    orig_pred = model._predictor.forward

    def hook(decimator_state, sat_problem, last_call=False):
        out = orig_pred(decimator_state, sat_problem, last_call)
        preds.append(out[0].detach().clone().numpy().reshape(-1))
        return out

    model._predictor.forward = hook
    with torch.no_grad():
        init = model.get_init_state(gm, bvm, bfm, ef, None, randomized=True, batch_replication=1)
        init_np = [[t.clone().numpy() for t in st] for st in init]
        (vp, fp), (ps, ds) = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm,
                                   edge_feature=ef, meta_data=None, is_training=False, iteration_num=T_iters,
                                   check_termination=cb, simplify=True, batch_replication=1)
    model._predictor.forward = orig_pred
    arrays = dict(graph_map=batch[0], bvm=batch[1], bfm=batch[2], ef=batch[3], T=T_iters, dims=np.array(dims),
                  model_type=np.array(model_type), pred=vp.numpy().reshape(-1), preds=np.stack(preds).astype(np.float32),
                  init_p0=init_np[0][0], init_p1=init_np[0][1], init_d0=init_np[1][0], init_d1=init_np[1][1],
                  final_p0=ps[0].numpy(), final_p1=ps[1].numpy(), final_d0=ds[0].numpy(), final_d1=ds[1].numpy())
    # the reference registers every sub-module twice (attribute + _module_list): store each tensor once
    seen, alias = {}, []
    for k, v in model.state_dict().items():
        key = (v.data_ptr(), tuple(v.shape))
        if key in seen:
            alias.append("%s=%s" % (k, seen[key]))
        else:
            seen[key] = k
            arrays["w:" + k] = v.numpy()
    arrays["w_alias"] = np.array(";".join(alias))
    save(name, **arrays)


# ----------------------------------------------------------------------------------------------
# batch replication (-b): replicated forward, termination across replicas, WalkSAT, _deduplicate
# ----------------------------------------------------------------------------------------------
def gen_replicated(name, batch, T_iters, W, epsilon, seed, b, **kw):
    r = run_reference_forward(batch, T_iters, W, epsilon, False, seed, b=b, record_iters=False, **kw)
    gm, bvm, bfm, ef = batch
    V, B = bvm.shape[0] * b, (int(bvm.max()) + 1) * b
    draws = r["rec"]["draws"]
    fill = np.zeros(0, np.float32)
    rvl, rcl = [], []
    for j, d in enumerate(draws):
        if d.ndim == 2:
            rvl.append(d.reshape(-1))
        elif j > 0 and draws[j - 1].ndim == 2:
            rcl.append(d.reshape(-1))
        else:
            assert j == 0
            fill = d.reshape(-1)
    rv = np.stack(rvl) if rvl else np.zeros((0, V), np.float32)
    rc = np.stack(rcl) if rcl else np.zeros((0, B), np.float32)
    ev = np.array(r["rec"]["events"], dtype=np.int64).reshape(-1, 3)
    save(name, graph_map=gm, bvm=bvm, bfm=bfm, ef=ef, T=T_iters, W=W, epsilon=epsilon, b=b, tol=kw.get("tol", 0.02),
         t_max=kw.get("t_max", 100), pred=r["pred"], solved=r["solved"], n_unsat=r["n_unsat"], events=ev, fill=fill,
         rand_var=rv, rand_coin=rc, init_dq=r["init"][1][0], init_df=r["init"][1][1])

# ----------------------------------------------------------------------------------------------
# command-line / input-pipeline fixtures (reference dataset.py, dimacs2json.py, trainer.predict)
# ----------------------------------------------------------------------------------------------
def _cli_rows(seed):
    """a small mixed dataset in the compact row format, python ints only"""
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = []
    specs = [(20, 3, 3.4), (35, 3, 3.8), (50, 3, 4.1), (28, 4, 7.5), (60, 3, 3.5), (24, 3, 4.3), (45, 5, 15.0),
             (32, 3, 3.9), (55, 3, 3.7), (18, 3, 2.5), (40, 3, 4.0), (26, 4, 8.5), (48, 3, 3.6), (30, 3, 4.2)]
    for j, (n, k, alpha) in enumerate(specs):
        m = int(n * alpha)
        lits, cls = [], []
        for c in range(m):
            vs = rng.choice(n, size=k, replace=False)
            sg = rng.integers(0, 2, size=k) * 2 - 1
            for v, sgn in zip(vs, sg):
                lits.append(int((v + 1) * sgn))
                cls.append(c + 1)
        rows.append([[n, m], lits, cls, float(j % 2), ["prob_%02d.cnf" % j]])
    return rows


_DIMACS = {
    "alpha_1.cnf": "c a comment line\nc another\np cnf 9 7\n1 -2 3 0\n-1 4 0\n5 -5 6 0\n2 2 -7 0\n-3 -4 -6 0\n9 0\n1 9 -2 0\n",
    "beta_0.dimacs": "p cnf 12 6\n-12 3 5 0\n3 -5 0\n10 11 -12 0\n-3 0\n5 10 0\n-11 -10 -3 0\n%\n0\n",
    "gamma.cnf": "c header with unused variables and a short clause count\np cnf 20 9\n4 -8 15 0\n-4 8 0\n15 -16 20 0\n8 16 -20 0\n-15 -4 0\n",
}


def gen_cli():
    import io
    import json
    import yaml
    _, _, _, _, _, trainer = compat.load_reference()
    from pdp.factorgraph.dataset import FactorGraphDataset
    import dimacs2json as ref_d2j

    data_path = os.path.join(OUT, "cli_small.json")
    rows = _cli_rows(71)
    with open(data_path, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")

    # --- collate: the reference's dataset + divider, one segment and a forced split ---------------
    arrays = {}
    for tag, limit in (("one", 40000000), ("split", 25000)):
        ds = FactorGraphDataset(input_file=data_path, limit=limit, hidden_dim=3)
        out = ds.dag_collate_fn([ds[i] for i in range(len(ds))])
        gm, bvm, bfm, ef, gf, lab, misc = out
        arrays[tag + "_segments"] = len(gm)
        arrays[tag + "_limit"] = limit
        for s in range(len(gm)):
            arrays["%s_%d_gm" % (tag, s)] = gm[s].numpy()
            arrays["%s_%d_bvm" % (tag, s)] = bvm[s].numpy()
            arrays["%s_%d_bfm" % (tag, s)] = bfm[s].numpy()
            arrays["%s_%d_ef" % (tag, s)] = ef[s].numpy()
            arrays["%s_%d_label" % (tag, s)] = lab[s].numpy()
            arrays["%s_%d_ids" % (tag, s)] = np.array([m[0] for m in misc[s]])
    save("cli_collate", **arrays)

    # --- DIMACS: the reference's CompactDimacs on three small files --------------------------------
    ddir = os.path.join(OUT, "dimacs")
    os.makedirs(ddir, exist_ok=True)
    expected = {}
    for name, text in _DIMACS.items():
        path = os.path.join(ddir, name)
        with open(path, "w") as f:
            f.write(text)
        stem = os.path.splitext(path)[0]
        label = float(stem[-1]) if stem[-1].isdigit() else -1          # dimacs2json.py:111
        js = ref_d2j.CompactDimacs(path, label, False).to_json()
        expected[name] = [[int(js[0][0]), int(js[0][1])], [int(x) for x in js[1]], [int(x) for x in js[2]], js[3], js[4]]
    with open(os.path.join(OUT, "dimacs_expected.json"), "w") as f:
        json.dump(expected, f)

    # --- whole predict runs of the reference's trainer, two seeds (appendix A, steps 4-6) ----------
    lines = {}
    T_iters = 1000
    for seed in (1, 2):
        config = {"model_type": "p-d-p", "model_name": "sp", "tolerance": 0.02, "t_max": 100,
                  "test_path": data_path, "test_recurrence_num": T_iters, "batch_replication": 1, "batch_size": 5000,
                  "max_cache_size": 100000, "test_batch_limit": 25000, "local_search_iteration": 0, "epsilon": 0.5,
                  "verbose": False, "cpu_mode": True, "random_seed": seed, "model_path": None, "hidden_dim": 3,
                  "dropout": 0, "error_dim": 1, "exploration": 0}
        np.random.seed(seed)
        torch.manual_seed(seed)
        tr = trainer.SatFactorGraphTrainer(config=config, use_cuda=False, logger=compat._NullLogger())
        tr._num_cores = 0
        buf = io.StringIO()
        tr.predict(test_list=data_path, out_file=buf, import_path_base=None,
                   post_processor=tr._post_process_predictions, batch_replication=1)
        lines["seed%d" % seed] = buf.getvalue()
    with open(os.path.join(OUT, "cli_expected.json"), "w") as f:
        json.dump({"iterations": T_iters, "test_batch_limit": 25000, "outputs": lines}, f)
    print("wrote cli fixtures")

# ----------------------------------------------------------------------------------------------
# the remaining model types: reinforce (pi terms, coin-gated external forces) and np-d-np (neural
# propagator + sequential decimator with a neural scorer)
# ----------------------------------------------------------------------------------------------
def gen_reinforce(name, batch, T_iters, seed, pi, p_dec, randomized):
    solver, prop, dec, pred, util, trainer = compat.load_reference()
    gm, bvm, bfm, ef = tensors(batch)
    torch.manual_seed(seed)
    model = solver.ReinforceSurveyPropagatorSolver(DEV, "r", pi=pi, decimation_probability=p_dec,
                                                   local_search_iterations=0, epsilon=0.5)
    cb0 = compat.make_termination_callback(DEV)
    preds, actives, coins = [], [], []

    def cb(active, prediction, sat_problem):
        preds.append(prediction[0].numpy().reshape(-1).copy())
        cb0(active, prediction, sat_problem)
        actives.append(active[:, 0].numpy().copy())

    orig_rand = torch.rand

    def rand_hook(*a, **k):
        r = orig_rand(*a, **k)
        assert r.numel() == 1
        coins.append(float(r.reshape(-1)[0]))
        return r

    with torch.no_grad():
        init = model.get_init_state(gm, bvm, bfm, ef, None, randomized=randomized, batch_replication=1)
        init_np = [[t.clone().numpy() for t in st] for st in init]
        try:
            torch.rand = rand_hook
            (vp, _), (ps, ds) = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm,
                                      edge_feature=ef, meta_data=None, is_training=False, iteration_num=T_iters,
                                      check_termination=cb, simplify=True, batch_replication=1)
        finally:
            torch.rand = orig_rand
    save(name, graph_map=batch[0], bvm=batch[1], bfm=batch[2], ef=batch[3], T=T_iters, pi=pi, p_dec=p_dec,
         coins=np.array(coins, np.float32), preds=np.stack(preds).astype(np.float32), actives=np.stack(actives),
         pred=vp.numpy().reshape(-1), init_p0=init_np[0][0], init_p1=init_np[0][1], init_d0=init_np[1][0],
         init_d1=init_np[1][1], final_q=ps[0].numpy(), final_f=ps[1].numpy())


def gen_npdnp(name, batch, T_iters, seed, dims, tol, t_max, with_termination):
    solver, prop, dec, pred, util, trainer = compat.load_reference()
    gm, bvm, bfm, ef = tensors(batch)
    H, MH, AH, MAH, CH = dims
    torch.manual_seed(seed)
    model = solver.NeuralSequentialDecimatorSolver(DEV, "m", edge_dimension=1, meta_data_dimension=0, propagator_dimension=H,
                                                   decimator_dimension=H, mem_hidden_dimension=MH, agg_hidden_dimension=AH,
                                                   mem_agg_hidden_dimension=MAH, classifier_dimension=CH, dropout=0,
                                                   tolerance=tol, t_max=t_max, local_search_iterations=0, epsilon=0.5)
    model.eval()
    cb0 = compat.make_termination_callback(DEV)
    events, state, actives, draws = [], dict(it=0), [], []
    orig_dec = model._decimator.forward

    def dec_hook(*a, **k):
        state["it"] += 1
        state["problem"] = a[2]
        return orig_dec(*a, **k)

    model._decimator.forward = dec_hook
    orig_set = solver.SATProblem.set_variables

    def set_hook(self, assignment):
        for i in torch.nonzero(assignment[:, 0]).flatten().tolist():
            events.append((state["it"], i, int(assignment[i, 0].item())))
        return orig_set(self, assignment)

    def cb(active, prediction, sat_problem):
        cb0(active, prediction, sat_problem)
        actives.append(active[:, 0].numpy().copy())
        state["problem"] = sat_problem

    orig_rand = torch.rand

    def rand_hook(*a, **k):
        r = orig_rand(*a, **k)
        draws.append(r.numpy().copy().reshape(-1))
        return r

    with torch.no_grad():
        init = model.get_init_state(gm, bvm, bfm, ef, None, randomized=True, batch_replication=1)
        init_np = [[t.clone().numpy() for t in st] for st in init]
        try:
            solver.SATProblem.set_variables = set_hook
            torch.rand = rand_hook
            (vp, _), (ps, ds) = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm,
                                      edge_feature=ef, meta_data=None, is_training=False, iteration_num=T_iters,
                                      check_termination=cb if with_termination else None, simplify=True,
                                      batch_replication=1)
        finally:
            torch.rand = orig_rand
            solver.SATProblem.set_variables = orig_set
    sp = state["problem"]
    arrays = dict(graph_map=batch[0], bvm=batch[1], bfm=batch[2], ef=batch[3], T=T_iters, dims=np.array(dims), tol=tol,
                  t_max=t_max, events=np.array(events, dtype=np.int64).reshape(-1, 3), with_termination=with_termination,
                  actives=np.stack(actives) if actives else np.zeros((0, 0)),
                  iterations=state["it"], pred=vp.numpy().reshape(-1), av=sp._active_variables[:, 0].numpy(),
                  af=sp._active_functions[:, 0].numpy(), fill=np.concatenate(draws) if draws else np.zeros(0, np.float32),
                  init_p0=init_np[0][0], init_p1=init_np[0][1], init_d0=init_np[1][0], init_d1=init_np[1][1],
                  final_p0=ps[0].numpy(), final_p1=ps[1].numpy())
    seen, alias = {}, []
    for k, v in model.state_dict().items():
        key = (v.data_ptr(), tuple(v.shape))
        if key in seen:
            alias.append("%s=%s" % (k, seen[key]))
        else:
            seen[key] = k
            arrays["w:" + k] = v.numpy()
    arrays["w_alias"] = np.array(";".join(alias))
    save(name, **arrays)
    print("  %d decimation events in %d iterations" % (len(events), state["it"]))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "types":
        gen_reinforce("reinforce_a", cnfgen.random_batch(1, 40, 3, 3.8, 81), 60, 11, 0.01, 0.5, False)
        gen_reinforce("reinforce_b", cnfgen.random_batch(5, 30, 3, 3.5, 82), 80, 12, 0.1, 0.5, False)
        gen_reinforce("reinforce_c", cnfgen.mixed_batch([(30, 3, 4.0), (20, 5, 15.0), (24, 4, 8.0)], 83), 50, 13, 0.05, 0.7, True)
        gen_npdnp("npdnp_a", cnfgen.random_batch(3, 18, 3, 3.6, 91), 24, 21, (24, 16, 16, 8, 8), 0.02, 3, False)
        gen_npdnp("npdnp_b", cnfgen.mixed_batch([(16, 3, 3.0), (12, 4, 7.0)], 92), 20, 22, (16, 12, 10, 6, 8), 0.5, 4, False)
        gen_npdnp("npdnp_c", cnfgen.random_batch(4, 14, 3, 3.2, 93), 12, 23, (16, 12, 10, 6, 8), 0.1, 3, True)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "cli":
        gen_cli()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "replicated":
        gen_replicated("rep_b3_a", cnfgen.random_batch(5, 30, 3, 3.7, 61), 120, 25, 0.5, 9, 3, t_max=30)
        gen_replicated("rep_b2_b", cnfgen.mixed_batch([(30, 3, 3.9), (20, 5, 14.0), (25, 3, 3.0)], 62), 100, 30, 0.4, 10, 2, t_max=25)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "replicated8":   # BASELINE.json configs[4] in miniature: -b 8 on mixed 3-/5-SAT
        gen_replicated("rep_b8_mixed", cnfgen.mixed_batch([(40, 3, 4.2), (24, 5, 18.0), (60, 3, 4.0), (16, 5, 16.0), (30, 3, 4.2)], 63),
                       120, 40, 0.5, 11, 8, t_max=30)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "neural":
        gen_neural("neural_pndnp_a", "p-nd-np", cnfgen.random_batch(4, 20, 3, 3.8, 51), 12, 5, (24, 16, 16, 8, 8))
        gen_neural("neural_pndnp_b", "p-nd-np", ragged_batch(52), 8, 6, (20, 12, 12, 6, 6))
        gen_neural("neural_npndnp_a", "np-nd-np", cnfgen.random_batch(3, 16, 4, 6.5, 53), 10, 7, (24, 16, 16, 8, 8))
        gen_neural("neural_npndnp_b", "np-nd-np", cnfgen.mixed_batch([(20, 3, 2.0), (12, 5, 8.0), (16, 3, 4.0)], 54), 8, 8, (16, 12, 10, 6, 8))
        return
    gen_ops("ops_3sat", cnfgen.random_batch(6, 20, 3, 4.0, 11), seed=1, adversarial=False)
    gen_ops("ops_3sat_adv", cnfgen.random_batch(5, 24, 3, 4.2, 12), seed=2, adversarial=True)
    gen_ops("ops_ragged", ragged_batch(13), seed=3, adversarial=True)
    gen_ops("ops_mixed", cnfgen.mixed_batch([(30, 3, 4.2), (25, 5, 12.0), (10, 4, 6.0), (40, 3, 3.0)], 14), seed=4, adversarial=False)
    gen_simplify("simplify_crafted", crafted_simplify_batch())
    gen_simplify("simplify_ragged", ragged_batch(15))
    gen_simplify("simplify_3sat", cnfgen.random_batch(8, 30, 3, 2.5, 16))
    # whole-forward trajectories (batch coupling present -> compare with the oracle's strict mode)
    gen_traj("traj_det_a", cnfgen.random_batch(6, 30, 3, 3.6, 21), 150, 20, 0.5, False, 1)
    gen_traj("traj_rand_a", cnfgen.random_batch(6, 30, 3, 3.6, 22), 150, 20, 0.5, True, 2)
    gen_traj("traj_det_b", cnfgen.random_batch(4, 60, 3, 4.1, 23), 160, 30, 0.5, False, 3, tol=0.02, t_max=40)
    gen_traj("traj_mixed", cnfgen.mixed_batch([(40, 3, 4.0), (25, 5, 16.0), (30, 4, 8.0)], 24), 200, 25, 0.3, False, 4, t_max=30)
    # single-problem trajectories: the semantics every problem must have inside any batch
    for j in range(4):
        gen_traj("traj_single_%d" % j, cnfgen.random_batch(1, 50, 3, 4.0, 30 + j), 200, 30, 0.5, j % 2 == 1, 10 + j, t_max=50)
    # WalkSAT only
    gen_traj("walksat_a", cnfgen.random_batch(8, 40, 3, 4.0, 41), 0, 60, 0.5, False, 7, model_type="walk-sat")
    gen_traj("walksat_b", ragged_batch(42), 0, 40, 0.2, False, 8, model_type="walk-sat")


if __name__ == "__main__":
    main()
