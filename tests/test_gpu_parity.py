"""GPU parity tests: the CUDA path (through the C ABI, pdp_solver_b200.engine.Context) against
 (1) golden vectors produced by the reference itself (tests/golden, oracle/make_golden.py) and
 (2) the C oracle on seeded inputs.
Integer / index / verdict results must be bit exact; surveys within 1e-4 absolute (north star)."""
import numpy as np
import pytest
import torch

from tests.helpers import golden, load, maxdiff, name

pytestmark = pytest.mark.gpu

SURVEY_TOL = 1e-4     # north star: converged surveys within 1e-4 absolute in fp32
STEP_TOL = 2e-6       # one operator application from identical inputs


def dev():
    return torch.device("cuda:0")


def T(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dev())
    return t if dtype is None else t.to(dtype)


def make_ctx(z):
    from pdp_solver_b200.engine import Context
    return Context(T(z["graph_map"]), T(z["bvm"]), T(z["bfm"]), T(z["ef"]))


def C(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("path", golden("ops_*.npz"), ids=name)
def test_operators_vs_reference(path):
    z = load(path)
    ctx = make_ctx(z)
    q1, f1 = ctx.sp_step(T(z["dq"]), T(z["df"]), None, T(z["pq"]), T(z["pf"]), None)
    assert maxdiff(C(q1), z["sp1_q"]) < STEP_TOL and maxdiff(C(f1), z["sp1_f"]) < STEP_TOL
    q2, f2 = ctx.sp_step(T(z["dq"]), T(z["df"]), T(z["em"]), T(z["pq"]), T(z["pf"]), T(z["active"]))
    assert maxdiff(C(q2), z["sp2_q"]) < STEP_TOL and maxdiff(C(f2), z["sp2_f"]) < STEP_TOL
    assert maxdiff(C(ctx.score(T(z["df"]), T(z["af"]))), z["score"]) < STEP_TOL
    assert maxdiff(C(ctx.score(T(z["df"]), torch.ones_like(T(z["af"])))), z["score_all"]) < STEP_TOL
    solved, nun = ctx.cnf_eval(T(z["vp"]))
    assert maxdiff(C(solved), z["solved"]) == 0 and maxdiff(C(nun), z["n_unsat"]) == 0
    en, uf = ctx.energy(T(z["asg"]), T(z["av"]), T(z["af"]))
    assert maxdiff(C(en), z["energy"]) == 0 and maxdiff(C(uf), z["unsat_fn"]) == 0
    de = ctx.energy_diff(T(z["asg"]), T(z["av"]), T(z["em"]))
    assert maxdiff(C(de), z["delta"]) == 0


def test_operators_unsorted_edges():
    """graph_map in arbitrary (not clause-major) edge order: results follow the caller's order."""
    z = dict(load(golden("ops_mixed.npz")[0]))
    E = z["graph_map"].shape[1]
    perm = np.random.default_rng(0).permutation(E)
    zz = dict(z)
    zz["graph_map"] = z["graph_map"][:, perm]
    zz["ef"] = z["ef"][perm]
    ctx = make_ctx(zz)
    # integer operators do not depend on accumulation order: exact
    solved, nun = ctx.cnf_eval(T(z["vp"]))
    assert maxdiff(C(solved), z["solved"]) == 0 and maxdiff(C(nun), z["n_unsat"]) == 0
    de = ctx.energy_diff(T(z["asg"]), T(z["av"]), T(z["em"][perm]))
    assert maxdiff(C(de), z["delta"]) == 0
    q2, f2 = ctx.sp_step(T(z["dq"][perm]), T(z["df"][perm]), T(z["em"][perm]), T(z["pq"][perm]), T(z["pf"][perm]), T(z["active"]))
    assert maxdiff(C(q2), z["sp2_q"][perm]) < 1e-5 and maxdiff(C(f2), z["sp2_f"][perm]) < 1e-5


@pytest.mark.parametrize("path", golden("simplify_*.npz"), ids=name)
def test_simplify_vs_reference(path):
    z = load(path)
    ctx = make_ctx(z)
    ctx.simplify()
    m = ctx.get_masks()
    for k in ("av", "af", "sol", "is_sat"):
        assert maxdiff(C(m[k]), z[k]) == 0, k
    ctx.set_variables(T(z["asg"]))
    m = ctx.get_masks()
    for k in ("av", "af", "sol", "is_sat"):
        assert maxdiff(C(m[k]), z[k + "2"]) == 0, k


def _replay(z, fused):
    """Runs the whole forward of the p-d-p / walk-sat model on the GPU from the golden's injected
    initial messages and random draws and compares with what the reference recorded."""
    ctx = make_ctx(z)
    Tn, W = int(z["T"]), int(z["W"])
    ctx.simplify()
    worst = 0.0
    if Tn > 0:
        ctx.enable_trace()
        ctx.load_state((T(z["init_pq"]), T(z["init_pf"])), (T(z["init_dq"]), T(z["init_df"])))
        n_ref = z["eta"].shape[0]
        if fused:
            done = ctx.sp_run(Tn, float(z["tol"]), int(z["t_max"]), True, sync=True)
            assert done == n_ref
            q, fs = ctx.store_state()
            m = ctx.get_masks()
            assert (C(m["av"]) == z["av"][-1]).all() and (C(m["af"]) == z["af"][-1]).all()
            assert (C(m["active"]) == z["active"][-1]).all()
            assert maxdiff(C(m["sol"]), z["sol"][-1]) == 0
            _, counters, _ = ctx.problem_flags()
            act = z["active"][-1].astype(bool)
            assert maxdiff(C(counters)[act], z["counters"][-1][act]) == 0
            worst = max(maxdiff(C(fs[:, 0]), z["final_fs2"][:, 0]), maxdiff(C(q[:, 0]), z["final_q3"][:, 0]))
        else:
            for t in range(n_ref):
                done = ctx.sp_run(1, float(z["tol"]), int(z["t_max"]), True, sync=True)
                assert done == 1
                q, fs = ctx.store_state()
                m = ctx.get_masks()
                worst = max(worst, maxdiff(C(fs[:, 0]), z["eta"][t]), maxdiff(C(q[:, 0]), z["qu"][t]))
                assert (C(m["av"]) == z["av"][t]).all() and (C(m["af"]) == z["af"][t]).all(), t
                assert (C(m["active"]) == z["active"][t]).all(), t
                assert maxdiff(C(m["sol"]), z["sol"][t]) == 0, t
                _, counters, _ = ctx.problem_flags()
                act = z["active"][t].astype(bool)
                # counters of frozen problems are unobservable in the reference and not maintained here
                assert maxdiff(C(counters)[act], z["counters"][t][act]) == 0, t
        tr = C(ctx.trace()).astype(np.int64)
        ref = z["events"]
        assert tr.shape == ref.shape
        key = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
        assert (key(tr) == key(ref)).all()
        ctx.disable_trace()
    n_act = ctx.count_active_variables()
    assert n_act == z["fill"].shape[0]
    if n_act:
        ctx.random_fill(T(z["fill"]))
    n_w = z["rand_var"].shape[0]
    if n_w:
        pred, it = ctx.walksat(W, float(z["epsilon"]), T(z["rand_var"]), T(z["rand_coin"]), sync=True)
        assert it == n_w or it == W
    else:
        pred, it = ctx.walksat(0, float(z["epsilon"]), None, None, sync=True)
    assert maxdiff(C(pred), z["pred"]) == 0
    solved, nun = ctx.cnf_eval(pred)
    assert maxdiff(C(solved), z["solved"]) == 0 and maxdiff(C(nun), z["n_unsat"]) == 0
    return worst


@pytest.mark.parametrize("fused", [False, True], ids=["stepwise", "fused"])
@pytest.mark.parametrize("path", golden("traj_*.npz") + golden("walksat_*.npz"), ids=name)
def test_forward_vs_reference(path, fused):
    """identical decimation sequence, masks, solutions, counters, WalkSAT flips, verdicts; surveys
    within 5e-2 at every one of up to 200 free-running iterations (5e-3 at the end) (the C oracle shows the same libm-level drift
    on the non-convergent instances) -- see test_single_step_vs_reference for the 1e-4 bound."""
    z = load(path)
    worst = _replay(z, fused)
    assert worst < (5e-2 if not fused else 5e-3)


@pytest.mark.parametrize("path", golden("traj_*.npz"), ids=name)
def test_single_step_vs_reference(path):
    """Each recorded iteration replayed from the reference's own state one step back: surveys within
    1e-4 (observed ~1e-6)."""
    z = load(path)
    ctx = make_ctx(z)
    E = z["graph_map"].shape[1]
    worst = 0.0
    n_ref = z["eta"].shape[0]
    gm = z["graph_map"]
    for t in range(1, n_ref, max(1, n_ref // 25)):
        av, af = z["av"][t - 1].astype(np.float32), z["af"][t - 1].astype(np.float32)
        em = av[gm[0]] * af[gm[1]]
        dq = np.zeros((E, 3), np.float32)
        dq[:, 0] = z["qu"][t - 1]
        df = np.zeros((E, 2), np.float32)
        df[:, 0] = z["eta"][t - 1]
        active = z["active"][t - 1]
        q, f = ctx.sp_step(T(dq), T(df), T(em), T(dq), T(df), T(active))
        worst = max(worst, maxdiff(C(f[:, 0]), z["eta"][t]), maxdiff(C(q[:, 0]), z["qu"][t]))
    assert worst < SURVEY_TOL


# ------------------------------------------------------------------------------------------------
# against the C oracle on seeded inputs
# ------------------------------------------------------------------------------------------------
def _oracle_forward(batch, init, Tn, tol, t_max, W, eps, rng):
    from oracle import pdp_oracle as po
    gm, bvm, bfm, ef = batch
    o = po.Oracle(gm, bvm, bfm, ef, strict=False)
    o.simplify()
    o.set_state(*init)
    done = o.run(Tn, tol, t_max, True)
    o.after_run = (o.masks(), o.state(), o.trace())
    n_act = o.count_active_variables()
    fill = rng.random(max(n_act, 1), dtype=np.float32)
    if n_act:
        o.random_fill(fill)
    rv = rng.random((W, o.V), dtype=np.float32)
    rc = rng.random((W, o.B), dtype=np.float32)
    pred, wit = o.local_search(W, eps, rv, rc)
    return o, done, fill, rv, rc, pred, wit


ORACLE_SPECS = [(64, 100, 3, 4.2, 150, 7), (8, 600, 3, 4.0, 120, 8), (3, 4000, 3, 3.9, 60, 9), (6, 300, 5, 17.0, 80, 10),
                (200, 50, 3, 4.3, 200, 11), (5, 1000, 4, 9.5, 100, 12)]


@pytest.mark.parametrize("strict_math", [True, False], ids=["strictmath", "productmath"])
@pytest.mark.parametrize("spec", ORACLE_SPECS, ids=lambda s: "B%d_n%d_k%d" % (s[0], s[1], s[2]))
def test_forward_vs_oracle(spec, strict_math):
    """Whole forward() against the C oracle on seeded random k-SAT.

    strictmath: the TEST build of the library (correctly rounded fp32 log/exp, as the oracle's math
    mode 1) must reproduce the oracle's trajectory exactly -- decimation sequence, masks, iteration
    counts, WalkSAT flips, verdicts -- on every instance, including numerically chaotic ones (forced
    decimations of non-converged surveys amplify 1-ulp libm differences to O(1) within tens of
    iterations, so only identical arithmetic can be compared there).
    productmath: the shipped build (logf/expf) on the well-conditioned instances, same exact checks,
    surveys within 1e-3 free-running."""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    Bn, n, k, alpha, Tn, seed = spec
    if not strict_math and seed not in (8, 9):
        pytest.skip("chaotic instance: compared in strict-math mode only")
    batch = cnfgen.random_batch(Bn, n, k, alpha, seed)
    E = batch[0].shape[1]
    rng = np.random.default_rng(seed)
    init = po.init_state(E, randomized=(seed % 2 == 0), rng=rng)
    W, eps, tol, t_max = 30, 0.5, 0.02, 25
    po.set_math_mode(strict_math)
    try:
        o, done, fill, rv, rc, pred, wit = _oracle_forward(batch, init, Tn, tol, t_max, W, eps, rng)
    finally:
        po.set_math_mode(False)

    ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]), strict_math=strict_math)
    ctx.enable_trace()
    ctx.simplify()
    ctx.load_state((T(init[0][0]), T(init[0][1])), (T(init[1][0]), T(init[1][1])))
    gdone = ctx.sp_run(Tn, tol, t_max, True, sync=True)
    tr = C(ctx.trace()).astype(np.int64)
    ctx.disable_trace()
    om, (oq, ofs), ref = o.after_run
    key = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
    assert gdone == done
    assert tr.shape == ref.shape and (key(tr) == key(ref)).all(), "decimation sequence differs"
    m = ctx.get_masks()
    assert (C(m["av"]) == om["av"]).all() and (C(m["af"]) == om["af"]).all()
    assert (C(m["active"]) == om["active"]).all()
    assert maxdiff(C(m["sol"]), om["sol"]) == 0
    q, fs = ctx.store_state()
    # frozen problems stop being updated on both sides at the same iteration; NaN positions (SP
    # contradictions, sticky through the reference's arithmetic blend) must coincide
    tol_msg = 1e-6 if strict_math else 1e-3
    assert maxdiff(C(fs[:, 0]), ofs[:, 0]) <= tol_msg and maxdiff(C(q[:, 0]), oq[:, 0]) <= tol_msg
    n_act = ctx.count_active_variables()
    assert n_act == int(om["av"].sum())
    if n_act:
        ctx.random_fill(T(fill))
    gpred, git = ctx.walksat(W, eps, T(rv), T(rc), sync=True)
    assert git == wit
    assert maxdiff(C(gpred), pred) == 0
    s1, u1 = ctx.cnf_eval(gpred)
    s2, u2 = o.cnf_eval(pred)
    assert maxdiff(C(s1), s2) == 0 and maxdiff(C(u1), u2) == 0


def test_fused_equals_stepwise_bitwise():
    """T iterations in one persistent launch == T launches of one iteration, bit for bit."""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    batch = cnfgen.random_batch(16, 200, 3, 4.1, 77)
    E = batch[0].shape[1]
    init = po.init_state(E, False)
    outs = []
    for fused in (True, False):
        ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]))
        ctx.simplify()
        ctx.load_state((T(init[0][0]), T(init[0][1])), (T(init[1][0]), T(init[1][1])))
        if fused:
            ctx.sp_run(90, 0.02, 20, True, sync=True)
        else:
            for _ in range(90):
                if ctx.sp_run(1, 0.02, 20, True, sync=True) == 0:
                    break
        q, fs = ctx.store_state()
        m = ctx.get_masks()
        outs.append((C(q[:, 0]), C(fs[:, 0]), C(m["av"]), C(m["af"]), C(m["sol"]), C(m["active"])))
    for a, b in zip(*outs):
        assert np.array_equal(a, b, equal_nan=True)


def test_batch_composition_invariance():
    """A problem's trajectory does not depend on its batch mates (what makes sharding across GPUs
    result-invariant): problem 3 alone == problem 3 inside a batch, bit for bit."""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    rng = np.random.Generator(np.random.PCG64(5))
    probs = [(150,) + cnfgen.random_ksat(150, 3, 4.0, rng) for _ in range(6)]
    full = cnfgen.collate(probs)
    solo = cnfgen.collate([probs[3]])
    res = []
    for batch in (full, solo):
        E = batch[0].shape[1]
        init = po.init_state(E, False)
        ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]))
        ctx.simplify()
        ctx.load_state((T(init[0][0]), T(init[0][1])), (T(init[1][0]), T(init[1][1])))
        ctx.sp_run(120, 0.02, 30, True, sync=True)
        q, fs = ctx.store_state()
        m = ctx.get_masks()
        res.append((batch, C(fs[:, 0]), C(m["av"]), C(m["sol"])))
    (fb, feta, fav, fsol), (sb, seta, sav, ssol) = res
    esel = fb[1][fb[0][0]] == 3
    vsel = fb[1] == 3
    assert np.array_equal(feta[esel], seta, equal_nan=True)
    assert np.array_equal(fav[vsel], sav) and np.array_equal(fsol[vsel], ssol)


def test_empty_and_degenerate_inputs():
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    # a problem without clauses, a problem whose only clause is empty-after-UP, unit clauses
    batch = cnfgen.from_clauses([(3, []), (2, [[1], [-1, 2]]), (1, [[1], [-1]])])
    ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]))
    ctx.simplify()
    m = ctx.get_masks()
    assert C(m["av"]).sum() == 0
    solved, nun = ctx.cnf_eval(m["sol"])
    assert C(solved).tolist() == [1.0, 1.0, 0.0]


# ------------------------------------------------------------------------------------------------
# blocked message layout (pdp_layout.cu) and the shared-memory passes built on it
# ------------------------------------------------------------------------------------------------
LAYOUT_SPECS = [(64, 100, 3, 4.2, 1), (3, 30000, 3, 4.2, 2), (5, 2000, 5, 18.0, 3), (400, 20, 3, 4.0, 4), (1, 200000, 3, 4.2, 5)]


@pytest.mark.parametrize("spec", LAYOUT_SPECS, ids=lambda s: "B%d_n%d_k%d" % (s[0], s[1], s[2]))
def test_blocked_layout_selfcheck(spec):
    """every invariant of the block tables, position maps, scatter tables and write-out pieces"""
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    Bn, n, k, alpha, seed = spec
    batch = cnfgen.random_batch(Bn, n, k, alpha, seed)
    ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]))
    errs, info = ctx.check_layout()
    E = batch[0].shape[1]
    assert info["blocked"] == 1 and info["nvb"] >= 1 and info["ncb"] >= 1
    assert errs[:6] == [0] * 6, (errs, info)
    assert errs[6] == E and errs[7] == E, (errs, info)


@pytest.mark.parametrize("path", golden("ops_*.npz") + golden("simplify_*.npz"), ids=name)
def test_blocked_layout_selfcheck_goldens(path):
    z = load(path)
    ctx = make_ctx(z)
    errs, info = ctx.check_layout()
    assert errs[:6] == [0] * 6, (errs, info)
    if info["blocked"]:
        E = z["graph_map"].shape[1]
        assert errs[6] == E and errs[7] == E


@pytest.mark.parametrize("spec", [(48, 100, 3, 4.2, 150, 21), (2, 20000, 3, 4.1, 60, 22), (6, 500, 5, 17.0, 80, 23)],
                         ids=lambda s: "B%d_n%d_k%d" % (s[0], s[1], s[2]))
def test_blocked_equals_generic_bitwise(spec):
    """the blocked shared-memory passes and the generic thread-per-node passes are the same arithmetic in
    the same order: messages, masks and decimation sequence must agree bit for bit"""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    Bn, n, k, alpha, Tn, seed = spec
    batch = cnfgen.random_batch(Bn, n, k, alpha, seed)
    E = batch[0].shape[1]
    init = po.init_state(E, randomized=True, rng=np.random.default_rng(seed))
    outs = []
    for generic in (False, True):
        ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]))
        ctx.enable_trace()
        ctx.simplify()
        ctx.load_state((T(init[0][0]), T(init[0][1])), (T(init[1][0]), T(init[1][1])))
        done = ctx.sp_run(Tn, 0.02, 25, True, sync=True, generic=generic)
        q, fs = ctx.store_state()
        m = ctx.get_masks()
        tr = C(ctx.trace()).astype(np.int64)
        tr = tr[np.lexsort((tr[:, 1], tr[:, 0]))]
        flags, counters, freeze = ctx.problem_flags()
        outs.append((np.int64(done), C(q[:, 0]), C(fs[:, 0]), C(m["av"]), C(m["af"]), C(m["sol"]), C(m["active"]), tr,
                     C(flags), C(freeze)))
    for a, b in zip(*outs):
        assert np.array_equal(a, b, equal_nan=True)


def test_blocked_path_with_signed_input_surveys():
    """surveys that arrive with a sign bit (not probabilities) must not confuse the blocked variable pass,
    which borrows that bit: the iteration that reads them runs on the generic passes"""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    batch = cnfgen.random_batch(12, 150, 3, 4.0, 31)
    E = batch[0].shape[1]
    init = po.init_state(E, randomized=True, rng=np.random.default_rng(31))
    init[1][1][::7, 0] *= -1.0    # dec_state function_state[:, 0]
    outs = []
    for generic in (False, True):
        ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]))
        ctx.simplify()
        ctx.load_state((T(init[0][0]), T(init[0][1])), (T(init[1][0]), T(init[1][1])))
        ctx.sp_run(40, 0.02, 25, True, sync=True, generic=generic)
        q, fs = ctx.store_state()
        m = ctx.get_masks()
        outs.append((C(q[:, 0]), C(fs[:, 0]), C(m["av"]), C(m["sol"])))
    for a, b in zip(*outs):
        assert np.array_equal(a, b, equal_nan=True)


# ------------------------------------------------------------------------------------------------
# neural model types (p-nd-np, np-nd-np) against the reference's own outputs
# ------------------------------------------------------------------------------------------------
def _neural_model(z):
    from pdp_solver_b200.nn import solver as S
    from pdp_solver_b200.nn import util as U
    H, MH, AH, MAH, CH = [int(x) for x in z["dims"]]
    clf = U.Perceptron(H, CH, 1)
    if str(z["model_type"]) == "p-nd-np":
        model = S.NeuralSurveyPropagatorSolver(dev(), "m", edge_dimension=1, meta_data_dimension=0, decimator_dimension=H,
                                               mem_hidden_dimension=MH, agg_hidden_dimension=AH, mem_agg_hidden_dimension=MAH,
                                               prediction_dimension=1, variable_classifier=clf, function_classifier=None,
                                               dropout=0, local_search_iterations=0, epsilon=0.5)
    else:
        model = S.NeuralPropagatorDecimatorSolver(dev(), "m", edge_dimension=1, meta_data_dimension=0, propagator_dimension=H,
                                                  decimator_dimension=H, mem_hidden_dimension=MH, agg_hidden_dimension=AH,
                                                  mem_agg_hidden_dimension=MAH, prediction_dimension=1, variable_classifier=clf,
                                                  function_classifier=None, dropout=0, local_search_iterations=0, epsilon=0.5)
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    for pair in str(z["w_alias"]).split(";"):
        if pair:
            k, src = pair.split("=")
            sd[k] = sd[src]
    model.load_state_dict(sd, strict=True)       # same keys and shapes as the reference's state_dict
    return model.to(dev()).eval()


@pytest.mark.parametrize("path", golden("neural_*.npz"), ids=name)
def test_neural_forward_vs_reference(path):
    """p-nd-np / np-nd-np forward() with the reference's weights and injected initial states: per-iteration
    predictions, iteration count, final states and final prediction (dense layers are fp32 library GEMMs whose
    accumulation order differs from the CPU reference's: 2e-4)."""
    z = load(path)
    model = _neural_model(z)
    gm, bvm, bfm, ef = T(z["graph_map"]), T(z["bvm"]), T(z["bfm"]), T(z["ef"])
    init = ((T(z["init_p0"]), T(z["init_p1"])), (T(z["init_d0"]), T(z["init_d1"])))
    preds = []
    orig = model._predictor.forward

    def hook(decimator_state, sat_problem, last_call=False):
        out = orig(decimator_state, sat_problem, last_call)
        preds.append(C(out[0]).reshape(-1))
        return out

    model._predictor.forward = hook

    def termination(active, prediction, sat_problem):
        raise RuntimeError("unreachable")
    termination._pdp_standard_termination = True
    with torch.no_grad():
        (vp, _), (ps, ds) = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm, edge_feature=ef,
                                  meta_data=None, is_training=False, iteration_num=int(z["T"]), check_termination=termination,
                                  simplify=True, batch_replication=1)
    ref = z["preds"]
    assert len(preds) == ref.shape[0], "iteration count differs"
    tol = 2e-4
    assert max(maxdiff(p, r) for p, r in zip(preds, ref)) < tol
    assert maxdiff(C(vp).reshape(-1), z["pred"]) < tol
    for ours, key in ((ps[0], "final_p0"), (ps[1], "final_p1"), (ds[0], "final_d0"), (ds[1], "final_d1")):
        assert maxdiff(C(ours), z[key]) < tol, key


@pytest.mark.parametrize("path", golden("rep_*.npz"), ids=name)
def test_batch_replication_vs_reference(path):
    """-b replicas through the CUDA path: replicated batch, fused loop with cross-replica termination
    (trainer.py:157-160), WalkSAT with the reference's draws, _deduplicate -- final prediction bit exact"""
    from pdp_solver_b200.engine import Context
    from tests.test_oracle_golden import _replicate
    z = load(path)
    b = int(z["b"])
    gm, bvm, bfm, ef = _replicate(z)
    ctx = Context(T(gm), T(bvm), T(bfm), T(ef))
    ctx.simplify()
    ctx.load_state((T(z["init_dq"]), T(z["init_df"])), (T(z["init_dq"]), T(z["init_df"])))
    ctx.sp_run(int(z["T"]), float(z["tol"]), int(z["t_max"]), True, batch_replication=b, sync=True)
    n_act = ctx.count_active_variables()
    assert n_act == z["fill"].shape[0]
    if n_act:
        ctx.random_fill(T(z["fill"]))
    pred, _ = ctx.walksat(int(z["W"]), float(z["epsilon"]), T(z["rand_var"]), T(z["rand_coin"]), batch_replication=b, sync=True)
    out, _ = ctx.deduplicate(b, pred)
    assert maxdiff(C(out), z["pred"]) == 0
    o = Context(T(z["graph_map"]), T(z["bvm"]), T(z["bfm"]), T(z["ef"]))
    solved, nun = o.cnf_eval(out)
    assert maxdiff(C(solved), z["solved"]) == 0 and maxdiff(C(nun), z["n_unsat"]) == 0


def test_full_size_properties():
    """BASELINE.json's full size (random 3-SAT, n = 1 000 000, alpha = 4.2; the oracle needs ~1 s per iteration
    there) through size-independent properties: (1) the blocked shared-memory passes and the generic
    thread-per-node passes agree bit for bit after 6 iterations; (2) the per-problem unsatisfied-clause count of
    the CNF evaluator equals the energy of the same assignment with every node active (two independent kernels);
    (3) the verdict is "solved" exactly where no clause is unsatisfied, and the merged prediction is 0/1 valued
    except for variables that occur in no clause (peeled to 0.5, reference solver.py:180-203)."""
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    batch = cnfgen.random_batch(2, 1000000, 3, 4.2, 99)
    gm, bvm, bfm, ef = [T(x) for x in batch]
    E, V, F = gm.shape[1], bvm.shape[0], bfm.shape[0]
    q3 = torch.full((E, 3), 1.0 / 3.0, device=dev())
    fs2 = torch.zeros((E, 2), device=dev())
    fs2[:, 0] = 0.5
    outs = []
    for generic in (False, True):
        ctx = Context(gm, bvm, bfm, ef, batch_size=2)
        errs, info = ctx.check_layout()
        assert info["blocked"] == 1 and errs[:6] == [0] * 6 and errs[6] == E and errs[7] == E
        ctx.simplify()
        ctx.load_state((q3, fs2), (q3, fs2))
        assert ctx.sp_run(6, 0.02, 100, True, sync=True, generic=generic) == 6
        q, fs = ctx.store_state()
        m = ctx.get_masks()
        outs.append((q[:, 0].clone(), fs[:, 0].clone(), m["av"].clone(), m["af"].clone()))
        if generic:
            break
        # (2) + (3) on the blocked context
        n_act = ctx.count_active_variables()
        ctx.random_fill(torch.rand(max(n_act, 1), device=dev()))
        pred, _ = ctx.walksat(20, 0.5, None, None, seed=5, sync=True)
        deg = torch.bincount(gm[0].long(), minlength=V)
        assert bool(((pred == 0) | (pred == 1) | ((pred == 0.5) & (deg == 0))).all())
        solved, n_unsat = ctx.cnf_eval(pred)
        energy, _ = ctx.energy(2 * pred - 1, torch.ones(V, device=dev()), torch.ones(F, device=dev()))
        assert torch.equal(energy, n_unsat)
        assert bool(((solved == 1) == (n_unsat == 0)).all())
        del ctx
    for a, b in zip(*outs):
        assert torch.equal(a, b) or bool(((a == b) | (a.isnan() & b.isnan())).all())


def test_full_size_decimation_paths():
    """the decimation regime at full size (2 x n = 1 000 000, T = 170: the problems converge near iteration 95 and then
    decimate almost every iteration): the product path (scores from the variable pass, range scans, frontier closure,
    CNF count without gathers) against the full-scan closure and against the generic passes (scores gathered by
    score_variable): identical decimation sequence, masks, solution, counters and messages"""
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    batch = cnfgen.random_batch(2, 1000000, 3, 4.2, 123)
    gm, bvm, bfm, ef = [T(x) for x in batch]
    outs = []
    for generic, full_closure in ((False, False), (False, True), (True, True)):
        ctx = Context(gm, bvm, bfm, ef, batch_size=2)
        ctx.enable_trace()
        ctx.simplify()
        ctx.load_state_const(1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0, 0.5, 0.0)
        done = ctx.sp_run(170, 0.02, 100, True, sync=True, generic=generic, full_closure=full_closure)
        q, fs = ctx.store_state()
        m = ctx.get_masks()
        tr = ctx.trace().to(torch.int64)
        tr = tr[torch.argsort(tr[:, 0] * (1 << 32) + tr[:, 1])]
        _, counters, _ = ctx.problem_flags()
        outs.append((torch.tensor(done), q[:, 0].clone(), fs[:, 0].clone(), m["av"].clone(), m["af"].clone(), m["sol"].clone(),
                     m["active"].clone(), tr.clone(), counters.clone()))
        del ctx
    assert outs[0][7].shape[0] >= 40, "too few decimation steps: the test does not exercise the regime"
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert a.shape == b.shape and (torch.equal(a, b) or bool(((a == b) | (a.isnan() & b.isnan())).all()))


@pytest.mark.parametrize("spec", [(96, 100, 3, 4.2, 200, 41), (12, 400, 3, 4.0, 150, 42), (40, 60, 3, 3.5, 150, 43)],
                         ids=lambda s: "B%d_n%d_k%d" % (s[0], s[1], s[2]))
def test_local_decimation_equals_grid_decimation(spec, monkeypatch):
    """the CTA-local decimation (score, arg-max, fix, UP / peel closure, CNF check, termination inside one CTA per
    problem) against the grid-wide phases, the latter with the frontier closure (lists of touched nodes) and with the
    full-scan closure: identical decimation sequence, masks, solutions, flags, messages"""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    Bn, n, k, alpha, Tn, seed = spec
    batch = cnfgen.random_batch(Bn, n, k, alpha, seed)
    E = batch[0].shape[1]
    init = po.init_state(E, randomized=False)
    outs = []
    # last variant: frontier lists of 8 entries -> they overflow and the closure falls back to the full scans mid-way
    for grid_dec, full_closure, cap in ((False, False, None), (True, False, None), (True, True, None), (True, False, "8")):
        if cap is None:
            monkeypatch.delenv("PDP_B200_FR_CAP", raising=False)
        else:
            monkeypatch.setenv("PDP_B200_FR_CAP", cap)
        ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]))
        ctx.enable_trace()
        ctx.simplify()
        ctx.load_state((T(init[0][0]), T(init[0][1])), (T(init[1][0]), T(init[1][1])))
        done = ctx.sp_run(Tn, 0.02, 25, True, sync=True, grid_decimation=grid_dec, full_closure=full_closure)
        q, fs = ctx.store_state()
        m = ctx.get_masks()
        tr = C(ctx.trace()).astype(np.int64)
        tr = tr[np.lexsort((tr[:, 1], tr[:, 0]))]
        flags, counters, freeze = ctx.problem_flags()
        outs.append((np.int64(done), C(q[:, 0]), C(fs[:, 0]), C(m["av"]), C(m["af"]), C(m["sol"]), C(m["active"]), C(m["is_sat"]),
                     tr, C(flags), C(freeze)))
    assert outs[0][8].shape[0] > 0, "no decimation happened: the test does not exercise anything"
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("ctas", ["1", "2"])
def test_both_cta_configurations(ctas, monkeypatch):
    """the blocked passes exist for one 1024-thread CTA and for two 512-thread CTAs per SM (chosen per batch by
    problem size); both must reproduce the generic passes bit for bit on the same batch"""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    monkeypatch.setenv("PDP_B200_CTAS", ctas)
    batch = cnfgen.mixed_batch([(3000, 3, 4.1), (500, 3, 3.9), (800, 5, 16.0), (40000, 3, 4.2), (100, 3, 4.0)], 71)
    E = batch[0].shape[1]
    init = po.init_state(E, randomized=True, rng=np.random.default_rng(71))
    outs = []
    for generic in (False, True):
        ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]))
        errs, info = ctx.check_layout()
        assert info["ctas"] == int(ctas) and errs[:6] == [0] * 6 and errs[6] == E and errs[7] == E
        ctx.simplify()
        ctx.load_state((T(init[0][0]), T(init[0][1])), (T(init[1][0]), T(init[1][1])))
        ctx.sp_run(60, 0.02, 25, True, sync=True, generic=generic)
        q, fs = ctx.store_state()
        m = ctx.get_masks()
        outs.append((C(q[:, 0]), C(fs[:, 0]), C(m["av"]), C(m["af"]), C(m["sol"]), C(m["active"])))
    for a, b in zip(*outs):
        assert np.array_equal(a, b, equal_nan=True)


def test_interleaved_problems_take_the_generic_paths():
    """batch maps that are not sorted by problem (nodes of different problems interleaved): no blocked layout, no
    CTA-local decimation, grid-wide WalkSAT -- the result must equal the sorted batch's, node for node"""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    batch = cnfgen.random_batch(10, 80, 3, 4.0, 81)
    gm, bvm, bfm, ef = batch
    V, F, E = bvm.shape[0], bfm.shape[0], gm.shape[1]
    rng = np.random.default_rng(81)
    vperm, fperm = rng.permutation(V), rng.permutation(F)        # new index of old node
    gm2 = np.stack([vperm[gm[0]], fperm[gm[1]]]).astype(np.int32)
    bvm2 = np.empty_like(bvm); bvm2[vperm] = bvm
    bfm2 = np.empty_like(bfm); bfm2[fperm] = bfm
    # keep clause-major edge order with ascending edge index inside a clause (same accumulation order per node)
    order = np.argsort(gm2[1], kind="stable")
    init = po.init_state(E, randomized=True, rng=np.random.default_rng(82))
    res = []
    for (g_, bv_, bf_, ef_, eord) in ((gm, bvm, bfm, ef, np.arange(E)), (gm2[:, order], bvm2, bfm2, ef[order], order)):
        ctx = Context(T(g_), T(bv_), T(bf_), T(ef_))
        errs, info = ctx.check_layout()
        assert errs[0] == 0
        ctx.simplify()
        st = [[x[eord] for x in pair] for pair in init]
        ctx.load_state((T(st[0][0]), T(st[0][1])), (T(st[1][0]), T(st[1][1])))
        done = ctx.sp_run(120, 0.02, 25, True, sync=True)
        m = ctx.get_masks()
        n_act = ctx.count_active_variables()
        res.append((info["blocked"], done, C(m["av"]), C(m["af"]), C(m["sol"]), C(m["active"]), n_act))
    (b1, d1, av1, af1, sol1, act1, n1), (b2, d2, av2, af2, sol2, act2, n2) = res
    assert b1 == 1 and b2 == 0
    assert d1 == d2 and n1 == n2 and np.array_equal(act1, act2)
    assert np.array_equal(av1, av2[vperm]) and np.array_equal(af1, af2[fperm]) and np.array_equal(sol1, sol2[vperm])


# ------------------------------------------------------------------------------------------------
# the remaining model types: reinforce and np-d-np against the reference's own outputs
# ------------------------------------------------------------------------------------------------
def _standard_termination():
    def termination(active, prediction, sat_problem):
        raise RuntimeError("unreachable")
    termination._pdp_standard_termination = True
    return termination


@pytest.mark.parametrize("path", golden("reinforce_*.npz"), ids=name)
def test_reinforce_forward_vs_reference(path):
    """model type `reinforce` with the reference's coin draws injected: per-iteration merged predictions and the
    iteration at which every problem retired are exact, final messages within the survey tolerance"""
    from pdp_solver_b200.nn import solver as S
    z = load(path)
    model = S.ReinforceSurveyPropagatorSolver(dev(), "r", pi=float(z["pi"]), decimation_probability=float(z["p_dec"]),
                                              local_search_iterations=0, epsilon=0.5)
    coins = iter(z["coins"].tolist())
    model._decimator._coin_source = lambda: next(coins)
    gm, bvm, bfm, ef = T(z["graph_map"]), T(z["bvm"]), T(z["bfm"]), T(z["ef"])
    init = ((T(z["init_p0"]), T(z["init_p1"])), (T(z["init_d0"]), T(z["init_d1"])))
    merged = []
    orig = S.SATProblem.update_solution

    def hook(self, variable_prediction):
        out = orig(self, variable_prediction)
        merged.append(C(out).reshape(-1))
        return out

    S.SATProblem.update_solution = hook
    try:
        with torch.no_grad():
            (vp, _), (ps, ds) = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm,
                                      edge_feature=ef, meta_data=None, is_training=False, iteration_num=int(z["T"]),
                                      check_termination=_standard_termination(), simplify=True, batch_replication=1)
    finally:
        S.SATProblem.update_solution = orig
    ref = z["preds"]
    assert int(model.last_iterations.item()) == ref.shape[0] == z["coins"].shape[0], "iteration count differs"
    assert len(merged) == ref.shape[0] + 1                      # + the final merge of solver.py:342-348
    for it, (a, b) in enumerate(zip(merged, ref)):
        assert maxdiff(a, b) == 0, "prediction differs at iteration %d" % (it + 1)
    assert maxdiff(C(vp).reshape(-1), z["pred"]) == 0
    assert maxdiff(C(ps[0]), z["final_q"]) < SURVEY_TOL and maxdiff(C(ps[1]), z["final_f"]) < SURVEY_TOL
    assert ds[1] is ps[1]                                       # the decimator edits the propagator's state in place


@pytest.mark.parametrize("path", golden("npdnp_*.npz"), ids=name)
def test_npdnp_forward_vs_reference(path):
    """model type `np-d-np` with the reference's weights and injected initial states: identical decimation sequence,
    final masks, iteration count; final hidden states within the GEMM tolerance"""
    from pdp_solver_b200.nn import solver as S
    z = load(path)
    H, MH, AH, MAH, CH = [int(x) for x in z["dims"]]
    model = S.NeuralSequentialDecimatorSolver(dev(), "m", edge_dimension=1, meta_data_dimension=0, propagator_dimension=H,
                                              decimator_dimension=H, mem_hidden_dimension=MH, agg_hidden_dimension=AH,
                                              mem_agg_hidden_dimension=MAH, classifier_dimension=CH, dropout=0,
                                              tolerance=float(z["tol"]), t_max=int(z["t_max"]), local_search_iterations=0,
                                              epsilon=0.5)
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    for pair in str(z["w_alias"]).split(";"):
        if pair:
            k, src = pair.split("=")
            sd[k] = sd[src]
    model.load_state_dict(sd, strict=True)
    model = model.to(dev()).eval()
    gm, bvm, bfm, ef = T(z["graph_map"]), T(z["bvm"]), T(z["bfm"]), T(z["ef"])
    model.get_init_state(gm, bvm, bfm, ef, None, randomized=False, batch_replication=1)     # resets the decimator
    init = ((T(z["init_p0"]), T(z["init_p1"])), (T(z["init_d0"]), T(z["init_d1"])))
    cb = _standard_termination() if bool(z["with_termination"]) else None
    with torch.no_grad():
        (vp, _), (ps, ds) = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm,
                                  edge_feature=ef, meta_data=None, is_training=False, iteration_num=int(z["T"]),
                                  check_termination=cb, simplify=True, batch_replication=1)
    assert int(model.last_iterations.item()) == int(z["iterations"])
    ours = [(int(i), int(s)) for idx, sg in model._decimator.decimation_log for i, s in sorted(zip(C(idx).tolist(), C(sg).tolist()))]
    assert ours == [(int(i), int(s)) for _, i, s in z["events"].tolist()]
    m = model.last_problem._ctx.get_masks()
    assert maxdiff(C(m["av"]), z["av"]) == 0 and maxdiff(C(m["af"]), z["af"]) == 0
    decided = z["av"] == 0
    assert maxdiff(C(vp).reshape(-1)[decided], z["pred"][decided]) == 0
    assert int((~decided).sum()) == z["fill"].shape[0]
    assert maxdiff(C(ps[0]), z["final_p0"]) < 2e-4 and maxdiff(C(ps[1]), z["final_p1"]) < 2e-4


# ------------------------------------------------------------------------------------------------
# degenerate batches through the module interface
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["no_clauses", "unit", "mixed_with_empty", "contradiction", "replicated"])
def test_degenerate_batches(case):
    """problems without clauses (E = 0 for the whole batch, or one empty problem among others), a lone unit clause, a
    one-variable contradiction, batch replication over an empty problem: shapes, verdicts (against the C oracle's CNF check
    of the returned prediction) and the values the closure alone decides"""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.nn import solver as S
    probs, b = {
        "no_clauses": ([(5, [])], 1),
        "unit": ([(3, [[2]])], 1),
        "mixed_with_empty": ([(4, [[1, -2, 3], [-1, 2, 4]]), (3, []), (2, [[1, 2], [-1, -2]])], 1),
        "contradiction": ([(1, [[1], [-1]])], 1),
        "replicated": ([(4, [[1, -2, 3]]), (3, [])], 2),
    }[case]
    batch = cnfgen.from_clauses(probs)
    gm, bvm, bfm, ef = [T(x) for x in batch]
    model = S.SurveyPropagatorSolver(dev(), "x", tolerance=0.02, t_max=10, local_search_iterations=4, epsilon=0.5)
    init = model.get_init_state(gm, bvm, bfm, ef, None, randomized=False, batch_replication=b)
    (pred, _), (ps, ds) = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm, edge_feature=ef,
                                meta_data=None, is_training=False, iteration_num=6, check_termination=_standard_termination(),
                                batch_replication=b)
    E, V = batch[0].shape[1], batch[1].shape[0]
    p = C(pred).reshape(-1)
    assert p.shape == (V,) and tuple(ps[0].shape) == (E, 3) and tuple(ps[1].shape) == (E, 2)
    assert np.isin(p, [0.0, 0.5, 1.0]).all()
    from pdp_solver_b200.engine import Context
    solved, n_unsat = Context(gm, bvm, bfm, ef).cnf_eval(pred)
    o_solved, o_unsat = po.Oracle(*batch).cnf_eval(p)
    assert maxdiff(C(solved), o_solved) == 0 and maxdiff(C(n_unsat), o_unsat) == 0
    if case == "no_clauses":
        assert (p == 0.5).all() and C(solved).tolist() == [1.0]
    if case == "unit":
        assert p.tolist() == [0.5, 1.0, 0.5]
    if case == "contradiction":
        assert C(solved).tolist() == [0.0] and p.tolist() == [0.5]
    if case == "mixed_with_empty":
        assert C(solved)[:2].tolist() == [1.0, 1.0] and (p[4:7] == 0.5).all()


# ------------------------------------------------------------------------------------------------
# round 2: parity at BASELINE size against the oracle, product-math replay, stateless CNF check
# ------------------------------------------------------------------------------------------------
def test_full_size_vs_oracle():
    """One problem of BASELINE.json configs[3] (random 3-SAT, n = 1 000 000, m/n = 4.2, E = 12.6 M) through the
    shipped library against the C oracle: simplify (unit propagation + peeling) masks and solution exact, T = 6
    SP iterations from the predict path's initial messages (q = 1/3, eta = 0.5, the state bench.py starts from) with
    surveys within 1e-4, energy of a random assignment, 2 WalkSAT iterations with injected draws and the CNF verdict
    exact.  (From a RANDOM initial state a handful of the 12.6 M edges sit at eta ~ 1, where log(1 - eta) amplifies the
    1-ulp difference between CUDA's and glibc's logf/expf to 5e-3 within 4 iterations on both sides of the comparison;
    the chaotic cases are pinned step by step in test_product_math_single_step_replay.)"""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    n, Tn, W, tol, t_max, eps = 1000000, 6, 2, 0.02, 100, 0.5
    batch = cnfgen.random_batch(1, n, 3, 4.2, 4242)
    E = batch[0].shape[1]
    rng = np.random.default_rng(4242)
    init = po.init_state(E, randomized=False)
    po.set_num_threads(__import__("os").cpu_count() or 1)
    o, done, fill, rv, rc, pred, wit = _oracle_forward(batch, init, Tn, tol, t_max, W, eps, rng)
    om, (oq, ofs), _ = o.after_run

    ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]))
    ctx.simplify()
    ctx.load_state((T(init[0][0]), T(init[0][1])), (T(init[1][0]), T(init[1][1])))
    gdone = ctx.sp_run(Tn, tol, t_max, True, sync=True)
    assert gdone == done
    m = ctx.get_masks()
    assert (C(m["av"]) == om["av"]).all() and (C(m["af"]) == om["af"]).all() and maxdiff(C(m["sol"]), om["sol"]) == 0
    assert 0 < int(om["av"].sum()) < n          # the peel did something and left something
    q, fs = ctx.store_state()
    assert maxdiff(C(fs[:, 0]), ofs[:, 0]) <= SURVEY_TOL and maxdiff(C(q[:, 0]), oq[:, 0]) <= SURVEY_TOL
    # integer operators on a random assignment over the residual formula
    asg = (rng.integers(0, 2, size=n).astype(np.float32) * 2 - 1) * om["av"]
    oe, ouf = o.energy(asg, om["av"], om["af"])
    ge, guf = ctx.energy(T(asg), m["av"], m["af"])
    assert maxdiff(C(ge), oe) == 0 and maxdiff(C(guf), ouf) == 0
    n_act = ctx.count_active_variables()
    assert n_act == int(om["av"].sum())
    ctx.random_fill(T(fill))
    gpred, git = ctx.walksat(W, eps, T(rv), T(rc), sync=True)
    assert git == wit and maxdiff(C(gpred), pred) == 0
    s1, u1 = ctx.cnf_eval(gpred)
    s2, u2 = o.cnf_eval(pred)
    assert maxdiff(C(s1), s2) == 0 and maxdiff(C(u1), u2) == 0


@pytest.mark.parametrize("spec", [s for s in ORACLE_SPECS if s[5] not in (8, 9)], ids=lambda s: "B%d_n%d_k%d" % (s[0], s[1], s[2]))
def test_product_math_single_step_replay(spec):
    """The numerically chaotic instances (whole trajectories are only comparable in strict-math mode) with the SHIPPED
    library: every sampled iteration is replayed from the oracle's own state one step back through the persistent
    blocked kernel -- surveys and q_u within 1e-4 on the problems that were running, NaN onsets at the same edges."""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    Bn, n, k, alpha, Tn, seed = spec
    batch = cnfgen.random_batch(Bn, n, k, alpha, seed)
    gm, bvm = batch[0], batch[1]
    E = gm.shape[1]
    rng = np.random.default_rng(seed)
    init = po.init_state(E, randomized=(seed % 2 == 0), rng=rng)
    tol, t_max = 0.02, 25
    o = po.Oracle(*batch, strict=False)
    o.simplify()
    o.set_state(*init)
    states = []
    for t in range(Tn):
        masks = o.masks()
        states.append((o.state(), masks))
        if o.run(1, tol, t_max, True) == 0:
            break
    states.append((o.state(), o.masks()))
    ctx = Context(T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]))
    eprob = bvm[gm[0]]
    worst, checked = 0.0, 0
    for t in range(1, len(states) - 1, max(1, len(states) // 12)):
        (q0, f0), m0 = states[t]
        (q1, f1), _ = states[t + 1]
        running = m0["active"].astype(bool)
        clean = np.ones(o.B, bool)
        clean[eprob[np.isnan(f0[:, 0]) | np.isnan(q0[:, 0])]] = False      # sticky-NaN problems need their history
        sel = (running & clean)[eprob]
        if not sel.any():
            continue
        ctx.reset()
        ctx.set_masks(T(m0["av"]), T(m0["af"]), T(m0["sol"]))
        ctx.load_state((T(q0), T(f0)), (T(q0), T(f0)))
        assert ctx.sp_run(1, tol, t_max, False, sync=True) == 1
        q, fs = ctx.store_state()
        ge, gq = C(fs[:, 0])[sel], C(q[:, 0])[sel]
        oe, oq = f1[:, 0][sel], q1[:, 0][sel]
        assert (np.isnan(ge) == np.isnan(oe)).all() and (np.isnan(gq) == np.isnan(oq)).all(), t
        worst = max(worst, maxdiff(ge, oe), maxdiff(gq, oq))
        checked += 1
    assert checked >= 3 and worst <= SURVEY_TOL, (checked, worst)


def test_stateless_cnf_evaluator():
    """SatCNFEvaluator (reference util.py:203-236) runs on the caller's edge list without a context: int64 maps, permuted
    edge order and a second call on different tensors of the same shape (the stale-cache case of round 1) agree with the
    context's evaluator and the oracle."""
    from oracle import pdp_oracle as po
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.engine import Context
    from pdp_solver_b200.nn.util import SatCNFEvaluator
    ev = SatCNFEvaluator(dev())
    for seed in (1, 2):
        batch = cnfgen.mixed_batch([(60, 3, 4.2), (40, 5, 18.0), (80, 3, 3.0)], seed)
        gm, bvm, bfm, ef = batch
        rng = np.random.default_rng(seed)
        pred = rng.choice(np.array([0.0, 0.5, 1.0], np.float32), size=bvm.shape[0])
        o = po.Oracle(*batch, strict=False)
        s_ref, u_ref = o.cnf_eval(pred)
        perm = rng.permutation(gm.shape[1])
        s1, u1 = ev(T(pred).unsqueeze(1), T(gm[:, perm]).long(), T(bvm).long(), T(bfm).long(), T(ef[perm]).unsqueeze(1), None)
        ctx = Context(T(gm), T(bvm), T(bfm), T(ef))
        s2, u2 = ctx.cnf_eval(T(pred))
        assert maxdiff(C(s1[:, 0]), s_ref) == 0 and maxdiff(C(u1[:, 0]), u_ref) == 0
        assert maxdiff(C(s2), s_ref) == 0 and maxdiff(C(u2), u_ref) == 0


def test_custom_termination_callback_is_honoured():
    """A check_termination callable that is not the trainer's own method is called after every iteration (reference
    solver.py:376-384) instead of being replaced by the in-kernel check: a re-implementation of the trainer's rule gives
    the fused path's result bit for bit, a no-op callback (the base class's, reference base.py:303-305) keeps every
    problem running for all T iterations, and a subclass override of the trainer's method is not mistaken for it."""
    from pdp_solver_b200 import cnfgen
    from pdp_solver_b200.nn import solver as pdp_solver
    from pdp_solver_b200.nn.util import SatCNFEvaluator
    from pdp_solver_b200.trainer import SatFactorGraphTrainer
    batch = cnfgen.random_batch(24, 60, 3, 3.7, 19)
    gm, bvm, bfm, ef = T(batch[0]), T(batch[1]), T(batch[2]), T(batch[3]).unsqueeze(1)
    model = pdp_solver.SurveyPropagatorSolver(dev(), "p-d-p", tolerance=0.02, t_max=30, local_search_iterations=0, epsilon=0.5)
    ev = SatCNFEvaluator(dev())
    calls = {"n": 0}

    def own_rule(active, prediction, sat_problem):       # trainer.py:150-162 restated by a caller
        calls["n"] += 1
        out, _ = ev(prediction[0], sat_problem._graph_map, sat_problem._batch_variable_map, sat_problem._batch_function_map,
                    sat_problem._edge_feature, None)
        active[(active[:, 0] != 0) & (out[:, 0] > 0.5), 0] = 0

    def noop(active, prediction, sat_problem):
        calls["n"] += 1

    def tagged(active, prediction, sat_problem):
        raise RuntimeError("unreachable")
    tagged._pdp_standard_termination = True

    def run(cb, T_iters=120):
        torch.manual_seed(3)
        init = model.get_init_state(gm, bvm, bfm, ef, None, randomized=False, batch_replication=1)
        (pred, _), _ = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm, edge_feature=ef,
                             meta_data=None, is_training=False, iteration_num=T_iters, check_termination=cb, batch_replication=1)
        m = model.last_problem._ctx.get_masks()
        return C(pred), C(m["av"]), C(m["active"]), int(model.last_iterations.item())

    fused = run(tagged)
    stepped = run(own_rule)
    assert calls["n"] == stepped[3] and stepped[3] == fused[3]
    for a, b in zip(fused[:3], stepped[:3]):
        assert np.array_equal(a, b)
    calls["n"] = 0
    free = run(noop, 40)
    assert calls["n"] == 40 and free[3] == 40

    class Sub(SatFactorGraphTrainer):
        def _check_recurrence_termination(self, active, prediction, sat_problem):
            pass
    assert pdp_solver._is_standard_termination(SatFactorGraphTrainer._check_recurrence_termination)
    assert not pdp_solver._is_standard_termination(Sub._check_recurrence_termination)


TC_LINEAR_SPECS = [(1000, 150, 1, 100, 1), (128, 100, 0, 50, 1), (333, 50, 1, 100, 1), (20000, 100, 0, 150, 1), (513, 150, 0, 2, 0), (77, 150, 0, 1, 1)]


@pytest.mark.parametrize("spec", TC_LINEAR_SPECS, ids=lambda s: "E%d_%d+%d_to_%d" % s[:4])
def test_tensor_core_linear_vs_fp64(spec):
    """pdp_edge_mlp_forward (tcgen05 kind::tf32, three-term split) against the same layer in fp64: the dense layers of
    MessageAggregator.forward (reference util.py:51-77) with the edge feature read as a second source, logsigmoid and the
    edge mask fused.  fp32-grade accuracy: 2e-5 absolute on O(1)-O(10) pre-activations (torch's own fp32 layer: ~1e-6)."""
    from pdp_solver_b200.nn import tensor_ops as TO
    E, k1, k2, n, act = spec
    torch.manual_seed(E + n)
    lin = torch.nn.Linear(k1 + k2, n, bias=(n != 50)).to(dev())
    x1 = torch.randn(E, k1, device=dev())
    src = [x1] + ([torch.sign(torch.randn(E, k2, device=dev()))] if k2 else [])
    mask = (torch.rand(E, 1, device=dev()) > 0.2).float()
    out = TO.TensorLinear(lin)(src, act=act, row_mask=mask)
    ref = torch.cat(src, 1).double() @ lin.weight.double().t()
    if lin.bias is not None:
        ref = ref + lin.bias.double()
    if act:
        ref = torch.nn.functional.logsigmoid(ref)
    ref = ref * mask.double()
    assert out.shape == (E, n)
    assert (out.double() - ref).abs().max().item() < 2e-5
    # a changed parameter rebuilds the weight image
    with torch.no_grad():
        lin.weight.mul_(0.5)
    out2 = TO.TensorLinear(lin)(src, act=0)
    ref2 = torch.cat(src, 1).double() @ lin.weight.double().t() + (lin.bias.double() if lin.bias is not None else 0.0)
    assert (out2.double() - ref2).abs().max().item() < 2e-5


@pytest.mark.parametrize("spec", [(1000, 150, 1, 150), (257, 3, 1, 150), (129, 2, 1, 150), (20000, 150, 1, 150)], ids=lambda s: "E%d_%d+%d_h%d" % s)
def test_tensor_core_gru_vs_fp64(spec):
    """pdp_edge_gru_forward against torch.nn.GRUCell in fp64 (the two cells of NeuralDecimator.forward, reference
    pdp_decimate.py:51-87), with the frozen-problem blend of rows whose mask is 0"""
    from pdp_solver_b200.nn import tensor_ops as TO
    E, kx1, kx2, H = spec
    torch.manual_seed(E)
    cell = torch.nn.GRUCell(kx1 + kx2, H).to(dev())
    x1, x2 = torch.randn(E, kx1, device=dev()), torch.sign(torch.randn(E, kx2, device=dev()))
    h = torch.rand(E, H, device=dev()) * 2 - 1
    mask = (torch.rand(E, 1, device=dev()) > 0.2).float()
    out = TO.TensorGRU(cell)([x1, x2], h, row_mask=mask)
    c64 = torch.nn.GRUCell(kx1 + kx2, H).to(dev()).double()
    c64.load_state_dict({k: v.double() for k, v in cell.state_dict().items()})
    ref = c64(torch.cat((x1, x2), 1).double(), h.double())
    ref = mask.double() * ref + (1 - mask.double()) * h.double()
    assert (out.double() - ref).abs().max().item() < 2e-5
    assert torch.equal(out[mask[:, 0] == 0], h[mask[:, 0] == 0])      # frozen rows keep their state bit for bit


@pytest.mark.parametrize("path", golden("neural_*.npz"), ids=name)
def test_neural_tensor_core_path_equals_library_gemm_path(path, monkeypatch):
    """the same forward() with the dense layers on tcgen05 (the product path) and on the library GEMMs (PDP_B200_NN=torch):
    predictions and final states agree to 5e-5 -- the tensor-core path is an fp32-grade drop-in, not a different model"""
    z = load(path)
    gm, bvm, bfm, ef = T(z["graph_map"]), T(z["bvm"]), T(z["bfm"]), T(z["ef"])
    init = ((T(z["init_p0"]), T(z["init_p1"])), (T(z["init_d0"]), T(z["init_d1"])))

    def termination(active, prediction, sat_problem):
        raise RuntimeError("unreachable")
    termination._pdp_standard_termination = True
    outs = []
    for mode in ("tcgen05", "torch"):
        monkeypatch.setenv("PDP_B200_NN", mode)
        model = _neural_model(z)
        with torch.no_grad():
            (vp, _), (ps, ds) = model(init_state=init, graph_map=gm, batch_variable_map=bvm, batch_function_map=bfm, edge_feature=ef,
                                      meta_data=None, is_training=False, iteration_num=int(z["T"]), check_termination=termination,
                                      simplify=True, batch_replication=1)
        outs.append([C(vp), C(ps[0]), C(ps[1]), C(ds[0]), C(ds[1])])
    for a, b in zip(*outs):
        assert maxdiff(a, b) < 5e-5
