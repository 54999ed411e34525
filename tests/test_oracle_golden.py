"""CPU: the C oracle (oracle/pdp_oracle.c) against golden vectors produced by the reference itself
(oracle/make_golden.py).  This is what pins the oracle; the CUDA path is then checked against both."""
import numpy as np
import pytest

from oracle import pdp_oracle as po
from tests.helpers import golden, load, maxdiff, name

FP_TOL = 2e-6      # libm (glibc vs torch/sleef) differences on single operator applications


@pytest.mark.parametrize("path", golden("ops_*.npz"), ids=name)
def test_operators(path):
    z = load(path)
    o = po.Oracle(z["graph_map"], z["bvm"], z["bfm"], z["ef"], strict=True)
    # SurveyPropagator.forward (pdp_propagate.py:139-221)
    q1, f1 = o.sp_step(z["dq"], z["df"], None, z["pq"], z["pf"], None)
    assert maxdiff(q1, z["sp1_q"]) < FP_TOL and maxdiff(f1, z["sp1_f"]) < FP_TOL
    q2, f2 = o.sp_step(z["dq"], z["df"], z["em"], z["pq"], z["pf"], z["active"])
    assert maxdiff(q2, z["sp2_q"]) < FP_TOL and maxdiff(f2, z["sp2_f"]) < FP_TOL
    # SurveyScorer.forward (pdp_predict.py:155-192)
    assert maxdiff(o.score(z["df"], z["af"]), z["score"]) < FP_TOL
    assert maxdiff(o.score(z["df"], np.ones_like(z["af"])), z["score_all"]) < FP_TOL
    # integer-valued operators: bit exact
    solved, nun = o.cnf_eval(z["vp"])
    assert maxdiff(solved, z["solved"]) == 0 and maxdiff(nun, z["n_unsat"]) == 0
    en, uf = o.energy(z["asg"], z["av"], z["af"])
    assert maxdiff(en, z["energy"]) == 0 and maxdiff(uf, z["unsat_fn"]) == 0
    assert maxdiff(o.energy_diff(z["asg"], z["av"], z["em"]), z["delta"]) == 0


@pytest.mark.parametrize("path", golden("simplify_*.npz"), ids=name)
def test_simplify(path):
    z = load(path)
    o = po.Oracle(z["graph_map"], z["bvm"], z["bfm"], z["ef"], strict=True)
    o.simplify()
    m = o.masks()
    for k in ("av", "af", "sol", "is_sat"):
        assert maxdiff(m[k], z[k]) == 0, k
    o.set_variables(z["asg"])
    m = o.masks()
    for k in ("av", "af", "sol", "is_sat"):
        assert maxdiff(m[k], z[k + "2"]) == 0, k


def _replay(z, strict):
    o = po.Oracle(z["graph_map"], z["bvm"], z["bfm"], z["ef"], strict=strict)
    T, W = int(z["T"]), int(z["W"])
    o.simplify()
    worst = 0.0
    if T > 0:
        o.set_state((z["init_pq"], z["init_pf"]), (z["init_dq"], z["init_df"]))
        for t in range(T):
            na = o.iterate(float(z["tol"]), float(z["t_max"]), True)
            q, fs = o.state()
            m = o.masks()
            worst = max(worst, maxdiff(fs[:, 0], z["eta"][t]), maxdiff(q[:, 0], z["qu"][t]))
            assert (m["av"] == z["av"][t]).all() and (m["af"] == z["af"][t]).all(), t
            assert (m["active"] == z["active"][t]).all(), t
            assert maxdiff(m["sol"], z["sol"][t]) == 0, t
            assert maxdiff(o.counters(), z["counters"][t]) == 0, t
            if na <= 0:
                assert t + 1 == z["eta"].shape[0]
                break
        tr = o.trace()
        assert tr.shape == z["events"].shape and (tr == z["events"]).all()
    n_act = o.count_active_variables()
    assert z["fill"].shape[0] == n_act
    if n_act:
        o.random_fill(z["fill"])
    n_w = z["rand_var"].shape[0]
    pred, it = o.local_search(W if n_w else 0, float(z["epsilon"]), z["rand_var"] if n_w else None,
                              z["rand_coin"] if n_w else None)
    assert it == n_w or it == W
    assert maxdiff(pred, z["pred"]) == 0
    solved, nun = o.cnf_eval(pred)
    assert maxdiff(solved, z["solved"]) == 0 and maxdiff(nun, z["n_unsat"]) == 0
    return worst


@pytest.mark.parametrize("strict", [True, False], ids=["strict", "isolated"])
@pytest.mark.parametrize("path", golden("traj_*.npz") + golden("walksat_*.npz"), ids=name)
def test_forward_trajectory(path, strict):
    """Whole forward(): identical decimation sequence, masks, solutions, counters, WalkSAT flips and
    final prediction; surveys within 1e-3 free-running (ulp-level libm differences amplified by
    non-convergent instances -- operator-level tolerance is FP_TOL above)."""
    z = load(path)
    worst = _replay(z, strict)
    assert worst < 1e-3


def _replicate(z):
    """replica r of problem j gets problem id r*B + j (reference solver.py:56-82)"""
    gm, bvm, bfm, ef, b = z["graph_map"], z["bvm"], z["bfm"], z["ef"], int(z["b"])
    V, F, B = bvm.shape[0], bfm.shape[0], int(bvm.max()) + 1
    gmr = np.concatenate([gm + np.array([[r * V], [r * F]], dtype=gm.dtype) for r in range(b)], axis=1)
    return (gmr.astype(np.int32), np.concatenate([bvm + r * B for r in range(b)]).astype(np.int32),
            np.concatenate([bfm + r * B for r in range(b)]).astype(np.int32), np.concatenate([ef] * b).astype(np.float32))


@pytest.mark.parametrize("path", golden("rep_*.npz"), ids=name)
def test_batch_replication(path):
    """-b replicas: termination across replicas, WalkSAT with the recorded draws, _deduplicate (solver.py:401-431)"""
    from oracle import pdp_oracle as po
    z = load(path)
    b = int(z["b"])
    o = po.Oracle(*_replicate(z), strict=True)
    o.simplify()
    o.set_state((z["init_dq"], z["init_df"]), (z["init_dq"], z["init_df"]))
    o.run(int(z["T"]), float(z["tol"]), int(z["t_max"]), True, b)
    n_act = o.count_active_variables()
    assert n_act == z["fill"].shape[0]
    if n_act:
        o.random_fill(z["fill"])
    pred, _ = o.local_search(int(z["W"]), float(z["epsilon"]), z["rand_var"], z["rand_coin"], b)
    out, _ = o.deduplicate(b, pred)
    assert maxdiff(out, z["pred"]) == 0


@pytest.mark.parametrize("path", golden("reinforce_*.npz"), ids=name)
def test_reinforce_vs_reference(path):
    """model type `reinforce` restated on the oracle's operators (pins the pi terms of SP and of the scorer): merged
    predictions of every iteration and the stopping iteration exact, final messages to the survey tolerance"""
    z = load(path)
    o = po.Oracle(z["graph_map"], z["bvm"], z["bfm"], z["ef"])
    merged, q, f = po.reinforce_run(o, (z["init_p0"], z["init_p1"]), (z["init_d0"], z["init_d1"]), int(z["T"]),
                                            float(z["pi"]), float(z["p_dec"]), z["coins"])
    assert len(merged) == z["preds"].shape[0]
    for a, b in zip(merged, z["preds"]):
        assert maxdiff(a, b) == 0
    assert maxdiff(merged[-1], z["pred"]) == 0
    assert maxdiff(q, z["final_q"]) < 1e-4 and maxdiff(f, z["final_f"]) < 1e-4
