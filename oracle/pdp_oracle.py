"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the C oracle (oracle/pdp_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this module.  The product package (pdp_solver_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pdp_oracle.c")
_SO = os.path.join(_HERE, "libpdp_oracle.so")

_lib = None


def build(force=False):
    """Compiles the C oracle with gcc (OpenMP so the CPU baseline can use every host core)."""
    if (not force) and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-fopenmp", "-fno-fast-math", "-ffp-contract=off", "-shared", "-fPIC",
           "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
            build()
        L = ctypes.CDLL(_SO)
        P = ctypes.c_void_p
        I64 = ctypes.c_int64
        F32 = ctypes.c_float
        L.ora_create.restype = P
        L.ora_create.argtypes = [I64, I64, I64, I64, P, P, P, P, ctypes.c_int]
        L.ora_destroy.argtypes = [P]
        L.ora_set_state.argtypes = [P, P, P, P, P]
        L.ora_set_masks.argtypes = [P, P, P, P]
        L.ora_get_state.argtypes = [P, P, P]
        L.ora_get_masks.argtypes = [P, P, P, P, P, P, P]
        L.ora_get_counters.argtypes = [P, P]
        L.ora_iters_done.restype = I64
        L.ora_iters_done.argtypes = [P]
        L.ora_trace_len.restype = I64
        L.ora_trace_len.argtypes = [P]
        L.ora_get_trace.argtypes = [P, P]
        L.ora_edge_mask_set.argtypes = [P]
        L.ora_simplify.argtypes = [P]
        L.ora_set_variables.argtypes = [P, P]
        L.ora_sp_step.argtypes = [P, P, P, P, P, P, P, F32, P, P]
        L.ora_score.argtypes = [P, P, P, F32, P]
        L.ora_cnf_eval.argtypes = [P, P, P, P]
        L.ora_energy.argtypes = [P, P, P, P, P, P]
        L.ora_energy_diff.argtypes = [P, P, P, P, P]
        L.ora_compute_edge_mask.argtypes = [P]
        L.ora_iterate.restype = I64
        L.ora_iterate.argtypes = [P, F32, F32, ctypes.c_int, ctypes.c_int, F32]
        L.ora_run.restype = I64
        L.ora_run.argtypes = [P, I64, F32, F32, ctypes.c_int, ctypes.c_int, F32]
        L.ora_count_active_variables.restype = I64
        L.ora_count_active_variables.argtypes = [P]
        L.ora_random_fill.argtypes = [P, P]
        L.ora_local_search.restype = I64
        L.ora_local_search.argtypes = [P, I64, F32, ctypes.c_int, P, P, P]
        L.ora_deduplicate.argtypes = [P, ctypes.c_int, P, P, P]
        L.ora_num_threads.restype = ctypes.c_int
        L.ora_set_math_mode.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class Oracle(object):
    """One batch of CNFs plus the reference's mutable solver state (SATProblem + SequentialDecimator)."""

    def __init__(self, graph_map, batch_variable_map, batch_function_map, edge_feature, strict=False,
                 batch_size=None):
        L = lib()
        gm = np.ascontiguousarray(np.asarray(graph_map, dtype=np.int32))
        self.E = int(gm.shape[1])
        self.gm = gm
        self.bvm = np.ascontiguousarray(np.asarray(batch_variable_map, dtype=np.int32))
        self.bfm = np.ascontiguousarray(np.asarray(batch_function_map, dtype=np.int32))
        self.V = int(self.bvm.shape[0])
        self.F = int(self.bfm.shape[0])
        if batch_size is None:
            batch_size = int(self.bvm.max()) + 1 if self.V > 0 else 0   # solver.py:53
        self.B = int(batch_size)
        sign = _f32(np.asarray(edge_feature).reshape(-1))
        self._h = L.ora_create(self.E, self.V, self.F, self.B, _p(gm), _p(sign), _p(self.bvm), _p(self.bfm),
                               1 if strict else 0)
        self._L = L

    def __del__(self):
        try:
            self._L.ora_destroy(self._h)
        except Exception:
            pass

    # ---- state ---------------------------------------------------------------------------
    def set_state(self, prop_state, dec_state):
        pq, pf = _f32(prop_state[0]), _f32(prop_state[1])
        dq, df = _f32(dec_state[0]), _f32(dec_state[1])
        self._L.ora_set_state(self._h, _p(pq), _p(pf), _p(dq), _p(df))

    def set_masks(self, av=None, af=None, sol=None):
        av, af, sol = _f32(av), _f32(af), _f32(sol)
        self._L.ora_set_masks(self._h, _p(av), _p(af), _p(sol))

    def state(self):
        q = np.empty((self.E, 3), np.float32)
        f = np.empty((self.E, 2), np.float32)
        self._L.ora_get_state(self._h, _p(q), _p(f))
        return q, f

    def masks(self):
        av = np.empty(self.V, np.float32)
        af = np.empty(self.F, np.float32)
        sol = np.empty(self.V, np.float32)
        is_sat = np.empty(self.B, np.float32)
        active = np.empty(self.B, np.uint8)
        em = np.empty(self.E, np.float32)
        self._L.ora_get_masks(self._h, _p(av), _p(af), _p(sol), _p(is_sat), _p(active), _p(em))
        return dict(av=av, af=af, sol=sol, is_sat=is_sat, active=active, em=em,
                    em_set=bool(self._L.ora_edge_mask_set(self._h)))

    def counters(self):
        c = np.empty(self.B, np.float32)
        self._L.ora_get_counters(self._h, _p(c))
        return c

    def trace(self):
        n = self._L.ora_trace_len(self._h)
        t = np.empty((n, 3), np.int64)
        if n:
            self._L.ora_get_trace(self._h, _p(t))
        return t

    # ---- SATProblem ----------------------------------------------------------------------
    def simplify(self):
        self._L.ora_simplify(self._h)

    def set_variables(self, assignment):
        a = _f32(assignment).reshape(-1)
        self._L.ora_set_variables(self._h, _p(a))

    def compute_edge_mask(self):
        self._L.ora_compute_edge_mask(self._h)

    # ---- stateless operators -------------------------------------------------------------
    def sp_step(self, dec_q3, dec_fs2, edge_mask=None, prop_q3=None, prop_fs2=None, active=None, pi=0.0):
        dq, df = _f32(dec_q3), _f32(dec_fs2)
        pq = dq if prop_q3 is None else _f32(prop_q3)
        pf = df if prop_fs2 is None else _f32(prop_fs2)
        em = None if edge_mask is None else _f32(np.asarray(edge_mask).reshape(-1))
        act = None if active is None else np.ascontiguousarray(np.asarray(active, dtype=np.uint8).reshape(-1))
        oq = np.empty((self.E, 3), np.float32)
        of = np.empty((self.E, 2), np.float32)
        self._L.ora_sp_step(self._h, _p(dq), _p(df), _p(em), _p(pq), _p(pf), _p(act), pi, _p(oq), _p(of))
        return oq, of

    def score(self, fs2, af, pi=0.0):
        f, a = _f32(fs2), _f32(np.asarray(af).reshape(-1))
        s = np.empty(self.V, np.float32)
        self._L.ora_score(self._h, _p(f), _p(a), pi, _p(s))
        return s

    def cnf_eval(self, pred):
        p = _f32(np.asarray(pred).reshape(-1))
        solved = np.empty(self.B, np.float32)
        nun = np.empty(self.B, np.float32)
        self._L.ora_cnf_eval(self._h, _p(p), _p(solved), _p(nun))
        return solved, nun

    def energy(self, assignment, av, af):
        a, v, f = _f32(np.asarray(assignment).reshape(-1)), _f32(np.asarray(av).reshape(-1)), _f32(np.asarray(af).reshape(-1))
        en = np.empty(self.B, np.float32)
        uf = np.empty(self.F, np.float32)
        self._L.ora_energy(self._h, _p(a), _p(v), _p(f), _p(en), _p(uf))
        return en, uf

    def energy_diff(self, assignment, av, edge_mask):
        a, v, em = _f32(np.asarray(assignment).reshape(-1)), _f32(np.asarray(av).reshape(-1)), _f32(np.asarray(edge_mask).reshape(-1))
        d = np.empty(self.V, np.float32)
        self._L.ora_energy_diff(self._h, _p(a), _p(v), _p(em), _p(d))
        return d

    # ---- the loop ------------------------------------------------------------------------
    def iterate(self, tol, t_max, check_termination=True, batch_replication=1, pi=0.0):
        return int(self._L.ora_iterate(self._h, tol, t_max, 1 if check_termination else 0, batch_replication, pi))

    def run(self, T, tol, t_max, check_termination=True, batch_replication=1, pi=0.0):
        return int(self._L.ora_run(self._h, T, tol, t_max, 1 if check_termination else 0, batch_replication, pi))

    def count_active_variables(self):
        return int(self._L.ora_count_active_variables(self._h))

    def random_fill(self, draws):
        d = _f32(np.asarray(draws).reshape(-1))
        assert d.shape[0] >= self.count_active_variables()
        self._L.ora_random_fill(self._h, _p(d))

    def local_search(self, W, epsilon, rand_var, rand_coin, batch_replication=1):
        rv = _f32(rand_var).reshape(-1) if W > 0 else np.zeros(1, np.float32)
        rc = _f32(rand_coin).reshape(-1) if W > 0 else np.zeros(1, np.float32)
        pred = np.empty(self.V, np.float32)
        it = self._L.ora_local_search(self._h, W, epsilon, batch_replication, _p(rv), _p(rc), _p(pred))
        return pred, int(it)

    def deduplicate(self, batch_replication, prediction):
        p = _f32(np.asarray(prediction).reshape(-1))
        out = np.empty(self.V // batch_replication, np.float32)
        win = np.empty(self.B // batch_replication, np.int64)
        self._L.ora_deduplicate(self._h, batch_replication, _p(p), _p(out), _p(win))
        return out, win


def set_math_mode(correctly_rounded):
    """True: fp32 log/exp rounded from fp64 (matches the CUDA strict-math test build bit for bit)."""
    lib().ora_set_math_mode(1 if correctly_rounded else 0)


def num_threads():
    return int(lib().ora_num_threads())


def set_num_threads(n):
    """pins the OpenMP team size: torchrun exports OMP_NUM_THREADS=1 into every rank"""
    L = lib()
    L.ora_set_num_threads.argtypes = [ctypes.c_int]
    L.ora_set_num_threads(int(n))


def init_state(E, randomized=False, rng=None):
    """(propagator_state, decimator_state) as SurveyPropagatorSolver.get_init_state builds them
    (reference pdp_propagate.py:223-237 and pdp_predict.py:194-208); the random variant takes a numpy
    Generator -- parity tests inject the tensors instead of sharing an RNG stream."""
    if not randomized:
        q = np.full((E, 3), 1.0 / 3.0, np.float32)
        f = np.full((E, 2), 0.5, np.float32)
        f[:, 1] = 0
        return (q.copy(), f.copy()), (q.copy(), f.copy())
    pq = rng.random((E, 3), dtype=np.float32)
    pq = pq / pq.sum(1, keepdims=True)
    pf = rng.random((E, 2), dtype=np.float32)
    pf[:, 1] = 0
    dq = rng.random((E, 3), dtype=np.float32)
    df = rng.random((E, 2), dtype=np.float32)
    df[:, 1] = 0
    return (pq, pf), (dq, df)


# ---- model type `reinforce` ------------------------------------------------------------------------
def _smooth_max_by_variable(o, x):
    "sparse_smooth_max (reference util.py:282-286): fp32 sums in ascending edge order per variable"
    x = x.astype(np.float32)
    w = np.exp(np.minimum(np.float32(30.0) * x, np.float32(30.0))).astype(np.float32)
    num = np.zeros(o.V, np.float32)
    den = np.zeros(o.V, np.float32)
    np.add.at(num, o.gm[0], (x * w).astype(np.float32))
    np.add.at(den, o.gm[0], w)
    return (num / np.maximum(den, np.float32(1.0))).astype(np.float32)


def _sparse_max(o, x):
    "sparse_max (reference util.py:267-275): batch-global min, zero floor, fl(fl(max + min) - 1)"
    x = x.astype(np.float32)
    mn = np.float32(x.min())
    mx = np.zeros(o.B, np.float32)
    np.maximum.at(mx, o.bvm, ((x - mn).astype(np.float32) + np.float32(1.0)).astype(np.float32))
    return ((mx + mn).astype(np.float32) - np.float32(1.0)).astype(np.float32)


def reinforce_run(o, init_prop, init_dec, T, pi, p_dec, coins):
    """The reference's loop (solver.py:355-386) for ReinforceSurveyPropagatorSolver on the oracle's C operators:
    SurveyPropagator with pi (pdp_propagate.py:139-221), ReinforceDecimator (pdp_decimate.py:204-233),
    ReinforcePredictor (pdp_predict.py:221-226), _update_solution (solver.py:388-399) and the trainer's termination
    (trainer.py:150-162).  Returns (merged prediction per iteration, final q [E,3], final [eta, force] [E,2])."""
    o.simplify()
    o.compute_edge_mask()
    m = o.masks()
    av, af, sol, em = m["av"], m["af"], m["sol"].copy(), m["em"]
    var = o.gm[0]
    active = np.ones(o.B, np.uint8)
    ps = (np.asarray(init_prop[0], np.float32), np.asarray(init_prop[1], np.float32))
    ds = (np.asarray(init_dec[0], np.float32), np.asarray(init_dec[1], np.float32))
    prev, em_attr, merged_all = None, None, []
    for t in range(int(T)):
        q, f = o.sp_step(ds[0], ds[1], ds[2] if len(ds) == 3 else None, ps[0], ps[1], active, pi)
        eta = f[:, 0]
        if prev is not None and av.sum() > 0:
            diff = np.abs(prev - eta)
            if em_attr is not None:
                diff = diff * em_attr
            sd = _sparse_max(o, _smooth_max_by_variable(o, diff) * av)
            active[sd <= 0.01] = 0
        prev = eta.copy()
        if coins[t] < p_dec:
            mask = active[o.bvm[var]].astype(np.float32)
            score = o.score(f, af, pi)
            f[:, 1] = mask * np.sign(score)[var] + (1 - mask) * f[:, 1]
        ps = (q, f)
        em_attr = em
        ds = (q, f, em) if em.sum() < o.E else (q, f)
        force = np.zeros(o.V, np.float32)
        np.add.at(force, var, f[:, 1])
        merged = av * (force > 0).astype(np.float32) + (1 - av) * sol
        sol[av == 1] = merged[av == 1]
        merged_all.append(merged.copy())
        solved, _ = o.cnf_eval(merged)
        active[(active == 1) & (solved > 0.5)] = 0
        if active.sum() == 0:
            break
    return merged_all, ps[0], ps[1]
