"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import-and-patch harness for the *unmodified* reference (microsoft/PDP-Solver) so it
can be run on CPU in the authoring container.  The reference tree is read-only and
bit-rotted (SURVEY.md section 5 "bit-rot list"); nothing is copied, the live classes
are monkey-patched in place:

  S1  pdp/trainer.py:150-162   self-aliased masked write in _check_recurrence_termination
  S2  pdp/nn/solver.py:420,424 float division inside view() when batch_replication > 1
  S3  pdp/nn/solver.py:555     p-nd-np builds GRUCell(1+1, H) but SP returns 2 columns

The reference is looked up at /root/reference (authoring container) and at baseline/_ref (the pip
--target install of the unmodified reference, which travels to the GPU box).  Used by
`oracle/make_golden.py` (fixture generation), by CPU-side oracle validation tests that skip when the
tree is absent, and by `bench.py --impl reference` / the `cpu_baseline` leg (timing the reference's
own `--cpu_mode` forward).
"""
import os
import sys
import warnings

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference():
    """directory holding the reference's `pdp` package: $PDP_REFERENCE_ROOT/src, /root/reference/src (authoring
    container) or baseline/_ref (the unmodified reference installed with `pip install --no-deps --target baseline/_ref`,
    git-ignored; it travels to the GPU box with the snapshot)"""
    cands = []
    if os.environ.get("PDP_REFERENCE_ROOT"):
        cands += [os.path.join(os.environ["PDP_REFERENCE_ROOT"], "src"), os.environ["PDP_REFERENCE_ROOT"]]
    cands += ["/root/reference/src", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if os.path.isdir(os.path.join(c, "pdp", "nn")):
            return c
    return cands[-1]


REF_SRC = _find_reference()


def reference_available():
    return os.path.isdir(os.path.join(REF_SRC, "pdp", "nn"))


_patched = False


def load_reference():
    """Returns the reference's (solver, pdp_propagate, pdp_decimate, pdp_predict, util, trainer)
    modules with the shims applied."""
    global _patched
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_SRC)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    warnings.filterwarnings("ignore")

    import torch
    from pdp.nn import solver, pdp_propagate, pdp_decimate, pdp_predict, util
    from pdp import trainer

    if not _patched:
        # ---- S1: trainer.py:150-162 -------------------------------------------------
        def _check_recurrence_termination(self, active, prediction, sat_problem):
            output, _ = self._cnf_evaluator(
                variable_prediction=prediction[0], graph_map=sat_problem._graph_map,
                batch_variable_map=sat_problem._batch_variable_map,
                batch_function_map=sat_problem._batch_function_map,
                edge_feature=sat_problem._edge_feature, meta_data=sat_problem._meta_data)
            idx = active[:, 0].clone().bool()
            if sat_problem._batch_replication > 1:
                real_batch = torch.mm(sat_problem._replication_mask_tuple[1], (output > 0.5).float())
                dup_batch = torch.mm(sat_problem._replication_mask_tuple[0], (real_batch == 0).float())
                active[idx, 0] = (dup_batch[idx, 0] > 0).to(active.dtype)
            else:
                active[idx, 0] = (output[idx, 0] <= 0.5).to(active.dtype)

        trainer.SatFactorGraphTrainer._check_recurrence_termination = _check_recurrence_termination

        # ---- S2: solver.py:420,424 ---------------------------------------------------
        class _IntDiv(int):
            def __truediv__(self, other):
                return int(self) // int(other)

        _orig_setup = solver.SATProblem.setup_problem

        def _setup_problem(self, data_batch, batch_replication):
            _orig_setup(self, data_batch, batch_replication)
            self._edge_num = _IntDiv(self._edge_num)

        solver.SATProblem.setup_problem = _setup_problem

        # ---- S3: solver.py:555 -------------------------------------------------------
        _orig_nd_init = pdp_decimate.NeuralDecimator.__init__

        def _nd_init(self, device, message_dimension, *args, **kwargs):
            if message_dimension == (3, 1):
                message_dimension = (3, 2)
            _orig_nd_init(self, device, message_dimension, *args, **kwargs)

        pdp_decimate.NeuralDecimator.__init__ = _nd_init
        _patched = True

    return solver, pdp_propagate, pdp_decimate, pdp_predict, util, trainer


class _NullLogger(object):
    def info(self, *a, **k):
        pass


def make_termination_callback(device):
    """The trainer's per-iteration termination check (trainer.py:150-162, shim S1) bound to a
    bare SatCNFEvaluator, without constructing the whole trainer."""
    _, _, _, _, util, trainer = load_reference()

    class _Holder(object):
        pass

    h = _Holder()
    h._cnf_evaluator = util.SatCNFEvaluator(device=device)
    fn = trainer.SatFactorGraphTrainer._check_recurrence_termination
    return lambda active, prediction, sat_problem: fn(h, active, prediction, sat_problem)
