"""Decimators of the PDP framework, B200-native (reference src/pdp/nn/pdp_decimate.py)."""
import torch.nn as nn


class SequentialDecimator(nn.Module):
    """The greedy, convergence-gated sequential decimator (reference pdp_decimate.py:106-183): at most
    one variable per problem and iteration is fixed, once that problem's surveys have converged
    (max smooth-max |delta eta| < tolerance) or its counter reached t_max.

    Its statistics, per-problem decisions, scoring, arg-max, variable fixing and the following unit
    propagation / peeling closure all live inside the persistent kernel launched by
    PropagatorDecimatorSolverBase.forward (pdp_sp_run); the module keeps the hyper-parameters and the
    initial-state contract."""

    def __init__(self, device, message_dimension, scorer, tolerance, t_max):
        super(SequentialDecimator, self).__init__()
        self._device = device
        self._tolerance = tolerance
        self._scorer = scorer
        self._message_dimension = message_dimension
        self._t_max = t_max
        self._module_list = nn.ModuleList([self._scorer])

    def forward(self, init_state, message_state, sat_problem, is_training, active_mask=None):
        raise NotImplementedError("SequentialDecimator runs fused inside solver.forward (pdp_sp_run); "
                                  "call the solver, or Context.sp_run(1, ...) for a single iteration")

    def get_init_state(self, graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat,
                       randomized, batch_replication):
        "reference pdp_decimate.py:179-183: module state is reset, the scorer provides the messages"
        return self._scorer.get_init_state(graph_map, batch_variable_map, batch_function_map, edge_feature,
                                           graph_feat, randomized, batch_replication)
