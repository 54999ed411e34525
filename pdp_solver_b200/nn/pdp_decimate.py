"""Decimators of the PDP framework, B200-native (reference src/pdp/nn/pdp_decimate.py)."""
import torch
import torch.nn as nn

from .pdp_propagate import edge_problem_mask


class NeuralDecimator(nn.Module):
    """The neural (non-greedy) decimator of `p-nd-np` / `np-nd-np` (reference pdp_decimate.py:21-100): one GRU
    cell per message direction over the edges; same sub-module names (state-dict compatible).  The GRU cells are
    library GEMMs + pointwise kernels (torch), fp32 like the reference.

    message_dimension == (3, 1) is widened to (3, 2) as the reference needs to run at all: its SurveyPropagator
    returns a 2-column function state (pdp_propagate.py:221) while solver.py:555 sizes the cell for 1 (SURVEY.md
    section 5, bit-rot item 4)."""

    def __init__(self, device, message_dimension, meta_data_dimension, hidden_dimension, mem_hidden_dimension,
                 mem_agg_hidden_dimension, agg_hidden_dimension, edge_dimension, dropout):
        super(NeuralDecimator, self).__init__()
        self._device = device
        self._module_list = nn.ModuleList()
        self._drop_out = dropout
        if message_dimension == (3, 1):
            message_dimension = (3, 2)
        if isinstance(message_dimension, tuple):
            variable_message_dim, function_message_dim = message_dimension
        else:
            variable_message_dim = function_message_dim = message_dimension
        self._variable_rnn_cell = nn.GRUCell(variable_message_dim + edge_dimension + meta_data_dimension, hidden_dimension, bias=True)
        self._function_rnn_cell = nn.GRUCell(function_message_dim + edge_dimension + meta_data_dimension, hidden_dimension, bias=True)
        self._module_list.append(self._variable_rnn_cell)
        self._module_list.append(self._function_rnn_cell)
        self._hidden_dimension = hidden_dimension
        self._mem_hidden_dimension = mem_hidden_dimension
        self._agg_hidden_dimension = agg_hidden_dimension
        self._mem_agg_hidden_dimension = mem_agg_hidden_dimension

    def forward(self, init_state, message_state, sat_problem, is_training, active_mask=None):
        if sat_problem._meta_data is not None:
            raise NotImplementedError("meta_data features are not supported")
        mask = edge_problem_mask(sat_problem, active_mask)
        variable_state, function_state = message_state[0], message_state[1]
        ef = sat_problem._edge_feature
        new_v = self._variable_rnn_cell(torch.cat((variable_state, ef), 1), init_state[0])
        new_f = self._function_rnn_cell(torch.cat((function_state, ef), 1), init_state[1])
        if mask is not None:   # frozen problems keep their state (reference :75,83)
            new_v = mask * new_v + (1 - mask) * init_state[0]
            new_f = mask * new_f + (1 - mask) * init_state[1]
        return new_v, new_f

    def get_init_state(self, graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat,
                       randomized, batch_replication):
        "reference pdp_decimate.py:89-100"
        edge_num = graph_map.size(1) * batch_replication
        if randomized:
            variable_state = 2.0 * torch.rand(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device) - 1.0
            function_state = 2.0 * torch.rand(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device) - 1.0
        else:
            variable_state = torch.zeros(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device)
            function_state = torch.zeros(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device)
        return (variable_state, function_state)


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
