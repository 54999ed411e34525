"""Propagators of the PDP framework, B200-native (reference src/pdp/nn/pdp_propagate.py)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import tensor_ops, util


def edge_problem_mask(sat_problem, active_mask):
    """[E,1] float: 1 on the edges of active problems (reference: mm(variable_mask_transpose, mm(b_variable_mask,
    active)), pdp_propagate.py:147-148) -- a gather through the two index maps"""
    if active_mask is None:
        return None
    prob = sat_problem.edge_problem_index()
    return active_mask.reshape(-1)[prob].to(torch.float32).unsqueeze(1)


def _blend(mask, new, old):
    """mask * new + (1 - mask) * old for a 0/1 row mask [E,1] (frozen problems keep their state, reference
    pdp_propagate.py:75,87) in one elementwise kernel: lerp returns `new` where mask = 1 and `old` where mask = 0 exactly."""
    return torch.lerp(old, new, mask)


class NeuralMessagePasser(nn.Module):
    """The neural propagator of `np-nd-np` (reference pdp_propagate.py:17-108): two deep-set aggregators, one per
    message direction.  Dense layers are library GEMMs; the segmented sums are the library's kernels."""

    def __init__(self, device, edge_dimension, decimator_dimension, meta_data_dimension, hidden_dimension, mem_hidden_dimension,
                 mem_agg_hidden_dimension, agg_hidden_dimension, dropout):
        super(NeuralMessagePasser, self).__init__()
        self._device = device
        self._module_list = nn.ModuleList()
        self._drop_out = dropout
        self._variable_aggregator = util.MessageAggregator(
            device, decimator_dimension + edge_dimension + meta_data_dimension, hidden_dimension, mem_hidden_dimension,
            mem_agg_hidden_dimension, agg_hidden_dimension, edge_dimension, include_self_message=False)
        self._function_aggregator = util.MessageAggregator(
            device, decimator_dimension + edge_dimension + meta_data_dimension, hidden_dimension, mem_hidden_dimension,
            mem_agg_hidden_dimension, agg_hidden_dimension, edge_dimension, include_self_message=False)
        self._module_list.append(self._variable_aggregator)
        self._module_list.append(self._function_aggregator)
        self._hidden_dimension = hidden_dimension
        self._mem_hidden_dimension = mem_hidden_dimension
        self._agg_hidden_dimension = agg_hidden_dimension
        self._mem_agg_hidden_dimension = mem_agg_hidden_dimension

    def forward(self, init_state, decimator_state, sat_problem, is_training, active_mask=None):
        if sat_problem._meta_data is not None:
            raise NotImplementedError("meta_data features are not supported")
        ctx = sat_problem._ctx
        mask = edge_problem_mask(sat_problem, active_mask)
        if len(decimator_state) == 3:
            dvs, dfs, edge_mask = decimator_state
        else:
            (dvs, dfs), edge_mask = decimator_state, None
        variable_state, function_state = init_state
        ef = sat_problem._edge_feature
        # variables --> functions (reference :69-78)
        new_f = self._variable_aggregator((dvs, ef), ef, ctx, True, edge_mask)
        function_state = new_f if mask is None else _blend(mask, new_f, function_state)
        # functions --> variables (reference :80-89)
        new_v = self._function_aggregator((dfs, ef), ef, ctx, False, edge_mask)
        variable_state = new_v if mask is None else _blend(mask, new_v, variable_state)
        return variable_state, function_state

    def get_init_state(self, graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat,
                       randomized, batch_replication):
        "reference pdp_propagate.py:97-108"
        edge_num = graph_map.size(1) * batch_replication
        if randomized:
            variable_state = 2.0 * torch.rand(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device) - 1.0
            function_state = 2.0 * torch.rand(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device) - 1.0
        else:
            variable_state = torch.zeros(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device)
            function_state = torch.zeros(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device)
        return (variable_state, function_state)


class SurveyPropagator(nn.Module):
    """Survey Propagation as a PDP propagator (reference pdp_propagate.py:114-237).

    `forward` is one SP flooding sweep -- both message directions computed from the previous messages
    -- executed by the library's pdp_sp_step kernels on the batch's CSR/CSC; inside the p-d-p solver the
    sweep runs fused in the persistent loop instead.  With `include_adaptors` (model type p-nd-np) the two
    Linear adaptors (150 -> 1, 150 -> 2; library GEMMs) turn the decimator's hidden states into messages and
    pdp_sp_step_adapted does the message arithmetic."""

    def __init__(self, device, decimator_dimension, include_adaptors=False, pi=0.0):
        super(SurveyPropagator, self).__init__()
        self._device = device
        self._function_message_dim = 3
        self._variable_message_dim = 2
        self._include_adaptors = include_adaptors
        self._pi = torch.tensor([pi], dtype=torch.float32, device=device)
        self._pi_value = float(pi)
        if self._include_adaptors:
            self._variable_input_projector = nn.Linear(decimator_dimension, self._variable_message_dim, bias=False)
            self._function_input_projector = nn.Linear(decimator_dimension, 1, bias=False)
            self._module_list = nn.ModuleList([self._variable_input_projector, self._function_input_projector])

    def pi_value(self):
        return self._pi_value

    def forward(self, init_state, decimator_state, sat_problem, is_training, active_mask=None):
        if len(decimator_state) == 3:
            dq, df, edge_mask = decimator_state
        else:
            (dq, df), edge_mask = decimator_state, None
        ctx = sat_problem._ctx
        if not self._include_adaptors:
            return ctx.sp_step(dq, df, edge_mask, init_state[0], init_state[1], active_mask, self._pi_value)
        if tensor_ops.use_tensor_cores():
            tc = self.__dict__.setdefault("_tc_layers", {})
            if not tc:
                tc["f"] = tensor_ops.TensorLinear(self._function_input_projector)
                tc["v"] = tensor_ops.TensorLinear(self._variable_input_projector)
            x_log = tc["f"]([dq], act=tensor_ops.ACT_LOGSIGMOID)           # reference :163-164
            proj = tc["v"]([df])                                            # reference :179-182
        else:
            x_log = F.logsigmoid(self._function_input_projector(dq))
            proj = self._variable_input_projector(df)
        eta_in = torch.sigmoid(proj[:, 0])
        ext = torch.sign(proj[:, 1])
        return ctx.sp_step_adapted(x_log, eta_in, ext, edge_mask, init_state[0], init_state[1], active_mask, self._pi_value)

    def get_init_state(self, graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat,
                       randomized, batch_replication):
        "reference pdp_propagate.py:223-237 (same torch.rand call order and shapes)"
        edge_num = graph_map.size(1) * batch_replication
        if randomized:
            variable_state = torch.rand(edge_num, self._function_message_dim, dtype=torch.float32, device=self._device)
            variable_state = variable_state / torch.sum(variable_state, 1).unsqueeze(1)
            function_state = torch.rand(edge_num, self._variable_message_dim, dtype=torch.float32, device=self._device)
            function_state[:, 1] = 0
        else:
            variable_state = torch.ones(edge_num, self._function_message_dim, dtype=torch.float32,
                                        device=self._device) / self._function_message_dim
            function_state = 0.5 * torch.ones(edge_num, self._variable_message_dim, dtype=torch.float32, device=self._device)
            function_state[:, 1] = 0
        return (variable_state, function_state)
