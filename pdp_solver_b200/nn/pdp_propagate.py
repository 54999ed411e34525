"""Propagators of the PDP framework, B200-native (reference src/pdp/nn/pdp_propagate.py)."""
import torch
import torch.nn as nn


class SurveyPropagator(nn.Module):
    """Survey Propagation as a PDP propagator (reference pdp_propagate.py:114-237), adaptors off.

    `forward` is one SP flooding sweep -- both message directions computed from the previous messages
    -- executed by the library's pdp_sp_step kernels on the batch's CSR/CSC; inside solver.forward the
    sweep runs fused in the persistent loop instead."""

    def __init__(self, device, decimator_dimension, include_adaptors=False, pi=0.0):
        super(SurveyPropagator, self).__init__()
        if include_adaptors:
            raise NotImplementedError("neural adaptors (p-nd-np) are not part of the accelerated path yet")
        self._device = device
        self._function_message_dim = 3
        self._variable_message_dim = 2
        self._include_adaptors = include_adaptors
        self._pi = torch.tensor([pi], dtype=torch.float32, device=device)
        self._pi_value = float(pi)

    def pi_value(self):
        return self._pi_value

    def forward(self, init_state, decimator_state, sat_problem, is_training, active_mask=None):
        if len(decimator_state) == 3:
            dq, df, edge_mask = decimator_state
        else:
            (dq, df), edge_mask = decimator_state, None
        return sat_problem._ctx.sp_step(dq, df, edge_mask, init_state[0], init_state[1], active_mask, self._pi_value)

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
