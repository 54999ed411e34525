"""Utilities of the PDP framework, B200-native (reference src/pdp/nn/util.py)."""
import torch.nn as nn

from ..engine import Context


class SatCNFEvaluator(nn.Module):
    """Verdict and number of unsatisfied clauses per problem for a variable prediction
    (reference util.py:203-236).  Integer-exact: a literal is true iff s*p + (1-s)/2 > 0.5 in fp32."""

    def __init__(self, device):
        super(SatCNFEvaluator, self).__init__()
        self._device = device
        self._cache = None

    def forward(self, variable_prediction, graph_map, batch_variable_map, batch_function_map, edge_feature, meta_data):
        key = (graph_map.data_ptr(), batch_variable_map.data_ptr(), batch_function_map.data_ptr(),
               edge_feature.data_ptr(), tuple(graph_map.shape))
        if self._cache is None or self._cache[0] != key:
            self._cache = (key, Context(graph_map, batch_variable_map, batch_function_map, edge_feature))
        solved, n_unsat = self._cache[1].cnf_eval(variable_prediction)
        return solved.unsqueeze(1), n_unsat.unsqueeze(1)
