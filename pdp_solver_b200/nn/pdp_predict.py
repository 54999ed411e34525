"""Predictors and scorers of the PDP framework, B200-native (reference src/pdp/nn/pdp_predict.py)."""
import torch
import torch.nn as nn

from . import util


class NeuralPredictor(nn.Module):
    """The neural predictor (reference pdp_predict.py:20-104): a self-inclusive deep-set aggregator over each
    variable's edges followed by the classifier.  Same sub-module names (state-dict compatible)."""

    def __init__(self, device, decimator_dimension, prediction_dimension, edge_dimension, meta_data_dimension,
                 mem_hidden_dimension, agg_hidden_dimension, mem_agg_hidden_dimension, variable_classifier=None,
                 function_classifier=None):
        super(NeuralPredictor, self).__init__()
        self._device = device
        self._module_list = nn.ModuleList()
        self._variable_classifier = variable_classifier
        self._function_classifier = function_classifier
        self._hidden_dimension = decimator_dimension
        if variable_classifier is not None:
            self._variable_aggregator = util.MessageAggregator(
                device, decimator_dimension + edge_dimension + meta_data_dimension, decimator_dimension, mem_hidden_dimension,
                mem_agg_hidden_dimension, agg_hidden_dimension, 0, include_self_message=True)
            self._module_list.append(self._variable_aggregator)
            self._module_list.append(self._variable_classifier)
        if function_classifier is not None:
            self._function_aggregator = util.MessageAggregator(
                device, decimator_dimension + edge_dimension + meta_data_dimension, decimator_dimension, mem_hidden_dimension,
                mem_agg_hidden_dimension, agg_hidden_dimension, 0, include_self_message=True)
            self._module_list.append(self._function_aggregator)
            self._module_list.append(self._function_classifier)

    def forward(self, decimator_state, sat_problem, last_call=False):
        if sat_problem._meta_data is not None:
            raise NotImplementedError("meta_data features are not supported")
        ctx = sat_problem._ctx
        if len(decimator_state) == 3:
            dvs, dfs, edge_mask = decimator_state
        else:
            (dvs, dfs), edge_mask = decimator_state, None
        ef = sat_problem._edge_feature
        variable_prediction = function_prediction = None
        if self._variable_classifier is not None:
            agg = self._variable_aggregator((dvs, ef), None, ctx, True, edge_mask)
            variable_prediction = self._variable_classifier(agg)
        if self._function_classifier is not None:
            agg = self._function_aggregator((dfs, ef), None, ctx, False, edge_mask)
            function_prediction = self._function_classifier(agg)
        return variable_prediction, function_prediction

    def get_init_state(self, graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat,
                       randomized, batch_replication):
        "reference pdp_predict.py:93-104"
        edge_num = graph_map.size(1) * batch_replication
        if randomized:
            variable_state = 2.0 * torch.rand(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device) - 1.0
            function_state = 2.0 * torch.rand(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device) - 1.0
        else:
            variable_state = torch.zeros(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device)
            function_state = torch.zeros(edge_num, self._hidden_dimension, dtype=torch.float32, device=self._device)
        return (variable_state, function_state)


class IdentityPredictor(nn.Module):
    """Prediction = the SATProblem's running solution; on the last call undecided variables are filled
    with uniform draws consumed in batch-global variable order (reference pdp_predict.py:110-128)."""

    def __init__(self, device, random_fill=False):
        super(IdentityPredictor, self).__init__()
        self._random_fill = random_fill
        self._device = device

    def forward(self, decimator_state, sat_problem, last_call=False):
        ctx = sat_problem._ctx
        if self._random_fill and last_call:
            n_active = ctx.count_active_variables()
            if n_active > 0:
                ctx.random_fill(torch.rand(n_active, device=ctx.device))   # same draw as pdp_predict.py:126
        return ctx.solution().unsqueeze(1), None


class SurveyScorer(nn.Module):
    "Per-variable SP bias W(+) - W(-) used for SP-guided decimation (reference pdp_predict.py:134-208)."

    def __init__(self, device, message_dimension, include_adaptors=False, pi=0.0):
        super(SurveyScorer, self).__init__()
        if include_adaptors:
            raise NotImplementedError("neural adaptors are not part of the accelerated path yet")
        self._device = device
        self._include_adaptors = include_adaptors
        self._pi_value = float(pi)

    def forward(self, message_state, sat_problem, last_call=False):
        ctx = sat_problem._ctx
        af = ctx.get_masks()["af"]
        return ctx.score(message_state[1], af, self._pi_value).unsqueeze(1), None

    def get_init_state(self, graph_map, batch_variable_map, batch_function_map, edge_feature, graph_feat,
                       randomized, batch_replication):
        "reference pdp_predict.py:194-208 (the random variant is NOT normalised, as in the reference)"
        edge_num = graph_map.size(1) * batch_replication
        if randomized:
            variable_state = torch.rand(edge_num, 3, dtype=torch.float32, device=self._device)
            function_state = torch.rand(edge_num, 2, dtype=torch.float32, device=self._device)
            function_state[:, 1] = 0
        else:
            variable_state = torch.full((edge_num, 3), 1.0 / 3.0, dtype=torch.float32, device=self._device)
            function_state = torch.zeros(edge_num, 2, dtype=torch.float32, device=self._device)
            function_state[:, 0] = 0.5
            # tag: every row is the same (checked against in-place edits through the version counters), so
            # the solver can fill its message arrays without gathering 20 bytes per edge
            variable_state._pdp_const = ((1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0), variable_state._version)
            function_state._pdp_const = ((0.5, 0.0), function_state._version)
        return (variable_state, function_state)


class ReinforcePredictor(nn.Module):
    "Prediction of the Reinforce algorithm: a variable is True iff its external forces sum to > 0 (reference pdp_predict.py:214-226)."

    def __init__(self, device):
        super(ReinforcePredictor, self).__init__()
        self._device = device

    def forward(self, decimator_state, sat_problem, last_call=False):
        s, _ = sat_problem._ctx.edge_aggregate(decimator_state[1][:, 1:2].contiguous(), by_variable=True)
        return (s > 0).float(), None
