"""Utilities of the PDP framework, B200-native (reference src/pdp/nn/util.py)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..engine import Context
from . import tensor_ops


class MessageAggregator(nn.Module):
    """Deep-set message aggregation at variable / function nodes (reference util.py:13-77), same sub-module
    names and shapes (state-dict compatible).  The dense layers are library GEMMs (cuBLAS through torch, fp32 like
    the reference); the segmented sum over a node's edges and the leave-one-out gather are the library's
    pdp_edge_aggregate kernel on the batch's CSR/CSC (ascending edge order) instead of two sparse mm."""

    def __init__(self, device, input_dimension, output_dimension, mem_hidden_dimension,
                 mem_agg_hidden_dimension, agg_hidden_dimension, feature_dimension, include_self_message):
        super(MessageAggregator, self).__init__()
        self._device = device
        self._include_self_message = include_self_message
        self._module_list = nn.ModuleList()
        if mem_hidden_dimension > 0 and mem_agg_hidden_dimension > 0:
            self._W1_m = nn.Linear(input_dimension, mem_hidden_dimension, bias=True)
            self._W2_m = nn.Linear(mem_hidden_dimension, mem_agg_hidden_dimension, bias=False)
            self._module_list.append(self._W1_m)
            self._module_list.append(self._W2_m)
        if agg_hidden_dimension > 0 and mem_agg_hidden_dimension > 0:
            if mem_hidden_dimension <= 0:
                mem_agg_hidden_dimension = input_dimension
            self._W1_a = nn.Linear(mem_agg_hidden_dimension + feature_dimension, agg_hidden_dimension, bias=True)
            self._W2_a = nn.Linear(agg_hidden_dimension, output_dimension, bias=False)
            self._module_list.append(self._W1_a)
            self._module_list.append(self._W2_a)
        self._agg_hidden_dimension = agg_hidden_dimension
        self._mem_hidden_dimension = mem_hidden_dimension
        self._mem_agg_hidden_dimension = mem_agg_hidden_dimension

    def forward(self, state, feature, ctx, by_variable, edge_mask=None):
        """`ctx`, `by_variable` replace the reference's (mask, mask_transpose) sparse matrices: the node side
        the edges are summed on (reference util.py:51-77).  `state` is the layer input or the list of row-major pieces
        whose concatenation it is (the tensor-core layers read the pieces in place: no torch.cat)."""
        pieces = list(state) if isinstance(state, (list, tuple)) else [state]
        tc = tensor_ops.use_tensor_cores()
        pre = self._mem_hidden_dimension > 0 and self._mem_agg_hidden_dimension > 0
        if pre and tc:
            # W1 -> logsigmoid -> W2 -> logsigmoid (-> * edge_mask) on the tensor cores (csrc/pdp_edge_nn.cu)
            hidden = self._tc("_W1_m")(pieces, act=tensor_ops.ACT_LOGSIGMOID)
            state = self._tc("_W2_m")([hidden], act=tensor_ops.ACT_LOGSIGMOID, row_mask=edge_mask)
        else:
            state = pieces[0] if len(pieces) == 1 else torch.cat(pieces, 1)
            if pre:
                state = F.logsigmoid(self._W2_m(F.logsigmoid(self._W1_m(state))))
            if edge_mask is not None:
                state = state * edge_mask
        node_sum, loo = ctx.edge_aggregate(state, by_variable, leave_one_out=not self._include_self_message)
        aggregated_state = node_sum if self._include_self_message else loo
        post = self._agg_hidden_dimension > 0 and self._mem_agg_hidden_dimension > 0
        if post and tc:
            srcs = [aggregated_state] + ([feature] if feature is not None else [])
            hidden = self._tc("_W1_a")(srcs, act=tensor_ops.ACT_LOGSIGMOID)
            return self._tc("_W2_a")([hidden], act=tensor_ops.ACT_LOGSIGMOID)
        if feature is not None:
            aggregated_state = torch.cat((aggregated_state, feature), 1)
        if post:
            aggregated_state = F.logsigmoid(self._W2_a(F.logsigmoid(self._W1_a(aggregated_state))))
        return aggregated_state

    def _tc(self, name):
        "the tensor-core image of one of the module's nn.Linear layers (rebuilt when its parameters change)"
        cache = self.__dict__.setdefault("_tc_layers", {})
        if name not in cache:
            cache[name] = tensor_ops.TensorLinear(getattr(self, name))
        return cache[name]


class Perceptron(nn.Module):
    "The 1-hidden-layer classifier the trainer builds (reference trainer.py:20-29)."

    def __init__(self, input_dimension, hidden_dimension, output_dimension):
        super(Perceptron, self).__init__()
        self._layer1 = nn.Linear(input_dimension, hidden_dimension)
        self._layer2 = nn.Linear(hidden_dimension, output_dimension, bias=False)

    def forward(self, inp):
        return torch.sigmoid(self._layer2(F.relu(self._layer1(inp))))


class PerceptronTanh(nn.Module):
    "1-hidden-layer perceptron with tanh output, the np-d-np scorer's classifier (reference util.py:240-250)."

    def __init__(self, input_dimension, hidden_dimension, output_dimension):
        super(PerceptronTanh, self).__init__()
        self._layer1 = nn.Linear(input_dimension, hidden_dimension)
        self._layer2 = nn.Linear(hidden_dimension, output_dimension, bias=False)

    def forward(self, inp):
        return torch.tanh(self._layer2(torch.relu(self._layer1(inp))))


class SatCNFEvaluator(nn.Module):
    """Verdict and number of unsatisfied clauses per problem for a variable prediction
    (reference util.py:203-236).  Integer-exact: a literal is true iff s*p + (1-s)/2 > 0.5 in fp32.
    Stateless: one pass over the caller's edge list (`pdp_cnf_eval_edges`), no solver context, nothing cached."""

    def __init__(self, device):
        super(SatCNFEvaluator, self).__init__()
        self._device = device

    def forward(self, variable_prediction, graph_map, batch_variable_map, batch_function_map, edge_feature, meta_data):
        return cnf_eval_edges(variable_prediction, graph_map, batch_variable_map, batch_function_map, edge_feature)


def cnf_eval_edges(variable_prediction, graph_map, batch_variable_map, batch_function_map, edge_feature, batch_size=None):
    import ctypes
    from .. import _lib
    dev = graph_map.device
    if dev.type != "cuda":
        raise _lib.PdpError("SatCNFEvaluator expects CUDA tensors (got %s); there is no CPU fallback" % dev)
    L = _lib.load()
    gm = graph_map.to(torch.int32).contiguous()
    bfm = batch_function_map.to(torch.int32).contiguous()
    ef = edge_feature.to(torch.float32).reshape(-1).contiguous()
    pred = variable_prediction.to(torch.float32).reshape(-1).contiguous()
    E, V, F = int(gm.shape[1]), int(batch_variable_map.shape[0]), int(bfm.shape[0])
    if batch_size is None:
        batch_size = int(batch_variable_map.max().item()) + 1 if V > 0 else 0
    B = int(batch_size)
    with torch.cuda.device(dev):
        solved = torch.empty(B, dtype=torch.float32, device=dev)
        n_unsat = torch.empty(B, dtype=torch.float32, device=dev)
        scratch = torch.empty(max(int(L.pdp_cnf_eval_edges_scratch_bytes(F, B)), 4), dtype=torch.uint8, device=dev)
        ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t.numel() else ctypes.c_void_p(0)
        _lib.check(L.pdp_cnf_eval_edges(ptr(gm), ptr(ef), ptr(bfm), E, V, F, B, ptr(pred), ptr(solved), ptr(n_unsat),
                                        ctypes.c_void_p(scratch.data_ptr()),
                                        ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "pdp_cnf_eval_edges", L)
    return solved.unsqueeze(1), n_unsat.unsqueeze(1)
