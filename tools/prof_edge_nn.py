#!/usr/bin/env python3
"""Profiling driver of the tensor-core edge layers: one GRU call and one MLP call on E rows (default 1.2 M)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdp_solver_b200.nn import tensor_ops as T
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1200000
dev = torch.device("cuda:0")
torch.manual_seed(0)
cell = torch.nn.GRUCell(151, 150).to(dev)
lin = torch.nn.Linear(151, 100).to(dev)
x1, x2, h = torch.randn(E, 150, device=dev), torch.sign(torch.randn(E, 1, device=dev)), torch.rand(E, 150, device=dev)
tg, tl = T.TensorGRU(cell), T.TensorLinear(lin)
for rep in range(3):
    e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e0.record(); tg([x1, x2], h); e1.record(); tl([x1, x2], act=1); e2.record(); torch.cuda.synchronize()
    print("E=%d  gru %.3f ms (%.1f TFLOP/s fp32-equivalent)   mlp 151->100 %.3f ms (%.1f TFLOP/s)" % (
        E, e0.elapsed_time(e1), 2.0 * E * 301 * 450 / e0.elapsed_time(e1) / 1e9, e1.elapsed_time(e2), 2.0 * E * 151 * 100 / e1.elapsed_time(e2) / 1e9))
for f, name in ((lambda: cell(torch.cat((x1, x2), 1), h), "torch GRUCell fp32"), (lambda: torch.nn.functional.logsigmoid(lin(torch.cat((x1, x2), 1))), "torch linear+logsigmoid")):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    print("   %s: %.3f ms" % (name, e0.elapsed_time(e1)))
