#!/usr/bin/env python3
"""Tensor-core edge layers against torch in fp64 (GPU):  python tools/test_edge_nn.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdp_solver_b200.nn import tensor_ops as T

dev = torch.device("cuda:0")
torch.manual_seed(0)
for (E, k1, k2, n, act) in [(1000, 150, 1, 100, 1), (128, 100, 0, 50, 1), (333, 50, 1, 100, 1), (70000, 100, 0, 150, 1), (513, 150, 0, 2, 0)]:
    lin = torch.nn.Linear(k1 + k2, n, bias=True).to(dev)
    x1 = torch.randn(E, k1, device=dev)
    x2 = torch.sign(torch.randn(E, k2, device=dev)) if k2 else None
    mask = (torch.rand(E, device=dev) > 0.2).float()
    tl = T.TensorLinear(lin)
    src = [x1] + ([x2] if k2 else [])
    out = tl(src, act=act, row_mask=mask)
    torch.cuda.synchronize()
    xin = torch.cat(src, 1).double()
    ref = xin @ lin.weight.double().t() + lin.bias.double()
    if act: ref = torch.nn.functional.logsigmoid(ref)
    ref = ref * mask.double().unsqueeze(1)
    ref32 = torch.nn.functional.linear(torch.cat(src, 1), lin.weight, lin.bias)
    if act: ref32 = torch.nn.functional.logsigmoid(ref32)
    ref32 = ref32 * mask.unsqueeze(1)
    print("linear E=%d %d->%d: max |tc - fp64| = %.3e   max |torch fp32 - fp64| = %.3e" % (
        E, k1 + k2, n, (out.double() - ref).abs().max().item(), (ref32.double() - ref).abs().max().item()), flush=True)

for (E, kx1, kx2, H) in [(1000, 150, 1, 150), (257, 3, 1, 150), (70000, 150, 1, 150)]:
    cell = torch.nn.GRUCell(kx1 + kx2, H).to(dev)
    x1 = torch.randn(E, kx1, device=dev)
    x2 = torch.sign(torch.randn(E, kx2, device=dev))
    h = torch.rand(E, H, device=dev) * 2 - 1
    mask = (torch.rand(E, device=dev) > 0.2).float()
    tg = T.TensorGRU(cell)
    out = tg([x1, x2], h, row_mask=mask)
    torch.cuda.synchronize()
    c64 = torch.nn.GRUCell(kx1 + kx2, H).to(dev).double()
    c64.load_state_dict({k: v.double() for k, v in cell.state_dict().items()})
    ref = c64(torch.cat((x1, x2), 1).double(), h.double())
    ref = mask.double().unsqueeze(1) * ref + (1 - mask.double().unsqueeze(1)) * h.double()
    ref32 = cell(torch.cat((x1, x2), 1), h)
    ref32 = mask.unsqueeze(1) * ref32 + (1 - mask.unsqueeze(1)) * h
    print("gru E=%d %d|%d: max |tc - fp64| = %.3e   max |torch fp32 - fp64| = %.3e" % (
        E, kx1 + kx2, H, (out.double() - ref).abs().max().item(), (ref32.double() - ref).abs().max().item()), flush=True)
    if E >= 70000:
        for f, name in ((lambda: tg([x1, x2], h, row_mask=mask), "tcgen05"), (lambda: cell(torch.cat((x1, x2), 1), h), "torch fp32")):
            f(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): f()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("   %s: %.3f ms per call, %.1f TFLOP/s (fp32-equivalent)" % (name, ms, 2.0 * E * (kx1 + kx2 + H) * 3 * H / ms / 1e9))
