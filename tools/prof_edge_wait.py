#!/usr/bin/env python3
"""Where the roles of k_edge_nn wait (profiling build -DPDP_NN_TIMING, PDP_B200_LIB=...alt_nnt.so)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdp_solver_b200 import _lib
from pdp_solver_b200.nn import tensor_ops as T
E = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
dev = torch.device("cuda:0")
cell = torch.nn.GRUCell(151, 150).to(dev)
x1, x2, h = torch.randn(E, 150, device=dev), torch.sign(torch.randn(E, 1, device=dev)), torch.rand(E, 150, device=dev)
tg = T.TensorGRU(cell)
tg([x1, x2], h); torch.cuda.synchronize()
L = _lib.load()
buf = (ctypes.c_ulonglong * 8)()
L.pdp_edge_nn_wait_counters(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); tg([x1, x2], h); e1.record(); torch.cuda.synchronize()
L.pdp_edge_nn_wait_counters(buf, 0)
ms = e0.elapsed_time(e1)
cyc = ms * 1e-3 * 1.965e9
names = ["A producers on empty (8 warps)", "B producer on empty", "issuer on full_a", "issuer on full_b", "issuer on acc_empty", "epilogue on acc_full (8 warps)", "issuer: MMA issue", "issuer: commit"]
print("GRU E=%d: %.3f ms = %.0f cycles per CTA" % (E, ms, cyc))
for i, n in enumerate(names):
    per = buf[i] / 148.0 / (8 if "8 warps" in n else 1)
    print("  %-34s %10.0f cycles per CTA (%.0f %% of the kernel)" % (n, per, 100 * per / cyc))

log = (ctypes.c_longlong * (96 * 8))()
L.pdp_edge_nn_event_log(log)
import numpy as np
ev = np.array(list(log), dtype=np.int64).reshape(96, 8)
t0 = ev[0, 0]
print("CTA 0, chunks 40..71 (cycles since the first event): producer sees empty | producer arrives | B copy issued | issuer sees full | MMAs issued | committed")
for c in range(40, 72):
    print("  chunk %2d: " % c + "  ".join("%8d" % (ev[c, k] - t0) for k in range(6)) + "   full-empty %6d  mma-full %6d   empty(c+4) - issued(c) %6d" % (ev[c, 3] - ev[c, 0], ev[c, 4] - ev[c, 3], ev[c + 4, 0] - ev[c, 4]))
print("chunk period (MMAs issued, chunks 40..71): %.0f cycles" % ((ev[71, 4] - ev[40, 4]) / 31.0))
try:
    plog = (ctypes.c_longlong * (96 * 32))()
    L.pdp_edge_nn_producer_log(plog)
    pv = np.array(list(plog), dtype=np.int64).reshape(96, 4, 8)
    print("producer warps 0..7, chunks 44..56, cycles: previous arrive -> operands in registers | -> next loads issued (reached the wait) | -> saw empty | -> arrived")
    for c in range(44, 57):
        print("  chunk %2d  data " % c + " ".join("%5d" % (pv[c, 3, w] - pv[c - 1, 2, w]) for w in range(8)) + "   loads " +
              " ".join("%5d" % (pv[c, 0, w] - pv[c, 3, w]) for w in range(8)) + "   empty " + " ".join("%5d" % (pv[c, 1, w] - pv[c, 0, w]) for w in range(8)) +
              "   arrive " + " ".join("%5d" % (pv[c, 2, w] - pv[c, 1, w]) for w in range(8)))
except Exception as e:
    print("no producer log:", e)
