import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pdp_oracle as po
from pdp_solver_b200 import cnfgen
from pdp_solver_b200.engine import Context
dev = torch.device("cuda:0")
T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
Bn, n, k, alpha, Tn, seed = (64, 100, 3, 4.2, 150, 7)
batch = cnfgen.random_batch(Bn, n, k, alpha, seed)
gm, bvm = batch[0], batch[1]
E = gm.shape[1]
rng = np.random.default_rng(seed)
init = po.init_state(E, randomized=(seed % 2 == 0), rng=rng)
o = po.Oracle(*batch, strict=False); o.simplify(); o.set_state(*init)
states = []
for t in range(Tn):
    states.append((o.state(), o.masks()))
    if o.run(1, 0.02, 25, True) == 0: break
states.append((o.state(), o.masks()))
ctx = Context(*[T(x) for x in batch])
eprob = bvm[gm[0]]
for t in range(1, len(states) - 1):
    (q0, f0), m0 = states[t]; (q1, f1), _ = states[t + 1]
    running = m0["active"].astype(bool)
    clean = np.ones(o.B, bool); clean[eprob[np.isnan(f0[:, 0]) | np.isnan(q0[:, 0])]] = False
    sel = (running & clean)[eprob]
    if not sel.any(): continue
    ctx.reset(); ctx.set_masks(T(m0["av"]), T(m0["af"]), T(m0["sol"])); ctx.load_state((T(q0), T(f0)), (T(q0), T(f0)))
    ctx.sp_run(1, 0.02, 25, False, sync=True)
    q, fs = ctx.store_state()
    gq = q[:, 0].cpu().numpy(); ge = fs[:, 0].cpu().numpy()
    bad = sel & (np.isnan(gq) != np.isnan(q1[:, 0]))
    if bad.any():
        idx = np.nonzero(bad)[0]
        print("t", t, "mismatch", len(idx), "of", sel.sum())
        for e in idx[:6]:
            v = gm[0][e]
            ev = np.nonzero(gm[0] == v)[0]
            print(" edge", e, "var", v, "prob", eprob[e], "gpu q", gq[e], "oracle q", q1[e, 0], "em", m0["em"][e], "av", m0["av"][v])
            print("   eta(t-1) on var edges", f0[ev, 0], "signs", batch[3][ev], "em", m0["em"][ev])
        break
