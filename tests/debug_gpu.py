"""Ad-hoc GPU-vs-oracle divergence finder (not a test)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import pdp_oracle as po
from pdp_solver_b200 import cnfgen
from pdp_solver_b200.engine import Context

def T(x): return torch.from_numpy(np.ascontiguousarray(x)).cuda()
def C(t): return t.detach().cpu().numpy()

def run(spec):
    Bn, n, k, alpha, Tn, seed = spec
    batch = cnfgen.random_batch(Bn, n, k, alpha, seed)
    E = batch[0].shape[1]
    rng = np.random.default_rng(seed)
    init = po.init_state(E, randomized=(seed % 2 == 0), rng=rng)
    tol, t_max = 0.02, 25
    o = po.Oracle(*batch, strict=False)
    o.simplify(); o.set_state(*init)
    ctx = Context(*[T(x) for x in batch])
    ctx.enable_trace()
    ctx.simplify(); ctx.load_state((T(init[0][0]), T(init[0][1])), (T(init[1][0]), T(init[1][1])))
    m, om = ctx.get_masks(), o.masks()
    print(spec, 'after simplify av eq', (C(m['av']) == om['av']).all(), 'af eq', (C(m['af']) == om['af']).all())
    for t in range(Tn):
        na = o.iterate(tol, t_max, True)
        ctx.sp_run(1, tol, t_max, True, sync=True)
        m, om = ctx.get_masks(), o.masks()
        q, fs = ctx.store_state(); oq, ofs = o.state()
        ok = (C(m['av']) == om['av']).all() and (C(m['af']) == om['af']).all() and (C(m['active']) == om['active']).all()
        d = np.nanmax(np.abs(C(fs[:, 0]) - ofs[:, 0])); dq = np.nanmax(np.abs(C(q[:, 0]) - oq[:, 0]))
        _, cnt, _ = ctx.problem_flags()
        if not ok or t % 20 == 0:
            print('iter', t + 1, 'ok', ok, 'deta', d, 'dq', dq, 'na', na, 'gpu active', int(C(m['active']).sum()))
        if not ok:
            bad_v = np.nonzero(C(m['av']) != om['av'])[0]; bad_b = np.nonzero(C(m['active']) != om['active'])[0]
            print(' bad vars', bad_v[:10], 'problems', np.unique(batch[1][bad_v])[:10], 'bad active', bad_b[:10])
            tr = C(ctx.trace()); otr = o.trace()
            print(' gpu events this iter', tr[tr[:, 0] == t + 1].tolist()[:10]); print(' ora events this iter', otr[otr[:, 0] == t + 1].tolist()[:10])
            b = int(np.unique(batch[1][bad_v])[0]) if len(bad_v) else int(bad_b[0])
            print(' counters gpu', C(cnt)[b], 'ora', o.counters()[b])
            return
        if na <= 0: break
    print(' all iterations agree')

for spec in [(64, 100, 3, 4.2, 150, 7), (8, 600, 3, 4.0, 120, 8), (3, 4000, 3, 3.9, 60, 9), (6, 300, 5, 17.0, 80, 10)]:
    run(spec)
