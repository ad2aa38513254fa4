import sys, pickle
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools/analysis')
import numpy as np
from qp_sqrt_riccati import Robust, rs
cap = pickle.load(open('/root/repo/tools/analysis/h100_qps.pkl', 'rb'))
def mask_of(U, lo, hi): return np.where(U <= lo + 1e-13, 1, np.where(U >= hi - 1e-13, 2, 0))

def solve(q, warm, n_ref=2, mtol_noise=1e-5, max_rounds=40, verbose=False):
    a = q['args']
    prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
    lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
    rb = Robust(prob, lo, hi)
    mask = warm.copy(); flips = np.zeros_like(mask)
    for rnd in range(max_rounds):
        vals = np.where(mask == 1, lo, np.where(mask == 2, hi, 0.0))
        fac = rb.factor(mask)
        X, U = rb.solve(fac, mask, vals)
        for it in range(n_ref):
            g = prob.gradient(X, U)
            X, U = rb.refine(fac, mask, X, U, g)
        g = prob.gradient(X, U)
        gs = max(1.0, np.abs(g).max())
        free = mask == 0
        nm = mask.copy()
        nm[free & (U < lo - 1e-9)] = 1
        nm[free & (U > hi + 1e-9)] = 2
        gn = np.where(mask == 1, -g, np.where(mask == 2, g, 0.0))
        tol = np.where(flips >= 2, 1e-2, mtol_noise) * gs
        rel = (mask != 0) & (gn > tol)
        nm[rel] = 0; flips[rel] += 1
        if verbose: print('   rnd', rnd, 'pinned', (mask != 0).sum(), 'changed', (nm != mask).sum(), 'gmax %.2e' % np.abs(g).max())
        if (nm == mask).all(): break
        mask = nm
    return U, rnd + 1, mask

prev = None
for qi, q in enumerate(cap):
    a = q['args']
    lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
    Uo = q['U'].T
    if prev is not None:
        warm = mask_of(np.clip(np.vstack([prev[1:], prev[-1:]]), lo, hi), lo, hi)
        for mt in (1e-10, 1e-5, 1e-3):
            U, rounds, mask = solve(q, warm, mtol_noise=mt)
            print('QP %d mtol %.0e: rounds %2d  |U-Uo| %.2e first col %.2e  pinned %d (oracle %d)' % (qi, mt, rounds, np.abs(U - Uo).max(), np.abs(U[0] - Uo[0]).max(), (mask != 0).sum(), (mask_of(Uo, lo, hi) != 0).sum()))
    prev = Uo
