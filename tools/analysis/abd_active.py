"""Primal-dual active-set rounds with the stage-ordered pivoted KKT solve, multipliers from the KKT lambda."""
import sys, pickle
sys.path.insert(0, '/root/repo')
import numpy as np, scipy.linalg as sl
from oracle import restate as rs
from tools.analysis.abd_full import build

def kkt_solve(prob, fixed, vals, refine=0):
    n, m, H = prob.n, prob.m, prob.H
    K, b = build(prob, fixed, vals)
    lu = sl.lu_factor(K)
    z = sl.lu_solve(lu, b)
    for _ in range(refine):
        z = z + sl.lu_solve(lu, b - K @ z)
    z = z.reshape(H, -1)
    U = z[:, :m].copy(); X = np.vstack([prob.x0, z[:, m:m + n]]); lam = z[:, m + n:]
    g = np.zeros((H, m))
    for t in range(H):
        g[t] = 2 * prob.R[t] @ (U[t] - prob.ub[t]) + prob.B[t].T @ lam[t]
    return X, U, g

def rounds(prob, lo, hi, mask, max_rounds=60, refine=0, verbose=False):
    flips = np.zeros_like(mask)
    for rnd in range(max_rounds):
        fixed = mask != 0
        vals = np.where(mask == 1, lo, np.where(mask == 2, hi, 0.0))
        X, U, g = kkt_solve(prob, fixed, vals, refine)
        gs = max(1.0, np.abs(g).max())
        free = mask == 0
        viol_lo = free & (U < lo - 1e-12); viol_hi = free & (U > hi + 1e-12)
        gn = np.where(mask == 1, -g, np.where(mask == 2, g, 0.0))
        rel = (mask != 0) & (gn > 1e-10 * gs) & ((flips < 2) | (gn > 1e-5 * gs))
        if verbose: print('   round', rnd, 'viol', viol_lo.sum() + viol_hi.sum(), 'rel', rel.sum(), 'gs %.2e' % gs, 'free g %.1e' % np.abs(g[free]).max())
        if not (viol_lo.any() or viol_hi.any() or rel.any()):
            return rnd + 1, U, mask
        mask = mask.copy()
        mask[viol_lo] = 1; mask[viol_hi] = 2; mask[rel] = 0; flips[rel] += 1
    return None, U, mask

if __name__ == '__main__':
    cap = pickle.load(open('/root/repo/tools/analysis/h%s_qps.pkl' % sys.argv[1], 'rb'))
    prev = None
    for qi, q in enumerate(cap):
        a = q['args']
        prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
        lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
        Uo = q['U'].T
        mo = np.where(Uo <= lo + 1e-13, 1, np.where(Uo >= hi - 1e-13, 2, 0))
        cold = np.zeros_like(mo)
        r0, U0, m0 = rounds(prob, lo, hi, cold)
        out = 'cold %s err %.1e' % (r0, np.abs(U0 - Uo).max())
        if prev is not None:
            warm = np.vstack([prev[1:], prev[-1:]])
            r1, U1, m1 = rounds(prob, lo, hi, warm, verbose=(qi in (3, 4) and len(sys.argv) > 2))
            out += ' | warm-shift %s err %.1e u0 err %.1e' % (r1, np.abs(U1 - Uo).max(), np.abs(U1[0] - Uo[0]).max())
        print('QP %2d pinned %3d | %s' % (qi, (mo != 0).sum(), out))
        prev = mo
