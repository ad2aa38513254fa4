"""Interior-point prototype for the box QP on top of the stage-ordered pivoted KKT solve: every iteration is ONE solve
with a diagonal shift on R (all controls free), then an active-set polish (primal-dual rounds started from the
interior-point estimate of the working set)."""
import sys, pickle, glob
sys.path.insert(0, '/root/repo')
import numpy as np, scipy.linalg as sl, warnings
warnings.simplefilter('ignore')
from oracle import restate as rs

def build(prob, fixed, vals, sig=None, shift=None):
    n, m, H = prob.n, prob.m, prob.H
    W = 2 * n + m
    K = np.zeros((W * H, W * H)); b = np.zeros(W * H)
    for t in range(1, H + 1):
        c = (t - 1) * W; r = c
        for i in range(m):
            if fixed[t - 1, i]:
                K[r + i, c + i] = 1.0; b[r + i] = vals[t - 1, i]
            else:
                K[r + i, c:c + m] = 2 * prob.R[t - 1][i]; K[r + i, c + m + n:c + W] = prob.B[t - 1][:, i]
                b[r + i] = 2 * prob.R[t - 1][i] @ prob.ub[t - 1]
                if sig is not None:
                    K[r + i, c + i] += sig[t - 1, i]; b[r + i] += shift[t - 1, i]
        r += m
        K[r:r + n, c:c + m] = -prob.B[t - 1]; K[r:r + n, c + m:c + m + n] = np.eye(n)
        b[r:r + n] = prob.D[t - 1]
        if t > 1: K[r:r + n, c - W + m:c - W + m + n] = -prob.A[t - 1]
        else: b[r:r + n] += prob.A[0] @ prob.x0
        r += n
        K[r:r + n, c + m:c + m + n] = -2 * prob.Q[t]; K[r:r + n, c + m + n:c + W] = np.eye(n)
        b[r:r + n] = -2 * prob.Q[t] @ prob.r[t]
        if t < H: K[r:r + n, c + W + m + n:c + 2 * W] = -prob.A[t].T
    return K, b

def solve(prob, fixed, vals, sig=None, shift=None):
    n, m, H = prob.n, prob.m, prob.H
    K, b = build(prob, fixed, vals, sig, shift)
    z = sl.lu_solve(sl.lu_factor(K), b).reshape(H, -1)
    U = z[:, :m].copy(); lam = z[:, m + n:]
    g = np.array([2 * prob.R[t] @ (U[t] - prob.ub[t]) + prob.B[t].T @ lam[t] for t in range(H)])
    return U, g

def ipm(prob, lo, hi, sigma=0.1, tau=0.995, max_it=60, mu_tol=1e-11, verbose=False):
    H, m = lo.shape
    nofix = np.zeros((H, m), bool); zeros = np.zeros((H, m))
    u = 0.5 * (lo + hi)
    sl_, su = u - lo, hi - u
    zl = np.ones((H, m)); zu = np.ones((H, m))
    n_solve = 0
    for it in range(max_it):
        mu = (np.sum(sl_ * zl) + np.sum(su * zu)) / (2 * H * m)
        if mu < mu_tol: break
        Sig = zl / sl_ + zu / su
        shift = Sig * u + sigma * mu * (1 / sl_ - 1 / su)
        up, g = solve(prob, nofix, zeros, Sig, shift); n_solve += 1
        du = up - u
        dzl = sigma * mu / sl_ - zl - (zl / sl_) * du
        dzu = sigma * mu / su - zu + (zu / su) * du
        def maxstep(v, dv):
            neg = dv < 0
            return min(1.0, (tau * (-v[neg] / dv[neg])).min()) if neg.any() else 1.0
        ap = min(maxstep(sl_, du), maxstep(su, -du)); ad = min(maxstep(zl, dzl), maxstep(zu, dzu))
        u = u + ap * du; sl_, su = u - lo, hi - u
        zl = zl + ad * dzl; zu = zu + ad * dzu
        if verbose: print('   it %2d mu %.2e ap %.3f ad %.3f' % (it, mu, ap, ad))
    return u, zl, zu, n_solve

def polish(prob, lo, hi, u, zl, zu, max_rounds=20):
    H, m = lo.shape
    mask = np.where((u - lo) < zl, 1, np.where((hi - u) < zu, 2, 0))
    flips = np.zeros((H, m), int)
    for rnd in range(max_rounds):
        fixed = mask != 0
        vals = np.where(mask == 1, lo, np.where(mask == 2, hi, 0.0))
        U, g = solve(prob, fixed, vals)
        U = np.where(fixed, vals, U)
        gs = max(1.0, np.abs(g).max())
        vl = ~fixed & (U < lo - 1e-12); vh = ~fixed & (U > hi + 1e-12)
        gn = np.where(mask == 1, -g, np.where(mask == 2, g, 0.0))
        rel = fixed & (gn > 1e-10 * gs) & ((flips < 2) | (gn > 1e-5 * gs))
        if not (vl.any() or vh.any() or rel.any()): return U, rnd + 1
        mask = mask.copy(); mask[vl] = 1; mask[vh] = 2; mask[rel] = 0; flips[rel] += 1
    return U, None

if __name__ == '__main__':
    files = sorted(glob.glob('/root/repo/tools/analysis/h100_m*_q*.pkl') + glob.glob('/root/repo/tools/analysis/h100_fail_m*.pkl'))
    items = []
    for f in files:
        obj = pickle.load(open(f, 'rb'))
        a = obj[0]; Uo = obj[2][1].T if len(obj) > 2 else None
        items.append((f.split('/')[-1], a, Uo))
    cap = pickle.load(open('/root/repo/tools/analysis/h100_qps.pkl', 'rb'))
    for qi in (0, 3, 4, 5, 9, 13):
        items.append(('nominal q%d' % qi, cap[qi]['args'], cap[qi]['U'].T))
    for name, a, Uo in items:
        prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
        lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
        if Uo is None:
            Uo = rs.qp_exact(*a)[1].T
        u, zl, zu, n = ipm(prob, lo, hi, verbose=(len(sys.argv) > 1))
        U, r = polish(prob, lo, hi, u, zl, zu)
        print('%-22s IPM solves %d (|u-U*| %.1e), polish rounds %s, final err %.1e' % (name, n, np.abs(u - Uo).max(), r, np.abs(U - Uo).max()))
