"""Full KKT (x, u, lambda unknowns) in stage-wise banded ordering: dense pivoted LU + refinement."""
import sys, pickle
sys.path.insert(0, '/root/repo')
import numpy as np, scipy.linalg as sl
from oracle import restate as rs
from tools.analysis.abd_proto import abd_solve

def build(prob, fixed, vals, lam_scale=1.0):
    n, m, H = prob.n, prob.m, prob.H
    W = 2 * n + m
    K = np.zeros((W * H, W * H)); b = np.zeros(W * H)
    for t in range(1, H + 1):
        c = (t - 1) * W          # columns: u_{t-1} [m], x_t [n], lam_t [n]
        r = c
        up = np.where(fixed[t - 1], vals[t - 1], 0.0)
        # stationarity / pin rows
        for i in range(m):
            if fixed[t - 1, i]:
                K[r + i, c + i] = 1.0; b[r + i] = vals[t - 1, i]
            else:
                K[r + i, c:c + m] = 2 * prob.R[t - 1][i]; K[r + i, c + m + n:c + W] = prob.B[t - 1][:, i] * lam_scale
                b[r + i] = 2 * prob.R[t - 1][i] @ prob.ub[t - 1]
        r += m
        K[r:r + n, c:c + m] = -prob.B[t - 1]; K[r:r + n, c + m:c + m + n] = np.eye(n)
        b[r:r + n] = prob.D[t - 1]
        if t > 1:
            K[r:r + n, c - W + m:c - W + m + n] = -prob.A[t - 1]
        else:
            b[r:r + n] += prob.A[0] @ prob.x0
        r += n
        K[r:r + n, c + m:c + m + n] = -2 * prob.Q[t]; K[r:r + n, c + m + n:c + W] = np.eye(n) * lam_scale
        b[r:r + n] = -2 * prob.Q[t] @ prob.r[t]
        if t < H:
            K[r:r + n, c + W + m + n:c + 2 * W] = -prob.A[t].T * lam_scale
    return K, b

if __name__ == '__main__':
    cap = pickle.load(open('/root/repo/tools/analysis/h%s_qps.pkl' % sys.argv[1], 'rb'))
    for qi, q in enumerate(cap):
        if qi < 2: continue
        a = q['args']
        prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
        n, m, H = prob.n, prob.m, prob.H
        lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
        Uo = q['U'].T
        fixed = (Uo <= lo + 1e-13) | (Uo >= hi - 1e-13)
        vals = np.where(Uo <= lo + 1e-13, lo, hi)
        Xl, Ul, laml, grl = abd_solve(prob, fixed, vals, dtype=np.longdouble)
        for ls in (1.0, 1e-3):
            K, b = build(prob, fixed, vals, ls)
            lu = sl.lu_factor(K)
            z = sl.lu_solve(lu, b)
            errs = [np.abs(z.reshape(H, -1)[:, :m] - Ul).max()]
            for it in range(3):
                z = z + sl.lu_solve(lu, b - K @ z)
                errs.append(np.abs(z.reshape(H, -1)[:, :m] - Ul).max())
            print('QP %d lam_scale %g cond %.1e: errs' % (qi, ls, np.linalg.cond(K)), ' '.join('%.1e' % e for e in errs), 'U growth', np.abs(sl.lu(K)[2]).max())
