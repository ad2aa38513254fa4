"""ABD system as one matrix: pivoted LU + iterative refinement (residual on the block equations, fp64)."""
import sys, pickle
sys.path.insert(0, '/root/repo')
import numpy as np, scipy.linalg as sl
from oracle import restate as rs
from tools.analysis.abd_proto import abd_solve
from tools.analysis.abd_qr import stage_data

def build(prob, fixed, vals):
    n, m, H = prob.n, prob.m, prob.H
    N2 = 2 * n
    M, Dt, ufix = stage_data(prob, fixed, vals)
    K = np.zeros((N2 * H, N2 * H)); b = np.zeros(N2 * H)
    K[:n, :N2] = np.hstack([np.eye(n), M[0]]); b[:n] = Dt[0] + prob.A[0] @ prob.x0
    r = n
    for t in range(1, H):
        c = (t - 1) * N2
        K[r:r + n, c:c + n] = -prob.A[t]
        K[r:r + n, c + N2:c + N2 + n] = np.eye(n); K[r:r + n, c + N2 + n:c + 2 * N2] = M[t]
        b[r:r + n] = Dt[t]
        r += n
        K[r:r + n, c:c + n] = -2 * prob.Q[t]; K[r:r + n, c + n:c + N2] = np.eye(n)
        K[r:r + n, c + N2 + n:c + 2 * N2] = -prob.A[t].T
        b[r:r + n] = -2 * prob.Q[t] @ prob.r[t]
        r += n
    c = (H - 1) * N2
    K[r:r + n, c:c + n] = -2 * prob.Q[H]; K[r:r + n, c + n:c + N2] = np.eye(n); b[r:r + n] = -2 * prob.Q[H] @ prob.r[H]
    return K, b, ufix

def recover(prob, z, ufix):
    n, m, H = prob.n, prob.m, prob.H
    z = z.reshape(H, 2 * n)
    lam = np.vstack([np.zeros(n), z[:, n:]])
    U = np.zeros((H, m))
    for t in range(H):
        f, p, Si, Bf, c0, up = ufix[t]
        U[t] = up
        if f.any():
            U[t, f] = c0 - Si @ (Bf.T @ lam[t + 1])
    return U

if __name__ == '__main__':
    cap = pickle.load(open('/root/repo/tools/analysis/h%s_qps.pkl' % sys.argv[1], 'rb'))
    for qi, q in enumerate(cap):
        if qi < 2: continue
        a = q['args']
        prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
        lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
        Uo = q['U'].T
        fixed = (Uo <= lo + 1e-13) | (Uo >= hi - 1e-13)
        vals = np.where(Uo <= lo + 1e-13, lo, hi)
        Xl, Ul, laml, grl = abd_solve(prob, fixed, vals, dtype=np.longdouble)
        K, b, ufix = build(prob, fixed, vals)
        lu = sl.lu_factor(K)
        z = sl.lu_solve(lu, b)
        errs = [np.abs(recover(prob, z, ufix) - Ul).max()]
        for it in range(3):
            z = z + sl.lu_solve(lu, b - K @ z)
            errs.append(np.abs(recover(prob, z, ufix) - Ul).max())
        print('QP %d cond %.1e: LU+refine errs' % (qi, np.linalg.cond(K)), ' '.join('%.1e' % e for e in errs),
              '| |z|max %.1e' % np.abs(z).max())
