"""ABD elimination with Householder QR per stage block instead of pivoted LU."""
import sys, pickle
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import restate as rs
from tools.analysis.abd_proto import abd_solve, mults

def stage_data(prob, fixed, vals):
    n, m, H = prob.n, prob.m, prob.H
    M = []; Dt = []; ufix = []
    for t in range(H):
        f = ~fixed[t]; p = fixed[t]
        R = prob.R[t]; B = prob.B[t]
        up = np.where(p, vals[t], 0.0)
        S = 2 * R[np.ix_(f, f)]
        Bf = B[:, f]
        Si = np.linalg.inv(S) if f.any() else np.zeros((0, 0))
        c0 = prob.ub[t][f] - Si @ (2 * R[np.ix_(f, p)] @ (up[p] - prob.ub[t][p])) if f.any() else np.zeros(0)
        M.append(Bf @ Si @ Bf.T)
        Dt.append(prob.D[t] + B[:, p] @ up[p] + Bf @ c0)
        ufix.append((f, p, Si, Bf, c0, up))
    return M, Dt, ufix

def abd_qr(prob, fixed, vals, dtype=np.float64, scale_rows=False):
    n, m, H = prob.n, prob.m, prob.H
    N2 = 2 * n
    M, Dt, ufix = stage_data(prob, fixed, vals)
    top = np.hstack([np.eye(n), M[0]]).astype(dtype)
    trhs = (Dt[0] + prob.A[0] @ prob.x0).astype(dtype)
    Rs = []; rhs_s = []
    for t in range(1, H + 1):
        if t < H:
            E = np.block([[-prob.A[t], np.zeros((n, n))], [-2 * prob.Q[t], np.eye(n)]])
            F = np.block([[np.eye(n), M[t]], [np.zeros((n, n)), -prob.A[t].T]])
            g = np.concatenate([Dt[t], -2 * prob.Q[t] @ prob.r[t]])
            W = np.vstack([np.hstack([top, np.zeros((n, N2))]), np.hstack([E, F])]).astype(dtype)
            b = np.concatenate([trhs, g]).astype(dtype)
        else:
            E = np.hstack([-2 * prob.Q[H], np.eye(n)])
            W = np.vstack([top, E]).astype(dtype)
            b = np.concatenate([trhs, -2 * prob.Q[H] @ prob.r[H]]).astype(dtype)
        Wb = np.hstack([W, b[:, None]])
        Qm, Rm = np.linalg.qr(Wb[:, :N2].astype(np.float64), mode='complete')
        Wb = Qm.T @ Wb
        Rs.append(Wb[:N2, :-1].copy()); rhs_s.append(Wb[:N2, -1].copy())
        if t < H:
            top = Wb[N2:, N2:-1].copy(); trhs = Wb[N2:, -1].copy()
            if scale_rows:
                s = np.abs(top).max(axis=1); top /= s[:, None]; trhs /= s
    z = np.zeros((H + 1, N2)); znext = None
    for t in range(H, 0, -1):
        Wt = Rs[t - 1]; b = rhs_s[t - 1].copy()
        if t < H:
            b -= Wt[:, N2:] @ znext
        zt = np.linalg.solve(np.triu(Wt[:, :N2]), b)
        z[t] = zt; znext = zt
    X = np.zeros((H + 1, n)); X[0] = prob.x0; X[1:] = z[1:, :n]
    lam = z[:, n:]
    U = np.zeros((H, m))
    for t in range(H):
        f, p, Si, Bf, c0, up = ufix[t]
        U[t] = up
        if f.any():
            U[t, f] = c0 - Si @ (Bf.T @ lam[t + 1])
    return X, U, lam

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
        X, U, lam = abd_qr(prob, fixed, vals)
        X2, U2, lam2 = abd_qr(prob, fixed, vals, scale_rows=True)
        print('QP %d: QR |U-Uld| %.2e  scaled %.2e' % (qi, np.abs(U - Ul).max(), np.abs(U2 - Ul).max()))
