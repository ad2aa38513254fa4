"""Prototype: stable solve of the equality-constrained QP (given working set) as a two-point BVP in (x, lambda)
by almost-block-diagonal elimination with row partial pivoting.  Compare with the oracle's sparse KKT solve."""
import sys, pickle
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import restate as rs

def abd_solve(prob, fixed, vals, dtype=np.float64):
    n, m, H = prob.n, prob.m, prob.H
    N2 = 2 * n
    # stage data
    M = []; Dt = []; ufix = []
    for t in range(H):
        f = ~fixed[t]; p = fixed[t]
        R = prob.R[t]; B = prob.B[t]
        up = np.where(p, vals[t], 0.0)
        S = 2 * R[np.ix_(f, f)]
        Bf = B[:, f]
        Si = np.linalg.inv(S) if f.any() else np.zeros((0, 0))
        # free: 2R_ff(u_f - ub_f) + 2R_fp(u_p - ub_p) + Bf^T lam = 0
        c0 = prob.ub[t][f] - Si @ (2 * R[np.ix_(f, p)] @ (up[p] - prob.ub[t][p])) if f.any() else np.zeros(0)
        M.append(Bf @ Si @ Bf.T)
        Dt.append(prob.D[t] + B[:, p] @ up[p] + Bf @ c0)
        ufix.append((f, p, Si, Bf, c0, up))
    # top block on z_1: x_1 + M_0 lam_1 = Dt_0 + A_0 x_0
    top = np.hstack([np.eye(n), M[0]]).astype(dtype)
    trhs = (Dt[0] + prob.A[0] @ prob.x0).astype(dtype)
    Us = []; rhs_s = []
    growth = 0.0
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
        rows = W.shape[0]
        for k in range(N2):
            piv = k + np.argmax(np.abs(W[k:, k]))
            if piv != k:
                W[[k, piv]] = W[[piv, k]]; b[[k, piv]] = b[[piv, k]]
            l = W[k + 1:, k] / W[k, k]
            W[k + 1:, k:] -= np.outer(l, W[k, k:])
            b[k + 1:] -= l * b[k]
        growth = max(growth, np.abs(W).max())
        Us.append(W[:N2].copy()); rhs_s.append(b[:N2].copy())
        if t < H:
            top = W[N2:, N2:].copy(); trhs = b[N2:].copy()
    # back substitution
    z = np.zeros((H + 1, N2), dtype=dtype)
    znext = None
    for t in range(H, 0, -1):
        Wt = Us[t - 1]; b = rhs_s[t - 1].copy()
        if t < H:
            b -= Wt[:, N2:] @ znext
        zt = np.zeros(N2, dtype=dtype)
        for k in range(N2 - 1, -1, -1):
            zt[k] = (b[k] - Wt[k, k + 1:N2] @ zt[k + 1:]) / Wt[k, k]
        z[t] = zt; znext = zt
    X = np.zeros((H + 1, n)); X[0] = prob.x0; X[1:] = z[1:, :n]
    lam = z[:, n:]
    U = np.zeros((H, m))
    for t in range(H):
        f, p, Si, Bf, c0, up = ufix[t]
        U[t] = up
        if f.any():
            U[t, f] = c0 - Si @ (Bf.T @ lam[t + 1])
    return X, U, lam.astype(float), growth

def mults(prob, U, lam):
    """gradient dJ/du_t = 2R(u-ub) + B^T lam_{t+1}"""
    g = np.zeros_like(U)
    for t in range(prob.H):
        g[t] = 2 * prob.R[t] @ (U[t] - prob.ub[t]) + prob.B[t].T @ lam[t + 1]
    return g

if __name__ == '__main__':
    cap = pickle.load(open('/root/repo/tools/analysis/h%s_qps.pkl' % sys.argv[1], 'rb'))
    for qi, q in enumerate(cap):
        a = q['args']
        prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
        lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
        Uo = q['U'].T
        fixed = (Uo <= lo + 1e-13) | (Uo >= hi - 1e-13)
        vals = np.where(Uo <= lo + 1e-13, lo, hi)
        Xs, Us_ = prob.solve_fixed(fixed, vals)
        X, U, lam, gr = abd_solve(prob, fixed, vals)
        g = mults(prob, U, lam)
        go = prob.gradient(Xs, Us_)
        print('QP %d pinned %d: |U-Uoracle| %.2e |U-Usparse| %.2e |X-Xsparse| %.2e rel %.2e growth %.1e  |g-g_adj(oracle)| %.2e gmax %.2e free-g %.2e'
              % (qi, fixed.sum(), np.abs(U - Uo).max(), np.abs(U - Us_).max(), np.abs(X - Xs).max(), np.abs(X - Xs).max() / np.abs(Xs).max(),
                 gr, np.abs(g - go).max(), np.abs(g).max(), np.abs(g[~fixed]).max() if (~fixed).any() else 0))
    print('--- long double reference')
    for qi, q in enumerate(cap):
        if qi < 3: continue
        a = q['args']
        prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
        lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
        Uo = q['U'].T
        fixed = (Uo <= lo + 1e-13) | (Uo >= hi - 1e-13)
        vals = np.where(Uo <= lo + 1e-13, lo, hi)
        Xs, Us_ = prob.solve_fixed(fixed, vals)
        X, U, lam, gr = abd_solve(prob, fixed, vals)
        Xl, Ul, laml, grl = abd_solve(prob, fixed, vals, dtype=np.longdouble)
        print('QP %d: |U64-Uld| %.2e  |Usparse-Uld| %.2e' % (qi, np.abs(U - Ul).max(), np.abs(Us_ - Ul).max()))
