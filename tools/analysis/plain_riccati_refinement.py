"""Plain fp64 Riccati + DDP-style refinement: how far does it get?"""
import sys, pickle
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import restate as rs
cap = pickle.load(open('/root/repo/tools/analysis/h%s_qps.pkl' % sys.argv[1], 'rb'))

def factor(prob, mask):
    Hn, m, n = prob.H, prob.m, prob.n
    P = prob.Q[Hn].copy(); fac = [None] * Hn
    for t in reversed(range(Hn)):
        free = mask[t] == 0
        Bf = prob.B[t] * free[None, :]
        PA = P @ prob.A[t]; PB = P @ Bf
        Sm = prob.R[t] + Bf.T @ PB
        for i in range(m):
            if not free[i]: Sm[i, :] = 0; Sm[:, i] = 0; Sm[i, i] = 1
        Si = np.linalg.inv(Sm); T21 = Bf.T @ PA; K = Si @ T21
        fac[t] = (K, Si, Bf, P.copy())
        Pn = prob.Q[t] + prob.A[t].T @ PA - T21.T @ K
        P = 0.5 * (Pn + Pn.T)
    return fac

def backward(prob, fac, mask, D, ql, qlf, h):
    """cost sum x'Qx - 2 ql'x + u'Ru - 2h'u; returns kk."""
    Hn = prob.H; p = qlf.copy(); kk = [None] * Hn
    for t in reversed(range(Hn)):
        K, Si, Bf, P1 = fac[t]
        free = mask[t] == 0
        v = P1 @ D[t] - p
        g = (Bf.T @ v - h[t]) * free
        kk[t] = Si @ g
        p = -(prob.A[t].T @ v - ql[t] - K.T @ g)
    return kk

for qi in range(len(cap)):
    q = cap[qi]; a = q['args']
    prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
    lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
    Uo = q['U'].T
    mask = np.where(Uo <= lo + 1e-13, 1, np.where(Uo >= hi - 1e-13, 2, 0))
    vals = np.where(mask == 1, lo, np.where(mask == 2, hi, 0.0))
    Hn, n, m = prob.H, prob.n, prob.m
    fac = factor(prob, mask)
    Dt = [prob.D[t] + (prob.B[t] * (mask[t] != 0)[None, :]) @ vals[t] for t in range(Hn)]
    ql = [prob.Q[t] @ prob.r[t] for t in range(Hn)]; qlf = prob.Q[Hn] @ prob.r[Hn]
    h = [prob.R[t] @ prob.ub[t] for t in range(Hn)]
    kk = backward(prob, fac, mask, Dt, ql, qlf, h)
    X = np.zeros((Hn + 1, n)); U = np.zeros((Hn, m)); X[0] = prob.x0
    for t in range(Hn):
        K, Si, Bf, P1 = fac[t]
        U[t] = np.where(mask[t] == 0, -(K @ X[t]) - kk[t], vals[t])
        X[t + 1] = prob.A[t] @ X[t] + prob.B[t] @ U[t] + prob.D[t]
    zD = [np.zeros(n)] * Hn
    hist = []
    for it in range(6):
        g = prob.gradient(X, U); gf = np.where(mask == 0, g, 0.0)
        hist.append((np.abs(U - Uo).max(), np.abs(gf).max()))
        kk = backward(prob, fac, mask, zD, zD, np.zeros(n), [-0.5 * gf[t] for t in range(Hn)])
        Xn = np.zeros_like(X); Un = np.zeros_like(U); Xn[0] = X[0]
        for t in range(Hn):
            K, Si, Bf, P1 = fac[t]
            du = -(K @ (Xn[t] - X[t])) - kk[t]
            Un[t] = np.where(mask[t] == 0, U[t] + du, U[t])
            Xn[t + 1] = prob.A[t] @ Xn[t] + prob.B[t] @ Un[t] + prob.D[t]
        X, U = Xn, Un
    print('QP %d: (|U-Uo|, |g_free|) per it: %s' % (qi, ' '.join('(%.1e,%.1e)' % x for x in hist)))
