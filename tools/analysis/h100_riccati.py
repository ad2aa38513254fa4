"""Masked Riccati (as the device does) on captured QPs: accuracy vs the sparse-KKT oracle, in fp64 / longdouble / scaled."""
import sys, pickle
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import restate as rs
np.set_printoptions(linewidth=200, precision=4)
H = sys.argv[1]
cap = pickle.load(open('/root/repo/tools/analysis/h%s_qps.pkl' % H, 'rb'))

def riccati(prob, lo, hi, mask, dtype=np.float64, scale=None, refine=0):
    """mask [H, m]: 0 free, 1 pinned lo, 2 pinned hi. Returns U [H, m], X [H+1, n]."""
    Hn, m, n = prob.H, prob.m, prob.n
    f = lambda a: np.asarray(a, dtype=dtype)
    A = [f(a) for a in prob.A]; B = [f(b) for b in prob.B]; D = [f(d) for d in prob.D]
    Q = [f(q) for q in prob.Q]; R = [f(r) for r in prob.R]; r = [f(x) for x in prob.r]; ub = [f(u) for u in prob.ub]
    # scaling: x_t = S_t xs_t
    if scale is not None:
        S = [f(s) for s in scale]   # list of H+1 diagonal vectors
        A = [ (A[t] * S[t][None, :]) / S[t + 1][:, None] for t in range(Hn)]
        B = [ B[t] / S[t + 1][:, None] for t in range(Hn)]
        D = [ D[t] / S[t + 1] for t in range(Hn)]
        Q = [ Q[t] * S[t][None, :] * S[t][:, None] for t in range(Hn + 1)]
        r = [ r[t] / S[t] for t in range(Hn + 1)]
        x0 = f(prob.x0) / S[0]
    else:
        x0 = f(prob.x0)
    # cost: sum (x-r)'Q(x-r) + (u-ub)'R(u-ub); V_t(x) = x'P x - 2 p'x
    P = Q[Hn].copy(); p = Q[Hn] @ r[Hn]
    K = [None] * Hn; kk = [None] * Hn; Sinv = [None] * Hn
    Bt = [None] * Hn; Dt = [None] * Hn
    for t in reversed(range(Hn)):
        free = mask[t] == 0
        bnd = np.where(mask[t] == 1, lo[t], hi[t]).astype(dtype)
        Bf = B[t] * free[None, :]
        Df = D[t] + (B[t] * (~free)[None, :]) @ np.where(free, 0, bnd).astype(dtype)
        Bt[t], Dt[t] = Bf, Df
        PA = P @ A[t]; PB = P @ Bf
        Sm = R[t] + Bf.T @ PB
        for i in range(m):
            if not free[i]:
                Sm[i, :] = 0; Sm[:, i] = 0; Sm[i, i] = 1
        Si = np.linalg.inv(Sm.astype(np.float64)).astype(dtype) if dtype != np.float64 else np.linalg.inv(Sm)
        if dtype != np.float64:   # newton refine inverse in extended precision
            for _ in range(3): Si = Si @ (2 * np.eye(m, dtype=dtype) - Sm @ Si)
        T21 = Bf.T @ PA
        Kt = Si @ T21
        # linear term: v = P D - p ; h = R ub (free rows)
        v = P @ Df - p
        h = (R[t] @ ub[t]) * free
        g = Bf.T @ v - h
        kt = Si @ (g * free)
        K[t], kk[t], Sinv[t] = Kt, kt, Si
        pn = A[t].T @ v - Q[t] @ r[t] - Kt.T @ (g * free)
        Pn = Q[t] + A[t].T @ PA - T21.T @ Kt
        P = 0.5 * (Pn + Pn.T); p = -pn
    X = np.zeros((Hn + 1, n), dtype=dtype); U = np.zeros((Hn, m), dtype=dtype)
    X[0] = x0
    for t in range(Hn):
        free = mask[t] == 0
        bnd = np.where(mask[t] == 1, lo[t], hi[t]).astype(dtype)
        u = -(K[t] @ X[t]) - kk[t]
        u = np.where(free, u, bnd)
        U[t] = u
        X[t + 1] = A[t] @ X[t] + B[t] @ u + D[t]
    if scale is not None:
        X = X * np.array(S)
    return U.astype(np.float64), X.astype(np.float64)

for qi, q in enumerate(cap):
    a = q['args']
    prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
    lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
    Uo = q['U'].T
    # active set of the oracle: at bound
    Xr = prob.rollout(Uo); g = prob.gradient(Xr, Uo)
    mask = np.where(Uo <= lo + 1e-13, 1, np.where(Uo >= hi - 1e-13, 2, 0))
    print('QP', qi, 'pinned', (mask != 0).sum(), 'grad on free max', np.abs(g[mask == 0]).max(), 'gs', np.abs(g).max())
    for name, kw in [('fp64', {}), ('longdouble', dict(dtype=np.longdouble))]:
        U, X = riccati(prob, lo, hi, mask, **kw)
        print('  %-12s |U - Uo| max %.3e   free-only %.3e' % (name, np.abs(U - Uo).max(), np.abs((U - Uo)[mask == 0]).max()))
    # diagonal scaling from column-norm growth of backward product
    n = prob.n
    g_ = np.ones(n); S = [None] * (prob.H + 1); S[prob.H] = np.ones(n)
    Phi = np.eye(n)
    for t in reversed(range(prob.H)):
        Phi = Phi @ prob.A[t]
        cn = np.maximum(np.linalg.norm(Phi, axis=0), 1.0)
        S[t] = 1.0 / 2.0 ** np.round(np.log2(cn))
        Phi = Phi * (S[t] / (S[t+1] if False else 1.0))[None, :] if False else Phi
    U, X = riccati(prob, lo, hi, mask, scale=S)
    print('  %-12s |U - Uo| max %.3e   free-only %.3e' % ('diag-scaled', np.abs(U - Uo).max(), np.abs((U - Uo)[mask == 0]).max()))

print('---- device-like active-set rounds from a cold start')
for qi, q in enumerate(cap[:1]):
    a = q['args']
    prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
    lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
    Uo = q['U'].T
    for solver in ('riccati', 'riccati_ld', 'kkt'):
        mask = np.zeros_like(Uo, dtype=int)
        print(' solver', solver)
        for rnd in range(14):
            if solver == 'kkt':
                fixed = mask != 0
                vals = np.where(mask == 1, lo, np.where(mask == 2, hi, 0.0))
                X, U = prob.solve_fixed(fixed, vals)
                Xr = X
            else:
                U, X = riccati(prob, lo, hi, mask, dtype=np.longdouble if solver.endswith('ld') else np.float64)
                Xr = X
            g = prob.gradient(Xr, U)
            gs = max(1.0, np.abs(g).max())
            free = mask == 0
            nm = mask.copy()
            nm[free & (U < lo - 1e-12)] = 1
            nm[free & (U > hi + 1e-12)] = 2
            nm[(mask == 1) & (g < -1e-10 * gs)] = 0
            nm[(mask == 2) & (g > 1e-10 * gs)] = 0
            vis = np.abs(g[free]).max() if free.any() else 0.0
            print('  round %2d pinned %3d |U|max %.3e |X|max %.3e gmax %.3e free-grad %.3e changed %3d  |U-Uo| %.2e' % (
                rnd, (mask != 0).sum(), np.abs(U).max(), np.abs(X).max(), np.abs(g).max(), vis, (nm != mask).sum(), np.abs(U - Uo).max()))
            if (nm == mask).all():
                break
            mask = nm
