"""numpy prototype of the robust fallback QP: square-root (QR array) Riccati + DDP-style refinement + primal-dual active set."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import restate as rs

def psd_sqrt(q):
    w, v = np.linalg.eigh(q)
    return (v * np.sqrt(np.maximum(w, 0))) @ v.T

class Robust:
    def __init__(self, prob, lo, hi):
        self.p, self.lo, self.hi = prob, lo, hi
        self.Qh = [psd_sqrt(q) for q in prob.Q]
        self.Rh = [np.linalg.cholesky(r).T for r in prob.R]
        self.nfac = 0

    def factor(self, mask):
        p = self.p; Hn, m, n = p.H, p.m, p.n
        L = self.Qh[Hn].copy(); fac = [None] * Hn
        for t in reversed(range(Hn)):
            free = mask[t] == 0
            Bf = p.B[t] * free[None, :]
            Rt = self.Rh[t].copy()
            if not free.all():
                Rf = p.R[t].copy()
                for i in range(m):
                    if not free[i]: Rf[i, :] = 0; Rf[:, i] = 0; Rf[i, i] = 1.0
                Rt = np.linalg.cholesky(Rf).T
            pre = np.block([[L @ Bf, L @ p.A[t]], [Rt, np.zeros((m, n))], [np.zeros((n, m)), self.Qh[t]]])
            Qm, Rm = np.linalg.qr(pre, mode='reduced')
            fac[t] = (Qm, Rm[:m, :m], Rm[:m, m:], L, Rt)
            L = Rm[m:, m:]
        self.nfac += 1
        return fac

    def backward(self, fac, mask, lH, d, rho, kap):
        p = self.p; Hn, m = p.H, p.m
        l = lH; ls = [None] * Hn
        for t in reversed(range(Hn)):
            Qm, S12, Kh, Lp, Rt = fac[t]
            z = Qm.T @ np.concatenate([l - Lp @ d[t], rho[t], kap[t]])
            ls[t] = z[:m]; l = z[m:]
        return ls

    def solve(self, fac, mask, vals):
        """initial solve of the equality-constrained problem (pinned = vals)."""
        p = self.p; Hn, m, n = p.H, p.m, p.n
        d = [p.D[t] + (p.B[t] * (mask[t] != 0)[None, :]) @ vals[t] for t in range(Hn)]
        rho = []
        for t in range(Hn):
            free = mask[t] == 0
            Rt = fac[t][4]
            # cost (u-ub)'R(u-ub) restricted to free with pinned fixed: linear term h_F = R_FF ub_F - R_F,fix (b - ub_fix)
            ufix = np.where(free, 0.0, vals[t] - p.ub[t])
            h = (p.R[t] @ p.ub[t] - p.R[t] @ ufix) * free
            rho.append(np.linalg.solve(Rt.T, h) * free)
        kap = [self.Qh[t] @ p.r[t] for t in range(Hn)]
        ls = self.backward(fac, mask, self.Qh[Hn] @ p.r[Hn], d, rho, kap)
        X = np.zeros((Hn + 1, n)); U = np.zeros((Hn, m)); X[0] = p.x0
        for t in range(Hn):
            Qm, S12, Kh, Lp, Rt = fac[t]
            u = np.linalg.solve(S12, ls[t] - Kh @ X[t])
            U[t] = np.where(mask[t] == 0, u, vals[t])
            X[t + 1] = p.A[t] @ X[t] + p.B[t] @ U[t] + p.D[t]
        return X, U

    def refine(self, fac, mask, X, U, g):
        p = self.p; Hn, m, n = p.H, p.m, p.n
        z = [np.zeros(n)] * Hn
        rho = [np.linalg.solve(fac[t][4].T, -0.5 * g[t] * (mask[t] == 0)) * (mask[t] == 0) for t in range(Hn)]
        ls = self.backward(fac, mask, np.zeros(n), z, rho, z)
        Xn = np.zeros_like(X); Un = np.zeros_like(U); Xn[0] = X[0]
        for t in range(Hn):
            Qm, S12, Kh, Lp, Rt = fac[t]
            du = np.linalg.solve(S12, ls[t] - Kh @ (Xn[t] - X[t]))
            Un[t] = np.where(mask[t] == 0, U[t] + du, U[t])
            Xn[t + 1] = p.A[t] @ Xn[t] + p.B[t] @ Un[t] + p.D[t]
        return Xn, Un

def qp_robust(x_init, X_bm, U_bm, Q_ls, R_ls, A_ls, B_ls, D_ls, u_prev=None, sat=None, du=None, warm=None, n_ref=2, log=None):
    prob = rs._SparseQP(np.asarray(x_init).reshape(-1), X_bm, U_bm, Q_ls, R_ls, A_ls, B_ls, D_ls)
    lo, hi = rs.qp_bounds(U_bm, u_prev, sat, du); lo, hi = lo.T.copy(), hi.T.copy()
    rb = Robust(prob, lo, hi)
    Hn, m = prob.H, prob.m
    mask = np.zeros((Hn, m), dtype=int) if warm is None else warm.copy()
    for rnd in range(40):
        vals = np.where(mask == 1, lo, np.where(mask == 2, hi, 0.0))
        fac = rb.factor(mask)
        X, U = rb.solve(fac, mask, vals)
        dUs = []
        for it in range(n_ref):
            g = prob.gradient(X, U)
            Xn, Un = rb.refine(fac, mask, X, U, g)
            dUs.append(np.abs(Un - U).max())
            X, U = Xn, Un
        g = prob.gradient(X, U)
        gs = max(1.0, np.abs(g).max())
        nm = mask.copy()
        free = mask == 0
        nm[free & (U < lo - 1e-12)] = 1
        nm[free & (U > hi + 1e-12)] = 2
        nm[(mask == 1) & (g < -1e-10 * gs)] = 0
        nm[(mask == 2) & (g > 1e-10 * gs)] = 0
        if log is not None:
            log.append((rnd, int((mask != 0).sum()), int((nm != mask).sum()), dUs))
        if (nm == mask).all():
            break
        mask = nm
    c = X_bm.shape[0]
    Xc = (X[:, :c] + 1j * X[:, c:]).T
    return Xc, U.T.copy(), prob.cost(X, U), dict(mask=mask, rounds=rnd + 1, dU=dUs)

if __name__ == '__main__':
    import pickle
    Hh = sys.argv[1]
    cap = pickle.load(open('/root/repo/tools/analysis/h%s_qps.pkl' % Hh, 'rb'))
    for qi, q in enumerate(cap):
        log = []
        out = qp_robust(*q['args'], log=log)
        print('QP %d rounds %d |U-Uo| %.3e first col %.3e  last dU %s' % (qi, out[3]['rounds'], np.abs(out[1] - q['U']).max(), np.abs(out[1][:, 0] - q['U'][:, 0]).max(), ['%.1e' % x for x in out[3]['dU']]))
