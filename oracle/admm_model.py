"""numpy model of the DEVICE algorithm for the horizon QP -- TEST/DESIGN INFRASTRUCTURE ONLY.

This is not the reference's algorithm (that is OSQP behind cvxpy, optimize.py:59) and not the exact oracle
(oracle/restate.py:qp_exact).  It is a line-by-line numpy statement of what the CUDA kernel in
mpc4quantum_b200/csrc does (ADMM on the control box with a Riccati inner solve, then an active-set polish with a
KKT certificate), used to choose rho / iteration limits on the CPU and to debug the kernel.
"""
import numpy as np

from .restate import realify_op, realify_vec, qp_bounds


def _riccati_factor(Ar, Br, Dl, Qb, Qfb, R, rho_half, mask=None, bfix=None):
    """Backward matrix sweep.  Returns per-stage K, Sinv, d = P_{t+1} Delta'_t, Btil, Dtil and the P list."""
    H = len(Ar)
    n, m = Br[0].shape
    P = Qfb.copy()
    K = [None] * H
    Sinv = [None] * H
    dvec = [None] * H
    Bt_ls = [None] * H
    Dt_ls = [None] * H
    for t in reversed(range(H)):
        B = Br[t].copy()
        D = Dl[t].copy()
        Rt = R.copy() + rho_half * np.eye(m)
        if mask is not None:
            for i in range(m):
                if mask[t, i]:
                    D = D + B[:, i] * bfix[t, i]
                    B[:, i] = 0
                    Rt[i, :] = 0
                    Rt[:, i] = 0
                    Rt[i, i] = 1
        W = P @ np.hstack([Ar[t], B])
        T = np.hstack([Ar[t], B]).T @ W
        S = Rt + T[n:, n:]
        Si = np.linalg.inv(S)
        K[t] = Si @ T[n:, :n]
        Sinv[t] = Si
        dvec[t] = P @ D
        Bt_ls[t] = B
        Dt_ls[t] = D
        P = Qb + T[:n, :n] - T[n:, :n].T @ K[t]
        P = 0.5 * (P + P.T)
    return K, Sinv, dvec, Bt_ls, Dt_ls


def _sweeps(x0, Ar, Bt, Dt, K, Sinv, dvec, qlin, qlin_f, h):
    """Backward vector sweep + forward rollout.  h[t] = linear control term; returns X [H+1,n], U [H,m]."""
    H = len(Ar)
    n, m = Bt[0].shape
    p = qlin_f.copy()
    kk = np.zeros((H, m))
    for t in reversed(range(H)):
        v = dvec[t] + p
        g = Bt[t].T @ v - h[t]
        kk[t] = Sinv[t] @ g
        p = qlin[t] + Ar[t].T @ v - K[t].T @ g
    X = np.zeros((H + 1, n))
    U = np.zeros((H, m))
    X[0] = x0
    for t in range(H):
        U[t] = -K[t] @ X[t] - kk[t]
        X[t + 1] = Ar[t] @ X[t] + Bt[t] @ U[t] + Dt[t]
    return X, U


def qp_admm(x_init, X_bm, U_bm, Q_ls, R_ls, A_ls, B_ls, Delta_ls, u_prev=None, sat=None, du=None,
            rho=None, alpha=1.6, eps=1e-4, max_iter=400, polish=True, warm=None, stats=None, max_polish=8):
    m, H = U_bm.shape
    c = X_bm.shape[0]
    n = 2 * c
    Ar = [realify_op(a) for a in A_ls]
    Br = [np.vstack([b.real, b.imag]) for b in B_ls]
    Dl = [realify_vec(np.asarray(d).reshape(-1)) for d in Delta_ls]
    Qb = realify_op(Q_ls[0]); Qb = 0.5 * (Qb + Qb.T)
    Qfb = realify_op(Q_ls[-1]); Qfb = 0.5 * (Qfb + Qfb.T)
    R = np.real(np.asarray(R_ls[0], dtype=complex)); R = 0.5 * (R + R.T)
    r = np.array([realify_vec(X_bm[:, t]) for t in range(H + 1)])
    ub = np.real(U_bm.T)                  # [H, m]
    qlin = [-Qb @ r[t] for t in range(H)]
    qlin_f = -Qfb @ r[H]
    lo, hi = qp_bounds(U_bm, u_prev, sat, du)
    lo, hi = lo.T, hi.T                   # [H, m]
    x0 = realify_vec(np.asarray(x_init).reshape(-1))
    if rho is None:
        rho = 0.1
    K, Sinv, dvec, Bt, Dt = _riccati_factor(Ar, Br, Dl, Qb, Qfb, R, 0.5 * rho)
    if warm is not None:
        z, y = warm[0].copy(), warm[1].copy()
        z = np.clip(z, lo, hi)
    else:
        z = np.clip(np.zeros((H, m)), lo, hi)
        y = np.zeros((H, m))
    Ru = ub @ R.T
    it = 0
    n_pol = 0
    done = False
    X = U = None
    while not done:
        # ---- ADMM block
        for _ in range(max_iter):
            h = Ru + 0.5 * rho * (z - y)
            X, U = _sweeps(x0, Ar, Bt, Dt, K, Sinv, dvec, qlin, qlin_f, h)
            uh = alpha * U + (1 - alpha) * z
            z_new = np.clip(uh + y, lo, hi)
            y = y + uh - z_new
            r_prim = np.abs(U - z_new).max()
            r_dual = rho * np.abs(z_new - z).max()
            z = z_new
            it += 1
            if r_prim < eps and r_dual < eps:
                break
        if not polish:
            U = z
            break
        # ---- polish: primal-dual active set from the ADMM estimate
        at_lo = (z <= lo) & (y < 0)
        at_hi = (z >= hi) & (y > 0)
        ok = False
        for _ in range(max_polish):
            n_pol += 1
            mask = at_lo | at_hi
            bfix = np.where(at_lo, lo, np.where(at_hi, hi, 0.0))
            Kp, Sp, dp, Btp, Dtp = _riccati_factor(Ar, Br, Dl, Qb, Qfb, R, 0.0, mask, bfix)
            # h_F = R_FF ub_F - R_F,fix (b - ub_fix); fixed rows get 0 so that the Riccati control comes out 0
            hp = np.zeros((H, m))
            for t in range(H):
                fr = ~mask[t]
                fx = mask[t]
                hp[t, fr] = R[np.ix_(fr, fr)] @ ub[t, fr] - R[np.ix_(fr, fx)] @ (bfix[t, fx] - ub[t, fx])
            Xp, Up = _sweeps(x0, Ar, Btp, Dtp, Kp, Sp, dp, qlin, qlin_f, hp)
            Up = np.where(mask, bfix, Up)
            # adjoint gradient certificate
            lam = 2 * Qfb @ (Xp[H] - r[H])
            grad = np.zeros((H, m))
            for t in reversed(range(H)):
                grad[t] = 2 * R @ (Up[t] - ub[t]) + Br[t].T @ lam
                lam = 2 * Qb @ (Xp[t] - r[t]) + Ar[t].T @ lam
            gs = max(1.0, np.abs(grad).max())
            viol_lo = (~mask) & (Up < lo - 1e-12)
            viol_hi = (~mask) & (Up > hi + 1e-12)
            rel_lo = at_lo & (grad < -1e-10 * gs)
            rel_hi = at_hi & (grad > 1e-10 * gs)
            if not (viol_lo.any() or viol_hi.any() or rel_lo.any() or rel_hi.any()):
                ok = True
                break
            at_lo = (at_lo & ~rel_lo) | viol_lo
            at_hi = (at_hi & ~rel_hi) | viol_hi
        if ok:
            X, U = Xp, Up
            # refresh the ADMM state so the next warm start carries the right active set
            z = np.clip(Up, lo, hi)
            done = True
        else:
            eps = eps * 0.1
            if eps < 1e-10:
                U = z
                done = True
    if stats is not None:
        stats['iters'] = stats.get('iters', 0) + it
        stats['polish'] = stats.get('polish', 0) + n_pol
        stats['solves'] = stats.get('solves', 0) + 1
        stats['warm'] = (z, y)
    Xc = (X[:, :c] + 1j * X[:, c:]).T
    # objective
    obj = 0.0
    for t in range(H + 1):
        e = X[t] - r[t]
        obj += e @ (Qfb if t == H else Qb) @ e
    for t in range(H):
        e = U[t] - ub[t]
        obj += e @ R @ e
    return Xc, U.T.copy(), float(obj), None
