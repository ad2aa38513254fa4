"""numpy model of the DEVICE's pivoted stage-wise KKT solve and its interior-point working set -- TEST/DESIGN
INFRASTRUCTURE ONLY.

Not the reference's algorithm (OSQP behind cvxpy, optimize.py:59) and not the exact oracle (oracle/restate.py:qp_exact):
a numpy statement of what mpc4quantum_b200/csrc/m4q_kkt.cuh does, used to choose its parameters on the CPU
(tools/analysis/ipm_proto.py, ipm_variants.py) and to pin the algorithm itself in tests/test_oracle.py.

The KKT system of the working set, unknowns ordered stage by stage -- block s = [u_{s-1} | x_s | lam_s] -- and rows
[stationarity or pin | dynamics | costate]:
    2 R_t (u_t - ub_t) + B_t^T lam_{t+1} = 0      (free controls; pinned: u_t[i] = bound;
                                                   interior point: + Sigma_t u_t on the left, + shift_t on the right)
    x_{t+1} - A_t x_t - B_t u_t = D_t
    lam_t - 2 Q_t x_t - A_t^T lam_{t+1} = -2 Q_t r_t
is banded; a dense LU with row partial pivoting makes the same pivot choices as the device's block-column elimination.
"""
import numpy as np
import scipy.linalg as sl


def build(prob, fixed, vals, sig=None, shift=None):
    """The banded KKT matrix and right-hand side for a restate._SparseQP."""
    n, m, H = prob.n, prob.m, prob.H
    W = 2 * n + m
    K = np.zeros((W * H, W * H))
    b = np.zeros(W * H)
    for t in range(1, H + 1):
        c = (t - 1) * W
        r = c
        for i in range(m):
            if fixed[t - 1, i]:
                K[r + i, c + i] = 1.0
                b[r + i] = vals[t - 1, i]
            else:
                K[r + i, c:c + m] = 2 * prob.R[t - 1][i]
                K[r + i, c + m + n:c + W] = prob.B[t - 1][:, i]
                b[r + i] = 2 * prob.R[t - 1][i] @ prob.ub[t - 1]
                if sig is not None:
                    K[r + i, c + i] += sig[t - 1, i]
                    b[r + i] += shift[t - 1, i]
        r += m
        K[r:r + n, c:c + m] = -prob.B[t - 1]
        K[r:r + n, c + m:c + m + n] = np.eye(n)
        b[r:r + n] = prob.D[t - 1]
        if t > 1:
            K[r:r + n, c - W + m:c - W + m + n] = -prob.A[t - 1]
        else:
            b[r:r + n] += prob.A[0] @ prob.x0
        r += n
        K[r:r + n, c + m:c + m + n] = -2 * prob.Q[t]
        K[r:r + n, c + m + n:c + W] = np.eye(n)
        b[r:r + n] = -2 * prob.Q[t] @ prob.r[t]
        if t < H:
            K[r:r + n, c + W + m + n:c + 2 * W] = -prob.A[t].T
    return K, b


def solve(prob, fixed, vals, sig=None, shift=None, refine=0):
    """(U [H, m], gradient of the objective w.r.t. U from the solve's own costates)."""
    n, m, H = prob.n, prob.m, prob.H
    K, b = build(prob, fixed, vals, sig, shift)
    lu = sl.lu_factor(K)
    z = sl.lu_solve(lu, b)
    for _ in range(refine):
        z = z + sl.lu_solve(lu, b - K @ z)
    z = z.reshape(H, -1)
    U = z[:, :m].copy()
    lam = z[:, m + n:]
    g = np.array([2 * prob.R[t] @ (U[t] - prob.ub[t]) + prob.B[t].T @ lam[t] for t in range(H)])
    return U, g


def interior_point(prob, lo, hi, sigma=0.1, tau=0.995, max_it=80, mu_tol=1e-8, state=None):
    """Primal-dual interior point on the box: every iteration is ONE KKT solve with all controls free and the barrier
    terms as a diagonal shift of R.  Returns (u, z_lo, z_hi, solves)."""
    H, m = lo.shape
    nofix = np.zeros((H, m), bool)
    zeros = np.zeros((H, m))
    if state is None:
        u, zl, zu = 0.5 * (lo + hi), np.ones((H, m)), np.ones((H, m))
    else:
        u, zl, zu = state
    n_solve = 0

    def maxstep(v, dv):
        neg = dv < 0
        return min(1.0, (tau * (-v[neg] / dv[neg])).min()) if neg.any() else 1.0
    for _ in range(max_it):
        s_lo, s_hi = u - lo, hi - u
        mu = (np.sum(s_lo * zl) + np.sum(s_hi * zu)) / (2 * H * m)
        if mu < mu_tol:
            break
        Sig = zl / s_lo + zu / s_hi
        up, _ = solve(prob, nofix, zeros, Sig, Sig * u + sigma * mu * (1 / s_lo - 1 / s_hi))
        n_solve += 1
        du = up - u
        dzl = sigma * mu / s_lo - zl - (zl / s_lo) * du
        dzu = sigma * mu / s_hi - zu + (zu / s_hi) * du
        ap = min(maxstep(s_lo, du), maxstep(s_hi, -du))
        ad = min(maxstep(zl, dzl), maxstep(zu, dzu))
        u, zl, zu = u + ap * du, zl + ad * dzl, zu + ad * dzu
    return u, zl, zu, n_solve


def polish(prob, lo, hi, u, zl, zu, max_rounds=16):
    """Working set = bounds whose slack is smaller than their multiplier; primal-dual rounds with refined solves and
    the hysteresis for weakly active bounds.  Returns (U or None, rounds)."""
    H, m = lo.shape
    mask = np.where((u - lo) < zl, 1, np.where((hi - u) < zu, 2, 0))
    flips = np.zeros((H, m), int)
    for rnd in range(max_rounds):
        fixed = mask != 0
        vals = np.where(mask == 1, lo, np.where(mask == 2, hi, 0.0))
        U, g = solve(prob, fixed, vals, refine=1)
        U = np.where(fixed, vals, U)
        gs = max(1.0, np.abs(g).max())
        v_lo = ~fixed & (U < lo - 1e-12)
        v_hi = ~fixed & (U > hi + 1e-12)
        gn = np.where(mask == 1, -g, np.where(mask == 2, g, 0.0))
        rel = fixed & (gn > 1e-10 * gs) & ((flips < 2) | (gn > 1e-5 * gs))
        if not (v_lo.any() or v_hi.any() or rel.any()):
            return U, rnd + 1
        mask = mask.copy()
        mask[v_lo] = 1
        mask[v_hi] = 2
        mask[rel] = 0
        flips[rel] += 1
    return None, max_rounds


def qp_kkt_ipm(prob, lo, hi):
    """The device's sequence: interior point to mu = 1e-8, polish; if that does not settle, on to 1e-11 and again.
    Returns (U [H, m] or None, stats)."""
    state, solves = None, 0
    for mu_tol in (1e-8, 1e-11):
        u, zl, zu, n = interior_point(prob, lo, hi, mu_tol=mu_tol, state=state)
        solves += n
        U, rounds = polish(prob, lo, hi, u, zl, zu)
        if U is not None:
            return U, dict(ipm_solves=solves, polish_rounds=rounds)
        state = (u, zl, zu)
    return None, dict(ipm_solves=solves, polish_rounds=None)
