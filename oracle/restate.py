"""numpy/scipy restatement of the reference's MPC hot path -- TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines whose semantics it restates
(paths relative to /root/reference).  Written from the mathematics, not
transliterated; checked against the reference itself (run through
oracle/refshim.py) by oracle/make_golden.py and tests/test_oracle.py.

Leaves restated from third-party packages that are absent here:

* ``qp_exact``      <- optimize.py:12-60 (cvxpy 1.1.13 -> OSQP 0.6.2.post0).
  The QP is strictly convex in U (R > 0), so the optimum is unique; it is
  found by an active-set iteration on the sparse (X, U) form and certified by KKT residuals
  (``qp_kkt``).
* ``expm_plant_segment`` / ``ExpmPlant`` <- experiment.py:202-212 (qutip 4.6.2
  mesolve with a piecewise-constant control from mpc.py:256-260): exact
  ``rho <- U rho U^dagger`` with ``U = expm(-i H(u) dt)``.
"""
import itertools
import math

import numpy as np
from scipy.linalg import expm


# ----------------------------------------------------------------------------
# Monomial library (linearize.py:92-164)
# ----------------------------------------------------------------------------
def power_table(order, dim_u):
    """All control monomials of total degree <= order, in the reference's order.

    linearize.py:92-116 enumerates them by stars-and-bars; the resulting order is
    "ascending lexicographic on the reversed exponent tuple" (last control most
    significant), row 0 being the constant.  Returns int array [p+1, dim_u].
    """
    rows = [pw for pw in itertools.product(range(order + 1), repeat=dim_u) if sum(pw) <= order]
    rows.sort(key=lambda pw: pw[::-1])
    return np.array(rows, dtype=int).reshape(len(rows), dim_u)


def monomials(powers, u):
    """phi_p(u) = prod_i u_i**powers[p, i]; negative exponent -> 0 (linearize.py:123-128)."""
    u = np.asarray(u, dtype=float).reshape(-1)
    out = np.ones(len(powers))
    for k, pw in enumerate(powers):
        for i, e in enumerate(pw):
            out[k] *= 0.0 if e < 0 else u[i] ** e
    return out


# ----------------------------------------------------------------------------
# Model construction (vectorize.py)
# ----------------------------------------------------------------------------
def liouvillian(H):
    """Matrix of rho -> -i[H, rho] on ROW-major vec(rho)  (vectorize.py:52-75 in the |a><b| basis)."""
    H = np.asarray(H, dtype=complex)
    eye = np.eye(H.shape[0])
    return -1j * (np.kron(H, eye) - np.kron(eye, H.T))


def liouvillian_in_basis(H, basis):
    """General-basis statement of vectorize.py:52-75.

    L[k, j] = -i * sum_{i != k} tr(H^dag s_i) * tr([s_i, s_k]^dag s_j).
    """
    H = np.asarray(H, dtype=complex)
    dim_m = len(basis)
    h = np.array([np.trace(H.conj().T @ s) for s in basis])
    L = np.zeros((dim_m, dim_m), dtype=complex)
    for k in range(dim_m):
        for i in range(dim_m):
            if i == k:
                continue
            comm = basis[i] @ basis[k] - basis[k] @ basis[i]
            for j in range(dim_m):
                L[k, j] += -1j * h[i] * np.trace(comm.conj().T @ basis[j])
    return L


def taylor_discretize(L_list, dt, order):
    """Order-`order` Taylor blocks of exp((L0 + sum u_i L_i) dt) grouped by control monomial.

    vectorize.py:8-49.  Returns [c, c*(p+1)] = hstack of blocks in power_table order.
    Recurrence on word length instead of the reference's enumeration of all words.
    """
    c = L_list[0].shape[0]
    m = len(L_list) - 1
    table = power_table(order, m)
    index = {tuple(pw): k for k, pw in enumerate(table)}
    blocks = np.zeros((len(table), c, c), dtype=complex)
    level = {tuple([0] * m): np.eye(c, dtype=complex)}
    blocks[0] += level[tuple([0] * m)]
    for k in range(1, order + 1):
        nxt = {}
        for pw, W in level.items():
            for j, L in enumerate(L_list):
                key = pw if j == 0 else tuple(e + (1 if i == j - 1 else 0) for i, e in enumerate(pw))
                nxt[key] = nxt.get(key, 0) + W @ L
        level = nxt
        for pw, W in level.items():
            blocks[index[pw]] += (dt ** k / math.factorial(k)) * W
    return np.hstack(list(blocks))


# ----------------------------------------------------------------------------
# Local linearisation (linearize.py:37-70)
# ----------------------------------------------------------------------------
class BilinearModel:
    """x+ = A x + N (phi(u) (x) x)   (linearize.py:13-35; model.py:95-103 slices [A | N])."""

    def __init__(self, A_full, dim_u, order):
        A_full = np.asarray(A_full, dtype=complex)
        self.c = A_full.shape[0]
        self.m = dim_u
        self.order = order
        self.table = power_table(order, dim_u)[1:]
        self.p = len(self.table)
        if A_full.shape[1] != self.c * (self.p + 1):
            raise ValueError('Dimension mismatch when wrapping a model operator.')
        self.A = A_full[:, :self.c]
        self.N = A_full[:, self.c:].reshape(self.c, self.p, self.c).transpose(1, 0, 2)  # [p, c, c]

    def step(self, x, u):
        """f(x,u)  (linearize.py:37-41; also model.py:81-93 via mpc.py:264-267)."""
        phi = monomials(self.table, u)
        return self.A @ x + np.einsum('p,pij,j->i', phi, self.N, x)

    def jac_x(self, u):
        """df/dx = A + sum_p phi_p N_p  (linearize.py:43-48)."""
        return self.A + np.einsum('p,pij->ij', monomials(self.table, u), self.N)

    def jac_u(self, x, u):
        """df/du[:, i] = sum_p (N_p x) * e_{p,i} * u**(e_p - 1_i)  (linearize.py:50-59, 143-164)."""
        Nx = np.einsum('pij,j->pi', self.N, x)      # [p, c]
        B = np.zeros((self.c, self.m), dtype=complex)
        for i in range(self.m):
            lowered = self.table.copy()
            lowered[:, i] -= 1
            w = self.table[:, i] * monomials(lowered, u)
            B[:, i] = w @ Nx
        return B

    def along(self, Xg, Ug, H):
        """(A_t, B_t, Delta_t) for t < H  (linearize.py:61-70)."""
        A_ls, B_ls, D_ls = [], [], []
        for t in range(H):
            At = self.jac_x(Ug[:, t])
            Bt = self.jac_u(Xg[:, t], Ug[:, t])
            D_ls.append(self.step(Xg[:, t], Ug[:, t]) - At @ Xg[:, t] - Bt @ Ug[:, t])
            A_ls.append(At)
            B_ls.append(Bt)
        return A_ls, B_ls, D_ls


# ----------------------------------------------------------------------------
# Horizon QP (optimize.py:12-60)
# ----------------------------------------------------------------------------
class ExactModel:
    """Exact-discretisation model (an extension, no reference counterpart):
    x+ = expm(G(u) dt) x, G(u) = L_0 + sum u_i L_i.  A_t = expm(G dt); B_t[:, i] = L_expm(G dt, L_i dt) x_t with scipy's
    Frechet derivative; Delta_t = f - A_t x_t - B_t u_t = -B_t u_t.  Same interface as BilinearModel."""

    def __init__(self, generators, dt):
        self.gen = [np.asarray(g, dtype=complex) for g in generators]
        self.dt = dt
        self.dim_u = len(self.gen) - 1

    def _G(self, u):
        G = np.array(self.gen[0])
        for Li, ui in zip(self.gen[1:], np.asarray(u).reshape(-1)):
            G = G + ui * Li
        return G * self.dt

    def step(self, x, u):
        return expm(self._G(u)) @ x

    def along(self, Xg, Ug, H):
        from scipy.linalg import expm_frechet
        A_ls, B_ls, D_ls = [], [], []
        for t in range(H):
            G = self._G(Ug[:, t])
            A = expm(G)
            B = np.stack([expm_frechet(G, Li * self.dt, compute_expm=False) @ Xg[:, t] for Li in self.gen[1:]], axis=1)
            A_ls.append(A)
            B_ls.append(B)
            D_ls.append(-B @ Ug[:, t])
        return A_ls, B_ls, D_ls


def realify_vec(z):
    return np.concatenate([np.real(z), np.imag(z)])     # mpc.py:87-89


def realify_op(P):
    P = np.asarray(P, dtype=complex)
    return np.block([[P.real, -P.imag], [P.imag, P.real]])   # mpc.py:92-93


def qp_bounds(U_bm, u_prev, sat, du):
    m, H = U_bm.shape
    lo = -sat * np.ones((m, H))
    hi = sat * np.ones((m, H))
    if u_prev is not None and du is not None:           # optimize.py:29-30
        up = np.asarray(u_prev, dtype=float).reshape(-1)
        lo[:, 0] = np.maximum(lo[:, 0], up - du)
        hi[:, 0] = np.minimum(hi[:, 0], up + du)
    return lo, hi


def qp_exact(x_init, X_bm, U_bm, Q_ls, R_ls, A_ls, B_ls, Delta_ls, u_prev=None, sat=None, du=None,
             verbose=False):
    """Exact optimum of the QP stated at optimize.py:12-60; same signature and returns.

    min  sum_t Re[(x_t-r_t)^H Q_t (x_t-r_t)] + (u_t-ub_t)^T R_t (u_t-ub_t) + terminal   (no 1/2; :34-35,:54)
    s.t. x_0 = x_init; x_{t+1} = Delta_t + A_t x_t + B_t u_t (:41); |u_t| <= sat (:43);
         |u_0 - u_prev| <= du (:29-30).
    Returns (X complex [c,H+1], U real [m,H], obj_val, info) -- ``info`` replaces the cvxpy problem and
    carries the KKT certificate (``info['kkt']`` = (stationarity/complementarity residual, infeasibility)).

    Method: the problem is kept in its sparse (X, U) form -- the condensed Hessian is useless at long horizons
    because the Taylor model is not norm preserving (cond > 1e10 at H = 50) -- and solved by an active-set
    iteration whose equality-constrained subproblems are sparse KKT solves; the answer is certified by the
    adjoint-gradient KKT residual, so its correctness does not rest on the iteration that found it.
    """
    prob = _SparseQP(np.asarray(x_init).reshape(-1), X_bm, U_bm, Q_ls, R_ls, A_ls, B_ls, Delta_ls)
    lo, hi = qp_bounds(U_bm, u_prev, sat, du)
    lo, hi = lo.T.copy(), hi.T.copy()          # [H, m]
    Xr, U, grad = _active_set(prob, lo, hi)
    c = X_bm.shape[0]
    X = (Xr[:, :c] + 1j * Xr[:, c:]).T
    info = {'prob': prob, 'lo': lo, 'hi': hi, 'grad': grad, 'kkt': qp_kkt_from_grad(U, grad, lo, hi)}
    return X, U.T.copy(), prob.cost(Xr, U), info


class _SparseQP:
    """Realified stage data of the QP and its equality-constrained (fixed-control) solves."""

    def __init__(self, x_init, X_bm, U_bm, Q_ls, R_ls, A_ls, B_ls, D_ls):
        self.m, self.H = U_bm.shape
        self.c = X_bm.shape[0]
        self.n = 2 * self.c
        H = self.H
        self.x0 = realify_vec(x_init)
        self.A = [realify_op(a) for a in A_ls]
        self.B = [np.vstack([np.real(b), np.imag(b)]) for b in B_ls]
        self.D = [realify_vec(np.asarray(d).reshape(-1)) for d in D_ls]
        self.Q = [0.5 * (realify_op(q) + realify_op(q).T) for q in Q_ls]
        self.R = [0.5 * (np.real(np.asarray(r, dtype=complex)) + np.real(np.asarray(r, dtype=complex)).T)
                  for r in R_ls]
        self.r = [realify_vec(X_bm[:, t]) for t in range(H + 1)]
        self.ub = [np.real(U_bm[:, t]).astype(float) for t in range(H)]

    def cost(self, X, U):
        val = 0.0
        for t in range(self.H + 1):
            e = X[t] - self.r[t]
            val += e @ self.Q[t] @ e
        for t in range(self.H):
            e = U[t] - self.ub[t]
            val += e @ self.R[t] @ e
        return float(val)

    def rollout(self, U):
        X = np.zeros((self.H + 1, self.n))
        X[0] = self.x0
        for t in range(self.H):
            X[t + 1] = self.A[t] @ X[t] + self.B[t] @ U[t] + self.D[t]
        return X

    def gradient(self, X, U):
        """d cost / d U through the dynamics (adjoint recursion); independent of how (X, U) was found."""
        lam = 2 * self.Q[self.H] @ (X[self.H] - self.r[self.H])
        g = np.zeros((self.H, self.m))
        for t in reversed(range(self.H)):
            g[t] = 2 * self.R[t] @ (U[t] - self.ub[t]) + self.B[t].T @ lam
            lam = 2 * self.Q[t] @ (X[t] - self.r[t]) + self.A[t].T @ lam
        return g

    def _base(self):
        """Constant blocks of the KKT matrix: 2C = blockdiag(2Q_t, 2R_t), E (dynamics), and their right-hand sides."""
        if getattr(self, '_cache', None) is None:
            import scipy.sparse as sp
            n, m, H = self.n, self.m, self.H
            C2 = sp.block_diag([2 * q for q in self.Q] + [2 * r for r in self.R], format='csc')
            lin = np.concatenate([2 * q @ r for q, r in zip(self.Q, self.r)] +
                                 [2 * r @ u for r, u in zip(self.R, self.ub)])
            # E = [I_x - shift(A) | -shift(B)]: row block t+1 has I at x_{t+1}, -A_t at x_t, -B_t at u_t
            Ablk = sp.block_diag(self.A, format='csc')                 # [nH, nH]
            Bblk = sp.block_diag(self.B, format='csc')                 # [nH, mH]
            top = sp.csc_matrix((n, n * H))
            shiftA = sp.hstack([sp.vstack([top, Ablk]), sp.csc_matrix((n * (H + 1), n))], format='csc')
            shiftB = sp.vstack([sp.csc_matrix((n, m * H)), Bblk], format='csc')
            E = sp.hstack([sp.eye(n * (H + 1), format='csc') - shiftA, -shiftB], format='csc')
            e = np.concatenate([self.x0] + list(self.D))
            self._cache = (C2, lin, E, e)
        return self._cache

    def solve_fixed(self, fixed, values, want_grad=False):
        """Minimise with U[fixed] = values[fixed] and the rest free: one sparse KKT solve.

        KKT matrix [[2C, E^T, G^T], [E, 0, 0], [G, 0, 0]] with G the selector of the fixed controls.
        """
        import scipy.sparse as sp
        from scipy.sparse.linalg import splu
        n, m, H = self.n, self.m, self.H
        C2, lin, E, e = self._base()
        nx = n * (H + 1)
        nv = nx + m * H
        fidx = np.flatnonzero(fixed.reshape(-1))
        nfix = len(fidx)
        G = sp.csc_matrix((np.ones(nfix), (np.arange(nfix), nx + fidx)), shape=(nfix, nv))
        K = sp.bmat([[C2, E.T, G.T], [E, None, None], [G, None, None]], format='csc')
        rhs = np.concatenate([lin, e, values.reshape(-1)[fidx]])
        lu = splu(K)
        sol = lu.solve(rhs)
        sol += lu.solve(rhs - K @ sol)          # one step of iterative refinement
        U = sol[nx:nv].reshape(H, m).copy()
        U.reshape(-1)[fidx] = values.reshape(-1)[fidx]
        # states from the KKT solution itself (a rollout would amplify round-off by ||A||^H at long horizons)
        X = sol[:nx].reshape(H + 1, n).copy()
        if not want_grad:
            return X, U
        # d cost / d U from the solve's own costates nu (the multipliers of E z = e; row block t+1 is the dynamics of
        # stage t): 2 R (u - ub) - B^T nu_{t+1}.  Unlike gradient(), nothing is propagated through prod A_t.
        nu = sol[nv:nv + nx].reshape(H + 1, n)
        g = np.array([2 * self.R[t] @ (U[t] - self.ub[t]) - self.B[t].T @ nu[t + 1] for t in range(H)])
        return X, U, g


def _active_set(prob, lo, hi, max_rounds=40):
    """Primal-dual active-set rounds; if a working set repeats, fall back to the textbook primal method."""
    H, m = lo.shape
    X, U = prob.solve_fixed(np.zeros((H, m), dtype=bool), np.zeros((H, m)))
    at_lo = U < lo
    at_hi = U > hi
    seen = set()
    for _ in range(max_rounds):
        fixed = at_lo | at_hi
        vals = np.where(at_lo, lo, np.where(at_hi, hi, 0.0))
        X, U = prob.solve_fixed(fixed, vals)
        grad = prob.gradient(X, U)
        gs = max(1.0, float(np.abs(grad).max()))
        viol_lo = ~fixed & (U < lo - 1e-13)
        viol_hi = ~fixed & (U > hi + 1e-13)
        rel_lo = at_lo & (grad < -1e-11 * gs)
        rel_hi = at_hi & (grad > 1e-11 * gs)
        if not (viol_lo.any() or viol_hi.any() or rel_lo.any() or rel_hi.any()):
            return X, U, grad
        key = (at_lo.tobytes(), at_hi.tobytes())
        if key in seen:
            break
        seen.add(key)
        at_lo = (at_lo & ~rel_lo) | viol_lo
        at_hi = (at_hi & ~rel_hi) | viol_hi
    try:
        return _primal_active_set(prob, lo, hi, np.clip(U, lo, hi))
    except RuntimeError:
        return _active_set_kkt_multipliers(prob, lo, hi)


def _active_set_kkt_multipliers(prob, lo, hi, max_rounds=40):
    """Last resort for the order-1 model at H = 100 (||prod A_t|| ~ 1e13): the adjoint gradient used above has a noise
    floor of eps ||prod A_t||^2 and the multiplier signs cannot be read from it.  Same primal-dual rounds from the empty
    working set, with the multipliers taken from the costates of the sparse KKT solve itself (a weakly active bound
    that has been released twice stays pinned unless its multiplier is wrong beyond 1e-5 relative); if those do not
    settle either (unconstrained optima ten box widths outside the box make the rounds erratic), the textbook primal
    method, which cannot cycle, with the same multipliers."""
    H, m = lo.shape
    mask = np.zeros((H, m), dtype=int)
    flips = np.zeros((H, m), dtype=int)
    for _ in range(max_rounds):
        fixed = mask != 0
        vals = np.where(mask == 1, lo, np.where(mask == 2, hi, 0.0))
        X, U, g = prob.solve_fixed(fixed, vals, want_grad=True)
        gs = max(1.0, float(np.abs(g).max()))
        viol_lo = ~fixed & (U < lo - 1e-12)
        viol_hi = ~fixed & (U > hi + 1e-12)
        gn = np.where(mask == 1, -g, np.where(mask == 2, g, 0.0))
        rel = fixed & (gn > 1e-10 * gs) & ((flips < 2) | (gn > 1e-5 * gs))
        if not (viol_lo.any() or viol_hi.any() or rel.any()):
            return X, U, np.where(fixed, g, 0.0)
        mask = mask.copy()
        mask[viol_lo] = 1
        mask[viol_hi] = 2
        mask[rel] = 0
        flips[rel] += 1
    return _primal_active_set(prob, lo, hi, np.clip(np.zeros((H, m)), lo, hi), kkt_multipliers=True)


def _primal_active_set(prob, lo, hi, U, kkt_multipliers=False):
    """Nocedal & Wright alg. 16.3 for the box: feasible iterates, one constraint added per round.
    kkt_multipliers: multipliers from the costates of the KKT solve instead of the adjoint gradient, and every
    constraint with a wrong-signed multiplier dropped at once (the order-1 model at H = 100)."""
    H, m = lo.shape
    at_lo = U <= lo
    at_hi = (U >= hi) & ~at_lo
    for _ in range(20 * H * m + 100):
        fixed = at_lo | at_hi
        vals = np.where(at_lo, lo, np.where(at_hi, hi, 0.0))
        if kkt_multipliers:
            Xe, Ue, grad = prob.solve_fixed(fixed, vals, want_grad=True)
        else:
            Xe, Ue = prob.solve_fixed(fixed, vals)
        step = Ue - U
        if np.abs(step).max() <= 1e-11 * max(1.0, float(np.abs(U).max())):
            if not kkt_multipliers:
                grad = prob.gradient(Xe, Ue)
            gs = max(1.0, float(np.abs(grad).max()))
            w = np.where(at_lo, -grad, np.where(at_hi, grad, 0.0))
            if w.max() <= 1e-9 * gs:
                return Xe, Ue, (np.where(fixed, grad, 0.0) if kkt_multipliers else grad)
            if kkt_multipliers:
                drop = w > 1e-9 * gs
                at_lo[drop] = False
                at_hi[drop] = False
                continue
            k = np.unravel_index(np.argmax(w), w.shape)
            at_lo[k] = False
            at_hi[k] = False
            continue
        with np.errstate(divide='ignore', invalid='ignore'):
            ratio = np.where(step < 0, (lo - U) / step, np.where(step > 0, (hi - U) / step, np.inf))
        ratio = np.where(fixed, np.inf, ratio)
        k = np.unravel_index(np.argmin(ratio), ratio.shape)
        if ratio[k] >= 1.0:
            U = Ue
        else:
            U = U + ratio[k] * step
            if step[k] < 0:
                at_lo[k] = True
                U[k] = lo[k]
            else:
                at_hi[k] = True
                U[k] = hi[k]
    raise RuntimeError('oracle QP active set did not settle')


def qp_kkt_from_grad(U, grad, lo, hi, tol=1e-9):
    on_lo = U <= lo + tol
    on_hi = U >= hi - tol
    r = np.where(on_lo & ~on_hi, np.minimum(grad, 0.0),
                 np.where(on_hi & ~on_lo, np.maximum(grad, 0.0),
                          np.where(on_lo & on_hi, 0.0, grad)))
    infeas = max(0.0, float((lo - U).max()), float((U - hi).max()))
    return float(np.abs(r).max()), infeas


def qp_kkt(info, U):
    """KKT certificate of ANY candidate U [m, H] for the problem in ``info`` (from qp_exact)."""
    Ut = np.asarray(U, dtype=float).T
    X = info['prob'].rollout(Ut)
    return qp_kkt_from_grad(Ut, info['prob'].gradient(X, Ut), info['lo'], info['hi'])


# ----------------------------------------------------------------------------
# Line search (mpc.py:101-125) -- including the reference's ordering quirk
# ----------------------------------------------------------------------------
def line_search(Q_ls, R_ls, X_ref, U_ref, Xg, Ug, Xo, Uo):
    """alpha and ||alpha*DZ|| exactly as mpc.py:101-122 computes them.

    The reference pairs a TIME-major block-diagonal metric (:103-104) with
    STATE-major flattened vectors (:105-107).  That pairing is reproduced here on
    purpose: M = sym(blockdiag(realify(Q_0..Q_H), realify(R_0..)))
    acting on Z = [Re X.ravel(), Im X.ravel(), Re U.ravel(), Im U.ravel()].
    """
    blocks = [realify_op(q) for q in Q_ls] + [realify_op(r) for r in R_ls]
    size = sum(b.shape[0] for b in blocks)
    M = np.zeros((size, size))
    o = 0
    for b in blocks:
        M[o:o + b.shape[0], o:o + b.shape[0]] = b
        o += b.shape[0]
    M = 0.5 * (M + M.T)

    def zvec(X, U):
        return np.concatenate([realify_vec(np.asarray(X).ravel()), realify_vec(np.asarray(U).ravel())])
    Zt, Zg, Zo = zvec(X_ref, U_ref), zvec(Xg, Ug), zvec(Xo, Uo)
    DZ = Zo - Zg
    alpha = -(M @ (Zg - Zt)) @ DZ / (DZ @ M @ DZ)
    return alpha, float(np.linalg.norm(alpha * DZ))


# ----------------------------------------------------------------------------
# Plant (experiment.py:202-212 + mpc.py:256-260) and observable maps (experiment.py:29-37, 225-306)
# ----------------------------------------------------------------------------
def expm_plant_segment(rho_vec, H0, H1_list, u, dt):
    d = H0.shape[0]
    Ham = np.array(H0, dtype=complex)
    for Hk, uk in zip(H1_list, np.asarray(u).reshape(-1)):
        Ham = Ham + uk * np.asarray(Hk, dtype=complex)
    U = expm(-1j * Ham * dt)
    rho = np.asarray(rho_vec, dtype=complex).reshape(d, d)
    return (U @ rho @ U.conj().T).reshape(-1)


def lift_identity(x):
    return x


def lift_coupled(rho_vec):
    """Stacked partial traces [vec(tr_B rho), vec(tr_A rho)]  (experiment.py:248-285)."""
    dAB = math.isqrt(len(rho_vec))
    dA = math.isqrt(dAB)
    r = np.asarray(rho_vec, dtype=complex).reshape(dA, dA, dA, dA)     # [a, b, a', b']
    rhoA = np.einsum('abcb->ac', r)
    rhoB = np.einsum('abad->bd', r)
    return np.concatenate([rhoA.reshape(-1), rhoB.reshape(-1)])


def proj_coupled(stacked):
    """vec(rho_A (x) rho_B)  (experiment.py:287-306)."""
    half = len(stacked) // 2
    dA = math.isqrt(half)
    return np.kron(stacked[:half].reshape(dA, dA), stacked[half:].reshape(dA, dA)).reshape(-1)


def lift_32(rho33_vec):
    """Truncate to the qubit block and trace-normalise (experiment.py:225-229; Qobj.unit() = /trace norm)."""
    blk = np.asarray(rho33_vec, dtype=complex).reshape(3, 3)[:2, :2]
    s = np.linalg.svd(blk, compute_uv=False).sum()
    return (blk / s).reshape(-1)


class ExpmPlant:
    """Duck-typed experiment for the MPC loop: lift/proj/simulate(x0, ts, us)  (experiment.py:8-49, 175-212)."""

    def __init__(self, H0, H1_list, lift=lift_identity, proj=lift_identity):
        self.H0 = np.asarray(H0, dtype=complex)
        self.H1_list = [np.asarray(h, dtype=complex) for h in H1_list]
        self.lift = lift
        self.proj = proj

    def simulate(self, x0, ts, us):
        """Piecewise-constant control: segment i uses us(ts[i]) (interp1d kind='previous', mpc.py:258)."""
        out = [np.asarray(x0, dtype=complex).reshape(-1)]
        for i in range(len(ts) - 1):
            u = us(ts[i]) if callable(us) else np.atleast_2d(us)[:, i]
            out.append(expm_plant_segment(out[-1], self.H0, self.H1_list, u, ts[i + 1] - ts[i]))
        return np.array(out).T


def lift_process(U_vec):
    """vec(U (x) U^*)  (experiment.py:357-369)."""
    n = math.isqrt(len(U_vec))
    U = np.asarray(U_vec, dtype=complex).reshape(n, n)
    return np.kron(U, U.conj()).reshape(-1)


def proj_process(P_vec):
    """A propagator equal to U up to a global phase from vec(U (x) U^*)  (experiment.py:371-388)."""
    P_vec = np.asarray(P_vec, dtype=complex)
    n = math.isqrt(math.isqrt(len(P_vec)))
    P4 = P_vec.reshape(n, n, n, n)            # [i, k, j, l] = U[i, j] conj(U[k, l])
    for blk in range(n * n):
        i, j = divmod(blk, n)
        b = P4[i, :, j, :]                    # U[i, j] conj(U)
        if np.any(b):
            return (b.conj() / np.lib.scimath.sqrt(b.reshape(-1)[blk])).reshape(-1)
    return np.zeros(n * n)


class ProcessPlant:
    """Gate-synthesis plant in process-vector coordinates (experiment.py:336-417 wired as test_NOT_gate intends:
    the loop sees vec(U (x) U^*) with identity observable maps).  One segment is P <- (V (x) V^*) P with
    V = expm(-i H(u) dt), which equals lift(V proj(P)) without fixing a phase."""
    lift = staticmethod(lift_identity)
    proj = staticmethod(lift_identity)

    def __init__(self, H0, H1_list):
        self.H0 = np.asarray(H0, dtype=complex)
        self.H1_list = [np.asarray(h, dtype=complex) for h in H1_list]

    def segment(self, P_vec, u, dt):
        Ham = np.array(self.H0, dtype=complex)
        for Hk, uk in zip(self.H1_list, np.asarray(u).reshape(-1)):
            Ham = Ham + uk * Hk
        V = expm(-1j * Ham * dt)
        n2 = V.shape[0] ** 2
        return (np.kron(V, V.conj()) @ np.asarray(P_vec, dtype=complex).reshape(n2, n2)).reshape(-1)

    def simulate(self, x0, ts, us):
        out = [np.asarray(x0, dtype=complex).reshape(-1)]
        for i in range(len(ts) - 1):
            u = us(ts[i]) if callable(us) else np.atleast_2d(us)[:, i]
            out.append(self.segment(out[-1], u, ts[i + 1] - ts[i]))
        return np.array(out).T


# ----------------------------------------------------------------------------
# The closed loop (mpc.py:128-304)
# ----------------------------------------------------------------------------
def mpc_loop(x0, dim_u, order, X_targ, U_targ, dt, horizon, n_steps, plant, A_full, Q, R, Qf, sat, du,
             max_iter=100, warm_start=True, measure_freq=1, qp=qp_exact, stats=None, exit_condition=None, model=None):
    """Restatement of the reference loop with its parity-critical behaviours:

    * guess initialised to the lifted x0 and zero controls (mpc.py:141-142)
    * u_prev = us[step-1] if step > 1 else U_ref[:, 0]            (mpc.py:185)
    * line search on steps <= 1 (warm_start) or always; stop when ||alpha DZ|| < 1e-4 (mpc.py:208-225)
    * us[step] = first column of the LAST QP solution             (mpc.py:250)
    * plant window uses [us[step], us[step-1], ...] newest first   (mpc.py:257)
    * reference windows lag: X_targ[:, step:step+H+1] is installed AFTER step `step` (mpc.py:276-277)
    Returns xs [dim, S+1], us [m, S], exit_code, and fills stats['qp_per_step'].
    """
    H, S, mf = horizon, n_steps, measure_freq
    model = model if model is not None else BilinearModel(A_full, dim_u, order)
    lx0 = np.asarray(plant.lift(np.asarray(x0, dtype=complex)), dtype=complex)
    Xg = np.tile(lx0.reshape(-1, 1), (1, H + 1))
    Ug = np.zeros((dim_u, H))
    X_ref = np.atleast_2d(X_targ[:, :H + 1])
    U_ref = np.atleast_2d(U_targ[:, :H])
    Q_ls = [Q] * H + [Qf]
    R_ls = [R] * H
    xs = [None] * (S + 1)
    us = [None] * S
    xs[0] = np.asarray(x0, dtype=complex)
    qp_counts = []
    exit_code = 0
    trace = stats.setdefault('trace', []) if (stats is not None and stats.get('want_trace')) else None
    for step in range(S):
        n_iter, done = 0, False
        if trace is not None:       # what the controller knows when step `step` starts (teacher-forcing fixtures)
            trace.append(dict(Xg=Xg.copy(), Ug=Ug.copy(), x=np.asarray(plant.lift(xs[step]), dtype=complex).copy()))
        while not done and n_iter < max_iter:
            A_ls, B_ls, D_ls = model.along(Xg, Ug, H)
            u_prev = us[step - 1] if step > 1 else U_ref[:, 0]
            Xo, Uo, obj, _ = qp(plant.lift(xs[step]), X_ref, U_ref, Q_ls, R_ls, A_ls, B_ls, D_ls,
                                np.real(u_prev), sat, du)
            if np.isinf(obj):
                exit_code = 3
                break
            if step > (1 if warm_start else np.inf):
                alpha, done = 1.0, True
            else:
                alpha, new_step = line_search(Q_ls, R_ls, X_ref, U_ref, Xg, Ug, Xo, Uo)
                done = new_step < 1e-4
            Xg = Xg + alpha * (Xo - Xg)
            Ug = Ug + alpha * (Uo - Ug)
            n_iter += 1
        if exit_code:
            break
        qp_counts.append(n_iter)
        us[step] = Uo[:, 0].copy()
        if (step + 1) % mf == 0:
            window = [us[step - j] for j in range(mf)]
            x = xs[step + 1 - mf]
            for useg in window:
                x = plant.segment(x, useg, dt) if hasattr(plant, 'segment') else \
                    expm_plant_segment(x, plant.H0, plant.H1_list, useg, dt)
            xs[step + 1] = x
        else:
            xs[step + 1] = np.asarray(plant.proj(model.step(plant.lift(xs[step]), us[step])), dtype=complex)
        Xg = np.hstack([Xg[:, 1:], Xg[:, -1:]])
        Ug = np.hstack([Ug[:, 1:], Ug[:, -1:]])
        X_ref = np.atleast_2d(X_targ[:, step:step + H + 1])
        U_ref = np.atleast_2d(U_targ[:, step:step + H])
        if exit_condition is not None and exit_condition(xs[step + 1], xs[step], us[step]):   # mpc.py:289-292
            exit_code = 1
            break
    if stats is not None:
        stats['qp_per_step'] = qp_counts
    if exit_code == 1:      # the reference returns xs[:step + 1], us[:step] with step the index it broke at (mpc.py:298-304)
        return np.array(xs[:step + 1]).T, (np.array(us[:step]).T if step > 0 else None), exit_code
    if exit_code:
        return None, None, exit_code
    return np.array(xs).T, np.array(us).T, exit_code
