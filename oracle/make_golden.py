"""Generate tests/golden/*.npz -- TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).

Every vector here is an output of the REFERENCE's own code imported through oracle/refshim.py, with the two leaves
that live in absent third-party packages (cvxpy/OSQP, qutip.mesolve) replaced by the exact CPU restatements of
oracle/restate.py.  The restatement (oracle/restate.py) is checked against the same runs before anything is
written, so a fixture that exists has passed "reference == restatement".

    python -m oracle.make_golden            # from the repository root
"""
import os
import sys

import numpy as np
from scipy.linalg import expm

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import refshim, restate as rs, admm_model as am   # noqa: E402
from mpc4quantum_b200 import systems               # noqa: E402  (pure-numpy system definitions only)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def reference_loop(cfg, plant=None):
    """Run the reference's mpc() (mpc.py:128-304) verbatim with the restated QP / plant leaves."""
    m4q = refshim.load()
    mpc_mod = refshim.module('mpc')
    exp_mod = refshim.module('experiment')
    counts = []

    def qp(*a, **k):
        counts[-1] += 1
        return rs.qp_exact(*a, **k)
    refshim.inject_qp(qp)

    base = {'coupled': exp_mod.QCoupledExperiment, 'trunc32': exp_mod.QExperiment32}.get(cfg.get('kind'), exp_mod.Experiment)
    src = plant if plant is not None else cfg['experiment']
    H0, H1_list = src.H0, src.H1_list

    class Plant(base):
        def __init__(self):
            exp_mod.Experiment.__init__(self)

        def f(self, t, x, u):
            raise NotImplementedError

        def simulate(self, x0, ts, us):
            out = [np.asarray(x0, dtype=complex).reshape(-1)]
            for i in range(len(ts) - 1):
                out.append(rs.expm_plant_segment(out[-1], H0, H1_list, us(ts[i]), ts[i + 1] - ts[i]))
            return np.array(out).T

    clock = mpc_mod.StepClock(cfg['clock'].dt, cfg['clock'].horizon, cfg['clock'].n_steps)
    clock.measure_freq = cfg['clock'].measure_freq
    c = cfg['model'].A.shape[0]
    model = m4q.DMDc(c, c, cfg['model'].A.shape[1] - c, cfg['model'].A)

    # QP solves per MPC step: every step ends with exactly one plant simulate() or one model.predict() (mpc.py:252-267)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        counts.append(0)
        plant_obj = Plant()
        sim = plant_obj.simulate

        def simulate(x0, ts, us):
            r = sim(x0, ts, us)
            counts.append(0)
            return r
        plant_obj.simulate = simulate
        pred = model.predict

        def predict(x, u):
            r = pred(x, u)
            counts.append(0)
            return r
        model.predict = predict
        data, _, exit_code = m4q.mpc(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], clock,
                                     plant_obj, model, cfg['Q'], cfg['R'], cfg['Qf'], sat=cfg['sat'], du=cfg['du'],
                                     warm_start=cfg['warm_start'], progress_bar=False)
    return data[0], data[1], exit_code, np.array(counts[:cfg['clock'].n_steps])


def restated_loop(cfg, plant=None, qp=rs.qp_exact):
    src = plant if plant is not None else cfg['experiment']
    lift, proj = {'coupled': (rs.lift_coupled, rs.proj_coupled), 'trunc32': (rs.lift_32, None)}.get(
        cfg.get('kind'), (rs.lift_identity, rs.lift_identity))
    pl = rs.ExpmPlant(src.H0, src.H1_list, lift, proj)
    stats = {}
    xs, us, ec = rs.mpc_loop(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'].dt,
                             cfg['clock'].horizon, cfg['clock'].n_steps, pl, cfg['model'].A, cfg['Q'], cfg['R'],
                             cfg['Qf'], cfg['sat'], cfg['du'], warm_start=cfg['warm_start'],
                             measure_freq=cfg['clock'].measure_freq, stats=stats, qp=qp)
    return xs, us, ec, np.array(stats['qp_per_step'])


def closed_loop_fixture(name, cfg, ensemble=None, n_members=0):
    print('== %s' % name)
    xs, us, ec, counts = reference_loop(cfg)
    xs2, us2, ec2, counts2 = restated_loop(cfg)
    assert ec == ec2 == 0
    assert np.array_equal(counts, counts2), (counts, counts2)
    gap = max(np.abs(xs - xs2).max(), np.abs(us - us2).max())
    # the slowly converging SQP of step 0 (39 iterations, stop at step < 1e-4) amplifies round-off differences in the
    # QP data: reference and restatement agree to 1e-9 .. 2e-7 on the qubit, 1e-12 elsewhere
    assert gap < 1e-6, gap
    fid = float(np.real(np.vdot(cfg['target'], xs[:, -1])))
    print('   reference == restatement (gap %.1e); %d QP solves, per step %s..., final <target|rho|target> = %.9f'
          % (gap, counts.sum(), counts[:4], fid))
    out = dict(xs=xs, us=us, exit_code=ec, qp_per_step=counts, fidelity=fid, A_full=cfg['model'].A, x0=cfg['x0'],
               restatement_gap=gap)
    if ensemble is not None and n_members:
        exps, params = ensemble
        e_xs, e_us, e_fid, e_cnt, s_us, s_fid = [], [], [], [], [], []
        for k in range(n_members):
            member = exps.member(k)
            x, u, e, cnt = restated_loop(cfg, plant=member)
            assert e == 0
            # conditioning of this member's closed loop: the same loop with a second exact QP solver (the numpy model
            # of the device algorithm; single QPs agree with qp_exact to ~1e-14).  Badly mismatched plants amplify that
            # round-off by up to 1e9 over the trajectory, which bounds how well ANY implementation can match.
            warm = {}

            def qp2(*a, **kw):
                return am.qp_admm(*a, rho=0.1, eps=1e-2, warm=warm.get('warm'), stats=warm, **kw)
            x2, u2, e2, cnt2 = restated_loop(cfg, plant=member, qp=qp2)
            s_us.append(np.abs(u2 - u).max())
            s_fid.append(abs(float(np.real(np.vdot(cfg['target'], x2[:, -1] - x[:, -1])))))
            if k == 0:      # anchor the ensemble path on the reference loop as well
                xr, ur, er, cr = reference_loop(cfg, plant=member)
                assert np.abs(xr - x).max() < 1e-6 and np.abs(ur - u).max() < 1e-6 and np.array_equal(cr, cnt)
            e_xs.append(x)
            e_us.append(u)
            e_cnt.append(cnt)
            e_fid.append(float(np.real(np.vdot(cfg['target'], x[:, -1]))))
        print('   ensemble members 0..%d: fidelity %s' % (n_members - 1, np.round(e_fid, 6)))
        print('   closed-loop conditioning (second exact solver): |du| %s |dfid| %s'
              % (np.array2string(np.array(s_us), precision=1), np.array2string(np.array(s_fid), precision=1)))
        out.update(ens_xs=np.array(e_xs), ens_us=np.array(e_us), ens_fidelity=np.array(e_fid),
                   ens_qp_per_step=np.array(e_cnt), ens_us_sensitivity=np.array(s_us),
                   ens_fid_sensitivity=np.array(s_fid))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)


def unit_fixtures():
    """Function-level vectors straight from the reference's functions."""
    m4q = refshim.load()
    lin = refshim.module('linearize')
    vec = refshim.module('vectorize')
    mpc_mod = refshim.module('mpc')
    rng = np.random.default_rng(7)
    out = {}
    # --- discretize_homogeneous (vectorize.py:8-49) on the three systems, orders 1..3; vectorize_me (52-75)
    tr = systems.RWA_Transmon(alpha=-2 * np.pi * 0.4)
    qb = systems.RWA_Qubit(2 * np.pi * 3.9, 2 * np.pi * 4, 2 * np.pi * 4)
    cp = systems.RWA_Coupled()
    for tag, Hs, dt, orders in (('qubit', qb.H_list, 1.0, (1, 2, 3)), ('transmon', tr.H_list, 0.25, (1, 2, 3)),
                                ('coupled', cp.H_list, 0.25, (1, 2))):
        Ls = [rs.liouvillian(h) for h in Hs]
        d = Hs[0].shape[0]
        basis = [refshim.Qobj(np.outer(np.eye(d)[a], np.eye(d)[b])) for a in range(d) for b in range(d)]
        for h, L in zip(Hs, Ls):
            assert np.abs(vec.vectorize_me(refshim.Qobj(h), basis) - L).max() < 1e-13
        out['disc_%s_L' % tag] = np.array(Ls)
        out['disc_%s_dt' % tag] = dt
        for o in orders:
            ref = vec.discretize_homogeneous(Ls, dt, o)
            assert np.abs(ref - rs.taylor_discretize(Ls, dt, o)).max() < 1e-13
            out['disc_%s_o%d' % (tag, o)] = ref
    # vectorize_me in a non-trivial (Pauli) basis
    paulis = [np.eye(2), systems.SX, systems.SY, systems.SZ]
    Hq = 0.3 * systems.SX - 0.7 * systems.SY + 1.1 * systems.SZ
    out['vecme_H'] = Hq
    out['vecme_basis'] = np.array(paulis)
    out['vecme_out'] = vec.vectorize_me(refshim.Qobj(Hq), [refshim.Qobj(p) for p in paulis])
    assert np.abs(out['vecme_out'] - rs.liouvillian_in_basis(Hq, paulis)).max() < 1e-13
    # --- WrapModel.get_model_along_traj (linearize.py:61-70)
    for tag, Hs, dt, o, m, H in (('qubit', qb.H_list, 1.0, 2, 1, 10), ('transmon', tr.H_list, 0.25, 2, 2, 16),
                                 ('transmon1', tr.H_list, 0.25, 1, 2, 16), ('coupled', cp.H_list, 0.25, 1, 3, 12)):
        A_full = vec.discretize_homogeneous([rs.liouvillian(h) for h in Hs], dt, o)
        c = A_full.shape[0]
        wm = lin.WrapModel(A_full[:, :c], A_full[:, c:], m, o)
        X = rng.normal(size=(c, H + 1)) + 1j * rng.normal(size=(c, H + 1))
        U = rng.normal(size=(m, H))
        A_ls, B_ls, D_ls = wm.get_model_along_traj(X, U, np.arange(H))
        bm = rs.BilinearModel(A_full, m, o)
        A2, B2, D2 = bm.along(X, U, H)
        assert max(np.abs(np.array(A_ls) - np.array(A2)).max(), np.abs(np.array(B_ls) - np.array(B2)).max(),
                   np.abs(np.array(D_ls)[:, :, 0] - np.array(D2)).max()) < 1e-12
        out.update({'lin_%s_A_full' % tag: A_full, 'lin_%s_X' % tag: X, 'lin_%s_U' % tag: U,
                    'lin_%s_order' % tag: o, 'lin_%s_A' % tag: np.array(A_ls), 'lin_%s_B' % tag: np.array(B_ls),
                    'lin_%s_D' % tag: np.array(D_ls)[:, :, 0]})
    # --- iqp_line_search (mpc.py:101-125), diagonal and full Hermitian costs
    for tag, c, m, H, full in (('qubit', 4, 1, 10, False), ('transmon', 9, 2, 16, False), ('transmon_full', 9, 2, 16, True),
                               ('cross', 8, 2, 20, True)):
        def herm(n):
            a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
            return a @ a.conj().T / n
        Q = herm(c) if full else np.diag(rng.uniform(0, 1, c)).astype(complex)
        Qf = herm(c) if full else np.diag(rng.uniform(0, 1, c)).astype(complex)
        R = np.real(herm(m)) if full else np.diag(rng.uniform(0.1, 1, m))
        Q_ls, R_ls = [Q] * H + [Qf], [R] * H
        Xs = [rng.normal(size=(c, H + 1)) + 1j * rng.normal(size=(c, H + 1)) for _ in range(3)]
        Us = [rng.normal(size=(m, H)) for _ in range(3)]
        alpha, step, _, _ = mpc_mod.iqp_line_search(Q_ls, R_ls, Xs[0], Us[0], Xs[1], Us[1], Xs[2], Us[2])
        a2, s2 = rs.line_search(Q_ls, R_ls, Xs[0], Us[0], Xs[1], Us[1], Xs[2], Us[2])
        assert abs(alpha - a2) < 1e-12 and abs(step - s2) < 1e-10
        out.update({'ls_%s_Q' % tag: np.array(Q_ls), 'ls_%s_R' % tag: np.array(R_ls), 'ls_%s_X' % tag: np.array(Xs),
                    'ls_%s_U' % tag: np.array(Us), 'ls_%s_alpha' % tag: alpha, 'ls_%s_step' % tag: step})
    # --- QCoupledExperiment.lift / proj (experiment.py:248-306) -- the reference test_partialTrace cases
    exp_mod = refshim.module('experiment')
    rho = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    rho = rho @ rho.conj().T
    rho /= np.trace(rho)
    out['lift_rho'] = rho.reshape(-1)
    out['lift_out'] = exp_mod.QCoupledExperiment.lift(rho.reshape(-1))
    out['proj_out'] = exp_mod.QCoupledExperiment.proj(out['lift_out'])
    assert np.abs(rs.lift_coupled(rho.reshape(-1)) - out['lift_out']).max() < 1e-14
    assert np.abs(rs.proj_coupled(out['lift_out']) - out['proj_out']).max() < 1e-14
    # --- plant propagators: scipy.linalg.expm on perturbed Hamiltonians (restates qutip mesolve, experiment.py:202-212)
    for tag, ens, dt in (('qubit', systems.ensemble_qubit(16)[0], 1.0), ('transmon', systems.ensemble_transmon(16)[0], 0.25),
                         ('cross', systems.ensemble_crosstalk(16)[0], 0.5)):
        n, m = len(ens), ens.dim_u
        u = rng.uniform(-1.5, 1.5, size=(n, 3, m))
        d = ens.d
        props = np.zeros((n, 3, d, d), dtype=complex)
        for k in range(n):
            for sgm in range(3):
                Ham = ens.H0[k] + sum(u[k, sgm, i] * ens.H1[k, i] for i in range(m))
                props[k, sgm] = expm(-1j * Ham * dt)
        out.update({'expm_%s_u' % tag: u, 'expm_%s_props' % tag: props, 'expm_%s_dt' % tag: dt})
    np.savez_compressed(os.path.join(OUT, 'unit.npz'), **out)
    print('== unit fixtures: %d arrays' % len(out))
    del m4q


def qp_fixtures():
    """QP instances harvested from closed loops (so they look like what the loop solves) + exact solutions."""
    rng = np.random.default_rng(11)
    out = {}
    for tag, cfg in (('qubit', systems.config_qubit(1, discretize=rs.taylor_discretize)),
                     ('transmon', systems.config_transmon(2, discretize=rs.taylor_discretize)),
                     ('cross', systems.config_crosstalk(0.05, n_steps=4, discretize=rs.taylor_discretize))):
        c, m, H = cfg['model'].A.shape[0], cfg['dim_u'], cfg['clock'].horizon
        bm = rs.BilinearModel(cfg['model'].A, m, cfg['order'])
        n_inst = 6
        keys = ('x_init', 'X_bm', 'U_bm', 'A', 'B', 'D', 'u_prev', 'X', 'U', 'obj')
        acc = {k: [] for k in keys}
        lift = rs.lift_coupled if cfg.get('kind') == 'coupled' else rs.lift_identity
        for i in range(n_inst):
            scale = [0.0, 0.2, 0.5, 1.0, 1.5, 2.5][i]
            Ug = np.clip(rng.normal(size=(m, H)) * scale * cfg['sat'], -cfg['sat'], cfg['sat'])
            x0 = lift(cfg['x0'])
            Xg = np.zeros((c, H + 1), dtype=complex)
            Xg[:, 0] = x0
            for t in range(H):
                Xg[:, t + 1] = bm.step(Xg[:, t], Ug[:, t])
            A_ls, B_ls, D_ls = bm.along(Xg, Ug, H)
            u_prev = rng.uniform(-0.5, 0.5, m) * cfg['sat']
            X_bm, U_bm = cfg['X_targ'][:, :H + 1], cfg['U_targ'][:, :H]
            X, U, obj, info = rs.qp_exact(x0, X_bm, U_bm, [cfg['Q']] * H + [cfg['Qf']], [cfg['R']] * H, A_ls, B_ls,
                                          D_ls, u_prev, cfg['sat'], cfg['du'])
            assert info['kkt'][0] < 1e-7 and info['kkt'][1] < 1e-12, info['kkt']
            for k, v in zip(keys, (x0, X_bm, U_bm, np.array(A_ls), np.array(B_ls), np.array(D_ls), u_prev, X, U, obj)):
                acc[k].append(v)
        for k in keys:
            out['%s_%s' % (tag, k)] = np.array(acc[k])
        out['%s_Q' % tag], out['%s_Qf' % tag], out['%s_R' % tag] = cfg['Q'], cfg['Qf'], cfg['R']
        out['%s_sat' % tag], out['%s_du' % tag] = cfg['sat'], cfg['du']
        nb = [(np.abs(np.abs(u) - cfg['sat']) < 1e-9).sum() for u in acc['U']]
        print('== qp %s: controls on the saturation bound per instance %s' % (tag, nb))
    np.savez_compressed(os.path.join(OUT, 'qp.npz'), **out)


def main():
    if not refshim.available():
        raise SystemExit('the reference tree is not present: fixtures can only be generated in the build container')
    os.makedirs(OUT, exist_ok=True)
    disc = rs.taylor_discretize
    unit_fixtures()
    qp_fixtures()
    closed_loop_fixture('loop_qubit_o1', systems.config_qubit(1, discretize=disc), systems.ensemble_qubit(4096), 6)
    closed_loop_fixture('loop_qubit_o2', systems.config_qubit(2, discretize=disc))
    closed_loop_fixture('loop_transmon_o1', systems.config_transmon(1, discretize=disc),
                        systems.ensemble_transmon(65536), 4)
    closed_loop_fixture('loop_transmon_o2', systems.config_transmon(2, discretize=disc))
    closed_loop_fixture('loop_transmon_o1_h50', systems.config_transmon(1, horizon=50, n_steps=6, discretize=disc))
    closed_loop_fixture('loop_crosstalk', systems.config_crosstalk(0.05, n_steps=12, discretize=disc),
                        systems.ensemble_crosstalk(65536), 3)


if __name__ == '__main__':
    main()
