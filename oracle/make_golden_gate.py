"""Generate tests/golden/gate.npz and loop_not_gate_*.npz -- TEST INFRASTRUCTURE ONLY; build container only.

Gate synthesis: the reference's ``QSynthesis`` (experiment.py:336-417) and the closed
loop of ``test_NOT_gate`` (tests/test_mpc4quantum.py:48-97).

* ``QSynthesis.lift`` / ``QSynthesis.proj`` are static numpy functions: the vectors come from the reference itself.
* ``QSynthesis.simulate`` needs qutip.propagator (absent): restated as expm per constant segment; the plant handed to
  the reference's ``mpc()`` composes the reference's own ``proj`` and ``lift`` around that restated factor.
* The reference test cannot run as written (``RWA_Qubit(**{'w0', 'w1', 'wR'})`` is not its signature; the target has
  H + 1 columns only; ``QSynthesis.lift`` applied to a 16-vector at mpc.py:135 yields 256 entries).  The loop fixture
  uses the wiring the test evidently intends -- process vectors with identity observable maps -- through the
  reference's unmodified ``mpc()``, and is cross-checked against the independent restatement before it is written.

    python -m oracle.make_golden_gate
"""
import os
import sys
import warnings

import numpy as np
from scipy.linalg import expm
from scipy.stats import unitary_group

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import refshim, restate as rs          # noqa: E402
from mpc4quantum_b200 import systems               # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def unit_vectors():
    exp_mod = refshim.module('experiment')
    QS = exp_mod.QSynthesis
    out = {}
    for n in (2, 3):
        Us = [unitary_group.rvs(n, random_state=100 * n + i) for i in range(4)]
        if n == 2:
            Us += [systems.SX.copy(), 1j * systems.SY, systems.rx(1e-3)]      # zero leading blocks, near identity
        else:
            Us.append(np.roll(np.eye(3), 1, axis=1).astype(complex))          # permutation: first block is zero
        lifted = np.array([QS.lift(U.flatten()) for U in Us])
        back = np.array([QS.proj(p) for p in lifted])
        for U, p, b in zip(Us, lifted, back):
            assert np.abs(rs.lift_process(U.flatten()) - p).max() < 1e-15
            assert np.abs(rs.proj_process(p) - b).max() < 1e-14
            assert np.abs(QS.lift(b) - p).max() < 1e-13          # proj returns U up to a phase
        out['U%d' % n], out['lift%d' % n], out['proj%d' % n] = np.array(Us), lifted, back
    # process propagation: lift(V proj(P)) over a few segments
    rng = np.random.default_rng(5)
    H0, H1 = 0.3 * systems.SZ, [0.5 * systems.SX, 0.5 * systems.SY]
    u = rng.uniform(-1, 1, size=(2, 6))
    ts = 0.1 * np.arange(7)
    P = [QS.lift(systems.rx(0.2).flatten())]
    for i in range(6):
        V = expm(-1j * (H0 + u[0, i] * H1[0] + u[1, i] * H1[1]) * 0.1)
        P.append(QS.lift((V @ QS.proj(P[-1]).reshape(2, 2)).flatten()))
    pl = rs.ProcessPlant(H0, H1)
    assert np.abs(pl.simulate(P[0], ts, u) - np.array(P).T).max() < 1e-13
    out.update(sim_H0=H0, sim_H1=np.array(H1), sim_u=u, sim_ts=ts, sim_P=np.array(P).T)
    np.savez_compressed(os.path.join(OUT, 'gate.npz'), **out)
    print('== gate unit vectors: %d arrays' % len(out))


def reference_loop(cfg, H0, H1_list, exit_condition=None):
    m4q = refshim.load()
    mpc_mod = refshim.module('mpc')
    exp_mod = refshim.module('experiment')
    QS = exp_mod.QSynthesis
    counts = [0]

    def qp(*a, **k):
        counts[-1] += 1
        return rs.qp_exact(*a, **k)
    refshim.inject_qp(qp)

    class Plant(exp_mod.Experiment):           # identity lift / proj (experiment.py:29-37)
        def __init__(self):
            exp_mod.Experiment.__init__(self)

        def f(self, t, x, u):
            raise NotImplementedError

        def simulate(self, x0, ts, us):        # experiment.py:390-417 with propagator() restated as expm
            n = H0.shape[0]
            U = QS.proj(np.asarray(x0, dtype=complex).reshape(-1)).reshape(n, n)
            cols = [QS.lift(U.flatten())]
            for i in range(len(ts) - 1):
                Ham = H0 + sum(uk * Hk for uk, Hk in zip(np.atleast_1d(us(ts[i])), H1_list))
                U = expm(-1j * Ham * (ts[i + 1] - ts[i])) @ U
                cols.append(QS.lift(U.flatten()))
            counts.append(0)
            return np.array(cols).T

    clock = mpc_mod.StepClock(cfg['clock'].dt, cfg['clock'].horizon, cfg['clock'].n_steps)
    c = cfg['model'].A.shape[0]
    model = m4q.DMDc(c, c, cfg['model'].A.shape[1] - c, cfg['model'].A)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        data, _, exit_code = m4q.mpc(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], clock, Plant(),
                                     model, cfg['Q'], cfg['R'], cfg['Qf'], sat=cfg['sat'], du=cfg['du'],
                                     exit_condition=exit_condition, warm_start=cfg['warm_start'], progress_bar=False)
    n_done = data[0].shape[1] - 1 + (1 if exit_code == 1 else 0)
    return data[0], data[1], exit_code, np.array(counts[:n_done])


def restated_loop(cfg, H0, H1_list, exit_condition=None):
    stats = {}
    xs, us, ec = rs.mpc_loop(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'].dt,
                             cfg['clock'].horizon, cfg['clock'].n_steps, rs.ProcessPlant(H0, H1_list), cfg['model'].A,
                             cfg['Q'], cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'], warm_start=cfg['warm_start'],
                             stats=stats, exit_condition=exit_condition)
    return xs, us, ec, np.array(stats['qp_per_step'])


def gate_fidelity(cfg, p):
    return float(np.real(np.vdot(cfg['target'], p)))


def loop_fixture(name, cfg, use_exit, n_members=0):
    print('== %s' % name)
    nominal = cfg['experiment']
    ex = cfg['exit_condition'] if use_exit else None
    xs, us, ec, counts = reference_loop(cfg, nominal.H0, nominal.H1_list, ex)
    xs2, us2, ec2, counts2 = restated_loop(cfg, nominal.H0, nominal.H1_list, ex)
    assert ec == ec2, (ec, ec2)
    assert xs.shape == xs2.shape and us.shape == us2.shape, (xs.shape, xs2.shape, us.shape, us2.shape)
    assert np.array_equal(counts, counts2), (counts, counts2)
    gap = max(np.abs(xs - xs2).max(), np.abs(us - us2).max())
    assert gap < 1e-6, gap
    fid = gate_fidelity(cfg, xs[:, -1])
    print('   reference == restatement (gap %.1e); exit code %d after %d steps, %d QP solves, gate fidelity %.9f'
          % (gap, ec, us.shape[1], counts.sum(), fid))
    out = dict(xs=xs, us=us, exit_code=ec, qp_per_step=counts, fidelity=fid, A_full=cfg['model'].A, x0=cfg['x0'],
               restatement_gap=gap)
    if n_members:
        ens, params = systems.ensemble_not_gate(4096)
        e_xs, e_us, e_fid, e_cnt = [], [], [], []
        for k in range(n_members):
            H0, H1 = ens.H0[k], list(ens.H1[k])
            x, u, e, cnt = restated_loop(cfg, H0, H1)
            assert e == 0
            if k == 0:
                xr, ur, er, cr = reference_loop(cfg, H0, H1)
                assert np.abs(xr - x).max() < 1e-6 and np.abs(ur - u).max() < 1e-6 and np.array_equal(cr, cnt)
            e_xs.append(x)
            e_us.append(u)
            e_cnt.append(cnt)
            e_fid.append(gate_fidelity(cfg, x[:, -1]))
        print('   ensemble members 0..%d: gate fidelity %s' % (n_members - 1, np.round(e_fid, 6)))
        out.update(ens_xs=np.array(e_xs), ens_us=np.array(e_us), ens_fidelity=np.array(e_fid),
                   ens_qp_per_step=np.array(e_cnt))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)


def main():
    if not refshim.available():
        raise SystemExit('the reference tree is not present: fixtures can only be generated in the build container')
    os.makedirs(OUT, exist_ok=True)
    disc = rs.taylor_discretize
    unit_vectors()
    loop_fixture('loop_not_gate_o1', systems.config_not_gate(1, discretize=disc), use_exit=False, n_members=4)
    loop_fixture('loop_not_gate_o2', systems.config_not_gate(2, discretize=disc), use_exit=False)
    loop_fixture('loop_not_gate_o1_exit', systems.config_not_gate(1, n_steps=90, discretize=disc), use_exit=True)


if __name__ == '__main__':
    main()
