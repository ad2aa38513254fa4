"""Generate tests/golden/streaming.npz -- TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).

Vectors for the data-driven models (model.py:109-313) and for mpc(streaming=True) (mpc.py:281-285), all produced by the
REFERENCE's own classes and loop imported through oracle/refshim.py (QP and plant leaves as in oracle/make_golden.py).

    python -m oracle.make_golden_streaming            # from the repository root
"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import refshim, restate as rs          # noqa: E402
from mpc4quantum_b200 import systems               # noqa: E402  (pure-numpy system definitions only)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def crandn(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def main():
    m4q = refshim.load()
    mpc_mod = refshim.module('mpc')
    exp_mod = refshim.module('experiment')
    rng = np.random.default_rng(7)
    out = {}

    # ---- OnlineDMDc: bootstrap + rank-1 updates with a discount
    dy, dx, du_ = 4, 4, 8
    A0 = crandn(rng, dy, dx + du_)
    ys, xs_, us_ = crandn(rng, 6, dy), crandn(rng, 6, dx), crandn(rng, 6, du_)
    mdl = m4q.OnlineDMDc.from_bootstrap(dy, dx, du_, A0.copy(), alpha=1e2)
    mdl.discount = 0.95
    for k in range(6):
        mdl.fit_iteration(ys[k], xs_[k], us_[k])
    out.update(on_A0=A0, on_y=ys, on_x=xs_, on_u=us_, on_A=mdl.A, on_P=mdl.P, on_pred=mdl.predict(xs_[0], us_[0]))
    Y, X, U = crandn(rng, dy, 20), crandn(rng, dx, 20), crandn(rng, du_, 20)
    mdl = m4q.OnlineDMDc.from_data(Y, X, U)
    mdl.fit_iteration(ys[0], xs_[0], us_[0])
    out.update(on_Y=Y, on_X=X, on_U=U, on_data_A=mdl.A, on_data_P=mdl.P)

    # ---- DiscrepDMDc: offline fit, then discrepancy updates; bootstrap below the rank threshold does nothing
    mdl = m4q.DiscrepDMDc.from_data(Y, X, U, rcond=1e-8)
    out['di_A0'] = mdl.A.copy()
    mdl.discount = 0.9
    for k in range(3):
        mdl.fit_iteration(ys[k], xs_[k], us_[k])
    out.update(di_A=mdl.A, di_Ystack=mdl.Y)
    mdl = m4q.DiscrepDMDc.from_bootstrap(dy, dx, du_, A0.copy())
    mdl.fit_iteration(ys[0], xs_[0], us_[0])
    mdl.fit_iteration(ys[1], xs_[1], us_[1])
    out['di_boot_A_rank_deficient'] = mdl.A.copy()
    for k in range(2, 6):
        mdl.fit_iteration(ys[k], xs_[k], us_[k])
    out['di_boot_A'] = mdl.A

    # ---- mpc(streaming=True): qubit, one plant measurement every 5 steps, OnlineDMDc model
    cfg = systems.config_qubit_freq(1, n_steps=15, discretize=rs.taylor_discretize)
    refshim.inject_qp(rs.qp_exact)
    H0, H1_list = cfg['experiment'].H0, cfg['experiment'].H1_list

    class Plant(exp_mod.Experiment):
        def __init__(self):
            exp_mod.Experiment.__init__(self)

        def f(self, t, x, u):
            raise NotImplementedError

        def simulate(self, x0, ts, us):
            o = [np.asarray(x0, dtype=complex).reshape(-1)]
            for i in range(len(ts) - 1):
                o.append(rs.expm_plant_segment(o[-1], H0, H1_list, us(ts[i]), ts[i + 1] - ts[i]))
            return np.array(o).T

    clock = mpc_mod.StepClock(cfg['clock'].dt, cfg['clock'].horizon, cfg['clock'].n_steps)
    clock.measure_freq = cfg['clock'].measure_freq
    c = cfg['model'].A.shape[0]
    A_init = cfg['model'].A.copy()
    model = m4q.OnlineDMDc.from_bootstrap(c, c, A_init.shape[1] - c, A_init.copy(), alpha=1e2)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        data, model2, ec = m4q.mpc(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], clock, Plant(),
                                   model, cfg['Q'], cfg['R'], cfg['Qf'], sat=cfg['sat'], du=cfg['du'], streaming=True,
                                   progress_bar=False)
    assert ec == 0 and model2 is model
    assert np.abs(model.A - A_init).max() > 1e-6          # the model did move ...
    # ... but the controller never saw it: same controls as the non-streaming run up to the first model step
    print('streaming loop: exit', ec, '|A - A0| max %.3e' % np.abs(model.A - A_init).max())
    out.update(loop_xs=data[0], loop_us=data[1], loop_A=model.A, loop_P=model.P, loop_A0=A_init)
    np.savez_compressed(os.path.join(OUT, 'streaming.npz'), **out)
    print('wrote streaming.npz', {k: np.shape(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
