"""CNOT state preparation (tests/test_mpc4quantum.py:399-466; c = 16, m = 3, H = 50, ramped target) through the
reference's own mpc() -- TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).

    python -m oracle.make_golden_cnot [n_steps]        # closed loop through the reference's mpc()
    python -m oracle.make_golden_cnot 40 tf            # teacher-forcing data of the same run (restated loop)

Writes tests/golden/loop_cnot.npz: xs, us, SQP counts per step of the first n_steps (default 40) of the 200-step ramp,
checked against the restated loop before it is written.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import refshim, restate as rs                    # noqa: E402
from oracle.make_golden import reference_loop, restated_loop, OUT   # noqa: E402
from mpc4quantum_b200 import systems                         # noqa: E402


def main():
    if not refshim.available():
        raise SystemExit('the reference tree is not present: fixtures can only be generated in the build container')
    n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    cfg = systems.config_cnot(n_steps=n_steps, horizon=50, ramp_steps=200, discretize=rs.taylor_discretize)
    xs, us, ec, counts = reference_loop(cfg)
    xs2, us2, ec2, counts2 = restated_loop(cfg)
    assert ec == ec2 == 0 and np.array_equal(counts, counts2), (ec, ec2, counts, counts2)
    gap = max(np.abs(xs - xs2).max(), np.abs(us - us2).max())
    assert gap < 2e-6, gap      # 40 closed-loop steps of a 39-iteration SQP start: two CPU runs part by 7e-7
    print('== loop_cnot: %d steps, reference == restatement (gap %.1e), QPs per step %s' % (n_steps, gap, counts))
    np.savez_compressed(os.path.join(OUT, 'loop_cnot.npz'), xs=xs, us=us, exit_code=ec, qp_per_step=counts,
                        A_full=cfg['model'].A, x0=cfg['x0'], restatement_gap=gap, n_steps=n_steps)


if __name__ == '__main__' and not (len(sys.argv) > 2 and sys.argv[2] == 'tf'):
    main()


def teacher_forcing(n_steps=40):
    """tests/golden/ens64_cnot.npz: what the controller knew at the start of every step of the run above and what it
    answered (the format of oracle/make_golden_ens64.py, one member: the nominal plant), for the per-step parity test."""
    cfg = systems.config_cnot(n_steps=n_steps, horizon=50, ramp_steps=200, discretize=rs.taylor_discretize)
    stats = {'want_trace': True}
    plant = rs.ExpmPlant(cfg['experiment'].H0, cfg['experiment'].H1_list)
    xs, us, ec = rs.mpc_loop(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'].dt,
                             cfg['clock'].horizon, cfg['clock'].n_steps, plant, cfg['model'].A, cfg['Q'], cfg['R'],
                             cfg['Qf'], cfg['sat'], cfg['du'], warm_start=cfg['warm_start'],
                             measure_freq=cfg['clock'].measure_freq, stats=stats)
    assert ec == 0
    tr = stats['trace']
    np.savez_compressed(os.path.join(OUT, 'ens64_cnot.npz'), us=us[None], xs=xs[None],
                        qp_per_step=np.array(stats['qp_per_step'])[None],
                        tf_x=np.array([t['x'] for t in tr])[None], tf_Xg=np.array([t['Xg'] for t in tr])[None],
                        tf_Ug=np.array([t['Ug'] for t in tr])[None], tf_us=us[None],
                        tf_qp_per_step=np.array(stats['qp_per_step'])[None])
    print('== ens64_cnot: teacher-forcing data of %d steps' % n_steps)


if __name__ == '__main__' and len(sys.argv) > 2 and sys.argv[2] == 'tf':
    teacher_forcing(int(sys.argv[1]))
