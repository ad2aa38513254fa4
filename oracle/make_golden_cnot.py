"""CNOT state preparation (tests/test_mpc4quantum.py:399-466; c = 16, m = 3, H = 50, ramped target) through the
reference's own mpc() -- TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).

    python -m oracle.make_golden_cnot [n_steps]

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


if __name__ == '__main__':
    main()
