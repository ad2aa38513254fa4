"""Generate tests/golden/loop_transmon_models.npz -- TEST INFRASTRUCTURE ONLY; build container only.

Ensembles of perturbed MODELS: member k controls plant k (systems.ensemble_transmon) with its own model k
(systems.transmon_model_liouvillians).  Each member is one run of the reference's unmodified mpc() through the shim,
with that member's DMDc and plant, cross-checked against the restatement before it is written.

    python -m oracle.make_golden_models
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import refshim, restate as rs                      # noqa: E402
from oracle.make_golden import reference_loop, restated_loop, OUT   # noqa: E402
from mpc4quantum_b200 import systems                           # noqa: E402
from mpc4quantum_b200.model import DMDc                        # noqa: E402


def main(n_members=4):
    if not refshim.available():
        raise SystemExit('the reference tree is not present: fixtures can only be generated in the build container')
    out = {}
    for order in (1, 2):
        cfg = systems.config_transmon(order, discretize=rs.taylor_discretize)
        plants, _ = systems.ensemble_transmon(65536)
        L, params = systems.transmon_model_liouvillians(65536)
        A, xs_all, us_all, fid, cnt, sens, fsens = [], [], [], [], [], [], []
        for k in range(n_members):
            A_k = rs.taylor_discretize(list(L[k]), cfg['clock'].dt, order)
            cfg_k = dict(cfg)
            cfg_k['model'] = DMDc(9, 9, A_k.shape[1] - 9, A_k)
            member = plants.member(k)
            xs, us, ec, counts = reference_loop(cfg_k, plant=member)
            xs2, us2, ec2, counts2 = restated_loop(cfg_k, plant=member)
            assert ec == ec2 == 0 and np.array_equal(counts, counts2)
            # two exact CPU runs of the same member (reference loop vs restatement) differ by round-off amplified through
            # the closed loop of a mismatched model; recorded per member as the conditioning of that trajectory
            gap = max(np.abs(xs - xs2).max(), np.abs(us - us2).max())
            assert gap < 1e-4, gap
            assert np.abs(us[:, :3] - us2[:, :3]).max() < 1e-9
            sens.append(np.abs(us - us2).max())
            fsens.append(abs(float(np.real(np.vdot(cfg['target'], xs[:, -1] - xs2[:, -1])))))
            A.append(A_k)
            xs_all.append(xs)
            us_all.append(us)
            cnt.append(counts)
            fid.append(float(np.real(np.vdot(cfg['target'], xs[:, -1]))))
        # the perturbed models really differ from the nominal one
        assert np.abs(np.array(A) - cfg['model'].A[None]).max() > 1e-3
        print('== order %d: fidelities %s, QP solves %s, conditioning |du| %s |dfid| %s'
              % (order, np.round(fid, 6), [int(c.sum()) for c in cnt], np.array2string(np.array(sens), precision=1),
                 np.array2string(np.array(fsens), precision=1)))
        out.update({'o%d_A' % order: np.array(A), 'o%d_xs' % order: np.array(xs_all), 'o%d_us' % order: np.array(us_all),
                    'o%d_fidelity' % order: np.array(fid), 'o%d_qp_per_step' % order: np.array(cnt),
                    'o%d_us_sensitivity' % order: np.array(sens), 'o%d_fid_sensitivity' % order: np.array(fsens)})
    out['L'] = L[:n_members]
    np.savez_compressed(os.path.join(OUT, 'loop_transmon_models.npz'), **out)


if __name__ == '__main__':
    main()
