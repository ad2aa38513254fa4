"""64-member ensemble fixtures for BASELINE configs 2, 3 and 4 -- TEST INFRASTRUCTURE ONLY (build container, needs
/root/reference).

For each config the first 64 members of the seeded ensemble are run through the REFERENCE's own
``mpc()`` (mpc.py:128-304, imported by oracle/refshim.py, exact-QP and expm-plant leaves injected); the restated loop
(oracle/restate.py) is run beside it and must agree.  Two kinds of data are written to tests/golden/ens64_<config>.npz:

* closed loop:    us [64, m, S], fidelity [64], qp_per_step [64, S], xs_final [64, d*d]  (outputs of the reference loop)
* teacher forcing: for every member and every MPC step, what the controller knew when the step started --
  the lifted measured state x [64, S, c], the guesses Xg [64, S, c, H+1] / Ug [64, S, m, H] -- and what it answered,
  us[:, :, step] and the SQP count of the step.  A GPU test that feeds these inputs step by step compares every
  QP sequence of every member without closed-loop amplification (tests/test_gpu_parity64.py).

    python -m oracle.make_golden_ens64 [qubit transmon crosstalk]
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

N_MEMBERS = 64
MEMBERS = {'transmon_h50': 16, 'transmon_h100': 16, 'transmon_o2_h100': 8}      # the long-horizon oracle QPs cost ~0.4 s each
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def make_cfg(name):
    from oracle import restate as rs
    from mpc4quantum_b200 import systems
    disc = rs.taylor_discretize
    if name == 'qubit':
        return systems.config_qubit(1, discretize=disc), systems.ensemble_qubit(4096)[0]
    if name == 'transmon':
        return systems.config_transmon(1, discretize=disc), systems.ensemble_transmon(65536)[0]
    if name == 'transmon_h50':  # BASELINE config 3 horizon sweep, order-1 model at H = 50 (ill-conditioned cost-to-go)
        return systems.config_transmon(1, horizon=50, n_steps=20, discretize=disc), systems.ensemble_transmon(65536)[0]
    if name == 'transmon_h100':  # order-1 model at H = 100: cost-to-go beyond fp64, the device solves a pivoted KKT system
        return systems.config_transmon(1, horizon=100, n_steps=20, discretize=disc), systems.ensemble_transmon(65536)[0]
    if name == 'transmon_o2_h100':  # the same horizon with the order-2 model (||A_t|| ~ 1.02: the Riccati path certifies)
        return systems.config_transmon(2, horizon=100, n_steps=20, discretize=disc), systems.ensemble_transmon(65536)[0]
    if name == 'crosstalk':     # the config's own S = 50 (round 1 pinned S = 12 only)
        return systems.config_crosstalk(0.0, discretize=disc), systems.ensemble_crosstalk(65536)[0]
    raise SystemExit('unknown config %s' % name)


def _member(job):
    name, k = job
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle import make_golden as mg, restate as rs
    cfg, ens = make_cfg(name)
    member = ens.member(k)
    xs_r, us_r, ec_r, cnt_r = mg.reference_loop(cfg, plant=member)
    lift, proj = {'coupled': (rs.lift_coupled, rs.proj_coupled)}.get(cfg.get('kind'), (rs.lift_identity, rs.lift_identity))
    stats = {'want_trace': True}
    xs, us, ec = rs.mpc_loop(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'].dt,
                             cfg['clock'].horizon, cfg['clock'].n_steps, rs.ExpmPlant(member.H0, member.H1_list, lift, proj),
                             cfg['model'].A, cfg['Q'], cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'],
                             warm_start=cfg['warm_start'], measure_freq=cfg['clock'].measure_freq, stats=stats)
    assert ec == ec_r == 0
    cnt = np.array(stats['qp_per_step'])
    gap = max(np.abs(xs - xs_r).max(), np.abs(us - us_r).max())
    tr = stats['trace']
    return dict(k=k, us=us_r, xs=xs_r, cnt=cnt_r, gap=gap, cnt_same=bool(np.array_equal(cnt, cnt_r)),
                fid=float(np.real(np.vdot(cfg['target'], xs_r[:, -1]))),
                tf_x=np.array([t['x'] for t in tr]), tf_Xg=np.array([t['Xg'] for t in tr]),
                tf_Ug=np.array([t['Ug'] for t in tr]), tf_us=us, tf_cnt=cnt)


def build(name, pool):
    t0 = time.time()
    res = pool.map(_member, [(name, k) for k in range(MEMBERS.get(name, N_MEMBERS))])
    res.sort(key=lambda r: r['k'])
    gaps = np.array([r['gap'] for r in res])
    same = np.array([r['cnt_same'] for r in res])
    print('== ens64_%s: %d members in %.0f s; reference vs restatement: max gap %.2e, SQP counts equal for %d/%d'
          % (name, len(res), time.time() - t0, gaps.max(), same.sum(), len(res)))
    out = dict(us=np.array([r['us'] for r in res]), xs=np.array([r['xs'] for r in res]),
               fidelity=np.array([r['fid'] for r in res]), qp_per_step=np.array([r['cnt'] for r in res]),
               restatement_gap=gaps, restatement_counts_equal=same,
               tf_x=np.array([r['tf_x'] for r in res]), tf_Xg=np.array([r['tf_Xg'] for r in res]),
               tf_Ug=np.array([r['tf_Ug'] for r in res]), tf_us=np.array([r['tf_us'] for r in res]),
               tf_qp_per_step=np.array([r['tf_cnt'] for r in res]))
    np.savez_compressed(os.path.join(OUT, 'ens64_%s.npz' % name), **out)
    print('   fidelity min %.6f median %.6f max %.6f; %.1f MB' % (
        out['fidelity'].min(), np.median(out['fidelity']), out['fidelity'].max(),
        os.path.getsize(os.path.join(OUT, 'ens64_%s.npz' % name)) / 1e6))


def main():
    from oracle import refshim
    if not refshim.available():
        raise SystemExit('the reference tree is not present: fixtures can only be generated in the build container')
    names = sys.argv[1:] or ['qubit', 'transmon', 'crosstalk']
    with mp.get_context('spawn').Pool(min(os.cpu_count() or 1, 16)) as pool:
        for name in names:
            build(name, pool)


if __name__ == '__main__':
    main()
