#!/usr/bin/env python
"""CPU calibration (build container only: needs /root/reference): the reference's own mpc() through oracle/refshim.py
against the numpy port that bench.py times as `cpu_baseline` / `--impl reference`, same members, same leaves, one core.
Writes profiles/r2_cpu_calibration.json."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
os.environ['OMP_NUM_THREADS'] = '1'
import numpy as np
from oracle import make_golden as mg, restate as rs
from mpc4quantum_b200 import systems

out = {}
for name, cfg, maker, n_total in (('transmon_h16', systems.config_transmon(1, discretize=rs.taylor_discretize), systems.ensemble_transmon, 65536),
                                  ('qubit', systems.config_qubit(1, discretize=rs.taylor_discretize), systems.ensemble_qubit, 4096)):
    ens, _ = maker(n_total, lo=0, hi=4)
    t_ref, t_port, gap = [], [], 0.0
    for k in range(4):
        member = ens.member(k)
        t0 = time.perf_counter(); xs_r, us_r, ec, cnt = mg.reference_loop(cfg, plant=member); t_ref.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); xs_p, us_p, ec2, cnt2 = mg.restated_loop(cfg, plant=member); t_port.append(time.perf_counter() - t0)
        gap = max(gap, float(np.abs(us_r - us_p).max()))
    out[name] = {'members': 4, 'reference_mpc_through_shim_s_per_trajectory': float(np.mean(t_ref)),
                 'port_s_per_trajectory': float(np.mean(t_port)), 'port_over_reference': float(np.mean(t_port) / np.mean(t_ref)),
                 'max_us_gap': gap, 'qp_solves_per_trajectory': int(cnt.sum())}
    print(name, out[name])
out['note'] = ('single core of the build container; both loops use the exact active-set QP and scipy expm leaves (cvxpy, OSQP and '
               'qutip are not installable offline).  The port is the faster of the two, so timing the port as the CPU arm does '
               'not inflate the GPU / CPU ratio.')
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'profiles', 'r2_cpu_calibration.json'), 'w') as fh:
    json.dump(out, fh, indent=1)
