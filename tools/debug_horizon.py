#!/usr/bin/env python
"""Debug aid (GPU box): nominal transmon at a long horizon -- GPU loop vs the CPU oracle loop, QP by QP."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems
from oracle import restate as rs

H = int(sys.argv[1]) if len(sys.argv) > 1 else 100
S = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = systems.config_transmon(1, horizon=H, n_steps=S)
args, kw = systems.mpc_args(cfg)
(xs, us), model, ec = m4q.mpc(*args, **kw)
print('GPU exit', ec, 'us', None if us is None else us[:, :S])
t0 = time.time()
mem = cfg['nominal']
plant = rs.ExpmPlant(cfg['experiment'].H0, cfg['experiment'].H1_list, rs.lift_identity, rs.lift_identity)
stats = {}
xs_c, us_c, ec_c = rs.mpc_loop(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'].dt, H, S,
                               plant, cfg['model'].A, cfg['Q'], cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'],
                               warm_start=cfg['warm_start'], measure_freq=1, stats=stats)
print('CPU exit', ec_c, 'qp per step', stats['qp_per_step'], 'time', time.time() - t0)
print('CPU us', us_c[:, :S])
if us is not None:
    print('max |du|', np.abs(us - us_c).max())
# settings sweep on the GPU
from mpc4quantum_b200 import _lib
from mpc4quantum_b200.mpc import ClosedLoopPlan
ex = cfg['experiment']
for st in [dict(), dict(admm_first=1), dict(max_polish=40), dict(rho=1e-3), dict(rho=10.0), dict(rho=1e-3, admm_first=1, max_polish=40)]:
    plan = ClosedLoopPlan(cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'], cfg['model'], cfg['Q'],
                          cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'], ex.H0.shape[0], ex.lift_mode, 100, cfg['warm_start'],
                          capacity=1, settings=_lib.qp_settings(**st))
    res = plan.run(_lib.dev(cfg['x0'][None], np.complex128), _lib.dev(ex.H0[None], np.complex128),
                   _lib.dev(np.stack(ex.H1_list)[None], np.complex128), n=1).numpy()
    sd = int(res.steps_done[0])
    print(st, 'exit', int(res.exit_code[0]), 'steps', sd, 'qp_count', res.qp_count[0][:S], 'counters', res.counters[0],
          'max|du|', np.abs(res.us[0][:, :sd] - us_c[:, :sd]).max() if sd else None)
