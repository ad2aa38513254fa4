#!/usr/bin/env python
"""Debug aid (GPU box): replay every QP of the CPU oracle loop on the stand-alone GPU QP kernel."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems, optimize, _lib
from oracle import restate as rs

H = int(sys.argv[1]) if len(sys.argv) > 1 else 100
S = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = systems.config_transmon(1, horizon=H, n_steps=S)
rec = []
def qp_rec(*a):
    out = rs.qp_exact(*a)
    rec.append((a, out))
    return out
plant = rs.ExpmPlant(cfg['experiment'].H0, cfg['experiment'].H1_list, rs.lift_identity, rs.lift_identity)
rs.mpc_loop(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'].dt, H, S, plant,
            cfg['model'].A, cfg['Q'], cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'], warm_start=True, qp=qp_rec)
for k, (a, out) in enumerate(rec):
    x_init, X_ref, U_ref, Q_ls, R_ls, A_ls, B_ls, D_ls, u_prev, sat, du = a
    Uc = out[1]
    lo, hi = rs.qp_bounds(U_ref, u_prev, sat, du)
    nact = int((np.abs(Uc - lo) < 1e-9).sum() + (np.abs(Uc - hi) < 1e-9).sum())
    for st in [dict(), dict(admm_first=1), dict(max_polish=60)]:
        X, U, obj, info = optimize.quad_program(x_init, X_ref, U_ref, Q_ls, R_ls, A_ls, B_ls, D_ls, u_prev, sat, du,
                                                settings=_lib.qp_settings(**st))
        print('qp %d active %d/%d %s status %d iters %s max|dU| %.3e obj %.6e (cpu %.6e)' % (
            k, nact, Uc.size, st, info.status_code, (getattr(info, "admm_iterations", None), getattr(info, "factorizations", None)), np.abs(U - Uc).max(), obj, out[2]))
