#!/usr/bin/env python
"""OSQP-equivalent mode (polish = 0) on the golden QPs: accuracy against the exact oracle vs eps / iteration cap."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import optimize, _lib
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'qp.npz'))
for adaptive in (0, -1):
    for eps, cap in ((1e-5, 500), (1e-6, 500), (1e-7, 2000), (1e-8, 5000)):
        for tag in ('qubit', 'transmon', 'cross'):
            n = g['%s_x_init' % tag].shape[0]
            H = g['%s_U' % tag].shape[2]
            Q_ls = [g['%s_Q' % tag]] * H + [g['%s_Qf' % tag]]
            R_ls = [g['%s_R' % tag]] * H
            st = _lib.qp_settings(polish=0, max_admm=cap, eps=eps, adaptive_rho=adaptive)
            worst, its, facs = 0.0, [], []
            for i in range(n):
                X, U, obj, info = optimize.quad_program(
                    g['%s_x_init' % tag][i], g['%s_X_bm' % tag][i], g['%s_U_bm' % tag][i], Q_ls, R_ls,
                    list(g['%s_A' % tag][i]), list(g['%s_B' % tag][i]), list(g['%s_D' % tag][i]), g['%s_u_prev' % tag][i],
                    float(g['%s_sat' % tag]), float(g['%s_du' % tag]), settings=st)
                worst = max(worst, np.abs(U - g['%s_U' % tag][i]).max())
                its.append(info.admm_iterations); facs.append(info.factorizations)
            print('adaptive %2d eps %.0e cap %5d %-8s max|dU| %.2e  iterations %s  factorizations %s' % (adaptive, eps, cap, tag, worst, its, facs))
