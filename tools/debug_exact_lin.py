#!/usr/bin/env python
"""Accuracy of the device exact linearisation against scipy (expm, expm_frechet) and an mpmath-free refinement:
scipy's own result is cross-checked by evaluating the augmented block-triangular exponential expm([[G, L],[0, G]])."""
import os
import sys
import numpy as np
from scipy.linalg import expm

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mpc4quantum_b200 as m4q                 # noqa: E402
from mpc4quantum_b200 import systems           # noqa: E402
from oracle import restate as rs               # noqa: E402

cfg = systems.config_transmon_exact()
L = list(cfg['model'].generators)
dt = cfg['clock'].dt
c, m, H = 9, 2, 16
rng = np.random.default_rng(3)
for scale in (0.1, 0.5, 1.0, 1.57):
    X = np.tile(cfg['x0'][:, None], (1, H + 1)) + 0.1 * (rng.normal(size=(c, H + 1)) + 1j * rng.normal(size=(c, H + 1)))
    U = rng.uniform(-scale, scale, size=(m, H))
    A, B, D = cfg['model'].get_model_along_traj(X, U)
    A2, B2, D2 = rs.ExactModel(L, dt).along(X, U, H)
    # augmented exponential as a third opinion
    B3 = []
    for t in range(H):
        G = (L[0] + U[0, t] * L[1] + U[1, t] * L[2]) * dt
        cols = []
        for i in range(m):
            aug = np.block([[G, L[1 + i] * dt], [np.zeros_like(G), G]])
            cols.append(expm(aug)[:c, c:] @ X[:, t])
        B3.append(np.stack(cols, axis=1))
    print('|u| <= %.2f: A gpu-scipy %.1e | B gpu-scipy %.1e  gpu-aug %.1e  scipy-aug %.1e | ||G dt||_1 max %.2f' % (
        scale, np.abs(np.array(A) - np.array(A2)).max(), np.abs(np.array(B) - np.array(B2)).max(),
        np.abs(np.array(B) - np.array(B3)).max(), np.abs(np.array(B2) - np.array(B3)).max(),
        max(np.abs((L[0] + U[0, t] * L[1] + U[1, t] * L[2]) * dt).sum(axis=0).max() for t in range(H))))
