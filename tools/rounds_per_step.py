#!/usr/bin/env python
"""Measurement aid (GPU box): active-set rounds per QP, MPC step by MPC step (transmon_h16)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems

ens, _ = systems.ensemble_transmon(65536)
prev = None
for S in (1, 2, 3, 5, 10, 20):
    cfg = systems.config_transmon(1, horizon=16, n_steps=S)
    args, kw = systems.mpc_args(cfg)
    kw.pop('progress_bar')
    res = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, 512), *args[7:], fid_target=cfg['target'], **kw)
    tot = res.counters.sum(axis=0).astype(float)
    cur = np.array([tot[3], tot[2], tot[1], tot[0]]) / 512
    d = cur if prev is None else cur - prev[1]
    print('steps 0..%2d: per member: QPs %6.1f rounds %6.1f factors %6.1f admm %5.1f | added since S=%s: QPs %5.1f rounds %5.1f -> rounds/QP %.2f'
          % (S - 1, cur[0], cur[1], cur[2], cur[3], 0 if prev is None else prev[0], d[0], d[1], d[1] / max(d[0], 1e-9)))
    prev = (S, cur)
