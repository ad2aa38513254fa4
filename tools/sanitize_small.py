#!/usr/bin/env python
"""Small closed-loop run for compute-sanitizer (memcheck / racecheck): a few members, a few MPC steps."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems

which = sys.argv[1] if len(sys.argv) > 1 else 'transmon'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
if which == 'transmon':
    cfg, maker = systems.config_transmon(1, horizon=8, n_steps=3), systems.ensemble_transmon
elif which == 'crosstalk':
    cfg, maker = systems.config_crosstalk(0.0, n_steps=4), systems.ensemble_crosstalk
else:
    cfg, maker = systems.config_qubit(1), systems.ensemble_qubit
ens, _ = maker(4096)
args, kw = systems.mpc_args(cfg)
kw.pop('progress_bar')
res = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, n), *args[7:], fid_target=cfg['target'], **kw)
print(which, 'exit codes', np.unique(res.exit_code), 'qp', res.qp_count.sum(axis=1)[:4], 'fid', res.fidelity[:3])
