"""Order-1 transmon ensemble at H = 100: exit codes, solver counters and wall time (n members, S steps)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems, _lib
n, S = int(sys.argv[1]), int(sys.argv[2])
H = int(sys.argv[3]) if len(sys.argv) > 3 else 100
cfg = systems.config_transmon(1, horizon=H, n_steps=S)
ens, _ = systems.ensemble_transmon(65536)
args, kw = systems.mpc_args(cfg)
kw.pop('progress_bar')
torch = _lib.torch()
t0 = time.time()
res = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, n), *args[7:], fid_target=cfg['target'], **kw)
print('%d members, %d steps, H = %d: %.1f s' % (n, S, H, time.time() - t0))
print('exit codes', np.bincount(res.exit_code, minlength=4), 'steps done', np.bincount(res.steps_done, minlength=S + 1))
bad = np.flatnonzero(res.exit_code != 0)
print('bad members', bad[:20], 'their steps', res.steps_done[bad][:20])
c = res.counters
print('counters mean [admm, factor, polish+kkt, solves]', c.mean(axis=0), 'max', c.max(axis=0))
print('fidelity median %.5f min %.5f' % (np.median(res.fidelity), res.fidelity.min()))
