#!/usr/bin/env python
"""Debug aid (GPU box): which members of the order-1 H=50 transmon ensemble need the ADMM fallback / fail, and where."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems, _lib
H = int(sys.argv[1]) if len(sys.argv) > 1 else 50
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
cfg = systems.config_transmon(1, horizon=H, n_steps=20)
ens, _ = systems.ensemble_transmon(n)
args, kw = systems.mpc_args(cfg); kw.pop('progress_bar')
import warnings; warnings.simplefilter('ignore')
res = m4q.mpc_ensemble(args[0], *args[1:6], ens, *args[7:], fid_target=cfg['target'], **kw)
bad = np.flatnonzero(res.exit_code != 0)
print('exit codes', {int(k): int((res.exit_code == k).sum()) for k in np.unique(res.exit_code)})
adm = res.counters[:, 0]
print('members with ADMM iterations > 0: %d; > 100: %d; > 1000: %d; total admm %d' % ((adm > 0).sum(), (adm > 100).sum(), (adm > 1000).sum(), adm.sum()))
print('factor/qp hist:', np.percentile(res.counters[:, 1] / np.maximum(res.counters[:, 3], 1), [5, 50, 95, 99, 100]))
for k in bad[:14]:
    print('member %d exit %d steps_done %d qp_count %s counters %s' % (k, res.exit_code[k], res.steps_done[k], res.qp_count[k][:res.steps_done[k] + 1], res.counters[k]))
top = np.argsort(-adm)[:8]
for k in top:
    print('heavy member %d exit %d qp_count %s counters %s' % (k, res.exit_code[k], res.qp_count[k], res.counters[k]))
# replay one failing member step by step (host-stepped) to see per-step counters
from mpc4quantum_b200.mpc import ClosedLoopPlan
for k in list(bad[:2]) + list(top[:2]):
    mem = ens.member(int(k))
    for st in (dict(), dict(max_polish=30), dict(admm_first=1)):
        plan = ClosedLoopPlan(cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'], cfg['model'], cfg['Q'],
                              cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'], 3, 0, 100, True, capacity=1, settings=_lib.qp_settings(**st))
        x0 = _lib.dev(cfg['x0'][None], np.complex128); H0 = _lib.dev(mem.H0[None], np.complex128); H1 = _lib.dev(np.stack(mem.H1_list)[None], np.complex128)
        prev = np.zeros(4, dtype=np.int64); rows = []
        for s in range(20):
            r = plan.run(x0, H0, H1, n=1, step_begin=s, step_end=s + 1)
            c = r.counters[0].cpu().numpy().astype(np.int64)
            rows.append((s, int(r.exit_code[0]), int(r.qp_count[0, s]), tuple(c - prev)))
            prev = c
            if int(r.exit_code[0]): break
        print('member %d settings %s:' % (k, st), rows)
