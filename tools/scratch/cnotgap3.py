import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems, _lib
from conftest import load_golden
np.set_printoptions(linewidth=220, precision=1)
g = load_golden('loop_cnot'); cfg = systems.config_cnot(n_steps=40, horizon=50, ramp_steps=200)
args, kw = systems.mpc_args(cfg)
kw = {k: v for k, v in kw.items() if k != 'progress_bar'}
ens = m4q.EnsembleQExperiment(np.asarray(cfg['experiment'].H0)[None], np.array(cfg['experiment'].H1_list)[None], 'identity')
for mode in (1, 3):
    t0 = time.time()
    res = m4q.mpc_ensemble(args[0], *args[1:6], ens, *args[7:], settings=_lib.qp_settings(kkt_fallback=mode), **kw)
    print('kkt_fallback', mode, 'exit', res.exit_code, '%.1f s' % (time.time() - t0), 'counters', res.counters[0])
    print('  us gap per step', np.abs(res.us[0] - g['us']).max(axis=0))
# cost on the order-1 transmon at H = 50
cfg = systems.config_transmon(1, horizon=50, n_steps=20)
ens, _ = systems.ensemble_transmon(65536)
args, kw = systems.mpc_args(cfg); kw = {k: v for k, v in kw.items() if k != 'progress_bar'}
g50 = load_golden('ens64_transmon_h50')
for mode in (1, 3):
    for n in (16, 2048):
        t0 = time.time()
        res = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, n), *args[7:], fid_target=cfg['target'], settings=_lib.qp_settings(kkt_fallback=mode), **kw)
        dt = time.time() - t0
        extra = ''
        if n == 16:
            extra = 'closed-loop us gap vs reference fixture (16 members) %.2e' % np.abs(res.us - g50['us'][:16]).max()
        print('transmon_h50 mode', mode, 'members', n, 'exit', np.bincount(res.exit_code), '%.2f s' % dt, 'mean polish+kkt', res.counters[:, 2].mean(), extra)
