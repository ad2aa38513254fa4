"""Teacher-forced replay of the H = 100 order-1 fixture: per member / step control gap, SQP counts, solver counters."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
from mpc4quantum_b200 import systems, _lib
from mpc4quantum_b200.mpc import ClosedLoopPlan
from conftest import load_golden
from test_gpu_parity64 import _realify_traj
np.set_printoptions(linewidth=250, precision=1)
name = sys.argv[1] if len(sys.argv) > 1 else 'transmon_h100'
H = int(name.split('_h')[-1])
g = load_golden('ens64_' + name)
cfg = systems.config_transmon(1, horizon=H, n_steps=20)
k, S, c, H1 = g['tf_Xg'].shape
m = cfg['dim_u']; N = 2 * c
plan = ClosedLoopPlan(cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'], cfg['model'], cfg['Q'],
                      cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'], d=0, max_iter=100, warm_start=cfg['warm_start'],
                      capacity=k, external_plant=True)
torch = _lib.torch()
state = plan.state.view(torch.float64).view(k, -1)
x0 = _lib.dev(g['tf_x'][:, 0], np.complex128)
gap = np.zeros((k, S)); ec = np.zeros((k, S), int); cnt = np.zeros((k, S, 4), int)
import time
for s in range(S):
    plan.xs[:, :, s] = _lib.dev(g['tf_x'][:, s], np.complex128)
    if s > 0:
        state[:, :(H + 1) * N] = _lib.dev(_realify_traj(g['tf_Xg'][:, s]).reshape(k, -1), np.float64)
        state[:, (H + 1) * N:(H + 1) * N + H * m] = _lib.dev(np.transpose(g['tf_Ug'][:, s], (0, 2, 1)).reshape(k, -1), np.float64)
        plan.us[:, :, s - 1] = _lib.dev(g['tf_us'][:, :, s - 1], np.float64)
    plan.counters.zero_()
    plan.exit_code.zero_()
    torch.cuda.synchronize(); t0 = time.time()
    res = plan.run(x0, n=k, step_begin=s, step_end=s + 1)
    torch.cuda.synchronize(); dt = time.time() - t0
    gap[:, s] = np.abs(res.us[:, :, s].cpu().numpy() - g['tf_us'][:, :, s]).max(axis=1)
    ec[:, s] = res.exit_code.cpu().numpy()
    cnt[:, s] = res.counters.cpu().numpy()
    print('step %2d %.3f s  gap max %.1e  exit %s  kkt+polish rounds %s' % (s, dt, gap[:, s].max(), ec[:, s], cnt[:, s, 2]))
print('gap per member/step'); print(gap)
