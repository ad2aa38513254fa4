"""Host-stepped run of single members (order-1 transmon, H = 100); on a non-zero exit code dump what the controller held
when the failing step started (guesses, measured state, previous controls) to gpurun_out/ for analysis on the CPU."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems, _lib
from mpc4quantum_b200.mpc import ClosedLoopPlan
members = [int(v) for v in sys.argv[1:]]
H, S = 100, 20
cfg = systems.config_transmon(1, horizon=H, n_steps=S)
ens, _ = systems.ensemble_transmon(65536)
torch = _lib.torch()
for k in members:
    mem = ens.slice(k, k + 1)
    plan = ClosedLoopPlan(cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'], cfg['model'], cfg['Q'],
                          cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'], d=3, max_iter=100, warm_start=cfg['warm_start'],
                          capacity=1)
    H0 = _lib.dev(mem.H0, np.complex128); H1 = _lib.dev(mem.H1, np.complex128)
    x0 = _lib.dev(np.asarray(cfg['x0']).reshape(1, -1), np.complex128)
    for s in range(S):
        before = plan.state.clone()
        res = plan.run(x0, H0, H1, n=1, x0_shared=True, step_begin=s, step_end=s + 1)
        ec = int(res.exit_code[0])
        if ec != 0:
            st = before.view(torch.float64).cpu().numpy()
            N = 18
            Xg = st[:(H + 1) * N].reshape(H + 1, N); Ug = st[(H + 1) * N:(H + 1) * N + H * 2].reshape(H, 2)
            np.savez(os.path.join(ROOT, 'gpurun_out', 'h100_fail_dev_m%d.npz' % k), step=s, Xg=Xg, Ug=Ug,
                     xs=res.xs[0].cpu().numpy(), us=res.us[0].cpu().numpy(), exit_code=ec,
                     counters=res.counters[0].cpu().numpy())
            print('member %d: exit %d at step %d, counters %s' % (k, ec, s, res.counters[0].cpu().numpy()))
            break
    else:
        print('member %d: all %d steps exit 0 (host-stepped)' % (k, S))
