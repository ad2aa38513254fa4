#!/usr/bin/env python
"""Short profiling target: one closed loop of N members (default 3,552 = two waves of 148 x 12 warps) of a workload,
run twice (the first launch warms the library and the caches).  Used under ncu -k regex:mpc_kernel -s 1 -c 1."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench
import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import _lib
name = sys.argv[1] if len(sys.argv) > 1 else 'transmon_h16'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3552
cfg, maker = bench.workload(name)
ens, _ = maker(65536)
ens = ens.slice(0, n)
plan = m4q.ClosedLoopPlan(cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'], cfg['model'], cfg['Q'],
                          cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'], d=ens.d, lift_mode=ens.lift_mode,
                          warm_start=cfg['warm_start'], fid_target=cfg['target'], capacity=n)
torch = _lib.torch()
H0, H1 = _lib.dev(ens.H0, np.complex128), _lib.dev(ens.H1, np.complex128)
x0 = _lib.dev((cfg['u0'] if cfg.get('kind') == 'process' else cfg['x0']).reshape(1, -1), np.complex128)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = plan.run(x0, H0, H1, n=n, x0_shared=True)
    e1.record()
    torch.cuda.synchronize()
    print('%s %d members: %.2f ms, exit codes %s, launch %s' % (name, n, e0.elapsed_time(e1), np.unique(res.exit_code.cpu().numpy()), plan.launch_info(n)))
