#!/usr/bin/env python
"""Per-member comparison of the exact-model closed loop (GPU) with the CPU oracle: fidelity gap, SQP counts."""
import os
import sys
import multiprocessing as mp

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench                                   # noqa: E402
import mpc4quantum_b200 as m4q                 # noqa: E402
from mpc4quantum_b200 import systems           # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 144
    name = 'transmon_exact_h16'
    cfg, maker = bench.workload(name)
    ens, _ = maker(16384)
    args, kw = systems.mpc_args(cfg)
    kw.pop('progress_bar')
    res = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, n), *args[7:], fid_target=cfg['target'], **kw)
    with mp.get_context('spawn').Pool(16) as pool:
        dt, cpu = bench.cpu_pass(name, range(n), 16384, pool)
    gaps = np.array([abs(res.fidelity[k] - cpu[k][0]) for k in range(n)])
    qg = res.qp_count.sum(axis=1)
    for k in np.argsort(-gaps)[:10]:
        print('member %4d  gap %.3e  fid gpu %.9f cpu %.9f  qp gpu %d cpu %d  exit %d' %
              (k, gaps[k], res.fidelity[k], cpu[k][0], qg[k], cpu[k][1], res.exit_code[k]))
    print('members with different SQP counts:', int(sum(qg[k] != cpu[k][1] for k in range(n))), 'of', n)
    print('max gap among members with equal counts: %.3e' % max([gaps[k] for k in range(n) if qg[k] == cpu[k][1]] + [0]))


if __name__ == '__main__':
    main()
