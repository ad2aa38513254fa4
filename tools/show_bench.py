#!/usr/bin/env python
"""One formatted line per bench.py JSON file."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    r = d.get('roofline', {})
    print('%-28s traj/s %9.0f e2e %9.0f qp/s %10.0f warps %2s qp/traj %6.1f fac/qp %.2f admm/qp %6.2f frac %.3f exit %s fid med %.6f' % (
        f.split('/')[-1], d['value'], d.get('e2e', {}).get('value', 0), d.get('qp_solves_per_s', 0), d['config'].get('launch', {}).get('warps_per_cta'),
        d.get('qp_solves_per_trajectory', 0), d.get('factorizations_per_qp', 0), d.get('admm_iterations_per_qp', 0), r.get('frac', 0), d.get('exit_codes'), d.get('fidelity', {}).get('median', 0)))
