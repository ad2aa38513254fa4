#!/bin/bash
# bench.py over the secondary workloads (16,384 members, no CPU leg); one formatted line per workload.
for w in qubit crosstalk not_gate transmon_h10 transmon_h16 transmon_h20 transmon_h50 transmon_o2_h10 transmon_o2_h16 transmon_o2_h20 transmon_o2_h50 transmon_o2_h100 transmon_models_h16 transmon_exact_h16; do
  python bench.py --workload $w --members 16384 --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-20s traj/s %9.0f  qp/s %10.0f  warps/SM %2d  qp/traj %6.1f  fac/qp %.2f  admm/qp %6.2f  fp64 frac %.3f  exit %s  median fid %.6f' % (
  '$w', d['value'], d['qp_solves_per_s'], d['config']['launch']['warps_per_cta'], d['qp_solves_per_trajectory'], d['factorizations_per_qp'],
  d['admm_iterations_per_qp'], d['roofline']['frac'], d['exit_codes'], d['fidelity']['median']))"
done
