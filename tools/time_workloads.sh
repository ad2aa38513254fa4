#!/bin/bash
# wall time and throughput of single bench workloads (1 timed pass), to keep the default bench run within minutes
for spec in "$@"; do
  name=${spec%%:*}; n=${spec##*:}
  s=$(date +%s.%N)
  python bench.py --workload $name --members-total $n --steps 1 --warmup 1 --no-extra --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-18s members %6d traj/s %9.0f ms %9.1f frac %.3f fac/qp %.2f admm/qp %.2f exit %s' % ('$name', $n, d['value'], d['ms_per_step'], d['roofline']['frac'], d['factorizations_per_qp'], d['admm_iterations_per_qp'], d['exit_codes']))"
  e=$(date +%s.%N); python -c "print(\"   wall %.1f s\" % ($e - $s))"
done
