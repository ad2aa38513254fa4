#!/bin/bash
# throughput vs resident warps per SM (M4Q_MAX_WARPS caps the CTA width) at 65,536 and 8,192 members
for n in 65536 8192; do
for w in 8 10 12 13 14 15 16; do
  M4Q_MAX_WARPS=$w python bench.py --workload transmon_h16 --members $n --steps 3 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('members %6d warps %2d traj/s %9.0f ms %8.3f' % ($n, d['config']['launch']['warps_per_cta'], d['value'], d['ms_per_step']))"
done; done
