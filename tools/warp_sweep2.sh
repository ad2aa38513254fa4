#!/bin/bash
# throughput vs resident warps per SM at two horizons: does the 16-warp configuration lose to L2 capacity?
# (per-warp workspace 42 KB at H = 16, 22 KB at H = 8; 148 x 16 warps x 42 KB = 100 MB)
for h in 8 16; do
for w in 4 8 12 16; do
  M4Q_MAX_WARPS=$w python bench.py --workload transmon_h$h --members 65536 --steps 3 --warmup 2 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('H %3d warps %2d traj/s %9.0f ms %8.3f frac %.3f' % ($h, d['config']['launch']['warps_per_cta'], d['value'], d['ms_per_step'], d['roofline']['frac']))"
done; done
