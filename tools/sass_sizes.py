#!/usr/bin/env python
"""SASS instruction counts per device function of one kernel instantiation (hot code must fit the 32 KB L1.5 I-cache).

    python tools/sass_sizes.py [Li9ELi2]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, '..', 'mpc4quantum_b200', 'libm4q.so')
key = sys.argv[1] if len(sys.argv) > 1 else 'Li9ELi2'
with tempfile.TemporaryDirectory() as d:
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=d, capture_output=True)
    cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith('.cubin')]
    txt = subprocess.run(['nvdisasm'] + cub, capture_output=True, text=True).stdout
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.match(r'^(\S+):\s*$', line)
    if m and (m.group(1).startswith('_Z') or m.group(1).startswith('$')):
        cur = m.group(1)
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        cnt[cur][m.group(1).split('.')[0]] += 1
        cnt[cur]['_all'] += 1
tot = 0
for k, c in cnt.items():
    if key in k and 'mpc_kernel' in k:
        nm = k.split('$')[-1] if '$' in k else 'KERNEL'
        nm = re.sub(r'^_ZN3m4q\d+', '', nm)[:34]
        print('%-36s %5d  %s' % (nm, c['_all'], {o: c[o] for o in ['LDS', 'STS', 'LD', 'DFMA', 'DMMA', 'SHFL', 'BRA', 'CALL'] if c[o]}))
        tot += c['_all']
print('total %d instructions = %.1f KB' % (tot, tot * 16 / 1024))
