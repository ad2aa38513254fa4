#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: instruction mix (DMMA / DFMA / LDS / STS / LDGSTS / tcgen05 / TMA) of every kernel
in libm4q.so, with registers and spills from the ptxas log of the same build (make -B > build.log 2>&1)."""
import collections, os, re, subprocess, sys, tempfile
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, '..', 'mpc4quantum_b200', 'libm4q.so')
log = sys.argv[1] if len(sys.argv) > 1 else None
regs = {}
if log:
    pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info    : Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info    : Used (\d+) registers")
    for m in pat.finditer(open(log).read()):
        regs[m.group(1)] = (m.group(5), m.group(3), m.group(4))
with tempfile.TemporaryDirectory() as d:
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=d, capture_output=True)
    cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith('.cubin')]
    arch = subprocess.run(['cuobjdump', '-lelf', os.path.abspath(so)], capture_output=True, text=True).stdout.strip()
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
print('# %s' % arch)
print('# per kernel (entry + the __noinline__ device functions it calls, summed): SASS instruction counts; registers / spill bytes from ptxas -v')
kern = collections.defaultdict(collections.Counter)
for k, c in cnt.items():
    entry = k.split('$')[1] if k.startswith('$') else k
    kern[entry].update(c)
ops = ['DMMA', 'DFMA', 'DMUL', 'DADD', 'LDS', 'STS', 'LDGSTS', 'LDG', 'STG', 'LD', 'ST', 'SHFL', 'UTCMMA', 'UTMALDG', 'UBLKCP', 'HMMA']
print('%-64s %7s %5s %9s  %s' % ('kernel', 'instr', 'regs', 'spill st/ld', ' '.join('%6s' % o for o in ops)))
for k in sorted(kern):
    c = kern[k]
    r = regs.get(k, ('?', '?', '?'))
    print('%-64s %7d %5s %4s/%-4s  %s' % (re.sub(r'^_ZN3m4q\d+', '', k)[:64], c['_all'], r[0], r[1], r[2], ' '.join('%6d' % c[o] for o in ops)))
print('# fp64 contraction: mma.sync m8n8k4 f64 (DMMA); tcgen05 (UTCMMA) has no fp64 type, so none is expected; cp.async = LDGSTS')
