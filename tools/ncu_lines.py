#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo and --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]
"""
import collections
import re
import csv
import io
import subprocess
import sys


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, fname, out = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] != '':
            d = {}
            for k, v in zip(hdr, r):
                d.setdefault(k, v)
            d['file'] = fname
            out.append(d)
    stalls = [h for h in dict.fromkeys(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(f(d['# Samples']) for d in out)
    inst = sum(f(d['Instructions Executed']) for d in out)
    agg = collections.Counter()
    for d in out:
        for s in stalls:
            agg[s] += f(d[s])
    print('samples %d  warp-instructions %.3e' % (tot, inst))
    print('stall mix:', ' '.join('%s=%.1f%%' % (k[6:], 100 * v / tot) for k, v in agg.most_common(9)))
    # per-function buckets: nearest preceding `__device__` / `__global__` definition in the current source
    import os
    starts = {}
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'mpc4quantum_b200', 'csrc')
    for fn in ('m4q_core.cuh', 'm4q_kernels.cu'):
        st = []
        for i, line in enumerate(open(os.path.join(here, fn)), 1):
            m = re.match(r'^(?:__device__|__global__).*?(\w+)\(', line)
            if m and not line.startswith('    '):
                st.append((i, m.group(1)))
        starts[fn] = st
    fb, fi = collections.Counter(), collections.Counter()
    for d in out:
        name = d['file']
        for i, nm in starts.get(d['file'], []):
            if int(d['Line No']) >= i:
                name = nm
        fb[name] += f(d['# Samples'])
        fi[name] += f(d['Instructions Executed'])
    print('by function (line ranges of the CURRENT source):')
    for k, v in fb.most_common(16):
        print('   %-22s samp %5.1f%%  inst %5.1f%%' % (k, 100 * v / tot, 100 * fi[k] / inst))
    out.sort(key=lambda d: -f(d['# Samples']))
    for d in out[:top]:
        st = sorted(((f(d[s]), s) for s in stalls), reverse=True)[:3]
        ns = max(f(d['# Samples']), 1)
        print('%-15s %5s samp %5.2f%% inst %5.2f%% conf %6.0fk | %-44s | %s' % (
            d['file'], d['Line No'], 100 * f(d['# Samples']) / tot, 100 * f(d['Instructions Executed']) / inst,
            f(d.get('L1 Wavefronts Shared Excessive', '0')) / 1e3,
            ' '.join('%s:%.0f%%' % (n[6:], 100 * v / ns) for v, n in st), d['Source'].strip()[:64]))


if __name__ == '__main__':
    main()
