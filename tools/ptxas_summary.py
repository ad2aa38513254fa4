#!/usr/bin/env python
"""Registers / stack / spills per kernel from a `make -B` log (nvcc -Xptxas -v)."""
import re, sys
txt = open(sys.argv[1]).read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info    : Function properties for \S+\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info    : Used (\d+) registers")
for m in pat.finditer(txt):
    name = m.group(1)
    short = re.sub(r'^_ZN3m4q\d+', '', name)
    print('%-60s regs %3s stack %4s spill st/ld %3s/%3s' % (short[:60], m.group(5), m.group(2), m.group(3), m.group(4)))
