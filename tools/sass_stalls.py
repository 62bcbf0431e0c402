#!/usr/bin/env python3
"""Static single-warp issue-time estimate from cuobjdump -sass text: sums the stall field (bits 105..108
of each 128-bit instruction) per address range.  Usage: sass_stalls.py file.sass start_hex end_hex"""
import re, sys
lines = open(sys.argv[1]).read().splitlines()
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
ins = []
i = 0
pat = re.compile(r'/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/')
while i < len(lines):
    m = pat.search(lines[i])
    if m and i + 1 < len(lines):
        m2 = re.search(r'/\* (0x[0-9a-f]{16}) \*/', lines[i + 1])
        if m2:
            addr = int(m.group(1), 16); w1 = int(m2.group(1), 16)
            stall = (w1 >> 41) & 0xf; yld = (w1 >> 45) & 1; wbar = (w1 >> 46) & 7; rbar = (w1 >> 49) & 7; wait = (w1 >> 52) & 0x3f
            ins.append((addr, m.group(2).strip(), stall, yld, wbar, rbar, wait))
            i += 2; continue
    i += 1
sel = [x for x in ins if lo <= x[0] < hi]
tot = sum(x[2] for x in sel)
print(f"{len(sel)} instructions, sum of stall counts {tot}, avg {tot/max(len(sel),1):.2f}")
from collections import Counter
c = Counter()
for x in sel: c[x[1].split()[0].split('.')[0] if not x[1].startswith('@') else x[1].split()[1].split('.')[0]] += x[2]
print(c.most_common(12))
if len(sys.argv) > 4:
    for x in sel: print(f"{x[0]:05x} st={x[2]:2d} y={x[3]} wb={x[4]} rb={x[5]} wait={x[6]:02x}  {x[1][:80]}")
