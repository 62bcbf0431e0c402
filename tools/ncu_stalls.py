#!/usr/bin/env python3
"""Per-region stall-reason breakdown from an `ncu --page source --csv` export.
Usage: python tools/ncu_stalls.py file.csv [nframes]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
nframes = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def num(r, n):
    try: return float(r[col[n]])
    except Exception: return 0.0
regions = []; cur = None
def new(): return {"start": None, "n": 0, "inst": 0.0, "samples": 0.0, **{k: 0.0 for k in reasons}}
cur = new()
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    src = r[col['Source']]; op = src.split()[0] if src else ''
    if op.startswith('@'): op = src.split()[1]
    base = op.split('.')[0]
    if cur["start"] is None: cur["start"] = r[col['Address']][-6:]
    cur["n"] += 1; cur["inst"] += num(r, 'Instructions Executed'); cur["samples"] += num(r, '# Samples')
    for k in reasons: cur[k] += num(r, k)
    if base in ('BAR', 'EXIT') or (base == 'BRA' and num(r, 'Instructions Executed') > 0 and cur["n"] > 40):
        regions.append(cur); cur = new()
if cur["n"]: regions.append(cur)
tot = sum(r["samples"] for r in regions)
print("region   inst/frame samp%  " + " ".join(k[6:12].rjust(6) for k in reasons))
for r in regions:
    if r["samples"] < 0.004 * tot: continue
    print(f"{r['start']:>8} {r['inst']/nframes:9.1f} {100*r['samples']/tot:5.1f}  " + " ".join(f"{100*r[k]/max(r['samples'],1):6.1f}" for k in reasons))
