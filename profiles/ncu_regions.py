#!/usr/bin/env python3
"""Aggregates an `ncu --page source --csv` export (SASS view) of the log-mel kernel into
address regions, so that instruction counts, shared-memory wavefronts and stall samples can be
read per phase.  Usage: ncu -i X.ncu-rep --page source --csv | python profiles/ncu_regions.py [nframes]
"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = rows[hdr_i + 1:]
nframes = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0


def num(r, name):
    try:
        return float(r[col[name]])
    except Exception:
        return 0.0


# region boundaries: split at barriers and at loop back-edges so phases separate
regions = []
cur = {"start": None, "inst": 0, "samples": 0, "wf": 0, "wf_ex": 0, "n": 0, "ops": {}}
for r in data:
    if len(r) < len(hdr):
        continue
    src = r[col["Source"]]
    op = src.split()[0] if src else ""
    if op.startswith("@"):
        op = src.split()[1] if len(src.split()) > 1 else op
    if cur["start"] is None:
        cur["start"] = r[col["Address"]]
    cur["inst"] += num(r, "Instructions Executed")
    cur["samples"] += num(r, "# Samples")
    cur["wf"] += num(r, "L1 Wavefronts Shared")
    cur["wf_ex"] += num(r, "L1 Wavefronts Shared Excessive")
    cur["n"] += 1
    base = op.split(".")[0]
    cur["ops"][base] = cur["ops"].get(base, 0) + num(r, "Instructions Executed")
    if base in ("BAR", "EXIT") or (base == "BRA" and num(r, "Instructions Executed") > 0 and cur["n"] > 40):
        cur["end"] = r[col["Address"]]
        regions.append(cur)
        cur = {"start": None, "inst": 0, "samples": 0, "wf": 0, "wf_ex": 0, "n": 0, "ops": {}}
if cur["n"]:
    cur["end"] = data[-1][col["Address"]]
    regions.append(cur)

tot_inst = sum(r["inst"] for r in regions)
tot_s = sum(r["samples"] for r in regions)
print(f"total warp-instructions {tot_inst:.0f}  ({tot_inst / nframes:.1f} per frame), samples {tot_s:.0f}")
print(f"{'region':>22} {'sass':>5} {'inst/frame':>10} {'inst%':>6} {'samp%':>6} {'wf/frame':>9} {'wf_exc/frame':>12}  top ops (per frame)")
for r in regions:
    if r["inst"] < 0.002 * tot_inst and r["samples"] < 0.002 * tot_s:
        continue
    top = sorted(r["ops"].items(), key=lambda kv: -kv[1])[:7]
    tops = " ".join(f"{k}:{v / nframes:.0f}" for k, v in top)
    print(f"{r['start'][-6:]:>10}-{r['end'][-6:]:>10} {r['n']:5d} {r['inst'] / nframes:10.1f} {100 * r['inst'] / tot_inst:6.1f} "
          f"{100 * r['samples'] / max(tot_s, 1):6.1f} {r['wf'] / nframes:9.1f} {r['wf_ex'] / nframes:12.1f}  {tops}")
