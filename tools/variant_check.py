#!/usr/bin/env python3
"""GPU box: a variant library (tools/variants/liblogmel_<name>.bin) against the shipped one on the headline batch, a ragged
batch, and small batches: max |difference| and time.  Usage: python tools/variant_check.py name [name ...]"""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if not (len(sys.argv) == 2 and sys.argv[1].startswith("--one=")):
    for name in ["ship"] + sys.argv[1:]:
        r = subprocess.run([sys.executable, __file__, f"--one={name}"], capture_output=True, text=True, timeout=120)
        print(r.stdout.strip() or r.stderr[-800:], flush=True)
    sys.exit(0)
name = sys.argv[1].split("=", 1)[1]
sys.path.insert(0, ROOT)
import numpy as np, torch
from audio_classification_icbhi_b200 import _lib
if name != "ship":
    _lib.LIB_PATH = os.path.join(ROOT, "tools", "variants", f"liblogmel_{name}.bin")
from audio_classification_icbhi_b200.plan import LogMelPlan
plan = LogMelPlan(device="cuda:0")
T = 80000
g = torch.Generator(device="cuda").manual_seed(7)
res = {}
for tag, B in (("headline", 4096), ("ragged", 3000), ("small", 32), ("one", 1)):
    if tag == "ragged":
        rs = np.random.RandomState(0)
        lens = (np.clip(rs.lognormal(np.log(2.5), 0.5, B), 0.0, 16.2) * 16000).astype(np.int64); lens[::97] = 0
        starts = np.concatenate([[0], np.cumsum((lens + 3) // 4 * 4)[:-1]])
        wave = torch.randn(int(starts[-1] + lens[-1]) + 8, generator=g, device="cuda") * 0.1
        off = torch.from_numpy(starts).cuda(); ln = torch.from_numpy(lens.astype(np.int32)).cuda()
    else:
        wave = torch.randn(B * T, generator=g, device="cuda") * 0.1
        off = torch.arange(B, device="cuda", dtype=torch.int64) * T
        ln = torch.full((B,), T, device="cuda", dtype=torch.int32)
    out = torch.empty(plan.out_shape(B), device="cuda")
    for _ in range(3): plan.forward(wave, off, ln, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): plan.forward(wave, off, ln, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    path = f"/tmp/variant_ref_{tag}.pt"
    if name == "ship":
        torch.save(out.cpu(), path); diff = 0.0
    else:
        diff = float((out.cpu() - torch.load(path)).abs().max())
    res[tag] = f"{ms:.4f} ms  max|diff| {diff:.2e}  finite {bool(torch.isfinite(out).all())}"
print(name, json.dumps(res))
