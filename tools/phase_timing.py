#!/usr/bin/env python3
"""GPU box: run the LM_TIMING variant on the headline batch and print per-phase cycle shares."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_classification_icbhi_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "variants", f"liblogmel_{os.environ.get('LM_VARIANT', 'timing')}.bin")
from audio_classification_icbhi_b200.plan import LogMelPlan
plan = LogMelPlan(device="cuda:0")
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 80000
clips = torch.randn(B, T, device="cuda") * 0.1
off = torch.arange(B, device="cuda", dtype=torch.int64) * T
ln = torch.full((B,), T, device="cuda", dtype=torch.int32)
out = torch.empty(plan.out_shape(B), device="cuda")
for _ in range(3):
    plan.forward(clips.view(-1), off, ln, out=out)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
n = 148 * 16 * 8
buf = (ctypes.c_longlong * n)()
assert lib.lm_debug_timing(buf, n) == 0
a = np.array(buf, dtype=np.int64).reshape(148, 16, 8)
tot = a.sum(axis=2)
print("mean cycles per warp:", tot.mean(), " min/max", tot.min(), tot.max())
items = (B + 147) // 148 * 20
for role, sl, names in (("mel warps (0-7)", slice(0, 8), ["sched+stage", "wait rows full", "mel/silent", "-", "norm", "end barrier", "-", "-"]),
                        ("FFT warps (8-15)", slice(8, 16), ["wait staged", "load + part 1", "wait rows empty", "part 2", "-", "-", "-", "-"])):
    print(role)
    for i, nm in enumerate(names):
        if nm == "-": continue
        x = a[:, sl, i]
        print(f"  {nm:16s} {100 * x.sum() / a[:, sl, :].sum():5.1f} %   {x.mean() / items:8.0f} cycles/item   (per-warp min {x.min()/items:.0f} max {x.max()/items:.0f})")

tr = (ctypes.c_longlong * 512)()
if hasattr(lib, "lm_debug_trace") and lib.lm_debug_trace(tr, 512) == 0:
    t = np.array(tr, dtype=np.int64).reshape(64, 8)
    t0 = t[0, 0]
    print("CTA 0 trace (cycles since the first bulk copy): item | TMA issued | FFT sees tile | loaded | rows free | rows written | mel sees rows | mel done | iteration end")
    for i in range(24):
        print(f"{i:3d} " + " ".join(f"{(x - t0):9d}" for x in t[i]))
