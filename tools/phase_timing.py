#!/usr/bin/env python3
"""GPU box: run the LM_TIMING variant on the headline batch and print per-phase cycle shares."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_classification_icbhi_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "variants", "liblogmel_timing.bin")
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
names = ["stage wait", "FFT part 1", "barrier A", "FFT part 2", "barrier B", "mel", "gather/C", "norm"]
tot = a.sum(axis=2)
print("mean cycles per warp:", tot.mean(), " min/max", tot.min(), tot.max())
items = (B + 295) // 296 * 20
for i, nm in enumerate(names):
    print(f"{nm:12s} {100 * a[:, :, i].sum() / a.sum():5.1f} %   {a[:, :, i].mean() / items:8.0f} cycles/item   (per-warp min {a[:,:,i].min()/items:.0f} max {a[:,:,i].max()/items:.0f})")
print("per-warp of CTA 0 (cycles/item):")
for w in range(16):
    print(w, " ".join(f"{a[0, w, i] / items:7.0f}" for i in range(8)))
