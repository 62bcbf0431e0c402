#!/usr/bin/env python3
"""How much do the two 8-warp groups of a CTA slow each other down?  B = 148 clips keeps group 1 idle,
B = 296 gives every group exactly one clip."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_classification_icbhi_b200.plan import LogMelPlan
plan = LogMelPlan(device="cuda:0")
T = 80000
for st in (0, 1000, 2000, 3500):
    plan.set("stagger_ns", st)
    for B in (296, 4096):
        clips = torch.randn(B, T, device="cuda") * 0.1
        off = torch.arange(B, device="cuda", dtype=torch.int64) * T
        ln = torch.full((B,), T, device="cuda", dtype=torch.int32)
        out = torch.empty(plan.out_shape(B), device="cuda")
        for _ in range(5):
            plan.forward(clips.view(-1), off, ln, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            plan.forward(clips.view(-1), off, ln, out=out)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        print(f"stagger {st:5d}  B {B:5d}: {us:9.1f} us  -> {us / ((B + 295) // 296):8.1f} us per clip-round", flush=True)
