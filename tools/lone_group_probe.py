#!/usr/bin/env python3
"""GPU box: time a library (shipped or tools/variants/<name>.bin) at B = 148 (one 8-warp group per SM busy, one clip
each), B = 296 (both groups, one clip each) and the headline batch.  Usage: python tools/lone_group_probe.py [variant ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_classification_icbhi_b200 import _lib
default_path = _lib.LIB_PATH
T = 80000
for name in ["ship"] + sys.argv[1:]:
    _lib._lib = None
    _lib.LIB_PATH = default_path if name == "ship" else os.path.join(ROOT, "tools", "variants", f"liblogmel_{name}.bin")
    from audio_classification_icbhi_b200.plan import LogMelPlan
    plan = LogMelPlan(device="cuda:0")
    plan.set("split", 1)   # no small-batch splitting: one clip per group
    for B in (148, 296, 4096):
        clips = torch.randn(B, T, device="cuda") * 0.1
        off = torch.arange(B, device="cuda", dtype=torch.int64) * T
        ln = torch.full((B,), T, device="cuda", dtype=torch.int32)
        out = torch.empty(plan.out_shape(B), device="cuda")
        for _ in range(5):
            plan.forward(clips.view(-1), off, ln, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            plan.forward(clips.view(-1), off, ln, out=out)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        rounds = (B + 295) // 296
        print(f"{name:10s} B {B:5d}: {us:9.1f} us  -> {us / rounds:8.1f} us per clip-round = {us / rounds * 1965 / 20:7.0f} cycles per tile", flush=True)
    del plan
