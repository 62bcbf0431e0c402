#!/usr/bin/env python3
"""GPU box: end-to-end host API throughput vs chunk size, fp32 and int16 PCM input."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_classification_icbhi_b200.plan import LogMelPlan
plan = LogMelPlan(device="cuda:0")
B, T = 4096, 80000
x = (torch.randn(B, T) * 0.1)
f32 = x.pin_memory()
pcm = (x * 32768.0).round_().clamp_(-32768, 32767).to(torch.int16).pin_memory()
off = torch.arange(B, dtype=torch.int64) * T
ln = torch.full((B,), T, dtype=torch.int32)
out = torch.empty(plan.out_shape(B), dtype=torch.float32).pin_memory()
for chunk in (0, 148, 296, 592, 1184, 2048):
    plan.set("host_chunk_clips", chunk)
    for name, w in (("f32", f32), ("pcm16", pcm)):
        for _ in range(2):
            plan.forward_host(w.view(-1), off, ln, out=out)
        t0 = time.perf_counter()
        for _ in range(5):
            plan.forward_host(w.view(-1), off, ln, out=out)
        dt = (time.perf_counter() - t0) / 5
        print(f"chunk {chunk:5d} {name:6s}: {dt*1e3:7.2f} ms  {B/dt:10,.0f} clips/s", flush=True)
