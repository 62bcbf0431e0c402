#!/usr/bin/env python3
"""GPU box: the headline batch as 16-bit PCM, device-resident: decode kernel + log-mel kernel vs the fused lm_forward_pcm16."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_classification_icbhi_b200.plan import LogMelPlan
plan = LogMelPlan(device="cuda:0")
B, T = 4096, 80000
pcm = (torch.randn(B * T, device="cuda") * 3000).clamp(-32768, 32767).to(torch.int16)
off = torch.arange(B, device="cuda", dtype=torch.int64) * T
ln = torch.full((B,), T, device="cuda", dtype=torch.int32)
out = torch.empty(plan.out_shape(B), device="cuda")
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
dec = plan.pcm16_decode(pcm)
t_f32 = timeit(lambda: plan.forward(dec, off, ln, out=out))
t_two = timeit(lambda: plan.forward(plan.pcm16_decode(pcm), off, ln, out=out))
want = out.clone()
t_fused = timeit(lambda: plan.forward_pcm16(pcm, off, ln, out=out))
print(f"fp32 input, one kernel           : {t_f32:.3f} ms  ({B / t_f32 * 1e3 / 1e6:.3f} M clips/s)")
print(f"int16 input, decode + log-mel     : {t_two:.3f} ms  ({B / t_two * 1e3 / 1e6:.3f} M clips/s)")
print(f"int16 input, fused lm_forward_pcm16: {t_fused:.3f} ms  ({B / t_fused * 1e3 / 1e6:.3f} M clips/s)   bit-identical: {torch.equal(out, want)}")
