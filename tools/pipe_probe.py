#!/usr/bin/env python3
"""GPU box: the pipeline kernel (plan.set("pipeline", 1)) against the single-kernel path (0): bit-identity and time.
Usage: python tools/pipe_probe.py [B] [ragged]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from audio_classification_icbhi_b200.plan import LogMelPlan
plan = LogMelPlan(device="cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = 80000
ragged = len(sys.argv) > 2
g = torch.Generator(device="cuda").manual_seed(7)
if ragged:
    rs = np.random.RandomState(0)
    lens = (np.clip(rs.lognormal(np.log(2.5), 0.5, B), 0.0, 16.2) * 16000).astype(np.int64)
    lens[::97] = 0
    starts = np.concatenate([[0], np.cumsum((lens + 3) // 4 * 4)[:-1]])
    wave = torch.randn(int(starts[-1] + lens[-1]) + 8, generator=g, device="cuda") * 0.1
    off = torch.from_numpy(starts).cuda(); ln = torch.from_numpy(lens.astype(np.int32)).cuda()
else:
    wave = torch.randn(B * T, generator=g, device="cuda") * 0.1
    off = torch.arange(B, device="cuda", dtype=torch.int64) * T
    ln = torch.full((B,), T, device="cuda", dtype=torch.int32)
def run(mode, n=20):
    plan.set("pipeline", mode)
    out = torch.full(plan.out_shape(B), float("nan"), device="cuda")
    for _ in range(3): plan.forward(wave, off, ln, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): plan.forward(wave, off, ln, out=out)
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n
ref, t0 = run(0)
print(f"single kernel : {t0:.3f} ms  ({B / t0 * 1e3 / 1e6:.3f} M clips/s)", flush=True)
got, t1 = run(1)
same = torch.equal(got, ref)
print(f"pipeline      : {t1:.3f} ms  ({B / t1 * 1e3 / 1e6:.3f} M clips/s)   bit-identical: {same}", flush=True)
if not same:
    d = (got - ref).abs()
    bad = torch.nonzero(~torch.isclose(got, ref, rtol=0, atol=0, equal_nan=True))
    print("mismatches:", bad.shape[0], "max abs diff", float(torch.nan_to_num(d, nan=1e9).max()), "nan in pipeline:", int(torch.isnan(got).sum()))
    print("first:", bad[:5].tolist())
