#!/usr/bin/env python3
"""GPU box: BASELINE configs[2] -- 6900 respiratory cycles of lognormal length (0.2 .. 16.2 s), pad / crop to 5 s --
timed on one GPU.  Usage: python tools/ragged_probe.py [variant]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_classification_icbhi_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.join(ROOT, "tools", "variants", f"liblogmel_{sys.argv[1]}.bin")
from audio_classification_icbhi_b200.plan import LogMelPlan
plan = LogMelPlan(device="cuda:0"); dev = plan.device
n = 6900
rs = np.random.RandomState(0)
secs = np.clip(rs.lognormal(np.log(2.5), 0.5, n), 0.2, 16.2)
lens = (secs * 16000).astype(np.int64)
starts = np.concatenate([[0], np.cumsum((lens + 3) // 4 * 4)[:-1]])
wave = torch.randn(int(starts[-1] + lens[-1]) + 4, device=dev) * 0.1
offset = torch.from_numpy(starts).to(dev); length = torch.from_numpy(lens.astype(np.int32)).to(dev)
out = torch.empty(plan.out_shape(n), device=dev)
for _ in range(3):
    plan.forward(wave, offset, length, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    plan.forward(wave, offset, length, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
frac = np.minimum(lens, 80000).sum() / (n * 80000.0)
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'ship'}: 6900 ragged clips (mean {secs.mean():.2f} s, {100*frac:.0f} % of the padded samples are signal): "
      f"{ms:.3f} ms -> {n / ms * 1e3:,.0f} clips/s; checksum {float(out.double().abs().mean()):.9f}")
