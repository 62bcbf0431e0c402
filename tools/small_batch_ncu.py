"""A handful of small-batch launches for `ncu --metrics gpu__time_duration.sum` (kernel-only durations of BASELINE
configs[0] and [3]): B = 1 (5 s), B = 32 and 64 (3 s, augmented with on-device noise), 5 launches each after warm-up."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from audio_classification_icbhi_b200 import LogMelPlan, draw_fast_augmentation

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
for T, B, aug in ((80000, 1, False), (48000, 32, True), (48000, 64, True), (48000, 32, False)):
    plan = LogMelPlan(target_length=T, device=dev)
    clips = torch.randn(B * T, generator=g, device=dev) * 0.1
    off = torch.arange(B, device=dev, dtype=torch.int64) * T
    ln = torch.full((B,), T, device=dev, dtype=torch.int32)
    out = torch.empty(plan.out_shape(B), device=dev)
    a = plan.upload_aug(draw_fast_augmentation(B, T, 128, plan.frames, rng=np.random.default_rng(1), gain_db=6.0)) if aug else None
    for _ in range(8):
        plan.forward(clips, off, ln, aug=a, out=out)
    torch.cuda.synchronize()
    print("config", T, B, aug, flush=True)
