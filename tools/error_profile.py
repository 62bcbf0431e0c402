#!/usr/bin/env python3
"""GPU box: signed relative error of the CUDA mel power against the float64 oracle, per mel band."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_classification_icbhi_b200 import LogMelPlan
from oracle import logmel_oracle as O
plan = LogMelPlan(device="cuda:0"); cfg = O.OracleConfig(); dev = plan.device
rs = np.random.RandomState(0)
x = (rs.standard_normal(80000) * 0.1).astype(np.float32)
w = torch.from_numpy(x).to(dev)
mp = torch.empty(plan.out_shape(1), device=dev)
plan.forward(w, torch.zeros(1, dtype=torch.int64, device=dev), torch.tensor([80000], dtype=torch.int32, device=dev), out_melpow=mp)
torch.cuda.synchronize()
ref = O.logmel(x, cfg, return_stages=True, fb=O.golden_filterbank(2048))["mel_power"]
rel = (mp[0, 0].cpu().numpy().astype(np.float64) - ref) / ref
print("all: mean %.3e  std %.3e  min %.3e  max %.3e" % (rel.mean(), rel.std(), rel.min(), rel.max()))
for m in (0, 1, 2, 3, 8, 16, 32, 64, 96, 127):
    print("mel %3d: mean %.3e std %.3e max|.| %.3e" % (m, rel[m].mean(), rel[m].std(), np.abs(rel[m]).max()))
print("frames 0,1,78,155,156 mean:", [float("%.3e" % rel[:, t].mean()) for t in (0, 1, 78, 155, 156)])
