#!/usr/bin/env python3
"""GPU box: worst-case distance of the CUDA path to the float64 oracle on the BASELINE tolerances
(log-mel dB abs, mel power rel against max(|ref|, 1e-6 clip peak), normalised features abs)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_classification_icbhi_b200 import _lib
if len(sys.argv) > 1:      # tools/variants/liblogmel_<name>.bin instead of the shipped library
    _lib.LIB_PATH = os.path.join(ROOT, "tools", "variants", f"liblogmel_{sys.argv[1]}.bin")
    import ctypes
    _probe = ctypes.CDLL(_lib.LIB_PATH)      # older variants lack the newer entry points
    for _name in list(_lib.EXPORTS):
        if not hasattr(_probe, _name):
            del _lib.EXPORTS[_name]
from audio_classification_icbhi_b200 import LogMelPlan
from oracle import logmel_oracle as O

plan = LogMelPlan(device="cuda:0")
cfg = O.OracleConfig()
rs = np.random.RandomState(0)
t = np.arange(80000) / 16000.0
clips = {
    "gaussian 0.1": rs.standard_normal(80000) * 0.1,
    "uniform [-1,1]": rs.uniform(-1, 1, 80000),
    "gaussian 1e-3": rs.standard_normal(80000) * 1e-3,
    "padded 1.3 s": rs.standard_normal(20800) * 0.1,
    "cropped 9 s": rs.standard_normal(144000) * 0.1,
    "440 Hz + 1e-3 noise": np.sin(2 * np.pi * 440 * t) * 0.5 + rs.standard_normal(80000) * 1e-3,
    "chirp + noise": np.sin(2 * np.pi * (100 * t + 300 * t * t)) * 0.3 + rs.standard_normal(80000) * 1e-2,
}
dev = plan.device
print(f"{'clip':24s} {'dB abs':>10s} {'mel rel':>10s} {'norm abs':>10s}")
for name, x in clips.items():
    x = x.astype(np.float32)
    w = torch.from_numpy(x).to(dev)
    off = torch.zeros(1, dtype=torch.int64, device=dev)
    ln = torch.tensor([len(x)], dtype=torch.int32, device=dev)
    db = torch.empty(plan.out_shape(1), device=dev); mp = torch.empty(plan.out_shape(1), device=dev)
    out = plan.forward(w, off, ln, out_db=db, out_melpow=mp)
    torch.cuda.synchronize()
    st = O.logmel(x, cfg, return_stages=True, fb=O.golden_filterbank(2048))
    ref_mp, ref_db, ref_norm = st["mel_power"], st["db"], st["out"]
    floor = max(1e-6 * np.abs(ref_mp).max(), 1e-30)
    rel = np.abs(mp[0, 0].cpu().numpy() - ref_mp) / np.maximum(np.abs(ref_mp), floor)
    print(f"{name:24s} {np.abs(db[0,0].cpu().numpy() - ref_db).max():10.2e} {rel.max():10.2e} "
          f"{np.abs(out[0,0].cpu().numpy() - ref_norm).max():10.2e}")
