import os, sys
sys.path.insert(0, "/root/repo")
import torch
from audio_classification_icbhi_b200 import _lib
name = sys.argv[1]
_lib.LIB_PATH = f"/root/repo/tools/variants/liblogmel_{name}.bin"
from audio_classification_icbhi_b200.plan import LogMelPlan
plan = LogMelPlan(device="cuda:0")
B, T = 4096, 80000
wave = torch.randn(B * T, device="cuda") * 0.1
off = torch.arange(B, device="cuda", dtype=torch.int64) * T
ln = torch.full((B,), T, device="cuda", dtype=torch.int32)
plan.set("pipeline", 1)
out = torch.empty(plan.out_shape(B), device="cuda")
for _ in range(3): plan.forward(wave, off, ln, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): plan.forward(wave, off, ln, out=out)
e1.record(); torch.cuda.synchronize()
print(name, e0.elapsed_time(e1) / 10, "ms")
