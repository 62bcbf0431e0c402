#!/usr/bin/env python3
"""GPU box: run the LM_PIPE_DEBUG build of the pipeline kernel on a small batch and print where waits got stuck."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from audio_classification_icbhi_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "variants", "liblogmel_pipedbg.bin")
from audio_classification_icbhi_b200.plan import LogMelPlan
plan = LogMelPlan(device="cuda:0")
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 80000
wave = torch.randn(B * T, device="cuda") * 0.1
off = torch.arange(B, device="cuda", dtype=torch.int64) * T
ln = torch.full((B,), T, device="cuda", dtype=torch.int32)
plan.set("pipeline", 1)
out = plan.forward(wave, off, ln)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_int * 4096)()
lib.lm_debug_pipe(buf, 4096)
a = np.array(buf)
n = int(a[0]); print("stuck waits:", n)
sites = {1: "stager ticket", 2: "stop marker ring space", 3: "FFT warp waits staged tile", 4: "FFT warp waits ring slot", 5: "loader loop", 6: "loader waits smem buffer", 7: "loader final", 8: "mel warp waits rows"}
from collections import Counter
c = Counter()
for i in range(min(n, 600)):
    e = a[8 + 6 * i: 14 + 6 * i]
    c[(int(e[2]),)] += 1
    if i < 40: print(f"block {e[0]:3d} role {e[0] % 4} warp {e[1]:2d} site {e[2]} ({sites.get(int(e[2]))}) a={e[3]} b={e[4]} c={e[5]}")
print(c)
plan.set("pipeline", 0)
ref = plan.forward(wave, off, ln); torch.cuda.synchronize()
print("bit-identical:", torch.equal(out, ref))

# per-role time breakdown (cycles summed over the warps of a CTA)
t = (ctypes.c_ulonglong * 1280)()
lib.lm_debug_pipe_time(t, 1280, 1)
plan.set("pipeline", 1)
import time
torch.cuda.synchronize(); t0 = time.time()
out = plan.forward(wave, off, ln); torch.cuda.synchronize()
print("one launch wall ms:", (time.time() - t0) * 1e3)
lib.lm_debug_pipe_time(t, 1280, 0)
tt = np.array(t, dtype=np.float64).reshape(160, 8)[:148]
fft = tt[np.arange(148) % 4 != 0]; mel = tt[np.arange(148) % 4 == 0]
items_team = B * 20 / (111 * 2)
print("transform CTAs, cycles per item per warp: wait staged %.0f | load + part 1 %.0f | restage %.0f | part 2 + ring wait %.0f | bulk store issue %.0f | untangle %.0f | syncwarp %.0f" % tuple(fft[:, :7].mean(axis=0) / 16 / items_team))
items_mel = B * 20 / 37
print("mel CTAs, per item: loader polling %.0f | loader waits buffer %.0f | loader issue %.0f ;  mel warps (per warp): wait rows %.0f | mel %.0f | tail + norm %.0f" % (mel[:, 5].mean() / items_mel, mel[:, 6].mean() / items_mel, mel[:, 7].mean() / items_mel, mel[:, 0].mean() / 15 / items_mel, mel[:, 1].mean() / 15 / items_mel, mel[:, 2].mean() / 15 / items_mel))
