#!/usr/bin/env python3
"""GPU box: copy-only pipeline (int16 in, fp32 features out) in chunks over 3 streams, like lm_forward_host_pcm16."""
import time, torch
dev = torch.device("cuda:0")
B, T, F = 4096, 80000, 128 * 157
h_in = torch.empty(B, T, dtype=torch.int16).pin_memory(); d_in = torch.empty(B, T, dtype=torch.int16, device=dev)
h_out = torch.empty(B, F, dtype=torch.float32).pin_memory(); d_out = torch.empty(B, F, dtype=torch.float32, device=dev)
d_f = torch.empty(B, T, dtype=torch.float32, device=dev)
streams = [torch.cuda.Stream() for _ in range(3)]
def run(chunk, compute, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        k = 0
        for c0 in range(0, B, chunk):
            s = streams[k % 3]; k += 1
            s.synchronize()
            with torch.cuda.stream(s):
                d_in[c0:c0 + chunk].copy_(h_in[c0:c0 + chunk], non_blocking=True)
                if compute:
                    d_f[c0:c0 + chunk].copy_(d_in[c0:c0 + chunk])       # int16 -> fp32 on device (1.3 GB written)
                h_out[c0:c0 + chunk].copy_(d_out[c0:c0 + chunk], non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
for chunk in (148, 296, 1024, 4096):
    for compute in (False, True):
        t = run(chunk, compute)
        print(f"chunk {chunk:5d} compute={compute}: {t*1e3:7.2f} ms -> {B/t:,.0f} clips/s")
