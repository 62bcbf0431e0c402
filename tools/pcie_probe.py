#!/usr/bin/env python3
"""GPU box: pinned-memory PCIe bandwidth, one direction at a time and both at once."""
import time, torch
dev = torch.device("cuda:0")
n_in, n_out = 1310720000 // 4, 329252864 // 4
h_in = torch.empty(n_in, dtype=torch.float32).pin_memory(); d_in = torch.empty(n_in, dtype=torch.float32, device=dev)
h_out = torch.empty(n_out, dtype=torch.float32).pin_memory(); d_out = torch.empty(n_out, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5, chunks=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        for c in range(chunks):
            a, b = c * n_in // chunks, (c + 1) * n_in // chunks
            a2, b2 = c * n_out // chunks, (c + 1) * n_out // chunks
            if h2d:
                with torch.cuda.stream(s1): d_in[a:b].copy_(h_in[a:b], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[a2:b2].copy_(d_out[a2:b2], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return dt
for chunks in (1, 20):
    t = run(True, False, chunks=chunks); print(f"chunks {chunks:2d}  H2D only : {t*1e3:7.2f} ms  {n_in*4/t/1e9:6.1f} GB/s")
    t = run(False, True, chunks=chunks); print(f"chunks {chunks:2d}  D2H only : {t*1e3:7.2f} ms  {n_out*4/t/1e9:6.1f} GB/s")
    t = run(True, True, chunks=chunks); print(f"chunks {chunks:2d}  both     : {t*1e3:7.2f} ms  H2D {n_in*4/t/1e9:6.1f} + D2H {n_out*4/t/1e9:6.1f} GB/s  -> {4096/t:,.0f} clips/s ceiling")
