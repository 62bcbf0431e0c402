"""Latency / throughput of small batches (BASELINE configs[0] and [3]): one launch of B clips, CUDA-event timed,
with the small-batch split on (default) and off.  Run on the GPU box: python tools/small_batch_probe.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
from audio_classification_icbhi_b200 import LogMelPlan


def time_launch(plan, clips, reps=200):
    B, n = clips.shape
    offset = torch.arange(B, device=clips.device, dtype=torch.int64) * n
    length = torch.full((B,), n, device=clips.device, dtype=torch.int32)
    out = torch.empty(plan.out_shape(B), device=clips.device)
    flat = clips.view(-1)
    for _ in range(20):
        plan.forward(flat, offset, length, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.forward(flat, offset, length, out=out)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[int(len(ts) * 0.9)]


def main():
    dev = torch.device("cuda:0")
    for T in (80000, 48000):
        plan = LogMelPlan(target_length=T, device=dev)
        g = torch.Generator(device=dev).manual_seed(3)
        for B in (1, 2, 8, 32, 64, 148, 296, 1024):
            clips = torch.randn(B, T, generator=g, device=dev) * 0.1
            row = {"T": T, "B": B}
            for name, split in (("split", 0), ("nosplit", 1)):
                plan.set("split", split)
                p50, p90 = time_launch(plan, clips)
                row[name + "_us_p50"] = round(p50, 2)
                row[name + "_us_p90"] = round(p90, 2)
                row[name + "_clips_per_s"] = round(B / p50 * 1e6)
            plan.set("split", 0)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
