#!/usr/bin/env python3
"""On the GPU box: time the shipped library (and any tools/variants/*.bin named on the command line)
on the headline batch, sweeping run-time knobs.  Usage: python tools/sweep.py [variant ...] [--stagger a,b,c]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from audio_classification_icbhi_b200 import _lib

def time_plan(plan, clips, off, ln, out, iters=20):
    for _ in range(3):
        plan.forward(clips.view(-1), off, ln, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        plan.forward(clips.view(-1), off, ln, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def main():
    args = sys.argv[1:]
    staggers = [2000]
    names = []
    i = 0
    while i < len(args):
        if args[i] == "--stagger":
            staggers = [int(x) for x in args[i + 1].split(",")]; i += 2
        else:
            names.append(args[i]); i += 1
    B, T = 4096, 80000
    g = torch.Generator(device="cuda").manual_seed(1234)
    clips = torch.randn(B, T, generator=g, device="cuda") * 0.1
    off = torch.arange(B, device="cuda", dtype=torch.int64) * T
    ln = torch.full((B,), T, device="cuda", dtype=torch.int32)
    res = {}
    default_path = _lib.LIB_PATH
    for name in ["ship"] + names:
        _lib._lib = None
        _lib.LIB_PATH = default_path if name == "ship" else os.path.join(ROOT, "tools", "variants", f"liblogmel_{name}.bin")
        from audio_classification_icbhi_b200.plan import LogMelPlan
        plan = LogMelPlan(device="cuda:0")
        out = torch.empty(plan.out_shape(B), device="cuda")
        for st in (staggers if name != "v4" else [0]):
            try:
                plan.set("stagger_ns", st)
            except Exception:
                pass
            ms = time_plan(plan, clips, off, ln, out)
            key = f"{name}@st{st}"
            res[key] = {"ms": ms, "clips_per_s": B / ms * 1e3, "checksum": float(out.double().abs().mean())}
            print(key, res[key], flush=True)
        del plan
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
