#!/usr/bin/env python3
"""Builds kernel variants (compile-time experiment switches) into build/variants/*.so and, on a GPU
box, times each on the headline batch.  Usage:
    python tools/build_variants.py build            # here (no GPU)
    python tools/build_variants.py bench            # on the GPU box (gpurun)
"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "audio_classification_icbhi_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "variants")
VARIANTS = {
    "fftonly": ["-DLM_EXP=1"],
    "melonly": ["-DLM_EXP=2"],
}
if len(sys.argv) > 2:   # python tools/build_variants.py build name=-DFLAG,-DFLAG ...
    VARIANTS = {a.split("=", 1)[0]: a.split("=", 1)[1].split(",") for a in sys.argv[2:]}

def build():
    os.makedirs(OUT, exist_ok=True)
    for name, flags in VARIANTS.items():
        so = os.path.join(OUT, f"liblogmel_{name}.bin")   # not *.so: keep it out of the driver's .so census
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xptxas", "-v",
               "-shared", "-Xcompiler", "-fPIC", *flags, "-o", so, os.path.join(CSRC, "logmel_capi.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            print(r.stderr); raise SystemExit(1)
        lines = (r.stdout + r.stderr).splitlines()
        info = [l.strip() for i, l in enumerate(lines) if "spill" in l and "2048ELb0" in lines[i - 1] + lines[i - 2]]
        print(name, flags, info)

def bench():
    sys.path.insert(0, ROOT)
    import torch
    from audio_classification_icbhi_b200 import _lib
    res = {}
    for name in VARIANTS:
        _lib._lib = None
        _lib.LIB_PATH = os.path.join(OUT, f"liblogmel_{name}.bin")
        from audio_classification_icbhi_b200.plan import LogMelPlan
        plan = LogMelPlan(device="cuda:0")
        B, T = 4096, 80000
        g = torch.Generator(device="cuda").manual_seed(1234)
        clips = torch.randn(B, T, generator=g, device="cuda") * 0.1
        off = torch.arange(B, device="cuda", dtype=torch.int64) * T
        ln = torch.full((B,), T, device="cuda", dtype=torch.int32)
        out = torch.empty(plan.out_shape(B), device="cuda")
        for _ in range(3):
            plan.forward(clips.view(-1), off, ln, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            plan.forward(clips.view(-1), off, ln, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res[name] = {"ms": ms, "clips_per_s": B / ms * 1e3, "checksum": float(out.double().abs().mean())}
        print(name, res[name], flush=True)
        del plan
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)

if __name__ == "__main__":
    {"build": build, "bench": bench}[sys.argv[1]]()
