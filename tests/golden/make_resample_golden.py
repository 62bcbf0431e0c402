#!/usr/bin/env python3
"""Golden vectors for the resampling step of AudioPreprocessor.load_audio
(R/src/data/preprocessing.py:63-65): `T.Resample(sr, self.sample_rate)(waveform)` executed here with the
installed torchaudio, exactly as the reference calls it (defaults: sinc_interp_hann, width 6, rolloff 0.99).
ICBHI recordings come at 4 kHz, 10 kHz and 44.1 kHz; 8 / 22.05 / 48 kHz are added as further rate pairs.

    python tests/golden/make_resample_golden.py      ->  tests/golden/resample_golden.npz
"""
import os

import numpy as np
import torch
import torchaudio.transforms as T

HERE = os.path.dirname(os.path.abspath(__file__))
RATES = (4000, 10000, 44100, 8000, 22050, 48000)
TARGET = 16000


def golden_resample_input(sr: int) -> np.ndarray:
    rs = np.random.RandomState(1000 + sr % 997)
    n = int(0.21 * sr) + 3           # odd lengths: the last partial output block is exercised
    t = np.arange(n) / sr
    x = 0.1 * rs.standard_normal(n) + 0.2 * np.sin(2 * np.pi * 220.0 * t)
    return x.astype(np.float32)


def main() -> None:
    out = {}
    for sr in RATES:
        x = golden_resample_input(sr)
        y = T.Resample(sr, TARGET)(torch.from_numpy(x).unsqueeze(0))[0].numpy()
        out[f"in/{sr}"] = x
        out[f"out/{sr}"] = y.astype(np.float32)
        print(sr, x.shape, "->", y.shape)
    path = os.path.join(HERE, "resample_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
