#!/usr/bin/env python3
"""The reference's mel filterbanks, bit for bit: `T.MelSpectrogram(...).mel_scale.fb` as built by torchaudio
for the two configurations of the path (R/src/data/preprocessing.py:38-47: n_fft 2048; R/data/
preprocessing_flexible.py:33-36: n_fft 1024), stored sparsely (2024 / 1960 non-zeros).  The numpy restatement in
oracle/logmel_oracle.py reproduces the float32 arithmetic of `melscale_fbanks` only to ~1e-5 (powf / linspace
rounding), which is fine for the 1e-4 bar but hides how close the CUDA path really is; with these weights the
oracle isolates the arithmetic of the path itself.

    python tests/golden/make_fb_golden.py   ->  tests/golden/fb_golden.npz
"""
import os

import numpy as np
import torchaudio.transforms as T

HERE = os.path.dirname(os.path.abspath(__file__))


def main() -> None:
    out = {}
    for n_fft, hop in ((2048, 512), (1024, 256)):
        ms = T.MelSpectrogram(sample_rate=16000, n_fft=n_fft, hop_length=hop, n_mels=128, f_min=0, f_max=8000)
        fb = ms.mel_scale.fb.numpy()                      # [n_freqs, n_mels] float32
        k, m = np.nonzero(fb)
        out[f"{n_fft}/shape"] = np.array(fb.shape, dtype=np.int64)
        out[f"{n_fft}/k"] = k.astype(np.int32)
        out[f"{n_fft}/m"] = m.astype(np.int32)
        out[f"{n_fft}/v"] = fb[k, m].astype(np.float32)
        print(n_fft, fb.shape, len(k), "non-zeros")
    path = os.path.join(HERE, "fb_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
