"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE CLASSES from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Only ``load_audio`` is bypassed (torchaudio.load needs torchcodec, absent here; SURVEY.md
section 8c): every other step calls the reference's own methods verbatim, in the order of
``AudioPreprocessor.preprocess`` (R/src/data/preprocessing.py:118-151) and
``FlexibleAudioPreprocessor.preprocess`` (R/data/preprocessing_flexible.py:156-192).

Inputs are NOT stored: they are regenerated from numpy's legacy RandomState (stable by contract)
by ``golden_input`` below, which the tests import.  Outputs are stored as float32.
"""

from __future__ import annotations

import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def golden_input(seed: int, n: int, kind: str = "noise") -> np.ndarray:
    """Deterministic synthetic clip, float32, |x| well inside [-1, 1]."""
    rs = np.random.RandomState(seed)
    if kind == "noise":
        return (rs.standard_normal(n) * 0.1).astype(np.float32)
    if kind == "uniform":
        return rs.uniform(-1.0, 1.0, n).astype(np.float32)
    if kind == "tone_noise":  # 440 Hz + weak noise, 16 kHz
        t = np.arange(n, dtype=np.float64) / 16000.0
        return (0.5 * np.sin(2 * np.pi * 440.0 * t) + 1e-3 * rs.standard_normal(n)).astype(np.float32)
    if kind == "zeros":
        return np.zeros(n, dtype=np.float32)
    if kind == "chirp_burst":  # silence + a short burst: exercises the -100 dB floor
        x = np.zeros(n, dtype=np.float32)
        m = min(n, 20800)
        x[:m] = (rs.standard_normal(m) * 0.05).astype(np.float32)
        return x
    raise ValueError(kind)


# (name, class, ctor kwargs, input seed, input length, input kind)
PLAIN_CASES = [
    ("headline_5s", "std", dict(duration=5.0), 0, 80000, "noise"),
    ("uniform_5s", "std", dict(duration=5.0), 1, 80000, "uniform"),
    ("short_1p3s_pad_5s", "std", dict(duration=5.0), 2, 20800, "noise"),
    ("long_7s_crop_5s", "std", dict(duration=5.0), 3, 112001, "noise"),
    ("tone_5s", "std", dict(duration=5.0), 4, 80000, "tone_noise"),
    ("zeros_5s", "std", dict(duration=5.0), 5, 80000, "zeros"),
    ("seg_3s", "std", dict(duration=3.0), 6, 48000, "noise"),
    ("cfg_8s", "std", dict(duration=8.0), 7, 128000, "noise"),
    ("flex_8s_resize", "flex", dict(duration=8.0), 7, 128000, "noise"),
    ("flex_1s", "flex", dict(duration=1.0), 8, 16000, "noise"),
    ("flex_0p5s", "flex", dict(duration=0.5), 9, 8000, "noise"),
    ("flex_1s_short_pad", "flex", dict(duration=1.0), 10, 9000, "noise"),
]


def main() -> None:
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    import torch
    from src.data.preprocessing import AudioPreprocessor
    from data.preprocessing_flexible import FlexibleAudioPreprocessor

    torch.set_num_threads(1)
    out = {}

    # ---- plain (augment=False) cases: every intermediate stage the reference exposes -------
    for name, cls, kw, seed, n, kind in PLAIN_CASES:
        P = (AudioPreprocessor if cls == "std" else FlexibleAudioPreprocessor)(augment=False, **kw)
        w = torch.from_numpy(golden_input(seed, n, kind)).unsqueeze(0)
        w = P.pad_or_crop(w)
        melp = P.mel_spectrogram(w)
        db = P.amplitude_to_db(melp)
        if cls == "flex":
            db = P.resize_spectrogram(db)
        norm = P.normalize(db)
        out[f"{name}/mel_power"] = melp[0].numpy().astype(np.float32)
        out[f"{name}/db"] = db[0].numpy().astype(np.float32)
        out[f"{name}/norm"] = norm[0].numpy().astype(np.float32)
        out[f"{name}/meta"] = np.array([P.n_fft, P.hop_length, P.target_length, melp.shape[-1], db.shape[-1]],
                                       dtype=np.int64)
        if name == "headline_5s":
            out["const/window"] = P.mel_spectrogram.spectrogram.window.numpy().astype(np.float32)
            out["const/fb"] = P.mel_spectrogram.mel_scale.fb.numpy().astype(np.float32)
        if name == "flex_0p5s":
            out["const/fb_513"] = P.mel_spectrogram.mel_scale.fb.numpy().astype(np.float32)
        print(name, tuple(norm.shape), "n_fft", P.n_fft, "hop", P.hop_length)

    # ---- seeded augmentation trace: config 4 (3 s) and the headline 5 s, seeds = 42 --------
    # aug64_3s is BASELINE configs[3] at its stated size (config_segmented.yaml:21 batch 32, README 64): 64 clips; to keep
    # the fixture small only the trace and 64 sampled output values per clip are stored (positions in aug64_3s/pos)
    sample_pos = np.random.RandomState(99).randint(0, 128 * 94, size=64)
    out["aug64_3s/pos"] = sample_pos.astype(np.int64)
    for tag, dur, n_clips in (("aug_3s", 3.0, 6), ("aug_5s", 5.0, 6), ("aug64_3s", 3.0, 64)):
        random.seed(42)
        np.random.seed(42)
        torch.manual_seed(42)  # R/src/utils/config.py:31-33 (set_seed)
        P = AudioPreprocessor(duration=dur, augment=True)
        T_len = P.target_length
        trace = []
        real_roll = torch.roll
        for c in range(n_clips):
            w0 = P.pad_or_crop(torch.from_numpy(golden_input(100 + c, T_len, "noise")).unsqueeze(0))
            log = {"shift": 0, "noise": None}

            def logging_roll(x, shifts, dims=None, _log=log):
                _log["shift"] = int(shifts)
                return real_roll(x, shifts, dims)

            real_add_noise = P.add_noise

            def logging_add_noise(x, noise_factor=0.005, _log=log):
                y = real_add_noise(x, noise_factor)
                _log["noise"] = ((y - x) / noise_factor).numpy().astype(np.float32)[0]
                return y

            torch.roll = logging_roll
            P.add_noise = logging_add_noise
            try:
                w1 = P.augment_waveform(w0)
            finally:
                torch.roll = real_roll
                P.add_noise = real_add_noise
            db = P.amplitude_to_db(P.mel_spectrogram(w1))
            masked = P.augment_spectrogram(db)
            norm = P.normalize(masked)
            # recover the mask intervals from the data (dB is never exactly 0.0 unmasked)
            z = (masked[0] == 0.0)
            rows = torch.nonzero(z.all(dim=1)).flatten().tolist()
            cols = torch.nonzero(z.all(dim=0)).flatten().tolist()
            f0, f1 = (rows[0], rows[-1] + 1) if rows else (0, 0)
            t0, t1 = (cols[0], cols[-1] + 1) if cols else (0, 0)
            assert rows == list(range(f0, f1)) and cols == list(range(t0, t1))
            noise = log["noise"]
            trace.append([int(noise is not None), log["shift"], f0, f1, t0, t1])
            if tag == "aug64_3s":
                out.setdefault(f"{tag}/samples", np.zeros((n_clips, 64), dtype=np.float32))[c] = norm[0].numpy().reshape(-1)[sample_pos]
                out.setdefault(f"{tag}/absmean", np.zeros(n_clips, dtype=np.float64))[c] = float(norm[0].abs().double().mean())
                continue
            out[f"{tag}/clip{c}/norm"] = norm[0].numpy().astype(np.float32)
            if noise is not None:
                out[f"{tag}/clip{c}/noise_head"] = noise[:32].copy()
                out[f"{tag}/clip{c}/noise_sum"] = np.array([noise.astype(np.float64).sum(),
                                                            np.abs(noise.astype(np.float64)).sum()])
            if c < 6:
                print(tag, c, trace[-1])
        out[f"{tag}/trace"] = np.array(trace, dtype=np.int64)

    # ---- RNG known answers ---------------------------------------
    torch.manual_seed(42)
    out["rng/torch_rand_seed42"] = torch.cat([torch.rand(1) for _ in range(8)]).numpy()
    torch.manual_seed(7)
    out["rng/torch_randn_seed7_head"] = torch.randn(1, 48000)[0, :64].numpy()
    out["rng/torch_rand_after_randn"] = torch.rand(1).numpy()
    np.random.seed(42)
    out["rng/numpy_random_seed42"] = np.array([np.random.random() for _ in range(4)])

    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
