"""The reference's per-clip CPU pipeline, restated on top of the SAME third-party kernels it
calls (torchaudio.transforms on the host CPU) -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference (R/src/data/preprocessing.py:20-53, :118-151) is 40 lines of glue around
``T.MelSpectrogram``, ``T.AmplitudeToDB`` and a mean/std normalisation; all arithmetic is in
torchaudio / torch, which are installed libraries, not files under /root/reference.  This module
is that glue written again so it can travel to the GPU box (where /root/reference does not
exist).  It is used by

* ``bench.py``: the ``cpu_baseline`` leg and ``--impl reference`` (kind = "port"), and
* ``tests/``: as a second checker beside the numpy oracle when torchaudio imports.

It is never imported by the product package.
"""

from __future__ import annotations

import os
import time
from typing import Optional, Tuple

import torch

try:  # torchaudio is part of the image; keep the import failure readable if it is not
    import torchaudio.transforms as T
except Exception as exc:  # pragma: no cover
    T = None
    _IMPORT_ERROR = exc


class ReferencePipeline:
    """pad/crop -> MelSpectrogram(power=2) -> AmplitudeToDB -> (masks) -> normalise, on CPU.

    Mirrors the object graph built by ``AudioPreprocessor.__init__``
    (R/src/data/preprocessing.py:37-53); ``flexible=True`` applies the n_fft/hop override of
    R/data/preprocessing_flexible.py:33-36 and the bilinear resize of :118-154."""

    def __init__(self, sample_rate: int = 16000, n_mels: int = 128, n_fft: int = 2048,
                 hop_length: int = 512, duration: float = 5.0, flexible: bool = False):
        if T is None:  # pragma: no cover
            raise RuntimeError(f"torchaudio is not importable: {_IMPORT_ERROR}")
        if flexible and duration < 1.0:
            n_fft = min(1024, int(sample_rate * duration / 2))
            hop_length = n_fft // 4
        self.sample_rate, self.n_mels, self.n_fft, self.hop_length = sample_rate, n_mels, n_fft, hop_length
        self.duration, self.flexible = duration, flexible
        self.target_length = int(sample_rate * duration)
        self.mel = T.MelSpectrogram(sample_rate=sample_rate, n_fft=n_fft, hop_length=hop_length,
                                    n_mels=n_mels, power=2.0)
        self.to_db = T.AmplitudeToDB()

    # -- stages ------------------------------------------------------------------------
    def pad_or_crop(self, w: torch.Tensor) -> torch.Tensor:
        n = w.shape[-1]
        if n < self.target_length:
            return torch.nn.functional.pad(w, (0, self.target_length - n))
        if n > self.target_length:
            s = (n - self.target_length) // 2
            return w[..., s:s + self.target_length]
        return w

    def mel_power(self, w: torch.Tensor) -> torch.Tensor:
        return self.mel(self.pad_or_crop(w))

    def db(self, w: torch.Tensor) -> torch.Tensor:
        return self.to_db(self.mel_power(w))

    def resize(self, db: torch.Tensor) -> torch.Tensor:
        import math
        target = max(int(math.ceil(self.target_length / self.hop_length)), 32)
        if db.shape[-1] == target:
            return db
        x = db.reshape(-1, 1, db.shape[-2], db.shape[-1])
        x = torch.nn.functional.interpolate(x, size=(self.n_mels, target), mode="bilinear",
                                            align_corners=False)
        return x.reshape(db.shape[:-1] + (target,))

    @staticmethod
    def normalize(x: torch.Tensor) -> torch.Tensor:
        return (x - x.mean()) / (x.std() + 1e-8)

    def __call__(self, w: torch.Tensor, masks: Optional[Tuple[int, int, int, int]] = None) -> torch.Tensor:
        """One clip ``[1, len]`` -> ``[1, n_mels, frames]`` (the reference's call pattern)."""
        x = self.db(w)
        if self.flexible:
            x = self.resize(x)
        if masks is not None:
            f0, f1, t0, t1 = masks
            x = x.clone()
            x[..., f0:f1, :] = 0.0
            x[..., :, t0:t1] = 0.0
        return self.normalize(x)

    def batched(self, w: torch.Tensor) -> torch.Tensor:
        """``[B, len]`` -> ``[B, 1, n_mels, frames]`` with per-clip statistics: one batched
        torchaudio call (bit-identical to the per-clip loop)."""
        x = self.db(w)
        if self.flexible:
            x = self.resize(x)
        mean = x.mean(dim=(-2, -1), keepdim=True)
        std = x.std(dim=(-2, -1), keepdim=True)
        return ((x - mean) / (std + 1e-8)).unsqueeze(1)


def time_reference(clips: torch.Tensor, cfg: dict, budget_s: float = 20.0, threads: Optional[int] = None) -> dict:
    """Times the reference call pattern (per-clip Python loop) and the batched form on
    ``clips`` ``[n, len]`` with all host threads; returns the faster as clips/s.

    Bounded: stops adding repetitions once ``budget_s`` of CPU time has been spent."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    pipe = ReferencePipeline(**cfg)
    n = clips.shape[0]
    with torch.no_grad():
        for i in range(min(3, n)):  # warm-up
            pipe(clips[i:i + 1])
        t_loop, reps_loop = 0.0, 0
        while t_loop < budget_s / 2 and reps_loop < 5:
            t0 = time.perf_counter()
            for i in range(n):
                pipe(clips[i:i + 1])
            t_loop += time.perf_counter() - t0
            reps_loop += 1
        pipe.batched(clips[:min(8, n)])
        t_b, reps_b = 0.0, 0
        while t_b < budget_s / 2 and reps_b < 5:
            t0 = time.perf_counter()
            pipe.batched(clips)
            t_b += time.perf_counter() - t0
            reps_b += 1
    loop_cps = n * reps_loop / t_loop
    batched_cps = n * reps_b / t_b
    return {"per_clip_loop": loop_cps, "batched": batched_cps, "best": max(loop_cps, batched_cps),
            "threads": threads, "clips": n, "reps": (reps_loop, reps_b)}
