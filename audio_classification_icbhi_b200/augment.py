"""Host-side replay of the reference's seeded augmentation choices.

The reference draws, per clip and in this order (R/src/data/preprocessing.py:95-109):

    np.random.random() > 0.5        -> add_noise:   torch.randn_like(waveform) * 0.005
    np.random.random() > 0.5        -> time_shift:  int(np.random.uniform(-0.2, 0.2) * T), torch.roll
    torch.rand(1), torch.rand(1)    -> FrequencyMasking(15)   (torchaudio mask_along_axis, fp32)
    torch.rand(1), torch.rand(1)    -> TimeMasking(35)

from numpy's *global* RandomState and torch's *global* CPU generator.  `draw_reference_augmentation`
consumes exactly those streams in exactly that order, so after `set_seed(s)` the GPU path applies
the same shifts, noise values and mask intervals as the reference would.  Only scalars (and the
noise tensor when it fires) are produced here; applying them is the CUDA kernel's job.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from .plan import make_aug_array

__all__ = ["draw_reference_augmentation", "draw_fast_augmentation", "mask_interval"]


def mask_interval(mask_param: int, axis_len: int) -> Tuple[int, int]:
    """torchaudio.functional.mask_along_axis's interval (functional.py:939-944), fp32 on CPU,
    drawn from torch's global generator."""
    if mask_param < 1:
        return 0, 0
    value = torch.rand(1) * mask_param
    min_value = torch.rand(1) * (axis_len - value)
    start = int(min_value.long())
    end = start + int(value.long())
    return start, end


def draw_reference_augmentation(n_clips: int, target_length: int, n_mels: int, frames: int, *,
                                noise_factor: float = 0.005, shift_max: float = 0.2,
                                freq_mask_param: int = 15, time_mask_param: int = 35,
                                waveform: bool = True, spectrogram: bool = True
                                ) -> Tuple[np.ndarray, Optional[torch.Tensor]]:
    """Returns ([n] lm_aug array, noise [n, T] float32 CPU tensor or None if no clip drew noise).

    `frames` is the length of the time axis the masks act on (after FlexibleAudioPreprocessor's
    resize when that applies)."""
    aug = make_aug_array(n_clips)
    noise = None
    for i in range(n_clips):
        if waveform:
            if np.random.random() > 0.5:
                if noise is None:
                    noise = torch.zeros(n_clips, target_length, dtype=torch.float32)
                noise[i] = torch.randn(1, target_length)[0]
                aug[i]["noise_scale"] = noise_factor
            if np.random.random() > 0.5:
                aug[i]["shift"] = int(np.random.uniform(-shift_max, shift_max) * target_length)
        if spectrogram:
            aug[i]["f0"], aug[i]["f1"] = mask_interval(freq_mask_param, n_mels)
            aug[i]["t0"], aug[i]["t1"] = mask_interval(time_mask_param, frames)
    return aug, noise


def draw_fast_augmentation(n_clips: int, target_length: int, n_mels: int, frames: int, *,
                           rng: Optional[np.random.Generator] = None, noise_factor: float = 0.005,
                           shift_max: float = 0.2, freq_mask_param: int = 15, time_mask_param: int = 35,
                           gain_db: float = 0.0) -> np.ndarray:
    """Throughput mode: same distributions, vectorised draws from a private numpy Generator, and the
    noise itself generated on the GPU (Philox keyed by `seed`), so nothing but 40 bytes per clip
    crosses PCIe.  Not stream-compatible with the reference (by design).  `gain_db` > 0 adds the
    uniform +-gain_db gain augmentation BASELINE.json mentions (absent from the reference)."""
    rng = rng or np.random.default_rng()
    aug = make_aug_array(n_clips)
    aug["noise_scale"] = np.where(rng.random(n_clips) > 0.5, noise_factor, 0.0)
    do_shift = rng.random(n_clips) > 0.5
    aug["shift"] = np.where(do_shift, (rng.uniform(-shift_max, shift_max, n_clips) * target_length).astype(np.int64), 0)
    fv = (rng.random(n_clips).astype(np.float32) * np.float32(freq_mask_param))
    f0 = (rng.random(n_clips).astype(np.float32) * (np.float32(n_mels) - fv)).astype(np.int64)
    tv = (rng.random(n_clips).astype(np.float32) * np.float32(time_mask_param))
    t0 = (rng.random(n_clips).astype(np.float32) * (np.float32(frames) - tv)).astype(np.int64)
    aug["f0"], aug["f1"] = f0, f0 + fv.astype(np.int64)
    aug["t0"], aug["t1"] = t0, t0 + tv.astype(np.int64)
    aug["seed"] = rng.integers(1, 2 ** 63 - 1, n_clips, dtype=np.int64).astype(np.uint64)
    if gain_db > 0.0:
        aug["gain"] = 10.0 ** (rng.uniform(-gain_db, gain_db, n_clips) / 20.0)
    return aug
