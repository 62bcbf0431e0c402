"""GPU polyphase sinc resampler with torchaudio.transforms.Resample's defaults: the resampling step of
AudioPreprocessor.load_audio (R/src/data/preprocessing.py:63-65), on the device.

ICBHI ships 4 kHz, 10 kHz and 44.1 kHz recordings; the reference resamples each to 16 kHz on the host with
a dense conv1d (TA/functional/functional.py `_apply_sinc_resample_kernel`).  Here the per-phase taps are
built once per (orig, new) pair inside the library (`lm_resampler_create`) and one kernel produces the
resampled waveform on the device, ready for `LogMelPlan.forward`.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple, Union

import torch

from . import _lib


class Resampler:
    """`Resampler(orig_freq, new_freq)(waveform)`; waveform `[len]` or `[channels, len]` fp32 (host or device).
    Returns a device tensor of the same rank with `ceil(new * len / orig)` samples per channel."""

    def __init__(self, orig_freq: int, new_freq: int, device: Union[str, torch.device] = "cuda:0"):
        if int(orig_freq) != orig_freq or int(new_freq) != new_freq or orig_freq < 1 or new_freq < 1:
            raise ValueError("frequencies must be positive integers")
        self.orig_freq, self.new_freq = int(orig_freq), int(new_freq)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("Resampler needs a CUDA device (there is no CPU fallback)")
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.lm_resampler_create(self.orig_freq, self.new_freq, self.device.index or 0, C.byref(h)))
        self._h = h

    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.lm_resampler_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def out_len(self, in_len: int) -> int:
        return int(self._lib.lm_resampler_out_len(self._h, int(in_len)))

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        if waveform.dtype != torch.float32:
            raise TypeError(f"Expected floating point type for waveform tensor, but received {waveform.dtype}.")
        x = waveform.to(self.device).contiguous()
        n_in = int(x.shape[-1])
        rows = 1
        for d in x.shape[:-1]:
            rows *= int(d)
        flat = x.view(rows, n_in)
        n_out = self.out_len(n_in)
        y = torch.empty((flat.shape[0], n_out), dtype=torch.float32, device=self.device)
        s = torch.cuda.current_stream(self.device)
        _lib.check(self._lib.lm_resample_rows(self._h, flat.data_ptr(), n_in, n_in, rows, y.data_ptr(), n_out,
                                              C.c_void_p(s.cuda_stream)))
        return y.view(x.shape[:-1] + (n_out,))


_cache: Dict[Tuple[int, int, str], Resampler] = {}


def get_resampler(orig_freq: int, new_freq: int, device: Union[str, torch.device] = "cuda:0") -> Resampler:
    """One resampler per (orig, new, device): the reference builds a fresh T.Resample per file
    (R/src/data/preprocessing.py:64); the taps only depend on the rate pair."""
    key = (int(orig_freq), int(new_freq), str(torch.device(device)))
    r = _cache.get(key)
    if r is None:
        r = _cache[key] = Resampler(orig_freq, new_freq, device)
    return r
