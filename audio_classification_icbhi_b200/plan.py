"""LogMelPlan: Python handle on one ``lm_plan`` (include/logmel_b200.h).

The plan owns the constants ``AudioPreprocessor.__init__`` builds in the reference
(R/src/data/preprocessing.py:37-47): the periodic Hann window and the HTK mel filterbank of
``T.MelSpectrogram`` and the ``T.AmplitudeToDB`` scalars.  They are produced with the same torch
calls torchaudio makes, so the uploaded bits equal the reference's registered buffers.

torch is used for memory and streams only; all arithmetic runs in liblogmel_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib

__all__ = ["LogMelPlan", "reference_window", "reference_filterbank", "make_aug_array"]


def reference_window(n_fft: int) -> torch.Tensor:
    """``torch.hann_window(n_fft)`` -- the ``window`` buffer of T.Spectrogram
    (torchaudio/transforms/_transforms.py:86-87)."""
    return torch.hann_window(n_fft, dtype=torch.float32)


def reference_filterbank(n_freqs: int, n_mels: int, sample_rate: int,
                         f_min: float = 0.0, f_max: Optional[float] = None) -> torch.Tensor:
    """``fb`` buffer of T.MelScale: melscale_fbanks(n_freqs, f_min, f_max, n_mels, sr, None, "htk")
    (torchaudio/transforms/_transforms.py:402-405).  Uses torchaudio's own builder when it
    imports (bit-identical to the reference); otherwise the same torch float32 ops in the same
    order (torchaudio/functional/functional.py:563-576)."""
    f_max = float(sample_rate // 2) if f_max is None else float(f_max)
    try:
        from torchaudio.functional import melscale_fbanks
        return melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, None, "htk").to(torch.float32)
    except Exception:  # torchaudio absent: identical torch op sequence
        all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
        m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
        m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
        m_pts = torch.linspace(m_min, m_max, n_mels + 2)
        f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
        f_diff = f_pts[1:] - f_pts[:-1]
        slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
        down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
        up = slopes[:, 2:] / f_diff[1:]
        return torch.max(torch.zeros(1), torch.min(down, up)).to(torch.float32)


def make_aug_array(n: int) -> np.ndarray:
    """[n] structured array laid out as ``lm_aug`` with identity defaults."""
    a = np.zeros(n, dtype=_lib.AUG_DTYPE)
    a["gain"] = 1.0
    return a


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class LogMelPlan:
    """One preprocessor configuration on one GPU."""

    def __init__(self, sample_rate: int = 16000, n_fft: int = 2048, hop_length: int = 512, n_mels: int = 128,
                 target_length: int = 80000, device: Union[int, str, torch.device, None] = None,
                 db_multiplier: float = 10.0, amin: float = 1e-10, db_offset: float = 0.0,
                 norm_eps: float = 1e-8, fb: Optional[torch.Tensor] = None,
                 window: Optional[torch.Tensor] = None):
        self._lib = _lib.load()          # raises if the .so is missing
        if not torch.cuda.is_available():
            raise RuntimeError("LogMelPlan needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"LogMelPlan needs a CUDA device, got {dev}")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.sample_rate, self.n_fft, self.hop_length, self.n_mels = sample_rate, n_fft, hop_length, n_mels
        self.target_length = int(target_length)
        self.n_freqs = n_fft // 2 + 1
        win = reference_window(n_fft) if window is None else window.to(torch.float32).cpu()
        fbt = reference_filterbank(self.n_freqs, n_mels, sample_rate) if fb is None else fb.to(torch.float32).cpu()
        if tuple(fbt.shape) != (self.n_freqs, n_mels) or win.numel() != n_fft:
            raise ValueError("window / fb shape mismatch")
        self._win = np.ascontiguousarray(win.numpy())
        self._fb = np.ascontiguousarray(fbt.numpy())
        cfg = _lib.LmConfig(n_fft, hop_length, n_mels, self.target_length,
                            self._win.ctypes.data_as(C.POINTER(C.c_float)),
                            self._fb.ctypes.data_as(C.POINTER(C.c_float)),
                            db_multiplier, amin, db_offset, norm_eps)
        handle = C.c_void_p()
        _lib.check(self._lib.lm_plan_create(C.byref(cfg), dev.index, C.byref(handle)))
        self._h = handle
        self.frames = self._lib.lm_plan_frames(self._h)

    # -- lifetime --------------------------------------------------------------------------
    def close(self) -> None:
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.lm_plan_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection -----------------------------------------------------------------------
    def info(self) -> dict:
        i = _lib.LmInfo()
        _lib.check(self._lib.lm_plan_info(self._h, C.byref(i)))
        return {k: getattr(i, k) for k, _ in _lib.LmInfo._fields_}

    def set(self, key: str, value: int) -> None:
        _lib.check(self._lib.lm_plan_set(self._h, key.encode(), int(value)))

    @property
    def bytes_per_clip(self) -> int:
        return 4 * self.target_length + 4 * self.n_mels * self.frames

    def out_shape(self, batch: int) -> Tuple[int, int, int, int]:
        return (batch, 1, self.n_mels, self.frames)

    # -- device path -----------------------------------------------------------------------
    def forward(self, wave: torch.Tensor, offset: torch.Tensor, length: torch.Tensor,
                aug: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                normalize: bool = True, out: Optional[torch.Tensor] = None,
                out_db: Optional[torch.Tensor] = None, out_melpow: Optional[torch.Tensor] = None,
                stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """Batched log-mel on device tensors.

        wave: fp32 1-D packed samples; offset int64 [B]; length int32 [B];
        aug: uint8 [B*40] device tensor holding ``lm_aug`` records (see ``upload_aug``);
        noise: fp32 [B, target_length].  Returns ``out`` = [B, 1, n_mels, frames] fp32."""
        B = int(offset.numel())
        for name, t, dt in (("wave", wave, torch.float32), ("offset", offset, torch.int64), ("length", length, torch.int32)):
            if t.device != self.device or t.dtype != dt or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous {dt} tensor on {self.device}")
        if length.numel() != B:
            raise ValueError("offset / length size mismatch")
        if out is None:
            out = torch.empty(self.out_shape(B), dtype=torch.float32, device=self.device)
        for name, t in (("out", out), ("out_db", out_db), ("out_melpow", out_melpow)):
            if t is not None and (t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous()
                                  or t.numel() != B * self.n_mels * self.frames):
                raise ValueError(f"{name} must be contiguous fp32 [B,1,n_mels,frames] on {self.device}")
        if aug is not None and (aug.device != self.device or aug.numel() * aug.element_size() != 40 * B):
            raise ValueError("aug must hold B lm_aug records (40 bytes each) on the plan's device")
        if noise is not None and (noise.device != self.device or noise.dtype != torch.float32
                                  or noise.numel() != B * self.target_length or not noise.is_contiguous()):
            raise ValueError("noise must be contiguous fp32 [B, target_length] on the plan's device")
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        _lib.check(self._lib.lm_forward(self._h, wave.data_ptr(), offset.data_ptr(), length.data_ptr(), B,
                                        _ptr(aug), _ptr(noise), out.data_ptr(), _ptr(out_db), _ptr(out_melpow),
                                        1 if normalize else 0, C.c_void_p(s.cuda_stream)))
        return out

    def forward_pcm16(self, pcm: torch.Tensor, offset: torch.Tensor, length: torch.Tensor,
                      aug: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                      normalize: bool = True, out: Optional[torch.Tensor] = None,
                      stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """`forward` on packed int16 samples (the ICBHI wav sample format; R/src/data/preprocessing.py:55-68 is
        torchaudio.load, i.e. x / 32768): one kernel, 2 bytes per sample across HBM, bit-identical to decoding
        first.  Clip starts that are multiples of 8 samples take the bulk-copy staging path."""
        B = int(offset.numel())
        for name, t, dt in (("pcm", pcm, torch.int16), ("offset", offset, torch.int64), ("length", length, torch.int32)):
            if t.device != self.device or t.dtype != dt or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous {dt} tensor on {self.device}")
        if out is None:
            out = torch.empty(self.out_shape(B), dtype=torch.float32, device=self.device)
        elif (out.device != self.device or out.dtype != torch.float32 or not out.is_contiguous()
              or out.numel() != B * self.n_mels * self.frames):
            raise ValueError(f"out must be contiguous fp32 [B,1,n_mels,frames] on {self.device}")
        if aug is not None and (aug.device != self.device or aug.numel() * aug.element_size() != 40 * B):
            raise ValueError("aug must hold B lm_aug records (40 bytes each) on the plan's device")
        if noise is not None and (noise.device != self.device or noise.dtype != torch.float32
                                  or noise.numel() != B * self.target_length or not noise.is_contiguous()):
            raise ValueError("noise must be contiguous fp32 [B, target_length] on the plan's device")
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        _lib.check(self._lib.lm_forward_pcm16(self._h, pcm.data_ptr(), offset.data_ptr(), length.data_ptr(), B,
                                              _ptr(aug), _ptr(noise), out.data_ptr(), 1 if normalize else 0,
                                              C.c_void_p(s.cuda_stream)))
        return out

    def forward_gather(self, wave: torch.Tensor, offset: torch.Tensor, length: torch.Tensor, out_slice_ptr: int,
                       peer_slice_ptrs=(), mc_slice_ptr: int = 0, aug: Optional[torch.Tensor] = None,
                       noise: Optional[torch.Tensor] = None, stream: Optional[torch.cuda.Stream] = None) -> None:
        """`forward` fused with the feature all-gather: the normalised features also go to the same slice of
        the other ranks' gathered buffers (`peer_slice_ptrs`: device addresses mapped into this process, e.g.
        `torch.distributed._symmetric_memory` buffer pointers + slice offset) or through the multicast address
        `mc_slice_ptr`.  See `sharding.FusedGather`."""
        B = int(offset.numel())
        for name, t, dt in (("wave", wave, torch.float32), ("offset", offset, torch.int64), ("length", length, torch.int32)):
            if t.device != self.device or t.dtype != dt or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous {dt} tensor on {self.device}")
        peers = [int(x) for x in peer_slice_ptrs]
        arr = (C.c_void_p * max(len(peers), 1))(*peers) if peers else None
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        _lib.check(self._lib.lm_forward_gather(self._h, wave.data_ptr(), offset.data_ptr(), length.data_ptr(), B,
                                               _ptr(aug), _ptr(noise), C.c_void_p(int(out_slice_ptr)), arr, len(peers),
                                               C.c_void_p(int(mc_slice_ptr)) if mc_slice_ptr else None,
                                               C.c_void_p(s.cuda_stream)))

    @property
    def launches(self) -> int:
        """Kernels launched through this plan so far (counted inside the library)."""
        return int(self._lib.lm_plan_launch_count(self._h))

    def upload_aug(self, aug: np.ndarray) -> torch.Tensor:
        """Host ``lm_aug`` structured array -> device byte tensor."""
        if aug.dtype != np.dtype(_lib.AUG_DTYPE):
            raise ValueError("aug must use the lm_aug dtype (make_aug_array)")
        raw = torch.from_numpy(np.ascontiguousarray(aug).view(np.uint8).copy())
        return raw.to(self.device, non_blocking=False)

    def forward_dense(self, clips: torch.Tensor, **kw) -> torch.Tensor:
        """Convenience: ``clips`` [B, len] (same length each) already on the device."""
        if clips.dim() != 2:
            raise ValueError("clips must be [B, len]")
        B, n = clips.shape
        clips = clips.contiguous()
        offset = torch.arange(B, device=self.device, dtype=torch.int64) * n
        length = torch.full((B,), n, device=self.device, dtype=torch.int32)
        return self.forward(clips.view(-1), offset, length, **kw)

    # -- host path (the call a reference user makes: CPU in, CPU out) ------------------------------
    def forward_host(self, wave: torch.Tensor, offset: torch.Tensor, length: torch.Tensor,
                     aug: Optional[np.ndarray] = None, noise: Optional[torch.Tensor] = None,
                     normalize: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host tensors in, host features out; H2D / kernel / D2H are pipelined inside the
        library on plan-owned streams.  Pinned tensors make the copies asynchronous.
        ``wave`` is fp32 samples or int16 PCM (expanded to x / 32768 inside the kernel: half the PCIe bytes)."""
        B = int(offset.numel())
        if wave.dtype not in (torch.float32, torch.int16):
            raise ValueError("wave must be fp32 samples or int16 PCM")
        for name, t, dt in (("wave", wave, wave.dtype), ("offset", offset, torch.int64), ("length", length, torch.int32)):
            if t.device.type != "cpu" or t.dtype != dt or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous CPU {dt} tensor")
        if out is None:
            out = torch.empty(self.out_shape(B), dtype=torch.float32, pin_memory=True)
        if out.device.type != "cpu" or out.dtype != torch.float32 or out.numel() != B * self.n_mels * self.frames:
            raise ValueError("out must be CPU fp32 [B,1,n_mels,frames]")
        aug_ptr = None
        if aug is not None:
            if aug.dtype != np.dtype(_lib.AUG_DTYPE) or aug.shape != (B,):
                raise ValueError("aug must be a [B] lm_aug array")
            aug = np.ascontiguousarray(aug)
            aug_ptr = aug.ctypes.data
        if noise is not None and (noise.device.type != "cpu" or noise.dtype != torch.float32
                                  or noise.numel() != B * self.target_length or not noise.is_contiguous()):
            raise ValueError("noise must be contiguous CPU fp32 [B, target_length]")
        fn = self._lib.lm_forward_host_pcm16 if wave.dtype == torch.int16 else self._lib.lm_forward_host
        _lib.check(fn(self._h, wave.data_ptr(), int(wave.numel()), offset.data_ptr(), length.data_ptr(), B, aug_ptr,
                      _ptr(noise), out.data_ptr(), 1 if normalize else 0))
        return out

    def pcm16_decode(self, pcm: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device int16 PCM -> fp32 in [-1, 1) (x / 32768), what ``torchaudio.load`` returns for a 16-bit wav."""
        if pcm.device != self.device or pcm.dtype != torch.int16 or not pcm.is_contiguous():
            raise ValueError(f"pcm must be a contiguous int16 tensor on {self.device}")
        if out is None:
            out = torch.empty(pcm.shape, dtype=torch.float32, device=self.device)
        s = torch.cuda.current_stream(self.device)
        _lib.check(self._lib.lm_pcm16_decode(pcm.data_ptr(), out.data_ptr(), int(pcm.numel()), C.c_void_p(s.cuda_stream)))
        return out
