"""Drop-in `AudioPreprocessor` (reference: R/src/data/preprocessing.py:9-151) on the B200 path.

Same constructor, attributes and method names as the reference class.  `preprocess` returns what
the reference returns -- a fresh float32 CPU tensor `[1, n_mels, frames]` -- but the arithmetic
(pad/crop, noise, roll, STFT, mel, dB, masks, normalisation) is one launch of the fused CUDA
kernel behind `lm_forward`.  New, batched entry points (`preprocess_waveform`,
`preprocess_batch`) keep features on the GPU for the training loop.

Seeded behaviour: with `augment=True` the random choices are drawn on the host from the same
global numpy / torch generators, in the same order, as the reference (see augment.py), so
`set_seed(42)` reproduces the reference's shifts, noise and mask intervals exactly.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .augment import draw_fast_augmentation, draw_reference_augmentation
from .plan import LogMelPlan
from .wavio import read_wav

__all__ = ["AudioPreprocessor"]

Wave = Union[torch.Tensor, np.ndarray]


def _as_mono_1d(w: Wave) -> torch.Tensor:
    t = torch.as_tensor(w)
    if t.dim() == 2:   # [C, len] as torchaudio.load returns it
        t = t.mean(dim=0) if t.shape[0] > 1 else t[0]
    if t.dim() != 1:
        raise ValueError(f"waveform must be [len] or [channels, len], got {tuple(t.shape)}")
    return t.to(torch.float32)


class _DeviceTransform:
    """Base of the callable attributes the reference class exposes (`mel_spectrogram`, `amplitude_to_db`,
    `freq_mask`, `time_mask`: R/src/data/preprocessing.py:38-53).  Tensors go to the GPU and come back on the
    device they arrived on, so a caller that treats them like the torchaudio transforms keeps working."""

    def __init__(self, owner: "AudioPreprocessor"):
        self._owner = owner


class _MelSpectrogram(_DeviceTransform):
    """`T.MelSpectrogram(sample_rate, n_fft, hop_length, n_mels, power=2.0)` (preprocessing.py:38-44): `[..., len]`
    -> `[..., n_mels, 1 + len // hop]` mel power.  One launch of the fused kernel with `out_melpow` and no
    normalisation; plans are cached per waveform length."""

    def __call__(self, waveform: torch.Tensor) -> torch.Tensor:
        o = self._owner
        w = torch.as_tensor(waveform, dtype=torch.float32)
        lead, n = w.shape[:-1], int(w.shape[-1])
        plan = o._plan_for_length(n)
        rows = w.reshape(-1, n).to(plan.device).contiguous()
        B = rows.shape[0]
        melp = torch.empty(plan.out_shape(B), dtype=torch.float32, device=plan.device)
        scratch = torch.empty_like(melp)
        plan.forward_dense(rows, normalize=False, out=scratch, out_melpow=melp)
        return melp.reshape(*lead, o.n_mels, plan.frames).to(w.device)


class _AmplitudeToDB(_DeviceTransform):
    """`T.AmplitudeToDB()` (preprocessing.py:46): 10 log10(max(x, 1e-10)), no top_db."""

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        import ctypes as C
        o = self._owner
        src = torch.as_tensor(x, dtype=torch.float32)
        dev = o.plan.device
        d = src.to(dev).contiguous()
        out = torch.empty_like(d)
        with torch.cuda.device(dev):
            s = torch.cuda.current_stream(dev)
            _lib.check(_lib.load().lm_amplitude_to_db(d.data_ptr(), out.data_ptr(), d.numel(), 10.0, 1e-10, 0.0,
                                                      C.c_void_p(s.cuda_stream)))
        return out.to(src.device)


class _AxisMask(_DeviceTransform):
    """`T.FrequencyMasking(15)` / `T.TimeMasking(35)` (preprocessing.py:52-53): one interval per call, drawn from torch's
    global generator exactly as torchaudio.functional.mask_along_axis draws it, filled with 0.0."""

    def __init__(self, owner, mask_param: int, axis: int):
        super().__init__(owner)
        self.mask_param, self.axis = mask_param, axis

    def __call__(self, spec: torch.Tensor, mask_value: float = 0.0) -> torch.Tensor:
        from .augment import mask_interval
        a, b = mask_interval(self.mask_param, spec.shape[self.axis])
        out = spec.clone()
        a = max(a, 0)
        if b > a:
            out.narrow(self.axis, a, min(b, spec.shape[self.axis]) - a).fill_(mask_value)
        return out


class AudioPreprocessor:
    """Audio preprocessing for respiratory sounds: resampling, log-mel, normalisation, augmentation."""

    freq_mask_param = 15   # T.FrequencyMasking(freq_mask_param=15)   preprocessing.py:52
    time_mask_param = 35   # T.TimeMasking(time_mask_param=35)        preprocessing.py:53
    flexible = False

    def __init__(self, sample_rate=16000, n_mels=128, n_fft=2048, hop_length=512, duration=5.0,
                 augment=False, device=None):
        self.sample_rate = sample_rate
        self.n_mels = n_mels
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.duration = duration
        self.augment = augment
        self.target_length = int(sample_rate * duration)
        self._device = device
        self._plan: Optional[LogMelPlan] = None
        self._length_plans = {}
        _lib.load()   # fail loudly at construction if the CUDA library is not built
        # the transform objects of the reference class (preprocessing.py:38-53), GPU-backed
        self.mel_spectrogram = _MelSpectrogram(self)
        self.amplitude_to_db = _AmplitudeToDB(self)
        if augment:
            self.time_stretch = None          # T.TimeStretch is constructed by the reference but never called
            self.freq_mask = _AxisMask(self, self.freq_mask_param, -2)
            self.time_mask = _AxisMask(self, self.time_mask_param, -1)

    # -- plan -------------------------------------------------------------------------------
    @property
    def plan(self) -> LogMelPlan:
        """The GPU plan (window, filterbank, dB constants), created on first use so that the
        object can be constructed in a DataLoader parent and used after CUDA is initialised."""
        if self._plan is None:
            self._plan = LogMelPlan(sample_rate=self.sample_rate, n_fft=self.n_fft, hop_length=self.hop_length,
                                    n_mels=self.n_mels, target_length=self.target_length, device=self._device)
        return self._plan

    def _plan_for_length(self, n: int) -> LogMelPlan:
        """Plan whose target length is the waveform's own length (the transform attributes neither pad nor crop)."""
        if n == self.target_length:
            return self.plan
        if n not in self._length_plans:
            if len(self._length_plans) >= 8:
                self._length_plans.pop(next(iter(self._length_plans))).close()
            self._length_plans[n] = LogMelPlan(sample_rate=self.sample_rate, n_fft=self.n_fft, hop_length=self.hop_length,
                                               n_mels=self.n_mels, target_length=n, device=self._device)
        return self._length_plans[n]

    @property
    def stft_frames(self) -> int:
        return 1 + self.target_length // self.hop_length

    @property
    def frames(self) -> int:
        """Time steps of the returned feature map."""
        return self.stft_frames

    # -- reference helper methods (host-side glue, same semantics) ---------------------------------
    def decode_audio(self, audio_path):
        """File -> (`[1, len]` float32 mono at the FILE's rate, sample_rate).  Host only -- never touches CUDA, so it
        is what forked DataLoader workers call (raw mode); resampling to `sample_rate` happens on the GPU in the
        process that owns the context (`load_audio`, `GpuCollate`)."""
        try:
            import torchaudio
            waveform, sr = torchaudio.load(audio_path)
        except Exception:
            data, sr = read_wav(str(audio_path))
            waveform = torch.from_numpy(data)
        if waveform.shape[0] > 1:
            waveform = torch.mean(waveform, dim=0, keepdim=True)
        return waveform.to(torch.float32), int(sr)

    def resample_to_target(self, waveform, sr):
        """`T.Resample(sr, self.sample_rate)` of the reference (preprocessing.py:63-65) as a CUDA kernel; returns a
        device tensor when it resamples, the input unchanged when the rates agree."""
        if int(sr) == int(self.sample_rate):
            return waveform
        from .resample import get_resampler
        return get_resampler(int(sr), int(self.sample_rate), self.plan.device)(waveform.float())

    def load_audio(self, audio_path):
        """File -> `[1, len]` float32 mono at `sample_rate` (preprocessing.py:55-68)."""
        waveform, sr = self.decode_audio(audio_path)
        return self.resample_to_target(waveform, sr).cpu()

    def pad_or_crop(self, waveform):
        """Right zero-pad or centre-crop to `target_length` (preprocessing.py:70-83)."""
        n = waveform.shape[-1]
        if n < self.target_length:
            return torch.nn.functional.pad(waveform, (0, self.target_length - n))
        if n > self.target_length:
            start = (n - self.target_length) // 2
            return waveform[..., start:start + self.target_length]
        return waveform

    def add_noise(self, waveform, noise_factor=0.005):
        return waveform + torch.randn_like(waveform) * noise_factor

    def time_shift(self, waveform, shift_max=0.2):
        shift = int(np.random.uniform(-shift_max, shift_max) * waveform.shape[-1])
        return torch.roll(waveform, shift, dims=-1)

    def augment_waveform(self, waveform):
        if np.random.random() > 0.5:
            waveform = self.add_noise(waveform)
        if np.random.random() > 0.5:
            waveform = self.time_shift(waveform)
        return waveform

    def augment_spectrogram(self, mel_spec):
        """FrequencyMasking(15) then TimeMasking(35), fill 0.0 (preprocessing.py:105-109)."""
        from .augment import mask_interval
        f0, f1 = mask_interval(self.freq_mask_param, mel_spec.shape[-2])
        t0, t1 = mask_interval(self.time_mask_param, mel_spec.shape[-1])
        out = mel_spec.clone()
        out[..., f0:f1, :] = 0.0
        out[..., :, t0:t1] = 0.0
        return out

    def normalize(self, mel_spec):
        return (mel_spec - mel_spec.mean()) / (mel_spec.std() + 1e-8)

    # -- the hot path ----------------------------------------------------------------------------
    def _draw(self, n: int, fast: bool):
        if not self.augment:
            return None, None
        if fast:
            return draw_fast_augmentation(n, self.target_length, self.n_mels, self.frames,
                                          freq_mask_param=self.freq_mask_param,
                                          time_mask_param=self.time_mask_param), None
        return draw_reference_augmentation(n, self.target_length, self.n_mels, self.frames,
                                           freq_mask_param=self.freq_mask_param,
                                           time_mask_param=self.time_mask_param)

    def _pinned_stage(self, n: int) -> torch.Tensor:
        """One of two grow-only pinned host buffers for packing a batch.  Each carries an event recorded after the copy
        out of it (`_stage_release`); the buffer is reused only once that copy is done, so packing batch k+1 overlaps the
        GPU work of batch k and never scribbles over a copy in flight."""
        slots = self.__dict__.setdefault("_stages", [None, None])
        self._stage_ix = 1 - getattr(self, "_stage_ix", 1)
        slot = slots[self._stage_ix]
        if slot is not None:
            slot[1].synchronize()
        if slot is None or slot[0].numel() < n:
            slot = slots[self._stage_ix] = [torch.empty(max(n, 1 << 16), dtype=torch.float32).pin_memory(), torch.cuda.Event()]
        return slot[0]

    def _stage_release(self) -> None:
        self._stages[self._stage_ix][1].record(torch.cuda.current_stream(self.plan.device))

    @staticmethod
    def _upload_noise(plan: LogMelPlan, aug, noise) -> Optional[torch.Tensor]:
        """Host-drawn noise `[B, T]` -> device, moving only the rows of clips that drew noise (the kernel never reads
        the row of a clip whose noise_scale is 0, so the rest of the device tensor stays uninitialised)."""
        if noise is None:
            return None
        rows = np.nonzero(aug["noise_scale"] != 0)[0] if aug is not None else np.arange(noise.shape[0])
        if len(rows) == noise.shape[0]:
            return noise.to(plan.device, non_blocking=True)
        noise_d = torch.empty(tuple(noise.shape), dtype=torch.float32, device=plan.device)
        if len(rows):
            ix = torch.from_numpy(rows)
            noise_d.index_copy_(0, ix.to(plan.device), noise.index_select(0, ix).to(plan.device, non_blocking=True))
        return noise_d

    def _finish(self, plan: LogMelPlan, wave, offset, length, aug, noise, B: int) -> torch.Tensor:
        """Runs the kernel(s).  Overridden by FlexibleAudioPreprocessor when a resize is needed."""
        aug_d = plan.upload_aug(aug) if aug is not None else None
        return plan.forward(wave, offset, length, aug=aug_d, noise=self._upload_noise(plan, aug, noise))

    def preprocess_batch(self, waveforms: Union[Sequence[Wave], torch.Tensor],
                         lengths: Optional[Sequence[int]] = None, fast_augment: bool = False) -> torch.Tensor:
        """Batched hot path.  `waveforms`: list of 1-D / [C, len] waveforms of any lengths (host
        or device), or a dense `[B, len]` tensor (optionally with per-row `lengths`).
        Returns `[B, 1, n_mels, frames]` float32 on the GPU."""
        plan = self.plan
        dev = plan.device
        if isinstance(waveforms, torch.Tensor) and waveforms.dim() == 2:
            B, n = waveforms.shape
            wave = waveforms.to(device=dev, dtype=torch.float32).contiguous().view(-1)
            offset = torch.arange(B, device=dev, dtype=torch.int64) * n
            lens = torch.full((B,), n, dtype=torch.int32) if lengths is None else torch.as_tensor(lengths, dtype=torch.int32)
            if int(lens.max()) > n:
                raise ValueError("lengths exceed the row length")
            length = lens.to(dev)
        else:
            clips = [_as_mono_1d(w) for w in waveforms]
            B = len(clips)
            lens = [int(c.numel()) for c in clips]
            starts, pos = [], 0
            for n in lens:           # 16-byte aligned clip starts: interior tiles go through TMA
                starts.append(pos)
                pos += (n + 3) & ~3
            packed = torch.empty(max(pos, 4), dtype=torch.float32, device=dev)   # alignment gaps are never read
            # host clips: packed into ONE pinned staging buffer (kept on the object, grow-only) and sent with ONE
            # asynchronous copy; clips that already live on the device (e.g. resampled by GpuCollate) are copied in place
            host_ix = [i for i, c in enumerate(clips) if not c.is_cuda]
            if host_ix:
                lo, hi = starts[host_ix[0]], starts[host_ix[-1]] + lens[host_ix[-1]]
                stage = self._pinned_stage(hi - lo)
                for i in host_ix:
                    if lens[i]:
                        stage[starts[i] - lo:starts[i] - lo + lens[i]].copy_(clips[i])
                if hi > lo:
                    packed[lo:hi].copy_(stage[:hi - lo], non_blocking=True)
                self._stage_release()
            for i, c in enumerate(clips):        # after the block copy: it also covered their (stale) slots
                if c.is_cuda and lens[i]:
                    packed[starts[i]:starts[i] + lens[i]].copy_(c, non_blocking=True)
            wave = packed
            offset = torch.tensor(starts, dtype=torch.int64).to(dev, non_blocking=True)
            length = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
        aug, noise = self._draw(B, fast_augment)
        return self._finish(plan, wave, offset, length, aug, noise, B)

    def preprocess_waveform(self, waveform: Wave) -> torch.Tensor:
        """One in-memory waveform -> `[1, n_mels, frames]` float32 CPU tensor (what `preprocess`
        returns after `load_audio`)."""
        return self.preprocess_batch([waveform])[0].cpu()

    def preprocess(self, audio_path):
        """Complete pipeline for one file (preprocessing.py:118-151): load, pad/crop, waveform
        augmentation, log-mel, dB, SpecAugment, normalise."""
        return self.preprocess_waveform(self.load_audio(audio_path))
