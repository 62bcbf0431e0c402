"""CPU oracle for the ICBHI log-mel front end -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's hot path.  It is the *checker*:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
(``audio_classification_icbhi_b200``) never imports anything from ``oracle/`` and has no CPU
fallback.

What is restated, and where it lives in the reference (R/ = /root/reference):

* the 40-line pipeline ``AudioPreprocessor.preprocess``      R/src/data/preprocessing.py:118-151
  and its twin ``FlexibleAudioPreprocessor.preprocess``      R/data/preprocessing_flexible.py:156-192
* the sliding-window cutter ``segment_audio``                R/realtime_analyzer_parallel.py:134-161

The arithmetic itself lives in third-party dependencies that are NOT under /root/reference:
torchaudio (pinned 2.8.0 / 2.9.1 in R/uv.lock:3671,3708; 2.11.0 installed) and torch
(pinned 2.8.0 / 2.9.1, R/uv.lock:3547,3604; 2.11.0 installed).  The published algorithms are
restated here from (TA/ = site-packages/torchaudio):

* spectrogram (reflect pad, framing, window, rFFT, |.|^2)    TA/functional/functional.py:106-145
  + torch.stft's ``center`` padding                          torch/functional.py (stft, ``if center:``)
* HTK mel filterbank                                         TA/functional/functional.py:425-588
* mel projection                                             TA/transforms/_transforms.py:407-419
* amplitude_to_DB (power, top_db=None)                       TA/functional/functional.py:390-391
* mask_along_axis                                            TA/functional/functional.py:885-958
* torch CPU generator: mt19937 -> 24-bit uniform floats, Box-Muller ``normal_fill``
  (aten/src/ATen/native/cpu/DistributionTemplates.h, aten/src/ATen/core/DistributionsHelper.h)

Pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference classes themselves, executed in the
build container from /root/reference by ``tests/golden/make_golden.py`` (committed) and stored
as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this file against them.
The PCM16 temp-wav round trip (soundfile) and librosa resampling used by the analyzers and by
preprocess_icbhi.py are "parity unpinned": those libraries are absent here and no reference
test covers them.

All spectral math is done in float64 unless ``dtype=np.float32`` is requested for the
projection/log steps; the float64 result is the centre of the error ball the reference's own
fp32 result sits in (~1e-6 relative on noise-like inputs).
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

__all__ = [
    "OracleConfig",
    "hann_periodic",
    "torch_linspace_f32",
    "melscale_fbanks_htk",
    "pad_or_crop",
    "roll",
    "reflect_pad",
    "frame_count",
    "stft_power",
    "mel_power",
    "amplitude_to_db",
    "apply_masks",
    "normalize",
    "resize_bilinear_time",
    "logmel",
    "segment_offsets",
    "TorchCpuGenerator",
    "AugDraw",
    "replay_augmentation",
    "flexible_fft_params",
]


# --------------------------------------------------------------------------------------
# configuration (R/src/data/preprocessing.py:20-35, R/data/preprocessing_flexible.py:14-36)
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class OracleConfig:
    sample_rate: int = 16000
    n_mels: int = 128
    n_fft: int = 2048
    hop_length: int = 512
    duration: float = 5.0

    @property
    def target_length(self) -> int:  # preprocessing.py:35
        return int(self.sample_rate * self.duration)

    @property
    def n_freqs(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def frames(self) -> int:
        return frame_count(self.target_length, self.hop_length)


def flexible_fft_params(sample_rate: int, n_fft: int, hop_length: int, duration: float) -> Tuple[int, int]:
    """n_fft / hop override for short segments -- R/data/preprocessing_flexible.py:33-36."""
    if duration < 1.0:
        n_fft = min(1024, int(sample_rate * duration / 2))
        hop_length = n_fft // 4
    return n_fft, hop_length


# --------------------------------------------------------------------------------------
# constants
# --------------------------------------------------------------------------------------
def hann_periodic(n: int) -> np.ndarray:
    """torch.hann_window(n) (periodic=True), the default window of T.Spectrogram
    (TA/transforms/_transforms.py:86-87).  Returned as float32 like the registered buffer."""
    # ATen builds it in float32 as arange(n).mul_(2*pi/n).cos_().mul_(-0.5).add_(0.5)
    # (aten/src/ATen/native/TensorFactories.cpp, hamming_window with alpha = beta = 0.5); the
    # float32 cancellation near the ends (w[1] = 2.3544e-6 vs 2.3531e-6 exact) is part of the
    # reference's arithmetic, so it is kept.
    k = np.arange(n, dtype=np.float32)
    x = (k * np.float32(np.pi * 2.0 / n)).astype(np.float32)
    return (np.cos(x).astype(np.float32) * np.float32(-0.5) + np.float32(0.5)).astype(np.float32)


def torch_linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace(start, end, steps) for float32 on CPU.

    ATen computes ``step = (end - start) / (steps - 1)`` in float32 and fills the first half
    as ``start + step * i`` and the second half as ``end - step * (steps - 1 - i)``
    (aten/src/ATen/native/cpu/RangeFactoriesKernel.cpp, linspace_kernel).
    """
    start32 = np.float32(start)
    end32 = np.float32(end)
    if steps == 1:
        return np.array([start32], dtype=np.float32)
    step = np.float32((end32 - start32) / np.float32(steps - 1))
    i = np.arange(steps, dtype=np.int64)
    half = steps // 2
    lo = (start32 + step * i.astype(np.float32)).astype(np.float32)
    hi = (end32 - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < half, lo, hi).astype(np.float32)


def _hz_to_mel_htk(freq: float) -> float:  # TA/functional/functional.py:438-439
    return 2595.0 * math.log10(1.0 + (freq / 700.0))


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> np.ndarray:
    """HTK triangular filterbank, norm=None -- TA/functional/functional.py:518-588 with
    _create_triangular_filterbank (:491-515).  float32 arithmetic in the same order as torch."""
    all_freqs = torch_linspace_f32(0.0, float(sample_rate // 2), n_freqs)
    m_min = _hz_to_mel_htk(f_min)
    m_max = _hz_to_mel_htk(f_max)
    m_pts = torch_linspace_f32(m_min, m_max, n_mels + 2)
    # _mel_to_hz (htk): 700 * (10 ** (mels / 2595) - 1), elementwise float32
    f_pts = (np.float32(700.0) * (np.power(np.float32(10.0), m_pts / np.float32(2595.0), dtype=np.float32)
                                  - np.float32(1.0))).astype(np.float32)
    f_diff = (f_pts[1:] - f_pts[:-1]).astype(np.float32)
    slopes = (f_pts[None, :] - all_freqs[:, None]).astype(np.float32)
    down = ((np.float32(-1.0) * slopes[:, :-2]) / f_diff[:-1]).astype(np.float32)
    up = (slopes[:, 2:] / f_diff[1:]).astype(np.float32)
    return np.maximum(np.float32(0.0), np.minimum(down, up)).astype(np.float32)


# --------------------------------------------------------------------------------------
# waveform stages
# --------------------------------------------------------------------------------------
def pad_or_crop(wave: np.ndarray, target_length: int) -> np.ndarray:
    """R/src/data/preprocessing.py:70-83 -- right zero-pad, or crop from the centre."""
    wave = np.asarray(wave)
    n = wave.shape[-1]
    if n < target_length:
        pad = [(0, 0)] * (wave.ndim - 1) + [(0, target_length - n)]
        return np.pad(wave, pad)
    if n > target_length:
        start = (n - target_length) // 2
        return wave[..., start:start + target_length]
    return wave


def roll(wave: np.ndarray, shift: int) -> np.ndarray:
    """torch.roll(w, shift, dims=1) -- R/src/data/preprocessing.py:93.  out[i] = w[(i-shift) % T]."""
    return np.roll(wave, shift, axis=-1)


def reflect_pad(wave: np.ndarray, pad: int) -> np.ndarray:
    """F.pad(..., (pad, pad), 'reflect') as used by torch.stft(center=True): the edge sample is
    not repeated."""
    return np.pad(wave, [(0, 0)] * (wave.ndim - 1) + [(pad, pad)], mode="reflect")


def frame_count(length: int, hop: int) -> int:
    """center=True STFT: 1 + length // hop frames."""
    return 1 + length // hop


def stft_power(wave: np.ndarray, n_fft: int, hop: int, window: Optional[np.ndarray] = None) -> np.ndarray:
    """|STFT|^2, shape (..., n_fft//2+1, frames), float64.

    TA/functional/functional.py:123-145 (torch.stft, center=True, reflect, onesided,
    then ``abs().pow(2.0)``)."""
    wave = np.asarray(wave, dtype=np.float64)
    if window is None:
        window = hann_periodic(n_fft)
    window = np.asarray(window, dtype=np.float64)
    padded = reflect_pad(wave, n_fft // 2)
    frames = 1 + (padded.shape[-1] - n_fft) // hop
    idx = (np.arange(frames) * hop)[:, None] + np.arange(n_fft)[None, :]
    tiles = padded[..., idx] * window  # (..., frames, n_fft)
    spec = np.fft.rfft(tiles, axis=-1)
    power = spec.real ** 2 + spec.imag ** 2
    return np.swapaxes(power, -1, -2)


def mel_power(power: np.ndarray, fb: np.ndarray) -> np.ndarray:
    """matmul(spec^T, fb)^T -- TA/transforms/_transforms.py:417.  (..., n_mels, frames)."""
    return np.swapaxes(np.swapaxes(power, -1, -2) @ fb.astype(power.dtype), -1, -2)


def amplitude_to_db(x: np.ndarray, multiplier: float = 10.0, amin: float = 1e-10,
                    db_multiplier: float = 0.0) -> np.ndarray:
    """TA/functional/functional.py:390-391 with the T.AmplitudeToDB() defaults
    (stype='power' -> multiplier 10, ref 1.0 -> db_multiplier 0, top_db None;
    TA/transforms/_transforms.py:324-333)."""
    return multiplier * np.log10(np.maximum(x, amin)) - multiplier * db_multiplier


def apply_masks(db: np.ndarray, f0: int, f1: int, t0: int, t1: int, value: float = 0.0) -> np.ndarray:
    """FrequencyMasking then TimeMasking with already-drawn intervals [f0,f1) x [t0,t1)
    (masked_fill with 0.0 in the dB domain, TA/functional/functional.py:939-953)."""
    out = np.array(db, copy=True)
    out[..., f0:f1, :] = value
    out[..., :, t0:t1] = value
    return out


def normalize(x: np.ndarray) -> np.ndarray:
    """(x - mean) / (std + 1e-8), std unbiased, over the whole clip tensor --
    R/src/data/preprocessing.py:111-116."""
    x = np.asarray(x)
    mean = x.mean(dtype=np.float64)
    std = x.std(dtype=np.float64, ddof=1)
    return (x - mean) / (std + 1e-8)


def resize_bilinear_time(db: np.ndarray, target_steps: int) -> np.ndarray:
    """F.interpolate(size=(n_mels, target), mode='bilinear', align_corners=False) when only the
    time axis changes -- R/data/preprocessing_flexible.py:118-154.  With an unchanged mel axis
    the bilinear kernel degenerates to 1-D linear interpolation along time
    (aten upsample_bilinear2d: src = (dst + 0.5) * scale - 0.5, clamped at 0)."""
    n_in = db.shape[-1]
    if n_in == target_steps:
        return db
    scale = n_in / target_steps
    dst = np.arange(target_steps, dtype=np.float64)
    src = np.maximum((dst + 0.5) * scale - 0.5, 0.0)
    i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    lam = src - i0
    return db[..., i0] * (1.0 - lam) + db[..., i1] * lam


def flexible_target_steps(target_length: int, hop: int) -> int:
    """R/data/preprocessing_flexible.py:129-134."""
    return max(int(np.ceil(target_length / hop)), 32)


# --------------------------------------------------------------------------------------
# the pipeline
# --------------------------------------------------------------------------------------
def logmel(wave: np.ndarray, cfg: OracleConfig = OracleConfig(), *, shift: int = 0,
           noise: Optional[np.ndarray] = None, noise_scale: float = 0.0, gain: float = 1.0,
           masks: Optional[Tuple[int, int, int, int]] = None, do_normalize: bool = True,
           flexible: bool = False, fb: Optional[np.ndarray] = None,
           return_stages: bool = False):
    """R/src/data/preprocessing.py:118-151 on an in-memory 1-D waveform (load_audio bypassed).

    Stage order: pad/crop -> (+ noise_scale*noise) -> roll(shift) -> mel power -> dB ->
    [flexible: resize] -> masks -> normalise.  ``gain`` is not in the reference
    (SURVEY.md headline fact 5); it multiplies the padded/cropped clip and defaults to 1.
    Returns float64 arrays of shape (n_mels, frames)."""
    w = pad_or_crop(np.asarray(wave, dtype=np.float64).reshape(-1), cfg.target_length)
    if gain != 1.0:
        w = w * gain
    if noise is not None and noise_scale != 0.0:
        w = w + np.asarray(noise, dtype=np.float64).reshape(-1) * noise_scale  # :87-88
    if shift:
        w = roll(w, shift)  # :92-93
    if fb is None:
        fb = melscale_fbanks_htk(cfg.n_freqs, 0.0, float(cfg.sample_rate // 2), cfg.n_mels, cfg.sample_rate)
    power = stft_power(w, cfg.n_fft, cfg.hop_length)
    melp = mel_power(power, fb.astype(np.float64))
    db = amplitude_to_db(melp)
    if flexible:
        db = resize_bilinear_time(db, flexible_target_steps(cfg.target_length, cfg.hop_length))
    if masks is not None:
        db = apply_masks(db, *masks)
    out = normalize(db) if do_normalize else db
    if return_stages:
        return {"power": power, "mel_power": melp, "db": db, "out": out}
    return out


# --------------------------------------------------------------------------------------
# sliding windows (R/realtime_analyzer_parallel.py:134-161)
# --------------------------------------------------------------------------------------
def segment_offsets(n_samples: int, sample_rate: int, segment_duration: float,
                    overlap: float) -> List[Tuple[int, int, float, float]]:
    """(start_sample, valid_length, start_time, end_time) per window.  Full windows while
    start+S <= n, then one zero-padded tail window if start < n."""
    seg = int(segment_duration * sample_rate)
    hop = int(seg * (1 - overlap))
    out = []
    start = 0
    while start + seg <= n_samples:
        out.append((start, seg, start / sample_rate, (start + seg) / sample_rate))
        start += hop
    if start < n_samples:
        out.append((start, n_samples - start, start / sample_rate, n_samples / sample_rate))
    return out


# --------------------------------------------------------------------------------------
# RNG replay: the reference's seeded augmentation choices
# --------------------------------------------------------------------------------------
class TorchCpuGenerator:
    """torch's default CPU generator: mt19937 seeded with init_genrand(seed); float32 uniforms
    are ``(u32 & (2**24-1)) * 2**-24`` (ATen/core/DistributionsHelper.h, uniform_real_distribution).

    numpy's legacy RandomState(seed) runs the same mt19937 with the same seeding, so its raw
    32-bit outputs are the torch stream."""

    def __init__(self, seed: int):
        self._rs = np.random.RandomState(seed)

    def _raw(self, n: int) -> np.ndarray:
        return self._rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)

    def rand(self, n: int = 1) -> np.ndarray:
        return ((self._raw(n) & np.uint32((1 << 24) - 1)).astype(np.float32)
                * np.float32(1.0 / (1 << 24))).astype(np.float32)

    def randn(self, n: int) -> np.ndarray:
        """at::normal_ on a contiguous float tensor of >= 16 elements (normal_fill): n uniforms,
        then Box-Muller over blocks of 16 (first 8 = radius source, last 8 = angle source)."""
        if n < 16:
            raise NotImplementedError("scalar normal path (double-precision draws) not restated")
        data = self.rand(n).astype(np.float32)

        def fill16(block: np.ndarray) -> None:
            u1 = np.float32(1.0) - block[:8]
            u2 = block[8:16].copy()
            radius = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
            theta = (np.float32(2.0 * math.pi) * u2).astype(np.float32)
            block[:8] = radius * np.cos(theta)
            block[8:16] = radius * np.sin(theta)

        for i in range(0, n - 15, 16):
            fill16(data[i:i + 16])
        if n % 16 != 0:
            data[n - 16:] = self.rand(16)
            fill16(data[n - 16:])
        return data


@dataclass
class AugDraw:
    """One clip's augmentation decisions, in the reference's draw order."""
    noise: bool
    shift: int
    f0: int
    f1: int
    t0: int
    t1: int
    noise_values: Optional[np.ndarray] = None


def _mask_interval(gen: TorchCpuGenerator, mask_param: int, axis_len: int) -> Tuple[int, int]:
    """TA/functional/functional.py:939-944 in float32."""
    value = np.float32(gen.rand(1)[0] * np.float32(mask_param))
    min_value = np.float32(gen.rand(1)[0] * (np.float32(axis_len) - value))
    start = int(min_value)  # .long(): truncation
    end = start + int(value)
    return start, end


def replay_augmentation(np_rs: np.random.RandomState, torch_gen: TorchCpuGenerator, n_clips: int,
                        target_length: int, n_mels: int, frames: int, *, shift_max: float = 0.2,
                        freq_mask_param: int = 15, time_mask_param: int = 35,
                        want_noise_values: bool = False) -> List[AugDraw]:
    """Replays, clip by clip, the draws of ``augment_waveform`` (R/src/data/preprocessing.py:95-103)
    followed by ``augment_spectrogram`` (:105-109): numpy global stream decides noise / shift,
    torch global stream produces the noise (target_length uniforms) and the four mask draws."""
    out = []
    for _ in range(n_clips):
        noise = bool(np_rs.random_sample() > 0.5)
        noise_values = None
        if noise:
            vals = torch_gen.randn(target_length)
            noise_values = vals if want_noise_values else None
        shift = 0
        if np_rs.random_sample() > 0.5:
            shift = int(np_rs.uniform(-shift_max, shift_max) * target_length)
        f0, f1 = _mask_interval(torch_gen, freq_mask_param, n_mels)
        t0, t1 = _mask_interval(torch_gen, time_mask_param, frames)
        out.append(AugDraw(noise, shift, f0, f1, t0, t1, noise_values))
    return out


# ----------------------------------------------------------------------------------------------
# Resampling step of load_audio (R/src/data/preprocessing.py:63-65: T.Resample(sr, sample_rate))
# ----------------------------------------------------------------------------------------------
def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """TA/functional/functional.py `_get_sinc_resample_kernel` (sinc_interp_hann), float64 throughout.
    Returns (kernel [new/gcd, 2*width + orig/gcd], width, orig/gcd, new/gcd)."""
    g = int(np.gcd(int(orig_freq), int(new_freq)))
    o, q = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(o, q) * rolloff
    width = int(np.ceil(lowpass_filter_width * o / base_freq))
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    t = np.arange(0, -q, -1, dtype=np.float64)[:, None] / q + idx
    t = np.clip(t * base_freq, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * np.pi / lowpass_filter_width / 2.0) ** 2
    t = t * np.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        sinc = np.where(t == 0.0, 1.0, np.sin(t) / t)
    return sinc * window * (base_freq / o), width, o, q


def resample(x: np.ndarray, orig_freq: int, new_freq: int) -> np.ndarray:
    """TA/functional/functional.py `_apply_sinc_resample_kernel` on a 1-D waveform: zero-pad by (width,
    width + orig), correlate with each of the `new` phase kernels at stride `orig`, interleave the phases,
    keep ceil(new * len / orig) samples.  float64 (the reference accumulates in float32)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    if int(orig_freq) == int(new_freq):
        return x.copy()
    kernel, width, o, q = sinc_resample_kernel(orig_freq, new_freq)
    n = x.shape[0]
    xp = np.concatenate([np.zeros(width), x, np.zeros(width + o)])
    K = kernel.shape[1]
    n_blocks = (xp.shape[0] - K) // o + 1
    starts = np.arange(n_blocks) * o
    windows = np.lib.stride_tricks.sliding_window_view(xp, K)[starts]     # [blocks, K]
    y = (windows @ kernel.T).reshape(-1)                                  # block-major, phase-minor
    target = int(np.ceil(q * n / o))
    return y[:target]


def golden_filterbank(n_fft: int = 2048) -> np.ndarray:
    """torchaudio's own float32 filterbank for the path's two configurations (tests/golden/fb_golden.npz,
    written by tests/golden/make_fb_golden.py), as float64 [n_freqs, n_mels]: pass it as ``fb=`` to `logmel`
    to take the ~1e-5 restatement error of `melscale_fbanks_htk` out of a comparison."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fb_golden.npz"))
    fb = np.zeros(tuple(g[f"{n_fft}/shape"]), dtype=np.float64)
    fb[g[f"{n_fft}/k"], g[f"{n_fft}/m"]] = g[f"{n_fft}/v"].astype(np.float64)
    return fb
