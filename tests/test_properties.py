"""Property tests of the host-side helpers of the hot path (pad/crop idempotence, roll
composition, mask bounds), with hypothesis.  CPU only: the oracle functions and the host replay of the reference's
random draws (augment.py) are what the GPU kernel is checked against, so their invariants are pinned here."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st

from oracle import logmel_oracle as O


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 400), target=st.integers(1, 300), seed=st.integers(0, 2**31 - 1))
def test_pad_or_crop_is_idempotent_and_centred(n, target, seed):
    """R/src/data/preprocessing.py:70-83: right zero-pad or centre crop; applying it twice changes nothing."""
    x = np.random.RandomState(seed).standard_normal(n).astype(np.float32)
    y = O.pad_or_crop(x, target)
    assert y.shape == (target,)
    np.testing.assert_array_equal(O.pad_or_crop(y, target), y)
    if n <= target:
        np.testing.assert_array_equal(y[:n], x)
        assert not y[n:].any()
    else:
        s = (n - target) // 2
        np.testing.assert_array_equal(y, x[s:s + target])


@settings(max_examples=200, deadline=None)
@given(n=st.integers(1, 300), a=st.integers(-1000, 1000), b=st.integers(-1000, 1000), seed=st.integers(0, 2**31 - 1))
def test_roll_composes_and_wraps(n, a, b, seed):
    """torch.roll semantics (R/src/data/preprocessing.py:90-93): rolls compose additively and wrap modulo the length --
    the kernel's staging relies on both (it reduces the shift modulo T once)."""
    x = np.random.RandomState(seed).standard_normal(n).astype(np.float32)
    np.testing.assert_array_equal(O.roll(O.roll(x, a), b), O.roll(x, a + b))
    np.testing.assert_array_equal(O.roll(x, a), O.roll(x, a % n))
    np.testing.assert_array_equal(O.roll(x, a), np.roll(x, a))
    assert O.roll(x, a)[(0 + a) % n] == x[0]          # +shift = delay


@settings(max_examples=300, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), param=st.integers(1, 80), axis_len=st.integers(1, 400))
def test_mask_intervals_stay_inside_the_axis(seed, param, axis_len):
    """mask_along_axis (torchaudio/functional/functional.py:885-958) as replayed on the host: width < mask_param always;
    0 <= start <= end <= axis length whenever mask_param fits the axis (the reference's 15 of 128 rows and 35 of >= 32
    frames).  With p = 1.0 torchaudio does not clamp mask_param: on a shorter axis the start may be negative (the mask
    then covers [0, end)), which the kernel's interval test (m >= f0 && m < f1) reproduces as is."""
    gen = O.TorchCpuGenerator(seed)
    a, b = O._mask_interval(gen, param, axis_len)
    assert a <= b and b - a < param and a > -param
    if param <= axis_len:
        assert 0 <= a and b <= axis_len


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 2**31 - 1), n_clips=st.integers(1, 5))
def test_replayed_draws_are_reference_shaped(seed, n_clips):
    """Shifts stay within +-0.2 T (truncated toward zero), masks inside the feature map, noise is drawn iff decided."""
    T, n_mels, frames = 4000, 128, 40
    draws = O.replay_augmentation(np.random.RandomState(seed % (2**32)), O.TorchCpuGenerator(seed), n_clips, T, n_mels,
                                  frames, want_noise_values=True)
    assert len(draws) == n_clips
    for d in draws:
        assert abs(d.shift) <= int(0.2 * T)
        assert 0 <= d.f0 <= d.f1 <= n_mels and d.f1 - d.f0 < 15
        assert 0 <= d.t0 <= d.t1 and d.t1 - d.t0 < 35
        assert (d.noise_values is not None) == bool(d.noise)
        if d.noise:
            assert d.noise_values.shape == (T,)


@settings(max_examples=50, deadline=None)
@given(n=st.integers(0, 3 * 16000), seg=st.sampled_from([0.5, 1.0, 2.5, 5.0]), ov=st.sampled_from([0.0, 0.5, 0.75]))
def test_segment_offsets_cover_the_recording(n, seg, ov):
    """R/realtime_analyzer_parallel.py:134-161: windows start every H samples while a full one fits, then one padded
    tail window; together they cover every sample exactly as the reference's loop does."""
    segs = O.segment_offsets(n, 16000, seg, ov)
    S, H = int(seg * 16000), int(int(seg * 16000) * (1 - ov))
    if n == 0:
        assert segs == []
        return
    starts = [s[0] for s in segs]
    assert starts == sorted(starts) and starts[0] == 0
    assert all(b - a == H for a, b in zip(starts, starts[1:]))
    assert all(0 < s[1] <= S for s in segs) and all(s[1] == S for s in segs[:-1])
    assert segs[-1][0] + segs[-1][1] == n or segs[-1][1] == S
    assert max(s[0] + s[1] for s in segs) == n or n > segs[-1][0] + S
