"""Pins the numpy oracle (oracle/logmel_oracle.py) against outputs of the reference classes
(tests/golden/reference_golden.npz, produced by tests/golden/make_golden.py in the build
container).  CPU only."""
import os

import numpy as np
import pytest

from oracle import logmel_oracle as O
from tests.golden.make_golden import PLAIN_CASES, golden_input

# Tolerances (BASELINE.json north_star): log-mel dB within 1e-3 absolute, mel power within 1e-4
# relative.  Relative error is taken against max(|ref|, 1e-6 * clip peak):
# the reference's own fp32 FFT carries an absolute error proportional to the frame peak.
DB_ATOL = 1e-3
MEL_RTOL = 1e-4
NORM_ATOL = 2e-4


def rel_err(x, ref):
    floor = max(1e-6 * np.abs(ref).max(), 1e-30)
    return np.abs(x - ref) / np.maximum(np.abs(ref), floor)


def cfg_for(cls, kw):
    n_fft, hop = 2048, 512
    if cls == "flex":
        n_fft, hop = O.flexible_fft_params(16000, n_fft, hop, kw["duration"])
    return O.OracleConfig(n_fft=n_fft, hop_length=hop, duration=kw["duration"])


def test_constants_match_reference(golden):
    w = O.hann_periodic(2048)
    assert w.shape == (2048,) and w[0] == 0.0 and w[1024] == 1.0
    np.testing.assert_allclose(w, golden["const/window"], atol=1e-7)
    assert abs(float((w.astype(np.float64) ** 2).sum()) - 768.0) < 1e-3
    fb = O.melscale_fbanks_htk(1025, 0.0, 8000.0, 128, 16000)
    ref = golden["const/fb"]
    assert fb.shape == ref.shape == (1025, 128)
    assert int((ref != 0).sum()) == 2024  # SURVEY.md section 8c known answer
    assert not ref[0].any() and not ref[1024].any()
    np.testing.assert_allclose(fb, ref, atol=3e-5)  # torch pow (Sleef) vs libm: f_pts differ by 1 ulp
    assert ((fb != 0) == (ref != 0)).mean() > 0.9999
    fb513 = O.melscale_fbanks_htk(513, 0.0, 8000.0, 128, 16000)
    np.testing.assert_allclose(fb513, golden["const/fb_513"], atol=3e-5)


@pytest.mark.parametrize("case", PLAIN_CASES, ids=[c[0] for c in PLAIN_CASES])
def test_pipeline_stages_match_reference(golden, case):
    name, cls, kw, seed, n, kind = case
    cfg = cfg_for(cls, kw)
    meta = golden[f"{name}/meta"]
    assert (cfg.n_fft, cfg.hop_length, cfg.target_length) == tuple(meta[:3])
    assert cfg.frames == meta[3]  # frame counts are exact
    st = O.logmel(golden_input(seed, n, kind), cfg, flexible=(cls == "flex"), return_stages=True)
    ref_mel, ref_db, ref_norm = (golden[f"{name}/{k}"] for k in ("mel_power", "db", "norm"))
    assert st["mel_power"].shape == ref_mel.shape
    assert st["db"].shape == ref_db.shape == ref_norm.shape
    if kind in ("noise", "uniform"):
        assert rel_err(st["mel_power"], ref_mel).max() < MEL_RTOL
        assert np.abs(st["db"] - ref_db).max() < DB_ATOL
        assert np.abs(st["out"] - ref_norm).max() < NORM_ATOL
    elif kind == "zeros":
        assert (ref_db == -100.0).all() and (st["db"] == -100.0).all()
        assert (ref_norm == 0.0).all() and (st["out"] == 0.0).all()
    else:
        # tonal input: the reference's own fp32 FFT is ~4e-2 relative away from fp64 in the
        # leakage skirts; compare where the reference is trustworthy.
        strong = ref_mel > 1e-4 * ref_mel.max()
        assert rel_err(st["mel_power"], ref_mel)[strong].max() < 5e-3
        assert np.abs(st["db"] - ref_db)[strong].max() < 5e-2


def test_floor_is_exactly_minus_100(golden):
    db = golden["short_1p3s_pad_5s/db"]
    frac = float((db == -100.0).mean())
    assert 0.70 < frac < 0.75  # 72.6 % of bins at the floor
    st = O.logmel(golden_input(2, 20800), O.OracleConfig(), return_stages=True)
    assert ((st["db"] == -100.0) == (db == -100.0)).mean() > 0.999


def test_rng_streams(golden):
    g = O.TorchCpuGenerator(42)
    np.testing.assert_array_equal(g.rand(8), golden["rng/torch_rand_seed42"])
    np.testing.assert_allclose(golden["rng/torch_rand_seed42"][:4],
                               [0.8822692633, 0.9150039554, 0.3828637600, 0.9593056440], atol=1e-9)
    g = O.TorchCpuGenerator(7)
    z = g.randn(48000)
    np.testing.assert_allclose(z[:64], golden["rng/torch_randn_seed7_head"], atol=2e-6)
    # randn(48000) consumed exactly 48000 draws of the stream
    np.testing.assert_array_equal(g.rand(1), golden["rng/torch_rand_after_randn"])
    rs = np.random.RandomState(42)
    np.testing.assert_array_equal([rs.random_sample() for _ in range(4)], golden["rng/numpy_random_seed42"])


@pytest.mark.parametrize("tag,dur", [("aug_3s", 3.0), ("aug_5s", 5.0)])
def test_seeded_augmentation_trace_is_bit_exact(golden, tag, dur):
    cfg = O.OracleConfig(duration=dur)
    trace = golden[f"{tag}/trace"]
    draws = O.replay_augmentation(np.random.RandomState(42), O.TorchCpuGenerator(42), len(trace),
                                  cfg.target_length, cfg.n_mels, cfg.frames, want_noise_values=True)
    # an empty interval has no recoverable position in the reference's output: canonicalise
    canon = lambda a, b: (a, b) if b > a else (0, 0)
    got = np.array([[int(d.noise), d.shift, *canon(d.f0, d.f1), *canon(d.t0, d.t1)] for d in draws])
    np.testing.assert_array_equal(got, trace)
    for c, d in enumerate(draws):
        if d.noise:
            np.testing.assert_allclose(d.noise_values[:32], golden[f"{tag}/clip{c}/noise_head"], atol=2e-6)
            s = golden[f"{tag}/clip{c}/noise_sum"]
            assert abs(d.noise_values.astype(np.float64).sum() - s[0]) < 1e-2
        out = O.logmel(golden_input(100 + c, cfg.target_length), cfg, shift=d.shift,
                       noise=d.noise_values, noise_scale=0.005 if d.noise else 0.0,
                       masks=(d.f0, d.f1, d.t0, d.t1))
        assert np.abs(out - golden[f"{tag}/clip{c}/norm"]).max() < NORM_ATOL


def test_aug_5s_trace_matches_survey_table(golden):
    # SURVEY.md section 8c table, captured independently during the survey
    expect = [[0, 7423, 105, 118, 137, 150], [1, 0, 52, 57, 140, 148], [0, 0, 69, 71, 0, 0],
              [1, 6658, 27, 40, 59, 69], [0, 10638, 64, 69, 37, 42], [0, 0, 114, 119, 9, 31]]
    np.testing.assert_array_equal(golden["aug_5s/trace"], expect)


def test_segment_offsets_known_answers():
    # R/analysis_results/test_audio_1_results.csv: 15 s, 0.5 s windows, 75 % overlap -> 118 rows,
    # starts every 0.125 s, last row 14.625,15.000
    segs = O.segment_offsets(15 * 16000, 16000, 0.5, 0.75)
    assert len(segs) == 118
    assert segs[1][2] == 0.125 and segs[-1][2:] == (14.625, 15.0)
    assert len(O.segment_offsets(15 * 16000, 16000, 1.0, 0.5)) == 30
    hour = O.segment_offsets(3600 * 16000, 16000, 1.0, 0.5)
    assert len(hour) == 7200 and hour[-1][1] == 8000  # 7199 full + 1 padded
    assert len(O.segment_offsets(3600 * 16000, 16000, 5.0, 0.5)) == 1440
    assert O.segment_offsets(100, 16000, 1.0, 0.5) == [(0, 100, 0.0, 100 / 16000)]
    assert O.segment_offsets(0, 16000, 1.0, 0.5) == []


@pytest.mark.parametrize("sr", [4000, 10000, 44100, 8000, 22050, 48000])
def test_oracle_resample_matches_torchaudio_golden(sr):
    """oracle.resample (float64) against T.Resample(sr, 16000) run by tests/golden/make_resample_golden.py --
    the reference's own call (R/src/data/preprocessing.py:63-65).  The reference sums up to 475 float32
    products per sample: its own distance to float64 is ~3e-6."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "resample_golden.npz"))
    y = O.resample(g[f"in/{sr}"], sr, 16000)
    assert y.shape == g[f"out/{sr}"].shape
    assert np.abs(y - g[f"out/{sr}"]).max() < 1e-5
    k, width, o, q = O.sinc_resample_kernel(sr, 16000)
    assert k.shape == (q, 2 * width + o)


def test_seeded_64_clip_trace_is_bit_exact(golden):
    """BASELINE configs[3] at its stated size (batch 32 / 64, R/config_segmented.yaml:21): the draw replay stays in
    step with the reference over 64 consecutive clips (28 of them consume 48 000 normal draws each), and the
    oracle reproduces the 64 sampled output values the fixture keeps per clip."""
    cfg = O.OracleConfig(duration=3.0)
    trace = golden["aug64_3s/trace"]
    assert trace.shape == (64, 6) and int(trace[:, 0].sum()) == 28
    draws = O.replay_augmentation(np.random.RandomState(42), O.TorchCpuGenerator(42), 64, cfg.target_length,
                                  cfg.n_mels, cfg.frames, want_noise_values=True)
    canon = lambda a, b: (a, b) if b > a else (0, 0)
    got = np.array([[int(d.noise), d.shift, *canon(d.f0, d.f1), *canon(d.t0, d.t1)] for d in draws])
    np.testing.assert_array_equal(got, trace)
    np.testing.assert_array_equal(trace[:6], golden["aug_3s/trace"])   # same seeds: the 6-clip trace is its prefix
    pos = golden["aug64_3s/pos"]
    for c in (0, 1, 31, 32, 63):
        d = draws[c]
        out = O.logmel(golden_input(100 + c, cfg.target_length), cfg, shift=d.shift, noise=d.noise_values,
                       noise_scale=0.005 if d.noise else 0.0, masks=(d.f0, d.f1, d.t0, d.t1))
        assert np.abs(out.reshape(-1)[pos] - golden["aug64_3s/samples"][c]).max() < NORM_ATOL
        assert abs(np.abs(out).mean() - golden["aug64_3s/absmean"][c]) < 1e-4


@pytest.mark.parametrize("case", PLAIN_CASES, ids=[c[0] for c in PLAIN_CASES])
def test_torchaudio_port_reproduces_reference_outputs(golden, case):
    """oracle/torchaudio_port.py is bench.py's CPU baseline and `--impl reference` arm (the reference classes cannot
    travel to the GPU box).  It must BE the reference's arithmetic: same torchaudio objects, same call order --
    so its outputs equal the reference-generated goldens exactly."""
    torch = pytest.importorskip("torch")
    pytest.importorskip("torchaudio")
    from oracle.torchaudio_port import ReferencePipeline
    name, cls, kw, seed, n, kind = case
    torch.set_num_threads(1)
    ref = ReferencePipeline(duration=kw["duration"], flexible=(cls == "flex"))
    meta = golden[f"{name}/meta"]
    assert (ref.n_fft, ref.hop_length, ref.target_length) == tuple(meta[:3])
    w = torch.from_numpy(golden_input(seed, n, kind)).unsqueeze(0)
    np.testing.assert_array_equal(ref.mel_power(w)[0].numpy(), golden[f"{name}/mel_power"])
    np.testing.assert_array_equal(ref(w)[0].numpy(), golden[f"{name}/norm"])
    np.testing.assert_array_equal(ref.batched(w)[0, 0].numpy(), golden[f"{name}/norm"])
