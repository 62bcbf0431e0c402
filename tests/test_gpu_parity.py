"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle and the golden
fixtures produced by the reference classes.  Run on the B200 box: pytest -m gpu.

Tolerances are BASELINE.json's: log-mel dB within 1e-3 absolute, mel power within 1e-4
relative (against max(|ref|, 1e-6 * clip peak)), frame counts and mask
indices exact.  The normalised output is (dB - mean)/std with std ~ 10 dB, so 1e-3 dB maps to
~1e-4; 2e-4 absolute is used.
"""
import numpy as np
import pytest
import torch

from oracle import logmel_oracle as O
from tests.golden.make_golden import PLAIN_CASES, golden_input

pytestmark = pytest.mark.gpu

DB_ATOL = 1e-3
MEL_RTOL = 1e-4
NORM_ATOL = 2e-4


def rel_err(x, ref):
    floor = max(1e-6 * np.abs(ref).max(), 1e-30)
    return np.abs(x - ref) / np.maximum(np.abs(ref), floor)


@pytest.fixture(scope="module")
def A():
    import audio_classification_icbhi_b200 as pkg
    return pkg


_plans = {}


def get_plan(A, n_fft=2048, hop=512, target_length=80000, n_mels=128):
    key = (n_fft, hop, target_length, n_mels)
    if key not in _plans:
        _plans[key] = A.LogMelPlan(n_fft=n_fft, hop_length=hop, n_mels=n_mels, target_length=target_length,
                                   device="cuda:0")
    return _plans[key]


def run_clips(plan, clips, aug=None, noise=None, normalize=True, want_stages=True):
    """clips: list of 1-D float32 numpy arrays (ragged)."""
    dev = plan.device
    lengths = np.array([len(c) for c in clips], dtype=np.int32)
    # keep every clip start 16-byte aligned so interior tiles take the TMA path
    starts, pos = [], 0
    for n in lengths:
        starts.append(pos)
        pos += (int(n) + 3) & ~3
    packed = np.zeros(max(pos, 4), dtype=np.float32)
    for s, c in zip(starts, clips):
        packed[s:s + len(c)] = c
    wave = torch.from_numpy(packed).to(dev)
    offset = torch.tensor(starts, dtype=torch.int64, device=dev)
    length = torch.from_numpy(lengths).to(dev)
    B = len(clips)
    shape = plan.out_shape(B)
    out_db = torch.full(shape, float("nan"), device=dev) if want_stages else None
    out_mp = torch.full(shape, float("nan"), device=dev) if want_stages else None
    aug_d = plan.upload_aug(aug) if aug is not None else None
    noise_d = torch.from_numpy(noise).to(dev) if noise is not None else None
    out = plan.forward(wave, offset, length, aug=aug_d, noise=noise_d, normalize=normalize,
                       out_db=out_db, out_melpow=out_mp)
    torch.cuda.synchronize()
    res = {"out": out.cpu().numpy()[:, 0]}
    if want_stages:
        res["db"] = out_db.cpu().numpy()[:, 0]
        res["mel_power"] = out_mp.cpu().numpy()[:, 0]
    return res


def test_library_is_loaded_and_reports_geometry(A):
    plan = get_plan(A)
    info = plan.info()
    assert plan.frames == 157 and info["frames"] == 157 and info["n_freqs"] == 1025
    assert info["fb_nnz"] == 2024
    assert info["bytes_per_clip"] == 400384
    assert info["threads_per_cta"] == 512
    import os
    with open("/proc/self/maps") as f:
        assert "liblogmel_b200.so" in f.read()


STD_CASES = [c for c in PLAIN_CASES if c[1] == "std" or c[0] in ("flex_1s", "flex_0p5s", "flex_1s_short_pad")]


@pytest.mark.parametrize("case", STD_CASES, ids=[c[0] for c in STD_CASES])
def test_golden_cases(A, golden, case):
    """CUDA path vs the outputs of the reference classes (tests/golden) and vs the fp64 oracle."""
    name, cls, kw, seed, n, kind = case
    n_fft, hop = (2048, 512) if cls == "std" else O.flexible_fft_params(16000, 2048, 512, kw["duration"])
    cfg = O.OracleConfig(n_fft=n_fft, hop_length=hop, duration=kw["duration"])
    plan = get_plan(A, n_fft, hop, cfg.target_length)
    assert plan.frames == cfg.frames == golden[f"{name}/meta"][3]
    x = golden_input(seed, n, kind)
    got = run_clips(plan, [x])
    ref_mel, ref_db, ref_norm = (golden[f"{name}/{k}"] for k in ("mel_power", "db", "norm"))
    st = O.logmel(x, cfg, return_stages=True)
    assert got["out"].shape[1:] == ref_norm.shape
    assert np.isfinite(got["out"]).all() and np.isfinite(got["db"]).all()
    if kind in ("noise", "uniform"):
        assert rel_err(got["mel_power"][0], ref_mel).max() < MEL_RTOL
        assert rel_err(got["mel_power"][0], st["mel_power"]).max() < MEL_RTOL
        assert np.abs(got["db"][0] - ref_db).max() < DB_ATOL
        assert np.abs(got["db"][0] - st["db"]).max() < DB_ATOL
        assert np.abs(got["out"][0] - ref_norm).max() < NORM_ATOL
        assert np.abs(got["out"][0] - st["out"]).max() < NORM_ATOL
    elif kind == "zeros":
        assert (got["db"][0] == -100.0).all()
        assert (got["out"][0] == 0.0).all()
    else:  # tonal: only as good as the reference's own distance to fp64
        strong = st["mel_power"] > 1e-4 * st["mel_power"].max()
        assert rel_err(got["mel_power"][0], st["mel_power"])[strong].max() < 1e-3
        assert np.abs(got["db"][0] - st["db"])[strong].max() < 1e-2
    # the -100 dB floor is exact wherever the reference sits on it
    floor_ref = ref_db == -100.0
    if floor_ref.any():
        assert (got["db"][0][floor_ref] == -100.0).mean() > 0.999


@pytest.mark.parametrize("tag,dur", [("aug_3s", 3.0), ("aug_5s", 5.0)])
def test_seeded_augmentation_matches_reference(A, golden, tag, dur):
    """Config 4: the reference's seeded choices replayed on the host, applied on the GPU."""
    cfg = O.OracleConfig(duration=dur)
    plan = get_plan(A, 2048, 512, cfg.target_length)
    trace = golden[f"{tag}/trace"]
    n = len(trace)
    draws = O.replay_augmentation(np.random.RandomState(42), O.TorchCpuGenerator(42), n, cfg.target_length,
                                  cfg.n_mels, cfg.frames, want_noise_values=True)
    aug = A.make_aug_array(n)
    noise = np.zeros((n, cfg.target_length), dtype=np.float32)
    clips = []
    for c, d in enumerate(draws):
        aug[c]["shift"] = d.shift
        aug[c]["noise_scale"] = 0.005 if d.noise else 0.0
        aug[c]["f0"], aug[c]["f1"], aug[c]["t0"], aug[c]["t1"] = d.f0, d.f1, d.t0, d.t1
        if d.noise:
            noise[c] = d.noise_values
        clips.append(golden_input(100 + c, cfg.target_length))
    got = run_clips(plan, clips, aug=aug, noise=noise)
    for c, d in enumerate(draws):
        ref = golden[f"{tag}/clip{c}/norm"]
        assert np.abs(got["out"][c] - ref).max() < NORM_ATOL
        # mask indices are bit-exact: exactly the reference's rows / frames are zero in dB
        z = got["db"][c] == 0.0
        rows = np.nonzero(z.all(axis=1))[0]
        cols = np.nonzero(z.all(axis=0))[0]
        f0, f1 = (rows[0], rows[-1] + 1) if len(rows) else (0, 0)
        t0, t1 = (cols[0], cols[-1] + 1) if len(cols) else (0, 0)
        assert [f0, f1, t0, t1] == list(trace[c][2:])


def test_ragged_batch_matches_oracle(A):
    """Config 3 in miniature: variable lengths (pad and centre-crop), unaligned starts,
    empty clip, more clips than one wave of CTAs would hold."""
    cfg = O.OracleConfig()
    plan = get_plan(A)
    rs = np.random.RandomState(11)
    lens = [80000, 1, 0, 3200, 41234, 79999, 80001, 112001, 259200, 1025, 2047, 65537]
    clips = [(rs.standard_normal(n) * 0.1).astype(np.float32) for n in lens]
    got = run_clips(plan, clips)
    for i, x in enumerate(clips):
        st = O.logmel(x, cfg, return_stages=True)
        assert rel_err(got["mel_power"][i], st["mel_power"]).max() < MEL_RTOL, (i, lens[i])
        assert np.abs(got["db"][i] - st["db"]).max() < DB_ATOL, (i, lens[i])
        assert np.abs(got["out"][i] - st["out"]).max() < NORM_ATOL, (i, lens[i])


def test_unaligned_offsets_take_the_gather_path(A):
    cfg = O.OracleConfig(duration=3.0)
    plan = get_plan(A, 2048, 512, cfg.target_length)
    dev = plan.device
    rs = np.random.RandomState(5)
    base = (rs.standard_normal(48000 * 3 + 7) * 0.1).astype(np.float32)
    starts = [1, 48003, 96006]
    wave = torch.from_numpy(base).to(dev)
    out = plan.forward(wave, torch.tensor(starts, dtype=torch.int64, device=dev),
                       torch.full((3,), 48000, dtype=torch.int32, device=dev))
    torch.cuda.synchronize()
    for i, s in enumerate(starts):
        ref = O.logmel(base[s:s + 48000], cfg)
        assert np.abs(out[i, 0].cpu().numpy() - ref).max() < NORM_ATOL


def test_tma_and_plain_staging_agree_bitwise(A):
    plan = get_plan(A)
    rs = np.random.RandomState(3)
    clips = [(rs.standard_normal(80000) * 0.1).astype(np.float32) for _ in range(5)]
    plan.set("tma", 1)
    a = run_clips(plan, clips, want_stages=False)["out"]
    plan.set("tma", 0)
    b = run_clips(plan, clips, want_stages=False)["out"]
    plan.set("tma", 1)
    np.testing.assert_array_equal(a, b)


def test_determinism_and_batch_independence(A):
    """Same clip alone, in a batch, and on a 1-CTA grid: bit-identical features."""
    plan = get_plan(A)
    rs = np.random.RandomState(8)
    clips = [(rs.standard_normal(80000) * 0.1).astype(np.float32) for _ in range(300)]
    full = run_clips(plan, clips, want_stages=False)["out"]
    again = run_clips(plan, clips, want_stages=False)["out"]
    np.testing.assert_array_equal(full, again)
    solo = run_clips(plan, [clips[217]], want_stages=False)["out"]
    np.testing.assert_array_equal(full[217], solo[0])
    plan.set("max_ctas", 1)
    one = run_clips(plan, clips[:3], want_stages=False)["out"]
    plan.set("max_ctas", 0)
    np.testing.assert_array_equal(full[:3], one)


def test_gain_noise_shift_and_philox(A):
    cfg = O.OracleConfig(duration=3.0)
    plan = get_plan(A, 2048, 512, cfg.target_length)
    rs = np.random.RandomState(21)
    T = cfg.target_length
    clips = [(rs.standard_normal(n) * 0.1).astype(np.float32) for n in (T, T - 5000, T + 901, T)]
    noise = rs.standard_normal((4, T)).astype(np.float32)
    aug = A.make_aug_array(4)
    aug["shift"] = [-9599, 9599, 1, 0]
    aug["noise_scale"] = [0.005, 0.0, 0.01, 0.0]
    aug["gain"] = [1.0, 0.5, 2.0, 1.0]
    aug["f0"], aug["f1"], aug["t0"], aug["t1"] = [0, 127, 5, 0], [14, 128, 5, 0], [0, 93, 10, 0], [34, 94, 11, 0]
    got = run_clips(plan, clips, aug=aug, noise=noise)
    for i in range(4):
        ref = O.logmel(clips[i], cfg, shift=int(aug[i]["shift"]), noise=noise[i], noise_scale=float(aug[i]["noise_scale"]),
                       gain=float(aug[i]["gain"]), masks=tuple(int(aug[i][k]) for k in ("f0", "f1", "t0", "t1")))
        assert np.abs(got["out"][i] - ref).max() < NORM_ATOL, i
    # throughput-mode noise: on-device Philox N(0,1).  Statistical check on a silent clip: the
    # mel power of pure noise of variance s^2 is s^2 * sum(w^2) * sum_k fb[k, m].
    aug2 = A.make_aug_array(1)
    aug2["noise_scale"] = 0.01
    aug2["seed"] = 1234
    r = run_clips(plan, [np.zeros(T, dtype=np.float32)], aug=aug2, noise=None)
    fb = O.melscale_fbanks_htk(1025, 0.0, 8000.0, 128, 16000).astype(np.float64)
    expect = (0.01 ** 2) * 768.0 * fb.sum(axis=0)
    mean_mp = r["mel_power"][0][:, 4:-4].mean(axis=1)
    assert np.abs(mean_mp / expect - 1.0)[32:].max() < 0.35  # ~86 overlapping frames: chi-square scatter
    r2 = run_clips(plan, [np.zeros(T, dtype=np.float32)], aug=aug2, noise=None)
    np.testing.assert_array_equal(r["out"], r2["out"])


def test_headline_batch_properties(A):
    """BASELINE configs[1] at full size (4096 x 5 s): size-independent properties.
    Per-clip normalisation => every clip has mean 0 / unbiased std 1; duplicated clips give
    identical features wherever they sit in the batch; spot clips match the oracle."""
    plan = get_plan(A)
    dev = plan.device
    B, T = 4096, 80000
    g = torch.Generator(device=dev).manual_seed(1234)
    clips = torch.randn(B, T, generator=g, device=dev) * 0.1
    clips[4095] = clips[0]
    clips[2048] = clips[1]
    out = plan.forward_dense(clips)
    torch.cuda.synchronize()
    assert out.shape == (B, 1, 128, 157)
    assert torch.isfinite(out).all()
    flat = out.view(B, -1).double()
    assert flat.mean(dim=1).abs().max().item() < 1e-5
    assert (flat.std(dim=1) - 1.0).abs().max().item() < 1e-5
    assert torch.equal(out[4095], out[0]) and torch.equal(out[2048], out[1])
    cfg = O.OracleConfig()
    for i in (0, 1, 777, 4094):
        ref = O.logmel(clips[i].cpu().numpy(), cfg)
        assert np.abs(out[i, 0].cpu().numpy() - ref).max() < NORM_ATOL


def test_linearity_in_db(A):
    """Scaling the waveform by g shifts un-normalised dB by 20*log10(g) and leaves the
    normalised features unchanged (away from the -100 dB floor)."""
    plan = get_plan(A)
    rs = np.random.RandomState(2)
    x = (rs.standard_normal(80000) * 0.1).astype(np.float32)
    r = run_clips(plan, [x, 4.0 * x])
    np.testing.assert_allclose(r["db"][1] - r["db"][0], 20 * np.log10(4.0), atol=2e-4)
    np.testing.assert_allclose(r["out"][1], r["out"][0], atol=1e-4)


def test_host_path_matches_device_path(A):
    """lm_forward_host (host buffers, chunked pipeline) == lm_forward, bit for bit."""
    plan = get_plan(A)
    rs = np.random.RandomState(4)
    B, T = 700, 80000
    host = torch.from_numpy((rs.standard_normal((B, T)) * 0.1).astype(np.float32)).pin_memory()
    offset = torch.arange(B, dtype=torch.int64) * T
    length = torch.full((B,), T, dtype=torch.int32)
    out_h = plan.forward_host(host.view(-1), offset, length)
    out_d = plan.forward_dense(host.to(plan.device))
    torch.cuda.synchronize()
    assert torch.equal(out_h, out_d.cpu())


def test_pcm16_host_path_matches_decoded_floats(A):
    """lm_forward_host_pcm16 (int16 over PCIe, decoded on the device) == lm_forward on pcm / 32768, bit for
    bit, for ragged clips that start at odd sample offsets; lm_pcm16_decode itself is exact."""
    plan = get_plan(A)
    rs = np.random.RandomState(5)
    lens = [80000, 12345, 80001, 0, 99999, 80000, 7, 64000]
    starts, pos = [], 3
    for n in lens:
        starts.append(pos)
        pos += n + 5
    pcm = torch.from_numpy(rs.randint(-32768, 32768, size=pos).astype(np.int16)).pin_memory()
    offset = torch.tensor(starts, dtype=torch.int64)
    length = torch.tensor(lens, dtype=torch.int32)
    out_h = plan.forward_host(pcm, offset, length)
    dec = plan.pcm16_decode(pcm.to(plan.device))
    assert torch.equal(dec.cpu(), pcm.float() / 32768.0)
    out_d = plan.forward(dec, offset.to(plan.device), length.to(plan.device))
    torch.cuda.synchronize()
    assert torch.equal(out_h, out_d.cpu())
    # against the oracle on the decoded floats
    ref = O.logmel((pcm[starts[0]:starts[0] + lens[0]].float() / 32768.0).numpy(), O.OracleConfig())
    assert np.abs(out_h[0, 0].numpy() - ref).max() < NORM_ATOL


@pytest.mark.parametrize("n_fft,hop", [(2048, 512), (1024, 256)])
def test_fused_pcm16_kernel_is_bit_identical_to_decode_then_forward(A, n_fft, hop):
    """lm_forward_pcm16 (raw 16-bit samples staged by the bulk copy and expanded in shared memory, or converted sample
    by sample on the gather path) == lm_pcm16_decode + lm_forward, bit for bit: clips that start on 8-sample boundaries
    (bulk path) and on odd ones (gather path), short / empty / cropped clips, augmented clips (roll, gain, host noise,
    on-device noise, masks), normalised and dB outputs, and a small batch that is split over the CTAs."""
    plan = get_plan(A, n_fft, hop)
    dev = plan.device
    rs = np.random.RandomState(31)
    lens = [80000, 80000, 12345, 80001, 0, 99999, 80000, 7, 64000, 80000, 80000, 3000]
    starts, pos = [], 0
    for i, n in enumerate(lens):
        pos = (pos + 7) // 8 * 8 + (0 if i % 3 else 0)   # 16-byte aligned starts ...
        if i in (3, 6, 10):
            pos += 3                                        # ... except these: the gather path
        starts.append(pos)
        pos += n
    pcm = torch.from_numpy(rs.randint(-32768, 32768, size=pos + 8).astype(np.int16)).to(dev)
    offset = torch.tensor(starts, dtype=torch.int64, device=dev)
    length = torch.tensor(lens, dtype=torch.int32, device=dev)
    dec = plan.pcm16_decode(pcm)
    B = len(lens)
    aug = A.make_aug_array(B)
    aug["shift"][1], aug["gain"][1] = 4001, 0.5
    aug["noise_scale"][5], aug["seed"][5] = 0.01, 99
    aug["noise_scale"][9] = 0.005
    aug["f0"][2], aug["f1"][2], aug["t0"][2], aug["t1"][2] = 3, 17, 10, 40
    noise = torch.from_numpy(rs.standard_normal((B, plan.target_length)).astype(np.float32)).to(dev)
    d_aug = plan.upload_aug(aug)
    for kw in (dict(), dict(normalize=False), dict(aug=d_aug, noise=noise), dict(aug=d_aug)):
        want = plan.forward(dec, offset, length, **kw)
        got = plan.forward_pcm16(pcm, offset, length, **kw)
        torch.cuda.synchronize()
        assert torch.equal(got, want), kw.keys()
    # small batch: the clips' tiles are dealt to many CTAs; also one launch per clip
    for i in (0, 3, 8):
        want = plan.forward(dec, offset[i:i + 1], length[i:i + 1])
        got = plan.forward_pcm16(pcm, offset[i:i + 1], length[i:i + 1])
        torch.cuda.synchronize()
        assert torch.equal(got, want), i
    # a few hundred full-length clips: several clips per group, dynamic scheduling
    Bn, T = 600, plan.target_length
    big = torch.from_numpy(rs.randint(-20000, 20000, size=Bn * T).astype(np.int16)).to(dev)
    off = torch.arange(Bn, device=dev, dtype=torch.int64) * T
    ln = torch.full((Bn,), T, device=dev, dtype=torch.int32)
    want = plan.forward(plan.pcm16_decode(big), off, ln)
    got = plan.forward_pcm16(big, off, ln)
    torch.cuda.synchronize()
    assert torch.equal(got, want)


def test_icbhi_sized_ragged_corpus(A):
    """BASELINE configs[2] at full size on one GPU: ~6900 respiratory cycles of lognormal length
    (0.2 .. 16.2 s), packed back to back, pad / centre-crop to 5 s.  Properties that do not
    depend on the size (per-clip mean 0 / std 1, frame count, exact -100 dB floor in the padding of short
    clips) plus spot clips -- the shortest, the longest, a cropped and a padded one -- against the oracle."""
    plan = get_plan(A)
    dev = plan.device
    n, T = 6900, 80000
    rs = np.random.RandomState(0)
    secs = np.clip(rs.lognormal(np.log(2.5), 0.5, n), 0.2, 16.2)
    lens = (secs * 16000).astype(np.int64)
    starts = np.concatenate([[0], np.cumsum((lens + 3) // 4 * 4)[:-1]])
    g = torch.Generator(device=dev).manual_seed(1)
    wave = torch.randn(int(starts[-1] + lens[-1]) + 4, generator=g, device=dev) * 0.1
    offset = torch.from_numpy(starts).to(dev)
    length = torch.from_numpy(lens.astype(np.int32)).to(dev)
    db = torch.empty(plan.out_shape(n), device=dev)
    out = plan.forward(wave, offset, length, out_db=db)
    torch.cuda.synchronize()
    assert out.shape == (n, 1, 128, 157) and torch.isfinite(out).all()
    flat = out.view(n, -1).double()
    assert flat.mean(dim=1).abs().max().item() < 1e-5
    assert (flat.std(dim=1) - 1.0).abs().max().item() < 1e-5
    # frames that lie entirely in the zero padding sit exactly at the floor
    short = int(np.argmin(lens))
    first_silent = (int(lens[short]) + 1024) // 512 + 1
    assert (db[short, 0, :, first_silent:] == -100.0).all()
    cfg = O.OracleConfig()
    w = wave.cpu().numpy()
    cropped = int(np.argmax((lens > T) & (lens < 2 * T)))
    for i in (short, int(np.argmax(lens)), cropped, 0, n - 1):
        ref = O.logmel(w[starts[i]:starts[i] + lens[i]], cfg)
        assert np.abs(out[i, 0].cpu().numpy() - ref).max() < NORM_ATOL, i


def test_accuracy_with_the_reference_filterbank_is_at_fp32_level(A):
    """With torchaudio's own float32 filterbank in the oracle (tests/golden/fb_golden.npz) instead of the numpy
    restatement of melscale_fbanks, the float64 oracle isolates the arithmetic of the path: on noise-like
    clips the CUDA mel power is within 5e-6 relative (bar: 1e-4), dB within 5e-5 (bar: 1e-3) -- the level
    of the reference's own float32 pipeline (1.0e-6 / 4.5e-6)."""
    plan = get_plan(A)
    fb = O.golden_filterbank(2048)
    rs = np.random.RandomState(11)
    for x in (rs.standard_normal(80000) * 0.1, rs.uniform(-1, 1, 80000), rs.standard_normal(30000) * 1e-3):
        x = x.astype(np.float32)
        r = run_clips(plan, [x])
        st = O.logmel(x, O.OracleConfig(), return_stages=True, fb=fb)
        assert rel_err(r["mel_power"][0], st["mel_power"]).max() < 5e-6
        assert np.abs(r["db"][0] - st["db"]).max() < 5e-5
        assert np.abs(r["out"][0] - st["out"]).max() < 1e-5


def test_silent_tiles_of_short_clips(A):
    """Tiles that lie entirely in a plain clip's zero padding skip FFT and mel (silent_from in the clip context):
    mel power exactly 0 and dB exactly the floor there, masks still applied, statistics (hence the normalised
    output) as if they had been computed; rolled / noisy / gained clips of the same length take the full path."""
    plan = get_plan(A)
    rs = np.random.RandomState(21)
    lens = [0, 1, 3000, 20000, 41000, 79000, 80000]
    clips = [(rs.standard_normal(n) * 0.1).astype(np.float32) for n in lens]
    aug = A.make_aug_array(len(clips))
    aug["f0"], aug["f1"], aug["t0"], aug["t1"] = 10, 21, 100, 131          # masks only: the clips stay "plain"
    r = run_clips(plan, clips, aug=aug)
    cfg = O.OracleConfig()
    for i, x in enumerate(clips):
        st = O.logmel(x, cfg, return_stages=True, masks=(10, 21, 100, 131), fb=O.golden_filterbank(2048))
        assert np.abs(r["out"][i] - st["out"]).max() < NORM_ATOL, lens[i]
        assert np.abs(r["db"][i] - st["db"]).max() < DB_ATOL, lens[i]
        first_silent = (lens[i] + 1024) // 512 + 1                            # first frame that sees no sample
        if first_silent < 157:
            tail_db, tail_mp = r["db"][i][:, first_silent:], r["mel_power"][i][:, first_silent:]
            assert (tail_mp == 0.0).all()
            expect = np.full_like(tail_db, -100.0)
            expect[10:21, :] = 0.0
            lo = max(100 - first_silent, 0)
            expect[:, lo:max(131 - first_silent, 0)] = 0.0
            np.testing.assert_array_equal(tail_db, expect)
    # the same short clip with a roll is not plain: full path, still equal to the oracle
    aug2 = A.make_aug_array(1)
    aug2["shift"] = 5000
    r2 = run_clips(plan, [clips[3]], aug=aug2)
    ref2 = O.logmel(clips[3], cfg, shift=5000, fb=O.golden_filterbank(2048))
    assert np.abs(r2["out"][0] - ref2).max() < NORM_ATOL


def test_forward_is_cuda_graph_capturable(A):
    """lm_forward allocates nothing and only enqueues on the caller's stream, so a serving loop can capture it
    in a CUDA graph: replaying the graph on new samples in the same buffers gives the eager result bit for bit."""
    plan = get_plan(A)
    dev = plan.device
    B, T = 64, 80000
    g = torch.Generator(device=dev).manual_seed(9)
    clips = torch.randn(B, T, generator=g, device=dev) * 0.1
    offset = torch.arange(B, device=dev, dtype=torch.int64) * T
    length = torch.full((B,), T, device=dev, dtype=torch.int32)
    out = torch.empty(plan.out_shape(B), device=dev)
    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        plan.forward(clips.view(-1), offset, length, out=out)       # warm-up outside the capture
    torch.cuda.current_stream(dev).wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        plan.forward(clips.view(-1), offset, length, out=out)
    clips.copy_(torch.randn(B, T, generator=g, device=dev) * 0.1)    # new samples, same buffers
    graph.replay()
    torch.cuda.synchronize()
    eager = plan.forward(clips.view(-1), offset, length)
    torch.cuda.synchronize()
    assert torch.equal(out, eager)


def test_small_batch_split_is_bit_identical(A):
    """Small batches cut every clip into tile ranges handled by different groups; the
    statistics are exact integer sums, so the features equal the one-group-per-clip path bit for bit --
    ragged clips, silent tiles, masks and the un-normalised outputs included."""
    plan = get_plan(A)
    rs = np.random.RandomState(21)
    T = plan.target_length
    for lens in ([T], [T, 30000, T + 777], [int(rs.randint(2000, 2 * T)) for _ in range(32)]):
        clips = [(rs.standard_normal(n) * 0.1).astype(np.float32) for n in lens]
        aug = A.plan.make_aug_array(len(clips))
        aug["f0"], aug["f1"], aug["t0"], aug["t1"] = 10, 17, 100, 131
        aug["shift"][::2] = 1234
        aug["shift"][1::2] = -3 * T - 17      # any shift is legal for torch.roll
        plan.set("split", 1)
        ref = run_clips(plan, clips, aug=aug)
        plan.set("split", 0)
        got = run_clips(plan, clips, aug=aug)
        for key in ("out", "db", "mel_power"):
            np.testing.assert_array_equal(got[key], ref[key])
        plan.set("split", 3)
        got3 = run_clips(plan, clips, aug=aug, want_stages=False)
        plan.set("split", 0)
        np.testing.assert_array_equal(got3["out"], ref["out"])
    # the roll wraps: a shift of -3T - 17 is a shift of -17
    x = (rs.standard_normal(T) * 0.1).astype(np.float32)
    a1, a2 = A.plan.make_aug_array(1), A.plan.make_aug_array(1)
    a1["shift"], a2["shift"] = -3 * T - 17, -17
    np.testing.assert_array_equal(run_clips(plan, [x], aug=a1)["out"], run_clips(plan, [x], aug=a2)["out"])


def test_non_hann_window_is_refused(A):
    """The kernel relies on w[n + N/2] = 1 - w[n] (periodic Hann): any other window must fail at plan creation."""
    with pytest.raises(RuntimeError, match="unsupported"):
        A.LogMelPlan(window=torch.hamming_window(2048), device="cuda:0")
    A.LogMelPlan(window=torch.hann_window(2048), device="cuda:0").close()


def test_seeded_training_batches_of_32_and_64_in_one_launch(A, golden):
    """BASELINE configs[3] at its stated size: batch 32 (R/config_segmented.yaml:21) and 64 (README) of 3 s clips with
    the reference's seeded noise / roll / SpecAugment choices, each as ONE kernel launch, against values the reference
    classes produced for the same 64 clips (tests/golden/make_golden.py, aug64_3s) and against the float64 oracle."""
    cfg = O.OracleConfig(duration=3.0)
    plan = get_plan(A, 2048, 512, cfg.target_length)
    trace, pos = golden["aug64_3s/trace"], golden["aug64_3s/pos"]
    draws = O.replay_augmentation(np.random.RandomState(42), O.TorchCpuGenerator(42), 64, cfg.target_length,
                                  cfg.n_mels, cfg.frames, want_noise_values=True)
    aug = A.make_aug_array(64)
    noise = np.zeros((64, cfg.target_length), dtype=np.float32)
    clips = []
    for c, d in enumerate(draws):
        aug[c]["shift"] = d.shift
        aug[c]["noise_scale"] = 0.005 if d.noise else 0.0
        aug[c]["f0"], aug[c]["f1"], aug[c]["t0"], aug[c]["t1"] = d.f0, d.f1, d.t0, d.t1
        if d.noise:
            noise[c] = d.noise_values
        clips.append(golden_input(100 + c, cfg.target_length))
    launches0 = plan.launches
    got64 = run_clips(plan, clips, aug=aug, noise=noise)
    got32 = run_clips(plan, clips[:32], aug=aug[:32], noise=noise[:32])
    assert plan.launches - launches0 == 2          # one launch per batch
    np.testing.assert_array_equal(got32["out"], got64["out"][:32])   # a clip does not depend on its batch
    for c, d in enumerate(draws):
        flat = got64["out"][c].reshape(-1)
        assert np.abs(flat[pos] - golden["aug64_3s/samples"][c]).max() < NORM_ATOL
        assert abs(np.abs(flat).mean() - golden["aug64_3s/absmean"][c]) < 1e-4
        z = got64["db"][c] == 0.0
        rows, cols = np.nonzero(z.all(axis=1))[0], np.nonzero(z.all(axis=0))[0]
        f0, f1 = (rows[0], rows[-1] + 1) if len(rows) else (0, 0)
        t0, t1 = (cols[0], cols[-1] + 1) if len(cols) else (0, 0)
        assert [f0, f1, t0, t1] == list(trace[c][2:])                 # mask indices bit-exact
    for c in (5, 17, 40, 63):                                          # full feature maps against the oracle
        d = draws[c]
        ref = O.logmel(clips[c], cfg, shift=d.shift, noise=d.noise_values, noise_scale=0.005 if d.noise else 0.0,
                       masks=(d.f0, d.f1, d.t0, d.t1))
        assert np.abs(got64["out"][c] - ref).max() < NORM_ATOL
