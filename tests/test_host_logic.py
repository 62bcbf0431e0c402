"""Host-side logic of the drop-in layer, CPU only: RNG replay, window tables, datasets, segmenter,
wav I/O, sharding arithmetic.  No CUDA compute is called."""
import json
import os
import random

import numpy as np
import pytest
import torch

import audio_classification_icbhi_b200 as A
from audio_classification_icbhi_b200 import wavio
from audio_classification_icbhi_b200.augment import draw_fast_augmentation, draw_reference_augmentation
from oracle import logmel_oracle as O


def set_seed(seed):   # R/src/utils/config.py:31-33
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


@pytest.mark.parametrize("tag,dur", [("aug_3s", 3.0), ("aug_5s", 5.0)])
def test_reference_augmentation_replay_is_bit_exact(golden, tag, dur):
    """After set_seed(42) the host replay makes the reference's choices (tests/golden trace)."""
    cfg = O.OracleConfig(duration=dur)
    trace = golden[f"{tag}/trace"]
    set_seed(42)
    aug, noise = draw_reference_augmentation(len(trace), cfg.target_length, cfg.n_mels, cfg.frames)
    canon = lambda a, b: (int(a), int(b)) if b > a else (0, 0)
    got = [[int(a["noise_scale"] > 0), int(a["shift"]), *canon(a["f0"], a["f1"]), *canon(a["t0"], a["t1"])] for a in aug]
    np.testing.assert_array_equal(got, trace)
    assert np.all(aug["gain"] == 1.0)
    for c in range(len(trace)):
        if trace[c][0]:
            # golden noise was recovered as (noisy - clean) / 0.005: a few ulp of rounding
            np.testing.assert_allclose(noise[c, :32].numpy(), golden[f"{tag}/clip{c}/noise_head"], atol=1e-5)
        elif noise is not None:
            assert not noise[c].any()
    # and it agrees with the oracle's own restatement of both generators
    draws = O.replay_augmentation(np.random.RandomState(42), O.TorchCpuGenerator(42), len(trace),
                                  cfg.target_length, cfg.n_mels, cfg.frames)
    for a, d in zip(aug, draws):
        assert (int(a["shift"]), int(a["f0"]), int(a["f1"]), int(a["t0"]), int(a["t1"])) == (d.shift, d.f0, d.f1, d.t0, d.t1)


def test_drop_in_helper_methods_consume_the_same_streams():
    """augment_waveform / augment_spectrogram called one by one (the reference's own call pattern)
    draw from the same global generators as the batched replay."""
    p = A.AudioPreprocessor(duration=3.0, augment=True)
    w = torch.from_numpy((np.random.RandomState(0).standard_normal(48000) * 0.1).astype(np.float32)).unsqueeze(0)
    set_seed(42)
    w1 = p.augment_waveform(w)
    m = p.augment_spectrogram(torch.ones(1, 128, 94))
    set_seed(42)
    aug, noise = draw_reference_augmentation(1, 48000, 128, 94)
    expect = w + (noise[0] * 0.005 if noise is not None else 0)
    expect = torch.roll(expect, int(aug[0]["shift"]), dims=-1)
    assert torch.equal(w1, expect)
    z = m[0] == 0
    rows, cols = torch.nonzero(z.all(dim=1)).flatten().tolist(), torch.nonzero(z.all(dim=0)).flatten().tolist()
    assert rows == list(range(int(aug[0]["f0"]), int(aug[0]["f1"])))
    assert cols == list(range(int(aug[0]["t0"]), int(aug[0]["t1"])))


def test_fast_augmentation_respects_the_reference_bounds():
    aug = draw_fast_augmentation(5000, 48000, 128, 94, rng=np.random.default_rng(1), gain_db=6.0)
    assert np.all((aug["f0"] >= 0) & (aug["f0"] <= aug["f1"]) & (aug["f1"] <= 128) & (aug["f1"] - aug["f0"] < 15))
    assert np.all((aug["t0"] >= 0) & (aug["t0"] <= aug["t1"]) & (aug["t1"] <= 94) & (aug["t1"] - aug["t0"] < 35))
    assert np.all(np.abs(aug["shift"]) < 0.2 * 48000 + 1)
    assert 0.4 < (aug["noise_scale"] > 0).mean() < 0.6 and 0.4 < (aug["shift"] != 0).mean() < 0.6
    assert np.all((aug["gain"] > 10 ** (-6.01 / 20)) & (aug["gain"] < 10 ** (6.01 / 20)))
    assert len(set(aug["seed"].tolist())) == 5000


def test_segment_offsets_match_oracle_and_known_answers():
    for n, seg, ov in [(15 * 16000, 0.5, 0.75), (15 * 16000, 1.0, 0.5), (3600 * 16000, 1.0, 0.5), (3600 * 16000, 5.0, 0.5),
                       (100, 1.0, 0.5), (0, 1.0, 0.5), (16000, 1.0, 0.5), (16001, 1.0, 0.0), (40000, 0.7, 0.33)]:
        starts, lengths, times = A.segment_offsets(n, 16000, seg, ov)
        ref = O.segment_offsets(n, 16000, seg, ov)
        assert [(int(s), int(l)) for s, l in zip(starts, lengths)] == [(r[0], r[1]) for r in ref]
        assert times == [(r[2], r[3]) for r in ref]
    starts, lengths, times = A.segment_offsets(15 * 16000, 16000, 0.5, 0.75)
    assert len(starts) == 118 and times[-1] == (14.625, 15.0)      # R/analysis_results/test_audio_1_results.csv
    assert len(A.segment_offsets(3600 * 16000, 16000, 1.0, 0.5)[0]) == 7200


def test_flexible_geometry_table():
    """SURVEY.md section 8 shapes table."""
    rows = [(0.5, 1024, 256, 32, 32), (1.0, 2048, 512, 32, 32), (3.0, 2048, 512, 94, 94),
            (5.0, 2048, 512, 157, 157), (8.0, 2048, 512, 251, 250)]
    for dur, n_fft, hop, stft_frames, out_frames in rows:
        f = A.FlexibleAudioPreprocessor(duration=dur)
        assert (f.n_fft, f.hop_length, f.stft_frames, f.frames) == (n_fft, hop, stft_frames, out_frames)
        assert f.target_length == int(16000 * dur)
    p = A.AudioPreprocessor(duration=8.0)
    assert p.frames == 251 and p.target_length == 128000


def test_pad_or_crop_semantics():
    p = A.AudioPreprocessor(duration=0.001 * 10)   # target 160 samples
    x = torch.arange(100, dtype=torch.float32).unsqueeze(0)
    y = p.pad_or_crop(x)
    assert y.shape == (1, 160) and torch.equal(y[0, :100], x[0]) and not y[0, 100:].any()
    x = torch.arange(201, dtype=torch.float32).unsqueeze(0)
    assert torch.equal(p.pad_or_crop(x)[0], x[0, 20:180])        # centre crop, start = (201-160)//2
    assert torch.equal(p.pad_or_crop(p.pad_or_crop(x)), p.pad_or_crop(x))
    np.testing.assert_array_equal(O.pad_or_crop(x.numpy(), 160), p.pad_or_crop(x).numpy())


def test_wav_round_trip(tmp_path):
    rs = np.random.RandomState(0)
    x = (rs.standard_normal(4000) * 0.3).clip(-1, 1).astype(np.float32)
    path = str(tmp_path / "a.wav")
    wavio.write_wav_pcm16(path, x, 16000)
    y, sr = wavio.read_wav(path)
    assert sr == 16000 and y.shape == (1, 4000)
    np.testing.assert_array_equal(y[0], np.rint(x.astype(np.float64) * 32767).astype(np.float32) / 32768.0)
    p = A.AudioPreprocessor()
    w = p.load_audio(path)
    assert w.shape == (1, 4000) and w.dtype == torch.float32


def make_icbhi_tree(root, n=10):
    d = root / "audio_and_txt_files"
    d.mkdir(parents=True)
    rs = np.random.RandomState(1)
    for i in range(n):
        wavio.write_wav_pcm16(str(d / f"{100 + i}_rec.wav"), rs.standard_normal(16000 * 3) * 0.1, 16000)
        rows = [(0.0, 1.2, i % 2, (i // 2) % 2), (1.2, 1.5, 0, 0), (1.5, 2.9, 0, 0)]
        (d / f"{100 + i}_rec.txt").write_text("".join(f"{a}\t{b}\t{c}\t{w}\n" for a, b, c, w in rows))
    return d


def test_icbhi_dataset_discovery_and_splits(tmp_path):
    make_icbhi_tree(tmp_path, 10)
    cfg = {"data": dict(sample_rate=16000, n_mels=128, n_fft=2048, hop_length=512, duration=3.0)}
    tr = A.ICBHIDataset(tmp_path, "train", cfg, augment=True)
    va = A.ICBHIDataset(tmp_path, "val", cfg, augment=True)
    te = A.ICBHIDataset(tmp_path, "test", cfg)
    assert (len(tr), len(va), len(te)) == (7, 1, 2)                  # int(0.7*10), int(0.15*10), rest
    assert tr.augment and tr.preprocessor.augment and not va.augment  # augment only on the train split
    assert tr.preprocessor.target_length == 48000
    assert [l for _, l in tr.data] == [0, 1, 2, 3, 0, 1, 2]          # union of cycle flags per recording
    assert A.ICBHIDataset.CLASS_MAP == {"normal": 0, "crackles": 1, "wheezes": 2, "both": 3}
    w, sr, y = tr.raw_item(1)
    assert w.shape == (1, 48000) and sr == 16000 and y == 1
    assert len(tr.raw()) == 7
    with pytest.raises(ValueError, match="Audio directory not found"):
        A.ICBHIDataset(tmp_path / "nope", "train")


def test_segmenter_and_segmented_dataset(tmp_path):
    src = make_icbhi_tree(tmp_path / "raw", 6)
    out = tmp_path / "seg"
    seg = A.ICBHISegmenter(src, out, sample_rate=16000, min_duration=0.5)
    assert seg.parse_annotation(src / "100_rec.txt") == [(0.0, 1.2, 0, 0), (1.2, 1.5, 0, 0), (1.5, 2.9, 0, 0)]
    assert [seg.get_label(c, w) for c, w in ((0, 0), (1, 0), (0, 1), (1, 1))] == ["normal", "crackle", "wheeze", "both"]
    seg.process_all()
    assert seg.stats["total_files"] == 6 and seg.stats["skipped_segments"] == 6   # the 0.3 s cycles
    assert seg.stats["total_segments"] == 12
    assert json.loads((out / "segmentation_stats.json").read_text()) == seg.stats
    y, sr = wavio.read_wav(str(next((out / "normal").glob("100_rec_seg000_normal.wav"))))
    assert sr == 16000 and y.shape[1] == int(1.2 * 16000)
    cfg = {"data": dict(sample_rate=16000, n_mels=128, n_fft=2048, hop_length=512, duration=3.0,
                        train_split=0.75, val_split=0.45)}
    tr = A.ICBHISegmentedDataset(out, "train", cfg, augment=True)
    va = A.ICBHISegmentedDataset(out, "val", cfg)
    te = A.ICBHISegmentedDataset(out, "test", cfg)
    assert (len(tr), len(va), len(te)) == (9, 3, 0)                  # int(.75*12), min(int(.45*12), rest)
    assert A.ICBHISegmentedDataset.CLASS_MAP == {"normal": 0, "crackle": 1, "wheeze": 2, "both": 3}
    again = A.ICBHISegmentedDataset(out, "train", cfg)
    assert [p for p, _ in again.data] == [p for p, _ in tr.data]     # random.seed(42) shuffle is reproducible
    with pytest.raises(ValueError, match="No audio files found"):
        A.ICBHISegmentedDataset(tmp_path / "empty", "train")


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 8, 6900, 4096, 7200):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = A.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo <= A.shard_size(n, world)
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
    assert A.shard_bounds(6900, 7, 8) == (6041, 6900) and A.shard_size(6900, 8) == 863


def test_raw_items_never_touch_cuda_even_when_the_file_needs_resampling(tmp_path):
    """raw mode is what forked DataLoader workers run, and raw ICBHI recordings are 4 / 10 / 44.1 kHz.
    raw_item must therefore only decode: it hands back the file's own rate and leaves resampling to GpuCollate in
    the process that owns the CUDA context.  This test runs where there is no GPU at all, through real forked
    workers and the identity collate."""
    import torch
    d = tmp_path / "audio_and_txt_files"
    d.mkdir()
    rs = np.random.RandomState(5)
    rates = [44100, 4000, 10000, 16000]
    for i, sr in enumerate(rates):
        wavio.write_wav_pcm16(str(d / f"{100 + i}_r.wav"), (rs.standard_normal(sr * 2) * 0.1).clip(-1, 1), sr)
        (d / f"{100 + i}_r.txt").write_text("0.0\t1.0\t0\t1\n")
    cfg = {"data": dict(sample_rate=16000, n_mels=128, n_fft=2048, hop_length=512, duration=3.0)}
    ds = A.ICBHIDataset(tmp_path, "train", cfg)            # 70 % of 4 = 2 recordings... take the raw view of all
    ds.data = [(str(d / f"{100 + i}_r.wav"), 2) for i in range(4)]
    w, sr, y = ds.raw_item(0)
    assert (tuple(w.shape), sr, y) == ((1, 88200), 44100, 2) and w.device.type == "cpu"
    assert ds.preprocessor._plan is None                    # no plan, no CUDA context was created
    loader = torch.utils.data.DataLoader(ds.raw(), batch_size=2, num_workers=2, collate_fn=A.raw_collate,
                                         multiprocessing_context="fork")
    got = [item for batch in loader for item in batch]
    assert [(int(w.shape[-1]), sr) for w, sr, _ in got] == [(2 * r, r) for r in rates]
