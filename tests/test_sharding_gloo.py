"""The N > 1 path on CPU: two gloo ranks shard a ragged clip list by index, run a stand-in
extractor on their block and all-gather; the result must equal the single-process run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audio_classification_icbhi_b200.sharding import ShardedLogMel, shard_bounds

SHAPE = (1, 8, 5)


def fake_extract(clips):
    """Deterministic per-clip 'features' that depend on the clip content only."""
    out = torch.zeros((len(clips),) + SHAPE)
    for i, c in enumerate(clips):
        c = torch.as_tensor(c, dtype=torch.float32)
        s = float(c.sum()) if c.numel() else 0.0
        out[i] = torch.arange(40, dtype=torch.float32).reshape(SHAPE) * 0.01 + s + c.numel()
    return out


def make_clips(n):
    rs = np.random.RandomState(5)
    return [rs.standard_normal(int(k)).astype(np.float32) for k in rs.randint(0, 50, n)]


def worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clips = make_clips(n)
        sh = ShardedLogMel(fake_extract, SHAPE)
        local, (lo, hi) = sh.local(clips)
        assert (lo, hi) == shard_bounds(n, rank, world)
        full = sh.gathered(clips, device="cpu")
        q.put((rank, full.numpy(), None if local is None else local.shape[0]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 8, 1])
def test_two_rank_gather_equals_single_process(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = fake_extract(make_clips(n)).numpy()
    for rank, full, n_local in results:
        assert full.shape == ref.shape
        np.testing.assert_array_equal(full, ref)
    assert sum(r[2] or 0 for r in results) == n
