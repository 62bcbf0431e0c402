"""Multi-GPU layout of the log-mel path: clips shard by index, one process per GPU.

Clips (and analyzer windows) are independent units; constants are replicated.  Rank r of R owns
the contiguous block [r*ceil(B/R), min(B, (r+1)*ceil(B/R))).  There is no data-path collective
when the consumer is itself data parallel; when one rank needs the whole batch the features are
all-gathered (`torch.distributed.all_gather_into_tensor`, NCCL over NVLink on the GPU box; gloo in
the CPU tests).  The last shard is padded with empty clips so every rank contributes the same
number of rows; the padding is trimmed after the gather.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_size", "ShardedLogMel"]


def shard_size(n_items: int, world: int) -> int:
    return (n_items + world - 1) // world


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's block; empty for trailing ranks when n_items < world * shard."""
    per = shard_size(n_items, world)
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


class ShardedLogMel:
    """Runs `extract(clips) -> [b, 1, n_mels, frames]` on this rank's shard and optionally gathers.

    `extract` is the per-rank feature extractor (e.g. `AudioPreprocessor.preprocess_batch`); it
    is injected so the index/gather logic can be exercised on CPU with gloo."""

    def __init__(self, extract: Callable[[Sequence], torch.Tensor], feature_shape: Tuple[int, int, int],
                 rank: Optional[int] = None, world: Optional[int] = None):
        self.extract = extract
        self.feature_shape = tuple(feature_shape)      # (1, n_mels, frames)
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world

    def local(self, clips: Sequence) -> Tuple[torch.Tensor, Tuple[int, int]]:
        """Features of this rank's block of `clips` (the full, identically ordered list)."""
        lo, hi = shard_bounds(len(clips), self.rank, self.world)
        feats = self.extract(clips[lo:hi]) if hi > lo else None
        return feats, (lo, hi)

    def gathered(self, clips: Sequence, device=None) -> torch.Tensor:
        """All ranks' features, `[len(clips), 1, n_mels, frames]`, identical on every rank."""
        n = len(clips)
        per = shard_size(n, self.world)
        feats, (lo, hi) = self.local(clips)
        if feats is None:
            if device is None:
                raise ValueError("device is required on a rank with an empty shard")
            feats = torch.zeros((0,) + self.feature_shape, dtype=torch.float32, device=device)
        block = torch.zeros((per,) + self.feature_shape, dtype=torch.float32, device=feats.device)
        block[:hi - lo] = feats
        full = torch.empty((per * self.world,) + self.feature_shape, dtype=torch.float32, device=feats.device)
        if self.world > 1:
            dist.all_gather_into_tensor(full, block)
        else:
            full.copy_(block)
        return full[:n]
