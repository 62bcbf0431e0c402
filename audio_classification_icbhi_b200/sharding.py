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

__all__ = ["shard_bounds", "shard_size", "ShardedLogMel", "FusedGather"]


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


class FusedGather:
    """Feature all-gather fused into the log-mel kernel (`lm_forward_gather`).

    Every rank allocates the gathered buffer `[world * per_rank, 1, n_mels, frames]` in symmetric memory
    (`torch.distributed._symmetric_memory`: the same allocation on every GPU of the NVSwitch domain, mapped
    into every process).  `run` launches the kernel once; its clip-end normalisation pass stores this
    rank's features into its slice of EVERY rank's buffer -- one `multimem.st` per 16 bytes when the group
    has a multicast address (the switch replicates it), plain stores to the peers' mapped buffers
    otherwise -- so the transfer overlaps the compute of the remaining clips.  `finish` is the rank
    barrier after which `self.full` holds all ranks' features.  NCCL (`ShardedLogMel.gathered`) remains
    the path for ranks without peer access.

    Ordering: `run` also STARTS with a rank barrier on the current stream.  The kernel writes straight into the
    other ranks' `full` buffers, so a rank must not begin step k+1 while a slower rank's consumer kernels (enqueued
    on that rank's stream after `finish`) still read step k: the entry barrier orders every rank's earlier stream
    work before anybody's stores.  `self.full` is therefore valid from `finish()` until this rank's next `run()`."""

    def __init__(self, plan, per_rank: int, group=None, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm
        self.plan = plan
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world - 1 > 7:
            raise ValueError("lm_forward_gather addresses at most 7 peers")
        self.per_rank = int(per_rank)
        self.clip_elems = plan.n_mels * plan.frames
        n = self.world * self.per_rank * self.clip_elems
        self._flat = symm.empty(n, dtype=torch.float32, device=plan.device)
        self._hdl = symm.rendezvous(self._flat, self.group)
        self.full = self._flat.view(self.world * self.per_rank, 1, plan.n_mels, plan.frames)
        off = self.rank * self.per_rank * self.clip_elems * 4
        ptrs = [int(x) for x in self._hdl.buffer_ptrs]
        self.out_slice_ptr = ptrs[self.rank] + off
        self.peer_slice_ptrs = [ptrs[r] + off for r in range(self.world) if r != self.rank]
        mc = int(self._hdl.multicast_ptr) if (use_multicast and self._hdl.has_multicast_support) else 0
        self.mc_slice_ptr = mc + off if mc else 0

    @property
    def mode(self) -> str:
        return "multimem.st (NVSwitch multicast)" if self.mc_slice_ptr else "st.global to peer-mapped buffers"

    def run(self, wave: torch.Tensor, offset: torch.Tensor, length: torch.Tensor, **kw) -> None:
        if int(offset.numel()) > self.per_rank:
            raise ValueError("more clips than the rank's slice holds")
        self._hdl.barrier()   # write-after-read: peers may still be reading the previous step's features
        self.plan.forward_gather(wave, offset, length, self.out_slice_ptr, self.peer_slice_ptrs, self.mc_slice_ptr, **kw)

    def finish(self) -> torch.Tensor:
        """Rank barrier on the current stream (symmetric-memory signal pads): afterwards every slice is complete."""
        self._hdl.barrier()
        return self.full
