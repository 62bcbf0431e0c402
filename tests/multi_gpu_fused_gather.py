"""Launched by tests/test_gpu_multi.py under torchrun with one rank per GPU: the fused feature all-gather
(lm_forward_gather: peer stores and NVSwitch multicast stores from the kernel's normalisation pass) must
reproduce NCCL's all_gather_into_tensor of the per-rank features bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> None:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from audio_classification_icbhi_b200 import FusedGather, LogMelPlan

    plan = LogMelPlan(target_length=48000, device=dev)      # config_segmented.yaml: 3 s clips, 94 frames
    B = 333                                                 # not a multiple of the persistent grid
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    lens = torch.randint(1000, 60000, (B,), generator=g, device=dev, dtype=torch.int64)
    starts = torch.cumsum((lens + 7) // 8 * 8, 0) - (lens + 7) // 8 * 8
    wave = torch.randn(int(starts[-1] + lens[-1]) + 8, generator=g, device=dev) * 0.1
    offset, length = starts.contiguous(), lens.to(torch.int32)
    own = plan.forward(wave, offset, length)
    ref = torch.empty((world * B, 1, plan.n_mels, plan.frames), device=dev)
    dist.all_gather_into_tensor(ref, own)
    modes = []
    for use_mc in (True, False):
        fg = FusedGather(plan, B, use_multicast=use_mc)
        fg.full.fill_(float("nan"))
        torch.cuda.synchronize()
        dist.barrier()
        fg.run(wave, offset, length)
        full = fg.finish()
        torch.cuda.synchronize()
        assert torch.equal(full, ref), f"rank {rank}: fused gather ({fg.mode}) differs from the NCCL gather"
        modes.append(fg.mode)
    # ---- BASELINE configs[4]: 1-hour recording, 1 s windows at 50 % overlap -> 7200 windows, sharded by index;
    #      the recording is replicated, every rank extracts its block of windows and the blocks are gathered
    #      in-kernel.  Rank 0 also extracts all 7200 windows alone: the gathered result must equal it bit for bit.
    from audio_classification_icbhi_b200 import segment_offsets, shard_bounds, shard_size
    plan1 = LogMelPlan(target_length=16000, device=dev)
    g2 = torch.Generator(device=dev).manual_seed(2)
    rec = torch.randn(3600 * 16000, generator=g2, device=dev) * 0.1
    starts_np, lens_np, times = segment_offsets(int(rec.numel()), 16000, 1.0, 0.5)
    assert len(starts_np) == 7200 and times[-1] == (3599.5, 3600.0)
    per = shard_size(len(starts_np), world)
    lo, hi = shard_bounds(len(starts_np), rank, world)
    fgw = FusedGather(plan1, per)
    fgw.full.fill_(float("nan"))
    torch.cuda.synchronize()
    dist.barrier()
    fgw.run(rec, torch.from_numpy(starts_np[lo:hi]).to(dev), torch.from_numpy(lens_np[lo:hi]).to(dev))
    allw = fgw.finish()
    torch.cuda.synchronize()
    if rank == 0:
        alone = plan1.forward(rec, torch.from_numpy(starts_np).to(dev), torch.from_numpy(lens_np).to(dev))
        torch.cuda.synchronize()
        for r in range(world):          # shards sit at r * per; the last one may be short
            a, b = shard_bounds(len(starts_np), r, world)
            assert torch.equal(allw[r * per:r * per + (b - a)], alone[a:b]), f"analyzer windows of rank {r} differ"
    dist.barrier()
    if rank == 0:
        print("FUSED_GATHER_OK", world, "ranks;", " | ".join(modes), "; 7200 analyzer windows sharded and gathered")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
