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
    dist.barrier()
    if rank == 0:
        print("FUSED_GATHER_OK", world, "ranks;", " | ".join(modes))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
