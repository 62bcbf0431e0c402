"""Why does the host-buffer (e2e) rate stop scaling past one GPU?  

Run under torchrun with N ranks on one box (N = 1, 2, 4, 8): every rank copies its own pinned host buffer to its own
GPU and back, all ranks at the same time (barrier before every timed region), with plain pinned and with
write-combined pinned source buffers, several chunk sizes, H2D only / D2H only / both directions at once.
Rank 0 prints one JSON line per configuration (per-rank min / mean GB/s and the aggregate) and, first, the host
topology (lscpu, numactl -H, nvidia-smi topo -m).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_scaling_probe.py
"""
import ctypes
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return f"({e})"


def pinned(nbytes, write_combined):
    """A pinned host tensor; write-combined memory comes straight from cudaHostAlloc (torch has no flag for it)."""
    if not write_combined:
        return torch.empty(nbytes, dtype=torch.uint8).pin_memory(), None
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    err = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(4))   # cudaHostAllocWriteCombined
    if err != 0:
        raise RuntimeError(f"cudaHostAlloc failed: {err}")
    return p.value, (rt, p)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        print(json.dumps({"world": world, "lscpu": sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread|Core'"),
                          "numactl": sh("numactl -H 2>/dev/null | head -12"), "topo": sh("nvidia-smi topo -m | head -14"),
                          "affinity": len(os.sched_getaffinity(0))}), flush=True)
    rt = ctypes.CDLL("libcudart.so.12")
    total = 512 << 20
    d_in = torch.empty(total, dtype=torch.uint8, device=dev)
    d_out = torch.empty(total, dtype=torch.uint8, device=dev)
    h_plain, _ = pinned(total, False)
    h_back = torch.empty(total, dtype=torch.uint8).pin_memory()
    try:
        h_wc, keep = pinned(total, True)
    except Exception as e:
        h_wc, keep = None, None
        if rank == 0:
            print(json.dumps({"write_combined": f"unavailable: {e}"}), flush=True)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def copy_async(dst_ptr, src_ptr, n, kind, stream):
        err = rt.cudaMemcpyAsync(ctypes.c_void_p(dst_ptr), ctypes.c_void_p(src_ptr), ctypes.c_size_t(n), ctypes.c_int(kind),
                                 ctypes.c_void_p(stream.cuda_stream))
        if err != 0:
            raise RuntimeError(f"cudaMemcpyAsync: {err}")

    for src_name, src_ptr in (("pinned", h_plain.data_ptr()), ("write_combined", h_wc)):
        if src_ptr is None:
            continue
        for chunk_mb in (4, 16, 64, 256):
            chunk = chunk_mb << 20
            n_chunks = total // chunk
            for mode in ("h2d", "d2h", "both"):
                def run():
                    for c in range(n_chunks):
                        o = c * chunk
                        if mode in ("h2d", "both"):
                            copy_async(d_in.data_ptr() + o, src_ptr + o, chunk, 1, s_in)
                        if mode in ("d2h", "both"):
                            copy_async(h_back.data_ptr() + o, d_out.data_ptr() + o, chunk, 2, s_out)
                run()
                barrier()
                t0 = time.perf_counter()
                for _ in range(3):
                    run()
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / 3
                gbs = total / dt / 1e9       # per direction
                t = torch.tensor([gbs], device=dev, dtype=torch.float64)
                if world > 1:
                    lo, su = t.clone(), t.clone()
                    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                    dist.all_reduce(su, op=dist.ReduceOp.SUM)
                else:
                    lo, su = t, t
                if rank == 0:
                    print(json.dumps({"world": world, "src": src_name, "chunk_mb": chunk_mb, "mode": mode,
                                      "per_rank_min_gbs": round(float(lo.item()), 2), "per_rank_mean_gbs": round(float(su.item()) / world, 2),
                                      "aggregate_gbs_per_direction": round(float(su.item()), 2)}), flush=True)
                barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
