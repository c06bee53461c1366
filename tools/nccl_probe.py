"""NCCL latency / bandwidth of this box as torch.distributed sees it (same libnccl the library loads): the sizes the
column-sharded path uses -- candidate all-gather (KBs), active-column all-reduce (~2 MB), screened-design all-reduce (40 MB)."""
import os

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    for name, nbytes, op in [("all_gather 4 KB", 4096, "ag"), ("all_reduce 1.76 MB", 1760000, "ar"),
                             ("all_reduce 40 MB", 40000000, "ar"), ("all_gather 40 MB total", 40000000 // world, "ag")]:
        x = torch.ones(nbytes // 8, dtype=torch.float64, device="cuda")
        out = torch.empty(world * x.numel(), dtype=torch.float64, device="cuda")
        for _ in range(5):
            dist.all_reduce(x) if op == "ar" else dist.all_gather_into_tensor(out, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 20
        for _ in range(reps):
            dist.all_reduce(x) if op == "ar" else dist.all_gather_into_tensor(out, x)
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            us = e0.elapsed_time(e1) / reps * 1e3
            print(f"{name}: {us:.1f} us  ({nbytes / us / 1e3:.1f} GB/s algorithmic)", flush=True)
    if rank == 0:
        print("can_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1) if torch.cuda.device_count() > 1 else None)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
