"""What the host link gives: pinned H2D / D2H alone and both directions at once, by transfer size.

    python tools/pcie_probe.py                                   one GPU
    torchrun --nproc-per-node 8 tools/pcie_probe.py              all GPUs at once (gloo barrier):
                                                                 the per-GPU rate when every rank's
                                                                 copies share the host's memory
"""
import os

import torch

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.distributed.init_process_group("gloo")

for mb in (16, 128):
    n = mb << 20
    a = torch.empty(n, dtype=torch.uint8).pin_memory()
    b = torch.empty(n, dtype=torch.uint8).pin_memory()
    da = torch.empty(n, dtype=torch.uint8, device="cuda")
    db = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    reps = max(8, 2048 // mb)

    def run(h2d, d2h):
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    da.copy_(a, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    b.copy_(db, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    run(True, True)
    t1, t2, t3 = run(True, False), run(False, True), run(True, True)
    gb = n * reps / 1e9
    print(f"rank {rank}/{world} {mb:4d} MiB x{reps}: H2D {gb / t1:6.1f} GB/s  D2H {gb / t2:6.1f} GB/s  "
          f"both {gb / t3:6.1f} GB/s each way", flush=True)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
