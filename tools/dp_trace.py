"""Debug: phase timeline of the fused gradient exchange (csrc/licv_dp.cu) under torchrun.

    LICV_EXTRA_NVCC_FLAGS=-DLICV_TRACE python -c "import __graft_entry__ as g; g.build()"
    python -m torch.distributed.run --nproc-per-node N tools/dp_trace.py

Replays bench.py's hot-path step as a CUDA graph and prints, per rank, the median time CTA 0 of
the exchange kernel spends between its trace points (globaltimer, ns):
entry -> dependency resolved -> packets pushed to every peer -> every peer's packet seen ->
sum written back -> CTA's norm partial ready."""
import ctypes
import os
import sys

# the trace points live in the all-to-all form of the kernel (the owner form is the default)
os.environ.setdefault("LICV_DP_ALGO", "all")

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from licv_vqa_b200 import _abi  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.distributed.init_process_group("nccl", device_id=dev)
    hp = bench.HotPath(dev, torch.float16, world)
    batch = bench.make_batch(dev, 1000 + rank, torch.float16)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            hp.step(batch)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            hp.step(batch)
        torch.distributed.barrier()
        for _ in range(200):
            g.replay()
        s.synchronize()
    raw = ctypes.CDLL(_abi.load()._name)
    n = 256 * 8
    buf = (ctypes.c_longlong * n)()
    raw.licv_debug_read_dp_trace(buf, n)
    t = np.array(buf, dtype=np.int64).reshape(256, 8)
    t = t[(t[:, 0] > 0) & (t[:, 5] > t[:, 0])]
    d = np.diff(t[:, :6], axis=1)
    names = ["pdl wait", "push", "peers' packets seen", "sum written", "warp sum"]
    med = np.median(d, axis=0)
    p90 = np.percentile(d, 90, axis=0)
    line = f"rank {rank}/{world}: " + "  ".join(f"{nm} {m/1e3:.1f} (p90 {q/1e3:.1f})" for nm, m, q in zip(names, med, p90))
    line += f"  | total {np.median(t[:,5]-t[:,0])/1e3:.1f} us, after dependency {np.median(t[:,5]-t[:,1])/1e3:.1f} us"
    for r in range(world):
        if r == rank:
            print(line, flush=True)
        torch.distributed.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
