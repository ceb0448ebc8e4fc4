"""e2e leg of bench.py alone (session depth from LICV_E2E_SLOTS, grid cap from LICV_HOST_GRID_CAP)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
hp = bench.HotPath(dev, torch.float16, 1)
dt, h2d, d2h = bench.run_e2e_host(hp, 20, 3, 1000)
print(f"slots={os.environ.get('LICV_E2E_SLOTS','8')} cap={os.environ.get('LICV_HOST_GRID_CAP','32')}: {dt*1e3:.3f} ms/step  {8/dt:.0f} samples/s  H2D {h2d/dt/1e9:.1f} GB/s D2H {d2h/dt/1e9:.1f} GB/s", flush=True)
