"""Debug: phase timeline of the cluster KD kernel (needs a build with LICV_EXTRA_NVCC_FLAGS=-DLICV_TRACE).

The trace points live in kd_loss_cluster_kernel, so the tensor-memory kernel (the default for
16-bit logits) is switched off for this process."""
import ctypes
import os
import sys

os.environ["LICV_KD_TMEM"] = "0"

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import _abi  # noqa: E402

lib = _abi.load()
R, V = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 32002
ce_only = len(sys.argv) > 2 and sys.argv[2] == "ce"
dt, code = torch.bfloat16, _abi.BF16
stu = (torch.randn(R, V, device="cuda") * 3).to(dt)
tea = (torch.randn(R, V, device="cuda") * 3).to(dt)
dst = torch.empty_like(stu)
lab = torch.randint(0, V, (R,), device="cuda")
ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8, device="cuda")
losses = torch.zeros(4, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    if ce_only:
        lib.licv_kd_loss_fwd_bwd(stu.data_ptr(), dst.data_ptr(), 0, 0, lab.data_ptr(), 0, 0, R, 1.0,
                                 1e-6, 0.5, 1, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, st)
    else:
        lib.licv_kd_loss_fwd_bwd(stu.data_ptr(), dst.data_ptr(), tea.data_ptr(), 0, lab.data_ptr(), 0, R, R,
                                 1.0, 1e-6, 0.5, 0, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, st)
    torch.cuda.synchronize()
raw = ctypes.CDLL(_abi.LIB_PATH if hasattr(_abi, "LIB_PATH") else lib._name)
n = 64 * 64 * 8
buf = (ctypes.c_longlong * n)()
raw.licv_debug_read_trace(buf, n)
t = np.array(buf, dtype=np.int64).reshape(64, 64, 8)
names = ["B", "arrive1+prefetch", "wait1", "C", "wait2", "D"]
for cta in (0, 1, 2, 3, 4, 17):
    print(f"CTA {cta}: rows x phases (cycles)")
    for row in range(2, 8):
        ts = t[cta, row]
        if ts[0] == 0:
            continue
        if ce_only:
            seg = [ts[1] - ts[0], ts[2] - ts[1], ts[3] - ts[2], ts[6] - ts[3]]
            print("   row", row, "B %6d  arr+pf %6d  wait1 %6d  D %6d | next-start gap %6d" %
                  (*seg, t[cta, row + 1, 0] - ts[6]))
        else:
            seg = [ts[i + 1] - ts[i] for i in range(6)]
            print("   row", row, " ".join(f"{nm} {v:6d}" for nm, v in zip(names, seg)),
                  "| total", t[cta, row + 1, 0] - ts[0])
