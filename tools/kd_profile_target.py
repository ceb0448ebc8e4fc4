"""Fixed workload for `ncu --set full` on the loss kernel: 2048 x 32002 bf16 rows, KL+CE then CE-only.
    python tools/kd_profile_target.py [rows]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import _abi  # noqa: E402

lib = _abi.load()
R, V = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 32002
dt, code = torch.bfloat16, _abi.BF16
st = torch.cuda.current_stream().cuda_stream
stu = (torch.randn(R, V, device="cuda") * 3).to(dt)
tea = (torch.randn(R, V, device="cuda") * 3).to(dt)
dst = torch.empty_like(stu)
lab = torch.randint(0, V, (R,), device="cuda")
ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8, device="cuda")
losses = torch.zeros(4, device="cuda")
for _ in range(2):
    _abi.check(lib.licv_kd_loss_fwd_bwd(stu.data_ptr(), dst.data_ptr(), tea.data_ptr(), 0, lab.data_ptr(), 0, R, R,
                                        1.0, 1e-6, 0.5, 0, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, st))
    _abi.check(lib.licv_kd_loss_fwd_bwd(stu.data_ptr(), dst.data_ptr(), 0, 0, lab.data_ptr(), 0, 0, R, 1.0,
                                        1e-6, 0.5, 1, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, st))
torch.cuda.synchronize()
print("done", losses.tolist())
