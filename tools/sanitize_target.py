"""Small workload for compute-sanitizer: every kernel family once, on shapes that exercise the
edge paths (row ends, misaligned teacher rows, partial token batches, idle cluster CTAs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import ops  # noqa: E402

torch.manual_seed(0)
for dt in (torch.bfloat16, torch.float32):
    for n_tok, d in ((1, 4096), (37, 4096), (300, 4096), (5, 512), (33, 520), (9, 8192), (7, 2048)):
        h = (torch.randn(n_tok, d, device="cuda") * 3).to(dt)
        g = torch.randn(n_tok, d, device="cuda").to(dt)
        s = torch.randn(d, device="cuda")
        ds = torch.zeros(d, device="cuda")
        ops.inject_forward(h, s, dt, 0)
        ops.inject_backward(h, g, s, ds, True, 0)
for dt, V, R, Rt in ((torch.bfloat16, 32002, 20, 9), (torch.float16, 32003, 11, 11), (torch.float32, 1003, 7, 7),
                     (torch.bfloat16, 50257, 5, 5), (torch.bfloat16, 33, 4, 4)):
    stu = (torch.randn(R, V, device="cuda") * 3).to(dt)
    tea = (torch.randn(Rt, V, device="cuda") * 3).to(dt)
    ktr = torch.full((R,), -1, dtype=torch.int32, device="cuda")
    ktr[::2] = torch.arange(0, (R + 1) // 2, dtype=torch.int32, device="cuda") % Rt
    lab = torch.randint(0, V, (R,), device="cuda")
    lab[1] = -100
    n_kl = int((ktr >= 0).sum())
    ops.kd_loss_raw(stu, tea, ktr, lab, None, n_kl, R - 1, 1.0, 1e-6, 0.5, in_place=False)
    ops.kd_loss_raw(stu, tea, ktr, lab, None, n_kl, R - 1, 1.0, 1e-6, 0.5, in_place=True)
    ops.kd_loss_raw(stu, None, None, lab, None, 0, R - 1, 2.0, 1e-6, 0.5, only_hard_loss=True, in_place=False)
ids = torch.randint(1, 50, (4, 12), device="cuda")
ops.kd_prepare_rows(ids, torch.tensor([3, 4, 5, 6], device="cuda"), torch.randint(1, 50, (4, 30), device="cuda"),
                    torch.tensor([21, 22, 23, 24], device="cuda"), 0)
p = torch.randn(1000 + 7, device="cuda")
ops.adamw_step(p, torch.randn_like(p), torch.zeros_like(p), torch.zeros_like(p), 1000, 7, 1e-3, 1e-2, 1)
torch.cuda.synchronize()
print("sanitize target done")
