"""Debug: phase timeline of the stream KD kernel (needs a build with
LICV_EXTRA_NVCC_FLAGS=-DLICV_TRACE).   python tools/kd_stream_trace.py [rows] [ce]"""
import ctypes
import os
import sys

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
n = 8 * 3 * 64 * 8
buf = (ctypes.c_longlong * n)()
lib.licv_debug_read_stream_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.licv_debug_read_stream_trace(buf, n)
t = np.array(buf, dtype=np.int64).reshape(8, 3, 64, 8)
r0 = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for cta in (0, 3):
    print(f"CTA {cta}")
    for row in range(r0, r0 + 7):
        for who, nm in ((0, "warp0 "), (1, "warp15")):
            ts = t[cta, who, row]
            nxt = t[cta, who, row + 1, 0]
            print(f"  it {row} {nm}: red1 {ts[1]-ts[0]:6d}  C+red2 {ts[2]-ts[1]:6d}  sweepDB {ts[3]-ts[2]:6d}"
                  f" (full-wait {ts[4]:6d})  total {nxt-ts[0]:6d}")
        ps = t[cta, 2, row]
        print(f"  row {row} producer: label {ps[1]-ps[0]:6d}  issue {ps[2]-ps[1]:6d} (empty-wait {ps[3]:6d})"
              f"  start rel. consumer it {ps[0]-t[cta,0,row,0]:8d}")
