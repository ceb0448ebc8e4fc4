"""Small fixed workload for `ncu --set full`: each hot kernel a few times at a bandwidth-bound
shape (configs[4] sizes) and at the training shape (configs[1])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import _abi  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    lib = _abi.load()
    st = torch.cuda.current_stream().cuda_stream
    dt, code = torch.bfloat16, _abi.BF16
    d, V = 4096, int(os.environ.get("LICV_PROFILE_VOCAB", "32002"))
    s = torch.randn(d, device="cuda")
    ds = torch.zeros(d, device="cuda")
    toks = [int(x) for x in os.environ.get("LICV_PROFILE_TOKENS", "32768,256").split(",")]
    for n_tok in toks:
        h = (torch.randn(n_tok, d, device="cuda") * 4).to(dt)
        g = torch.randn(n_tok, d, device="cuda").to(dt)
        o = torch.empty_like(h)
        for _ in range(3):
            if which in ("all", "inject"):
                lib.licv_inject_fwd(h.data_ptr(), s.data_ptr(), o.data_ptr(), n_tok, d, code, code, 0, st)
                lib.licv_inject_bwd(h.data_ptr(), g.data_ptr(), s.data_ptr(), o.data_ptr(), ds.data_ptr(),
                                    n_tok, d, code, code, 0, st)
        torch.cuda.synchronize()
    if which in ("all", "kd"):
        for R in (2048, 256):
            stu = (torch.randn(R, V, device="cuda") * 3).to(dt)
            tea = (torch.randn(R, V, device="cuda") * 3).to(dt)
            dst = torch.empty_like(stu)
            lab = torch.randint(0, V, (R,), device="cuda")
            ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8, device="cuda")
            losses = torch.zeros(4, device="cuda")
            for _ in range(3):
                lib.licv_kd_loss_fwd_bwd(stu.data_ptr(), dst.data_ptr(), tea.data_ptr(), 0, lab.data_ptr(),
                                         0, R, R, 1.0, 1e-6, 0.5, 0, 1.0, losses.data_ptr(), ws.data_ptr(),
                                         R, V, V, V, code, 16, st)
            torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
