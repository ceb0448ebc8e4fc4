"""The reference's op chain as stock eager PyTorch ON THE SAME B200, next to this repository's
kernels ("time the same reference functions in eager CUDA on the B200 - that is
the real comparator").

    python tools/eager_cuda_compare.py > gpurun_out/eager_cuda.jsonl

Per shape, forward + backward of
  * the injection hook body, `icv_intervention.py:66-72`: y = h + s; y / ||y|| * ||h||, backward by
    autograd (what the reference runs), against licv_inject_fwd + licv_inject_bwd;
  * the loss, `icv_module.py:121-134` + HF's CE: softmax x2, the eps-logarithms, mean, cross
    entropy, backward by autograd, against licv_kd_loss_fwd_bwd.
The eager chain is restated here in five lines (it is the comparator, not the product and not the
oracle); timing = CUDA events around `reps` iterations after warm-up, rotating inputs > L2.
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import _abi  # noqa: E402


def timed(fn, nbuf, reps=4, warm=2):
    for i in range(warm * nbuf):
        fn(i % nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps * nbuf):
        fn(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (reps * nbuf)


def main():
    lib = _abi.load()
    out = []
    d, V = 4096, 32002
    # ---- injection -------------------------------------------------------------------------
    for n_tok, dt, code in ((256, torch.float16, _abi.F16), (8192, torch.bfloat16, _abi.BF16),
                            (131072, torch.bfloat16, _abi.BF16)):
        nbuf = max(2, min(16, int(600e6 // (n_tok * d * 6)) + 1))
        hs = [(torch.randn(n_tok, d, device="cuda") * 4).to(dt) for _ in range(nbuf)]
        gs = [torch.randn(n_tok, d, device="cuda").to(dt) for _ in range(nbuf)]
        o = torch.empty_like(hs[0])
        dh = torch.empty_like(hs[0])
        s32 = torch.randn(d, device="cuda")
        ds = torch.zeros(d, device="cuda")
        shift = s32.to(dt).requires_grad_(True)       # the reference's DeepSpeed recipe: ICV in the tower's dtype

        def eager(k):
            h = hs[k].view(1, n_tok, d).requires_grad_(True)
            y = h + shift.view(1, 1, d)
            y = y / y.norm(dim=-1, keepdim=True) * h.norm(dim=-1, keepdim=True)
            y.backward(gs[k].view(1, n_tok, d))
            h.grad = None
            shift.grad = None

        def ours(k):
            st = torch.cuda.current_stream().cuda_stream
            _abi.check(lib.licv_inject_fwd(hs[k].data_ptr(), s32.data_ptr(), o.data_ptr(), n_tok, d,
                                           code, code, 0, st))
            _abi.check(lib.licv_inject_bwd(hs[k].data_ptr(), gs[k].data_ptr(), s32.data_ptr(),
                                           dh.data_ptr(), ds.data_ptr(), n_tok, d, code, code, 0, st))

        te, to = timed(eager, nbuf), timed(ours, nbuf)
        out.append({"op": "inject fwd+bwd", "n_tok": n_tok, "d": d, "dtype": str(dt).split(".")[-1],
                    "eager_cuda_us": round(te * 1e6, 1), "licv_us": round(to * 1e6, 1),
                    "speedup": round(te / to, 2),
                    "licv_gbs": round(5 * 2 * n_tok * d / to / 1e9, 1)})
        print(json.dumps(out[-1]), flush=True)
        del hs, gs
    # ---- loss ------------------------------------------------------------------------------
    for R, dt, code in ((64, torch.float16, _abi.F16), (2048, torch.bfloat16, _abi.BF16),
                        (8192, torch.bfloat16, _abi.BF16)):
        nbuf = max(2, min(4, int(600e6 // (R * V * 6)) + 1))
        stus = [(torch.randn(R, V, device="cuda") * 3).to(dt) for _ in range(nbuf)]
        teas = [(torch.randn(R, V, device="cuda") * 3).to(dt) for _ in range(nbuf)]
        lab = torch.randint(0, V, (R,), device="cuda")
        dst = torch.empty_like(stus[0])
        ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8, device="cuda")
        losses = torch.zeros(4, device="cuda")
        eps = 1e-6

        def eager(k):
            stu = stus[k].requires_grad_(True)
            sl, tl = stu.float(), teas[k].float()                 # autocast: softmax / log in fp32
            p, q = torch.softmax(tl, 1), torch.softmax(sl, 1)
            kl = (p * (torch.log(p + eps) - torch.log(q + eps))).sum(1).mean()
            loss = kl + 0.5 * F.cross_entropy(sl, lab)
            loss.backward()
            stu.grad = None

        def ours(k):
            st = torch.cuda.current_stream().cuda_stream
            _abi.check(lib.licv_kd_loss_fwd_bwd(
                stus[k].data_ptr(), dst.data_ptr(), teas[k].data_ptr(), 0, lab.data_ptr(), 0, R, R,
                1.0, eps, 0.5, 0, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, st))

        te, to = timed(eager, nbuf), timed(ours, nbuf)
        out.append({"op": "KL + 0.5 CE fwd+bwd", "rows": R, "V": V, "dtype": str(dt).split(".")[-1],
                    "eager_cuda_us": round(te * 1e6, 1), "licv_us": round(to * 1e6, 1),
                    "speedup": round(te / to, 2),
                    "licv_gbs": round(6 * R * V / to / 1e9, 1)})
        print(json.dumps(out[-1]), flush=True)
        del stus, teas


if __name__ == "__main__":
    main()
