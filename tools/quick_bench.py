"""Quick kernel timing sweep (CUDA events, rotating buffers larger than L2)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import ops  # noqa: E402


def timeit(fn, n_iter=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_iter):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n_iter * 1e-3


def main():
    peak = 6549.4
    res = []
    d = 4096
    for n_tok in [24, 256, 2048, 16384, 65536, 131072]:
        nbuf = max(2, min(32, int(400e6 // (n_tok * d * 2)) + 1))
        hs = [(torch.randn(n_tok, d, device="cuda") * 4).bfloat16() for _ in range(nbuf)]
        gs = [torch.randn(n_tok, d, device="cuda").bfloat16() for _ in range(nbuf)]
        outs = [torch.empty_like(hs[0]) for _ in range(nbuf)]
        s = torch.randn(d, device="cuda")
        ds = torch.zeros(d, device="cuda")
        from licv_vqa_b200 import _abi
        lib = _abi.load()
        st = torch.cuda.current_stream().cuda_stream

        def fwd(i):
            k = i % nbuf
            lib.licv_inject_fwd(hs[k].data_ptr(), s.data_ptr(), outs[k].data_ptr(), n_tok, d, 1, 1, 0, st)

        def bwd(i):
            k = i % nbuf
            lib.licv_inject_bwd(hs[k].data_ptr(), gs[k].data_ptr(), s.data_ptr(), outs[k].data_ptr(),
                                ds.data_ptr(), n_tok, d, 1, 1, 0, st)

        tf = timeit(fwd, 50)
        tb = timeit(bwd, 50)
        res.append(dict(kernel="inject_fwd", n_tok=n_tok, us=tf * 1e6, gbs=4 * n_tok * d / tf / 1e9,
                        frac=4 * n_tok * d / tf / 1e9 / peak))
        res.append(dict(kernel="inject_bwd", n_tok=n_tok, us=tb * 1e6, gbs=6 * n_tok * d / tb / 1e9,
                        frac=6 * n_tok * d / tb / 1e9 / peak))
        del hs, gs, outs
    V = 32002
    for R in [32, 256, 2048, 8192]:
        nbuf = max(2, min(8, int(400e6 // (R * V * 2)) + 1))
        stus = [(torch.randn(R, V, device="cuda") * 3).bfloat16() for _ in range(nbuf)]
        teas = [(torch.randn(R, V, device="cuda") * 3).bfloat16() for _ in range(nbuf)]
        lab = torch.randint(0, V, (R,), device="cuda")

        def kd(i):
            k = i % nbuf
            ops.kd_loss_raw(stus[k], teas[k], None, lab, None, R, R, 1.0, 1e-6, 0.5, in_place=True)

        def kd_ce(i):
            k = i % nbuf
            ops.kd_loss_raw(stus[k], None, None, lab, None, 0, R, 1.0, 1e-6, 0.5, only_hard_loss=True,
                            in_place=True)

        t = timeit(kd, 10)
        res.append(dict(kernel="kd_loss kl+ce", R=R, us=t * 1e6, gbs=6 * R * V / t / 1e9,
                        frac=6 * R * V / t / 1e9 / peak))
        t = timeit(kd_ce, 10)
        res.append(dict(kernel="kd_loss ce-only", R=R, us=t * 1e6, gbs=4 * R * V / t / 1e9,
                        frac=4 * R * V / t / 1e9 / peak))
        del stus, teas
    for r in res:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
