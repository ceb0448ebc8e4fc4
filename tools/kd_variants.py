"""A/B harness for loss-kernel variants: build alternative copies of the library from one changed
source file, then time and check each on the GPU in its own process (LICV_LIB selects the build).

    python tools/kd_variants.py build NAME path/to/licv_kd_loss_stream.cu [nvcc flags...]   (CPU box)
    python tools/kd_variants.py run [NAME ...]                                              (GPU box)
    python tools/kd_variants.py one                                                         (child)

Variants live in tools/bin/kdv/ (git-ignored, travels to the GPU box).  The check is a float64
torch restatement of icv_module.py:121-134 + the shifted CE on 96 rows (not the oracle: tools may
not import it); parity proper stays in tests/.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tools", "bin", "kdv")
sys.path.insert(0, ROOT)


def build(name, src, extra):
    from licv_vqa_b200 import build as b
    b.build()
    os.makedirs(OUT, exist_ok=True)
    obj = os.path.join(OUT, name + ".o")
    cmd = [b._nvcc(), *b.NVCC_FLAGS, *extra, "-I", b.INCLUDE, "-I", b.CSRC, "-c", src, "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    lines = [l for l in r.stdout.splitlines() if "stream_kernel" in l or "spill" in l or "Used" in l or "error" in l]
    print("\n".join(lines[-12:]))
    if r.returncode:
        print(r.stdout[-3000:])
        sys.exit(1)
    base = os.path.basename(src)[:-3] + ".o"
    others = [os.path.join(b.LIB_DIR, "obj", o) for o in sorted(os.listdir(os.path.join(b.LIB_DIR, "obj")))
              if o.endswith(".o") and o != "licv_kd_loss_stream.o" and o != base]
    so = os.path.join(OUT, name + ".so")
    subprocess.run([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so, obj, *others],
                   check=True)
    os.remove(obj)
    print(so)


def one():
    import torch
    from licv_vqa_b200 import _abi
    lib = _abi.load()
    lib.licv_debug_set_kd_stream(2)       # the stream kernel also for the check's few rows
    V = int(os.environ.get("KDV_V", "32002"))
    R = int(os.environ.get("KDV_ROWS", "8192"))
    res = {"lib": os.path.basename(_abi.LIB_PATH)}
    st = torch.cuda.current_stream().cuda_stream

    def call(stu, dst, tea, lab, n_kl, n_ce, only_hard, code, T=1.0, ws=None, losses=None):
        Rr = stu.shape[0]
        _abi.check(lib.licv_kd_loss_fwd_bwd(stu.data_ptr(), dst.data_ptr(), tea.data_ptr() if tea is not None else 0, 0,
                                            lab.data_ptr() if lab is not None else 0, 0, n_kl, n_ce, T, 1e-6, 0.5,
                                            only_hard, 1.0, losses.data_ptr(), ws.data_ptr(), Rr, V, stu.stride(0),
                                            tea.stride(0) if tea is not None else V, code, 16, st))

    # ---- check: 96 rows against float64 torch ------------------------------------------------
    torch.manual_seed(1)
    for dt, code, nm in ((torch.bfloat16, _abi.BF16, "bf16"), (torch.float16, _abi.F16, "fp16")):
        Rc = 444     # three rows per CTA: first / middle / last sweeps all run
        stu = (torch.randn(Rc, V, device="cuda") * 3).to(dt)
        tea = (stu.float() + torch.randn(Rc, V, device="cuda")).to(dt)
        lab = torch.randint(0, V, (Rc,), device="cuda")
        ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(Rc) + 64, dtype=torch.uint8, device="cuda")
        losses = torch.zeros(4, device="cuda")
        for T in (1.0,):
            for mode in ("klce", "ce"):
                dst = torch.empty_like(stu)
                s64 = stu.double().requires_grad_(True)
                ce = torch.nn.functional.cross_entropy(s64, lab, reduction="mean")
                if mode == "klce":
                    p = torch.softmax(tea.double() / T, -1)
                    q = torch.softmax(s64 / T, -1)
                    kl = (p * (torch.log(p + 1e-6) - torch.log(q + 1e-6))).sum(-1).mean() * T * T
                    total = kl + 0.5 * ce
                    call(stu, dst, tea, lab, Rc, Rc, 0, code, T, ws, losses)
                else:
                    kl = torch.zeros((), dtype=torch.float64, device="cuda")
                    total = ce
                    call(stu, dst, None, lab, 0, Rc, 1, code, T, ws, losses)
                total.backward()
                torch.cuda.synchronize()
                g = s64.grad
                got = losses.double()
                err_l = abs(got[2].item() - total.item()) / abs(total.item())
                err_g = ((dst.double() - g).abs().max() / g.abs().max()).item()
                # error beyond the rounding of the output format
                gr = g.to(dt).double()
                err_gr = ((dst.double() - gr).abs().max() / g.abs().max()).item()
                res[f"{nm}_{mode}_loss_rel"] = float("%.3g" % err_l)
                res[f"{nm}_{mode}_grad_relmax"] = float("%.3g" % err_g)
                res[f"{nm}_{mode}_grad_vs_rounded"] = float("%.3g" % err_gr)
        del stu, tea, dst

    # ---- timing: CUDA-graph replays over two buffer sets (each far beyond L2) -------------------
    for dt, code, nm in ((torch.bfloat16, _abi.BF16, "bf16"), (torch.float16, _abi.F16, "fp16")):
        stus = [(torch.randn(R, V, device="cuda") * 3).to(dt) for _ in range(2)]
        teas = [(torch.randn(R, V, device="cuda") * 3).to(dt) for _ in range(2)]
        dsts = [torch.empty_like(stus[0]) for _ in range(2)]
        lab = torch.randint(0, V, (R,), device="cuda")
        ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8, device="cuda")
        losses = torch.zeros(4, device="cuda")
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            sst = s.cuda_stream

            def kd(k, ce_only):
                if ce_only:
                    _abi.check(lib.licv_kd_loss_fwd_bwd(stus[k].data_ptr(), dsts[k].data_ptr(), 0, 0, lab.data_ptr(), 0,
                                                        0, R, 1.0, 1e-6, 0.5, 1, 1.0, losses.data_ptr(), ws.data_ptr(),
                                                        R, V, V, V, code, 16, sst))
                else:
                    _abi.check(lib.licv_kd_loss_fwd_bwd(stus[k].data_ptr(), dsts[k].data_ptr(), teas[k].data_ptr(), 0,
                                                        lab.data_ptr(), 0, R, R, 1.0, 1e-6, 0.5, 0, 1.0,
                                                        losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, sst))
            for ce_only in (False, True):
                for k in (0, 1):
                    kd(k, ce_only)
                s.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    for k in (0, 1):
                        kd(k, ce_only)
                g.replay()
                s.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                best = 1e9
                for _ in range(3):
                    a.record(s)
                    for _ in range(4):
                        g.replay()
                    b.record(s)
                    s.synchronize()
                    best = min(best, a.elapsed_time(b) * 1e3 / 8)
                res[f"{nm}_{'ce' if ce_only else 'klce'}_us"] = round(best, 1)
        del stus, teas, dsts
    print("KDV " + json.dumps(res), flush=True)


def run(names):
    sos = sorted(f for f in os.listdir(OUT) if f.endswith(".so"))
    if names:
        sos = [f for f in sos if f[:-3] in names]
    for f in sos:
        env = dict(os.environ, LICV_LIB=os.path.join(OUT, f))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, timeout=600)
        lines = [l for l in r.stdout.splitlines() if l.startswith("KDV ")]
        print(f, lines[-1] if lines else "FAILED rc=%d\n%s" % (r.returncode, r.stdout[-1500:]), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2], sys.argv[3], sys.argv[4:])
    elif sys.argv[1] == "run":
        run(sys.argv[2:])
    else:
        one()
