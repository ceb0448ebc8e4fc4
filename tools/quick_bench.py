"""Kernel timing sweep through the C ABI (CUDA events around CUDA-graph replays; rotating buffers
larger than L2, so every launch streams from HBM).

    python tools/quick_bench.py [inject] [kd] [--dtype bf16|fp16] [--tokens 256,131072] [--rows 32,2048]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import _abi  # noqa: E402

PEAK = 6549.4
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                       "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


class Clocks:
    """SM clock / power samples (NVML) while a timed region runs."""

    def __init__(self):
        import threading
        self.mhz, self.watts, self._stop = [], [], threading.Event()
        self.mem_mhz, self.reasons, self.temps = [], 0, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        except Exception:
            self.nv = None
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set() and self.nv is not None:
            try:
                self.mhz.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.watts.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                self.mem_mhz.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_MEM))
                self.reasons |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.temps.append(self.nv.nvmlDeviceGetTemperature(self.h, self.nv.NVML_TEMPERATURE_GPU))
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join()

    def summary(self):
        m = sorted(self.mhz)
        return {"sm_mhz": m[len(m) // 2] if m else None, "sm_mhz_min": m[0] if m else None,
                "watts_max": round(max(self.watts), 1) if self.watts else None,
                "mem_mhz_min": min(self.mem_mhz) if self.mem_mhz else None,
                "reasons": hex(self.reasons), "temp_max": max(self.temps) if self.temps else None}


LAST_CLOCKS = {}


def time_graph(launch, nbuf, reps=5, min_seconds=0.25):
    """launch(k) enqueues one kernel on buffer set k; returns seconds per launch."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for k in range(nbuf):
            launch(k)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for k in range(nbuf):
                launch(k)
        for _ in range(3):
            g.replay()
        s.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        s.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / (reps * nbuf)
        # a second, longer region with the clocks sampled: what the SM clock settles to under load
        n2 = max(reps, int(min_seconds / max(t * nbuf, 1e-6)))
        with Clocks() as ck:
            e0.record(s)
            for _ in range(n2):
                g.replay()
            e1.record(s)
            s.synchronize()
        LAST_CLOCKS.clear()
        LAST_CLOCKS.update(ck.summary(), us_sustained=round(e0.elapsed_time(e1) * 1e3 / (n2 * nbuf), 2))
    return t


def cta_stats(lib):
    """Debug builds (-DLICV_TRACE): duration of every CTA of the last stream-kernel launch."""
    if not hasattr(lib, "licv_debug_read_stream_trace"):
        return {}
    import ctypes
    import numpy as np
    buf = (ctypes.c_ulonglong * (256 * 4))()
    lib.licv_debug_read_stream_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.licv_debug_read_stream_trace(buf, 256 * 4)
    t = np.array(buf, dtype=np.int64).reshape(256, 4)[:148]
    dur = (t[:, 1] - t[:, 0]) / 1e3
    return {"cta_us_min": round(float(dur.min()), 1), "cta_us_med": round(float(np.median(dur)), 1),
            "cta_us_max": round(float(dur.max()), 1),
            "cta_start_spread_us": round(float((t[:, 0].max() - t[:, 0].min()) / 1e3), 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["inject", "kd"])
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--tokens", default="24,256,2048,16384,65536,131072")
    ap.add_argument("--rows", default="32,256,2048,8192")
    ap.add_argument("--vocab", type=int, default=32002)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--d", type=int, default=4096)
    args = ap.parse_args()
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    code = {"bf16": _abi.BF16, "fp16": _abi.F16, "fp32": _abi.F32}[args.dtype]
    e = 4 if args.dtype == "fp32" else 2
    lib = _abi.load()
    d = args.d
    if "inject" in args.which:
        for n_tok in [int(x) for x in args.tokens.split(",")]:
            torch.cuda.empty_cache()
            nbuf = max(2, min(32, int(600e6 // (n_tok * d * e * 3)) + 1))
            hs = [(torch.randn(n_tok, d, device="cuda") * 4).to(dt) for _ in range(nbuf)]
            gs = [torch.randn(n_tok, d, device="cuda").to(dt) for _ in range(nbuf)]
            outs = [torch.empty_like(hs[0]) for _ in range(nbuf)]
            s = torch.randn(d, device="cuda")
            ds = torch.zeros(d, device="cuda")

            def fwd(k):
                st = torch.cuda.current_stream().cuda_stream
                _abi.check(lib.licv_inject_fwd(hs[k].data_ptr(), s.data_ptr(), outs[k].data_ptr(),
                                               n_tok, d, code, code, args.flags, st))

            def bwd(k):
                st = torch.cuda.current_stream().cuda_stream
                _abi.check(lib.licv_inject_bwd(hs[k].data_ptr(), gs[k].data_ptr(), s.data_ptr(),
                                               outs[k].data_ptr(), ds.data_ptr(), n_tok, d, code,
                                               code, args.flags, st))

            for name, fn, nb in (("inject_fwd", fwd, 2 * e), ("inject_bwd", bwd, 3 * e)):
                t = time_graph(fn, nbuf)
                gbs = nb * n_tok * d / t / 1e9
                print(json.dumps(dict(kernel=name, n_tok=n_tok, d=d, dtype=args.dtype,
                                      us=round(t * 1e6, 2), gbs=round(gbs, 1),
                                      frac=round(gbs / PEAK, 4), **LAST_CLOCKS)), flush=True)
            del hs, gs, outs
    if "kd" in args.which:
        V = args.vocab
        for R in [int(x) for x in args.rows.split(",")]:
            # fresh segments for every size: buffers carved out of blocks that the caching allocator
            # kept from a smaller size have shown 2-4x lower bandwidth for every kernel, torch's own
            # copy included (profiles/README.md, r2q)
            torch.cuda.empty_cache()
            nbuf = max(2, min(8, int(600e6 // (R * V * e * 3)) + 1))
            stus = [(torch.randn(R, V, device="cuda") * 3).to(dt) for _ in range(nbuf)]
            teas = [(torch.randn(R, V, device="cuda") * 3).to(dt) for _ in range(nbuf)]
            dst = [torch.empty_like(stus[0]) for _ in range(nbuf)]
            lab = torch.randint(0, V, (R,), device="cuda")
            ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8,
                             device="cuda")
            losses = torch.zeros(4, device="cuda")

            def kd(k):
                st = torch.cuda.current_stream().cuda_stream
                _abi.check(lib.licv_kd_loss_fwd_bwd(
                    stus[k].data_ptr(), dst[k].data_ptr(), teas[k].data_ptr(), 0, lab.data_ptr(),
                    0, R, R, 1.0, 1e-6, 0.5, 0, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V, V,
                    code, 16, st))

            def kd_ce(k):
                st = torch.cuda.current_stream().cuda_stream
                _abi.check(lib.licv_kd_loss_fwd_bwd(
                    stus[k].data_ptr(), dst[k].data_ptr(), 0, 0, lab.data_ptr(), 0, 0, R, 1.0,
                    1e-6, 0.5, 1, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, st))

            for name, fn, nb in (("kd_loss kl+ce", kd, 3 * e), ("kd_loss ce-only", kd_ce, 2 * e)):
                t = time_graph(fn, nbuf)
                gbs = nb * R * V / t / 1e9
                print(json.dumps(dict(kernel=name, R=R, V=V, dtype=args.dtype,
                                      us=round(t * 1e6, 2), gbs=round(gbs, 1),
                                      frac=round(gbs / PEAK, 4), **LAST_CLOCKS, **cta_stats(lib))),
                      flush=True)
            del stus, teas, dst


if __name__ == "__main__":
    main()
