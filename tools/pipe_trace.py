"""Debug: per-CTA phase timeline of the TMA backward at the training shape, inside the step's chain
of programmatic launches (needs a build with LICV_EXTRA_NVCC_FLAGS=-DLICV_TRACE; LICV_LIB selects it).
    python tools/pipe_trace.py [n_tok]"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import _abi
lib = _abi.load()
n_tok = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d, L, dt, code = 4096, 32, torch.float16, _abi.F16
h = [torch.randn(n_tok, d, device="cuda").to(dt) for _ in range(L)]
g = [torch.randn(n_tok, d, device="cuda").to(dt) for _ in range(L)]
dh = [torch.empty_like(h[0]) for _ in range(L)]
icv = torch.randn(L, d, device="cuda") * 0.1
R = min(16, max(1, lib.licv_inject_bwd_rows(n_tok, d, code, code)))
rows = torch.zeros(L, R, d, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    st = s.cuda_stream
    def chain():
        for l in reversed(range(L)):
            _abi.check(lib.licv_inject_bwd_spread(h[l].data_ptr(), g[l].data_ptr(), icv[l].data_ptr(), dh[l].data_ptr(),
                                                  rows[l].data_ptr(), R, n_tok, d, code, code, 15, st))
    chain(); s.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        chain()
    for _ in range(3): gr.replay()
    s.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(10): gr.replay()
    b.record(s); s.synchronize()
    print("us per launch in the chain: %.2f" % (a.elapsed_time(b) * 1e3 / 10 / L))
buf = (ctypes.c_ulonglong * (512 * 8))()
lib.licv_debug_read_pipe_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.licv_debug_read_pipe_trace(buf, 512 * 8)
t = np.array(buf, dtype=np.int64).reshape(512, 8)
t = t[t[:, 0] > 0][:, :6]
t0 = t[:, 0].min()
names = ["entry", "after pdl_wait", "stage landed", "reduction done", "loop done", "d_shift out"]
print("CTAs traced:", len(t), " (times in ns from the first CTA's entry: min / median / max)")
for i, nm in enumerate(names):
    v = t[:, i] - t0
    print("  %-16s %7d %7d %7d" % (nm, v.min(), np.median(v), v.max()))
dur = t[:, 5] - t[:, 0]
print("per-CTA entry->exit: median %d max %d;  pdl_wait share: median %d ns" % (np.median(dur), dur.max(), np.median(t[:, 1] - t[:, 0])))
