import os, sys, torch
sys.path.insert(0, os.getcwd())
from licv_vqa_b200 import _abi
lib = _abi.load()
R, V = 8192, 32002
dt, code = torch.bfloat16, _abi.BF16
stus = [(torch.randn(R, V, device="cuda") * 3).to(dt) for _ in range(2)]
teas = [(torch.randn(R, V, device="cuda") * 3).to(dt) for _ in range(2)]
dsts = [torch.empty_like(stus[0]) for _ in range(2)]
lab = torch.randint(0, V, (R,), device="cuda")
ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8, device="cuda")
losses = torch.zeros(4, device="cuda")
def kd(k, st):
    _abi.check(lib.licv_kd_loss_fwd_bwd(stus[k].data_ptr(), dsts[k].data_ptr(), teas[k].data_ptr(), 0, lab.data_ptr(), 0, R, R,
                             1.0, 1e-6, 0.5, 0, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, st))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    st = s.cuda_stream
    for _ in range(3): kd(0, st)
    s.synchronize()
    for label, seq in (("same buffer x6", [0]*6), ("alternating x6", [0,1]*3)):
        evs = []
        for k in seq:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s); kd(k, st); b.record(s); evs.append((a, b))
        s.synchronize()
        print(label, [round(a.elapsed_time(b)*1e3, 1) for a, b in evs], flush=True)
    # back to back without events in between
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for k in [0,1]*5: kd(k, st)
    b.record(s); s.synchronize()
    print("back-to-back x10 us/launch", round(a.elapsed_time(b)*1e3/10, 1), flush=True)
    for k in [0,1]: kd(k, st)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for k in [0,1]: kd(k, st)
    g.replay(); s.synchronize()
    a.record(s)
    for _ in range(5): g.replay()
    b.record(s); s.synchronize()
    print("graph x10 us/launch", round(a.elapsed_time(b)*1e3/10, 1), flush=True)

if hasattr(lib, "licv_debug_read_stream_trace"):
    import ctypes, numpy as np
    buf = (ctypes.c_ulonglong * (256 * 4))()
    lib.licv_debug_read_stream_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.licv_debug_read_stream_trace(buf, 256 * 4)
    t = np.array(buf, dtype=np.int64).reshape(256, 4)[:148]
    dur = (t[:, 1] - t[:, 0]) / 1e3
    start = (t[:, 0] - t[:, 0].min()) / 1e3
    order = np.argsort(dur)
    print("CTA duration us: min %.1f median %.1f max %.1f | start spread %.1f us" %
          (dur.min(), np.median(dur), dur.max(), start.max()))
    print("slowest CTAs (cta, smid, us):", [(int(i), int(t[i, 2]), round(float(dur[i]), 1)) for i in order[-8:]])
    print("fastest CTAs (cta, smid, us):", [(int(i), int(t[i, 2]), round(float(dur[i]), 1)) for i in order[:8]])
    hist, edges = np.histogram(dur, bins=8)
    print("histogram:", list(zip([round(float(e), 0) for e in edges[:-1]], hist.tolist())))

if hasattr(lib, "licv_debug_read_stream_phases"):
    buf = (ctypes.c_longlong * (2 * 16 * 8))()
    lib.licv_debug_read_stream_phases.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.licv_debug_read_stream_phases(buf, 2 * 16 * 8)
    ph = np.array(buf, dtype=np.int64).reshape(2, 16, 8)
    names = ["red1", "C", "red2", "setup", "sweep", "tail"]
    for c in (0, 1):
        print("CTA", (0, 5)[c], "clocks per row (rows >= 2), per warp:")
        for w in (0, 1, 2, 3, 7, 11, 15):
            n = max(int(ph[c, w, 6]), 1)
            print("   warp %2d  " % w + "  ".join("%s %6d" % (nm, ph[c, w, i] // n) for i, nm in enumerate(names)),
                  " total", int(ph[c, w, :6].sum() // n))
