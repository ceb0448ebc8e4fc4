"""Debug: bisect what makes reused buffers slow after the loss kernel ran."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import _abi
lib = _abi.load()
V = 32002
code = _abi.BF16
which = sys.argv[1] if len(sys.argv) > 1 else "stream"
lib.licv_debug_set_kd_stream(1 if which == "stream" else 0)

def bw(dst, src, n=3):
    s = torch.cuda.current_stream()
    out = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); dst.copy_(src); b.record(s); torch.cuda.synchronize()
        out.append(round(2 * src.numel() * src.element_size() / a.elapsed_time(b) / 1e6))
    return out

def kd_us(stu, tea, dst, R, n=3):
    lab = torch.randint(0, V, (R,), device="cuda")
    ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8, device="cuda")
    losses = torch.zeros(4, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    out = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _abi.check(lib.licv_kd_loss_fwd_bwd(stu.data_ptr(), dst.data_ptr(), tea.data_ptr(), 0, lab.data_ptr(), 0, R, R,
                       1.0, 1e-6, 0.5, 0, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, st)); b.record()
        torch.cuda.synchronize()
        out.append(round(a.elapsed_time(b) * 1e3, 1))
    return out

def mk(R):
    return (torch.randn(R, V, device="cuda") * 3).to(torch.bfloat16)

R = 8192
stu, tea, dst = mk(R), mk(R), torch.empty(R, V, dtype=torch.bfloat16, device="cuda")
print("1 fresh: copy GB/s", bw(dst, tea), "kernel us", kd_us(stu, tea, dst, R), "copy after", bw(dst, tea), flush=True)
del stu, tea, dst
stu, tea, dst = mk(R), mk(R), torch.empty(R, V, dtype=torch.bfloat16, device="cuda")
print("2 reused, same size: copy GB/s", bw(dst, tea), "kernel us", kd_us(stu, tea, dst, R), "copy after", bw(dst, tea), flush=True)
del stu, tea, dst
R2 = 2048
s2, t2, d2 = mk(R2), mk(R2), torch.empty(R2, V, dtype=torch.bfloat16, device="cuda")
print("3 small (2048) in reused blocks: copy GB/s", bw(d2, t2), "kernel us", kd_us(s2, t2, d2, R2), "copy after", bw(d2, t2), flush=True)
del s2, t2, d2
stu, tea, dst = mk(R), mk(R), torch.empty(R, V, dtype=torch.bfloat16, device="cuda")
print("4 big again: copy GB/s", bw(dst, tea), "kernel us", kd_us(stu, tea, dst, R), "copy after", bw(dst, tea), flush=True)
print("  reserved GB", round(torch.cuda.memory_reserved() / 2**30, 2), "allocated GB", round(torch.cuda.memory_allocated() / 2**30, 2))
big = torch.empty(R * V * 3, dtype=torch.bfloat16, device="cuda")     # one fresh 1.5 GB segment
a, b, c = big[:R * V].view(R, V), big[R * V:2 * R * V].view(R, V), big[2 * R * V:].view(R, V)
a.copy_(stu); b.copy_(tea)
print("5 same data in one fresh segment: copy GB/s", bw(c, b), "kernel us", kd_us(a, b, c, R), flush=True)
