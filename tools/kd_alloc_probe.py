"""Debug: does the loss kernel's time depend on where the caching allocator puts the buffers?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import _abi
lib = _abi.load()
V = 32002
dt, code = torch.bfloat16, _abi.BF16

def run(R, tag):
    stus = [(torch.randn(R, V, device="cuda") * 3).to(dt) for _ in range(2)]
    teas = [(torch.randn(R, V, device="cuda") * 3).to(dt) for _ in range(2)]
    dsts = [torch.empty_like(stus[0]) for _ in range(2)]
    lab = torch.randint(0, V, (R,), device="cuda")
    ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8, device="cuda")
    losses = torch.zeros(4, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        st = s.cuda_stream
        def kd(k):
            _abi.check(lib.licv_kd_loss_fwd_bwd(stus[k].data_ptr(), dsts[k].data_ptr(), teas[k].data_ptr(), 0, lab.data_ptr(), 0, R, R,
                       1.0, 1e-6, 0.5, 0, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, st))
        for k in (0, 1): kd(k)
        s.synchronize()
        res = []
        for k in (0, 1, 0, 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s); kd(k); b.record(s); s.synchronize()
            res.append(round(a.elapsed_time(b) * 1e3, 1))
        lib.licv_debug_set_kd_stream(0)
        old = []
        for k in (0, 1, 0, 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s); kd(k); b.record(s); s.synchronize()
            old.append(round(a.elapsed_time(b) * 1e3, 1))
        lib.licv_debug_set_kd_stream(1)
        cp = []
        for k in (0, 1):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s); dsts[k].copy_(teas[k]); b.record(s); s.synchronize()
            cp.append(round(a.elapsed_time(b) * 1e3, 1))
        print("   round-1 kernel on the same buffers:", old, " torch copy tea->dst:", cp)
    mb = 1 << 21
    print(tag, "R", R, "us per launch (buf 0,1,0,1):", res, "| offsets in 2MB page (KB): stu",
          [(t.data_ptr() % mb) // 1024 for t in stus], "tea", [(t.data_ptr() % mb) // 1024 for t in teas],
          "dst", [(t.data_ptr() % mb) // 1024 for t in dsts], flush=True)
    print("    addresses GB: stu", [round(t.data_ptr() / 2**30, 3) for t in stus], "tea", [round(t.data_ptr() / 2**30, 3) for t in teas],
          "dst", [round(t.data_ptr() / 2**30, 3) for t in dsts], flush=True)

run(8192, "fresh")
run(256, "then")
run(2048, "then")
run(8192, "after small sizes")
torch.cuda.empty_cache()
run(8192, "after empty_cache")
