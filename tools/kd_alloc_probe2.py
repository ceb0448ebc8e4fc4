"""Debug: what makes buffers slow after the caching allocator reuses blocks?  Times a plain torch copy
(read 1 + write 1) at 524 MB per buffer in several allocation histories."""
import sys, time, torch

def bw(dst, src, n=4):
    s = torch.cuda.current_stream()
    out = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s); dst.copy_(src); b.record(s); torch.cuda.synchronize()
        out.append(round(2 * src.numel() * src.element_size() / a.elapsed_time(b) / 1e6, 0))
    return out

N = 8192 * 32002
mode = sys.argv[1]
def mk_like_bench():
    return (torch.randn(8192, 32002, device="cuda") * 3).to(torch.bfloat16)
if mode == "A":      # plain empty buffers, reuse after free
    a = torch.empty(N, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
    print("A fresh", bw(b, a))
    del a, b
    a = torch.empty(N, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
    print("A reused blocks", bw(b, a))
elif mode == "B":    # randn temporaries (1 GB fp32 blocks) then bf16 tensors carved from them
    a = mk_like_bench(); b = mk_like_bench()
    print("B fresh", bw(b, a), "reserved GB", torch.cuda.memory_reserved() / 2**30)
    del a, b
    a = mk_like_bench(); b = mk_like_bench()
    print("B second generation", bw(b, a), "reserved GB", torch.cuda.memory_reserved() / 2**30)
    c = torch.empty(N, dtype=torch.bfloat16, device="cuda")
    print("B copy into another cached block", bw(c, a), hex(c.data_ptr()), hex(a.data_ptr()), hex(b.data_ptr()))
    print(torch.cuda.memory_summary(abbreviated=True)[:1500])
elif mode == "C":    # small tensors in between, as quick_bench does
    a = mk_like_bench(); b = mk_like_bench(); c = torch.empty_like(a)
    print("C fresh", bw(c, a))
    del a, b, c
    small = [(torch.randn(2048, 32002, device="cuda") * 3).to(torch.bfloat16) for _ in range(6)]
    print("C small", bw(small[1], small[0]))
    del small
    a = mk_like_bench(); b = mk_like_bench(); c = torch.empty_like(a)
    print("C after small", bw(c, a), bw(b, a), bw(c, b))
    time.sleep(3)
    print("C after 3 s", bw(c, a))
    for t, nm in ((a, "a"), (b, "b"), (c, "c")):
        d = torch.empty_like(t)
        pass
    torch.cuda.empty_cache()
    print("C after empty_cache (same live tensors)", bw(c, a))
