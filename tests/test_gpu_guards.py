"""Home-made bounds checks (compute-sanitizer is closed on this pool): every kernel runs on a
payload embedded in a larger allocation whose surroundings are NaN (reads that stray show up as
NaN in the results, which are compared with the oracle) and whose output surroundings hold a
sentinel that must survive (writes that stray are caught bit-exactly)."""
import numpy as np
import pytest
import torch

from oracle import licv_oracle as O
from tests.util import EPS, rel_err

pytestmark = pytest.mark.gpu

TD = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}
GUARD = 4096      # elements on each side


@pytest.fixture(scope="module")
def ops():
    from licv_vqa_b200 import ops as _ops
    return _ops


def host(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


def embed(payload, fill):
    """payload [n, d] -> (view into a guarded flat buffer, the buffer)."""
    n = payload.numel()
    buf = torch.full((n + 2 * GUARD,), fill, dtype=payload.dtype, device="cuda")
    view = buf[GUARD:GUARD + n].view(payload.shape)
    view.copy_(payload)
    return view, buf


def guards_intact(buf, n, fill):
    a, b = buf[:GUARD].float(), buf[GUARD + n:].float()
    if np.isnan(fill):
        return bool(torch.isnan(a).all() and torch.isnan(b).all())
    return bool((a == fill).all() and (b == fill).all())


@pytest.mark.parametrize("dt", ["bf16", "fp16", "fp32"])
@pytest.mark.parametrize("n_tok,d", [(1, 4096), (37, 4096), (297, 4096), (5, 512), (33, 520), (9, 8192),
                                     (7, 2048), (600, 2048), (3, 72)])
def test_inject_stays_inside_its_buffers(ops, dt, n_tok, d):
    rng = np.random.default_rng(n_tok + d)
    tdt = TD[dt]
    h, hbuf = embed(torch.tensor(rng.normal(size=(n_tok, d)) * 3, dtype=torch.float32).to(tdt).cuda(), float("nan"))
    g, gbuf = embed(torch.tensor(rng.normal(size=(n_tok, d)), dtype=torch.float32).to(tdt).cuda(), float("nan"))
    s, sbuf = embed(torch.tensor(rng.normal(size=(1, d)), dtype=torch.float32).cuda(), float("nan"))
    s = s.view(d)
    out, obuf = embed(torch.zeros(n_tok, d, dtype=tdt, device="cuda"), 7.0)
    dh, dbuf = embed(torch.zeros(n_tok, d, dtype=tdt, device="cuda"), 7.0)
    ds, dsbuf = embed(torch.zeros(1, d, device="cuda"), 7.0)
    ds = ds.view(d)
    from licv_vqa_b200 import _abi
    lib, code = _abi.load(), {"bf16": _abi.BF16, "fp16": _abi.F16, "fp32": _abi.F32}[dt]
    st = torch.cuda.current_stream().cuda_stream
    _abi.check(lib.licv_inject_fwd(h.data_ptr(), s.data_ptr(), out.data_ptr(), n_tok, d, code, code, 0, st))
    _abi.check(lib.licv_inject_bwd(h.data_ptr(), g.data_ptr(), s.data_ptr(), dh.data_ptr(), ds.data_ptr(),
                                   n_tok, d, code, code, 0, st))
    torch.cuda.synchronize()
    n = n_tok * d
    assert guards_intact(obuf, n, 7.0) and guards_intact(dbuf, n, 7.0) and guards_intact(dsbuf, d, 7.0)
    assert guards_intact(hbuf, n, float("nan")) and guards_intact(gbuf, n, float("nan"))
    ref = O.inject_fwd(host(h), host(s), out_fmt=dt)
    tol = 2e-6 if dt == "fp32" else 2 * EPS[dt]
    assert np.all(np.abs(host(out) - ref) <= np.abs(ref) * tol + 1e-30)      # no NaN leaked in
    o_dh, o_ds = O.inject_bwd(host(h), host(s), host(g))
    assert rel_err(host(ds), o_ds) < 1e-5
    assert rel_err(host(dh), o_dh) < (1e-5 if dt == "fp32" else 1.2 * EPS[dt])


@pytest.mark.parametrize("dt,V,R,Rt,pad_s,pad_t", [
    ("bf16", 32002, 13, 9, 0, 0), ("bf16", 32002, 13, 9, 6, 3), ("fp16", 32003, 11, 11, 0, 0),
    ("fp16", 32003, 6, 8, 1, 5), ("fp32", 32002, 5, 5, 0, 1), ("bf16", 50257, 5, 5, 0, 0),
    ("bf16", 33, 4, 4, 0, 0), ("fp32", 1003, 7, 7, 3, 0), ("bf16", 4097, 9, 9, 0, 0)])
def test_kd_loss_stays_inside_its_rows(ops, dt, V, R, Rt, pad_s, pad_t):
    """Row padding, the space before the first row and after the last one are NaN / sentinel."""
    rng = np.random.default_rng(V + R)
    tdt = TD[dt]
    s_pay = torch.full((R, V + pad_s), float("nan"), dtype=tdt, device="cuda")
    t_pay = torch.full((Rt, V + pad_t), float("nan"), dtype=tdt, device="cuda")
    s_pay[:, :V] = torch.tensor(rng.normal(size=(R, V)) * 3, dtype=torch.float32).to(tdt).cuda()
    t_pay[:, :V] = torch.tensor(rng.normal(size=(Rt, V)) * 3, dtype=torch.float32).to(tdt).cuda()
    sv, sbuf = embed(s_pay, float("nan"))
    tv, tbuf = embed(t_pay, float("nan"))
    gv, gbuf = embed(torch.full((R, V + pad_s), 7.0, dtype=tdt, device="cuda"), 7.0)
    stu, tea, dstu = sv[:, :V], tv[:, :V], gv[:, :V]
    ktr = np.full(R, -1, np.int32)
    ktr[::2] = np.arange((R + 1) // 2) % Rt
    lab = rng.integers(0, V, size=R).astype(np.int64)
    lab[1] = -100
    lab[2] = V - 1
    lab[0] = 0
    n_kl, n_ce = int((ktr >= 0).sum()), int((lab != -100).sum())
    from licv_vqa_b200 import _abi
    lib, code = _abi.load(), {"bf16": _abi.BF16, "fp16": _abi.F16, "fp32": _abi.F32}[dt]
    ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8, device="cuda")
    losses = torch.zeros(4, device="cuda")
    d_ktr, d_lab = torch.tensor(ktr).cuda(), torch.tensor(lab).cuda()
    _abi.check(lib.licv_kd_loss_fwd_bwd(
        stu.data_ptr(), dstu.data_ptr(), tea.data_ptr(), d_ktr.data_ptr(), d_lab.data_ptr(), 0, n_kl, n_ce,
        1.0, 1e-6, 0.5, 0, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V + pad_s, V + pad_t, code, 16,
        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    # writes: the padding columns of the gradient rows and both guards still hold the sentinel
    assert guards_intact(gbuf, R * (V + pad_s), 7.0)
    if pad_s:
        assert bool((gv[:, V:].float() == 7.0).all())
    want = O.kd_loss_rows(host(stu), host(tea), ktr, lab, 1.0, 1e-6, 0.5,
                          logit_fmt=None if dt == "fp32" else dt)
    assert abs(float(losses[2]) - want["loss"]) <= 1e-5 * abs(want["loss"])    # no NaN leaked in
    assert rel_err(host(dstu), want["d_stu"]) < (1e-5 if dt == "fp32" else 1.2 * EPS[dt])


def test_two_threads_two_streams_give_the_single_thread_results():
    """The entry points are re-entrant (backward runs on autograd's own thread, on
    whatever stream is current there): two host threads, each on its own stream with its own
    tensors, running injection fwd/bwd and the loss at the same time give exactly what one thread
    gives."""
    import threading

    import numpy as np
    from licv_vqa_b200 import ops

    def work(seed, out, use_stream):
        g = torch.Generator(device="cuda").manual_seed(seed)
        h = (torch.randn(300, 4096, device="cuda", generator=g) * 4).to(torch.bfloat16)
        gr = torch.randn(300, 4096, device="cuda", generator=g).to(torch.bfloat16)
        s = torch.randn(4096, device="cuda", generator=g)
        stu = (torch.randn(600, 32002, device="cuda", generator=g) * 3).to(torch.bfloat16)
        tea = (torch.randn(600, 32002, device="cuda", generator=g) * 3).to(torch.bfloat16)
        lab = torch.randint(0, 32002, (600,), device="cuda", generator=g)
        torch.cuda.synchronize()
        stream = torch.cuda.Stream() if use_stream else torch.cuda.current_stream()
        with torch.cuda.stream(stream):
            res = []
            for _ in range(6):
                o = ops.inject_forward(h, s)
                ds = torch.zeros(4096, device="cuda")
                dh = ops.inject_backward(h, gr, s, ds, True, 0)
                losses, dstu = ops.kd_loss_raw(stu, tea, None, lab, None, 600, 600, 1.0, 1e-6, 0.5,
                                               in_place=False)
                res = [o, dh, ds, losses, dstu]
            stream.synchronize()
        out[seed] = [t.float().cpu().numpy() for t in res]

    ref, got = {}, {}
    for seed in (1, 2):
        work(seed, ref, False)
    threads = [threading.Thread(target=work, args=(seed, got, True)) for seed in (1, 2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for seed in (1, 2):
        for a, b, exact in zip(ref[seed], got[seed], (True, True, False, True, True)):
            if exact:
                assert np.array_equal(a, b)
            else:   # d_shift: fp32 atomics, summation order
                assert np.allclose(a, b, rtol=1e-5, atol=1e-5)
