"""GPU parity of the small kernels around the two hot ones and of the host-buffer entry points:
encoder product (a1+a2), get_mask (a6), row pairing / CE labels (a6+a7+a9), the fused clip+AdamW
step (f2), the data-parallel optimizer at world size 1, licv_*_host vs the device entry points."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import licv_oracle as O
from tests.util import load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from licv_vqa_b200 import ops as _ops
    return _ops


def host(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


ENC = load_golden("encoder_cases.npz")


@pytest.mark.parametrize("name", [str(n) for n in ENC["names"]])
def test_encoder_product_matches_reference_golden(ops, name):
    L, d, learn, sig = [int(x) for x in ENC[f"{name}/cfg"]]
    from licv_vqa_b200 import GlobalICVEncoder
    enc = GlobalICVEncoder(d, L, alpha_learnable=bool(learn), alpha_init_value=0.0,
                           use_sigmoid=bool(sig)).cuda()
    assert sorted(enc.state_dict().keys()) == ["alpha", "icv"]
    with torch.no_grad():
        enc.alpha.copy_(torch.tensor(ENC[f"{name}/alpha_raw"]))
        enc.icv.copy_(torch.tensor(ENC[f"{name}/vec"]))
    out = enc()
    assert out.in_context_feature is None
    assert rel_err(host(out.alpha), ENC[f"{name}/alpha_eff"]) < 1e-6
    icv = enc.scaled_icv()
    assert icv.shape == (1, L, d)
    assert rel_err(host(icv), ENC[f"{name}/icv"]) < 1e-6
    icv.backward(torch.tensor(ENC[f"{name}/g"]).cuda())
    assert rel_err(host(enc.icv.grad), ENC[f"{name}/dvec"]) < 1e-6
    if learn:
        assert rel_err(host(enc.alpha.grad), ENC[f"{name}/dalpha"]) < 1e-5
    else:
        assert enc.alpha.grad is None


MASK = load_golden("mask_cases.npz")


@pytest.mark.parametrize("name", [str(n) for n in MASK["names"]])
def test_get_mask_matches_reference_golden(ops, name):
    ids = torch.tensor(MASK[f"{name}/ids"]).cuda()
    lens = torch.tensor(MASK[f"{name}/len"]).cuda()
    m = ops.get_mask(ids, lens, int(MASK[f"{name}/pad"]))
    assert m.dtype == torch.bool
    assert np.array_equal(m.cpu().numpy(), MASK[f"{name}/mask"])     # bit-exact


@pytest.mark.parametrize("variant", ["idefics", "idefics2", "causal_lm"])
@pytest.mark.parametrize("B,Tq,Tc", [(4, 12, 40), (8, 32, 896), (1, 5, 3), (3, 700, 1300)])
def test_prepare_rows_vs_oracle(ops, variant, B, Tq, Tc):
    rng = np.random.default_rng(B * 1000 + Tq)
    qx = rng.integers(1, Tq, size=B)
    s_ids = rng.integers(3, 100, size=(B, Tq))
    s_att = np.ones((B, Tq), np.int64)
    for b in range(0, B, 2):                         # right padding on some rows
        npad = int(rng.integers(0, max(Tq // 3, 1)))
        if npad:
            s_ids[b, Tq - npad:] = 0
            s_att[b, Tq - npad:] = 0
    img = 32001 if variant == "idefics2" else -1
    if variant == "idefics2":
        s_ids[:, 2:4] = img
    ctx = rng.integers(3, 100, size=(B, Tc))
    t_ids = np.concatenate([ctx, s_ids[:, 1:]], axis=1)      # the collator's contract
    icl = Tc + qx - 1
    s_mask = O.get_mask(s_ids, qx, 0)
    t_mask = O.get_mask(t_ids, icl, 0)
    want_ktr = O.pair_rows(s_mask, t_mask)
    want_lab = O.ce_labels(s_ids, s_att, variant, img if img >= 0 else None).reshape(-1)
    ktr, lab, counts = ops.kd_prepare_rows(
        torch.tensor(s_ids).cuda(), torch.tensor(qx).cuda(), torch.tensor(t_ids).cuda(),
        torch.tensor(icl).cuda(), 0, torch.tensor(s_att).cuda(), variant, img)
    assert np.array_equal(ktr.cpu().numpy(), want_ktr)                  # bit-exact
    assert np.array_equal(lab.cpu().numpy(), want_lab)
    c = counts.cpu().numpy()
    assert c[0] == (want_ktr >= 0).sum() and c[1] == (want_lab != -100).sum() and c[2] == c[0]
    # compact form (f1): kl_tea_row holds ranks, tea_sel the flat teacher rows in rank order
    k3, _, c3, sel = ops.kd_prepare_rows(
        torch.tensor(s_ids).cuda(), torch.tensor(qx).cuda(), torch.tensor(t_ids).cuda(),
        torch.tensor(icl).cuda(), 0, torch.tensor(s_att).cuda(), variant, img, compact_teacher=True)
    k3, sel = k3.cpu().numpy(), sel.cpu().numpy()
    n = int(c3[0])
    assert sel.shape == (B * Tq,) and np.array_equal(sel[:n], np.sort(want_ktr[want_ktr >= 0]))
    assert not sel[n:].any()
    assert np.array_equal(k3 >= 0, want_ktr >= 0)
    assert np.array_equal(sel[k3[k3 >= 0]], want_ktr[want_ktr >= 0])
    # no CE wanted: labels are not produced
    ktr2, lab2, _ = ops.kd_prepare_rows(
        torch.tensor(s_ids).cuda(), torch.tensor(qx).cuda(), torch.tensor(t_ids).cuda(),
        torch.tensor(icl).cuda(), 0, None, variant, img, want_ce=False)
    assert lab2 is None and np.array_equal(ktr2.cpu().numpy(), want_ktr)


def test_prepare_rows_mismatched_masks_reports_counts(ops):
    """The reference raises a shape error when the two masks select different row counts
    (icv_module.py:126-131); the kernel reports both counts and the host side raises."""
    s_ids = torch.randint(3, 50, (2, 6)).cuda()
    t_ids = torch.randint(3, 50, (2, 9)).cuda()
    _, _, counts = ops.kd_prepare_rows(s_ids, torch.tensor([2, 2]).cuda(), t_ids,
                                       torch.tensor([2, 2]).cuda(), 0)
    c = counts.cpu().numpy()
    assert c[0] == 8 and c[2] == 14


@pytest.mark.parametrize("n_vec,n_alpha,clip", [(32 * 4096, 32, 1.0), (1000, 7, 0.0), (64, 0, 0.05)])
def test_adamw_step_vs_oracle(ops, n_vec, n_alpha, clip):
    rng = np.random.default_rng(n_vec)
    n = n_vec + n_alpha
    p = rng.normal(size=n)
    m = np.zeros(n)
    v = np.zeros(n)
    dp = torch.tensor(p, dtype=torch.float32).cuda()
    dm = torch.zeros(n, device="cuda")
    dv = torch.zeros(n, device="cuda")
    norm = torch.zeros(1, device="cuda")
    lr = np.concatenate([np.full(n_vec, 1e-3), np.full(n_alpha, 1e-1)])
    world = 4.0
    for step in range(1, 6):
        g = rng.normal(size=n) * (10.0 if step == 2 else 0.3)
        dg = torch.tensor(g, dtype=torch.float32).cuda()
        gg = g / world                                      # grad_prescale = 1 / world
        coef, tot = (1.0, np.linalg.norm(gg)) if clip <= 0 else O.clip_coef([gg], clip)
        p, m, v = O.adamw_step(p, gg * coef, m, v, step, lr)
        ops.adamw_step(dp, dg, dm, dv, n_vec, n_alpha, 1e-3, 1e-1, step, grad_prescale=1 / world,
                       max_grad_norm=clip, norm_out=norm)
        assert rel_err(host(norm), [tot]) < 1e-5
        assert rel_err(host(dp), p) < 2e-6
        # fp32 state: the clip coefficient comes from an fp32 sum of 131k squares, v sees it squared
        assert rel_err(host(dm), m) < 1e-5 and rel_err(host(dv), v) < 5e-5


@pytest.mark.parametrize("L,R,d,use_sigmoid,accumulate,train_alpha", [
    (32, 16, 4096, False, False, True),     # the training step's tail
    (2, 1, 512, True, True, True),          # configs[0] shape, sigmoid, gradient accumulation
    (5, 3, 1000, True, False, False),       # odd replica count, d not a multiple of the CTA, alpha frozen
    (3, 9, 8192, False, True, True),        # two sweeps per thread
])
def test_icv_grad_finish_equals_the_separate_tail(ops, L, R, d, use_sigmoid, accumulate, train_alpha):
    """licv_icv_grad_finish = licv_reduce_rows + licv_icv_scale_bwd + the optimizer's sum of squares,
    against float64 (icv_module.py:89-92 differentiated by hand) and, through
    licv_adamw_step_partials, against the two-launch licv_adamw_step on the same gradient."""
    rng = np.random.default_rng(L * 1000 + R)
    rows = rng.normal(size=(L, R, d))
    alpha = rng.normal(size=L)
    vec = rng.normal(size=(L, d)) * 0.01
    old_dv, old_da = rng.normal(size=(L, d)), rng.normal(size=L)
    prescale = 0.25
    f32 = dict(dtype=torch.float32, device="cuda")
    t_rows = torch.tensor(rows, **f32)
    t_alpha, t_vec = torch.tensor(alpha, **f32), torch.tensor(vec, **f32)
    grad = torch.tensor(np.concatenate([old_dv.ravel(), old_da]), **f32)
    d_vec, d_alpha = grad[:L * d].view(L, d), grad[L * d:]
    d_icv = torch.empty(L, d, **f32)
    partials = torch.full((L,), -1.0, **f32)
    rows32 = host(t_rows)                        # the fp32 replicas the kernel sees (it clears them)
    ops.icv_grad_finish(t_rows, t_alpha, t_vec, d_vec, d_alpha if train_alpha else None, d_icv, partials,
                        grad_prescale=prescale, use_sigmoid=use_sigmoid, accumulate=accumulate)
    # the oracle's autograd of a1 + a2 (pinned to the reference's golden in tests/test_oracle_golden.py)
    want_icv = rows32.sum(1)
    a_eff = O.encoder_alpha(host(t_alpha), use_sigmoid)
    d_a_eff, want_dv = O.icv_product_bwd(a_eff, host(t_vec), want_icv)
    want_da = O.encoder_alpha_bwd(host(t_alpha), use_sigmoid, d_a_eff)
    if accumulate:
        want_dv = want_dv + host(torch.tensor(old_dv, **f32))
        want_da = want_da + host(torch.tensor(old_da, **f32))
    assert rel_err(host(d_icv), want_icv) < 1e-6
    assert rel_err(host(d_vec), want_dv) < 1e-6
    if train_alpha:
        assert rel_err(host(d_alpha), want_da) < 2e-5
    else:
        assert np.array_equal(host(d_alpha), host(torch.tensor(old_da, **f32)))     # untouched
    want_sq = prescale ** 2 * (np.square(want_dv).sum(1) + (np.square(want_da) if train_alpha else 0.0))
    assert rel_err(host(partials), want_sq) < 1e-5
    assert float(t_rows.abs().max()) == 0.0      # the replicas are ready for the next pass
    # the optimizer on those partials = the optimizer that sums the squares itself
    n_vec, n_alpha = L * d, (L if train_alpha else 0)
    g_used = grad[:n_vec + n_alpha].clone()
    state = [torch.tensor(rng.normal(size=n_vec + n_alpha), **f32), torch.zeros(n_vec + n_alpha, **f32),
             torch.zeros(n_vec + n_alpha, **f32)]
    state2 = [t.clone() for t in state]
    n1, n2 = torch.zeros(1, **f32), torch.zeros(1, **f32)
    ops.adamw_step(state[0], g_used, state[1], state[2], n_vec, n_alpha, 1e-3, 1e-1, 1,
                   grad_prescale=prescale, max_grad_norm=0.5, norm_out=n1)
    ops.adamw_step(state2[0], g_used, state2[1], state2[2], n_vec, n_alpha, 1e-3, 1e-1, 1,
                   grad_prescale=prescale, max_grad_norm=0.5, norm_out=n2, norm_partials=partials)
    assert rel_err(host(n2), host(n1)) < 1e-6
    for a, b in zip(state, state2):
        assert rel_err(host(b), host(a)) < 1e-6


def test_dp_optimizer_world1_matches_torch_adamw(ops):
    """ICVDataParallelOptimizer (flat views + fused kernel) against torch.optim.AdamW with the
    reference's two parameter groups, clip 1.0 and cosine warm-up (icv_module.py:171-209)."""
    from transformers import get_cosine_schedule_with_warmup

    from licv_vqa_b200 import GlobalICVEncoder
    from licv_vqa_b200.dp import ICVDataParallelOptimizer
    L, d, total = 4, 256, 20
    torch.manual_seed(426)
    enc = GlobalICVEncoder(d, L, alpha_init_value=0.1).cuda()
    ref = GlobalICVEncoder(d, L, alpha_init_value=0.1).cuda()
    ref.load_state_dict(enc.state_dict())
    cfg = dict(icv_lr=1e-3, alpha_lr=1e-2, weight_decay=1e-3, warm_steps=0.1)
    opt = ICVDataParallelOptimizer(enc, cfg, total_steps=total)
    topt = torch.optim.AdamW([{"params": ref.alpha, "lr": 1e-2}, {"params": ref.icv}], lr=1e-3,
                             weight_decay=1e-3)
    sched = get_cosine_schedule_with_warmup(topt, num_warmup_steps=0.1 * total,
                                            num_training_steps=total)
    gen = torch.Generator(device="cuda").manual_seed(3)
    for step in range(8):
        x = torch.randn(1, L, d, device="cuda", generator=gen) * (30.0 if step == 3 else 1.0)
        for e in (enc, ref):
            loss = ((e.scaled_icv() if e is enc else e.alpha.unsqueeze(-1) * e.icv) - x).pow(2).mean()
            loss.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        topt.step()
        sched.step()
        topt.zero_grad()
        logs = opt.step({"loss": loss.detach(), "kl_loss": loss.detach()})
        assert float(logs["loss"]) == pytest.approx(float(loss), rel=1e-6)
        assert rel_err(host(enc.icv), host(ref.icv)) < 5e-6
        assert rel_err(host(enc.alpha), host(ref.alpha)) < 5e-6
        assert enc.icv.grad.data_ptr() == opt.state.grad.data_ptr() and not enc.icv.grad.any()


def test_host_entry_points_match_device_entry_points(ops):
    """licv_*_host (pinned host buffers in, pinned host buffers out) == the device calls."""
    from licv_vqa_b200 import _abi
    lib = _abi.load()
    rng = np.random.default_rng(11)
    n_tok, d, V, R, Rt = 37, 4096, 32002, 9, 5
    dt, code = torch.bfloat16, _abi.BF16
    h = (torch.tensor(rng.normal(size=(n_tok, d)), dtype=torch.float32) * 4).to(dt).pin_memory()
    g = torch.tensor(rng.normal(size=(n_tok, d)), dtype=torch.float32).to(dt).pin_memory()
    s = torch.tensor(rng.normal(size=d), dtype=torch.float32).pin_memory()
    out_h = torch.empty_like(h).pin_memory()
    dh_h = torch.empty_like(h).pin_memory()
    ds_h = torch.empty(d).pin_memory()
    sess = C.c_void_p()
    _abi.check(lib.licv_host_session_create(C.byref(sess), 64 << 20, 3), "create")
    try:
        _abi.check(lib.licv_inject_fwd_host(sess, h.data_ptr(), s.data_ptr(), out_h.data_ptr(), n_tok,
                                            d, code, code, 0), "fwd_host")
        _abi.check(lib.licv_inject_bwd_host(sess, h.data_ptr(), g.data_ptr(), s.data_ptr(),
                                            dh_h.data_ptr(), ds_h.data_ptr(), n_tok, d, code, code, 0),
                   "bwd_host")
        # saved-for-backward pair: h crosses the link once, the backward takes only g
        out2 = torch.empty_like(h).pin_memory()
        dh2 = torch.empty_like(h).pin_memory()
        ds2 = torch.empty(d).pin_memory()
        assert lib.licv_inject_bwd_host_saved(sess, 7, g.data_ptr(), s.data_ptr(), dh2.data_ptr(),
                                              ds2.data_ptr(), n_tok, d, code, code, 0) == -5  # nothing saved
        for key in (7, 7, 8):      # a second save under a live key replaces it
            _abi.check(lib.licv_inject_fwd_host_save(sess, key, h.data_ptr(), s.data_ptr(), out2.data_ptr(),
                                                     n_tok, d, code, code, 0), "fwd_host_save")
        _abi.check(lib.licv_inject_bwd_host_saved(sess, 7, g.data_ptr(), s.data_ptr(), dh2.data_ptr(),
                                                  ds2.data_ptr(), n_tok, d, code, code, 0), "bwd_host_saved")
        # pageable (not pinned) host buffers take the staged path
        h_pg, g_pg = h.clone(), g.clone()
        out3, dh3, ds3 = torch.empty_like(h_pg), torch.empty_like(h_pg), torch.empty(d)
        assert not h_pg.is_pinned()
        _abi.check(lib.licv_inject_fwd_host_save(sess, 9, h_pg.data_ptr(), s.data_ptr(), out3.data_ptr(),
                                                 n_tok, d, code, code, 0), "fwd_host_save pageable")
        _abi.check(lib.licv_inject_bwd_host_saved(sess, 9, g_pg.data_ptr(), s.data_ptr(), dh3.data_ptr(),
                                                  ds3.data_ptr(), n_tok, d, code, code, 0),
                   "bwd_host_saved pageable")
        stu = (torch.tensor(rng.normal(size=(R, V)), dtype=torch.float32) * 3).to(dt).pin_memory()
        tea = (torch.tensor(rng.normal(size=(Rt, V)), dtype=torch.float32) * 3).to(dt).pin_memory()
        ktr = torch.tensor([0, -1, 1, 2, -1, -1, 3, 4, -1], dtype=torch.int32).pin_memory()
        lab = torch.tensor(rng.integers(0, V, size=R)).pin_memory()
        lab[1] = -100
        dstu_h = torch.empty_like(stu).pin_memory()
        loss_h = torch.zeros(3).pin_memory()
        _abi.check(lib.licv_kd_loss_fwd_bwd_host(sess, stu.data_ptr(), dstu_h.data_ptr(), tea.data_ptr(),
                                                 ktr.data_ptr(), lab.data_ptr(), 5, 8, 1.0, 1e-6, 0.5, 0,
                                                 1.0, loss_h.data_ptr(), R, Rt, V, code, 16), "kd_host")
        _abi.check(lib.licv_host_sync(sess), "sync")
    finally:
        lib.licv_host_session_destroy(sess)
    sd = s.cuda()
    out_d = ops.inject_forward(h.cuda(), sd, dt, 0)
    ds_d = torch.zeros(d, device="cuda")
    dh_d = ops.inject_backward(h.cuda(), g.cuda(), sd, ds_d, True, 0)
    assert torch.equal(out_h.cuda(), out_d) and torch.equal(dh_h.cuda(), dh_d)
    assert rel_err(host(ds_h), host(ds_d)) < 1e-6          # atomics: order may differ
    assert torch.equal(out2.cuda(), out_d) and torch.equal(dh2.cuda(), dh_d)
    assert torch.equal(out3.cuda(), out_d) and torch.equal(dh3.cuda(), dh_d)
    assert rel_err(host(ds2), host(ds_d)) < 1e-6 and rel_err(host(ds3), host(ds_d)) < 1e-6
    losses, dstu_d = ops.kd_loss_raw(stu.cuda(), tea.cuda(), ktr.cuda(), lab.cuda(), None, 5, 8, 1.0,
                                     1e-6, 0.5, in_place=False)
    assert torch.equal(dstu_h.cuda(), dstu_d)
    assert np.allclose(host(loss_h), host(losses), rtol=1e-6)
    want = O.kd_loss_rows(host(stu), host(tea), ktr.numpy(), lab.numpy(), 1.0, 1e-6, 0.5)
    assert abs(float(loss_h[2]) - want["loss"]) <= 1e-5 * abs(want["loss"])


def test_no_cpu_path(ops):
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.inject_forward(torch.zeros(2, 8), torch.zeros(8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.kd_loss_raw(torch.zeros(2, 8), torch.zeros(2, 8))


def test_peer_exchange_world1_equals_adamw_step(ops):
    """csrc/licv_dp.cu with a single rank: exchange + optimizer == licv_adamw_step (the multi-rank
    protocol is checked under torchrun by tools/dp_p2p_check.py)."""
    from licv_vqa_b200.dp import PeerExchange
    n_vec, n_alpha, n_extra = 4096 * 3, 3, 5
    n = n_vec + n_alpha + n_extra
    npad = (n + 3) // 4 * 4
    g = torch.Generator(device="cuda").manual_seed(5)
    p0 = torch.randn(n_vec + n_alpha, device="cuda", generator=g)
    grad0 = torch.zeros(npad, device="cuda")
    grad0[:n] = torch.randn(n, device="cuda", generator=g)
    pa, pb = p0.clone(), p0.clone()
    ma, va, mb, vb = (torch.zeros_like(p0) for _ in range(4))
    na, nb = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    wa, wb = torch.zeros(16, dtype=torch.uint8, device="cuda"), torch.zeros(16, dtype=torch.uint8, device="cuda")
    ex = PeerExchange(npad)
    try:
        for step in (1, 2, 3):
            ga = grad0.clone() * step
            gb = ga.clone()
            ex.step(pa, ga, ma, va, n_vec, n_alpha, npad - n_vec - n_alpha, 1e-3, 1e-2, (0.9, 0.999), 1e-8,
                    1e-3, step, 1.0, na, wa)
            ops.adamw_step(pb, gb, mb, vb, n_vec, n_alpha, 1e-3, 1e-2, step, grad_prescale=1.0,
                           max_grad_norm=1.0, norm_out=nb, workspace=wb)
            assert torch.equal(ga, gb)                       # the sum over one rank is the gradient
            assert torch.allclose(na, nb, rtol=1e-5)
            assert rel_err(host(pa), host(pb)) < 1e-6
        assert not ex.timed_out()
    finally:
        ex.close()
