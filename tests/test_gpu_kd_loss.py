"""GPU parity: the fused distillation-loss kernel (through the C ABI) against the oracle and the
golden vectors of the real reference.  Tolerance (BASELINE.json north_star): loss and gradients
within 1e-4 relative with fp32 accumulation; d(student logits) is stored in the logits' dtype, so
for bf16/fp16 logits it is compared to the oracle's gradient at that format's roundoff."""
import numpy as np
import pytest
import torch

from oracle import licv_oracle as O
from tests.util import EPS, load_golden, rel_err

pytestmark = pytest.mark.gpu

TD = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


@pytest.fixture(scope="module", params=["auto", "stream", "round1"])
def ops(request):
    """The whole file runs three times: the dispatcher's own choice, the stream kernel wherever it
    can run (the small row counts of these tests would otherwise go to the tensor-memory kernel),
    and the round-1 kernels only."""
    from licv_vqa_b200 import _abi, ops as _ops
    lib = _abi.load()
    lib.licv_debug_set_kd_stream({"auto": 1, "stream": 2, "round1": 0}[request.param])
    yield _ops
    lib.licv_debug_set_kd_stream(-1)


def dev(a, dtype):
    return torch.tensor(np.asarray(a), dtype=torch.float32).to(dtype).cuda()


def host(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


KL = load_golden("kl_cases.npz")
KL_NAMES = [str(n) for n in KL["names"]]


@pytest.mark.parametrize("name", KL_NAMES)
def test_kl_matches_reference_golden(ops, name):
    dt, _ = [str(x) for x in KL[f"{name}/dtype"]]
    T, eps = KL[f"{name}/params"]
    up = dt.endswith("up")
    tdt = torch.float32 if up else TD[dt]
    stu = dev(KL[f"{name}/stu"], tdt)
    tea = dev(KL[f"{name}/tea"], tdt)
    N = stu.shape[0]
    losses, dstu = ops.kd_loss_raw(stu, tea, n_kl=N, temperature=T, kl_eps=eps, in_place=False)
    o_loss, o_d = O.kl_divergence(KL[f"{name}/stu"], KL[f"{name}/tea"], T, eps,
                                  logit_fmt=None if tdt == torch.float32 else dt)
    kl = float(losses[0])
    assert abs(kl - o_loss) <= 1e-5 * abs(o_loss) + 1e-7
    assert float(losses[1]) == 0.0 and float(losses[2]) == pytest.approx(kl, rel=1e-7)
    assert rel_err(host(dstu), o_d) < (1e-5 if tdt == torch.float32 else 1.2 * EPS[dt])
    if tdt == torch.float32:
        # against the reference's own fp32 results: the stated 1e-4
        assert abs(kl - float(KL[f"{name}/loss"])) <= 1e-4 * abs(kl) + 1e-7
        assert rel_err(host(dstu), KL[f"{name}/dstu"]) < 1e-4


CASES = [  # R, Rt, V, dtype, T, lambda
    (9, 14, 33, "fp32", 2.0, 0.5),
    (17, 40, 1003, "fp32", 1.0, 0.5),
    (17, 40, 1003, "bf16", 1.0, 0.5),
    (17, 40, 1003, "bf16", 2.0, 0.5),
    (12, 12, 32000, "bf16", 1.0, 0.5),
    (12, 30, 32002, "bf16", 1.0, 0.5),
    (12, 30, 32002, "fp16", 1.0, 0.0),
    (10, 21, 32003, "bf16", 1.0, 0.5),
    (10, 21, 32003, "fp32", 1.5, 0.25),
    (6, 6, 8, "fp32", 0.5, 1.0),
    (5, 9, 257, "fp16", 2.0, 0.5),
    (40, 64, 50257, "bf16", 1.0, 0.5),
    # KL + CE rows at T != 1 on the 32k vocabularies (decay_temperature makes T != 1 a live state,
    # icv_module.py:150-158): the CE term works on the raw logits next to the tempered KL
    (12, 30, 32002, "bf16", 2.0, 0.5),
    (12, 30, 32002, "fp16", 0.5, 0.5),
    (7, 9, 32003, "bf16", 1.5, 0.25),
    (6, 11, 20011, "bf16", 1.0, 0.5),
]


def make_rows(rng, R, Rt, V):
    ktr = np.full(R, -1, np.int32)
    n_kl = max(1, R // 3)
    rows = np.sort(rng.choice(R, n_kl, replace=False))
    ktr[rows] = np.sort(rng.choice(Rt, n_kl, replace=False)).astype(np.int32)
    lab = rng.integers(0, V, size=R).astype(np.int64)
    lab[rng.random(R) < 0.3] = -100
    lab[rows[0]] = V - 1            # last vocabulary entry as a label
    return ktr, lab


@pytest.mark.parametrize("R,Rt,V,dt,T,lam", CASES)
@pytest.mark.parametrize("in_place", [False, True])
def test_kd_loss_rows_vs_oracle(ops, R, Rt, V, dt, T, lam, in_place):
    rng = np.random.default_rng(R * 7919 + V)
    stu_np = rng.normal(size=(R, V)) * 3
    tea_np = rng.normal(size=(Rt, V)) * 3
    ktr, lab = make_rows(rng, R, Rt, V)
    # a confident teacher and a student that partly agrees (the +10 spike of SURVEY.md §8d)
    for r in np.flatnonzero(ktr >= 0):
        j = rng.integers(0, V)
        tea_np[ktr[r], j] += 10
        if r % 2:
            stu_np[r, j] += 8
    stu = dev(stu_np, TD[dt])
    tea = dev(tea_np, TD[dt])
    stu_h, tea_h = host(stu), host(tea)
    use_ce = lam != 0
    want = O.kd_loss_rows(stu_h, tea_h, ktr, lab, T, 1e-6, lam,
                          logit_fmt=None if dt == "fp32" else dt)
    d_ktr = torch.tensor(ktr).cuda()
    d_lab = torch.tensor(lab).cuda() if use_ce else None
    counts = torch.tensor([want["N"], want["M"], want["N"], 0], dtype=torch.int32).cuda()
    src = stu.clone()
    losses, dstu = ops.kd_loss_raw(src, tea, d_ktr, d_lab, counts, temperature=T, kl_eps=1e-6,
                                   hard_loss_weight=lam, in_place=in_place)
    if in_place:
        assert dstu.data_ptr() == src.data_ptr()
    else:
        assert torch.equal(src, stu)
    kl, ce, tot = [float(x) for x in losses]
    assert abs(kl - want["kl"]) <= 1e-5 * abs(want["kl"]) + 1e-7
    assert abs(ce - want["ce"]) <= 1e-5 * abs(want["ce"]) + 1e-7
    assert abs(tot - want["loss"]) <= 1e-5 * abs(want["loss"]) + 1e-7
    assert rel_err(host(dstu), want["d_stu"]) < (1e-5 if dt == "fp32" else 1.2 * EPS[dt])
    # rows in neither loss get an exactly zero gradient
    dead = (ktr < 0) & ((lab == -100) | (not use_ce))
    assert not host(dstu)[dead].any()
    # host-side counts give the same answer as device counts
    l2, _ = ops.kd_loss_raw(stu.clone(), tea, d_ktr, d_lab, None, want["N"], want["M"], T, 1e-6,
                            lam, in_place=False)
    assert torch.equal(l2, losses)


def test_kd_loss_strided_rows_and_misaligned_pairs(ops):
    """Row strides that put student and teacher rows on different 16-byte phases."""
    rng = np.random.default_rng(5)
    R, Rt, V = 11, 23, 32002
    for dt, pad_s, pad_t in [("bf16", 0, 3), ("bf16", 6, 0), ("fp32", 1, 2), ("fp16", 8, 8)]:
        sbuf = torch.zeros(R, V + pad_s, dtype=TD[dt], device="cuda")
        tbuf = torch.zeros(Rt, V + pad_t, dtype=TD[dt], device="cuda")
        sbuf[:, :V] = dev(rng.normal(size=(R, V)) * 3, TD[dt])
        tbuf[:, :V] = dev(rng.normal(size=(Rt, V)) * 3, TD[dt])
        stu, tea = sbuf[:, :V], tbuf[:, :V]
        ktr, lab = make_rows(rng, R, Rt, V)
        want = O.kd_loss_rows(host(stu), host(tea), ktr, lab, 1.0, 1e-6, 0.5)
        losses, dstu = ops.kd_loss_raw(stu, tea, torch.tensor(ktr).cuda(), torch.tensor(lab).cuda(),
                                       None, want["N"], want["M"], 1.0, 1e-6, 0.5, in_place=False)
        assert abs(float(losses[2]) - want["loss"]) <= 1e-5 * abs(want["loss"])
        assert rel_err(host(dstu), want["d_stu"]) < (1e-5 if dt == "fp32" else 1.2 * EPS[dt])


def test_kd_loss_in_place_keeps_row_padding(ops):
    """Rows that start and end inside 16-byte granules (V = 32002 / 32003, padded strides, a view
    that starts mid-buffer): the kernel reads whole granules at the row ends but must write only
    the row - the sentinels around every row survive an in-place launch - and the bytes it reads
    outside the row must not leak into the result."""
    rng = np.random.default_rng(11)
    for dt, V, pad, lead in [("bf16", 32002, 6, 3), ("fp16", 32003, 5, 1), ("bf16", 32002, 0, 5),
                             ("fp32", 32002, 3, 1)]:
        R = 9
        stride = V + pad
        flat = torch.full((lead + R * stride + 16,), 7.0, dtype=TD[dt], device="cuda")
        tflat = torch.full((lead + R * stride + 16,), float("nan"), dtype=TD[dt], device="cuda")
        stu = flat[lead:lead + R * stride].view(R, stride)[:, :V]
        tea = tflat[lead:lead + R * stride].view(R, stride)[:, :V]
        stu.copy_(dev(rng.normal(size=(R, V)) * 3, TD[dt]))
        tea.copy_(dev(rng.normal(size=(R, V)) * 3, TD[dt]))
        ktr = np.arange(R, dtype=np.int32)
        ktr[4] = -1
        lab = rng.integers(0, V, size=R).astype(np.int64)
        lab[2] = -100
        want = O.kd_loss_rows(host(stu), host(tea), ktr, lab, 1.0, 1e-6, 0.5)
        losses, dstu = ops.kd_loss_raw(stu, tea, torch.tensor(ktr).cuda(), torch.tensor(lab).cuda(),
                                       None, want["N"], want["M"], 1.0, 1e-6, 0.5, in_place=True)
        assert dstu.data_ptr() == stu.data_ptr()
        assert abs(float(losses[2]) - want["loss"]) <= 1e-5 * abs(want["loss"])
        assert rel_err(host(dstu), want["d_stu"]) < (1e-5 if dt == "fp32" else 1.2 * EPS[dt])
        keep = torch.ones_like(flat, dtype=torch.bool)
        keep[lead:lead + R * stride].view(R, stride)[:, :V] = False
        assert bool((flat[keep] == 7.0).all()), "wrote outside a row"


def test_kd_loss_only_hard_and_empty(ops):
    rng = np.random.default_rng(6)
    R, V = 13, 1003
    stu = dev(rng.normal(size=(R, V)) * 3, torch.float32)
    lab = rng.integers(0, V, size=R).astype(np.int64)
    lab[::4] = -100
    want = O.kd_loss_rows(host(stu), np.zeros((1, V)), np.full(R, -1), lab, 1.0, 1e-6, 0.5,
                          only_hard_loss=True)
    losses, dstu = ops.kd_loss_raw(stu, None, None, torch.tensor(lab).cuda(), None, 0, want["M"],
                                   1.0, 1e-6, 0.5, only_hard_loss=True, in_place=False)
    assert abs(float(losses[2]) - want["ce"]) <= 1e-5 * want["ce"]
    assert float(losses[0]) == 0.0
    assert rel_err(host(dstu), want["d_stu"]) < 1e-5
    # no KL rows at all: mean over an empty selection is NaN in torch and here
    tea = dev(rng.normal(size=(2, V)), torch.float32)
    ktr = torch.full((R,), -1, dtype=torch.int32, device="cuda")
    losses, dstu = ops.kd_loss_raw(stu.clone(), tea, ktr, None, None, 0, 0, in_place=False)
    assert np.isnan(float(losses[0])) and not host(dstu).any()
    # zero rows
    losses, _ = ops.kd_loss_raw(stu[:0], tea, None, None, None, 0, 0, in_place=False)
    assert np.isnan(float(losses[0]))


def test_kd_loss_autograd_and_upstream_scale(ops):
    """ops.kd_loss as an autograd op: gradient reaches the producer of the logits, an upstream
    scale (loss / accumulate_grad_batches) is applied, and the reference's drop-in compact form
    calculate_kl_divergence(stu[N,V], tea[N,V]) works."""
    rng = np.random.default_rng(8)
    N, V, dmodel = 6, 1003, 32
    hid = dev(rng.normal(size=(N, dmodel)), torch.float32).requires_grad_(True)
    w = dev(rng.normal(size=(V, dmodel)) * 0.3, torch.float32)
    tea = dev(rng.normal(size=(N, V)) * 2, torch.float32)
    logits = hid @ w.t()
    ref_logits = host(logits)
    total, kl, ce = ops.kd_loss(logits, tea, temperature=2.0)
    (total * 0.25).backward()
    o_loss, o_d = O.kl_divergence(ref_logits, host(tea), 2.0, 1e-6)
    assert abs(float(total) - o_loss) <= 1e-5 * abs(o_loss)
    assert rel_err(host(hid.grad), 0.25 * (o_d @ host(w))) < 1e-4
    assert not kl.requires_grad and float(ce) == 0.0


def test_kd_loss_full_size_properties(ops):
    """2048 x 32002 bf16 rows (config-5 loss sweep size): per-row gradient sums vanish (softmax
    gradients are tangent to the simplex), sampled rows match the oracle."""
    torch.manual_seed(1)
    R, V = 2048, 32002
    stu = (torch.randn(R, V, device="cuda") * 3).to(torch.bfloat16)
    tea = (torch.randn(R, V, device="cuda") * 3).to(torch.bfloat16)
    lab = torch.randint(0, V, (R,), device="cuda")
    keep = stu[:64].clone()
    losses, dstu = ops.kd_loss_raw(stu, tea, None, lab, None, R, R, 1.0, 1e-6, 0.5, in_place=True)
    assert dstu.data_ptr() == stu.data_ptr()
    row_sums = dstu.float().sum(dim=1)
    scale = dstu.float().abs().sum(dim=1)
    assert torch.max(row_sums.abs() / scale) < 2e-2      # bf16 rounding of 32002 terms
    idx = np.arange(0, 64, 7)
    want = O.kd_loss_rows(host(keep[idx]), host(tea[idx]), np.arange(len(idx)),
                          host(lab[idx]).astype(np.int64), 1.0, 1e-6, 0.5)
    # the means' denominators are R here, len(idx) in the oracle call
    got = host(dstu[idx]) * (R / len(idx))
    assert rel_err(got, want["d_stu"]) < 1.2 * EPS["bf16"]
    assert np.isfinite(float(losses[2]))


@pytest.mark.parametrize("dt,V", [("bf16", 32002), ("fp16", 32003)])
def test_kd_loss_several_rows_per_sm_every_sweep_kind(ops, dt, V):
    """600 rows = four per SM: the stream kernel's fused sweeps between neighbouring rows of one SM
    run in all their kinds - KL row after KL row (rows r, r+148, r+296 of the first 400), a CE-only
    row after a KL row, CE-only after CE-only, rows in no loss in between - with labels on the
    first / last element of a row and on both sides of the 4096-element group boundaries.  The
    whole gradient and the losses against the oracle."""
    rng = np.random.default_rng(600 + V)
    R = 600
    stu_np = rng.normal(size=(R, V)) * 3
    tea_np = stu_np[:400] + rng.normal(size=(400, V))
    ktr = np.full(R, -1, np.int32)
    ktr[:400] = np.arange(400, dtype=np.int32)
    lab = rng.integers(0, V, size=R).astype(np.int64)
    lab[0:8] = [0, 1, 7, 8, 4095, 4096, V - 1, V - 2]
    lab[148:152] = [4088, 4089, 28671, 28672]
    lab[450:454] = [0, V - 1, 4096, 8191]
    lab[500:520:3] = -100                   # CE-only rows without a label: rows in no loss
    lab[30:40:2] = -100                     # KL rows without CE
    stu = dev(stu_np, TD[dt])
    tea = dev(tea_np, TD[dt])
    want = O.kd_loss_rows(host(stu), host(tea), ktr, lab, 1.0, 1e-6, 0.5, logit_fmt=dt)
    # means over 400 / 600 rows put most fp16 gradient elements among the subnormals: a loss scale,
    # as fp16 training uses one, keeps the comparison about the kernel and not about the format
    scale = 4096.0 if dt == "fp16" else 1.0
    for in_place in (False, True):
        losses, dstu = ops.kd_loss_raw(stu.clone(), tea, torch.tensor(ktr).cuda(), torch.tensor(lab).cuda(),
                                       None, want["N"], want["M"], 1.0, 1e-6, 0.5, grad_scale=scale,
                                       in_place=in_place)
        kl, ce, tot = [float(x) for x in losses]
        assert abs(kl - want["kl"]) <= 1e-5 * abs(want["kl"]) + 1e-7
        assert abs(ce - want["ce"]) <= 1e-5 * abs(want["ce"]) + 1e-7
        assert abs(tot - want["loss"]) <= 1e-5 * abs(want["loss"]) + 1e-7
        got = host(dstu) / scale
        assert rel_err(got, want["d_stu"]) < 1.2 * EPS[dt]
        # row by row as well: one wrong row would drown in the norm of 600
        for r in (0, 5, 147, 148, 151, 296, 399, 400, 450, 453, 599):
            assert rel_err(got[r], want["d_stu"][r]) < 1.5 * EPS[dt], r
        dead = (ktr < 0) & (lab == -100)
        assert dead.any() and not got[dead].any()


@pytest.mark.parametrize("eps", [1e-10, 1e-4])
def test_kd_loss_other_kl_eps(ops, eps):
    """kl_eps is an option (icv_module.py:121-134).  The stream kernel shares one reciprocal among
    four q+eps only while their product stays a normal number (eps >= 1e-9): a smaller eps takes the
    plain reciprocals, a larger one the shared ones with other magnitudes.  300 rows: two per SM."""
    rng = np.random.default_rng(77)
    R, V = 300, 32002
    stu_np = rng.normal(size=(R, V)) * 3
    tea_np = stu_np + rng.normal(size=(R, V))
    ktr = np.arange(R, dtype=np.int32)
    lab = rng.integers(0, V, size=R).astype(np.int64)
    stu, tea = dev(stu_np, torch.bfloat16), dev(tea_np, torch.bfloat16)
    want = O.kd_loss_rows(host(stu), host(tea), ktr, lab, 1.0, eps, 0.5, logit_fmt="bf16")
    losses, dstu = ops.kd_loss_raw(stu.clone(), tea, torch.tensor(ktr).cuda(), torch.tensor(lab).cuda(), None,
                                   want["N"], want["M"], 1.0, eps, 0.5, in_place=False)
    assert abs(float(losses[0]) - want["kl"]) <= 1e-5 * abs(want["kl"]) + 1e-7
    assert abs(float(losses[2]) - want["loss"]) <= 1e-5 * abs(want["loss"]) + 1e-7
    assert rel_err(host(dstu), want["d_stu"]) < 1.2 * EPS["bf16"]


def test_kd_loss_far_from_first_vector_and_minus_inf(ops):
    """Rows whose large logits sit far (in octaves) above what a thread sees first: the stream
    kernel takes its exponentials relative to the thread's first vector and must rebuild such a
    row with exact maxima; -inf logits (masked vocabulary entries) are legal inputs."""
    rng = np.random.default_rng(21)
    for dt, V in [("bf16", 32002), ("fp16", 32003)]:
        R = 6
        stu_np = rng.normal(size=(R, V)) * 3
        tea_np = rng.normal(size=(R, V)) * 3
        stu_np[0, :4300] -= 200.0            # every thread's first vector is 288 octaves down
        tea_np[1, :4300] -= 150.0
        stu_np[2, :4300] = -np.inf           # first vectors hold nothing finite
        tea_np[2, 100:5000] = -np.inf
        stu_np[3, 7::13] = -np.inf           # scattered masked entries (the regular path)
        tea_np[4, 20000] += 120.0            # one teacher logit 170 octaves above the rest
        stu_np[4, 20000] += 60.0
        stu = dev(stu_np, TD[dt])
        tea = dev(tea_np, TD[dt])
        ktr = np.arange(R, dtype=np.int32)
        lab = rng.integers(5000, V, size=R).astype(np.int64)
        lab[3] = 5008                        # not one of row 3's masked entries
        want = O.kd_loss_rows(host(stu), host(tea), ktr, lab, 1.0, 1e-6, 0.5)
        assert np.isfinite(want["loss"])
        for in_place in (False, True):
            losses, dstu = ops.kd_loss_raw(stu.clone(), tea, torch.tensor(ktr).cuda(),
                                           torch.tensor(lab).cuda(), None, want["N"], want["M"],
                                           1.0, 1e-6, 0.5, in_place=in_place)
            assert abs(float(losses[0]) - want["kl"]) <= 1e-5 * abs(want["kl"]) + 1e-7
            assert abs(float(losses[1]) - want["ce"]) <= 1e-5 * abs(want["ce"]) + 1e-7
            assert rel_err(host(dstu), want["d_stu"]) < 1.2 * EPS[dt]


def test_kd_loss_config1_training_shape_in_place(ops):
    """BASELINE configs[1] (idefics-9B, VQAv2 32-shot, bs 8, fp16-mixed): 256 student rows of
    32002 fp16 logits, 32 KL rows paired with a compact teacher [32, V], 248 CE rows,
    hard_loss_weight 0.5, gradient written in place."""
    rng = np.random.default_rng(426)
    B, Tq, V = 8, 32, 32002
    R = B * Tq
    stu_np = rng.normal(size=(R, V)) * 3
    tea_np = rng.normal(size=(32, V)) * 3
    ktr = np.full(R, -1, np.int32)
    lab = rng.integers(3, V, size=R).astype(np.int64)
    n = 0
    for b in range(B):
        lab[b * Tq + Tq - 1] = -100                     # the shifted CE has no target for the last token
        for t in range(Tq - 5, Tq - 1):                 # 4 answer tokens per sample
            ktr[b * Tq + t] = n
            j = rng.integers(0, V)
            tea_np[n, j] += 10
            if n % 2:
                stu_np[b * Tq + t, j] += 8
            n += 1
    stu = dev(stu_np, torch.float16)
    tea = dev(tea_np, torch.float16)
    want = O.kd_loss_rows(host(stu), host(tea), ktr, lab, 1.0, 1e-6, 0.5, logit_fmt="fp16")
    assert (want["N"], want["M"]) == (32, 248)
    counts = torch.tensor([32, 248, 32, 0], dtype=torch.int32).cuda()
    losses, dstu = ops.kd_loss_raw(stu, tea, torch.tensor(ktr).cuda(), torch.tensor(lab).cuda(),
                                   counts, temperature=1.0, kl_eps=1e-6, hard_loss_weight=0.5,
                                   in_place=True)
    assert dstu.data_ptr() == stu.data_ptr()
    kl, ce, tot = [float(x) for x in losses]
    assert abs(kl - want["kl"]) <= 1e-5 * abs(want["kl"]) + 1e-7
    assert abs(ce - want["ce"]) <= 1e-5 * abs(want["ce"]) + 1e-7
    assert abs(tot - want["loss"]) <= 1e-5 * abs(want["loss"]) + 1e-7
    assert rel_err(host(dstu), want["d_stu"]) < 1.2 * EPS["fp16"]
    dead = (ktr < 0) & (lab == -100)
    assert not host(dstu)[dead].any()



DTEMP = load_golden("kl_dtemp_cases.npz")


@pytest.mark.parametrize("name", [str(n) for n in DTEMP["names"]])
def test_learnable_temperature_gradient_matches_reference(ops, name):
    """learnable_t=True (icv_module.py:49-52): d loss / d T through both in-place divides and the
    T**2 factor, against the reference's own autograd (tests/golden/kl_dtemp_cases.npz) at the
    stated 1e-4 and against the oracle's closed form."""
    T, eps = DTEMP[f"{name}/params"]
    stu = dev(DTEMP[f"{name}/stu"], torch.float32).requires_grad_(True)
    tea = dev(DTEMP[f"{name}/tea"], torch.float32)
    t_param = torch.nn.Parameter(torch.tensor(float(T), device="cuda"))
    logits = stu * 1.0
    total, kl, _ = ops.kd_loss(logits, tea, temperature=t_param, kl_eps=float(eps), in_place=False)
    (total * 3.0).backward()                      # an upstream factor reaches dT as well
    ref_loss, ref_dt = float(DTEMP[f"{name}/loss"]), float(DTEMP[f"{name}/dtemp"])
    assert abs(float(total) - ref_loss) <= 1e-4 * abs(ref_loss)
    assert abs(float(t_param.grad) / 3.0 - ref_dt) <= 1e-4 * abs(ref_dt) + 1e-7
    assert rel_err(host(stu.grad) / 3.0, DTEMP[f"{name}/dstu"]) < 1e-4
    _, _, o_dt = O.kl_divergence(DTEMP[f"{name}/stu"], DTEMP[f"{name}/tea"], T, eps, want_dtemp=True)
    assert abs(float(t_param.grad) / 3.0 - o_dt) <= 2e-5 * abs(o_dt) + 1e-7
