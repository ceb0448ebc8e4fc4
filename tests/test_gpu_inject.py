"""GPU parity: the injection kernels (through the C ABI) against the oracle and the golden vectors
of the real reference.  Tolerances (stated by BASELINE.json north_star):
  * injected hidden states: 1e-3 relative (||a-b||/||b||) for bf16/fp16 results; here the kernel
    restates where the reference's chain rounds, so low-precision results are required to be
    within ONE unit in the last place elementwise, and fp32 results within 2e-6 relative;
  * ICV gradient (d_shift) and dh: 1e-4 relative with fp32 accumulation (measured ~1e-6).
"""
import numpy as np
import pytest
import torch

from oracle import licv_oracle as O
from tests.util import EPS, load_golden, rel_err

pytestmark = pytest.mark.gpu

TD = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16,
      "float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}
NAME = {torch.float32: "fp32", torch.bfloat16: "bf16", torch.float16: "fp16"}


@pytest.fixture(scope="module")
def ops():
    from licv_vqa_b200 import ops as _ops
    return _ops


def dev(a, dtype):
    return torch.tensor(np.asarray(a), dtype=torch.float32).to(dtype).cuda()


def host(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


INJ = load_golden("inject_cases.npz")
INJ_NAMES = [str(n) for n in INJ["names"]]


@pytest.mark.parametrize("name", INJ_NAMES)
def test_inject_matches_reference_golden(ops, name):
    hdt, idt, odt = [str(x) for x in INJ[f"{name}/dtypes"]]
    layer, _ = INJ[f"{name}/meta"]
    up = hdt.endswith("up")           # fp32 tensors holding bf16/fp16 values == CUDA autocast
    h_t = TD["fp32"] if up else TD[hdt]
    h = dev(INJ[f"{name}/h"], h_t)
    icv = INJ[f"{name}/icv"]
    shift = dev(icv[0, layer], torch.float32)
    flags, ref_dtype = ops.reference_rounding(h_t, TD[idt], autocast=False)
    assert ref_dtype == TD[odt]
    out = ops.inject_forward(h, shift, ref_dtype, flags)
    assert out.dtype == TD[odt] and out.shape == h.shape
    ref = INJ[f"{name}/out"].astype(np.float64)
    if ref_dtype == torch.float32:
        assert rel_err(host(out), ref) < 2e-6
    else:
        ulp = np.abs(ref) * 2 * EPS[odt] + 1e-30
        assert np.all(np.abs(host(out) - ref) <= ulp)
        assert np.mean(host(out) != ref) < 0.02
    assert rel_err(host(out), ref) < 1e-3          # the stated tolerance, for the record
    # backward: g in the forward's output dtype
    g = dev(INJ[f"{name}/g"], ref_dtype)
    ds = torch.zeros_like(shift)
    dh = ops.inject_backward(h, g, shift, ds, True, flags)
    o_dh, o_ds = O.inject_bwd(INJ[f"{name}/h"], icv[0, layer], INJ[f"{name}/g"], flags,
                              None if up or hdt == "fp32" else hdt)
    lowp = h_t != torch.float32
    assert rel_err(host(ds), o_ds) < 1e-5                       # vs oracle, fp32 accumulation
    assert rel_err(host(dh), o_dh) < (1e-5 if not lowp else 1.2 * EPS[NAME[h_t]])
    if ref_dtype == torch.float32:
        assert rel_err(host(ds), INJ[f"{name}/dicv"][0, layer]) < 1e-4   # vs the reference itself
        assert rel_err(host(dh), INJ[f"{name}/dh"]) < (1e-4 if not lowp else 6 * EPS[NAME[h_t]])


SHAPES = [  # n_tok, d
    (1, 4096), (3, 4096), (257, 4096), (2500, 4096), (5, 512), (33, 520), (7, 72), (64, 8192),
    (19, 1024), (9, 16384), (4, 8), (300, 2048),
]


@pytest.mark.parametrize("n_tok,d", SHAPES)
@pytest.mark.parametrize("hdt,odt", [("bf16", "bf16"), ("bf16", "fp32"), ("fp16", "fp16"),
                                     ("fp16", "fp32"), ("fp32", "fp32")])
def test_inject_random_vs_oracle(ops, n_tok, d, hdt, odt):
    if hdt == "fp32" and d > 8192:
        pytest.skip("row > 32 KB: documented limit")
    rng = np.random.default_rng(n_tok * 131 + d)
    sigma = rng.uniform(1, 30)
    h = dev(rng.normal(size=(n_tok, d)) * sigma, TD[hdt])
    hn = host(h)
    for ratio in (1e-3, 0.1, 1.0, 10.0):
        s = rng.normal(size=d)
        s = s / np.linalg.norm(s) * np.linalg.norm(hn, axis=-1).mean() * ratio
        shift = dev(s, torch.float32)
        sn = host(shift)
        out = ops.inject_forward(h, shift, TD[odt], 0)
        ref = O.inject_fwd(hn, sn, out_fmt=odt)
        if odt == "fp32":
            assert rel_err(host(out), ref) < 2e-6
        else:
            assert np.all(np.abs(host(out) - ref) <= np.abs(ref) * 2 * EPS[odt] + 1e-30)
        # not vacuous: compare the MOVE as well (out - h), SURVEY.md §7 "vacuous parity"
        if ratio >= 0.1:
            assert rel_err(host(out) - hn, ref - hn) < (1e-4 if odt == "fp32" else 0.05)
        g = dev(rng.normal(size=(n_tok, d)), TD[odt])
        ds = torch.zeros(d, device="cuda")
        dh = ops.inject_backward(h, g, shift, ds, True, 0)
        o_dh, o_ds = O.inject_bwd(hn, sn, host(g))
        assert rel_err(host(ds), o_ds) < 1e-5
        assert rel_err(host(dh), o_dh) < (1e-5 if hdt == "fp32" else 1.2 * EPS[hdt])
        # d_shift accumulates (+=): a second call doubles it; dh may be skipped
        assert ops.inject_backward(h, g, shift, ds, False, 0) is None
        assert rel_err(host(ds), 2 * o_ds) < 1e-5


@pytest.mark.parametrize("flags_case", ["mixed_noautocast", "lowp_chain", "lowp_autocast"])
@pytest.mark.parametrize("hdt", ["bf16", "fp16"])
def test_inject_rounding_chains(ops, flags_case, hdt):
    """Every place the reference's eager chain rounds, at d=4096 (d_shift tolerance 1e-4)."""
    rng = np.random.default_rng(7)
    n_tok, d = 40, 4096
    h = dev(rng.normal(size=(n_tok, d)) * 4, TD[hdt])
    s = dev(rng.normal(size=d) * 2, TD[hdt]).float()   # representable in the low precision too
    icv_dtype = torch.float32 if flags_case == "mixed_noautocast" else TD[hdt]
    autocast = flags_case == "lowp_autocast"
    flags, ref_dtype = ops.reference_rounding(TD[hdt], icv_dtype, autocast=autocast)
    o_flags, o_fmt = O.chain_flags(hdt, NAME[icv_dtype], autocast=autocast)
    assert flags == o_flags and NAME[ref_dtype] == o_fmt
    out = ops.inject_forward(h, s, ref_dtype, flags)
    ref = O.inject_fwd(host(h), host(s), flags, hdt, out_fmt=o_fmt)
    if ref_dtype == torch.float32:
        assert rel_err(host(out), ref) < 2e-6
    else:
        bad = np.abs(host(out) - ref) > 0
        assert np.all(np.abs(host(out) - ref) <= np.abs(ref) * 2 * EPS[hdt] + 1e-30)
        assert bad.mean() < 0.02
    g = dev(rng.normal(size=(n_tok, d)), ref_dtype)
    ds = torch.zeros(d, device="cuda")
    dh = ops.inject_backward(h, g, s, ds, True, flags)
    o_dh, o_ds = O.inject_bwd(host(h), host(s), host(g), flags, hdt)
    assert rel_err(host(ds), o_ds) < 1e-4
    assert rel_err(host(dh), o_dh) < 1.2 * EPS[hdt]


def test_inject_nan_like_reference(ops):
    h = torch.randn(4, 512, device="cuda")
    shift = -h[1].clone()
    out = ops.inject_forward(h, shift, torch.float32, 0)
    assert torch.isnan(out[1]).all()
    assert torch.isfinite(out[[0, 2, 3]]).all()


def test_inject_empty_and_errors(ops):
    from licv_vqa_b200 import _abi
    lib = _abi.load()
    s = torch.zeros(512, device="cuda")
    out = ops.inject_forward(torch.zeros(0, 512, device="cuda"), s)
    assert out.shape == (0, 512)
    h = torch.zeros(4, 512, device="cuda")
    o = torch.empty_like(h)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.licv_inject_fwd(h.data_ptr(), s.data_ptr(), o.data_ptr(), 4, 512, 7, 7, 0, st) == -2
    assert lib.licv_inject_fwd(h.data_ptr(), s.data_ptr(), o.data_ptr(), 4, 510, 0, 0, 0, st) == -3
    assert lib.licv_inject_fwd(0, s.data_ptr(), o.data_ptr(), 4, 512, 0, 0, 0, st) == -1
    assert lib.licv_inject_fwd(h.data_ptr() + 4, s.data_ptr(), o.data_ptr(), 3, 512, 0, 0, 0, st) == -4
    assert lib.licv_inject_fwd(h.data_ptr(), s.data_ptr(), h.data_ptr(), 4, 512, 0, 0, 0, st) == -5
    assert lib.licv_inject_fwd(h.data_ptr(), s.data_ptr(), o.data_ptr(), 4, 512, 1, 2, 0, st) == -2
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.inject_forward(torch.zeros(2, 8), torch.zeros(8))
    with pytest.raises(TypeError):
        ops.inject_forward(torch.zeros(2, 8, device="cuda", dtype=torch.float64), s[:8])


def test_inject_full_size_properties(ops):
    """Config-5 size (64 x 2048 x 4096 bf16 = 1 GiB in): size-independent properties."""
    torch.manual_seed(0)
    n_tok, d = 64 * 2048, 4096
    h = (torch.randn(n_tok, d, device="cuda") * 5).to(torch.bfloat16)
    shift = torch.randn(d, device="cuda") * 3
    out = ops.inject_forward(h, shift, torch.bfloat16, 0)
    # 1. token norms are preserved (to bf16 rounding of the elements)
    nh = h.float().norm(dim=-1)
    no = out.float().norm(dim=-1)
    assert torch.max((no - nh).abs() / nh) < 2e-3
    # 2. direction is that of h + s
    idx = torch.randint(0, n_tok, (512,), device="cuda")
    y = h[idx].float() + shift
    cos = torch.nn.functional.cosine_similarity(out[idx].float(), y, dim=-1)
    assert torch.min(cos) > 0.9999
    # 3. a sample of rows against the oracle, exactly as in the small tests
    pick = idx[:64].cpu().numpy()
    ref = O.inject_fwd(host(h[idx[:64]]), host(shift), out_fmt="bf16")
    got = host(out[idx[:64]])
    assert np.all(np.abs(got - ref) <= np.abs(ref) * 2 * EPS["bf16"] + 1e-30)
    # 4. backward: linear in g, and d_shift is the token sum of g_y (checked on a slice with the
    #    oracle and on the whole with linearity)
    g1 = torch.randn(n_tok, d, device="cuda").to(torch.bfloat16)
    ds1 = torch.zeros(d, device="cuda")
    dh1 = ops.inject_backward(h, g1, shift, ds1, True, 0)
    ds2 = torch.zeros(d, device="cuda")
    g2 = (g1.float() * 2).to(torch.bfloat16)
    dh2 = ops.inject_backward(h, g2, shift, ds2, True, 0)
    assert rel_err(host(ds2), 2 * host(ds1)) < 1e-5
    assert rel_err(host(dh2[:4096]), 2 * host(dh1[:4096])) < 1.2 * EPS["bf16"]
    # g_y is orthogonal to y (rescaling kills the radial direction): <ds-contribution, y> = 0
    sl = slice(1000, 1000 + 257)
    ds_s = torch.zeros(d, device="cuda")
    dh_s = ops.inject_backward(h[sl], g1[sl], shift, ds_s, True, 0)
    o_dh, o_ds = O.inject_bwd(host(h[sl]), host(shift), host(g1[sl]))
    assert rel_err(host(ds_s), o_ds) < 1e-5
    assert rel_err(host(dh_s), o_dh) < 1.2 * EPS["bf16"]
    # whole-tensor d_shift equals the sum of slice d_shifts (fp32 atomics: order differs)
    acc = torch.zeros(d, device="cuda")
    step = n_tok // 8
    for i in range(8):
        ops.inject_backward(h[i * step:(i + 1) * step], g1[i * step:(i + 1) * step], shift, acc,
                            False, 0)
    assert rel_err(host(acc), host(ds1)) < 1e-4


# BASELINE.json configs 2-5 at their own shapes: the fp16 DeepSpeed recipe of
# idefics-9B, the bf16 MLP-output hook of idefics2 (1 and 5 image crops), and the inference sweep
# (beams x batch x T, prefill and single-token decode)
CONFIG_SHAPES = [
    ("cfg2_idefics9b_fp16_ds", 8 * 32, "fp16", "lowp_chain"),
    ("cfg3_idefics2_bf16_1crop", 8 * 96, "bf16", "mixed_autocast"),
    ("cfg3_idefics2_bf16_5crops", 8 * 352, "bf16", "mixed_autocast"),
    ("cfg5_decode_bs1_beams3", 3, "bf16", "none"),
    ("cfg5_decode_bs64_beams3", 192, "bf16", "none"),
    ("cfg5_prefill_bs8_beams3_t128", 24 * 128, "bf16", "none"),
    ("cfg5_prefill_bs2_beams3_t2048", 6 * 2048, "fp16", "none"),
]


@pytest.mark.parametrize("name,n_tok,hdt,chain", CONFIG_SHAPES)
def test_inject_baseline_config_shapes(ops, name, n_tok, hdt, chain):
    from licv_vqa_b200 import _abi
    d = 4096
    rng = np.random.default_rng(len(name) * 31 + n_tok)
    sigma = rng.uniform(1, 30, size=(n_tok, 1))
    hn = rng.normal(size=(n_tok, d)) * sigma
    hn[:, [7, 2041]] *= 100.0                      # LLaMA-style massive-activation channels
    h = dev(hn, TD[hdt])
    hn = host(h)
    s = rng.normal(size=d)
    s = s / np.linalg.norm(s) * np.linalg.norm(hn, axis=-1).mean() * 0.1
    flags = {"none": 0,
             "lowp_chain": _abi.ROUND_Y | _abi.ROUND_NH | _abi.ROUND_NY | _abi.ROUND_T,
             "mixed_autocast": 0}[chain]
    if chain == "lowp_chain":
        s = host(dev(s, TD[hdt]))                  # the ICV itself is low precision in this recipe
    odt = "fp32" if chain == "mixed_autocast" else hdt
    shift = dev(s, torch.float32)
    out = ops.inject_forward(h, shift, TD[odt], flags)
    ref = O.inject_fwd(hn, host(shift), flags, hdt if flags else None, out_fmt=odt)
    if odt == "fp32":
        assert rel_err(host(out), ref) < 2e-6
    else:
        assert np.all(np.abs(host(out) - ref) <= np.abs(ref) * 2 * EPS[odt] + 1e-30)
    assert rel_err(host(out), ref) < 1e-3                      # north_star's stated tolerance
    g = dev(rng.normal(size=(n_tok, d)), TD[odt])
    ds = torch.zeros(d, device="cuda")
    dh = ops.inject_backward(h, g, shift, ds, True, flags)
    o_dh, o_ds = O.inject_bwd(hn, host(shift), host(g), flags, hdt if flags else None)
    assert rel_err(host(ds), o_ds) < 1e-4                      # stated: 1e-4 with fp32 accumulation
    assert rel_err(host(dh), o_dh) < 1.2 * EPS[hdt]


@pytest.mark.parametrize("n_tok,d,hdt", [(256, 4096, "fp16"), (256, 4096, "bf16"), (24, 4096, "bf16"),
                                         (1000, 2048, "fp32"), (5000, 4096, "bf16"), (77, 1536, "bf16")])
def test_inject_backward_spread_replicas_equal_accumulated(ops, n_tok, d, hdt):
    """licv_inject_bwd_spread + licv_reduce_rows against licv_inject_bwd on the same launch: dh
    bit-identical, d_shift equal up to the summation order (fp32, 1e-6)."""
    from licv_vqa_b200 import _abi
    lib = _abi.load()
    torch.manual_seed(n_tok + d)
    dt = TD[hdt]
    h = (torch.randn(n_tok, d, device="cuda") * 4).to(dt)
    g = torch.randn(n_tok, d, device="cuda").to(dt)
    shift = torch.randn(d, device="cuda")
    ds = torch.zeros(d, device="cuda")
    dh = ops.inject_backward(h, g, shift, ds, True, 0)
    code = {"fp32": _abi.F32, "bf16": _abi.BF16, "fp16": _abi.F16}[hdt]
    st = torch.cuda.current_stream().cuda_stream
    R = lib.licv_inject_bwd_rows(n_tok, d, code, code)
    assert 1 <= R <= 16 and R & (R - 1) == 0
    rows = torch.zeros(3, R, d, device="cuda")
    dh2 = torch.empty_like(h)
    _abi.check(lib.licv_inject_bwd_spread(h.data_ptr(), g.data_ptr(), shift.data_ptr(), dh2.data_ptr(),
                                          rows[1].data_ptr(), R, n_tok, d, code, code, 0, st))
    assert torch.equal(dh, dh2)
    assert not rows[0].any() and not rows[2].any()
    if R > 1 and n_tok >= 64:
        assert bool(rows[1, R - 1].any())            # the CTAs really spread over the replicas
    out = torch.zeros(3, d, device="cuda")
    _abi.check(lib.licv_reduce_rows(rows[1].data_ptr(), out[1].data_ptr(), 1, R, R * d, d, 0, 0, st))
    assert rel_err(host(out[1]), host(ds)) < 1e-6
    assert not out[0].any() and not out[2].any()
    # a second launch (another micro-batch, or the same layer again) adds; the reduce can
    # accumulate into its output and leaves the replicas zero for the next pass
    _abi.check(lib.licv_inject_bwd_spread(h.data_ptr(), g.data_ptr(), shift.data_ptr(), 0,
                                          rows[1].data_ptr(), R, n_tok, d, code, code, 0, st))
    _abi.check(lib.licv_reduce_rows(rows[1].data_ptr(), out[1].data_ptr(), 1, R, R * d, d, 1, 1, st))
    assert rel_err(host(out[1]), 3 * host(ds)) < 1e-6
    assert not rows[1].any()
