"""CPU: the oracle (oracle/licv_oracle.py) against the golden vectors that the REAL reference
produced (oracle/make_golden.py -> tests/golden/*.npz).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import licv_oracle as O
from tests.util import EPS, load_golden, rel_err


def _cases(fname):
    z = load_golden(fname)
    return z, [str(n) for n in z["names"]]


INJ, INJ_NAMES = _cases("inject_cases.npz")
KL, KL_NAMES = _cases("kl_cases.npz")
MASK, MASK_NAMES = _cases("mask_cases.npz")
ENC, ENC_NAMES = _cases("encoder_cases.npz")


@pytest.mark.parametrize("name", INJ_NAMES)
def test_inject_fwd_bwd_matches_reference(name):
    h, icv, g = INJ[f"{name}/h"], INJ[f"{name}/icv"], INJ[f"{name}/g"]
    layer, _ = INJ[f"{name}/meta"]
    hdt, idt, odt = [str(x) for x in INJ[f"{name}/dtypes"]]
    s = icv[0, layer]
    # "..up" cases handed the reference fp32 tensors holding bf16/fp16 values (= CUDA autocast)
    lowp = hdt if hdt in ("bf16", "fp16") else None
    flags, ref_out_fmt = O.chain_flags(hdt, idt, autocast=False) if lowp else (0, "fp32")
    assert EPS[ref_out_fmt] == EPS[odt]          # the oracle predicts the reference's result dtype
    out = O.inject_fwd(h, s, flags, lowp, out_fmt=ref_out_fmt)
    dh, ds = O.inject_bwd(h, s, g, flags, lowp)
    ref_out = INJ[f"{name}/out"].astype(np.float64)
    if odt == "float32":
        assert rel_err(out, ref_out) < 2e-6
    else:
        # bit-faithful chain: identical except where a fp32 summation-order difference flips a
        # final rounding; never more than one unit in the last place
        ulp = np.abs(ref_out) * 2 * EPS[odt] + 1e-30
        assert np.all(np.abs(out - ref_out) <= ulp)
        assert np.mean(out != ref_out) < 0.02
    ref_ds = INJ[f"{name}/dicv"][0, layer]
    if odt == "float32":
        # ICV gradient: fp32 accumulation in the reference -> the stated 1e-4 with room to spare
        assert rel_err(ds, ref_ds) < 2e-5
        # dh is stored in h's dtype by autograd
        assert rel_err(dh, INJ[f"{name}/dh"]) < (2e-5 if lowp is None else 6 * EPS[hdt])
    else:
        # all-low-precision autograd chain (DeepSpeed recipe): the reference's own gradient is
        # only good to a few units of bf16/fp16 roundoff
        assert rel_err(ds, ref_ds) < 24 * EPS[odt]
        assert rel_err(dh, INJ[f"{name}/dh"]) < 24 * EPS[odt]
    # every other layer's icv row received no gradient
    other = np.delete(INJ[f"{name}/dicv"][0], layer, axis=0)
    assert not other.any()
    if flags == 0:
        # the injection preserves the token norm (icv_intervention.py:68-71)
        np.testing.assert_allclose(np.linalg.norm(out, axis=-1),
                                   np.linalg.norm(h.astype(np.float64), axis=-1), rtol=1e-6)
    # and it is not vacuous: the output moved
    assert rel_err(out, h) > 1e-4


def test_inject_dtype_promotion_recorded():
    """(bf16 h, fp32 icv) -> fp32 out; (bf16, bf16) -> bf16 (SURVEY.md §8a probe)."""
    got = {str(n): [str(x) for x in INJ[f"{n}/dtypes"]] for n in INJ_NAMES}
    assert got["bf16_fp32icv_r0.1"][2] == "float32"
    assert got["bf16_bf16icv_r1"][2] == "bfloat16"
    assert got["fp16_fp32icv_r0.1"][2] == "float32"
    assert got["fp16_fp16icv_r1"][2] == "float16"


@pytest.mark.parametrize("name", KL_NAMES)
def test_kl_matches_reference(name):
    stu, tea = KL[f"{name}/stu"], KL[f"{name}/tea"]
    T, eps = KL[f"{name}/params"]
    dt, odt = [str(x) for x in KL[f"{name}/dtype"]]
    loss, dstu = O.kl_divergence(stu, tea, T, eps)
    if dt in ("fp32", "bf16up"):
        assert abs(loss - KL[f"{name}/loss"]) <= 2e-5 * abs(loss) + 1e-7
        assert rel_err(dstu, KL[f"{name}/dstu"]) < 2e-5
    else:
        # the reference ran softmax/log/sum in bf16/fp16 here (no autocast): loose by design
        assert abs(loss - KL[f"{name}/loss"]) <= 0.08 * abs(loss) + 1e-3
        assert rel_err(dstu, KL[f"{name}/dstu"]) < 0.1


@pytest.mark.parametrize("name", MASK_NAMES)
def test_get_mask_bit_exact(name):
    m = O.get_mask(MASK[f"{name}/ids"], MASK[f"{name}/len"], int(MASK[f"{name}/pad"]))
    assert m.dtype == bool
    assert np.array_equal(m, MASK[f"{name}/mask"])


@pytest.mark.parametrize("name", ENC_NAMES)
def test_encoder_and_product(name):
    L, d, learn, sig = [int(x) for x in ENC[f"{name}/cfg"]]
    a_raw, vec = ENC[f"{name}/alpha_raw"], ENC[f"{name}/vec"]
    a_eff = O.encoder_alpha(a_raw, bool(sig))
    np.testing.assert_allclose(a_eff, ENC[f"{name}/alpha_eff"], rtol=2e-6)
    icv = O.icv_product(a_eff, vec)
    np.testing.assert_allclose(icv, ENC[f"{name}/icv"], rtol=3e-6, atol=1e-9)
    d_a_eff, d_vec = O.icv_product_bwd(a_eff, vec, ENC[f"{name}/g"])
    np.testing.assert_allclose(d_vec, ENC[f"{name}/dvec"], rtol=3e-6, atol=1e-9)
    if learn:
        d_a = O.encoder_alpha_bwd(a_raw, bool(sig), d_a_eff)
        np.testing.assert_allclose(d_a, ENC[f"{name}/dalpha"], rtol=2e-5, atol=1e-7)
    else:
        assert ENC[f"{name}/dalpha"].size == 0 and not bool(ENC[f"{name}/alpha_requires_grad"])
    a0, init_std, init_mean = ENC[f"{name}/init"]
    assert abs(init_mean - a0) < 1e-7           # alpha filled with alpha_init_value
    assert 0.005 < init_std < 0.02              # icv ~ N(0, 0.01^2)


def test_pair_rows_and_gather_agree():
    rng = np.random.default_rng(0)
    B, Tq, Tt, V = 3, 7, 19, 11
    sm = rng.random((B, Tq)) < 0.4
    tm = np.zeros((B, Tt), bool)
    for b in range(B):
        k = int(sm[b].sum())
        tm[b, rng.choice(Tt, k, replace=False)] = True
    stu = rng.normal(size=(B, Tq, V))
    tea = rng.normal(size=(B, Tt, V))
    ktr = O.pair_rows(sm, tm)
    rows = np.flatnonzero(ktr >= 0)
    np.testing.assert_array_equal(stu.reshape(-1, V)[rows], O.gather_rows(stu, sm))
    np.testing.assert_array_equal(tea.reshape(-1, V)[ktr[rows]], O.gather_rows(tea, tm))
    tm[0, :] = True
    with pytest.raises(ValueError):
        O.pair_rows(sm, tm)


def test_kd_loss_rows_composes_kl_and_ce():
    rng = np.random.default_rng(1)
    R, Rt, V = 9, 14, 33
    stu = rng.normal(size=(R, V)) * 2
    tea = rng.normal(size=(Rt, V)) * 2
    ktr = np.array([-1, 3, -1, 0, 13, -1, -1, 7, -1], np.int32)
    lab = np.array([4, 5, -100, 0, 32, 1, -100, 2, -100], np.int64)
    r = O.kd_loss_rows(stu, tea, ktr, lab, temperature=2.0, hard_loss_weight=0.5)
    rows = np.flatnonzero(ktr >= 0)
    kl, dkl = O.kl_divergence(stu[rows], tea[ktr[rows]], 2.0, 1e-6)
    ce, dce, M = O.cross_entropy_rows(stu, lab)
    assert r["N"] == 4 and r["M"] == 6 == M
    np.testing.assert_allclose(r["loss"], kl + 0.5 * ce, rtol=1e-14)
    exp = 0.5 * dce
    exp[rows] += dkl
    np.testing.assert_allclose(r["d_stu"], exp, rtol=1e-13, atol=1e-16)
    # numerical gradient of the total
    eps = 1e-6
    for (i, j) in [(1, 5), (3, 0), (5, 1), (2, 7), (7, 2)]:
        sp = stu.copy(); sp[i, j] += eps
        sm_ = stu.copy(); sm_[i, j] -= eps
        num = (O.kd_loss_rows(sp, tea, ktr, lab, 2.0, hard_loss_weight=0.5)["loss"]
               - O.kd_loss_rows(sm_, tea, ktr, lab, 2.0, hard_loss_weight=0.5)["loss"]) / (2 * eps)
        assert abs(num - r["d_stu"][i, j]) < 1e-7
    # hard_loss_weight = 0 switches the CE term off entirely; only_hard_loss returns ce alone
    r0 = O.kd_loss_rows(stu, tea, ktr, lab, 2.0, hard_loss_weight=0.0)
    assert r0["ce"] == 0.0 and r0["M"] == 0 and np.isclose(r0["loss"], kl)
    r1 = O.kd_loss_rows(stu, tea, ktr, lab, 2.0, hard_loss_weight=0.5, only_hard_loss=True)
    np.testing.assert_allclose(r1["loss"], ce, rtol=1e-14)
    np.testing.assert_allclose(r1["d_stu"], dce, rtol=1e-13, atol=1e-16)


def test_ce_label_variants_against_torch():
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(2)
    B, T, V = 3, 8, 17
    ids = rng.integers(3, V, size=(B, T))
    att = np.ones((B, T), np.int64)
    ids[1, 5:] = 0; att[1, 5:] = 0
    ids[2, :2] = 0; att[2, :2] = 0
    ids[0, 3] = 9  # the "image token"
    logits = rng.normal(size=(B, T, V))
    tl, ti, ta = torch.tensor(logits), torch.tensor(ids), torch.tensor(att)
    keep = ta[..., 1:] != 0
    want = {
        "idefics": F.cross_entropy(tl[..., :-1, :][keep], ti[..., 1:][keep]),
        "idefics2": F.cross_entropy(tl[..., :-1, :][keep], ti[..., 1:][keep], ignore_index=9),
        "causal_lm": F.cross_entropy(tl[..., :-1, :].reshape(-1, V), ti[..., 1:].reshape(-1)),
    }
    for variant, w in want.items():
        lab = O.ce_labels(ids, att, variant, image_token_id=9)
        loss, d, M = O.cross_entropy_rows(logits.reshape(-1, V), lab.reshape(-1))
        np.testing.assert_allclose(loss, float(w), rtol=1e-12)
        assert (lab[:, -1] == -100).all()


def test_inject_nan_when_shift_cancels_token():
    """no eps in the norm (icv_intervention.py:70): h + s = 0 -> NaN, like the reference."""
    h = np.array([[1.0, -2.0, 3.0], [0.5, 0.5, 0.5]])
    s = np.array([-1.0, 2.0, -3.0])
    out = O.inject_fwd(h, s)
    assert np.isnan(out[0]).all() and np.isfinite(out[1]).all()


def test_oracle_temperature_gradient_matches_reference_golden():
    """d loss / d T of the oracle against the reference's autograd (learnable_t, icv_module.py:49-52)."""
    G = load_golden("kl_dtemp_cases.npz")
    for name in [str(n) for n in G["names"]]:
        T, eps = G[f"{name}/params"]
        loss, d_stu, d_t = O.kl_divergence(G[f"{name}/stu"], G[f"{name}/tea"], T, eps, want_dtemp=True)
        assert abs(loss - float(G[f"{name}/loss"])) <= 1e-5 * abs(loss)
        assert abs(d_t - float(G[f"{name}/dtemp"])) <= 2e-5 * abs(d_t) + 1e-7
        assert rel_err(d_stu, G[f"{name}/dstu"]) < 1e-5
