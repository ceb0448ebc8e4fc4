"""Properties of the oracle itself (CPU, float64): its closed-form gradients against central
finite differences of its own forward, and the invariants the GPU parity tests lean on at full
size (norm preservation, vanishing row sums of the loss gradient, linearity of d_shift in g).

The golden vectors (tests/test_oracle_golden.py) pin the oracle to the reference on fixed inputs;
these pin its algebra on random ones.
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import licv_oracle as O

SEEDS = st.integers(min_value=0, max_value=2 ** 31 - 1)


def _fd(f, x, eps=1e-6):
    """Central finite-difference gradient of scalar f at x (x float64 array)."""
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        old = x[i]
        x[i] = old + eps
        fp = f(x)
        x[i] = old - eps
        fm = f(x)
        x[i] = old
        g[i] = (fp - fm) / (2 * eps)
    return g


@settings(max_examples=15, deadline=None)
@given(seed=SEEDS, n_tok=st.integers(1, 4), d=st.sampled_from([4, 8, 12]),
       shift_scale=st.sampled_from([1e-3, 0.3, 3.0]))
def test_inject_preserves_the_token_norm_and_points_along_h_plus_s(seed, n_tok, d, shift_scale):
    rng = np.random.default_rng(seed)
    h = rng.normal(size=(n_tok, d)) * 4
    s = rng.normal(size=d) * shift_scale
    out = O.inject_fwd(h, s)
    np.testing.assert_allclose(np.linalg.norm(out, axis=-1), np.linalg.norm(h, axis=-1), rtol=1e-12)
    y = h + s
    cos = (out * y).sum(-1) / (np.linalg.norm(out, axis=-1) * np.linalg.norm(y, axis=-1))
    np.testing.assert_allclose(cos, 1.0, rtol=1e-12)


@settings(max_examples=10, deadline=None)
@given(seed=SEEDS, n_tok=st.integers(1, 3), d=st.sampled_from([4, 6]))
def test_inject_bwd_is_the_derivative_of_inject_fwd(seed, n_tok, d):
    rng = np.random.default_rng(seed)
    h = rng.normal(size=(n_tok, d)) * 2
    s = rng.normal(size=d) * 0.5
    g = rng.normal(size=(n_tok, d))
    dh, ds = O.inject_bwd(h, s, g)
    fd_h = _fd(lambda x: float((O.inject_fwd(x, s) * g).sum()), h.copy())
    fd_s = _fd(lambda x: float((O.inject_fwd(h, x) * g).sum()), s.copy())
    np.testing.assert_allclose(dh, fd_h, rtol=2e-6, atol=2e-8)
    np.testing.assert_allclose(ds, fd_s, rtol=2e-6, atol=2e-8)


@settings(max_examples=10, deadline=None)
@given(seed=SEEDS)
def test_d_shift_is_linear_in_g_and_additive_over_tokens(seed):
    rng = np.random.default_rng(seed)
    h = rng.normal(size=(5, 8)) * 3
    s = rng.normal(size=8)
    g1, g2 = rng.normal(size=(5, 8)), rng.normal(size=(5, 8))
    _, a = O.inject_bwd(h, s, g1)
    _, b = O.inject_bwd(h, s, g2)
    _, c = O.inject_bwd(h, s, 2.0 * g1 - 0.5 * g2)
    np.testing.assert_allclose(c, 2.0 * a - 0.5 * b, rtol=1e-10, atol=1e-12)
    _, head = O.inject_bwd(h[:2], s, g1[:2])
    _, tail = O.inject_bwd(h[2:], s, g1[2:])
    np.testing.assert_allclose(head + tail, a, rtol=1e-10, atol=1e-12)


@settings(max_examples=10, deadline=None)
@given(seed=SEEDS, T=st.sampled_from([0.5, 1.0, 2.0]), eps=st.sampled_from([1e-6, 1e-3]))
def test_kl_gradient_is_the_derivative_and_its_rows_sum_to_zero(seed, T, eps):
    rng = np.random.default_rng(seed)
    n, v = 3, 7
    stu, tea = rng.normal(size=(n, v)) * 2, rng.normal(size=(n, v)) * 2
    loss, d = O.kl_divergence(stu, tea, T, eps)
    fd = _fd(lambda x: float(O.kl_divergence(x, tea, T, eps, need_grad=False)[0]), stu.copy())
    np.testing.assert_allclose(d, fd, rtol=5e-6, atol=5e-9)
    # softmax is shift invariant per row -> every row of the gradient sums to zero (the property
    # the full-size GPU test checks on 8192 x 32002 rows)
    np.testing.assert_allclose(d.sum(-1), 0.0, atol=1e-14)
    assert loss >= -1e-12 or eps > 0     # with eps inside the logs the loss is only ~non-negative


@settings(max_examples=10, deadline=None)
@given(seed=SEEDS, lam=st.sampled_from([0.25, 0.5, 1.0]), only_hard=st.booleans())
def test_fused_rows_gradient_is_the_derivative_of_the_fused_loss(seed, lam, only_hard):
    rng = np.random.default_rng(seed)
    R, Rt, V = 5, 4, 6
    stu, tea = rng.normal(size=(R, V)) * 2, rng.normal(size=(Rt, V)) * 2
    ktr = np.array([-1, 2, -1, 0, 3], np.int32)
    lab = np.array([1, -100, 5, 0, -100], np.int64)
    ref = O.kd_loss_rows(stu, tea, ktr, lab, 1.0, 1e-6, lam, only_hard)
    fd = _fd(lambda x: float(O.kd_loss_rows(x, tea, ktr, lab, 1.0, 1e-6, lam, only_hard)["loss"]),
             stu.copy())
    np.testing.assert_allclose(ref["d_stu"], fd, rtol=5e-6, atol=5e-9)
    np.testing.assert_allclose(ref["d_stu"].sum(-1), 0.0, atol=1e-14)
    # a row in neither loss gets a zero gradient (the kernel zero-fills it)
    dead = (ktr < 0) & (lab == -100) if not only_hard else (lab == -100)
    assert np.all(ref["d_stu"][dead] == 0.0)
    assert ref["M"] == int((lab != -100).sum())
    assert ref["N"] == (0 if only_hard else int((ktr >= 0).sum()))


@pytest.mark.parametrize("use_sigmoid", [False, True])
def test_alpha_and_product_backward_are_derivatives(use_sigmoid):
    rng = np.random.default_rng(3)
    L, d = 3, 5
    a_raw, vec, g = rng.normal(size=L), rng.normal(size=(L, d)), rng.normal(size=(L, d))

    def loss(a, v):
        return float((O.icv_product(O.encoder_alpha(a, use_sigmoid), v) * g).sum())

    a_eff = O.encoder_alpha(a_raw, use_sigmoid)
    d_alpha_eff, d_vec = O.icv_product_bwd(a_eff, vec, g)
    d_alpha_raw = O.encoder_alpha_bwd(a_raw, use_sigmoid, d_alpha_eff)
    np.testing.assert_allclose(d_vec, _fd(lambda v: loss(a_raw, v), vec.copy()), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.asarray(d_alpha_raw).reshape(-1),
                               _fd(lambda a: loss(a, vec), a_raw.copy()), rtol=1e-6, atol=1e-9)
