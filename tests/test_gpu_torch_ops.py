"""torch.ops.licv.* (licv_vqa_b200/torch_ops.py): the registered operators give the results of the
autograd.Function path, pass torch.library.opcheck, and a torch.compile graph holds them whole.
Replaces the hook body of icv_src/icv_model/icv_intervention.py:61-98 and the loss of
icv_src/icv_module.py:94-134."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    from licv_vqa_b200 import ops, torch_ops  # noqa: F401  (registers torch.ops.licv)
    return ops


def _inputs(dtype=torch.bfloat16, n_tok=96, d=4096, R=24, V=32002):
    g = torch.Generator(device="cuda").manual_seed(5)
    h = torch.randn(n_tok, d, device="cuda", generator=g).to(dtype)
    shift = torch.randn(d, device="cuda", generator=g) * 0.1
    stu = (torch.randn(R, V, device="cuda", generator=g) * 3).to(dtype)
    tea = (torch.randn(R, V, device="cuda", generator=g) * 3).to(dtype)
    lab = torch.randint(0, V, (R,), device="cuda", generator=g)
    return h, shift, stu, tea, lab


def test_registered_inject_equals_autograd_function(T):
    h, shift, *_ = _inputs()
    up = torch.randn_like(h)
    outs = []
    for fn in (lambda a, b: torch.ops.licv.inject(a, b, torch.bfloat16, 0),
               lambda a, b: T.inject(a, b, torch.bfloat16, 0)):
        a = h.clone().requires_grad_(True)
        b = shift.clone().requires_grad_(True)
        out = fn(a, b)
        out.backward(up)
        outs.append((out.detach(), a.grad, b.grad))
    (o0, dh0, ds0), (o1, dh1, ds1) = outs
    assert torch.equal(o0, o1) and torch.equal(dh0, dh1)
    # d_shift is accumulated with fp32 atomics: equal up to the order of the additions
    assert float((ds0 - ds1).abs().max()) <= 1e-5 * float(ds0.abs().max())


def test_registered_kd_loss_equals_raw_kernel(T):
    _, _, stu, tea, lab = _inputs()
    R = stu.shape[0]
    s = stu.clone().requires_grad_(True)
    total, kl, ce, dstu = torch.ops.licv.kd_loss(s, tea, None, lab, None, R, R, 1.0, 1e-6, 0.5, False, 16)
    (total * 2.0).backward()
    losses, want = T.kd_loss_raw(stu.clone(), tea, None, lab, None, R, R, 1.0, 1e-6, 0.5, in_place=False)
    assert torch.equal(torch.stack([kl, ce, total.detach()]), losses[:3])
    assert torch.equal(dstu, want)
    np.testing.assert_allclose(s.grad.float().cpu().numpy(), (want.float() * 2.0).cpu().numpy(), rtol=2 ** -7)
    assert torch.equal(stu, s.detach())          # functional form: the logits are untouched


def test_opcheck(T):
    h, shift, stu, tea, lab = _inputs(n_tok=32, d=2048, R=6, V=1003)
    torch.library.opcheck(torch.ops.licv.inject.default,
                          (h.requires_grad_(True), shift.requires_grad_(True), torch.bfloat16, 0),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    torch.library.opcheck(torch.ops.licv.kd_loss.default,
                          (stu.requires_grad_(True), tea, None, lab, None, 6, 6, 1.0, 1e-6, 0.5, False, 16),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


def test_compiled_graph_holds_the_ops_whole(T):
    """fullgraph=True fails on a graph break: the whole hook + loss step traces through."""
    h, shift, stu, tea, lab = _inputs()
    R = stu.shape[0]

    def step(h, shift, stu):
        out = torch.ops.licv.inject(h, shift, torch.bfloat16, 0)
        total, kl, ce, _ = torch.ops.licv.kd_loss(stu, tea, None, lab, None, R, R, 1.0, 1e-6, 0.5, False, 16)
        return out.float().square().mean() + total

    args = lambda: (h.clone().requires_grad_(True), shift.clone().requires_grad_(True), stu.clone().requires_grad_(True))
    a0 = args()
    want = step(*a0)
    want.backward()
    a1 = args()
    got = torch.compile(step, fullgraph=True, backend="aot_eager")(*a1)
    got.backward()
    assert torch.allclose(got, want, rtol=1e-6)
    for x, y in zip(a0, a1):
        assert torch.allclose(x.grad.float(), y.grad.float(), rtol=1e-5, atol=1e-8)
