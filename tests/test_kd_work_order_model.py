"""Model of the work order of the loss kernel's tensor-memory form (csrc/licv_kd_loss_cluster.cu,
`row_of` / `ordered`): rows sorted by cost class (KL / CE only / neither; stable) and dealt to the
G CTAs in snake order.  CPU only: the kernel's own parity tests are tests/test_gpu_kd_loss.py (its
results do not depend on the order); what is checked here is the scheduling claim of DESIGN 3.3(b)
- every row is worked exactly once, and for the training batch of BASELINE configs[1] no SM gets
more than two row-units of work, where dealing rows r, r + G, ... to CTA r gave three."""
import numpy as np
import pytest

COST = {0: 2, 1: 1, 2: 0}          # a KL row costs about two CE-only rows, a row in neither loss ~0


def row_classes(kl_tea_row, ce_label):
    return np.where(kl_tea_row >= 0, 0, np.where(ce_label != -100, 1, 2))


def deal(cls, G, ordered):
    """-> list over CTAs of the rows each works, in the kernel's order."""
    n = len(cls)
    order = np.argsort(cls, kind="stable") if ordered else np.arange(n)
    work = [[] for _ in range(G)]
    for b in range(G):
        it = 0
        while True:
            pos = it * G + ((G - 1 - b) if (ordered and it & 1) else b)
            if pos >= n:
                break
            work[b].append(int(order[pos]))
            it += 1
    return work


def configs1_batch(B=8, T=32, K=4):
    ktr = np.full(B * T, -1, np.int32)
    lab = np.full(B * T, -100, np.int64)
    for b in range(B):
        ktr[b * T + T - K:b * T + T] = np.arange(b * K, (b + 1) * K)
        lab[b * T:b * T + T - 1] = 7
    return ktr, lab


@pytest.mark.parametrize("n,G", [(256, 148), (149, 148), (600, 148), (1024, 148), (40, 148), (300, 7)])
def test_every_row_is_worked_exactly_once(n, G):
    rng = np.random.default_rng(n)
    cls = rng.integers(0, 3, n)
    for ordered in (False, True):
        work = deal(cls, min(G, n), ordered)
        seen = sorted(r for w in work for r in w)
        assert seen == list(range(n))


def test_training_batch_needs_two_row_units_per_sm_instead_of_three():
    ktr, lab = configs1_batch()
    cls = row_classes(ktr, lab)
    assert (cls == 0).sum() == 32 and (cls == 1).sum() == 224
    load = lambda work: max(sum(COST[int(cls[r])] for r in w) for w in work)
    assert load(deal(cls, 148, ordered=False)) == 3       # a KL row and a second row on one SM
    assert load(deal(cls, 148, ordered=True)) == 2        # = ceil(total / SMs): 280 units on 148 SMs
