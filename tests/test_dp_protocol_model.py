"""Executable model of the slot addressing of the fused gradient exchange (csrc/licv_dp.cu), CPU only.

The CUDA kernels cannot run here; what CAN be checked without a GPU is the protocol's bookkeeping,
which both forms of the kernel share: two parity slots per source rank, 16-byte packets addressed by
(parity, source rank, packet index), a step tag in every packet, and - for the owner form - which
positions carry contributions and which carry sums.  The model runs `world` ranks as cooperative
threads under a random scheduler (every interleaving the memory model allows at packet granularity:
a packet is either absent or complete, csrc/licv_dp.cu header), lets fast ranks run ahead as far as
the protocol permits, and asserts

  * no packet is overwritten while a reader of the step it belongs to still needs it,
  * a reader never accepts a packet of another step,
  * every rank ends every step with the rank-order sum (bit-identical replicas),
  * the owner form writes 2 (world - 1) / world of what the all-to-all form writes.

The GPU evidence for the kernels themselves is the bench line's `checks` at N = 2 / 4 / 8
(exchange vs NCCL, replicas bit-identical; profiles/r2e_*) and tools/dp_p2p_check.py.
"""
import random

import numpy as np
import pytest


class Region:
    """One rank's receive region: [parity][source rank][packet] -> (value, tag); counts writes."""

    def __init__(self, world, n_packets):
        self.val = np.zeros((2, world, n_packets), np.float32)
        self.tag = np.zeros((2, world, n_packets), np.int64)      # regions start zeroed, steps from 1
        self.pending = {}    # (parity, src, i) -> step whose reader has not consumed it yet
        self.writes = 0

    def store(self, parity, src, i, value, step):
        key = (parity, src, i)
        assert key not in self.pending, f"packet {key} of step {self.pending[key]} overwritten by step {step}"
        self.val[parity, src, i] = value
        self.tag[parity, src, i] = step
        self.pending[key] = step
        self.writes += 1

    def try_load(self, parity, src, i, step):
        t = self.tag[parity, src, i]
        assert t <= step, f"a packet of step {t} is visible to a reader of step {step}"
        if t != step:
            return None
        self.pending.pop((parity, src, i), None)
        return self.val[parity, src, i]


def packet_thread(form, rank, world, slice_len, i, step, grads, regions, out):
    """One thread of the exchange kernel for packet i of one step, as a generator that yields
    whenever it has to wait (dp_exchange_kernel / dp_exchange_owner_kernel)."""
    parity = step & 1
    mine = grads[rank][i]
    if form == "all":
        for p in range(world):
            if p != rank:
                regions[p].store(parity, rank, i, mine, step)
        got = {}
        while len(got) < world - 1:
            for p in range(world):
                if p != rank and p not in got:
                    v = regions[rank].try_load(parity, p, i, step)
                    if v is not None:
                        got[p] = v
            if len(got) < world - 1:
                yield
        s = np.float32(0)
        for p in range(world):
            s = np.float32(s + (mine if p == rank else got[p]))
        out[rank][i] = s
        return
    own = i // slice_len
    if own != rank:
        regions[own].store(parity, rank, i, mine, step)          # contribution: to the owner only
        while True:
            v = regions[rank].try_load(parity, own, i, step)      # the owner's sum, from local memory
            if v is not None:
                break
            yield
        out[rank][i] = v
    else:
        got = {}
        while len(got) < world - 1:
            for p in range(world):
                if p != rank and p not in got:
                    v = regions[rank].try_load(parity, p, i, step)
                    if v is not None:
                        got[p] = v
            if len(got) < world - 1:
                yield
        s = np.float32(0)
        for p in range(world):
            s = np.float32(s + (mine if p == rank else got[p]))
        for p in range(world):
            if p != rank:
                regions[p].store(parity, rank, i, s, step)       # the sum: owner's slot, owner's slice
        out[rank][i] = s


def run(form, world, n_packets, steps, seed):
    rng = random.Random(seed)
    nrng = np.random.default_rng(seed)
    slice_len = -(-n_packets // world)
    regions = [Region(world, n_packets) for _ in range(world)]
    step_of = [1] * world                      # the step each rank's kernel is running
    grads = {s: nrng.normal(size=(world, n_packets)).astype(np.float32) for s in range(1, steps + 1)}
    outs = {s: np.full((world, n_packets), np.nan, np.float32) for s in range(1, steps + 1)}

    def kernel(rank):
        s = step_of[rank]
        return [packet_thread(form, rank, world, slice_len, i, s, grads[s], regions, outs[s])
                for i in rng.sample(range(n_packets), n_packets)]

    live = {r: kernel(r) for r in range(world)}
    spins = 0
    while live:
        r = rng.choice(list(live))
        # a biased scheduler: sometimes one rank gets a long burst (it runs ahead of the others)
        for _ in range(rng.choice([1, 1, 1, 8, 64])):
            if r not in live:
                break
            th = rng.choice(live[r])
            try:
                next(th)
                spins += 1
            except StopIteration:
                live[r].remove(th)
            if not live[r]:                    # kernel boundary: the next step's launch may start
                step_of[r] += 1
                if step_of[r] <= steps:
                    live[r] = kernel(r)
                else:
                    del live[r]
        assert spins < 2_000_000, "the model does not make progress (deadlock)"
    for s in range(1, steps + 1):
        want = np.zeros(n_packets, np.float32)
        for p in range(world):
            want = (want + grads[s][p]).astype(np.float32)       # rank order, fp32
        for r in range(world):
            assert np.array_equal(outs[s][r], want), f"step {s} rank {r}: not the rank-order sum"
    return sum(reg.writes for reg in regions)


@pytest.mark.parametrize("form", ["all", "owner"])
@pytest.mark.parametrize("world,n_packets", [(2, 7), (3, 10), (4, 16), (8, 19), (8, 8)])
def test_exchange_protocol_model(form, world, n_packets):
    for seed in range(3):
        run(form, world, n_packets, steps=5, seed=seed)


def test_owner_form_moves_fewer_packets():
    world, n, steps = 8, 64, 2
    all_w = run("all", world, n, steps, 0)
    own_w = run("owner", world, n, steps, 0)
    assert all_w == steps * world * (world - 1) * n
    assert own_w == steps * 2 * (world - 1) * n          # = 2 (world - 1) / world of the above per rank
