"""Data-parallel host logic on CPU with `gloo`, world_size 2: the batch shards by
sample, the flat [vec | alpha | scalars] buffer is the ONLY thing exchanged, one all-reduce per
optimizer step, logged scalars ride along, and the optimizer kernel refuses CPU tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from licv_vqa_b200.dp import (FlatICVState, ICVDataParallelOptimizer, N_SCALARS,
                              cosine_warmup_factor, shard_batch)
from licv_vqa_b200.icv_encoder import GlobalICVEncoder
from oracle import licv_oracle as O

L, D = 3, 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_grads(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(1, L, D, generator=g), torch.randn(1, L, generator=g)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(426)                      # same initial parameters on every rank
        enc = GlobalICVEncoder(D, L, alpha_learnable=True, alpha_init_value=0.1)
        opt = ICVDataParallelOptimizer(enc, dict(icv_lr=1e-4, alpha_lr=1e-2, warm_steps=0.1),
                                       total_steps=100)
        st = opt.state
        assert opt.world_size == world
        # parameters / grads are views of the flat buffers
        assert enc.icv.data_ptr() == st.param.data_ptr()
        assert enc.icv.grad.data_ptr() == st.grad.data_ptr()
        assert enc.alpha.grad.data_ptr() == st.grad[st.n_vec:].data_ptr()
        # two micro-batches accumulate through autograd into the flat gradient, no collective
        gv, ga = _rank_grads(rank)
        for _ in range(2):
            ((enc.icv * gv).sum() + (enc.alpha * ga).sum()).backward()
        assert torch.allclose(st.grad[:st.n_vec].view(1, L, D), 2 * gv)
        logs = {"kl_loss": torch.tensor(1.0 + rank), "ce_loss": torch.tensor(10.0 * (rank + 1)),
                "loss": torch.tensor(1.0 + rank + 5.0 * (rank + 1))}
        opt.all_reduce_gradients(logs)               # ONE collective
        np.save(os.path.join(out_dir, f"grad{rank}.npy"), st.grad.numpy())
        synced = opt.synced_logs()
        np.save(os.path.join(out_dir, f"logs{rank}.npy"),
                np.array([float(synced[k]) for k in ("kl_loss", "ce_loss", "loss")]))
        # the optimizer kernel is CUDA only: no silent CPU optimizer
        with pytest.raises(RuntimeError, match="B200"):
            opt.step()
    finally:
        dist.destroy_process_group()


def test_flat_buffer_allreduce_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0 = np.load(tmp_path / "grad0.npy")
    g1 = np.load(tmp_path / "grad1.npy")
    assert np.array_equal(g0, g1)                    # every rank holds the same reduced buffer
    want_v = sum(2 * _rank_grads(r)[0] for r in range(world)).reshape(-1).numpy()
    want_a = sum(2 * _rank_grads(r)[1] for r in range(world)).reshape(-1).numpy()
    n = L * D + L
    assert g0.shape == ((n + N_SCALARS + 3) // 4 * 4,)
    np.testing.assert_allclose(g0[:L * D], want_v, rtol=1e-6)
    np.testing.assert_allclose(g0[L * D:n], want_a, rtol=1e-6)
    # scalars: SUM in the buffer, mean in synced_logs (log_dict(sync_dist=True) semantics)
    np.testing.assert_allclose(g0[n:n + 3], [3.0, 30.0, 18.0], rtol=1e-6)
    np.testing.assert_allclose(np.load(tmp_path / "logs0.npy"), [1.5, 15.0, 9.0], rtol=1e-6)


def test_mean_of_rank_grads_equals_single_process_full_batch():
    """DDP semantics (config 4): all-reduced mean of per-rank mean-loss gradients == gradient of
    the mean over ranks of the per-rank losses; with equal shard sizes that is the full-batch
    mean.  Checked with a quadratic loss through the real flat-buffer bookkeeping."""
    torch.manual_seed(0)
    B, world = 8, 2
    x = torch.randn(B, L, D)

    def loss_of(enc, xs):
        return ((enc.icv * enc.alpha.unsqueeze(-1) - xs) ** 2).mean()

    torch.manual_seed(426)
    ref = GlobalICVEncoder(D, L, alpha_init_value=0.3)
    loss_of(ref, x).backward()
    acc_v = torch.zeros_like(ref.icv)
    acc_a = torch.zeros_like(ref.alpha)
    for r in range(world):
        torch.manual_seed(426)
        enc = GlobalICVEncoder(D, L, alpha_init_value=0.3)
        st = FlatICVState(enc)
        shard = shard_batch({"x": x}, r, world)["x"]
        assert shard.shape[0] == B // world
        loss_of(enc, shard).backward()
        acc_v += st.grad[:st.n_vec].view_as(acc_v)
        acc_a += st.grad[st.n_vec:st.n].view_as(acc_a)
    torch.testing.assert_close(acc_v / world, ref.icv.grad, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(acc_a / world, ref.alpha.grad, rtol=1e-5, atol=1e-7)


def test_shard_batch_nested_and_errors():
    b = {"query_inputs": {"input_ids": torch.arange(12).view(4, 3)}, "len": torch.arange(4), "k": 7}
    s = shard_batch(b, 1, 2)
    assert s["query_inputs"]["input_ids"].tolist() == [[6, 7, 8], [9, 10, 11]]
    assert s["len"].tolist() == [2, 3] and s["k"] == 7
    with pytest.raises(ValueError):
        shard_batch({"x": torch.zeros(5, 2)}, 0, 2)


def test_schedule_matches_oracle_and_frozen_alpha():
    for step in (0, 1, 5, 10, 11, 50, 99, 100):
        assert cosine_warmup_factor(step, 10.0, 100) == pytest.approx(
            O.cosine_warmup_factor(step, 10.0, 100), rel=1e-12, abs=1e-15)
    enc = GlobalICVEncoder(D, L, alpha_learnable=False, alpha_init_value=0.2)
    st = FlatICVState(enc)
    assert not st.alpha_learnable and enc.alpha.grad is None
    assert float(enc.alpha.mean()) == pytest.approx(0.2)
    st.zero_grad()
    assert enc.icv.grad.data_ptr() == st.grad.data_ptr()


def test_rebind_keeps_gradients_that_autograd_put_in_fresh_tensors():
    """`module.zero_grad()` (set_to_none=True by default) detaches the flat views; the next
    backward then creates fresh `.grad` tensors.  The optimizer must add those into the flat
    buffer when it re-attaches its views - not drop them."""
    torch.manual_seed(1)
    enc = GlobalICVEncoder(D, L, alpha_learnable=True, alpha_init_value=0.1)
    st = FlatICVState(enc)
    gv, ga = _rank_grads(0)
    enc.zero_grad(set_to_none=True)                      # what torch users do
    assert enc.icv.grad is None
    ((enc.icv * gv).sum() + (enc.alpha * ga).sum()).backward()
    assert enc.icv.grad.data_ptr() != st.grad.data_ptr()   # autograd made its own tensors
    st.rebind()
    assert enc.icv.grad.data_ptr() == st.grad.data_ptr()
    assert torch.allclose(st.grad[:st.n_vec].view(1, L, D), gv)
    assert torch.allclose(st.grad[st.n_vec:st.n].view(1, L), ga)
    # with the views attached, further backwards accumulate in place
    ((enc.icv * gv).sum() + (enc.alpha * ga).sum()).backward()
    st.rebind()
    assert torch.allclose(st.grad[:st.n_vec].view(1, L, D), 2 * gv)


class _FakeLib:
    """liblicv_b200's peer-exchange set-up calls with a failure injected on one rank."""

    def __init__(self, fail_rank, stage, rank, log):
        self.fail_rank, self.stage, self.rank, self.log = fail_rank, stage, rank, log

    def licv_dp_region_alloc(self, n, region_ref, handle):
        if self.stage == "alloc" and self.rank == self.fail_rank:
            return 2          # cudaErrorMemoryAllocation
        region_ref._obj.value = 0x1000 + self.rank
        return 0

    def licv_dp_comm_create(self, comm_ref, rank, world, region, handles, n):
        if self.stage == "map" and self.rank == self.fail_rank:
            return 101        # an IPC handle that can not be opened
        comm_ref._obj.value = 0x2000 + self.rank
        return 0

    def licv_dp_region_free(self, region):
        self.log.append("freed")
        return 0


def _consensus_worker(rank, world, port, out_dir, stage):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from licv_vqa_b200 import _abi, dp
        log = []
        fake = _FakeLib(1, stage, rank, log)
        _abi_load, _abi_status = _abi.load, _abi.status_string
        _abi.load = lambda *a, **k: fake
        _abi.status_string = lambda rc: f"status {rc}"
        try:
            try:
                dp.PeerExchange(100)
                outcome = "created"
            except RuntimeError as exc:
                outcome = "raised: " + str(exc)
        finally:
            _abi.load, _abi.status_string = _abi_load, _abi_status
        with open(os.path.join(out_dir, f"{stage}{rank}.txt"), "w") as f:
            f.write(outcome + "|" + ",".join(log))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("stage", ["alloc", "map"])
def test_peer_exchange_setup_failure_on_one_rank_raises_on_every_rank(tmp_path, stage):
    """A rank that can not allocate or map its peers must not fall back on its own: every rank
    learns of it (one MIN all-reduce) and raises, so `exchange="auto"` takes the NCCL path on all
    of them - nobody is left spinning on packets that never come.  The region is given back."""
    world, port = 2, _free_port()
    mp.spawn(_consensus_worker, args=(world, port, str(tmp_path), stage), nprocs=world, join=True)
    out = [open(tmp_path / f"{stage}{r}.txt").read() for r in range(world)]
    assert all(o.startswith("raised: peer-memory exchange unavailable") for o in out), out
    assert "another rank failed" in out[0]                     # the healthy rank says why
    assert ("licv_dp_region_alloc" if stage == "alloc" else "licv_dp_comm_create") in out[1]
    if stage == "map":
        assert out[0].endswith("freed") and out[1].endswith("freed")
