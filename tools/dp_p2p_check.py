"""torchrun check of the peer-memory gradient exchange (csrc/licv_dp.cu) against NCCL:

    python -m torch.distributed.run --nproc-per-node 2 tools/dp_p2p_check.py

Every rank builds the same encoder, accumulates rank-dependent gradients, and steps twice-cloned
optimizers - one with exchange="p2p", one with exchange="nccl" - for several steps, eagerly and
through a CUDA graph of the fused launch; parameters, reduced gradients, norms and logged scalars
must agree, and every rank must hold bit-identical parameters."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licv_vqa_b200 import GlobalICVEncoder  # noqa: E402
from licv_vqa_b200.dp import ICVDataParallelOptimizer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L, d = 32, 4096
    cfg = dict(icv_lr=1e-3, alpha_lr=1e-2, weight_decay=1e-3, warm_steps=0.1)
    encs, opts = [], []
    for mode in ("p2p", "nccl"):
        torch.manual_seed(426)
        enc = GlobalICVEncoder(d, L, alpha_init_value=0.1).cuda()
        encs.append(enc)
        opts.append(ICVDataParallelOptimizer(enc, cfg, total_steps=50, exchange=mode))
    assert opts[0].peer is not None and opts[1].peer is None
    gen = torch.Generator(device="cuda").manual_seed(100 + rank)
    worst = 0.0
    for step in range(6):
        gv = torch.randn(1, L, d, device="cuda", generator=gen) * (20.0 if step == 2 else 0.05)
        ga = torch.randn(1, L, device="cuda", generator=gen)
        logs = {"loss": torch.tensor(1.0 + rank + step, device="cuda"),
                "kl_loss": torch.tensor(0.5 * rank, device="cuda")}
        outs = []
        for enc, opt in zip(encs, opts):
            ((enc.icv * gv).sum() + (enc.alpha * ga).sum()).backward()
            outs.append(opt.step(logs))
        for k in ("loss", "kl_loss"):
            assert abs(float(outs[0][k]) - float(outs[1][k])) < 1e-6, (k, outs)
        want_loss = sum(1.0 + r + step for r in range(world)) / world
        assert abs(float(outs[0]["loss"]) - want_loss) < 1e-6
        assert abs(float(opts[0].grad_norm) - float(opts[1].grad_norm)) <= 1e-5 * float(opts[1].grad_norm)
        for a, b in ((encs[0].icv, encs[1].icv), (encs[0].alpha, encs[1].alpha)):
            err = float((a - b).norm() / b.norm())
            worst = max(worst, err)
            assert err < 2e-6, err
        # every rank holds bit-identical parameters (sums are formed in rank order everywhere)
        mine = encs[0].icv.detach().clone()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(mine, ref)
    assert not opts[0].peer.timed_out()
    # the fused launch replayed from a CUDA graph (the step counter lives in device memory)
    opt, enc = opts[0], encs[0]
    st = opt.state
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        st.grad.fill_(0.01 * (rank + 1))
        opt.step()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        st.grad.fill_(0.01 * (rank + 1))
        with torch.cuda.graph(g, stream=s):
            opt.peer.step(st.param, st.grad, st.exp_avg, st.exp_avg_sq, st.n_vec, st.n_alpha,
                          st.grad.numel() - st.n_vec - st.n_alpha, 1e-3, 1e-2, opt.betas, opt.eps,
                          opt.weight_decay, 10, 1.0, opt.grad_norm, opt._ws)
        for _ in range(5):
            st.grad.fill_(0.01 * (rank + 1))
            g.replay()
        s.synchronize()
        want = 0.01 * sum(r + 1 for r in range(world))
        assert torch.allclose(st.grad[:st.n], torch.full_like(st.grad[:st.n], want), rtol=1e-6)
    assert not opts[0].peer.timed_out()
    dist.barrier()
    if rank == 0:
        print(f"dp p2p check ok: world={world}, worst param rel diff vs nccl {worst:.2e}")
    os._exit(0)


if __name__ == "__main__":
    main()
