"""Data-parallel optimizer step for the ICV parameters: ONE collective per optimizer step.

What the reference gets from Lightning (icv_src/icv_module.py:171-209, config/trainer/ddp.yaml:
`strategy: ddp_find_unused_parameters_true`, `gradient_clip_val: 1.0`, `accumulate_grad_batches`)
and from `self.log_dict(loss_dict, sync_dist=True)` (icv_module.py:163):

  * DDP's bucketed all-reduce (mean) of `icv_encoder.icv.grad` / `icv_encoder.alpha.grad`,
  * a second collective for the logged scalars,
  * `clip_grad_norm_(1.0)`, `torch.optim.AdamW` (or DeepSpeedCPUAdam) with two parameter groups
    (`alpha_lr`, `icv_lr`, `weight_decay`) and `get_cosine_schedule_with_warmup`,

is here one flat fp32 buffer  [ vec (L*d) | alpha (L) | logged scalars (4) ]:

  * the encoder's parameters are VIEWS of the flat parameter buffer and their `.grad`s are views
    of the flat gradient buffer, so autograd accumulates straight into it (gradient accumulation
    over micro-batches costs nothing and no collective),
  * `step()` issues one `all_reduce(SUM)` over the flat buffer - 131 104 floats = 0.5 MB for
    idefics shapes, a latency-bound NVLink/NVSwitch operation; the mean (1/world) is folded into
    the optimizer kernel's `grad_prescale`; the logged scalars ride in the tail,
  * then ONE fused kernel pair (`licv_adamw_step`): global-norm clip + AdamW with the two learning
    rates, lr schedule evaluated on the host (a scalar).

The batch is sharded by sample across ranks (independent samples: no data-path collective); each
rank's loss is a per-rank mean (icv_module.py:132-133), averaged across ranks - DDP semantics.

The collective and the buffer bookkeeping are backend-agnostic (tests run them on CPU tensors with
`gloo`, world_size 2); the optimizer kernel is CUDA only - on a CPU tensor `step()` raises, there
is no CPU optimizer here.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.distributed as dist

N_SCALARS = 4   # kl_loss, ce_loss, loss, spare


def cosine_warmup_factor(step: int, warm_steps: float, total_steps: int, num_cycles: float = 0.5):
    """`transformers.get_cosine_schedule_with_warmup` multiplier at optimizer step `step`."""
    if step < warm_steps:
        return float(step) / float(max(1, warm_steps))
    prog = float(step - warm_steps) / float(max(1, total_steps - warm_steps))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * num_cycles * 2.0 * prog)))


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int):
    """This rank's samples of a global batch (dim 0 split into `world` equal contiguous shards,
    like a DistributedSampler over an already-collated batch)."""
    out = {}
    for k, v in batch.items():
        if isinstance(v, dict):
            out[k] = shard_batch(v, rank, world)
        elif torch.is_tensor(v) and v.dim() > 0:
            n = v.shape[0]
            if n % world:
                raise ValueError(f"batch dimension {n} of '{k}' is not divisible by world size {world}")
            per = n // world
            out[k] = v[rank * per:(rank + 1) * per]
        else:
            out[k] = v
    return out


class FlatICVState:
    """Flat parameter / gradient / moment buffers behind a GlobalICVEncoder."""

    def __init__(self, encoder: torch.nn.Module):
        icv, alpha = encoder.icv, encoder.alpha
        if icv.dtype != torch.float32 or alpha.dtype != torch.float32:
            raise TypeError("the ICV parameters are fp32")
        self.n_vec, self.n_alpha = icv.numel(), alpha.numel()
        n = self.n_vec + self.n_alpha
        dev = icv.device
        self.param = torch.empty(n, dtype=torch.float32, device=dev)
        # rounded up to whole 16-byte vectors (the peer-memory exchange moves float4s)
        self.grad = torch.zeros((n + N_SCALARS + 3) // 4 * 4, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            self.param[:self.n_vec].copy_(icv.reshape(-1))
            self.param[self.n_vec:].copy_(alpha.reshape(-1))
        # parameters and their gradients become views of the flat buffers
        icv.data = self.param[:self.n_vec].view(icv.shape)
        alpha.data = self.param[self.n_vec:].view(alpha.shape)
        icv.grad = self.grad[:self.n_vec].view(icv.shape)
        self.alpha_learnable = bool(alpha.requires_grad)
        if self.alpha_learnable:
            alpha.grad = self.grad[self.n_vec:n].view(alpha.shape)
        self.encoder = encoder

    @property
    def n(self) -> int:
        return self.n_vec + self.n_alpha

    def scalars(self) -> torch.Tensor:
        return self.grad[self.n:self.n + N_SCALARS]

    def rebind(self):
        """Re-attach the gradient views.  After `module.zero_grad()` (set_to_none=True is torch's
        default) autograd has created FRESH `.grad` tensors: what they hold is this step's
        gradient and is added into the flat buffer before the views are put back - dropping it
        would make the next optimizer step a weight-decay-only step."""
        enc = self.encoder
        flat_v = self.grad[:self.n_vec].view(enc.icv.shape)
        g = enc.icv.grad
        if g is None or g.data_ptr() != flat_v.data_ptr():
            if g is not None:
                flat_v.add_(g.to(flat_v.dtype))
            enc.icv.grad = flat_v
        if self.alpha_learnable:
            flat_a = self.grad[self.n_vec:self.n].view(enc.alpha.shape)
            g = enc.alpha.grad
            if g is None or g.data_ptr() != flat_a.data_ptr():
                if g is not None:
                    flat_a.add_(g.to(flat_a.dtype))
                enc.alpha.grad = flat_a

    def zero_grad(self):
        self.grad.zero_()
        self.rebind()


class PeerExchange:
    """The fused gradient exchange + optimizer step over NVLink peer memory (csrc/licv_dp.cu).

    One instance per rank; `world` processes on one node, one GPU each.  Set-up allocates this
    rank's exchange region in the library (cudaMalloc + CUDA IPC handle), all-gathers the 64-byte
    handles over the existing process group, and maps the peers.  After that a step is two kernel
    launches and no collective call."""

    def __init__(self, n_floats: int, group=None):
        """Collective over `group`: raises on EVERY rank if the set-up failed on any of them (one
        rank falling back to NCCL on its own would leave the others spinning on its packets)."""
        import ctypes as C

        from . import _abi
        self._abi, self._C = _abi, C
        self.lib = _abi.load()
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n = int(n_floats)
        self.comm = None
        region, handle = C.c_void_p(), C.create_string_buffer(64)
        failure = None

        def agree(stage):
            """MIN over ranks of 'I am fine so far' - every rank takes the same branch."""
            nonlocal failure
            if self.world == 1:
                ok = failure is None
            else:
                where = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
                flag = torch.tensor([0.0 if failure else 1.0], device=where)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                ok = bool(flag.item() == 1.0)
            if not ok:
                if region:
                    self.lib.licv_dp_region_free(region)
                raise RuntimeError(f"peer-memory exchange unavailable ({stage}): "
                                   f"{failure or 'another rank failed'}")

        rc = self.lib.licv_dp_region_alloc(self.n, C.byref(region), handle)
        if rc != 0:
            failure = f"licv_dp_region_alloc: {_abi.status_string(rc)}"
            region = C.c_void_p()
        agree("region")
        handles = handle.raw
        if self.world > 1:
            mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8)
            if dist.get_backend(group) == "nccl":
                mine = mine.cuda()
            every = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(every, mine, group=group)
            handles = b"".join(bytes(t.cpu().tolist()) for t in every)
        self._handles = C.create_string_buffer(handles, len(handles))
        comm = C.c_void_p()
        rc = self.lib.licv_dp_comm_create(C.byref(comm), self.rank, self.world, region,
                                          self._handles, self.n)
        if rc != 0:
            failure = f"licv_dp_comm_create: {_abi.status_string(rc)}"
        agree("mapping the peers")     # also the barrier: every region is mapped before the first step
        self.comm = comm

    def step(self, param, grad, exp_avg, exp_avg_sq, n_vec, n_alpha, n_extra, lr_vec, lr_alpha,
             betas, eps, weight_decay, step, max_grad_norm, norm_out, workspace):
        self._abi.check(self.lib.licv_dp_allreduce_adamw(
            self.comm, param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
            int(n_vec), int(n_alpha), int(n_extra), float(lr_vec), float(lr_alpha), float(betas[0]),
            float(betas[1]), float(eps), float(weight_decay), int(step), float(max_grad_norm),
            norm_out.data_ptr(), workspace.data_ptr(), torch.cuda.current_stream().cuda_stream),
            "licv_dp_allreduce_adamw")

    def timed_out(self) -> bool:
        """Has a wait for a peer ever timed out?  Synchronises the device (a 4-byte read)."""
        return self.lib.licv_dp_comm_error(self.comm) == 1

    def reset_error(self):
        self._abi.check(self.lib.licv_dp_comm_reset_error(self.comm), "licv_dp_comm_reset_error")

    def close(self):
        if getattr(self, "comm", None) is not None and self.comm:
            self.lib.licv_dp_comm_destroy(self.comm)
            self.comm = None


class ICVDataParallelOptimizer:
    """all-reduce + clip + AdamW + cosine warm-up for the ICV parameters of one rank.

        opt = ICVDataParallelOptimizer(module.icv_encoder, module.module_cfg, total_steps)
        for micro_batches in loader:                       # each rank: its shard of the batch
            for mb in micro_batches:
                loss, logs = module.training_step(mb)
                (loss / len(micro_batches)).backward()     # grads accumulate in the flat buffer
            synced = opt.step(logs)                        # ONE collective, ONE optimizer kernel
    """

    def __init__(self, encoder: torch.nn.Module, module_cfg=None, total_steps: int = 1,
                 max_grad_norm: float = 1.0, process_group=None, betas=(0.9, 0.999),
                 eps: float = 1e-8, exchange: str = "auto", check_every: int = 50):
        """``exchange``: "p2p" = fused exchange + optimizer over NVLink peer memory (one node),
        "nccl" = `all_reduce` then the optimizer kernels, "auto" = p2p for CUDA parameters in a
        multi-rank group when EVERY rank can map its peers (decided collectively), else nccl.
        ``check_every``: the peer exchange's error flag (a peer that never delivered: the
        optimizer update is skipped on the device) is polled every that many steps - a 4-byte
        read that synchronises - and raises."""
        def get(name, default):
            if module_cfg is None:
                return default
            if isinstance(module_cfg, dict):
                return module_cfg.get(name, default)
            return getattr(module_cfg, name, default)

        self.state = FlatICVState(encoder)
        self.icv_lr = float(get("icv_lr", 1e-4))
        self.alpha_lr = float(get("alpha_lr", 1e-2))
        self.weight_decay = float(get("weight_decay", 1e-3))
        warm = get("warm_steps", 0.1)
        self.total_steps = int(total_steps)
        self.warm_steps = warm * total_steps if isinstance(warm, float) else int(warm)
        self.max_grad_norm = float(max_grad_norm)
        self.betas, self.eps = betas, float(eps)
        self.group = process_group
        self.step_no = 0
        dev = self.state.param.device
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._ws = torch.zeros(16, dtype=torch.uint8, device=dev)
        self.peer = None
        if exchange not in ("auto", "p2p", "nccl"):
            raise ValueError("exchange must be 'auto', 'p2p' or 'nccl'")
        self.check_every = int(check_every)
        if exchange != "nccl" and dev.type == "cuda" and self.world_size > 1:
            try:
                # collective: either every rank gets the peer path or every rank raises here
                self.peer = PeerExchange(self.state.grad.numel(), process_group)
            except RuntimeError:
                if exchange == "p2p":
                    raise
                self.peer = None

    # ------------------------------------------------------------------ distributed plumbing
    @property
    def world_size(self) -> int:
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def zero_grad(self):
        self.state.zero_grad()

    def stage_logged(self, logged: Optional[Dict[str, torch.Tensor]] = None) -> None:
        """The logged scalars ride in the tail of the flat gradient buffer."""
        st = self.state
        st.rebind()
        if logged:
            sc = st.scalars()
            for i, key in enumerate(("kl_loss", "ce_loss", "loss")):
                if key in logged and logged[key] is not None:
                    sc[i].copy_(logged[key].detach().to(sc.dtype).reshape(()))

    def all_reduce_gradients(self, logged: Optional[Dict[str, torch.Tensor]] = None) -> None:
        """SUM over ranks of [grads | logged scalars], in place, one collective.  The division by
        world size is folded into the optimizer kernel (`grad_prescale`) and into `synced_logs`."""
        st = self.state
        self.stage_logged(logged)
        if self.world_size > 1:
            dist.all_reduce(st.grad, op=dist.ReduceOp.SUM, group=self.group)

    def synced_logs(self) -> Dict[str, torch.Tensor]:
        """Cross-rank means of the logged scalars (device tensors: no host sync here)."""
        sc = self.state.scalars() / self.world_size
        return {"kl_loss": sc[0], "ce_loss": sc[1], "loss": sc[2]}

    # ------------------------------------------------------------------ the optimizer step
    def current_lrs(self):
        f = cosine_warmup_factor(self.step_no, self.warm_steps, self.total_steps)
        return self.icv_lr * f, self.alpha_lr * f

    def step(self, logged: Optional[Dict[str, torch.Tensor]] = None, module=None) -> Dict[str, torch.Tensor]:
        """``module``: the VQAICVModule whose `global_step` this optimizer step advances (what
        Lightning does for the reference; the temperature schedule counts in it)."""
        if module is not None and hasattr(module, "on_optimizer_step"):
            module.on_optimizer_step()
        st = self.state
        if not st.param.is_cuda:
            raise RuntimeError("ICVDataParallelOptimizer.step runs the fused sm_100a optimizer "
                               "kernel: the ICV parameters must live on a B200 (no CPU optimizer)")
        from . import ops
        lr_vec, lr_alpha = self.current_lrs()          # scheduler value BEFORE this step's update
        if self.peer is not None:
            # ONE fused exchange + optimizer step over peer memory: no collective call
            self.stage_logged(logged)
            self.step_no += 1
            self.peer.step(st.param, st.grad, st.exp_avg, st.exp_avg_sq, st.n_vec,
                           st.n_alpha if st.alpha_learnable else 0,
                           st.grad.numel() - st.n_vec - (st.n_alpha if st.alpha_learnable else 0),
                           lr_vec, lr_alpha, self.betas, self.eps, self.weight_decay, self.step_no,
                           self.max_grad_norm, self.grad_norm, self._ws)
            logs = {k: v.clone() for k, v in self.synced_logs().items()} if logged else {}
            self.zero_grad()
            if self.check_every > 0 and self.step_no % self.check_every == 0 and self.peer.timed_out():
                raise RuntimeError(
                    "ICV gradient exchange: a peer did not deliver its gradient within ~4 s; the "
                    "optimizer updates since then were skipped on this rank (PeerExchange.reset_error "
                    "re-arms it once the ranks are back in step)")
            return logs
        self.all_reduce_gradients(logged)
        self.step_no += 1
        ops.adamw_step(st.param, st.grad, st.exp_avg, st.exp_avg_sq, st.n_vec,
                       st.n_alpha if st.alpha_learnable else 0, lr_vec, lr_alpha, self.step_no,
                       beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                       weight_decay=self.weight_decay, grad_prescale=1.0 / self.world_size,
                       max_grad_norm=self.max_grad_norm, norm_out=self.grad_norm,
                       workspace=self._ws)
        logs = self.synced_logs() if logged else {}
        logs = {k: v.clone() for k, v in logs.items()}
        self.zero_grad()
        return logs

    # ------------------------------------------------------------------ checkpoint of the state
    def state_dict(self):
        st = self.state
        return {"step": self.step_no, "exp_avg": st.exp_avg.detach().cpu().clone(),
                "exp_avg_sq": st.exp_avg_sq.detach().cpu().clone()}

    def load_state_dict(self, sd):
        st = self.state
        self.step_no = int(sd["step"])
        st.exp_avg.copy_(sd["exp_avg"])
        st.exp_avg_sq.copy_(sd["exp_avg_sq"])
