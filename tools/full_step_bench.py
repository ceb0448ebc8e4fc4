"""End-to-end L-ICV training step on an idefics-9B-shaped tower (BASELINE configs[1] / [3] shapes).

    python tools/full_step_bench.py [--steps 5] [--layers 32] [--hooks fused|eager|none]
    torchrun --nproc-per-node N tools/full_step_bench.py ...        (data parallel, bs 8 per GPU)

The frozen tower is a random-init HF `LlamaForCausalLM` with idefics-9B's language-model shape
(32 layers, d = 4096, ffn 11008, 32 heads, V = 32002) in bf16; its GEMMs / attention / lm_head
are stock PyTorch + cuBLAS (out of scope by BASELINE.json north_star).  One step = the
reference's `training_step`: student pass on the zero-shot query (B x 32 tokens, ICV injected at
every decoder layer), teacher pass on the 32-shot prompt (B x 896 tokens, hooks off, no grad),
KL + 0.5 CE, backward to the ICV parameters, all-reduce, clip + AdamW.

`--hooks fused` is this repository's path (persistent hooks + licv kernels + fused loss +
flat-buffer optimizer); `--hooks eager` restates what the reference runs at the same place - its
hook body as five eager torch ops per layer with autograd's backward, boolean-mask gathers,
eager KL and HF's internal CE, torch.optim.AdamW - so the two can be timed on the same tower;
`--hooks none` is the tower alone (no injection, loss on plain logits) as the floor.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Interface(torch.nn.Module):
    """Duck-typed lmm_icl_interface.LMMInterface."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.tokenizer = type("Tok", (), {"pad_token_id": 0})()
        self.input_ids_field_name = "input_ids"

    @property
    def device(self):
        return next(self.model.parameters()).device

    def forward(self, **kw):
        return self.model(**kw)


def build_tower(layers, device, dtype):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=32002, hidden_size=4096, intermediate_size=11008,
                      num_hidden_layers=layers, num_attention_heads=32, num_key_value_heads=32,
                      max_position_embeddings=2048, pad_token_id=0, bos_token_id=1, eos_token_id=2,
                      tie_word_embeddings=False, attn_implementation="sdpa")
    torch.manual_seed(426)
    with torch.device(device):
        model = LlamaForCausalLM(cfg).to(dtype)
    model.eval()
    model.requires_grad_(False)
    return model


def make_batch(B, Tq, Tc, qx, V, device, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randint(3, V, (B, Tq), generator=g)
    q[:, 0] = 1
    ctx = torch.randint(3, V, (B, Tc), generator=g)
    ctx[:, 0] = 1
    t = torch.cat([ctx, q[:, 1:]], dim=1)
    ones = torch.ones
    return ({"input_ids": q.to(device), "attention_mask": ones(B, Tq, dtype=torch.long, device=device)},
            {"input_ids": t.to(device), "attention_mask": ones(B, Tc + Tq - 1, dtype=torch.long, device=device)},
            torch.full((B,), qx, dtype=torch.long, device=device),
            torch.full((B,), Tc + qx - 1, dtype=torch.long, device=device))


# ---- the reference's chain, restated as plain eager torch (timing comparator only) --------------
class EagerICV(torch.nn.Module):
    def __init__(self, tower, L, d):
        super().__init__()
        self.tower = tower
        self.alpha = torch.nn.Parameter(torch.full((1, L), 0.1))
        self.icv = torch.nn.Parameter(torch.randn(1, L, d) * 0.01)
        self.hooks_on = True
        self._icv = None
        for i, layer in enumerate(tower.model.layers):
            layer.register_forward_hook(self._hook(i))

    def _hook(self, i):
        def fn(_m, _a, out):
            if not self.hooks_on or self._icv is None:
                return None
            h = out[0] if isinstance(out, tuple) else out
            shift = self._icv[:, i].unsqueeze(dim=1)
            y = h + shift
            y = y / y.norm(dim=-1, keepdim=True) * h.norm(dim=-1, keepdim=True)
            return (y,) + tuple(out[1:]) if isinstance(out, tuple) else y
        return fn

    def step_loss(self, q, t, qxl, icl, hard_w=0.5, eps=1e-6):
        # bf16 ICV like the reference's DeepSpeed recipe (an fp32 ICV would promote the residual
        # stream to fp32, which a bf16 tower without autocast cannot consume)
        self._icv = (self.alpha.unsqueeze(-1) * self.icv).to(torch.bfloat16)
        self.hooks_on = True
        so = self.tower(**q, labels=q["input_ids"])
        with torch.no_grad():
            self.hooks_on = False
            tl = self.tower(**t).logits
        pos_q = torch.arange(q["input_ids"].shape[1], device=tl.device)[None]
        pos_t = torch.arange(t["input_ids"].shape[1], device=tl.device)[None]
        mq = (pos_q >= qxl[:, None]) & (q["input_ids"] != 0)
        mt = (pos_t >= icl[:, None]) & (t["input_ids"] != 0)
        stu = so.logits[mq].float()
        tea = tl[mt].float()
        p, qq = torch.softmax(tea, 1), torch.softmax(stu, 1)
        kl = (p * (torch.log(p + eps) - torch.log(qq + eps))).sum(1).mean()
        return kl + hard_w * so.loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--hooks", default="fused", choices=["fused", "eager", "none"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--tq", type=int, default=32)
    ap.add_argument("--tc", type=int, default=865)      # teacher length 896 = 32-shot prompt
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    tower = build_tower(args.layers, dev, torch.bfloat16)
    q, t, qxl, icl = make_batch(args.batch, args.tq, args.tc, args.tq - 4, 32002, dev, 1000 + rank)

    if args.hooks == "fused":
        from licv_vqa_b200 import LMMConfig, ModuleConfig, VQAICVModule
        from licv_vqa_b200.dp import ICVDataParallelOptimizer
        from licv_vqa_b200.icv_module import ICVEncoderConfig
        cfg = ModuleConfig(hard_loss_weight=0.5, ce_variant="causal_lm", residual_dtype="keep",
                           icv_encoder=ICVEncoderConfig(use_sigmoid=False, alpha_init_value=0.1))
        mod = VQAICVModule(Interface(tower), cfg, LMMConfig("llama", args.layers,
                                                            "model.model.layers.<LAYER_NUM>", -1, 4096)).to(dev)
        opt = ICVDataParallelOptimizer(mod.icv_encoder, cfg, total_steps=1000)

        def step():
            loss_dict, _ = mod(q, t, qxl, icl)
            loss_dict["loss"].backward()
            opt.step(loss_dict)
            return loss_dict["loss"]
    else:
        m = EagerICV(tower, args.layers, 4096).to(dev)
        if args.hooks == "none":
            m.hooks_on = False
            for p in (m.alpha, m.icv):
                p.requires_grad_(False)
            probe = torch.nn.Parameter(torch.zeros((), device=dev))

            def step():
                with torch.no_grad():
                    so = tower(**q).logits
                    tl = tower(**t).logits
                return so.float().mean() + tl[:, -1].float().mean() + probe
        else:
            topt = torch.optim.AdamW([{"params": m.alpha, "lr": 1e-2}, {"params": m.icv}], lr=1e-4,
                                     weight_decay=1e-3)

            def step():
                loss = m.step_loss(q, t, qxl, icl)
                loss.backward()
                if world > 1:
                    for p in (m.alpha, m.icv):
                        torch.distributed.all_reduce(p.grad)
                        p.grad /= world
                torch.nn.utils.clip_grad_norm_([m.alpha, m.icv], 1.0)
                topt.step()
                topt.zero_grad()
                return loss

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / args.steps
    if world > 1:
        tt = torch.tensor([sec], device=dev)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        sec = float(tt)
    if rank == 0:
        print(json.dumps({"what": "full L-ICV training step, idefics-9B-shaped LLaMA tower (bf16, random init)",
                          "hooks": args.hooks, "layers": args.layers, "n_gpus": world,
                          "batch_per_gpu": args.batch, "student_tokens": args.tq,
                          "teacher_tokens": args.tc + args.tq - 1, "ms_per_step": sec * 1e3,
                          "samples_per_s": args.batch * world / sec, "loss": float(loss),
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    if world > 1:
        torch.distributed.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
