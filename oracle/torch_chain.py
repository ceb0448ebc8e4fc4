"""TEST INFRASTRUCTURE ONLY - the reference's eager op chain, restated with torch on the host CPU.

Where ``licv_oracle.py`` is the closed-form float64 checker, this file restates the hot path the
way the reference *executes* it: the same sequence of eager torch ops with autograd doing the
backward, in whatever dtype the tensors carry.  Two uses:

* ``bench.py``'s ``cpu_baseline`` leg and ``--impl reference`` arm time it on the GPU box's host
  cores (``kind: "port"`` - the reference itself is Python that needs ``/root/reference`` plus
  absent packages, so it cannot travel to the GPU box);
* tests use it to see how far the reference's own low-precision chain sits from exact arithmetic,
  which is what the stated tolerances have to absorb.

Cites (relative to /root/reference):
  shift_and_rescale   icv_src/icv_model/icv_intervention.py:66-72 (tuple branch) / :76-82
  scaled_icv          icv_src/icv_module.py:89-92
  kd_term             icv_src/icv_module.py:121-134
  hard_term           transformers 4.38.2 IdeficsForVisionText2Text loss (not in /root/reference;
                      SURVEY.md §8c) consumed at icv_src/icv_module.py:94-98,115-117
  total               icv_src/icv_module.py:107-118
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def scaled_icv(alpha_eff: torch.Tensor, vec: torch.Tensor) -> torch.Tensor:
    return alpha_eff.unsqueeze(-1) * vec


def shift_and_rescale(hidden: torch.Tensor, icv: torch.Tensor, icv_index: int) -> torch.Tensor:
    """hidden [B,T,d]; icv [1,L,d]; 5 eager ops, result dtype by torch type promotion."""
    moved = hidden + icv[:, icv_index].unsqueeze(1)
    len_before = hidden.norm(dim=-1, keepdim=True)
    len_after = moved.norm(dim=-1, keepdim=True)
    return moved / len_after * len_before


def kd_term(stu_rows: torch.Tensor, tea_rows: torch.Tensor, temperature, kl_eps: float):
    """[N,V] x2 -> scalar.  Like the reference this evaluates softmax(tea) twice and divides its
    arguments by T in place (they must be gathered copies, not leaves)."""
    stu_rows /= temperature
    tea_rows /= temperature
    gap = (tea_rows.softmax(1) + kl_eps).log() - (stu_rows.softmax(1) + kl_eps).log()
    return (tea_rows.softmax(1) * gap).sum(1).mean() * temperature ** 2


def hard_term(logits: torch.Tensor, input_ids: torch.Tensor, attention_mask: torch.Tensor):
    """Shifted next-token CE with labels = input_ids, rows kept where attention_mask[...,1:] != 0."""
    keep = attention_mask[..., 1:] != 0
    return F.cross_entropy(logits[..., :-1, :][keep], input_ids[..., 1:][keep])


def hot_path_step(h_layers, g_layers, alpha_raw, vec, use_sigmoid, stu_logits, tea_rows,
                  stu_mask, input_ids, attention_mask, temperature, kl_eps, hard_loss_weight):
    """One pass of the hot path over one batch, the way the reference's eager chain does it.

    h_layers / g_layers: per hooked layer, the layer output [B,T,d] and the gradient arriving at
    the injected output.  stu_logits [B,Tq,V] (a leaf here, standing for lm_head's output),
    tea_rows [N,V] already gathered.  Returns (losses, d_alpha_raw, d_vec, dh per layer,
    d_stu_logits).
    """
    alpha_raw = alpha_raw.detach().requires_grad_(True)
    vec = vec.detach().requires_grad_(True)
    alpha_eff = torch.sigmoid(alpha_raw) if use_sigmoid else alpha_raw
    icv = scaled_icv(alpha_eff, vec)
    hs = [h.detach().requires_grad_(True) for h in h_layers]
    outs = [shift_and_rescale(h, icv, i) for i, h in enumerate(hs)]
    stu_logits = stu_logits.detach().requires_grad_(True)
    kl = kd_term(stu_logits[stu_mask].view(-1, stu_logits.shape[-1]), tea_rows.clone(),
                 temperature, kl_eps)
    loss = kl
    ce = None
    if hard_loss_weight:
        ce = hard_term(stu_logits, input_ids, attention_mask)
        loss = loss + hard_loss_weight * ce
    torch.autograd.backward([loss] + outs, [torch.ones_like(loss)] + list(g_layers))
    return (dict(kl=kl.detach(), ce=None if ce is None else ce.detach(), loss=loss.detach()),
            alpha_raw.grad, vec.grad, [h.grad for h in hs], stu_logits.grad)
