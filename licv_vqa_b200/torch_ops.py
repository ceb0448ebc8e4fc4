"""The hot path as registered torch operators: ``torch.ops.licv.*``.

``ops.py`` wraps the C ABI in ``autograd.Function``s, which is all eager training needs.  This
module registers the same launches with ``torch.library`` - a schema, a fake (meta) kernel and an
autograd formula each - so that ``torch.compile`` / ``torch.export`` graphs can hold them as single
nodes instead of breaking at a Python function (north_star: "registered as custom autograd ops that
replace the baukit TraceDict hooks", icv_src/icv_model/icv_intervention.py:88-113).  The kernels are
the same library entry points; there is still no CPU implementation (``device_types="cuda"``).

    out = torch.ops.licv.inject(h, shift, out_dtype, round_flags)            # differentiable
    total, kl, ce, dstu = torch.ops.licv.kd_loss(stu, tea, kl_tea_row, ce_label, counts, ...)

``licv::inject`` is the reference's ``intervention_function`` (icv_intervention.py:61-86) with its
closed-form backward; ``licv::kd_loss`` is ``calculate_kl_divergence`` + the shifted CE + the
combine (icv_module.py:94-134) in functional form (the gradient goes to a new tensor: a traced graph
must not see its input's storage change under it; the in-place form stays in ``ops.kd_loss``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _abi, ops

_lib = torch.library


# ---------------------------------------------------------------------------------------------
# a3 + a4: injection
# ---------------------------------------------------------------------------------------------
@_lib.custom_op("licv::inject", mutates_args=(), device_types="cuda")
def inject(h: Tensor, shift: Tensor, out_dtype: torch.dtype, round_flags: int) -> Tensor:
    return ops.inject_forward(h, shift, out_dtype, round_flags)


@inject.register_fake
def _(h, shift, out_dtype, round_flags):
    return torch.empty(h.shape, dtype=out_dtype, device=h.device)


@_lib.custom_op("licv::inject_bwd", mutates_args=(), device_types="cuda")
def inject_bwd(h: Tensor, g: Tensor, shift: Tensor, round_flags: int) -> Tuple[Tensor, Tensor]:
    """-> (dh, d_shift): d_shift = sum over tokens of the gradient of the normalised sum."""
    d_shift = torch.zeros_like(shift, dtype=torch.float32)
    dh = ops.inject_backward(h, g, shift.contiguous(), d_shift, True, round_flags)
    return dh, d_shift


@inject_bwd.register_fake
def _(h, g, shift, round_flags):
    return torch.empty_like(h, memory_format=torch.contiguous_format), torch.empty_like(shift, dtype=torch.float32)


def _inject_setup(ctx, inputs, output):
    h, shift, _out_dtype, round_flags = inputs
    ctx.save_for_backward(h, shift)          # h only: the reference keeps several fp32 [B,T,d] per layer
    ctx.round_flags = round_flags


def _inject_backward(ctx, g):
    h, shift = ctx.saved_tensors
    dh, d_shift = torch.ops.licv.inject_bwd(h, g.contiguous(), shift, ctx.round_flags)
    return dh, d_shift, None, None


inject.register_autograd(_inject_backward, setup_context=_inject_setup)


# ---------------------------------------------------------------------------------------------
# a7 - a10: distillation loss, functional form
# ---------------------------------------------------------------------------------------------
@_lib.custom_op("licv::kd_loss", mutates_args=(), device_types="cuda")
def kd_loss(stu: Tensor, tea: Optional[Tensor], kl_tea_row: Optional[Tensor], ce_label: Optional[Tensor],
            counts: Optional[Tensor], n_kl: int, n_ce: int, temperature: float, kl_eps: float,
            hard_loss_weight: float, only_hard_loss: bool, round_flags: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """-> (total, kl, ce, d total / d stu); one launch (see ops.kd_loss_raw)."""
    losses, dstu = ops.kd_loss_raw(stu, tea, kl_tea_row, ce_label, counts, n_kl, n_ce, temperature, kl_eps,
                                   hard_loss_weight, only_hard_loss, 1.0, False, True, round_flags)
    return losses[2].clone(), losses[0].clone(), losses[1].clone(), dstu


@kd_loss.register_fake
def _(stu, tea, kl_tea_row, ce_label, counts, n_kl, n_ce, temperature, kl_eps, hard_loss_weight, only_hard_loss,
      round_flags):
    s = torch.empty((), dtype=torch.float32, device=stu.device)
    return s, torch.empty_like(s), torch.empty_like(s), torch.empty_like(stu)


def _kd_setup(ctx, inputs, output):
    ctx.save_for_backward(output[3])


def _kd_backward(ctx, g_total, _g_kl, _g_ce, _g_dstu):
    (dstu,) = ctx.saved_tensors
    return (dstu * g_total.to(dstu.dtype),) + (None,) * 11


kd_loss.register_autograd(_kd_backward, setup_context=_kd_setup)


# ---------------------------------------------------------------------------------------------
# a6 / a7: masks and row lists (integer work, no gradient)
# ---------------------------------------------------------------------------------------------
@_lib.custom_op("licv::get_mask", mutates_args=(), device_types="cuda")
def get_mask(input_ids: Tensor, mask_length: Tensor, pad_token_id: int) -> Tensor:
    return ops.get_mask(input_ids, mask_length, pad_token_id)


@get_mask.register_fake
def _(input_ids, mask_length, pad_token_id):
    return torch.empty(input_ids.shape, dtype=torch.bool, device=input_ids.device)


REGISTERED = ("inject", "inject_bwd", "kd_loss", "get_mask")
ROUND_TEMPERED = _abi.ROUND_TEMPERED
