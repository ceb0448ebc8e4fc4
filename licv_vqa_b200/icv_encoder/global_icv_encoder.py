"""The learnable per-layer vectors v_l and scalars alpha_l.

Drop-in for the reference's GlobalICVEncoder (icv_src/icv_encoder/global_icv_encoder.py:5-43):
same constructor arguments and defaults, same parameter names (`alpha` [1,L], `icv` [1,L,d] - the
state-dict keys `icv_encoder.alpha` / `icv_encoder.icv` that inference.py:96-97 reads), same
initialisation (alpha = alpha_init_value, icv ~ N(0, 0.01^2)), same `forward()` / `get_alpha()`.

`scaled_icv()` is the B200 addition: the product `get_alpha().unsqueeze(-1) * icv` that every
caller forms next (icv_module.py:89-92, inference.py:311) as one fused kernel with a fused
backward (dv_l = a_l * ds_l, dalpha_l = ds_l . v_l, through the sigmoid when enabled).
"""
from __future__ import annotations

import torch

from .. import ops
from .base_icv_encoder import BaseICVEncoder, ICVEncoderOutput


class GlobalICVEncoder(BaseICVEncoder):
    def __init__(self, lmm_hidden_dim, lmm_layers, alpha_learnable=True, alpha_init_value=0.0,
                 use_sigmoid=False) -> None:
        super().__init__()
        self.alpha = torch.nn.Parameter(
            torch.full(size=(1, lmm_layers), fill_value=float(alpha_init_value)),
            requires_grad=alpha_learnable,
        )
        self.icv = torch.nn.Parameter(torch.empty(1, lmm_layers, lmm_hidden_dim))
        torch.nn.init.normal_(self.icv, mean=0.0, std=0.01)
        self.use_sigmoid = use_sigmoid

    def forward(self) -> ICVEncoderOutput:
        return ICVEncoderOutput(in_context_vector=self.icv, alpha=self.get_alpha(),
                                in_context_feature=None)

    def get_alpha(self):
        if self.use_sigmoid:
            return torch.sigmoid(self.alpha)
        return self.alpha

    def scaled_icv(self) -> torch.Tensor:
        """icv [1,L,d] fp32 = get_alpha()[..., None] * self.icv, one kernel forward / backward."""
        return ops.icv_scale(self.alpha, self.icv, self.use_sigmoid)
