"""ICV encoders: the learnable per-layer vectors v_l and scalars alpha_l, and their output record.

API-compatible with the reference's icv_src/icv_encoder package (base_icv_encoder.py:7-23,
global_icv_encoder.py:5-43): same class names, constructor arguments and defaults, the parameter
names `alpha` [1, L] and `icv` [1, L, d] (the state-dict keys `icv_encoder.alpha` /
`icv_encoder.icv` that inference.py:96-97 reads), the same initial values (alpha =
alpha_init_value everywhere, icv ~ N(0, 0.01^2)) and the same `forward()` / `get_alpha()`.

What is new here is `scaled_icv()`: the product `get_alpha().unsqueeze(-1) * icv` that every caller
of the reference forms next (icv_module.py:89-92, inference.py:311), as ONE kernel with a fused
backward (dv_l = a_l * ds_l, dalpha_l = ds_l . v_l, through the sigmoid when it is enabled).
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import torch

from .. import ops

_ICV_INIT_STD = 0.01


@dataclasses.dataclass
class ICVEncoderOutput:
    """What an encoder's forward returns; field names and order are the reference's."""

    in_context_feature: Optional[torch.Tensor] = None   # unused by the global encoder
    in_context_vector: Optional[torch.Tensor] = None    # v_l, [1, L, d]
    alpha: Optional[torch.Tensor] = None                # effective alpha_l (after the sigmoid), [1, L]

    def product(self) -> torch.Tensor:
        """alpha_l * v_l as the reference forms it (eager torch; `scaled_icv()` is the fused op)."""
        return self.alpha.unsqueeze(dim=-1) * self.in_context_vector


class BaseICVEncoder(torch.nn.Module):
    """Common root; subclasses own `alpha` and produce an ICVEncoderOutput."""

    def __init__(self) -> None:
        super().__init__()
        self.alpha = None
        self.icv_encoder = None

    def forward(self, *args, **kwargs) -> ICVEncoderOutput:
        raise NotImplementedError(f"{type(self).__name__} does not define forward()")


class GlobalICVEncoder(BaseICVEncoder):
    """One vector per hooked layer, shared by every query (the L-ICV of the paper)."""

    def __init__(self, lmm_hidden_dim, lmm_layers, alpha_learnable=True, alpha_init_value=0.0,
                 use_sigmoid=False) -> None:
        super().__init__()
        self.use_sigmoid = use_sigmoid
        self.alpha = torch.nn.Parameter(torch.empty(1, lmm_layers), requires_grad=alpha_learnable)
        self.icv = torch.nn.Parameter(torch.empty(1, lmm_layers, lmm_hidden_dim))
        self.reset_parameters(alpha_init_value)

    @torch.no_grad()
    def reset_parameters(self, alpha_init_value: float = 0.0) -> None:
        self.alpha.fill_(float(alpha_init_value))
        self.icv.normal_(mean=0.0, std=_ICV_INIT_STD)

    @property
    def n_layers(self) -> int:
        return self.icv.shape[1]

    @property
    def hidden_dim(self) -> int:
        return self.icv.shape[2]

    def get_alpha(self) -> torch.Tensor:
        return torch.sigmoid(self.alpha) if self.use_sigmoid else self.alpha

    def forward(self) -> ICVEncoderOutput:
        return ICVEncoderOutput(None, self.icv, self.get_alpha())

    def scaled_icv(self) -> torch.Tensor:
        """icv [1, L, d] fp32 = get_alpha()[..., None] * self.icv, one kernel forward / backward."""
        return ops.icv_scale(self.alpha, self.icv, self.use_sigmoid)

    def extra_repr(self) -> str:
        return (f"layers={self.n_layers}, hidden_dim={self.hidden_dim}, use_sigmoid={self.use_sigmoid}, "
                f"alpha_learnable={self.alpha.requires_grad}")
