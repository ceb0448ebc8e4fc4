"""Output record and base class of the ICV encoders.

Mirrors the reference's icv_src/icv_encoder/base_icv_encoder.py:7-23 (same names and fields) so
code written against the reference keeps working.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn


@dataclass
class ICVEncoderOutput:
    in_context_feature: Optional[torch.Tensor]
    in_context_vector: Optional[torch.Tensor]   # v_l            [1, L, d]
    alpha: Optional[torch.Tensor]               # effective a_l  [1, L]


class BaseICVEncoder(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.alpha = None
        self.icv_encoder = None

    def forward(self, *args, **kwargs) -> ICVEncoderOutput:
        raise NotImplementedError
