"""licv_vqa_b200 - B200-native (sm_100a) implementation of L-ICV's data-parallel hot path.

Drop-in for the reference's `icv_encoder` / `icv_model` hook API and `icv_module` loss
(ForJadeForest/LICV-VQA): the residual-stream injection and the KL + CE distillation loss are
hand-written CUDA kernels behind a C ABI (include/licv_b200.h, lib/liblicv_b200.so).
GPU only - there is no CPU path and no fallback.
"""
from .icv_encoder import BaseICVEncoder, GlobalICVEncoder, ICVEncoderOutput
from .icv_model import LearnableICVInterventionLMM
from .collate import check_batch_contract, collate_token_ids
from .icv_module import (ICVEncoderConfig, LMM_PRESETS, LMMConfig, ModuleConfig, VQAICVModule,
                         load_icv_for_inference)

__all__ = [
    "BaseICVEncoder", "GlobalICVEncoder", "ICVEncoderOutput", "LearnableICVInterventionLMM",
    "VQAICVModule", "ModuleConfig", "LMMConfig", "ICVEncoderConfig", "LMM_PRESETS",
    "load_icv_for_inference", "collate_token_ids", "check_batch_contract",
]
__version__ = "0.1.0"
