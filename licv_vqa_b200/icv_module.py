"""L-ICV task module: student/teacher passes, masks, the fused KL + CE distillation loss.

Mirrors the hot-path surface of the reference's VQAICVModule (icv_src/icv_module.py:15-216)
without Lightning/hydra/DeepSpeed: same constructor arguments (`interface, module_cfg, lmm_cfg`),
same attribute names (`interface`, `icv_model`, `icv_encoder`, `temperature`, `module_cfg`,
`lmm_cfg`), same `forward(query_inputs, inputs, query_x_length, in_context_length)` return value
`(loss_dict, icv_encoder_output)` with keys `kl_loss`, `ce_loss`, `loss`, same
`calculate_kl_divergence(stu_logits, tea_logits)` and `get_mask(inputs, mask_length)`, same option
names (including the reference's `min_tmeprature` spelling).

Under it the hot path is the B200 one: the injection hooks are fused kernels
(`icv_model.LearnableICVInterventionLMM`), the two boolean-mask gathers + `calculate_kl_divergence`
+ the HF-internal shifted CE + the combine are ONE kernel launch that reads each logits row once
and writes d(student logits) over the logits themselves (`ops.kd_loss`), with row selection done
by index lists computed on the device (`ops.kd_prepare_rows`) - no gathered copies, no host sync.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import torch
from torch import nn

from . import ops
from .icv_encoder.global_icv_encoder import GlobalICVEncoder
from .icv_model.icv_intervention import LearnableICVInterventionLMM


@dataclass
class ICVEncoderConfig:
    """config/icv_module/icv_encoder/global_icv_encoder.yaml:3-5"""
    use_sigmoid: bool = True
    alpha_learnable: bool = True
    alpha_init_value: float = 0.0


@dataclass
class ModuleConfig:
    """config/icv_module/icv_module.yaml:5-25 (names and defaults of the reference)."""
    hard_loss_weight: float = 0.0
    only_hard_loss: bool = False
    init_temperature: float = 1.0
    decay_ratio: float = -1
    decay_per_step: Any = -1
    min_tmeprature: float = 1.0  # (sic) the reference's key
    learnable_t: bool = False
    kl_eps: float = 1e-6
    log_alpha: bool = True
    alpha_lr: float = 1e-2
    icv_lr: float = 1e-4
    weight_decay: float = 1e-3
    warm_steps: Any = 0.1
    strategy: str = "ddp"
    icv_encoder: ICVEncoderConfig = field(default_factory=ICVEncoderConfig)
    # --- additions of this implementation (defaults keep the reference's semantics) ---
    ce_variant: str = "auto"            # "idefics" | "idefics2" | "causal_lm" | "auto" (by lmm name)
    image_token_id: int = -1            # idefics2: label id ignored by its CE (-1: the tower's config)
    # the reference enables activation checkpointing whenever the tower supports it
    # (icv_module.py:29-30); "reentrant" / "non_reentrant" pick torch.utils.checkpoint's mode
    # (True = the installed transformers' default), False turns it off
    gradient_checkpointing: Any = True
    check_row_counts: bool = False      # True: sync and raise when the two masks select != rows
    residual_dtype: str = "promote"     # see LearnableICVInterventionLMM
    # f1: run lm_head on the teacher's SELECTED rows only (the reference materialises
    # [B, ~900, V] teacher logits to use ~32 rows, icv_module.py:103-111).  "auto": when the tower
    # exposes the HF get_decoder() / get_output_embeddings() pair; False: full logits
    teacher_rows_before_lm_head: Any = "auto"


@dataclass
class LMMConfig:
    """config/lmm/*.yaml: the arguments of the intervention."""
    name: str = "idefics-9b"
    total_layers: int = 32
    layer_format: str = "model.model.layers.<LAYER_NUM>"
    intervention_layer: Any = -1
    hidden_size: int = 4096


LMM_PRESETS = {
    # config/lmm/idefics-9B.yaml:3-9
    "idefics-9b": LMMConfig("idefics-9b", 32, "model.model.layers.<LAYER_NUM>", -1, 4096),
    # config/lmm/idefics2-8B-base.yaml:3-10 (hook on the MLP output, before the residual add)
    "idefics2-8b-base": LMMConfig("idefics2-8b-base", 32,
                                  "model.model.text_model.layers.<LAYER_NUM>.mlp", -1, 4096),
    # config/lmm/openflamingov2-9B.yaml:3-9
    "openflamingov2-9B": LMMConfig("openflamingov2-9B", 32,
                                   "model.lang_encoder.transformer.blocks.<LAYER_NUM>", -1, 4096),
}


def _get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


class VQAICVModule(nn.Module):
    def __init__(self, interface: nn.Module, module_cfg=None, lmm_cfg=None) -> None:
        super().__init__()
        self.module_cfg = module_cfg if module_cfg is not None else ModuleConfig()
        self.lmm_cfg = lmm_cfg if lmm_cfg is not None else LMMConfig()
        self.interface = interface

        self.interface.requires_grad_(False)
        gc = _get(self.module_cfg, "gradient_checkpointing", True)
        if gc and hasattr(self.interface.model, "gradient_checkpointing_enable"):
            if gc in ("reentrant", "non_reentrant"):
                self.interface.model.gradient_checkpointing_enable(
                    gradient_checkpointing_kwargs={"use_reentrant": gc == "reentrant"})
            else:
                self.interface.model.gradient_checkpointing_enable()

        self.icv_model = LearnableICVInterventionLMM(
            interface,
            enable_intervention=True,
            intervention_layer=_get(self.lmm_cfg, "intervention_layer"),
            layer_format=_get(self.lmm_cfg, "layer_format"),
            total_layers=_get(self.lmm_cfg, "total_layers"),
            residual_dtype=_get(self.module_cfg, "residual_dtype", "promote"),
        )

        enc_cfg = _get(self.module_cfg, "icv_encoder", None) or ICVEncoderConfig()
        self.icv_encoder = GlobalICVEncoder(
            lmm_hidden_dim=_get(self.lmm_cfg, "hidden_size"),
            lmm_layers=len(self.icv_model.intervention_layer_names),
            alpha_learnable=_get(enc_cfg, "alpha_learnable", True),
            alpha_init_value=_get(enc_cfg, "alpha_init_value", 0.0),
            use_sigmoid=_get(enc_cfg, "use_sigmoid", False),
        )

        # icv_module.py:49-52.  learnable_t: d loss / d T comes out of the loss launch (the generic
        # kernel) and the value is read from the device once per step; otherwise the kernels take
        # T from the host mirror and nothing synchronises
        init_t = float(_get(self.module_cfg, "init_temperature", 1.0))
        self.learnable_t = bool(_get(self.module_cfg, "learnable_t", False))
        self.temperature = torch.nn.Parameter(torch.tensor(init_t), requires_grad=self.learnable_t)
        self._temperature_value = init_t  # host mirror: the kernels take T by value, no sync
        self.decay_per_step = None
        self.global_step = 0

    # ------------------------------------------------------------------ temperature schedule
    def set_temperature(self, value: float):
        self._temperature_value = float(value)
        with torch.no_grad():
            self.temperature.fill_(float(value))

    def _temperature(self):
        """What the loss takes as T: the Parameter itself when it is learnable (its gradient is
        wanted and its value lives on the device), else the host mirror."""
        return self.temperature if self.learnable_t else self._temperature_value

    def setup_temperature_decay(self, estimated_stepping_batches: int):
        """on_train_start (icv_module.py:54-69)."""
        dps = _get(self.module_cfg, "decay_per_step", -1)
        if dps < 0:
            return -1
        if isinstance(dps, int):
            self.decay_per_step = dps
        elif isinstance(dps, float) and 0 < dps < 1:
            self.decay_per_step = int(estimated_stepping_batches * dps)
        else:
            raise ValueError("decay_ratio must be an int or a float between 0 and 1")
        return self.decay_per_step

    def decay_temperature(self):
        """icv_module.py:150-158, on the host mirror (no device read)."""
        ratio = _get(self.module_cfg, "decay_ratio", -1)
        if ratio < 0:
            return
        if self.decay_per_step is None:
            if self.setup_temperature_decay(0) in (None, -1) or not self.decay_per_step:
                raise RuntimeError("decay_ratio is set but the decay period is not: call "
                                   "setup_temperature_decay(estimated_stepping_batches) first "
                                   "(on_train_start in the reference, icv_module.py:54-69)")
        if self.global_step % self.decay_per_step == 0 and self.global_step != 0:
            self.set_temperature(max(self._temperature_value * ratio,
                                     _get(self.module_cfg, "min_tmeprature", 1.0)))

    # ------------------------------------------------------------------ hot path
    def _ce_variant(self):
        v = _get(self.module_cfg, "ce_variant", "auto")
        if v != "auto":
            return v
        name = str(_get(self.lmm_cfg, "name", "")).lower()
        if "idefics2" in name:
            return "idefics2"
        if "idefics" in name:
            return "idefics"
        return "causal_lm"

    def _image_token_id(self):
        """idefics2's CE ignores the image token (HF: CrossEntropyLoss(ignore_index=image_token_id));
        taken from the tower's config unless the module config names it."""
        tid = int(_get(self.module_cfg, "image_token_id", -1))
        if tid >= 0:
            return tid
        conf = getattr(getattr(self.interface, "model", None), "config", None)
        tid = getattr(conf, "image_token_id", None)
        return int(tid) if isinstance(tid, int) and tid >= 0 else -1

    def forward(self, query_inputs, inputs, query_x_length, in_context_length):
        """One student pass (hooks on) + one teacher pass (hooks off, no grad) + the fused loss.

        query_inputs / inputs: the collator's dicts (icv_datamodule.py:125-130) holding at least
        `input_ids` [B,T] (+ `attention_mask`, pixel values, ...), passed through to the tower.
        """
        cfg = self.module_cfg
        ids_name = self.interface.input_ids_field_name
        pad_id = self.interface.tokenizer.pad_token_id
        hard_w = float(_get(cfg, "hard_loss_weight", 0.0) or 0.0)
        only_hard = bool(_get(cfg, "only_hard_loss", False))
        want_ce = bool(hard_w) or only_hard

        icv_encoder_output = self.icv_encoder()
        icv = self.icv_encoder.scaled_icv()

        # the reference sets labels = input_ids here so that HF computes the CE internally
        # (icv_module.py:94-95); the CE is fused into the loss kernel instead
        query_inputs = {k: v for k, v in query_inputs.items() if k != "labels"}
        self.icv_model.toggle_intervention(True)
        icv_logits = self.icv_model(**query_inputs, icv=icv)["logits"]
        V = icv_logits.shape[-1]
        stu_ids = query_inputs[ids_name]

        if only_hard:
            _, ce_label, counts = ops.kd_prepare_rows(
                stu_ids, query_x_length, stu_ids, query_x_length, pad_id,
                query_inputs.get("attention_mask"), self._ce_variant(),
                self._image_token_id(), want_ce=True)
            total, _, _ = ops.kd_loss(icv_logits.view(-1, V), None, None, ce_label, counts,
                                      temperature=self._temperature_value, only_hard_loss=True)
            return {"loss": total}, icv_encoder_output

        compact = self._teacher_head_parts() is not None
        prep = ops.kd_prepare_rows(
            stu_ids, query_x_length, inputs[ids_name], in_context_length, pad_id,
            query_inputs.get("attention_mask"), self._ce_variant(),
            self._image_token_id(), want_ce=want_ce, compact_teacher=compact)
        kl_tea_row, ce_label, counts = prep[:3]
        with torch.no_grad():
            self.icv_model.toggle_intervention(False)
            if compact:
                ice_logits = self._teacher_logits_of_rows(inputs, prep[3])
            else:
                ice_logits = self.icv_model(**inputs)["logits"]
        if ice_logits.dtype != icv_logits.dtype:
            ice_logits = ice_logits.to(icv_logits.dtype)
        if _get(cfg, "check_row_counts", False):
            n_s, _, n_t, _ = counts.tolist()
            if n_s != n_t:
                raise RuntimeError(f"The size of tensor a ({n_t}) must match the size of tensor b "
                                   f"({n_s}) at non-singleton dimension 0")
        total, kl, ce = ops.kd_loss(
            icv_logits.view(-1, V), ice_logits.view(-1, V), kl_tea_row, ce_label, counts,
            temperature=self._temperature(), kl_eps=float(_get(cfg, "kl_eps", 1e-6)),
            hard_loss_weight=hard_w)
        loss_dict = {"kl_loss": kl}
        if want_ce:
            loss_dict["ce_loss"] = ce
        loss_dict["loss"] = total
        return loss_dict, icv_encoder_output

    # ------------------------------------------------------------------ f1: teacher rows only
    def _teacher_head_parts(self):
        """(decoder, lm_head) when the teacher's logits can be restricted to the selected rows."""
        want = _get(self.module_cfg, "teacher_rows_before_lm_head", "auto")
        if want is False:
            return None
        model = getattr(self.interface, "model", None)
        dec = getattr(model, "get_decoder", None)
        head = getattr(model, "get_output_embeddings", None)
        try:
            dec, head = (dec() if dec else None), (head() if head else None)
        except Exception:
            dec = head = None
        if dec is None or head is None or dec is model:
            if want is True:
                raise RuntimeError("teacher_rows_before_lm_head=True needs a tower with "
                                   "get_decoder() and get_output_embeddings()")
            return None
        return dec, head

    def _teacher_logits_of_rows(self, inputs, tea_sel):
        """lm_head over the gathered hidden rows: [B*Tq, V] compact teacher logits (hooks are off:
        the decoder's layers carry them, disabled by toggle_intervention(False))."""
        dec, head = self._teacher_head_parts()
        previous = self.icv_model._active
        self.icv_model._active = None          # what LearnableICVInterventionLMM._run does when off
        try:
            hidden = dec(**{k: v for k, v in inputs.items() if k != "labels"})
        finally:
            self.icv_model._active = previous
        hidden = hidden[0] if isinstance(hidden, tuple) else hidden.last_hidden_state
        rows = hidden.reshape(-1, hidden.shape[-1]).index_select(0, tea_sel.long())
        return head(rows)

    def calculate_kl_divergence(self, stu_logits, tea_logits):
        """T^2 * mean_rows sum_v p (log(p+eps) - log(q+eps)) (icv_module.py:121-134).

        stu_logits / tea_logits [N,V].  Like the reference this consumes its arguments: the
        student logits' storage is overwritten (here with the gradient, there with logits/T)."""
        total, _, _ = ops.kd_loss(stu_logits, tea_logits.detach(),
                                  temperature=self._temperature(),
                                  kl_eps=float(_get(self.module_cfg, "kl_eps", 1e-6)))
        return total

    def get_mask(self, inputs, mask_length):
        """icv_module.py:136-148."""
        return ops.get_mask(inputs[self.interface.input_ids_field_name], mask_length,
                            self.interface.tokenizer.pad_token_id)

    def training_step(self, batch, batch_idx=0):
        """icv_module.py:160-169 minus the logger: returns (loss, loss_dict)."""
        self.decay_temperature()
        loss_dict, _ = self(**batch)
        return loss_dict["loss"], loss_dict

    def on_optimizer_step(self):
        """Lightning advances `global_step` once per optimizer step; without Lightning the training
        loop (or ICVDataParallelOptimizer.step(module=...)) calls this - the temperature schedule
        (icv_module.py:150-158) counts in it."""
        self.global_step += 1

    # ------------------------------------------------------------------ checkpoint (f3)
    def icv_checkpoint(self):
        """The dict the reference writes as icv_cpk.pth (train.py:97-106) and inference.py:95-100
        reads: keys `icv_encoder.icv`, `icv_encoder.alpha`, `use_sigmoid`, `lmm_args`."""
        lmm = self.lmm_cfg
        return {
            "icv_encoder.icv": self.icv_encoder.icv.detach().float().cpu().clone(),
            "icv_encoder.alpha": self.icv_encoder.alpha.detach().float().cpu().clone(),
            "temperature": self.temperature.detach().cpu().clone(),
            "use_sigmoid": self.icv_encoder.use_sigmoid,
            "lmm_args": {
                "name": _get(lmm, "name"),
                "total_layers": _get(lmm, "total_layers"),
                "layer_format": _get(lmm, "layer_format"),
                "intervention_layer": _get(lmm, "intervention_layer"),
                "hidden_size": _get(lmm, "hidden_size"),
            },
        }

    def save_icv_checkpoint(self, path):
        torch.save(self.icv_checkpoint(), path)

    def load_icv_checkpoint(self, path_or_dict):
        ck = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict,
                                                                            map_location="cpu")
        with torch.no_grad():
            self.icv_encoder.icv.copy_(ck["icv_encoder.icv"])
            self.icv_encoder.alpha.copy_(ck["icv_encoder.alpha"])
        if ck.get("use_sigmoid", None) is not None:
            self.icv_encoder.use_sigmoid = bool(ck["use_sigmoid"])
        return ck


def load_icv_for_inference(path_or_dict, device):
    """What inference.py:95-100 does with icv_cpk.pth: -> (icv [1,L,d], alpha_eff [1,L], lmm_args),
    alpha already passed through the sigmoid when the checkpoint says so."""
    ck = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict,
                                                                        map_location="cpu")
    icv = ck["icv_encoder.icv"].to(device)
    alpha = ck["icv_encoder.alpha"].to(device)
    if ck.get("use_sigmoid", None):
        alpha = torch.sigmoid(alpha)
    return icv, alpha, dict(ck["lmm_args"])
