"""TEST INFRASTRUCTURE ONLY - loader for the *real* reference (ForJadeForest/LICV-VQA).

This module imports the reference's own Python code from ``/root/reference`` so that
``oracle/make_golden.py`` can generate golden input/output vectors with it and so that the
restatement in ``oracle/licv_oracle.py`` can be validated against it.  ``/root/reference`` exists
only in the build container (never on the GPU box), so nothing in ``tests -m gpu``, ``smoke()`` or
``bench.py`` may import this file; the committed fixtures under ``tests/golden/`` are what travels.

Nothing from the reference is copied: the reference's files are imported / exec'd where they lie.

How (recipe from SURVEY.md Appendix B):

* ``icv_src/icv_model/icv_intervention.py`` and ``icv_src/icv_encoder/*`` import unmodified once
  ``sys.modules['baukit']`` holds a stand-in exposing ``TraceDict`` (the reference's only use of
  baukit is ``icv_intervention.py:7,90-97``).  The stand-in is a plain ``register_forward_hook``
  context manager - it contains no arithmetic.
* ``icv_src/icv_module.py`` needs hydra/lightning/deepspeed (absent), so the three methods on the
  hot path (``forward`` :71-119, ``calculate_kl_divergence`` :121-134, ``get_mask`` :136-148) are
  AST-extracted from the file and exec'd, then bound to a ``SimpleNamespace``.
"""
from __future__ import annotations

import ast
import contextlib
import os
import sys
import textwrap
import types
from types import SimpleNamespace

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("LICV_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "icv_src", "icv_module.py"))


class _TraceDictStandIn(contextlib.AbstractContextManager):
    """Minimal stand-in for ``baukit.TraceDict`` (hook plumbing only, no arithmetic).

    ``TraceDict(module, layers, edit_output=f(output, layer_name), retain_grad)`` registers one
    forward hook per named submodule whose return value replaces the module output.
    """

    def __init__(self, module, layers=None, edit_output=None, retain_grad=False, **_):
        self._handles = []
        named = dict(module.named_modules())
        for name in layers or []:
            sub = named[name]

            def hook(_m, _inp, out, _name=name):
                return edit_output(out, _name) if edit_output is not None else out

            self._handles.append(sub.register_forward_hook(hook))

    def __exit__(self, *exc):
        for h in self._handles:
            h.remove()
        self._handles = []
        return False


def _install_baukit_standin():
    if "baukit" not in sys.modules:
        mod = types.ModuleType("baukit")
        mod.TraceDict = _TraceDictStandIn
        sys.modules["baukit"] = mod


def load_reference_classes():
    """Return the reference's unmodified ``(LearnableICVInterventionLMM, GlobalICVEncoder,
    ICVEncoderOutput)`` classes."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_baukit_standin()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from icv_src.icv_model.icv_intervention import LearnableICVInterventionLMM
    from icv_src.icv_encoder.global_icv_encoder import GlobalICVEncoder
    from icv_src.icv_encoder.base_icv_encoder import ICVEncoderOutput

    return LearnableICVInterventionLMM, GlobalICVEncoder, ICVEncoderOutput


def load_reference_module_methods(names=("forward", "calculate_kl_divergence", "get_mask")):
    """AST-extract ``VQAICVModule.<name>`` from the reference file and return plain functions."""
    path = os.path.join(REFERENCE_ROOT, "icv_src", "icv_module.py")
    with open(path) as f:
        src = f.read()
    tree = ast.parse(src)
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "VQAICVModule":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in names:
                    seg = textwrap.dedent(ast.get_source_segment(src, item))
                    # `-> torch.Any` (icv_module.py:77) does not exist in torch 2.11
                    code = "from __future__ import annotations\n" + seg
                    ns = {"torch": torch}
                    exec(compile(code, f"{path}:{item.name}", "exec"), ns)
                    out[item.name] = ns[item.name]
    missing = set(names) - set(out)
    if missing:
        raise RuntimeError(f"could not extract {missing} from {path}")
    return out


def load_reference_collator():
    """The reference's own ``collator_data`` (icv_src/icv_datamodule.py:73-130), AST-extracted: the
    module imports pytorch_lightning / lmm_icl_interface at its top, the function itself needs
    only a prompt processor with ``prepare_input``, ``input_ids_field`` and ``tokenizer``."""
    path = os.path.join(REFERENCE_ROOT, "icv_src", "icv_datamodule.py")
    with open(path) as f:
        src = f.read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "collator_data":
            code = "from __future__ import annotations\n" + textwrap.dedent(ast.get_source_segment(src, node))
            ns = {"torch": torch}
            exec(compile(code, f"{path}:collator_data", "exec"), ns)
            return ns["collator_data"]
    raise RuntimeError(f"could not extract collator_data from {path}")


class InterfaceStandIn(nn.Module):
    """Duck-typed ``lmm_icl_interface.LMMInterface``: only the attributes the hot path touches
    (``icv_module.py:28-30,137-146``, ``icv_intervention.py:46,113,129``)."""

    def __init__(self, model, pad_token_id=0):
        super().__init__()
        self.model = model
        self.tokenizer = SimpleNamespace(pad_token_id=pad_token_id)
        self.input_ids_field_name = "input_ids"

    @property
    def device(self):
        return next(self.model.parameters()).device

    def forward(self, **kw):
        return self.model(**kw)

    def generate(self, **kw):
        return self.model.generate(**kw)


def build_reference_module(interface, icv_encoder, *, layer_format, total_layers,
                           intervention_layer=-1, hard_loss_weight=0.0, only_hard_loss=False,
                           kl_eps=1e-6, temperature=1.0):
    """A ``SimpleNamespace`` carrying the reference's own ``forward`` / ``calculate_kl_divergence``
    / ``get_mask`` bound to the reference's own ``LearnableICVInterventionLMM``."""
    LICV, _, _ = load_reference_classes()
    fns = load_reference_module_methods()
    icv_model = LICV(interface, enable_intervention=True, intervention_layer=intervention_layer,
                     layer_format=layer_format, total_layers=total_layers)
    self = SimpleNamespace(
        interface=interface,
        icv_model=icv_model,
        icv_encoder=icv_encoder,
        temperature=torch.nn.Parameter(torch.tensor(float(temperature)), requires_grad=False),
        module_cfg=SimpleNamespace(hard_loss_weight=hard_loss_weight,
                                   only_hard_loss=only_hard_loss, kl_eps=kl_eps),
    )
    for name, fn in fns.items():
        setattr(self, name, types.MethodType(fn, self))
    return self
