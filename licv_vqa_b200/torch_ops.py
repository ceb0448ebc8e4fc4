"""The hot path as registered torch operators: ``torch.ops.licv.*``.

``ops.py`` wraps the C ABI in ``autograd.Function``s over ctypes, which is all eager training
needs.  This module loads ``lib/liblicv_torch.so`` - ``csrc/licv_torch_ops.cpp``, a host-only
``TORCH_LIBRARY`` shim over the same C ABI - which registers the same launches as torch operators
from C++: a schema, a CUDA kernel, a Meta (fake) kernel and an autograd formula each.  A
``torch.compile`` / ``torch.export`` graph holds them as single nodes instead of breaking at a
Python function, and an eager call runs without a Python frame or a ctypes call in either
direction (north_star: "a thin C-ABI torch extension registered as custom autograd ops that replace
the baukit TraceDict hooks", icv_src/icv_model/icv_intervention.py:88-113).  There is still no CPU
implementation: the operators exist for the CUDA and Meta dispatch keys only.

    out = torch.ops.licv.inject(h, shift, out_dtype, round_flags)            # differentiable
    total, kl, ce, dstu = torch.ops.licv.kd_loss(stu, tea, kl_tea_row, ce_label, counts, ...)

``licv::inject`` is the reference's ``intervention_function`` (icv_intervention.py:61-86) with its
closed-form backward (``licv::inject_bwd``); ``licv::kd_loss`` is ``calculate_kl_divergence`` + the
shifted CE + the combine (icv_module.py:94-134) in functional form (the gradient goes to a new
tensor: a traced graph must not see its input's storage change under it; the in-place form stays
in ``ops.kd_loss``); ``licv::get_mask`` is icv_module.py:136-148.
"""
from __future__ import annotations

import os
import threading

import torch

from . import _abi

TORCH_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "liblicv_torch.so")
_lock = threading.Lock()
_loaded = False


def load() -> None:
    """Load the C++ registration (built in-tree by ``licv_vqa_b200.build.build_torch_ops``).  No
    Python re-implementation stands behind it: a missing library is built, or this raises."""
    global _loaded
    with _lock:
        if _loaded:
            return
        if not os.path.exists(TORCH_LIB_PATH):
            from . import build
            build.build_torch_ops()
        if not os.path.exists(_abi.LIB_PATH):
            _abi.load()        # builds liblicv_b200.so, which the registration links
        torch.ops.load_library(TORCH_LIB_PATH)
        _loaded = True


load()

inject = torch.ops.licv.inject
inject_bwd = torch.ops.licv.inject_bwd
kd_loss = torch.ops.licv.kd_loss
get_mask = torch.ops.licv.get_mask

REGISTERED = ("inject", "inject_bwd", "kd_loss", "get_mask")
ROUND_TEMPERED = _abi.ROUND_TEMPERED
