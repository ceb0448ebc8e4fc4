"""CPU-side checks of the drop-in boundary: liblicv_b200.so builds, loads and exports every symbol
include/licv_b200.h declares; the ctypes table mirrors the header; the host-side classes keep the
reference's API; and the product fails loudly instead of computing on the CPU."""
import ctypes as C
import os
import re

import pytest
import torch

from licv_vqa_b200 import (GlobalICVEncoder, ICVEncoderOutput, LMM_PRESETS,
                           LearnableICVInterventionLMM, ModuleConfig, _abi, ops)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "licv_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(licv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()                      # builds with nvcc if the .so is missing
    raw = C.CDLL(_abi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 23
    for name in names:
        assert hasattr(raw, name), f"{name} is declared in licv_b200.h but not exported"
    assert sorted(_abi.SIGNATURES) == names, "ctypes table and header disagree"
    assert lib.licv_abi_version() == 1
    assert b"dtype" in lib.licv_status_string(-2)
    assert lib.licv_kd_loss_workspace_bytes(10) == 16 + 80
    assert lib.licv_kd_loss_workspace_bytes(0) == 16


def test_no_device_is_an_error_not_a_fallback():
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    lib = _abi.load()
    sm, major, minor = C.c_int(), C.c_int(), C.c_int()
    assert lib.licv_device_info(C.byref(sm), C.byref(major), C.byref(minor)) == -7  # LICV_ERR_NO_DEVICE
    # argument checks come after the device check: every compute entry point refuses to run
    assert lib.licv_inject_fwd(0, 0, 0, 4, 512, 1, 1, 0, 0) == -7
    assert lib.licv_kd_loss_fwd_bwd(0, 0, 0, 0, 0, 0, 0, 0, 1.0, 1e-6, 0.0, 0, 1.0, 0, 0, 0, 8, 8, 8, 0, 0,
                                    0) == -7
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.inject_forward(torch.zeros(2, 8), torch.zeros(8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.kd_loss_raw(torch.zeros(2, 8), torch.zeros(2, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.get_mask(torch.zeros(2, 3, dtype=torch.long), torch.zeros(2, dtype=torch.long), 0)


def test_exchange_region_holds_two_packet_slots_per_source_rank():
    """licv_dp_region_bytes is host arithmetic: 2 parities x 16 possible source ranks x one 16-byte
    {float, tag, float, tag} packet per two gradient floats (slots rounded to 256 B), plus the control
    block and the per-CTA norm partials.  It must grow with the gradient and never be smaller than
    the packets it has to hold."""
    lib = _abi.load()
    assert lib.licv_dp_region_bytes(-1) == 0
    prev = 0
    for n in (1, 3, 4, 1000, 32 * 4096 + 32 + 4):
        got = lib.licv_dp_region_bytes(n)
        padded = (n + 3) // 4 * 4                     # FlatICVState pads the gradient to 4 floats
        packets = (padded + 1) // 2
        slot = (packets * 16 + 255) // 256 * 256
        assert got >= 2 * 16 * slot + 64
        assert got <= 2 * 16 * slot + 64 + 4096
        assert got >= prev
        prev = got


def test_reference_rounding_table():
    """Where the reference's eager chain rounds (probe-verified dtype rule)."""
    f, dt = ops.reference_rounding(torch.bfloat16, torch.float32, autocast=False)
    assert dt == torch.float32 and f == _abi.ROUND_NH
    f, dt = ops.reference_rounding(torch.bfloat16, torch.float32, autocast=True)
    assert dt == torch.float32 and f == 0
    f, dt = ops.reference_rounding(torch.float16, torch.float16, autocast=False)
    assert dt == torch.float16 and f == (_abi.ROUND_Y | _abi.ROUND_NH | _abi.ROUND_NY | _abi.ROUND_T)
    f, dt = ops.reference_rounding(torch.float32, torch.float32, autocast=False)
    assert dt == torch.float32 and f == 0


class Tower(torch.nn.Module):
    def __init__(self, n=3, d=8):
        super().__init__()
        self.layers = torch.nn.ModuleList(torch.nn.Linear(d, d) for _ in range(n))

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


def test_intervention_wrapper_keeps_the_reference_api():
    tower = Tower()
    m = LearnableICVInterventionLMM(tower, True, -1, "layers.<LAYER_NUM>", 3)
    assert m.intervention_layers == [0, 1, 2]
    assert m.intervention_layer_names == ["layers.0", "layers.1", "layers.2"]
    assert m.layer_to_icv_index == {0: 0, 1: 1, 2: 2}
    assert m.intervention_status is True and m.intervention_enabled is True
    assert len(m._hook_handles) == 3 and m.lmm is tower
    m.toggle_intervention(False)
    assert m.intervention_status is False
    with pytest.raises(ValueError, match="boolean"):
        m.intervention_status = 1
    # disabled: a plain forward of the tower, no ICV needed (the teacher pass)
    x = torch.randn(2, 8)
    assert torch.equal(m(x=x), tower(x))
    # enabled without an icv: the reference dies subscripting None; same exception type here
    m.toggle_intervention(True)
    with pytest.raises(TypeError):
        m(x=x)
    # enabled on CPU tensors: the injection has no CPU path and says so
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(icv=torch.zeros(1, 3, 8), x=x)
    # a single layer / a list of layers, like the reference's _prepare_layers
    one = LearnableICVInterventionLMM(Tower(), True, 1, "layers.<LAYER_NUM>", 3)
    assert one.intervention_layers == [1] and one.layer_to_icv_index == {1: 0}
    some = LearnableICVInterventionLMM(Tower(), True, [2, 0], "layers.<LAYER_NUM>", 3)
    assert some.layer_to_icv_index == {2: 0, 0: 1}
    with pytest.raises(LookupError):
        LearnableICVInterventionLMM(Tower(), True, -1, "blocks.<LAYER_NUM>", 3)
    # a second wrapper over the same tower retires the first one's hooks
    again = LearnableICVInterventionLMM(tower, True, -1, "layers.<LAYER_NUM>", 3)
    assert m._hook_handles == [] and len(again._hook_handles) == 3
    off = LearnableICVInterventionLMM(Tower(), enable_intervention=False)
    assert torch.is_tensor(off(x=x))


def test_encoder_defaults_and_presets():
    enc = GlobalICVEncoder(lmm_hidden_dim=16, lmm_layers=4)
    assert enc.alpha.shape == (1, 4) and enc.icv.shape == (1, 4, 16)
    assert float(enc.alpha.abs().max()) == 0.0 and enc.use_sigmoid is False
    assert 0.005 < float(enc.icv.std()) < 0.02          # N(0, 0.01^2)
    out = enc()
    assert isinstance(out, ICVEncoderOutput) and out.in_context_feature is None
    assert out.in_context_vector is enc.icv
    sig = GlobalICVEncoder(16, 4, use_sigmoid=True, alpha_init_value=0.0)
    assert torch.allclose(sig.get_alpha(), torch.full((1, 4), 0.5))
    frozen = GlobalICVEncoder(16, 4, alpha_learnable=False, alpha_init_value=0.3)
    assert not frozen.alpha.requires_grad and sorted(frozen.state_dict()) == ["alpha", "icv"]
    # config/lmm/*.yaml and config/icv_module/icv_module.yaml of the reference
    assert LMM_PRESETS["idefics-9b"].layer_format == "model.model.layers.<LAYER_NUM>"
    assert LMM_PRESETS["idefics2-8b-base"].layer_format.endswith(".mlp")
    cfg = ModuleConfig()
    assert (cfg.hard_loss_weight, cfg.kl_eps, cfg.alpha_lr, cfg.icv_lr, cfg.weight_decay,
            cfg.warm_steps, cfg.min_tmeprature) == (0.0, 1e-6, 1e-2, 1e-4, 1e-3, 0.1, 1.0)


def test_checkpoint_wire_format_round_trip(tmp_path):
    """f3: the dict the reference writes as icv_cpk.pth (train.py:97-106) and inference.py:95-100
    reads - keys icv_encoder.icv / icv_encoder.alpha / use_sigmoid / lmm_args."""
    from licv_vqa_b200 import LMMConfig, VQAICVModule, load_icv_for_inference
    from licv_vqa_b200.icv_module import ICVEncoderConfig

    class Iface(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.model = Tower(3, 8)
            self.tokenizer = type("Tok", (), {"pad_token_id": 0})()
            self.input_ids_field_name = "input_ids"

        def forward(self, **kw):
            return self.model(**kw)

    cfg = ModuleConfig(icv_encoder=ICVEncoderConfig(use_sigmoid=True, alpha_init_value=0.2))
    lmm = LMMConfig("toy", 3, "model.layers.<LAYER_NUM>", -1, 8)
    mod = VQAICVModule(Iface(), cfg, lmm)
    assert all(not p.requires_grad for p in mod.interface.parameters())     # frozen tower
    assert sorted(k for k, p in mod.named_parameters() if p.requires_grad) == \
        ["icv_encoder.alpha", "icv_encoder.icv"]
    with torch.no_grad():
        mod.icv_encoder.icv.normal_()
    path = tmp_path / "icv_cpk.pth"
    mod.save_icv_checkpoint(path)
    ck = torch.load(path, map_location="cpu")
    assert {"icv_encoder.icv", "icv_encoder.alpha", "use_sigmoid", "lmm_args"} <= set(ck)
    assert ck["lmm_args"]["layer_format"] == "model.layers.<LAYER_NUM>"
    assert ck["lmm_args"]["total_layers"] == 3 and ck["lmm_args"]["intervention_layer"] == -1
    # what inference.py does with it: alpha through the sigmoid when the checkpoint says so
    icv, alpha, lmm_args = load_icv_for_inference(path, "cpu")
    assert torch.equal(icv, mod.icv_encoder.icv.detach())
    assert torch.allclose(alpha, torch.sigmoid(mod.icv_encoder.alpha.detach()))
    assert lmm_args["hidden_size"] == 8
    # and back into a fresh module
    other = VQAICVModule(Iface(), cfg, lmm)
    other.load_icv_checkpoint(path)
    assert torch.equal(other.icv_encoder.icv, mod.icv_encoder.icv)
    # temperature schedule on the host mirror (icv_module.py:54-69,150-158)
    dec = VQAICVModule(Iface(), ModuleConfig(init_temperature=4.0, decay_ratio=0.5, decay_per_step=2,
                                             min_tmeprature=1.5), lmm)
    assert dec.setup_temperature_decay(100) == 2
    seen = []
    for step in range(7):
        dec.global_step = step
        dec.decay_temperature()
        seen.append(dec._temperature_value)
    assert seen == [4.0, 4.0, 2.0, 2.0, 1.5, 1.5, 1.5] and float(dec.temperature) == 1.5
    # the schedule counts optimizer steps (Lightning's global_step in the reference)
    dec2 = VQAICVModule(Iface(), ModuleConfig(init_temperature=4.0, decay_ratio=0.5, decay_per_step=2,
                                              min_tmeprature=1.5), lmm)
    dec2.setup_temperature_decay(100)
    seen = []
    for _ in range(5):
        dec2.decay_temperature()
        seen.append(dec2._temperature_value)
        dec2.on_optimizer_step()
    assert seen == [4.0, 4.0, 2.0, 2.0, 1.5]
    # a decay ratio without a period is a configuration error, said clearly
    bad = VQAICVModule(Iface(), ModuleConfig(init_temperature=4.0, decay_ratio=0.5, decay_per_step=-1), lmm)
    with pytest.raises(RuntimeError, match="decay period"):
        bad.decay_temperature()
    # learnable_t (icv_module.py:49-52): the temperature is a trainable parameter
    lt = VQAICVModule(Iface(), ModuleConfig(learnable_t=True, init_temperature=2.0), lmm)
    assert lt.temperature.requires_grad and float(lt.temperature) == 2.0
    assert "temperature" in [k for k, p in lt.named_parameters() if p.requires_grad]


def test_collator_contract_from_token_ids():
    """f4: the four-key batch of icv_datamodule.py:125-130 from pre-tokenised samples, and the
    invariant that both get_mask calls select the same rows."""
    from licv_vqa_b200.collate import check_batch_contract, collate_token_ids
    from oracle import licv_oracle as O
    BOS, EOS, PAD = 1, 2, 0
    query_x = [[BOS, 11, 12, 13], [BOS, 21, 22]]                      # question, no answer
    query = [[BOS, 11, 12, 13, 90, 91], [BOS, 21, 22, 95]]            # question + answer
    ice = [[BOS, 50, 51, 52, 53, 54], [BOS, 60, 61]]                  # in-context examples
    batch = collate_token_ids(query, query_x, ice, PAD, BOS, EOS)
    q, t = batch["query_inputs"]["input_ids"], batch["inputs"]["input_ids"]
    assert q.tolist() == [[1, 11, 12, 13, 90, 91, 2], [1, 21, 22, 95, 2, 0, 0]]
    assert t[0].tolist() == [1, 50, 51, 52, 53, 54, 11, 12, 13, 90, 91, 2]
    assert t[1].tolist() == [1, 60, 61, 21, 22, 95, 2, 0, 0, 0, 0, 0]
    assert batch["query_inputs"]["attention_mask"].tolist() == [[1] * 7, [1, 1, 1, 1, 1, 0, 0]]
    assert batch["query_x_length"].tolist() == [4, 3]                 # non-pad tokens of query_x
    assert batch["in_context_length"].tolist() == [6 + 3, 3 + 2]      # ICE + query_x without its BOS
    assert check_batch_contract(batch, PAD) == 3 + 2                  # answer tokens + EOS per sample
    # the same selection the oracle's get_mask / pair_rows make
    sm = O.get_mask(q.numpy(), batch["query_x_length"].numpy(), PAD)
    tm = O.get_mask(t.numpy(), batch["in_context_length"].numpy(), PAD)
    ktr = O.pair_rows(sm, tm)
    assert (ktr >= 0).sum() == 5
    flat_t = t.reshape(-1).numpy()
    assert flat_t[ktr[ktr >= 0]].tolist() == q.reshape(-1).numpy()[ktr >= 0].tolist()   # same tokens
    # a broken batch is refused
    bad = dict(batch, in_context_length=batch["in_context_length"] + 1)
    with pytest.raises(ValueError, match="rows"):
        check_batch_contract(bad, PAD)
    with pytest.raises(ValueError, match="missing"):
        check_batch_contract({"inputs": batch["inputs"]}, PAD)
    left = collate_token_ids(query, query_x, ice, PAD, BOS, EOS, padding_side="left")
    assert left["query_inputs"]["input_ids"][1].tolist() == [0, 0, 1, 21, 22, 95, 2]


def test_kd_loss_plan_fast_kernel_for_reference_vocabularies():
    """Every reference model's vocabulary (idefics-9b 32002, idefics2-8b 32003, config 1's 32000)
    runs on the stream kernel for 16-bit logits - also KL + CE rows at T != 1 (decay_temperature,
    icv_module.py:150-158), which round 1 left to the generic kernel."""
    from licv_vqa_b200 import _abi
    lib = _abi.load()
    for V in (32000, 32002, 32003):
        for code in (_abi.BF16, _abi.F16):
            for T in (1.0, 0.5, 2.0):
                assert lib.licv_kd_loss_plan(V, code, T, 1, 8192) == _abi.KD_KERNEL_STREAM
            # a handful of rows per SM: the tensor-memory kernel where it can (T = 1 or one loss)
            assert lib.licv_kd_loss_plan(V, code, 1.0, 1, 256) == _abi.KD_KERNEL_TMEM
            assert lib.licv_kd_loss_plan(V, code, 2.0, 0, 256) == _abi.KD_KERNEL_TMEM
            assert lib.licv_kd_loss_plan(V, code, 2.0, 1, 256) == _abi.KD_KERNEL_STREAM
        assert lib.licv_kd_loss_plan(V, _abi.F32, 1.0, 1, 8192) == _abi.KD_KERNEL_CLUSTER
    assert lib.licv_kd_loss_plan(1003, _abi.BF16, 1.0, 1, 8192) == _abi.KD_KERNEL_TMEM
    assert lib.licv_kd_loss_plan(50257, _abi.BF16, 1.0, 1, 8192) == _abi.KD_KERNEL_CLUSTER
    assert lib.licv_kd_loss_plan(50257, _abi.BF16, 2.0, 1, 8192) == _abi.KD_KERNEL_GENERIC
    assert lib.licv_kd_loss_plan(300000, _abi.BF16, 1.0, 0, 8192) == _abi.KD_KERNEL_GENERIC
    assert lib.licv_kd_loss_plan(0, _abi.BF16, 1.0, 0, 8192) < 0


def test_collator_matches_the_references_collator_over_a_real_tokenizer():
    """f4 pinned: tests/golden/collator_cases.npz is the output of the reference's own, unmodified
    collator_data (icv_datamodule.py:73-130) driven by a real transformers fast tokenizer (built
    in process by oracle/make_golden.py: LLaMA / idefics-style BOS, EOS, pad = <unk>, image
    tokens; right and left padding).  collate_token_ids, fed the per-sample id lists a dataset
    that tokenises each part once would hold, must reproduce all four outputs bit for bit."""
    import numpy as np
    from licv_vqa_b200.collate import check_batch_contract, collate_token_ids
    from tests.util import load_golden
    G = load_golden("collator_cases.npz")
    for name in [str(n) for n in G["names"]]:
        pad, bos, eos = [int(x) for x in G[f"{name}/special_ids"]]
        n = int(G[f"{name}/n_samples"])
        q = [G[f"{name}/sample{b}/query_ids"].tolist() for b in range(n)]
        qx = [G[f"{name}/sample{b}/query_x_ids"].tolist() for b in range(n)]
        ice = [G[f"{name}/sample{b}/ice_ids"].tolist() for b in range(n)]
        side = "left" if name.endswith("left") else "right"
        batch = collate_token_ids(q, qx, ice, pad, bos, eos, padding_side=side)
        assert np.array_equal(batch["query_inputs"]["input_ids"].numpy(), G[f"{name}/q_ids"])
        assert np.array_equal(batch["query_inputs"]["attention_mask"].numpy(), G[f"{name}/q_att"])
        assert np.array_equal(batch["inputs"]["input_ids"].numpy(), G[f"{name}/t_ids"])
        assert np.array_equal(batch["inputs"]["attention_mask"].numpy(), G[f"{name}/t_att"])
        assert np.array_equal(batch["in_context_length"].numpy(), G[f"{name}/in_context_length"])
        assert np.array_equal(batch["query_x_length"].numpy(), G[f"{name}/query_x_length"])
        if side == "right":
            # the lengths are token COUNTS used as positions (icv_module.py:136-148): with right
            # padding both masks select the answer tokens + EOS of every sample
            n_rows = check_batch_contract(batch, pad)
            want = sum(len(a) - len(b) + 1 for a, b in zip(q, qx))
            assert n_rows == want


def test_torch_library_ops_are_registered_with_fake_kernels():
    """north_star: the kernels sit behind registered torch operators.  Schemas exist and the fake
    (meta) kernels give the shapes a trace needs - no GPU, no compute call."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from licv_vqa_b200 import torch_ops
    for name in torch_ops.REGISTERED:
        assert hasattr(torch.ops.licv, name)
    assert "ScalarType out_dtype" in str(torch.ops.licv.inject.default._schema)
    with FakeTensorMode():
        h = torch.empty(8, 32, 4096, dtype=torch.bfloat16, device="cuda")
        out = torch.ops.licv.inject(h, torch.empty(4096, device="cuda"), torch.float32, 0)
        assert out.shape == h.shape and out.dtype == torch.float32
        dh, ds = torch.ops.licv.inject_bwd(h, h, torch.empty(4096, device="cuda"), 0)
        assert dh.shape == h.shape and ds.shape == (4096,) and ds.dtype == torch.float32
        stu = torch.empty(64, 32002, dtype=torch.float16, device="cuda")
        tot, kl, ce, d = torch.ops.licv.kd_loss(stu, stu, None, None, None, 64, 0, 1.0, 1e-6, 0.0, False, 16)
        assert tot.shape == () and d.shape == stu.shape and d.dtype == stu.dtype


def test_registered_ops_have_no_cpu_kernel_and_a_cpp_autograd_formula():
    """The operators come from C++ (csrc/licv_torch_ops.cpp: CUDA + Meta + Autograd keys).  A CPU
    tensor fails in the dispatcher - there is no CPU path to fall back to - and the C++ autograd
    formula is what a fake-tensor trace records (C++ nodes; shapes and dtypes of dh / d_shift / d stu)."""
    import pytest
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from licv_vqa_b200 import torch_ops  # noqa: F401
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.licv.inject(torch.zeros(2, 64), torch.zeros(64), torch.float32, 0)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.licv.get_mask(torch.zeros(2, 8, dtype=torch.long), torch.zeros(2, dtype=torch.long), 0)
    with FakeTensorMode():
        h = torch.empty(4, 16, 4096, dtype=torch.bfloat16, device="cuda", requires_grad=True)
        shift = torch.empty(4096, device="cuda", requires_grad=True)
        out = torch.ops.licv.inject(h, shift, torch.bfloat16, 0)
        assert out.requires_grad and "InjectFn" in out.grad_fn.name()      # the C++ node, not a Python one
        # (running the backward needs the engine's CUDA device thread: tests/test_gpu_torch_ops.py)
        dh, ds = torch.ops.licv.inject_bwd(h.detach(), out.detach(), shift.detach(), 0)
        assert dh.shape == h.shape and dh.dtype == torch.bfloat16
        assert ds.shape == shift.shape and ds.dtype == torch.float32
        stu = torch.empty(6, 32002, dtype=torch.float16, device="cuda", requires_grad=True)
        lab = torch.empty(6, dtype=torch.long, device="cuda")
        total, kl, ce, dstu = torch.ops.licv.kd_loss(stu, stu.detach(), None, lab, None, 6, 6, 1.0, 1e-6, 0.5,
                                                     False, 16)
        assert total.requires_grad and not kl.requires_grad and not dstu.requires_grad
        assert "KdLossFn" in total.grad_fn.name()
        assert dstu.shape == stu.shape and dstu.dtype == torch.float16
