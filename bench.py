#!/usr/bin/env python
"""bench.py - L-ICV hot path on B200: train samples/s, kernel roofline, CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One STEP = one pass of the hot path over one batch of BASELINE.json configs[1] (idefics-9B shape:
32 hooked layers, d = 4096, V = 32002, bs = 8 per GPU, 32 student tokens per sample, fp16 under
the README's DeepSpeed "16-mixed" recipe, 4 answer tokens per sample as KL rows, every shifted
non-pad position a CE row, hard_loss_weight = 0.5), on synthetic tensors:

    icv = sigmoid?(alpha) * v                                  licv_icv_scale
    32 x  out_l = inject(h_l, icv_l)                           licv_inject_fwd
    row pairing + labels, KL + 0.5 CE fwd+bwd on the logits     licv_kd_prepare_rows, licv_kd_loss_fwd_bwd
    32 x  dh_l, replicas_l += inject_bwd(h_l, g_l, icv_l)       licv_inject_bwd_spread
    d_icv = sum of the replicas; d_v, d_alpha; squared norms    licv_icv_grad_finish
    (N > 1) exchange of the flat ICV gradient (131 104 fp32) fused with, (N = 1) just
    clip + AdamW on the flat ICV parameters                     licv_dp_allreduce_adamw / licv_adamw_step_partials

The frozen tower's GEMMs that produce h_l, g_l and the logits are stock cuBLAS and out of scope
(BASELINE.json north_star); they are replaced here by resident synthetic tensors, a distinct
buffer per layer so one step streams ~290 MB (> 126 MB L2) through the kernels.

`value`  : samples/s over all ranks, inputs resident in HBM, whole step replayed as a CUDA graph.
`e2e`    : the same step through the host-buffer C-ABI entry points (licv_*_host): every input
           starts in pinned host memory and every result ends there.
`roofline`: the dominant kernel of the step (licv_inject_bwd), CUDA-event timed per launch.
`roofline_bw`: the same kernels at the bandwidth-bound shapes of configs[4] (inference sweep).
`cpu_baseline`: oracle/torch_chain.py (the reference's eager op chain) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "train samples/s (L-ICV hot path: ICV inject fwd/bwd x32 layers + KL/CE fwd/bwd)"
UNIT = "samples/s"

CFG = dict(workload="configs[1]: idefics-9B shape, VQAv2 32-shot teacher vs zero-shot+ICV student",
           layers=32, d=4096, vocab=32002, batch_per_gpu=8, student_tokens=32, kl_rows_per_sample=4,
           teacher_rows=32, hard_loss_weight=0.5, temperature=1.0, kl_eps=1e-6, use_sigmoid=False,
           alpha_init_value=0.1, dtype="fp16", recipe="DeepSpeed 16-mixed (fp16 ICV, no autocast)")


def config_of(n_gpus):
    """The `config` of the JSON line: the SAME dict for this arm and for --impl reference."""
    return dict(CFG, global_batch=CFG["batch_per_gpu"] * n_gpus, parallelism=f"dp{n_gpus}",
                flush="inputs of one step (~290 MB, a distinct buffer per layer) exceed L2")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------
# synthetic batch
# ------------------------------------------------------------------------------------------------
def make_batch(device, seed, dtype, pinned_host=False):
    L, d, V = CFG["layers"], CFG["d"], CFG["vocab"]
    B, T = CFG["batch_per_gpu"], CFG["student_tokens"]
    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = "cpu" if pinned_host else device

    def mk(*shape, scale=1.0, dt=dtype):
        t = (torch.randn(*shape, generator=g) * scale).to(dt)
        if pinned_host:
            return t.pin_memory()
        return t.to(device)

    batch = {}
    batch["h"] = [mk(B * T, d, scale=float(1 + 29 * l / (L - 1))) for l in range(L)]
    for h in batch["h"]:   # two "massive activation" channels, LLaMA-style
        h[:, 7] *= 50
        h[:, d // 3] *= -20
    batch["g"] = [mk(B * T, d, scale=1e-2) for _ in range(L)]
    stu = torch.randn(B * T, V, generator=g) * 3
    tea = torch.randn(B * CFG["kl_rows_per_sample"], V, generator=g) * 3
    # token ids: BOS, text, 4 answer tokens at the end of each sample; no padding in this batch
    ids = torch.randint(3, V, (B, T), generator=g)
    ids[:, 0] = 1
    qx = torch.full((B,), T - CFG["kl_rows_per_sample"], dtype=torch.long)
    # teacher sequence: only its answer rows are kept (the tower is out of scope): [B, 4]
    t_ids = ids[:, T - CFG["kl_rows_per_sample"]:].clone()
    t_len = torch.zeros(B, dtype=torch.long)
    # +10 spike on the label in half of the KL rows
    for b in range(B):
        for k in range(CFG["kl_rows_per_sample"]):
            r = b * T + (T - CFG["kl_rows_per_sample"]) + k
            j = int(torch.randint(0, V, (1,), generator=g))
            tea[b * CFG["kl_rows_per_sample"] + k, j] += 10
            if k % 2:
                stu[r, j] += 8
    batch["stu"] = stu.to(dtype).pin_memory() if pinned_host else stu.to(dtype).to(device)
    batch["tea"] = tea.to(dtype).pin_memory() if pinned_host else tea.to(dtype).to(device)
    for k, v in (("ids", ids), ("qx", qx), ("t_ids", t_ids), ("t_len", t_len),
                 ("att", torch.ones(B, T, dtype=torch.long))):
        batch[k] = v.to(dev)
    return batch


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# the step, on resident tensors (C ABI launches on torch's current stream)
# ------------------------------------------------------------------------------------------------
class HotPath:
    @property
    def launches_per_step(self):
        """icv_scale, 32 x fwd, row prep, loss, 32 x bwd, the tail (1 fused / 2 split), the optimizer
        (N = 1: AdamW alone after the fused tail, else sum of squares / exchange + AdamW); NCCL's own
        kernel comes on top when that path is selected at N > 1."""
        tail = 2 if self.split_tail else 1
        opt = 1 if (self.world == 1 and not self.split_tail) else 2
        return 1 + 32 + 1 + 1 + 32 + tail + opt

    def __init__(self, device, dtype, world):
        from licv_vqa_b200 import _abi, ops
        self.abi, self.ops, self.lib = _abi, ops, _abi.load()
        self.device, self.dtype, self.world = device, dtype, world
        self.code = {torch.float16: _abi.F16, torch.bfloat16: _abi.BF16,
                     torch.float32: _abi.F32}[dtype]
        L, d, V = CFG["layers"], CFG["d"], CFG["vocab"]
        B, T = CFG["batch_per_gpu"], CFG["student_tokens"]
        f32 = dict(dtype=torch.float32, device=device)
        # flat parameter / gradient / moment buffers: [vec (L*d) | alpha (L)]
        self.n_vec, self.n_alpha = L * d, L
        self.param = torch.empty(L * d + L, **f32)
        g = torch.Generator(device="cpu").manual_seed(426)
        self.param[:L * d] = (torch.randn(L * d, generator=g) * 0.01).to(device)
        self.param[L * d:] = CFG["alpha_init_value"]
        self.grad = torch.zeros(L * d + L + 4, **f32)      # + logged scalars ride along
        self.m = torch.zeros(L * d + L, **f32)
        self.v = torch.zeros(L * d + L, **f32)
        self.icv = torch.empty(L, d, **f32)
        self.sink = torch.zeros(L, d, **f32)
        # N > 1: the loss kernel writes (kl, ce, total) straight into the tail of the flat gradient
        # buffer, so the logged scalars ride along in the one exchange without a copy launch
        self.losses = self.grad[L * d + L:] if world > 1 else torch.zeros(4, **f32)
        self.norm = torch.zeros(1, **f32)
        self.norm_partials = torch.zeros(L, **f32)
        self.split_tail = env_int("LICV_BENCH_SPLIT_TAIL", 0) != 0
        self.kl_tea_row = torch.empty(B * T, dtype=torch.int32, device=device)
        self.ce_label = torch.empty(B * T, dtype=torch.int64, device=device)
        self.counts = torch.zeros(4, dtype=torch.int32, device=device)
        self.ws = torch.zeros(self.lib.licv_kd_loss_workspace_bytes(B * T) + 64, dtype=torch.uint8,
                              device=device)
        self.opt_ws = torch.zeros(16, dtype=torch.uint8, device=device)
        # the backward launches add their d_shift into replicas of the [L, d] gradient (zero between steps)
        self.n_rows = (env_int("LICV_BENCH_ROWS", 0) or
                       self.lib.licv_inject_bwd_rows(B * T, d, self.code, self.code))
        self.rows = torch.zeros(L, self.n_rows, d, **f32)
        self.out = [torch.empty(B * T, d, dtype=dtype, device=device) for _ in range(L)]
        self.dh = [torch.empty(B * T, d, dtype=dtype, device=device) for _ in range(L)]
        self.dstu = torch.empty(B * T, V, dtype=dtype, device=device)
        # the DeepSpeed recipe's fp16/bf16 chain: every op of the reference rounds
        self.flags = (_abi.ROUND_Y | _abi.ROUND_NH | _abi.ROUND_NY | _abi.ROUND_T
                      if dtype != torch.float32 else 0)
        self.step_no = 0
        # N > 1: the gradient exchange is fused with the optimizer step over NVLink peer memory
        # (csrc/licv_dp.cu); LICV_DP_EXCHANGE=nccl selects all_reduce + optimizer kernels instead
        self.peer, self.peer_note = None, ""
        if world > 1 and os.environ.get("LICV_DP_EXCHANGE", "p2p") != "nccl":
            from licv_vqa_b200.dp import PeerExchange
            ok = torch.ones(1, device=device)
            try:
                self.peer = PeerExchange(self.grad.numel())
            except Exception as exc:       # peers not mappable: every rank must take the same path
                self.peer_note = f" (p2p unavailable: {exc})"
                ok.zero_()
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
            if float(ok) == 0.0:
                self.peer = None

    def _chk(self, rc, what):
        if rc != 0:
            self.abi.check(rc, what)

    def compute(self, batch):
        """icv -> 32 x inject fwd -> row prep + loss fwd/bwd -> 32 x inject bwd -> d_vec, d_alpha
        (this rank's gradient in self.grad, logged scalars in its tail)."""
        lib, st = self.lib, torch.cuda.current_stream().cuda_stream
        L, d, V = CFG["layers"], CFG["d"], CFG["vocab"]
        B, T = CFG["batch_per_gpu"], CFG["student_tokens"]
        n_tok = B * T
        p = self.param.data_ptr()
        alpha_p, vec_p = p + 4 * self.n_vec, p
        self._chk(lib.licv_icv_scale(alpha_p, vec_p, self.icv.data_ptr(), L, d,
                                     int(CFG["use_sigmoid"]), st), "icv_scale")
        for l in range(L):
            self._chk(lib.licv_inject_fwd(batch["h"][l].data_ptr(), self.icv[l].data_ptr(),
                                          self.out[l].data_ptr(), n_tok, d, self.code, self.code,
                                          self.flags, st), "inject_fwd")
        self._chk(lib.licv_kd_prepare_rows(
            batch["ids"].data_ptr(), batch["qx"].data_ptr(), batch["att"].data_ptr(),
            batch["t_ids"].data_ptr(), batch["t_len"].data_ptr(), 0, -1, 0, B, T,
            CFG["kl_rows_per_sample"], self.kl_tea_row.data_ptr(), self.ce_label.data_ptr(),
            self.counts.data_ptr(), st), "kd_prepare_rows")
        self._chk(lib.licv_kd_loss_fwd_bwd(
            batch["stu"].data_ptr(), self.dstu.data_ptr(), batch["tea"].data_ptr(),
            self.kl_tea_row.data_ptr(), self.ce_label.data_ptr(), self.counts.data_ptr(), 0, 0,
            CFG["temperature"], CFG["kl_eps"], CFG["hard_loss_weight"], 0, 1.0,
            self.losses.data_ptr(), self.ws.data_ptr(), n_tok, V, V, V, self.code,
            self.abi.ROUND_TEMPERED, st), "kd_loss")
        R = self.n_rows
        for l in reversed(range(L)):
            self._chk(lib.licv_inject_bwd_spread(
                batch["h"][l].data_ptr(), batch["g"][l].data_ptr(), self.icv[l].data_ptr(),
                self.dh[l].data_ptr(), self.rows[l].data_ptr(), R, n_tok, d, self.code, self.code,
                self.flags, st), "inject_bwd_spread")
        # ONE launch: d_icv = sum of the replicas (left zero for the next step), d_vec, d_alpha and
        # the per-layer squared norms the optimizer clips by (LICV_BENCH_SPLIT_TAIL=1: the separate
        # licv_reduce_rows + licv_icv_scale_bwd launches, the optimizer sums the squares itself)
        g = self.grad.data_ptr()
        if self.split_tail:
            self._chk(lib.licv_reduce_rows(self.rows.data_ptr(), self.sink.data_ptr(), L, R, R * d, d,
                                           0, 1, st), "reduce_rows")
            self._chk(lib.licv_icv_scale_bwd(alpha_p, vec_p, self.sink.data_ptr(), g,
                                             g + 4 * self.n_vec, L, d, int(CFG["use_sigmoid"]), st),
                      "icv_scale_bwd")
            return
        self._chk(lib.licv_icv_grad_finish(self.rows.data_ptr(), R, R * d, alpha_p, vec_p,
                                           self.sink.data_ptr(), g, g + 4 * self.n_vec,
                                           self.norm_partials.data_ptr(), 1.0 / self.world, L, d,
                                           int(CFG["use_sigmoid"]), 0, 1, st), "icv_grad_finish")

    def optimize(self):
        """(N > 1: exchange of the flat gradient, fused with) clip + AdamW."""
        lib, st = self.lib, torch.cuda.current_stream().cuda_stream
        p, g = self.param.data_ptr(), self.grad.data_ptr()
        self.step_no += 1
        if self.world > 1 and self.peer is not None:
            self._chk(lib.licv_dp_allreduce_adamw(
                self.peer.comm, p, g, self.m.data_ptr(), self.v.data_ptr(), self.n_vec,
                self.n_alpha, 4, 1e-4, 1e-2, 0.9, 0.999, 1e-8, 1e-3, self.step_no, 1.0,
                self.norm.data_ptr(), self.opt_ws.data_ptr(), st), "dp_allreduce_adamw")
            return
        if self.world > 1:
            torch.distributed.all_reduce(self.grad)
        if self.world == 1 and not self.split_tail:
            # the squared norm comes per layer from licv_icv_grad_finish: no sum-of-squares launch
            self._chk(lib.licv_adamw_step_partials(
                p, g, self.m.data_ptr(), self.v.data_ptr(), self.n_vec, self.n_alpha, 1e-4, 1e-2, 0.9,
                0.999, 1e-8, 1e-3, self.step_no, 1.0, 1.0, self.norm.data_ptr(),
                self.opt_ws.data_ptr(), self.norm_partials.data_ptr(), CFG["layers"], st),
                "adamw_step_partials")
            return
        self._chk(lib.licv_adamw_step(p, g, self.m.data_ptr(), self.v.data_ptr(), self.n_vec,
                                      self.n_alpha, 1e-4, 1e-2, 0.9, 0.999, 1e-8, 1e-3,
                                      self.step_no, 1.0 / self.world, 1.0, self.norm.data_ptr(),
                                      self.opt_ws.data_ptr(), st), "adamw_step")

    def step(self, batch):
        self.compute(batch)
        self.optimize()


def time_kernel_launches(fn_list, stream):
    """CUDA-event time around each launch (events on the launching stream) -> list of seconds."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in fn_list]
    for (e0, e1), fn in zip(evs, fn_list):
        e0.record(stream)
        fn()
        e1.record(stream)
    torch.cuda.synchronize()
    return [e0.elapsed_time(e1) * 1e-3 for e0, e1 in evs]


def load_traffic(key):
    """dram__bytes_read + dram__bytes_write per launch of `key` from the committed ncu capture
    (profiles/r1_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)[key]
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


def load_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def esize(dtype):
    return 4 if dtype == torch.float32 else 2


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's eager op chain on the host cores
# ------------------------------------------------------------------------------------------------
def reference_chain_inputs(device, seed=0):
    """The batch of one step in the reference's own recipe - DeepSpeed "16-mixed": hidden states,
    incoming gradients, logits AND the ICV parameters in fp16 (README.md:126-190, zero2.yaml) -
    i.e. exactly the dtypes the B200 arm moves (2-byte h / g / out / dh / logits)."""
    from oracle import torch_chain  # noqa: F401  (the chain itself; imported by the callers)
    dt = torch.float16
    batch = make_batch(device, seed, dt)
    L, d, V = CFG["layers"], CFG["d"], CFG["vocab"]
    B, T = CFG["batch_per_gpu"], CFG["student_tokens"]
    alpha = torch.full((1, L), CFG["alpha_init_value"]).to(dt).to(device)
    gen = torch.Generator().manual_seed(426)
    vec = (torch.randn(1, L, d, generator=gen) * 0.01).to(dt).to(device)
    hs = [h.view(B, T, d) for h in batch["h"]]
    gs = [g.view(B, T, d) for g in batch["g"]]
    stu_mask = torch.zeros(B, T, dtype=torch.bool, device=device)
    stu_mask[:, T - CFG["kl_rows_per_sample"]:] = True
    temp = torch.tensor(CFG["temperature"], device=device)
    return dict(hs=hs, gs=gs, alpha=alpha, vec=vec, stu=batch["stu"].view(B, T, V), tea=batch["tea"],
                stu_mask=stu_mask, ids=batch["ids"], att=batch["att"], temp=temp)


def reference_chain_step(x):
    from oracle import torch_chain
    return torch_chain.hot_path_step(x["hs"], x["gs"], x["alpha"], x["vec"], CFG["use_sigmoid"],
                                     x["stu"], x["tea"], x["stu_mask"], x["ids"], x["att"], x["temp"],
                                     CFG["kl_eps"], CFG["hard_loss_weight"])


REFERENCE_RECIPE = ("oracle/torch_chain.py: the reference's eager op chain + autograd (icv_intervention.py:"
                    "61-86, icv_module.py:89-134, HF shifted CE), fp16 hidden states / gradients / logits / "
                    "ICV parameters like the B200 arm (DeepSpeed 16-mixed)")


def cpu_reference_steps(n_steps, warmup, seed=0):
    torch.set_num_threads(os.cpu_count() or 1)
    x = reference_chain_inputs("cpu", seed)
    for _ in range(warmup):
        reference_chain_step(x)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        reference_chain_step(x)
    dt_s = (time.perf_counter() - t0) / max(n_steps, 1)
    return dt_s, torch.get_num_threads()


def eager_cuda_steps(device, n_steps=10, warmup=3):
    """The same eager chain on the same B200 (stock PyTorch kernels): SURVEY 8(d)'s "real
    comparator".  -> seconds per step, CUDA events on the current stream."""
    x = reference_chain_inputs(device, 0)
    for _ in range(warmup):
        reference_chain_step(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_steps):
        reference_chain_step(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / n_steps


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    steps = max(1, args.steps)      # one step = 8 samples = ~50 ms of host work: K steps as asked
    warm = max(0, args.warmup)
    sec, cores = cpu_reference_steps(steps, warm)
    val = CFG["batch_per_gpu"] / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": CFG["dtype"] + " (fp32 accumulate)",
        "data": "synthetic", "config": config_of(args.gpus),
        "arm": {"recipe": REFERENCE_RECIPE, "where": "host CPU, rank 0 only (one process, one "
                "8-sample batch per step; world-size independent)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} full steps of the same workload (one 8-sample batch "
                                   "each) through " + REFERENCE_RECIPE},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# correctness evidence for the JSON line
# ------------------------------------------------------------------------------------------------
def dp_check(hp, batch):
    """One warm-up step at N > 1, twice: through the exchange this run uses, and - on copies of the
    same state - through torch.distributed.all_reduce + licv_adamw_step.  Also: every replica holds
    bit-identical parameters afterwards, and no wait for a peer timed out."""
    dist = torch.distributed
    hp.compute(batch)
    torch.cuda.synchronize()
    n = hp.n_vec + hp.n_alpha
    p2, g2, m2, v2 = hp.param.clone(), hp.grad.clone(), hp.m.clone(), hp.v.clone()
    norm2, ws2 = torch.zeros(1, device=hp.device), torch.zeros(16, dtype=torch.uint8, device=hp.device)
    hp.optimize()
    dist.all_reduce(g2)
    hp._chk(hp.lib.licv_adamw_step(p2.data_ptr(), g2.data_ptr(), m2.data_ptr(), v2.data_ptr(), hp.n_vec,
                                   hp.n_alpha, 1e-4, 1e-2, 0.9, 0.999, 1e-8, 1e-3, hp.step_no,
                                   1.0 / hp.world, 1.0, norm2.data_ptr(), ws2.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream), "adamw_step (check)")
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

    bits = hp.param.view(torch.int32).to(torch.int64)
    chk = torch.stack([bits.sum(), (bits * torch.arange(1, bits.numel() + 1, device=hp.device)).sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out = {"exchange_vs_nccl_grad_rel_err": rel(hp.grad[:n + 3], g2[:n + 3]),
           "exchange_vs_nccl_param_rel_err": rel(hp.param, p2),
           "exchange_vs_nccl_grad_norm_rel_err": rel(hp.norm, norm2),
           "replicas_bit_identical": bool(torch.equal(lo, hi)),
           "comm_error": (int(hp.peer.timed_out()) if hp.peer is not None else 0)}
    t = torch.tensor([out["exchange_vs_nccl_grad_rel_err"], out["exchange_vs_nccl_param_rel_err"],
                      out["exchange_vs_nccl_grad_norm_rel_err"], float(out["comm_error"])], device=hp.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)      # worst rank
    out["exchange_vs_nccl_grad_rel_err"], out["exchange_vs_nccl_param_rel_err"] = float(t[0]), float(t[1])
    out["exchange_vs_nccl_grad_norm_rel_err"], out["comm_error"] = float(t[2]), int(t[3])
    return out


def oracle_check(hp, batch):
    """The timed step's arithmetic against the float64 oracle (oracle/licv_oracle.py, the checker
    that tests/ pins to the reference's golden vectors) on the same batch: losses, d_icv of every
    layer, d_vec / d_alpha, sampled rows of d(logits).  CPU, outside every timed region."""
    import numpy as np
    from oracle import licv_oracle as O
    L, d, V = CFG["layers"], CFG["d"], CFG["vocab"]
    B, T, K4 = CFG["batch_per_gpu"], CFG["student_tokens"], CFG["kl_rows_per_sample"]
    hp.compute(batch)
    torch.cuda.synchronize()

    def host(t):
        return t.detach().float().cpu().numpy().astype(np.float64)

    def rel(a, b):
        den = np.linalg.norm(np.asarray(b).ravel())
        return float(np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / (den if den > 0 else 1.0))

    icv = host(hp.icv)
    worst_ds, worst_dh = 0.0, 0.0
    for l in range(L):
        o_dh, o_ds = O.inject_bwd(host(batch["h"][l]), icv[l], host(batch["g"][l]), hp.flags, CFG["dtype"])
        worst_ds = max(worst_ds, rel(host(hp.sink[l]), o_ds))
        if l in (0, L - 1):
            worst_dh = max(worst_dh, rel(host(hp.dh[l]), o_dh))
    alpha = host(hp.param[hp.n_vec:hp.n_vec + hp.n_alpha])
    vec = host(hp.param[:hp.n_vec]).reshape(L, d)
    sink = host(hp.sink)
    d_vec_o, d_alpha_o = alpha[:, None] * sink, (sink * vec).sum(1)
    ktr = np.full(B * T, -1, np.int32)
    lab = np.full(B * T, -100, np.int64)
    ids = batch["ids"].cpu().numpy()
    for b in range(B):
        for k in range(K4):
            ktr[b * T + T - K4 + k] = b * K4 + k
        lab[b * T:b * T + T - 1] = ids[b, 1:]
    want = O.kd_loss_rows(host(batch["stu"]), host(batch["tea"]), ktr, lab, CFG["temperature"],
                          CFG["kl_eps"], CFG["hard_loss_weight"], logit_fmt=CFG["dtype"])
    got = [float(x) for x in hp.losses[:3].cpu()]
    rows = [0, T - 2, T - 1, B * T - 3, B * T - 1]
    extra = {}
    if not hp.split_tail:   # the squared norm licv_icv_grad_finish hands to the optimizer
        want_sq = (np.square(d_vec_o).sum() + np.square(d_alpha_o).sum()) / hp.world ** 2
        extra["grad_sq_norm_rel_err"] = abs(float(host(hp.norm_partials).sum()) - want_sq) / want_sq
    return {**extra, "loss_rel_err": abs(got[2] - want["loss"]) / abs(want["loss"]),
            "kl_rel_err": abs(got[0] - want["kl"]) / abs(want["kl"]),
            "ce_rel_err": abs(got[1] - want["ce"]) / abs(want["ce"]),
            "d_icv_rel_err_worst_layer": worst_ds, "dh_rel_err": worst_dh,
            "d_vec_rel_err": rel(host(hp.grad[:hp.n_vec]).reshape(L, d), d_vec_o),
            "d_alpha_rel_err": rel(host(hp.grad[hp.n_vec:hp.n_vec + hp.n_alpha]), d_alpha_o),
            "dstu_rows_rel_err": rel(host(hp.dstu[rows]), want["d_stu"][rows]),
            "against": "oracle/licv_oracle.py (float64, same fp16 inputs and rounding points); "
                       "tolerances of tests/: losses 1e-5, d_icv / d_vec / d_alpha 1e-4, dh and "
                       "d(logits) the fp16 roundoff 6e-4"}


# ------------------------------------------------------------------------------------------------
# main arm
# ------------------------------------------------------------------------------------------------
def bandwidth_shapes(hp, peak):
    """configs[4]-sized launches of the same kernels: bandwidth-bound roofline points.

    Every shape is measured on TWO freshly allocated buffer sets (torch.cuda.empty_cache() first)
    and the faster one is reported, all samples listed: on this pool a buffer set occasionally
    lands on memory where EVERY kernel - torch's own copy included - runs 2-4x slower for the
    lifetime of the allocation (profiles/README.md, r2q); that is a property of the placement,
    not of the kernel."""
    lib, st = hp.lib, torch.cuda.current_stream().cuda_stream
    out = []
    d, V = CFG["d"], CFG["vocab"]
    dt, code, es = hp.dtype, hp.code, esize(hp.dtype)
    n_tok = 64 * 2048                                   # bs 64 x T 2048, 1 GiB in

    def measure(name, shape, nbytes, make, launch, n_launch, n_warm):
        samples = []
        for _ in range(2):
            torch.cuda.empty_cache()
            bufs = make()
            for i in range(n_warm):
                launch(bufs, i)
            ts = time_kernel_launches([lambda i=i: launch(bufs, i) for i in range(n_launch)],
                                      torch.cuda.current_stream())
            samples.append(sum(ts) / len(ts))
            del bufs
        t = min(samples)
        out.append({"kernel": name, "shape": shape, "achieved": nbytes / t / 1e9, "peak": peak,
                    "unit": "GB/s", "frac": nbytes / t / 1e9 / peak,
                    "frac_of_nominal_8tbs": nbytes / t / 8e12, "us": t * 1e6,
                    "us_per_allocation": [round(x * 1e6, 1) for x in samples]})

    s = hp.icv[0]
    ds = torch.zeros(d, device=hp.device)

    def make_inject():
        return ([(torch.randn(n_tok, d, device=hp.device) * 4).to(dt) for _ in range(2)],
                [torch.randn(n_tok, d, device=hp.device).to(dt) for _ in range(2)],
                torch.empty(n_tok, d, dtype=dt, device=hp.device))

    measure("licv_inject_fwd", f"n_tok={n_tok} d={d} {CFG['dtype']}", 2 * es * n_tok * d, make_inject,
            lambda b, i: lib.licv_inject_fwd(b[0][i % 2].data_ptr(), s.data_ptr(), b[2].data_ptr(), n_tok,
                                             d, code, code, hp.flags, st), 6, 3)
    measure("licv_inject_bwd", f"n_tok={n_tok} d={d} {CFG['dtype']}", 3 * es * n_tok * d, make_inject,
            lambda b, i: lib.licv_inject_bwd(b[0][i % 2].data_ptr(), b[1][i % 2].data_ptr(), s.data_ptr(),
                                             b[2].data_ptr(), ds.data_ptr(), n_tok, d, code, code,
                                             hp.flags, st), 6, 3)
    R = 8192
    lab = torch.randint(0, V, (R,), device=hp.device)
    ws = torch.zeros(lib.licv_kd_loss_workspace_bytes(R) + 64, dtype=torch.uint8, device=hp.device)
    losses = torch.zeros(4, device=hp.device)

    def make_kd():
        return ([(torch.randn(R, V, device=hp.device) * 3).to(dt) for _ in range(2)],
                [(torch.randn(R, V, device=hp.device) * 3).to(dt) for _ in range(2)],
                torch.empty(R, V, dtype=dt, device=hp.device))

    def kd(b, i):
        lib.licv_kd_loss_fwd_bwd(b[0][i % 2].data_ptr(), b[2].data_ptr(), b[1][i % 2].data_ptr(), 0,
                                 lab.data_ptr(), 0, R, R, 1.0, 1e-6, 0.5, 0, 1.0,
                                 losses.data_ptr(), ws.data_ptr(), R, V, V, V, code, 16, st)

    def kd_ce(b, i):      # only_hard_loss: every row is a CE row, no teacher
        lib.licv_kd_loss_fwd_bwd(b[0][i % 2].data_ptr(), b[2].data_ptr(), 0, 0, lab.data_ptr(), 0, 0, R,
                                 1.0, 1e-6, 0.5, 1, 1.0, losses.data_ptr(), ws.data_ptr(), R, V, V,
                                 V, code, 16, st)

    measure("licv_kd_loss_fwd_bwd", f"R={R} KL+CE rows V={V} {CFG['dtype']}", 3 * es * R * V, make_kd, kd, 4, 2)
    measure("licv_kd_loss_fwd_bwd", f"R={R} CE-only rows V={V} {CFG['dtype']}", 2 * es * R * V, make_kd,
            kd_ce, 4, 2)
    torch.cuda.empty_cache()
    return out


def bind_near_gpu(index):
    """Best effort: run this process, and place the pinned buffers it allocates, on the NUMA node
    the GPU hangs off (eight ranks that all pin on node 0 share one socket's memory and the
    inter-socket link).  -> what was done, for the JSON line."""
    note = {}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        note["gpu_numa_node"] = node
        if node < 0:
            return note
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            note["cpu_affinity"] = f"{len(use)} cpus of node {node}"
        else:
            note["cpu_affinity"] = f"node {node} cpus not in this container's cpuset ({len(allowed)} allowed)"
        # memory policy: prefer that node for what this thread allocates from here on
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))   # set_mempolicy(MPOL_PREFERRED)
        note["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy errno %d" % ctypes.get_errno()
    except Exception as exc:      # not fatal: the measurement simply runs where the process is
        note["error"] = repr(exc)[:120]
    return note


def run_e2e_host(hp, steps, warmup, seed):
    """The SAME step through the host-buffer entry points: every tensor of the step starts in
    pinned host memory and every result ends there - hidden states, gradients, logits through
    licv_*_host; the ICV product, the row lists, d_vec / d_alpha and the clip + AdamW update on
    131 104 floats are host-side work of a host plugin and run inside the timed region too."""
    import ctypes as C
    lib = hp.lib
    L, d, V = CFG["layers"], CFG["d"], CFG["vocab"]
    B, T = CFG["batch_per_gpu"], CFG["student_tokens"]
    n_tok, es = B * T, esize(hp.dtype)
    hb = make_batch(hp.device, seed, hp.dtype, pinned_host=True)
    out_h = [torch.empty(n_tok, d, dtype=hp.dtype).pin_memory() for _ in range(L)]
    dh_h = [torch.empty(n_tok, d, dtype=hp.dtype).pin_memory() for _ in range(L)]
    dstu_h = torch.empty(n_tok, V, dtype=hp.dtype).pin_memory()
    ds_h = torch.zeros(L, d).pin_memory()
    loss_h = torch.zeros(4).pin_memory()
    icv_h = torch.empty(L, d).pin_memory()
    T4 = CFG["kl_rows_per_sample"]
    ktr = torch.empty(n_tok, dtype=torch.int32).pin_memory()
    lab = torch.empty(n_tok, dtype=torch.int64).pin_memory()
    ids = hb["ids"]
    # host-resident parameters and optimizer state
    alpha = torch.full((L,), CFG["alpha_init_value"])
    vec = hp.param[:L * d].view(L, d).cpu().clone()
    m_v, v_v, m_a, v_a = torch.zeros(L, d), torch.zeros(L, d), torch.zeros(L), torch.zeros(L)
    step_no = [0]
    sess = C.c_void_p()
    scratch = max(3 * n_tok * d * es + 4 * d * 4 + 4096,
                  (n_tok + B * T4) * V * es + n_tok * 16 + 65536) + (1 << 20)
    n_slots = env_int("LICV_E2E_SLOTS", 16)
    hp.abi.check(lib.licv_host_session_create(C.byref(sess), scratch, n_slots), "host_session_create")

    def adamw(p, g, m, v, lr, coef, t):
        g = g * coef
        p.mul_(1.0 - lr * 1e-3)
        m.mul_(0.9).add_(g, alpha=0.1)
        v.mul_(0.999).addcmul_(g, g, value=0.001)
        p.addcdiv_(m, (v.sqrt() / (1.0 - 0.999 ** t) ** 0.5).add_(1e-8), value=-lr / (1.0 - 0.9 ** t))

    seg = [0.0, 0.0, 0.0, 0.0]    # LICV_E2E_SEGMENTS=1: host prologue / call issue / drain / host optimizer
    want_seg = env_int("LICV_E2E_SEGMENTS", 0) != 0
    phase = [0.0, 0.0, 0.0]       # LICV_E2E_PHASES=1: forward launches / loss / backward launches, synced apart
    want_phases = env_int("LICV_E2E_PHASES", 0) != 0

    def step():
        t_a = time.perf_counter()
        # a1 + a2: the product that crosses into the hooks; a6 + a7 + a9: row pairing and labels
        torch.mul(alpha.unsqueeze(-1), vec, out=icv_h)
        ktr.fill_(-1)
        ktr.view(B, T)[:, T - T4:] = (torch.arange(B).unsqueeze(1) * T4 + torch.arange(T4)).to(torch.int32)
        lab.fill_(-100)
        lab.view(B, T)[:, :T - 1] = ids[:, 1:]
        n_kl, n_ce = B * T4, B * (T - 1)
        t_b = time.perf_counter()
        for l in range(L):
            # the forward keeps h on the device for its backward (saved-for-backward), so the
            # backward moves only g in and dh out
            hp.abi.check(lib.licv_inject_fwd_host_save(sess, l, hb["h"][l].data_ptr(),
                                                       icv_h[l].data_ptr(), out_h[l].data_ptr(),
                                                       n_tok, d, hp.code, hp.code, hp.flags),
                         "inject_fwd_host_save")
        if want_phases:     # diagnostic only: a sync between the phases gives each phase's own link rate
            hp.abi.check(lib.licv_host_sync(sess), "host_sync")
            t_p1 = time.perf_counter()
        hp.abi.check(lib.licv_kd_loss_fwd_bwd_host(
            sess, hb["stu"].data_ptr(), dstu_h.data_ptr(), hb["tea"].data_ptr(), ktr.data_ptr(),
            lab.data_ptr(), n_kl, n_ce, CFG["temperature"], CFG["kl_eps"], CFG["hard_loss_weight"],
            0, 1.0, loss_h.data_ptr(), n_tok, B * T4, V, hp.code, 16), "kd_loss_host")
        if want_phases:
            hp.abi.check(lib.licv_host_sync(sess), "host_sync")
            t_p2 = time.perf_counter()
        for l in reversed(range(L)):
            hp.abi.check(lib.licv_inject_bwd_host_saved(sess, l, hb["g"][l].data_ptr(),
                                                        icv_h[l].data_ptr(), dh_h[l].data_ptr(),
                                                        ds_h[l].data_ptr(), n_tok, d, hp.code,
                                                        hp.code, hp.flags), "inject_bwd_host_saved")
        t_c = time.perf_counter()
        hp.abi.check(lib.licv_host_sync(sess), "host_sync")
        t_d = time.perf_counter()
        if want_phases:
            for i, v in enumerate((t_p1 - t_b, t_p2 - t_p1, t_d - t_p2)):
                phase[i] += v
        # autograd of a2, then f2: global-norm clip + AdamW with the two learning rates
        d_vec = alpha.unsqueeze(-1) * ds_h
        d_alpha = (ds_h * vec).sum(1)
        norm = float((d_vec.square().sum() + d_alpha.square().sum()).sqrt())
        coef = min(1.0, 1.0 / (norm + 1e-6))
        step_no[0] += 1
        adamw(vec, d_vec, m_v, v_v, 1e-4, coef, step_no[0])
        adamw(alpha, d_alpha, m_a, v_a, 1e-2, coef, step_no[0])
        if want_seg:
            t_e = time.perf_counter()
            for i, v in enumerate((t_b - t_a, t_c - t_b, t_d - t_c, t_e - t_d)):
                seg[i] += v
        return float(loss_h[2])   # the device->host read of the step's result

    for _ in range(warmup):
        step()
    seg[:] = [0.0, 0.0, 0.0, 0.0]
    phase[:] = [0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    if want_seg:
        print("e2e segments ms/step: prologue %.3f issue %.3f drain %.3f optimizer %.3f of %.3f" %
              tuple([1e3 * v / steps for v in seg] + [1e3 * dt]), file=sys.stderr, flush=True)
    if want_phases:
        fb = L * n_tok * d * es
        lb = (n_tok + B * T4) * V * es
        ms = [1e3 * v / steps for v in phase]
        print("e2e phases (synced apart) ms/step: fwd %.3f (%.1f GB/s each way) loss %.3f (in %.1f GB/s) "
              "bwd %.3f (%.1f GB/s each way) of %.3f" %
              (ms[0], fb / ms[0] / 1e6, ms[1], lb / ms[1] / 1e6, ms[2], fb / ms[2] / 1e6, 1e3 * dt),
              file=sys.stderr, flush=True)
    lib.licv_host_session_destroy(sess)
    # bytes that actually cross the link per step: h (once) and g per layer, the shift twice,
    # student + teacher logits, the two row lists; back: out and dh per layer, d_shift, d(logits)
    h2d = L * (2 * n_tok * d * es + 2 * d * 4) + (n_tok + B * T4) * V * es + n_tok * 12
    d2h = L * (2 * n_tok * d * es + d * 4) + n_tok * V * es + 12
    return dt, h2d, d2h


def e2e_child(args):
    """The e2e leg in its own process (its own CUDA context): a stall there cannot take the main
    measurement down with it."""
    numa = bind_near_gpu(args.e2e_child)      # before anything is pinned
    # the host-side part of the step is a few vector operations on 131 104 floats: a handful of
    # threads, and never more than this rank's share of the host's cores
    share = len(os.sched_getaffinity(0)) // max(1, env_int("LICV_E2E_WORLD", 1))
    torch.set_num_threads(max(1, min(4, share)))
    torch.cuda.set_device(args.e2e_child)
    device = torch.device("cuda", args.e2e_child)
    dtype = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[CFG["dtype"]]
    hp = HotPath(device, dtype, 1)
    sec, h2d, d2h = run_e2e_host(hp, args.steps, args.warmup, 1000 + env_int("RANK", 0))
    print(json.dumps({"e2e_child": True, "sec": sec, "h2d": h2d, "d2h": d2h, "numa": numa}), flush=True)
    return 0


def run_e2e_subprocess(device_index, steps, warmup, timeout_s=240):
    """-> (sec, h2d, d2h, path, numa note) or raises.  Tries the zero-copy host path, then the staged one."""
    import subprocess
    last = None
    for zero_copy in ("1", "0"):
        env = dict(os.environ, LICV_HOST_ZERO_COPY=zero_copy,
                   LICV_E2E_WORLD=os.environ.get("WORLD_SIZE", "1"))
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            if k != "RANK":
                env.pop(k, None)
        cmd = [sys.executable, os.path.abspath(__file__), "--e2e-child", str(device_index),
               "--steps", str(steps), "--warmup", str(warmup)]
        try:
            out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout_s)
            for ln in out.stdout.splitlines():
                if ln.startswith("{") and "e2e_child" in ln:
                    r = json.loads(ln)
                    return (r["sec"], r["h2d"], r["d2h"], ("zero-copy" if zero_copy == "1" else "staged"),
                            r.get("numa", {}))
            last = f"rc={out.returncode}: {out.stderr[-300:]}"
        except subprocess.TimeoutExpired:
            last = f"timed out after {timeout_s}s"
    raise RuntimeError(f"e2e leg failed: {last}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip cpu_baseline / e2e / bandwidth-shape legs (profiling runs)")
    ap.add_argument("--e2e-child", type=int, default=-1, metavar="DEVICE",
                    help="(internal) run only the host-buffer e2e leg on DEVICE, print its JSON")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    if args.e2e_child >= 0:
        return e2e_child(args)

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU path "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=device)
    dtype = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[CFG["dtype"]]
    hp = HotPath(device, dtype, world)
    batch = make_batch(device, 1000 + rank, dtype)
    peak, peak_src = load_peak()

    stream = torch.cuda.Stream()
    graph = None
    checks = {}
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            hp.step(batch)
        stream.synchronize()
        # correctness evidence, outside every timed region: the exchange against NCCL + the same
        # optimizer kernel on copies (N > 1), the step's arithmetic against the float64 oracle
        if world > 1:
            checks.update(dp_check(hp, batch))
        if rank == 0 and not args.no_extras:
            checks.update(oracle_check(hp, batch))
        hp.step(batch)          # every rank back in the same state of the step sequence
        stream.synchronize()
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                hp.step(batch)
            for _ in range(3):
                graph.replay()
            stream.synchronize()

        def barrier():
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()

        def one_step():
            if graph is not None:
                graph.replay()
            else:
                hp.step(batch)

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        clocks = ClockSampler(local)     # NVML set-up BEFORE the barrier: it takes milliseconds
        barrier()
        for _ in range(2):               # untimed: the first exchange after a barrier aligns the ranks
            one_step()
        with clocks:
            e0.record(stream)
            for k in range(args.steps):
                one_step()
                marks[k].record(stream)
            e1.record(stream)
            barrier()
        sec = e0.elapsed_time(e1) * 1e-3
        per_step = sorted(((marks[k - 1] if k else e0).elapsed_time(marks[k]) for k in range(args.steps)))
        step_stats = [per_step[len(per_step) // 2], per_step[min(len(per_step) - 1, int(0.9 * len(per_step)))],
                      per_step[-1]]
        if world > 1:
            t = torch.tensor([sec] + step_stats, device=device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            sec, step_stats = float(t[0]), [float(x) for x in t[1:]]
        ms_per_step = sec / args.steps * 1e3
        value = CFG["batch_per_gpu"] * world / (sec / args.steps)

        # ---- N > 1: what each GPU needs for the step's own work (no exchange, no optimizer) ------
        # The ranks meet once per step, so the job runs at the pace of the slowest GPU: its compute
        # time, next to the N = 1 line's, tells GPU-to-GPU spread apart from the cost of the exchange.
        per_rank_compute_ms, exchange_only_ms = None, None
        if world > 1:
            hp.compute(batch)
            stream.synchronize()
            gc = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gc, stream=stream):
                hp.compute(batch)
            for _ in range(3):
                gc.replay()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for _ in range(20):
                gc.replay()
            c1.record(stream)
            stream.synchronize()
            mine = torch.tensor([c0.elapsed_time(c1) / 20], device=device)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            torch.distributed.all_gather(allr, mine)
            per_rank_compute_ms = [round(float(x), 4) for x in allr]
            # the exchange + optimizer alone, back to back on every rank (same count everywhere)
            go = torch.cuda.CUDAGraph()
            hp.optimize()
            stream.synchronize()
            with torch.cuda.graph(go, stream=stream):
                hp.optimize()
            torch.distributed.barrier()
            for _ in range(3):
                go.replay()
            c0.record(stream)
            for _ in range(20):
                go.replay()
            c1.record(stream)
            stream.synchronize()
            exchange_only_ms = round(c0.elapsed_time(c1) / 20, 4)
            hp.step(batch)                       # parameters of all ranks advance together again
            stream.synchronize()

        # ---- dominant kernel, per launch, CUDA events on the launching stream -------------------
        L, d = CFG["layers"], CFG["d"]
        n_tok = CFG["batch_per_gpu"] * CFG["student_tokens"]
        es = esize(dtype)
        st = stream.cuda_stream
        def bwd_launch(l):
            return lambda: hp.lib.licv_inject_bwd_spread(
                batch["h"][l].data_ptr(), batch["g"][l].data_ptr(), hp.icv[l].data_ptr(),
                hp.dh[l].data_ptr(), hp.rows[l].data_ptr(), hp.n_rows, n_tok, d, hp.code, hp.code,
                hp.flags, st)

        def fwd_launch(l):
            return lambda: hp.lib.licv_inject_fwd(batch["h"][l].data_ptr(), hp.icv[l].data_ptr(),
                                                  hp.out[l].data_ptr(), n_tok, d, hp.code, hp.code,
                                                  hp.flags, st)

        def kd_launch():
            V = CFG["vocab"]
            return hp.lib.licv_kd_loss_fwd_bwd(
                batch["stu"].data_ptr(), hp.dstu.data_ptr(), batch["tea"].data_ptr(),
                hp.kl_tea_row.data_ptr(), hp.ce_label.data_ptr(), hp.counts.data_ptr(), 0, 0,
                CFG["temperature"], CFG["kl_eps"], CFG["hard_loss_weight"], 0, 1.0,
                hp.losses.data_ptr(), hp.ws.data_ptr(), n_tok, V, V, V, hp.code,
                hp.abi.ROUND_TEMPERED, st)

        def graph_time(launches, reps=20):
            """Seconds per launch: the launches captured as one CUDA graph (no host gaps between
            them), CUDA events on the launching stream around `reps` replays."""
            for fn in launches:
                fn()
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for fn in launches:
                    fn()
            for _ in range(3):
                g.replay()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                g.replay()
            b.record(stream)
            stream.synchronize()
            return a.elapsed_time(b) * 1e-3 / (reps * len(launches))

        t_bwd = graph_time([bwd_launch(l) for l in reversed(range(L))])
        t_fwd = graph_time([fwd_launch(l) for l in range(L)])
        t_kd = graph_time([kd_launch])
        bytes_bwd = 3 * es * n_tok * d
        n_kl_rows = CFG["batch_per_gpu"] * CFG["kl_rows_per_sample"]
        n_ce_rows = CFG["batch_per_gpu"] * (CFG["student_tokens"] - 1)
        n_dead = n_tok - n_ce_rows
        bytes_kd = es * CFG["vocab"] * (3 * n_kl_rows + 2 * (n_ce_rows - n_kl_rows) + n_dead)
        roofline = {"bound": "hbm", "kernel": "licv_inject_bwd",
                    "achieved": bytes_bwd / t_bwd / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": bytes_bwd / t_bwd / 1e9 / peak,
                    "frac_of_nominal_8tbs": bytes_bwd / t_bwd / 8e12,
                    "traffic": load_traffic("licv_inject_bwd@256tok"),
                    "traffic_note": "ncu --set full capture of the same launch shape (bf16): DRAM "
                                    "reads = the algorithmic h + g; the dh write-back is still in "
                                    "L2 when the kernel ends",
                    "peak_source": peak_src, "bytes_per_launch": bytes_bwd,
                    "us_per_launch": t_bwd * 1e6, "launches_per_step": L,
                    "share_of_step": L * t_bwd / (ms_per_step * 1e-3),
                    "how": "the step's 32 launches of the kernel captured as one CUDA graph "
                           "(distinct h/g/dh buffers per layer, 200 MB > L2 between reuses), "
                           "CUDA events on the launching stream around 20 replays, right after "
                           "the timed region; 256 tokens per launch = 6 MB: latency-bound, see "
                           "roofline_bw for the same kernel at bandwidth-bound sizes",
                    "other_kernels": [
                        {"kernel": "licv_inject_fwd", "us_per_launch": t_fwd * 1e6,
                         "launches_per_step": L, "bytes_per_launch": 2 * es * n_tok * d,
                         "achieved": 2 * es * n_tok * d / t_fwd / 1e9,
                         "share_of_step": L * t_fwd / (ms_per_step * 1e-3)},
                        {"kernel": "licv_kd_loss_fwd_bwd", "us_per_launch": t_kd * 1e6,
                         "launches_per_step": 1, "bytes_per_launch": bytes_kd,
                         "achieved": bytes_kd / t_kd / 1e9,
                         "share_of_step": t_kd / (ms_per_step * 1e-3)}]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": CFG["dtype"] + " (fp32 accumulate)",
        "data": "synthetic", "config": config_of(world),
        "arm": {"grad_exchange": ("none (1 GPU)" if world == 1 else
                                  "fused p2p exchange (%s form) + AdamW over NVLink peer memory"
                                  % ({"all": "all-to-all", "owner": "owner"}.get(os.environ.get("LICV_DP_ALGO", ""), "owner"))
                                  if hp.peer is not None else "nccl all_reduce" + hp.peer_note),
                "cuda_graph": graph is not None, "gpu": torch.cuda.get_device_name(local)},
        "step_ms": {"median": step_stats[0], "p90": step_stats[1], "max": step_stats[2],
                    "how": "CUDA event after every step of the timed region; max over ranks"},
        "roofline": roofline, "clocks": clocks.summary(),
        "gpu_launches": hp.launches_per_step * args.steps,
    }
    if per_rank_compute_ms:
        line["per_rank_compute_ms"] = {
            "values": per_rank_compute_ms, "slowest": max(per_rank_compute_ms), "fastest": min(per_rank_compute_ms),
            "exchange_and_adamw_alone_ms": exchange_only_ms,
            "how": "the step WITHOUT exchange and optimizer (icv_scale .. icv_grad_finish) as a CUDA graph on "
                   "every rank, 20 replays after the timed region: the ranks meet once per step, so "
                   "ms_per_step tracks the slowest GPU's value plus exchange + AdamW"}
    if checks:
        line["checks"] = checks

    if rank == 0 and not args.no_extras:
        with torch.cuda.stream(stream):
            line["roofline_bw"] = bandwidth_shapes(hp, peak)
        stream.synchronize()
    if not args.no_extras:
        e_steps = max(3, min(args.steps, 20))
        try:
            e_sec, h2d, d2h, e_path, numa = run_e2e_subprocess(local, e_steps, 3)
        except RuntimeError as exc:
            e_sec, h2d, d2h, e_path, numa = float("nan"), 0, 0, str(exc), {}
        e_mine = e_sec
        if world > 1:
            t = torch.tensor([e_sec], device=device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            e_sec = float(t)
            per_rank = [torch.zeros(1, device=device) for _ in range(world)]
            torch.distributed.all_gather(per_rank, torch.tensor([e_mine], device=device))
            per_rank = [float(x) for x in per_rank]
        else:
            per_rank = [e_mine]
        ok = e_sec == e_sec
        line["e2e"] = {"value": CFG["batch_per_gpu"] * world / e_sec if ok else None, "unit": UNIT,
                       "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e_sec * 1e3 if ok else None, "steps": e_steps, "path": e_path,
                       "per_rank_ms": [round(x * 1e3, 3) for x in per_rank],
                       "per_rank_link_gbs_each_way": [round((h2d + d2h) / 2 / x / 1e9, 1) if x == x and x > 0
                                                      else None for x in per_rank],
                       "numa": numa,
                       "how": "the same step through the licv_*_host entry points on one pipelined "
                              "session (own process, bound to the GPU's NUMA node where the "
                              "container allows): hidden states, gradients and logits start in pinned "
                              "host memory and end there, the ICV product, row lists, d_vec / d_alpha "
                              "and clip + AdamW are host-side work inside the timed region; zero-copy "
                              "= the kernels read and write the pinned host buffers over PCIe "
                              "themselves, staged = cudaMemcpyAsync through device scratch"}
    if rank == 0 and world == 1 and not args.no_extras:
        one, cores = cpu_reference_steps(1, 1)
        n = max(2, min(40, int(12.0 / max(one, 1e-3))))
        sec_cpu, cores = cpu_reference_steps(n, 0)
        line["cpu_baseline"] = {"value": CFG["batch_per_gpu"] / sec_cpu, "unit": UNIT,
                                "cores": cores, "kind": "port", "ms_per_step": sec_cpu * 1e3,
                                "sample": f"{n} full steps of the same workload through " + REFERENCE_RECIPE}
        with torch.cuda.stream(stream):
            sec_eager = eager_cuda_steps(device)
        line["eager_cuda_baseline"] = {
            "value": CFG["batch_per_gpu"] / sec_eager, "unit": UNIT, "ms_per_step": sec_eager * 1e3,
            "what": "the same eager chain (oracle/torch_chain.py) in stock PyTorch on this B200, inputs "
                    "resident, CUDA events around 10 steps - SURVEY 8(d)'s same-box comparator",
            "speedup_of_value": value / (CFG["batch_per_gpu"] / sec_eager)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        torch.distributed.barrier()
        # teardown must not be able to hang the launcher: the line is out, everything is
        # synchronised; a watchdog ends the process if NCCL / CUDA-graph destructors stall
        sys.stdout.flush()
        sys.stderr.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        graph = None
        torch.cuda.synchronize()
        torch.distributed.destroy_process_group()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
