"""Custom autograd ops over the C ABI (liblicv_b200.so): the injection and the distillation loss.

Each op launches hand-written sm_100a kernels on torch's current stream with raw device pointers
(PyTorch owns every buffer; it is plumbing, not the product).  CUDA tensors only: there is no CPU
path, no Triton, no multi-backend dispatch - a CPU tensor is a hard error.
"""
from __future__ import annotations

import torch

from . import _abi

_DT = {torch.float32: _abi.F32, torch.bfloat16: _abi.BF16, torch.float16: _abi.F16}
_LOWP = (torch.bfloat16, torch.float16)
IGNORE_INDEX = -100


def _code(dtype: torch.dtype) -> int:
    try:
        return _DT[dtype]
    except KeyError:
        raise TypeError(f"licv_vqa_b200 supports float32 / bfloat16 / float16, got {dtype}") from None


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "licv_vqa_b200 runs on B200 GPUs only (no CPU path, no fallback): got a "
                f"{t.device.type} tensor")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------
# where the reference's eager chain rounds (see include/licv_b200.h LICV_ROUND_*)
# ---------------------------------------------------------------------------------------------
def reference_rounding(h_dtype: torch.dtype, icv_dtype: torch.dtype, autocast: bool | None = None):
    """(round_flags, result dtype) the reference's ``intervention_function``
    (icv_intervention.py:66-72) produces for hidden states of ``h_dtype`` and an ``icv`` tensor of
    ``icv_dtype`` under torch type promotion, with CUDA autocast on or off (``norm`` runs in fp32
    under autocast; ``+ * /`` follow their operands)."""
    if h_dtype not in _LOWP:
        return 0, torch.float32
    if autocast is None:
        autocast = torch.is_autocast_enabled("cuda")
    ry = icv_dtype == h_dtype
    rnh = not autocast
    rny = ry and not autocast
    rt = ry and rny
    flags = ((_abi.ROUND_Y if ry else 0) | (_abi.ROUND_NH if rnh else 0)
             | (_abi.ROUND_NY if rny else 0) | (_abi.ROUND_T if rt else 0))
    return flags, (h_dtype if (rt and rnh) else torch.float32)


# ---------------------------------------------------------------------------------------------
# a1 + a2: icv = get_alpha().unsqueeze(-1) * in_context_vector
# ---------------------------------------------------------------------------------------------
class _ICVScale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha_raw, vec, use_sigmoid):
        _need_cuda(alpha_raw, vec)
        if alpha_raw.dtype != torch.float32 or vec.dtype != torch.float32:
            raise TypeError("icv_scale: the ICV parameters are fp32")
        _, L, d = vec.shape
        a = alpha_raw.contiguous()
        v = vec.contiguous()
        icv = torch.empty_like(v)
        _abi.check(_abi.load().licv_icv_scale(a.data_ptr(), v.data_ptr(), icv.data_ptr(), L, d,
                                              int(use_sigmoid), _stream()), "licv_icv_scale")
        ctx.save_for_backward(a, v)
        ctx.use_sigmoid = bool(use_sigmoid)
        return icv

    @staticmethod
    def backward(ctx, d_icv):
        a, v = ctx.saved_tensors
        _, L, d = v.shape
        g = d_icv.contiguous().float()
        d_vec = torch.empty_like(v)
        d_alpha = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        _abi.check(_abi.load().licv_icv_scale_bwd(a.data_ptr(), v.data_ptr(), g.data_ptr(),
                                                  d_vec.data_ptr(), _ptr(d_alpha), L, d,
                                                  int(ctx.use_sigmoid), _stream()),
                   "licv_icv_scale_bwd")
        return d_alpha, (d_vec if ctx.needs_input_grad[1] else None), None


def icv_scale(alpha_raw: torch.Tensor, vec: torch.Tensor, use_sigmoid: bool) -> torch.Tensor:
    """``sigmoid?(alpha)[..., None] * vec``: alpha_raw [1,L], vec [1,L,d] -> icv [1,L,d] (fp32).

    One kernel each way instead of sigmoid + unsqueeze + mul and their autograd
    (global_icv_encoder.py:40-43, icv_module.py:89-92, inference.py:311)."""
    return _ICVScale.apply(alpha_raw, vec, use_sigmoid)


# ---------------------------------------------------------------------------------------------
# a3 + a4: residual-stream injection
# ---------------------------------------------------------------------------------------------
def inject_forward(h: torch.Tensor, shift: torch.Tensor, out_dtype: torch.dtype | None = None,
                   round_flags: int = 0) -> torch.Tensor:
    """out = (h + shift)/||h + shift|| * ||h|| per token (no autograd).  h [..., d], shift [d] fp32."""
    _need_cuda(h, shift)
    d = h.shape[-1]
    if shift.dtype != torch.float32 or shift.numel() != d:
        raise ValueError("shift must be an fp32 vector of the hidden size")
    hc = h.contiguous()
    sc = shift.contiguous()
    out_dtype = out_dtype or h.dtype
    out = torch.empty(hc.shape, dtype=out_dtype, device=h.device)
    n_tok = hc.numel() // d if d else 0
    _abi.check(_abi.load().licv_inject_fwd(hc.data_ptr(), sc.data_ptr(), out.data_ptr(), n_tok, d,
                                           _code(h.dtype), _code(out_dtype), round_flags,
                                           _stream()), "licv_inject_fwd")
    return out


def inject_backward(h, g, shift, d_shift, want_dh=True, round_flags=0):
    """dh (or None) for g = dL/dout; accumulates sum_tokens g_y into the fp32 vector d_shift."""
    _need_cuda(h, g, shift, d_shift)
    d = h.shape[-1]
    hc = h.contiguous()
    gc = g.contiguous()
    dh = torch.empty_like(hc) if want_dh else None
    n_tok = hc.numel() // d if d else 0
    _abi.check(_abi.load().licv_inject_bwd(hc.data_ptr(), gc.data_ptr(), shift.data_ptr(),
                                           _ptr(dh), d_shift.data_ptr(), n_tok, d, _code(h.dtype),
                                           _code(g.dtype), round_flags, _stream()),
               "licv_inject_bwd")
    return dh


class _Inject(torch.autograd.Function):
    """h, shift -> out.  Saves h only (the reference's autograd keeps several fp32 [B,T,d]
    intermediates per layer plus baukit's clone)."""

    @staticmethod
    def forward(ctx, h, shift, out_dtype, round_flags, sink_row):
        out = inject_forward(h, shift, out_dtype, round_flags)
        ctx.save_for_backward(h, shift)
        ctx.round_flags = round_flags
        ctx.sink_row = sink_row
        return out

    @staticmethod
    def backward(ctx, g):
        h, shift = ctx.saved_tensors
        row = ctx.sink_row
        if row is None:
            row = torch.zeros_like(shift)
        else:
            ctx.sink_row = None  # a second backward through this node must not double-count
        dh = inject_backward(h, g, shift.contiguous(), row, ctx.needs_input_grad[0],
                             ctx.round_flags)
        return dh, (row if ctx.needs_input_grad[1] else None), None, None, None


def inject(h, shift, out_dtype=None, round_flags=0, sink_row=None):
    """Differentiable injection.  ``sink_row``: optional zero-initialised fp32 [d] buffer that
    receives d_shift (lets a whole model's per-layer gradients land in one [L,d] buffer)."""
    if not (torch.is_grad_enabled() and (h.requires_grad or shift.requires_grad)):
        return inject_forward(h, shift, out_dtype, round_flags)
    return _Inject.apply(h, shift, out_dtype, round_flags, sink_row)


class GradStore:
    """Where the per-layer injection backwards of a backward pass leave d_shift.

    Every layer owns ``R`` fp32 replicas of its [d] gradient; the CTAs of its backward launch add
    into replica ``cta mod R`` (``licv_inject_bwd_spread``: <= 16 CTAs per atomic address instead
    of all of them on one vector), and ``collect`` adds up the replicas of all layers in one
    launch that also zero-fills them for the next pass.  A layer that runs its backward twice in
    a pass (or with other shapes) simply adds again."""

    MAX_ROWS = 16

    def __init__(self, n_layers: int, d: int, device):
        self.L, self.d, self.device = n_layers, d, device
        self.rows = None             # [L, R, d] replicas, zero between passes
        self.r = 0
        self.deposited = [False] * n_layers
        self.expected = set()        # layers whose forward ran with a gradient wanted

    def deposit(self, layer, h, g, shift, want_dh, round_flags):
        lib = _abi.load()
        d = self.d
        hc, gc = h.contiguous(), g.contiguous()
        n_tok = hc.numel() // d
        dh = torch.empty_like(hc) if want_dh else None
        if self.rows is None:
            r = lib.licv_inject_bwd_rows(n_tok, d, _code(h.dtype), _code(g.dtype))
            if r < 0:
                _abi.check(r, "licv_inject_bwd_rows")
            self.r = min(int(r), self.MAX_ROWS)
            self.rows = torch.zeros(self.L, self.r, d, dtype=torch.float32, device=self.device)
        _abi.check(lib.licv_inject_bwd_spread(
            hc.data_ptr(), gc.data_ptr(), shift.data_ptr(), _ptr(dh), self.rows[layer].data_ptr(),
            self.r, n_tok, d, _code(h.dtype), _code(g.dtype), round_flags, _stream()),
            "licv_inject_bwd_spread")
        self.deposited[layer] = True
        return dh

    def collect(self) -> torch.Tensor:
        """-> d_shift of every layer, [L, d] fp32; the store is ready for the next pass."""
        missing = [l for l in self.expected if not self.deposited[l]]
        if missing:
            raise RuntimeError(
                f"ICV gradient collected before layers {missing} ran their backward: the hooked "
                "layers are expected to run backward in the reverse of their forward order")
        out = torch.empty(self.L, self.d, dtype=torch.float32, device=self.device)
        if self.rows is None:
            return out.zero_()
        _abi.check(_abi.load().licv_reduce_rows(self.rows.data_ptr(), out.data_ptr(), self.L, self.r,
                                                self.r * self.d, self.d, 0, 1, _stream()),
                   "licv_reduce_rows")
        self.deposited = [False] * self.L
        self.expected = set()
        return out


def icv_grad_finish(rows, alpha_raw, vec, d_vec, d_alpha=None, d_icv=None, norm_partials=None,
                    grad_prescale=1.0, use_sigmoid=False, accumulate=False, clear=True):
    """The tail of a backward pass in one launch: ``rows`` [L, R, d] (the replicas the spread
    backward launches left) -> d_icv (optional), d_vec (+)= alpha_eff * d_icv, d_alpha (+)=
    (d_icv . vec) * dsigmoid, per-layer squared norms of ``grad_prescale`` * the stored gradient
    (for ``adamw_step(norm_partials=...)``).  The autograd of icv_module.py:89-92 with the
    replica sum in front and the optimizer's norm behind."""
    _need_cuda(rows, alpha_raw, vec, d_vec)
    L, R, d = rows.shape
    for t in (rows, vec, d_vec):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError("icv_grad_finish: contiguous fp32 buffers")
    _abi.check(_abi.load().licv_icv_grad_finish(
        rows.data_ptr(), R, R * d, alpha_raw.data_ptr(), vec.data_ptr(), _ptr(d_icv), d_vec.data_ptr(),
        _ptr(d_alpha), _ptr(norm_partials), float(grad_prescale), L, d, int(use_sigmoid),
        int(accumulate), int(clear), _stream()), "licv_icv_grad_finish")


class _Anchor(torch.autograd.Function):
    """icv -> a 1-element token handed to the FIRST hooked layer's injection.  That layer runs its
    backward last, so the token's gradient arrives when every layer has deposited its d_shift:
    this node then collects them into d(icv).  The per-layer shifts themselves are detached - a
    nested backward (reentrant activation checkpointing runs one per layer) only deposits, and
    the graph behind ``icv`` (the encoder's alpha * v) is walked once."""

    @staticmethod
    def forward(ctx, icv32, store):
        ctx.store = store
        ctx.shape = icv32.shape
        return torch.zeros(1, dtype=torch.float32, device=icv32.device)

    @staticmethod
    def backward(ctx, _g):
        return ctx.store.collect().view(ctx.shape), None


class _InjectStored(torch.autograd.Function):
    """The hook's injection: h -> out, d_shift deposited in a GradStore (see _Anchor)."""

    @staticmethod
    def forward(ctx, h, shift, out_dtype, round_flags, store, layer, token):
        out = inject_forward(h, shift, out_dtype, round_flags)
        ctx.save_for_backward(h, shift)
        ctx.round_flags, ctx.store, ctx.layer = round_flags, store, layer
        ctx.has_token = token is not None
        store.expected.add(layer)
        return out

    @staticmethod
    def backward(ctx, g):
        h, shift = ctx.saved_tensors
        dh = ctx.store.deposit(ctx.layer, h, g, shift, ctx.needs_input_grad[0], ctx.round_flags)
        token_grad = torch.zeros(1, dtype=torch.float32, device=g.device) if ctx.has_token else None
        return dh, None, None, None, None, None, token_grad


def inject_stored(h, shift, out_dtype, round_flags, store, layer, token):
    return _InjectStored.apply(h, shift, out_dtype, round_flags, store, layer, token)


def fan_out_shifts(icv32: torch.Tensor):
    """-> (tuple of L detached shift vectors [d], GradStore)."""
    L, d = icv32.shape[-2], icv32.shape[-1]
    flat = icv32.detach().reshape(L, d)
    return tuple(flat[l] for l in range(L)), GradStore(L, d, icv32.device)


def anchor_token(icv32: torch.Tensor, store: GradStore) -> torch.Tensor:
    return _Anchor.apply(icv32, store)


# ---------------------------------------------------------------------------------------------
# a6: get_mask;  a6 + a7 + a9 label prep: row selection
# ---------------------------------------------------------------------------------------------
def get_mask(input_ids: torch.Tensor, mask_length: torch.Tensor, pad_token_id: int) -> torch.Tensor:
    """mask[b,t] = (t >= mask_length[b]) & (input_ids[b,t] != pad)  (icv_module.py:136-148)."""
    _need_cuda(input_ids, mask_length)
    B, T = input_ids.shape
    ids = input_ids.contiguous().long()
    ml = mask_length.contiguous().long()
    mask = torch.empty(B, T, dtype=torch.bool, device=ids.device)
    _abi.check(_abi.load().licv_get_mask(ids.data_ptr(), ml.data_ptr(), int(pad_token_id), B, T,
                                         mask.data_ptr(), _stream()), "licv_get_mask")
    return mask


CE_VARIANTS = {"idefics": 0, "idefics2": 1, "causal_lm": 2}


def kd_prepare_rows(stu_ids, stu_mask_length, tea_ids, tea_mask_length, pad_token_id,
                    stu_attention_mask=None, ce_variant="idefics", image_token_id=-1,
                    want_ce=True, compact_teacher=False):
    """Row pairing and next-token labels without gathers or host syncs.

    -> (kl_tea_row int32 [B*Tq], ce_label int64 [B*Tq] or None, counts int32 [4] = N, M, N_tea, 0)
    With ``compact_teacher`` a fourth value ``tea_sel`` int32 [B*Tq] is returned (flat teacher row
    of the n-th pair, 0 beyond the last) and ``kl_tea_row`` indexes the compact teacher logits
    ``lm_head(hidden.view(-1, d)[tea_sel])`` instead of the full [B*Tt, V] ones.
    """
    _need_cuda(stu_ids, stu_mask_length, tea_ids, tea_mask_length, stu_attention_mask)
    B, Tq = stu_ids.shape
    Bt, Tt = tea_ids.shape
    if B != Bt:
        raise ValueError("student and teacher batches differ")
    dev = stu_ids.device
    s_ids = stu_ids.contiguous().long()
    t_ids = tea_ids.contiguous().long()
    s_len = stu_mask_length.contiguous().long()
    t_len = tea_mask_length.contiguous().long()
    s_att = None if stu_attention_mask is None else stu_attention_mask.contiguous().long()
    kl_tea_row = torch.empty(B * Tq, dtype=torch.int32, device=dev)
    ce_label = torch.empty(B * Tq, dtype=torch.int64, device=dev) if want_ce else None
    counts = torch.empty(4, dtype=torch.int32, device=dev)
    tea_sel = torch.empty(B * Tq, dtype=torch.int32, device=dev) if compact_teacher else None
    _abi.check(_abi.load().licv_kd_select_rows(
        s_ids.data_ptr(), s_len.data_ptr(), _ptr(s_att), t_ids.data_ptr(), t_len.data_ptr(),
        int(pad_token_id), int(image_token_id), CE_VARIANTS[ce_variant], B, Tq, Tt,
        kl_tea_row.data_ptr(), _ptr(ce_label), counts.data_ptr(), _ptr(tea_sel), _stream()),
        "licv_kd_select_rows")
    if compact_teacher:
        return kl_tea_row, ce_label, counts, tea_sel
    return kl_tea_row, ce_label, counts


# ---------------------------------------------------------------------------------------------
# a7 + a8 + a9 + a10: distillation loss
# ---------------------------------------------------------------------------------------------
_WS_CACHE: dict = {}
_OPT_WS_CACHE: dict = {}
_WS_CACHE_MAX = 32      # (device, stream) pairs kept; the oldest entry goes first


def _cached(cache, key, need, device):
    """A zero-initialised scratch tensor per (device, stream).  It is allocated while that stream
    is current, so dropping it later is ordered after the stream's kernels; the calling thread
    owns its entry - two threads on two streams never share one."""
    ws = cache.get(key)
    if ws is None or ws.numel() < need:
        if ws is None and len(cache) >= _WS_CACHE_MAX:
            cache.pop(next(iter(cache)))
        ws = torch.zeros(max(need, 64), dtype=torch.uint8, device=device)
        cache[key] = ws
    return ws


def _workspace(device, n_rows, dtemp=False):
    """Loss-kernel workspace per (device, stream), grown on demand; its 16-byte header stays zero
    between calls (the kernel restores it)."""
    lib = _abi.load()
    need = (lib.licv_kd_loss_dtemp_workspace_bytes if dtemp else lib.licv_kd_loss_workspace_bytes)(int(n_rows))
    return _cached(_WS_CACHE, (device.index, _stream()), max(need, 4096), device)


def _optimizer_workspace(device):
    """The optimizer's own 16 bytes (sum of squares + ticket, zero between calls)."""
    return _cached(_OPT_WS_CACHE, (device.index, _stream()), 16, device)


def kd_loss_raw(stu, tea, kl_tea_row=None, ce_label=None, counts=None, n_kl=0, n_ce=0,
                temperature=1.0, kl_eps=1e-6, hard_loss_weight=0.0, only_hard_loss=False,
                grad_scale=1.0, in_place=True, want_grad=True, round_flags=_abi.ROUND_TEMPERED,
                want_dtemp=False):
    """One launch: losses [3] = (kl, ce, total) and d total / d stu.  ``want_dtemp``: losses [4],
    the last entry d total / d temperature (learnable_t, icv_module.py:49-52).

    stu [R,V] (last dim contiguous, rows strided), tea [Rt,V].  With ``in_place`` the gradient
    overwrites ``stu``'s storage (the student logits are dead after the loss; the reference
    already mutates its gathered copies in place, icv_module.py:122-123)."""
    _need_cuda(stu, tea, kl_tea_row, ce_label, counts)
    if stu.dim() != 2 or stu.stride(1) != 1:
        raise ValueError("stu must be [R,V] with a contiguous last dimension")
    R, V = stu.shape
    if tea is not None:
        if tea.dim() != 2 or tea.stride(1) != 1 or tea.shape[1] != V or tea.dtype != stu.dtype:
            raise ValueError("tea must be [Rt,V] of the student's dtype with a contiguous last dim")
    losses = torch.empty(4 if want_dtemp else 3, dtype=torch.float32, device=stu.device)
    dstu = None
    if want_grad:
        dstu = stu if in_place else torch.empty_strided(stu.shape, stu.stride(), dtype=stu.dtype,
                                                        device=stu.device)
    ws = _workspace(stu.device, R, want_dtemp)
    entry = _abi.load().licv_kd_loss_fwd_bwd_dtemp if want_dtemp else _abi.load().licv_kd_loss_fwd_bwd
    _abi.check(entry(
        stu.data_ptr(), _ptr(dstu), _ptr(tea), _ptr(kl_tea_row), _ptr(ce_label), _ptr(counts),
        int(n_kl), int(n_ce), float(temperature), float(kl_eps), float(hard_loss_weight),
        int(bool(only_hard_loss)), float(grad_scale), losses.data_ptr(), ws.data_ptr(), R, V,
        stu.stride(0) if R > 0 else V, (tea.stride(0) if tea is not None and tea.shape[0] > 0 else V),
        _code(stu.dtype), int(round_flags), _stream()), "licv_kd_loss_fwd_bwd")
    return losses, dstu


class _KDLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stu, tea, kl_tea_row, ce_label, counts, n_kl, n_ce, temperature, kl_eps,
                hard_loss_weight, only_hard_loss, in_place, round_flags, t_param):
        need = stu.requires_grad
        src = stu.detach()
        want_dtemp = t_param is not None and t_param.requires_grad
        losses, dstu = kd_loss_raw(src, tea, kl_tea_row, ce_label, counts, n_kl, n_ce, temperature,
                                   kl_eps, hard_loss_weight, only_hard_loss, 1.0, in_place, need,
                                   round_flags, want_dtemp)
        ctx.dstu = dstu
        ctx.dtemp = losses[3].clone() if want_dtemp else None
        if in_place and need:
            # stu's storage now holds the gradient: bump its version so autograd refuses, loudly,
            # any other backward that saved the logits themselves
            torch.autograd.graph.increment_version(stu)
        kl, ce = losses[0].clone(), losses[1].clone()
        ctx.mark_non_differentiable(kl, ce)
        return losses[2].clone(), kl, ce

    @staticmethod
    def backward(ctx, g_total, _g_kl, _g_ce):
        dstu = ctx.dstu
        ctx.dstu = None
        d_t = None
        if ctx.dtemp is not None:
            d_t = (ctx.dtemp * g_total.detach().float()).reshape(())
        if dstu is None:
            return (None,) * 13 + (d_t,)
        # upstream gradient of the scalar loss: applied in place, and skipped on the device
        # (no traffic, no host sync) when it is exactly 1
        g = g_total.detach().to(torch.float32).contiguous()
        if dstu.is_contiguous():
            _abi.check(_abi.load().licv_scale_inplace(dstu.data_ptr(), dstu.numel(), g.data_ptr(),
                                                      _code(dstu.dtype), _stream()),
                       "licv_scale_inplace")
        else:  # padded row stride: rare, plain torch
            dstu.mul_(g.to(dstu.dtype))
        return (dstu,) + (None,) * 12 + (d_t,)


def kd_loss(stu, tea, kl_tea_row=None, ce_label=None, counts=None, n_kl=None, n_ce=None,
            temperature=1.0, kl_eps=1e-6, hard_loss_weight=0.0, only_hard_loss=False,
            in_place=True, round_flags=_abi.ROUND_TEMPERED):
    """Differentiable fused loss -> (total, kl, ce) scalars (kl/ce detached, for logging).

    ``stu`` [R,V] is every student row; ``kl_tea_row``/``ce_label`` say which rows take part
    (None/None = the compact form ``calculate_kl_divergence`` takes: row r vs teacher row r, no
    CE).  ``counts`` (device) or ``n_kl``/``n_ce`` (host ints) give the means' denominators.
    With ``in_place`` (default) the gradient is written over ``stu``'s storage, so ``stu`` must
    not be read after this call.  ``temperature`` may be the module's Parameter: when it
    requires grad (learnable_t) its gradient is computed too, at the price of the generic kernel
    and of one host read of its value per call.
    """
    t_param = None
    if isinstance(temperature, torch.Tensor):
        if temperature.requires_grad and torch.is_grad_enabled():
            t_param = temperature
        temperature = float(temperature.detach())
    R = stu.shape[0]
    if counts is None:
        if n_kl is None:
            n_kl = R if kl_tea_row is None else None
        if n_ce is None:
            n_ce = 0 if ce_label is None else None
        if n_kl is None or n_ce is None:
            raise ValueError("give `counts` (device) or both n_kl and n_ce (host) with row lists")
    return _KDLoss.apply(stu, tea, kl_tea_row, ce_label, counts, n_kl or 0, n_ce or 0,
                         float(temperature), float(kl_eps), float(hard_loss_weight),
                         bool(only_hard_loss), bool(in_place), int(round_flags), t_param)


# ---------------------------------------------------------------------------------------------
# f2: fused clip + AdamW on the flat ICV buffer
# ---------------------------------------------------------------------------------------------
def adamw_step(param, grad, exp_avg, exp_avg_sq, n_vec, n_alpha, lr_vec, lr_alpha, step,
               beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-3, grad_prescale=1.0,
               max_grad_norm=1.0, norm_out=None, workspace=None, norm_partials=None):
    """``norm_partials``: fp32 partial sums of (grad * grad_prescale)^2 (what ``icv_grad_finish``
    leaves per layer) - the update then is one launch, without the sum-of-squares kernel."""
    _need_cuda(param, grad, exp_avg, exp_avg_sq)
    if workspace is None:
        workspace = _optimizer_workspace(param.device)
    ws_ptr = workspace.data_ptr()
    if norm_partials is not None:
        _need_cuda(norm_partials)
        if norm_partials.dtype != torch.float32 or not norm_partials.is_contiguous():
            raise TypeError("norm_partials: contiguous fp32")
        _abi.check(_abi.load().licv_adamw_step_partials(
            param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), int(n_vec),
            int(n_alpha), float(lr_vec), float(lr_alpha), float(beta1), float(beta2), float(eps),
            float(weight_decay), int(step), float(grad_prescale), float(max_grad_norm),
            _ptr(norm_out), ws_ptr, norm_partials.data_ptr(), norm_partials.numel(), _stream()),
            "licv_adamw_step_partials")
        return
    _abi.check(_abi.load().licv_adamw_step(
        param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), int(n_vec),
        int(n_alpha), float(lr_vec), float(lr_alpha), float(beta1), float(beta2), float(eps),
        float(weight_decay), int(step), float(grad_prescale), float(max_grad_norm), _ptr(norm_out),
        ws_ptr, _stream()), "licv_adamw_step")
