"""TEST INFRASTRUCTURE ONLY - CPU oracle for the L-ICV hot path (numpy, float64).

A restatement, in closed form, of the arithmetic the reference (ForJadeForest/LICV-VQA) performs on
its data-parallel hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and only as the checker.
The product (``licv_vqa_b200``) never imports it and has no CPU fallback.

Parity pinning: the reference ships NO tests, golden vectors or fixtures, so this
oracle is pinned the other way the contract allows - against outputs of the reference's own code
executed in the build container: ``oracle/make_golden.py`` imports the unmodified reference from
``/root/reference`` (``oracle/ref_loader.py``) and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those vectors.

Every function cites the reference lines it follows (paths relative to ``/root/reference``).
All maths is float64 on float inputs (bf16/fp16 inputs are passed as their exact float values).
"""
from __future__ import annotations

import numpy as np

IGNORE_INDEX = -100


# --------------------------------------------------------------------------------------------
# a1  GlobalICVEncoder.get_alpha / forward      icv_src/icv_encoder/global_icv_encoder.py:35-43
# --------------------------------------------------------------------------------------------
def encoder_alpha(alpha_raw, use_sigmoid):
    """``get_alpha``: sigmoid(alpha) iff use_sigmoid (global_icv_encoder.py:40-43)."""
    a = np.asarray(alpha_raw, dtype=np.float64)
    return 1.0 / (1.0 + np.exp(-a)) if use_sigmoid else a


def encoder_alpha_bwd(alpha_raw, use_sigmoid, d_alpha_eff):
    a = np.asarray(alpha_raw, dtype=np.float64)
    g = np.asarray(d_alpha_eff, dtype=np.float64)
    if not use_sigmoid:
        return g
    sg = 1.0 / (1.0 + np.exp(-a))
    return g * sg * (1.0 - sg)


# --------------------------------------------------------------------------------------------
# a2  icv = alpha.unsqueeze(-1) * in_context_vector        icv_src/icv_module.py:89-92,
#                                                           inference.py:311
# --------------------------------------------------------------------------------------------
def icv_product(alpha_eff, vec):
    """alpha_eff [1,L], vec [1,L,d] -> icv [1,L,d]."""
    return np.asarray(alpha_eff, np.float64)[..., None] * np.asarray(vec, np.float64)


def icv_product_bwd(alpha_eff, vec, d_icv):
    """-> (d_alpha_eff [1,L], d_vec [1,L,d])."""
    a = np.asarray(alpha_eff, np.float64)
    v = np.asarray(vec, np.float64)
    g = np.asarray(d_icv, np.float64)
    return (g * v).sum(-1), a[..., None] * g


# --------------------------------------------------------------------------------------------
# storage-format rounding, used to restate WHERE the reference's eager chain rounds
# --------------------------------------------------------------------------------------------
def round_to(x, fmt):
    """Round float values to ``fmt`` ("bf16" | "fp16" | "fp32") with round-to-nearest-even and
    return them as float64 (what a torch op producing a tensor of that dtype stores)."""
    x32 = np.asarray(x, np.float64).astype(np.float32)
    if fmt in ("fp32", "float32"):
        return x32.astype(np.float64)
    if fmt in ("fp16", "float16"):
        with np.errstate(over="ignore"):
            return x32.astype(np.float16).astype(np.float64)
    if fmt in ("bf16", "bfloat16"):
        u = np.ascontiguousarray(x32).view(np.uint32).astype(np.uint64)
        nan = np.isnan(x32)
        r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
        out = (r & 0xFFFFFFFF).astype(np.uint32).view(np.float32).astype(np.float64)
        return np.where(nan, np.nan, out)
    raise ValueError(fmt)


# Where the reference's five eager ops round when the hidden states are bf16/fp16
# (icv_intervention.py:66-72; dtype rules probed in SURVEY.md §8a and with CUDA autocast's op
# lists: `norm`, `softmax`, `log`, `sum` run in fp32 under autocast, `+ - * /` do not):
#   RY   the shift is itself low precision (DeepSpeed recipe casts the ICV parameters), so
#        `hidden_states + shift` is stored in the low-precision dtype
#   RNH  `hidden_states.norm()` is stored in the hidden states' dtype (no autocast, e.g.
#        inference.py with a bf16 tower and the fp32 checkpointed ICV)
#   RNY  `shifted_states.norm()` is stored in low precision (RY and no autocast)
#   RT   `shifted_states / norm` is stored in low precision (RY and RNY)
# The final product is low precision iff RT and RNH, else fp32 by type promotion.
RY, RNH, RNY, RT = 1, 2, 4, 8


def chain_flags(h_fmt, s_fmt, autocast=False):
    """Rounding flags + the reference's result dtype for a (hidden dtype, icv dtype) pair."""
    lowp = h_fmt in ("bf16", "fp16")
    if not lowp:
        return 0, "fp32"
    ry = s_fmt == h_fmt
    rnh = not autocast
    rny = ry and not autocast
    rt = ry and rny
    flags = (RY if ry else 0) | (RNH if rnh else 0) | (RNY if rny else 0) | (RT if rt else 0)
    return flags, (h_fmt if (rt and rnh) else "fp32")


def _inject_terms(h, s, flags, lowp):
    h = np.asarray(h, np.float64)
    s = np.asarray(s, np.float64)
    y = h + s
    if flags & RY:
        y = round_to(h + round_to(s, lowp), lowp)
    with np.errstate(divide="ignore", invalid="ignore"):
        n_y = np.sqrt((y * y).sum(-1, keepdims=True))
        n_h = np.sqrt((h * h).sum(-1, keepdims=True))
    if flags & RNY:
        n_y = round_to(n_y, lowp)
    if flags & RNH:
        n_h = round_to(n_h, lowp)
    return h, y, n_h, n_y


# --------------------------------------------------------------------------------------------
# a3  intervention_function                    icv_src/icv_model/icv_intervention.py:61-86
# --------------------------------------------------------------------------------------------
def inject_fwd(h, s, flags=0, lowp=None, out_fmt=None):
    """out = (h+s) / ||h+s||_2 * ||h||_2 over the last dim, no eps (icv_intervention.py:66-72).

    h [..., d], s [d] (the reference's ``icv[:, idx].unsqueeze(1)`` broadcast over B and T).
    ||h+s|| = 0 gives NaN exactly like the reference (0/0).  ``flags``/``lowp`` restate where the
    reference's chain rounds for low-precision hidden states (see ``chain_flags``); ``out_fmt``
    rounds the result to a storage format (None = leave as float64).
    """
    h, y, n_h, n_y = _inject_terms(h, s, flags, lowp)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = y / n_y
        if flags & RT:
            t = round_to(t, lowp)
        out = t * n_h
    return round_to(out, out_fmt) if out_fmt else out


# --------------------------------------------------------------------------------------------
# a4  autograd of a3 (closed form; checked against torch autograd of the
#     reference's own function in tests/test_oracle_golden.py)
# --------------------------------------------------------------------------------------------
def inject_bwd(h, s, g, flags=0, lowp=None):
    """g = dL/dout [..., d] -> (dh [..., d], ds [d] summed over every leading index).

    Exact (float64) derivative of the forward *as evaluated*: the rounded y, ||y||, ||h|| the
    forward used enter the closed form; the roundings themselves have identity derivative
    (that is how autograd treats them).
    """
    g = np.asarray(g, np.float64)
    h, y, n_h, n_y = _inject_terms(h, s, flags, lowp)
    with np.errstate(divide="ignore", invalid="ignore"):
        y_hat = y / n_y
        r = n_h / n_y
        c = (y_hat * g).sum(-1, keepdims=True)
        g_y = r * (g - y_hat * c)
        dh = g_y + c * h / n_h
    ds = g_y.reshape(-1, g_y.shape[-1]).sum(0)
    return dh, ds


# --------------------------------------------------------------------------------------------
# a6  VQAICVModule.get_mask                                 icv_src/icv_module.py:136-148
# --------------------------------------------------------------------------------------------
def get_mask(input_ids, mask_length, pad_token_id):
    """mask[b,t] = (t >= mask_length[b]) and (input_ids[b,t] != pad)."""
    ids = np.asarray(input_ids)
    ml = np.asarray(mask_length).reshape(-1, 1)
    t = np.arange(ids.shape[1])[None, :]
    return (t >= ml) & (ids != pad_token_id)


# --------------------------------------------------------------------------------------------
# a7  row gather by boolean mask                            icv_src/icv_module.py:108-111
# --------------------------------------------------------------------------------------------
def gather_rows(logits, mask):
    """logits [B,T,V], mask [B,T] -> [N,V] in row-major (b,t) order."""
    lg = np.asarray(logits)
    return lg[np.asarray(mask, bool)].reshape(-1, lg.shape[-1])


def pair_rows(stu_mask, tea_mask):
    """The pairing the two gathers imply: the n-th True of the student mask (row-major over the
    flattened [B*Tq]) meets the n-th True of the teacher mask (flattened [B*Tt]).

    -> kl_tea_row int32 [B*Tq]: teacher flat row for each student flat row, -1 where not a KL row.
    Raises like the reference would (shape mismatch in the subtraction, icv_module.py:126-131)
    when the two masks select different numbers of rows.
    """
    sm = np.asarray(stu_mask, bool).reshape(-1)
    tm = np.asarray(tea_mask, bool).reshape(-1)
    s_idx = np.flatnonzero(sm)
    t_idx = np.flatnonzero(tm)
    if s_idx.size != t_idx.size:
        raise ValueError(f"student mask selects {s_idx.size} rows, teacher mask {t_idx.size}")
    out = np.full(sm.shape[0], -1, np.int32)
    out[s_idx] = t_idx.astype(np.int32)
    return out


# --------------------------------------------------------------------------------------------
# a8  VQAICVModule.calculate_kl_divergence                  icv_src/icv_module.py:121-134
# --------------------------------------------------------------------------------------------
def _softmax(z):
    z = z - z.max(-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(-1, keepdims=True)


def _tempered(x, T, logit_fmt):
    """``logits /= temperature`` (icv_module.py:122-123) is an in-place divide: on bf16/fp16
    logits the quotient is stored back in that dtype before the (autocast fp32) softmax."""
    z = x / T
    if logit_fmt in ("bf16", "fp16", "bfloat16", "float16") and T != 1.0:
        z = round_to(z, logit_fmt)
    return z


def kl_divergence(stu, tea, temperature=1.0, kl_eps=1e-6, need_grad=True, logit_fmt=None,
                  want_dtemp=False):
    """T^2 * mean_n sum_v p (log(p+eps) - log(q+eps)),  p=softmax(tea/T), q=softmax(stu/T).

    stu, tea [N,V].  Returns (loss, d_stu [N,V] or None); d_stu[n,j] = (T/N)(q_j W_n - w_j),
    w = p q/(q+eps), W_n = sum_v w_v  (teacher carries no grad: icv_module.py:103-105).
    N = 0 gives NaN (mean of empty), like torch.  ``logit_fmt``: storage dtype of the logits
    (see ``_tempered``); everything after the divide is fp32 under the reference's autocast
    recipes and float64 here.
    """
    stu = np.asarray(stu, np.float64)
    tea = np.asarray(tea, np.float64)
    T = float(temperature)
    n = stu.shape[0]
    q = _softmax(_tempered(stu, T, logit_fmt))
    p = _softmax(_tempered(tea, T, logit_fmt))
    per_row = (p * (np.log(p + kl_eps) - np.log(q + kl_eps))).sum(-1)
    loss = (per_row.mean() if n else np.float64("nan")) * T * T
    if not need_grad:
        return loss, None
    w = p * q / (q + kl_eps)
    W = w.sum(-1, keepdims=True)
    d_stu = (T / max(n, 1)) * (q * W - w)
    if not want_dtemp:
        return loss, d_stu
    # learnable_t (icv_module.py:49-52): BOTH in-place divides (:122-123) and the T**2 factor (:133)
    # depend on T.  With z = x / T:  dKL_n/dT = -(1/T) [ sum_j dKL/dzs_j zs_j + sum_j dKL/dzt_j zt_j ],
    # dKL/dzs = q W - w,  dKL/dzt_j = p_j (a_j - sum_v p_v a_v),  a = log(p+eps) - log(q+eps) + p/(p+eps)
    zs, zt = _tempered(stu, T, logit_fmt), _tempered(tea, T, logit_fmt)
    a = np.log(p + kl_eps) - np.log(q + kl_eps) + p / (p + kl_eps)
    g_t = p * (a - (p * a).sum(-1, keepdims=True))
    dkl = -(((q * W - w) * zs).sum(-1) + (g_t * zt).sum(-1)) / T
    d_temp = 2.0 * T * per_row.mean() + T * T * dkl.mean() if n else np.float64("nan")
    return loss, d_stu, d_temp


# --------------------------------------------------------------------------------------------
# a9  HF-internal shifted CE consumed at icv_src/icv_module.py:94-98,115-117
#     (arithmetic lives in transformers, absent from /root/reference: SURVEY.md §8c)
# --------------------------------------------------------------------------------------------
def ce_labels(input_ids, attention_mask=None, variant="idefics", image_token_id=None,
              pad_token_id=None):
    """Per-row next-token label for ``labels = input_ids`` (icv_module.py:94-95); -100 = ignore.

    Row (b,t) predicts token t+1; the last position has no label.  Variants:

    * ``"idefics"``  - transformers 4.38.2 ``IdeficsForVisionText2Text`` (the reference's pin,
      requirements.txt:174): rows where ``attention_mask[b,t+1] != 0`` count, plain mean.
    * ``"idefics2"`` - transformers >=4.40 ``Idefics2ForConditionalGeneration``: same masking and
      ``CrossEntropyLoss(ignore_index=image_token_id)``.
    * ``"causal_lm"`` - transformers 5.x ``ForCausalLMLoss`` (what the installed 5.5.0 runs):
      every shifted position counts (pads too, labels are the ids), ignore_index -100.
    """
    ids = np.asarray(input_ids)
    B, T = ids.shape
    lab = np.full((B, T), IGNORE_INDEX, np.int64)
    lab[:, :-1] = ids[:, 1:]
    if variant in ("idefics", "idefics2"):
        if attention_mask is not None:
            am = np.asarray(attention_mask)
            keep = np.zeros((B, T), bool)
            keep[:, :-1] = am[:, 1:] != 0
            lab[~keep] = IGNORE_INDEX
        if variant == "idefics2" and image_token_id is not None:
            lab[lab == image_token_id] = IGNORE_INDEX
    elif variant != "causal_lm":
        raise ValueError(f"unknown CE variant {variant!r}")
    return lab


def cross_entropy_rows(logits, labels, need_grad=True):
    """mean over rows with label != -100 of (logsumexp(x) - x[label]).

    logits [R,V], labels [R].  Returns (loss, d_logits [R,V] or None, M).  M = 0 -> NaN loss
    (torch's mean-reduction CE over an empty selection).
    """
    x = np.asarray(logits, np.float64)
    lab = np.asarray(labels).reshape(-1)
    act = lab != IGNORE_INDEX
    M = int(act.sum())
    m = x.max(-1, keepdims=True)
    lse = (m + np.log(np.exp(x - m).sum(-1, keepdims=True))).reshape(-1)
    safe = np.where(act, lab, 0)
    nll = lse - x[np.arange(x.shape[0]), safe]
    loss = (nll[act].sum() / M) if M else np.float64("nan")
    if not need_grad:
        return loss, None, M
    sm = _softmax(x)
    sm[np.arange(x.shape[0]), safe] -= 1.0
    d = sm * (act[:, None] / max(M, 1))
    return loss, d, M


# --------------------------------------------------------------------------------------------
# a7+a8+a9+a10 fused, in the shape the C-ABI entry ``licv_kd_loss_fwd_bwd`` takes it
#     loss combine: icv_src/icv_module.py:100-101,107-119
# --------------------------------------------------------------------------------------------
def kd_loss_rows(stu, tea, kl_tea_row, ce_label, temperature=1.0, kl_eps=1e-6,
                 hard_loss_weight=0.0, only_hard_loss=False, logit_fmt=None):
    """stu [R,V] every student row; tea [Rt,V]; kl_tea_row int [R] (-1 = not a KL row);
    ce_label int [R] (-100 = not a CE row).

    Returns dict(kl, ce, loss, d_stu [R,V], N, M):
      loss = kl + hard_loss_weight * ce   (icv_module.py:107-118)
      only_hard_loss -> loss = ce         (icv_module.py:100-101)
    CE is on the un-tempered logits (HF computes it before the module divides its gathered copies
    by T, icv_module.py:122-123); KL on logits / T.  A zero ``hard_loss_weight`` disables the CE
    term entirely (icv_module.py:94,115 test its truthiness), reported as ce = 0.
    """
    stu = np.asarray(stu, np.float64)
    tea = np.asarray(tea, np.float64)
    ktr = np.asarray(kl_tea_row).reshape(-1)
    R = stu.shape[0]
    d = np.zeros_like(stu)
    kl = 0.0
    N = 0
    if not only_hard_loss:
        rows = np.flatnonzero(ktr >= 0)
        N = rows.size
        kl, d_kl = kl_divergence(stu[rows], tea[ktr[rows]], temperature, kl_eps,
                                 logit_fmt=logit_fmt)
        d[rows] += d_kl
    ce = 0.0
    M = 0
    use_ce = bool(hard_loss_weight) or only_hard_loss
    if use_ce:
        ce, d_ce, M = cross_entropy_rows(stu, ce_label)
        d += (1.0 if only_hard_loss else hard_loss_weight) * d_ce
    if only_hard_loss:
        loss = ce
    else:
        loss = kl + (hard_loss_weight * ce if use_ce else 0.0)
    return dict(kl=kl, ce=ce, loss=loss, d_stu=d, N=N, M=M)


# --------------------------------------------------------------------------------------------
# "next" row f2: AdamW + global-norm clip + cosine warm-up   icv_src/icv_module.py:171-209,
#     config/trainer/*.yaml gradient_clip_val, transformers.get_cosine_schedule_with_warmup
# --------------------------------------------------------------------------------------------
def cosine_warmup_factor(step, warm_steps, total_steps, num_cycles=0.5):
    """LR multiplier of ``get_cosine_schedule_with_warmup`` at optimizer step ``step``."""
    if step < warm_steps:
        return float(step) / float(max(1, warm_steps))
    prog = float(step - warm_steps) / float(max(1, total_steps - warm_steps))
    return max(0.0, 0.5 * (1.0 + np.cos(np.pi * num_cycles * 2.0 * prog)))


def clip_coef(grads, max_norm=1.0, eps=1e-6):
    """``torch.nn.utils.clip_grad_norm_`` coefficient over a list of arrays (L2, global)."""
    tot = np.sqrt(sum(float((np.asarray(g, np.float64) ** 2).sum()) for g in grads))
    return min(1.0, max_norm / (tot + eps)), tot


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-3):
    """One ``torch.optim.AdamW`` update (decoupled decay), ``step`` counted from 1."""
    p = np.asarray(p, np.float64) * (1.0 - lr * weight_decay)
    m = beta1 * np.asarray(m, np.float64) + (1 - beta1) * g
    v = beta2 * np.asarray(v, np.float64) + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v
