/*
 * licv_b200.h - C ABI of the B200-native L-ICV hot path (liblicv_b200.so, sm_100a only).
 *
 * Drop-in boundary for ForJadeForest/LICV-VQA's data-parallel hot path.  The reference is pure
 * Python; what crosses its "FFI" for this path are torch tensors handed to the forward hooks of
 * `LearnableICVInterventionLMM` and to `VQAICVModule.calculate_kl_divergence`.  Every entry point
 * below names the reference code it replaces (paths relative to the reference root).  Signatures
 * carry plain pointers and sizes only - no torch types.  The Python host side
 * (licv_vqa_b200/_abi.py) binds them with ctypes; INTEGRATION.md shows the stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the function name ends in `_host`;
 *  - tensors are row-major and contiguous in the last dimension; rows must start on a 16-byte
 *    boundary for the injection kernels (d % 8 == 0 for bf16/fp16, d % 4 == 0 for fp32); the loss
 *    kernels accept any V and any element-aligned row stride (V = 32002 / 32003 are the cases);
 *  - work is enqueued on `stream` and the call returns without synchronising; no allocation
 *    happens inside any non-`_host` call (workspaces are passed in);
 *  - return value: 0 = LICV_OK, otherwise a negative licv_status (argument errors, reported
 *    before anything is launched) or a positive cudaError_t from the launch;
 *  - there is NO CPU path: the library needs a device of compute capability 10.x.
 */
#ifndef LICV_B200_H_
#define LICV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LICV_ABI_VERSION 1

typedef struct CUstream_st* licv_stream_t; /* == cudaStream_t */

typedef enum {
    LICV_OK = 0,
    LICV_ERR_NULL_POINTER = -1,
    LICV_ERR_BAD_DTYPE = -2,
    LICV_ERR_BAD_DIM = -3,        /* d not a multiple of the 16-byte vector, or too large */
    LICV_ERR_MISALIGNED = -4,     /* base pointer not 16-byte aligned */
    LICV_ERR_BAD_ARGUMENT = -5,
    LICV_ERR_WORKSPACE = -6,      /* workspace too small */
    LICV_ERR_NO_DEVICE = -7       /* no sm_100 device / kernel image not loadable */
} licv_status;

typedef enum { LICV_F32 = 0, LICV_BF16 = 1, LICV_F16 = 2 } licv_dtype;

/* Where the reference's eager op chain rounds when hidden states are bf16/fp16
 * (icv_src/icv_model/icv_intervention.py:66-72 under torch type promotion / CUDA autocast).
 * 0 = all arithmetic in fp32, one rounding at the output. */
#define LICV_ROUND_Y 1u   /* `hidden_states + shift` stored in low precision (shift is low precision) */
#define LICV_ROUND_NH 2u  /* `hidden_states.norm()` stored in low precision (no autocast) */
#define LICV_ROUND_NY 4u  /* `shifted_states.norm()` stored in low precision */
#define LICV_ROUND_T 8u   /* `shifted_states / norm` stored in low precision */
/* loss kernel: `logits /= temperature` (icv_src/icv_module.py:122-123) stores the quotient in the
 * logits' own low-precision dtype before the fp32 softmax */
#define LICV_ROUND_TEMPERED 16u

const char* licv_status_string(int status);
int licv_abi_version(void);
/* SM count / compute capability of the current device; LICV_ERR_NO_DEVICE if it is not sm_10x. */
int licv_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * a1+a2  icv = get_alpha().unsqueeze(-1) * in_context_vector
 *   replaces GlobalICVEncoder.get_alpha (icv_src/icv_encoder/global_icv_encoder.py:40-43) and the
 *   product at icv_src/icv_module.py:89-92 / inference.py:311.
 *   alpha_raw [L], vec [L,d] -> icv [L,d] (all fp32).  use_sigmoid != 0 applies sigmoid first.
 * ------------------------------------------------------------------------------------------ */
int licv_icv_scale(const float* alpha_raw, const float* vec, float* icv, int n_layers, int d,
                   int use_sigmoid, licv_stream_t stream);

/* autograd of the above: d_icv [L,d] -> d_vec [L,d] = alpha_eff * d_icv,
 * d_alpha_raw [L] = (d_icv . vec) * (sigmoid' if use_sigmoid).  d_alpha_raw may be NULL
 * (alpha_learnable = False, global_icv_encoder.py:26-29). */
int licv_icv_scale_bwd(const float* alpha_raw, const float* vec, const float* d_icv, float* d_vec,
                       float* d_alpha_raw, int n_layers, int d, int use_sigmoid,
                       licv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a3  residual-stream injection, forward
 *   replaces intervention_function (icv_src/icv_model/icv_intervention.py:61-86), the body of
 *   the baukit TraceDict hook:  out = (h + s) / ||h + s||_2 * ||h||_2   per token, no eps.
 *   h [n_tokens, d] (h_dtype), shift [d] fp32 (= icv[0, layer_to_icv_index[layer]]),
 *   out [n_tokens, d] (out_dtype: h_dtype, or LICV_F32 = the reference's promoted result).
 *   round_flags: LICV_ROUND_* (ignored for fp32 h).  out may not alias h.
 * ------------------------------------------------------------------------------------------ */
int licv_inject_fwd(const void* h, const float* shift, void* out, int64_t n_tokens, int d,
                    int h_dtype, int out_dtype, unsigned round_flags, licv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a4  residual-stream injection, backward (the reference gets this from autograd)
 *   g [n_tokens, d] (g_dtype = the forward's out_dtype) -> dh [n_tokens, d] (h_dtype) and
 *   d_shift [d] fp32, ACCUMULATED (+=) over all tokens with fp32 atomics: zero it first.
 *   dh may alias g when g_dtype == h_dtype.  dh may be NULL (frozen input, only d_shift wanted).
 * ------------------------------------------------------------------------------------------ */
int licv_inject_bwd(const void* h, const void* g, const float* shift, void* dh, float* d_shift,
                    int64_t n_tokens, int d, int h_dtype, int g_dtype, unsigned round_flags,
                    licv_stream_t stream);

/* a4, spread form: the same backward, but d_shift is ADDED (+=, fp32 atomics) into one of n_rows
 * replicas of the [d] vector - CTA b of the launch takes replica b mod n_rows of `rows`
 * [n_rows, d] - and the replicas of all layers are added up by ONE licv_reduce_rows launch at the
 * end of the backward pass.  The L2 atomic units serialise per address: with one [d] vector per
 * layer the 256 CTAs of a training-shape launch queue for 3.5 us (licv_inject_bwd pre-reduces
 * across thread-block clusters instead, at the price of two cluster barriers); with <= 16 CTAs
 * per replica the atomics are free.  `rows` must be zero before the first launch of a pass
 * (licv_reduce_rows(..., clear = 1) leaves it so); n_rows is a power of two <= 64.
 * licv_inject_bwd_rows: the replica count recommended for a launch of this shape (a pure
 * function of its arguments and the device; 1 for rows the TMA kernel does not take);
 * negative = LICV_ERR_*. */
int licv_inject_bwd_rows(int64_t n_tokens, int d, int h_dtype, int g_dtype);
int licv_inject_bwd_spread(const void* h, const void* g, const float* shift, void* dh, float* rows,
                           int n_rows, int64_t n_tokens, int d, int h_dtype, int g_dtype,
                           unsigned round_flags, licv_stream_t stream);
/* out [n_layers, d] (+)= sum_p rows[l * layer_stride + p * d + c], p < n_rows  (fp32; layer_stride
 * in floats, a multiple of 4, >= n_rows * d).  accumulate != 0 adds to `out` (gradient
 * accumulation over micro-batches), 0 overwrites it; clear != 0 zero-fills the rows read. */
int licv_reduce_rows(float* rows, float* out, int n_layers, int n_rows, int64_t layer_stride, int d,
                     int accumulate, int clear, licv_stream_t stream);
/* The tail of a backward pass in ONE launch: licv_reduce_rows, licv_icv_scale_bwd (the autograd of
 * icv = alpha.unsqueeze(-1) * in_context_vector, icv_src/icv_module.py:89-92) and the optimizer's
 * sum of squares.  Per layer l: d_icv[l] = sum of the n_rows replicas (written if d_icv != NULL;
 * the replicas are zero-filled if clear != 0), d_vec[l] (+)= alpha_eff[l] * d_icv[l],
 * d_alpha_raw[l] (+)= (d_icv[l] . vec[l]) * dsigmoid (NULL: alpha not trained), and
 * norm_partials[l] (NULL: not wanted) = the squared L2 norm of grad_prescale * (d_vec[l],
 * d_alpha_raw[l]) as STORED (after accumulation), summed in a fixed order - what
 * licv_adamw_step_partials takes.  accumulate != 0: += into d_vec / d_alpha_raw (gradient
 * accumulation over micro-batches), 0: overwrite. */
int licv_icv_grad_finish(float* rows, int n_rows, int64_t layer_stride, const float* alpha_raw,
                         const float* vec, float* d_icv, float* d_vec, float* d_alpha_raw,
                         float* norm_partials, float grad_prescale, int n_layers, int d,
                         int use_sigmoid, int accumulate, int clear, licv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a6  VQAICVModule.get_mask (icv_src/icv_module.py:136-148)
 *   mask[b,t] = (t >= mask_length[b]) && (input_ids[b,t] != pad_token_id), uint8 0/1 (torch.bool)
 * ------------------------------------------------------------------------------------------ */
int licv_get_mask(const int64_t* input_ids, const int64_t* mask_length, int64_t pad_token_id,
                  int batch, int seq_len, uint8_t* mask, licv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a6+a7 (+ label prep of a9)  row selection without materialising gathered copies
 *   replaces the two get_mask calls and the boolean-mask gathers at icv_src/icv_module.py:84-85,
 *   108-111, and the shifted-label construction of the HF CE consumed at :94-98.
 *   The n-th selected student row (row-major over [B,Tq]) is paired with the n-th selected teacher
 *   row (row-major over [B,Tt]).
 *     kl_tea_row [B*Tq] int32 : flat teacher row for each student row, -1 = not a KL row
 *     ce_label   [B*Tq] int64 : next-token label, -100 = not a CE row (may be NULL)
 *     counts     [4]    int32 : {N = KL rows, M = CE rows, teacher rows selected, 0}
 *   ce_variant: 0 "idefics" (transformers 4.38.2: keep rows with attention_mask[b,t+1] != 0),
 *               1 "idefics2" (same and label != image_token_id), 2 "causal_lm" (transformers 5.x
 *               ForCausalLMLoss: every shifted position).  stu_attention_mask may be NULL
 *               (treated as all ones).  counts[0] != counts[2] is the reference's shape error;
 *               the host side checks it.
 * ------------------------------------------------------------------------------------------ */
int licv_kd_prepare_rows(const int64_t* stu_ids, const int64_t* stu_mask_length,
                         const int64_t* stu_attention_mask, const int64_t* tea_ids,
                         const int64_t* tea_mask_length, int64_t pad_token_id,
                         int64_t image_token_id, int ce_variant, int batch, int stu_len,
                         int tea_len, int32_t* kl_tea_row, int64_t* ce_label, int32_t* counts,
                         licv_stream_t stream);

/* f1 ("next" row)  the same, for teacher logits computed ONLY for the selected rows
 *   (icv_src/icv_module.py:103-111 materialises [B, ~900, V] teacher logits to use ~32 rows):
 *     tea_sel [B*Tq] int32 : flat teacher row of the n-th selected pair, n < N; entries n >= N
 *                            hold 0 (a valid row), so gathering a fixed B*Tq rows of teacher
 *                            hidden states needs no host sync
 *     kl_tea_row           : then holds n (the row of the COMPACT teacher logits [B*Tq, V] made
 *                            by lm_head(hidden.view(-1, d)[tea_sel])) instead of the flat row
 *   tea_sel == NULL is licv_kd_prepare_rows. */
int licv_kd_select_rows(const int64_t* stu_ids, const int64_t* stu_mask_length,
                        const int64_t* stu_attention_mask, const int64_t* tea_ids,
                        const int64_t* tea_mask_length, int64_t pad_token_id,
                        int64_t image_token_id, int ce_variant, int batch, int stu_len,
                        int tea_len, int32_t* kl_tea_row, int64_t* ce_label, int32_t* counts,
                        int32_t* tea_sel, licv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a7+a8+a9+a10  distillation loss, forward + backward in one pass over HBM
 *   replaces VQAICVModule.calculate_kl_divergence (icv_src/icv_module.py:121-134), the HF-internal
 *   shifted cross-entropy consumed at :94-98,115-117 and the combine at :100-101,107-119.
 *
 *   stu  [R, V] student logits, row stride stu_stride elements
 *   dstu [R, V] receives d loss / d stu (same dtype/stride); MAY ALIAS stu (in place); rows that
 *        are neither KL nor CE rows are zero-filled.  NULL = loss only.
 *   tea  [Rt, V] teacher logits, row stride tea_stride (rows addressed through kl_tea_row)
 *   kl_tea_row [R] int32 or NULL (NULL = row r pairs with teacher row r: the compact [N,V] form
 *        `calculate_kl_divergence` itself takes)
 *   ce_label [R] int64 or NULL (no CE term)
 *   counts: device int32 {N, M} as written by licv_kd_prepare_rows, or NULL to use the host
 *        values n_kl / n_ce
 *   loss = T^2/N * sum_rows KL + hard_loss_weight * (1/M) * sum_rows CE;  only_hard_loss != 0
 *        returns the CE alone (icv_module.py:100-101).  CE uses the un-tempered logits.
 *   grad_scale multiplies dstu (1.0, or 1/accumulate_grad_batches).
 *   out_losses [3] fp32 device: {kl_loss, ce_loss, loss}.
 *   workspace: licv_kd_loss_workspace_bytes(R) bytes, 16-byte aligned; its first 16 bytes must be
 *        zero before the first use (the kernel leaves them zero again).
 *   Rows may start on any element boundary (V = 32002 / 32003 do).  The kernel WRITES only the V
 *   elements of each dstu row; it may READ the whole 16-byte-aligned granules that hold the first
 *   and the last element of a stu / tea row (the extra bytes belong to the neighbouring row, the
 *   row padding or, for the first / last row, the same 16-byte granule of the caller's
 *   allocation; they never enter the result).
 * ------------------------------------------------------------------------------------------ */
int64_t licv_kd_loss_workspace_bytes(int64_t n_rows);
/* Which kernel licv_kd_loss_fwd_bwd runs for a vocabulary / dtype / temperature / row count
 * (kl_and_ce != 0: rows carry both a KL and a CE term).  A pure function of its arguments, the
 * device's SM count and the environment
 * switches LICV_KD_STREAM / LICV_KD_TMEM / LICV_KD_NO_CLUSTER; negative = LICV_ERR_*. */
#define LICV_KD_KERNEL_GENERIC 0 /* one CTA per row, sweeps re-read the row from L2 */
#define LICV_KD_KERNEL_CLUSTER 1 /* row pair cached across a thread-block cluster */
#define LICV_KD_KERNEL_TMEM 2    /* row pair cached in shared + tensor memory, one CTA per SM */
#define LICV_KD_KERNEL_STREAM 3  /* TMA-staged rows, gradient sweep fused with the next row's exponentials */
int licv_kd_loss_plan(int vocab, int dtype, float temperature, int kl_and_ce, int64_t n_rows);
/* Test / timing hook: 0 = never the stream kernel, 1 = where it is the faster one (default),
 * 2 = wherever it can run, -1 = back to the environment (LICV_KD_STREAM).  Process-wide. */
void licv_debug_set_kd_stream(int mode);
int licv_kd_loss_fwd_bwd(const void* stu, void* dstu, const void* tea, const int32_t* kl_tea_row,
                         const int64_t* ce_label, const int32_t* counts, int64_t n_kl, int64_t n_ce,
                         float temperature, float kl_eps, float hard_loss_weight,
                         int only_hard_loss, float grad_scale, float* out_losses, void* workspace,
                         int64_t n_rows, int vocab, int64_t stu_stride, int64_t tea_stride,
                         int dtype, unsigned round_flags, licv_stream_t stream);
/* The same with a learnable temperature (learnable_t, icv_src/icv_module.py:49-52): out_losses4
 * [4] = {kl_loss, ce_loss, loss, d loss / d temperature} - both in-place divides by T
 * (icv_module.py:122-123) and the T^2 factor (:133) are differentiated, like the reference's
 * autograd does.  Runs the generic kernel (it re-reads the logits it needs);
 * workspace: licv_kd_loss_dtemp_workspace_bytes(R). */
int64_t licv_kd_loss_dtemp_workspace_bytes(int64_t n_rows);
int licv_kd_loss_fwd_bwd_dtemp(const void* stu, void* dstu, const void* tea, const int32_t* kl_tea_row,
                               const int64_t* ce_label, const int32_t* counts, int64_t n_kl,
                               int64_t n_ce, float temperature, float kl_eps, float hard_loss_weight,
                               int only_hard_loss, float grad_scale, float* out_losses4,
                               void* workspace, int64_t n_rows, int vocab, int64_t stu_stride,
                               int64_t tea_stride, int dtype, unsigned round_flags,
                               licv_stream_t stream);

/* x[i] *= *scale for i < n, skipped entirely (no traffic) when *scale == 1: the upstream
 * gradient of the loss applied to the in-place dstu in the autograd backward. */
int licv_scale_inplace(void* x, int64_t n, const float* scale, int dtype, licv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * f2 ("next" row)  optimizer step on the flat ICV parameter buffer
 *   replaces torch.optim.AdamW / DeepSpeedCPUAdam + gradient_clip_val + the cosine warm-up
 *   schedule (icv_src/icv_module.py:171-209, config/trainer/{ddp,zero2}.yaml) for the L*d + L trainable
 *   floats.  flat layout: [vec (n_vec floats) | alpha (n_alpha floats)].
 *   grad is first multiplied by grad_prescale (1/world_size after the all-reduce sum), then
 *   clipped to global L2 norm max_grad_norm (<= 0 disables), then AdamW with lr_vec / lr_alpha.
 *   step counts from 1.  norm_out [1] fp32 device receives the pre-clip gradient norm (or NULL).
 * ------------------------------------------------------------------------------------------ */
int licv_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                    int64_t n_vec, int64_t n_alpha, float lr_vec, float lr_alpha, float beta1,
                    float beta2, float eps, float weight_decay, int64_t step, float grad_prescale,
                    float max_grad_norm, float* norm_out, void* workspace /* >= 16 B, zeroed */,
                    licv_stream_t stream);

/* The same update with the squared gradient norm handed over as n_partials partial sums of
 * (grad * grad_prescale)^2 - what licv_icv_grad_finish leaves per layer - added in a fixed order:
 * ONE launch instead of two (no sum-of-squares kernel). */
int licv_adamw_step_partials(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                             int64_t n_vec, int64_t n_alpha, float lr_vec, float lr_alpha,
                             float beta1, float beta2, float eps, float weight_decay, int64_t step,
                             float grad_prescale, float max_grad_norm, float* norm_out,
                             void* workspace /* >= 16 B, zeroed */, const float* norm_partials,
                             int n_partials, licv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * e  data-parallel gradient exchange fused with the optimizer step, over NVLink peer memory
 *   replaces Lightning DDP's NCCL all-reduce of icv_encoder.*.grad, the sync_dist collective of
 *   the logged scalars (icv_src/icv_module.py:163) and the optimizer step (config/trainer/
 *   ddp.yaml:5,7; icv_module.py:171-209) for one process per GPU on one NVSwitch node.
 *   Set-up (once): every rank allocates a region (licv_dp_region_alloc returns a 64-byte CUDA IPC
 *   handle), the ranks exchange the handles (any host channel), licv_dp_comm_create maps the
 *   peers.  Per optimizer step: licv_dp_allreduce_adamw - `grad` [n_vec + n_alpha + n_extra,
 *   rounded up to 4 floats] holds the local gradient (+ n_extra logged scalars) on entry and the
 *   SUM over ranks on return (summed in rank order: bit-identical on every rank); the parameters
 *   are updated with the mean gradient (clip at max_grad_norm, AdamW).  Two launches, no host
 *   synchronisation, replayable from a CUDA graph (the step counter lives in device memory).
 *   world == 1 degenerates to licv_adamw_step.  A region holds two slots per possible source
 *   rank (16) of 16-byte {float, step tag, float, step tag} packets (each 8-byte half guarded by
 *   its own tag): about 34 MB for idefics shapes.  Every rank must call once per step.  A peer
 *   that never delivers (~4 s) raises the error flag: that step's optimizer update is SKIPPED on
 *   this rank (parameters, moments and the local gradient stay as they were) and every later
 *   step too, until licv_dp_comm_reset_error.
 * ------------------------------------------------------------------------------------------ */
typedef struct licv_dp_comm licv_dp_comm;
int64_t licv_dp_region_bytes(int64_t n_floats);
int licv_dp_region_alloc(int64_t n_floats, void** region, void* ipc_handle_64);
int licv_dp_comm_create(licv_dp_comm** out, int rank, int world, void* region,
                        const void* all_handles /* world x 64 bytes, rank order */, int64_t n_floats);
int licv_dp_comm_destroy(licv_dp_comm* c);
int licv_dp_comm_error(licv_dp_comm* c); /* 1 if a wait for a peer ever timed out (synchronises) */
int licv_dp_comm_reset_error(licv_dp_comm* c);
int licv_dp_region_free(void* region);    /* a region that never made it into a comm */
int licv_dp_allreduce_adamw(licv_dp_comm* c, float* param, float* grad, float* exp_avg,
                            float* exp_avg_sq, int64_t n_vec, int64_t n_alpha, int64_t n_extra,
                            float lr_vec, float lr_alpha, float beta1, float beta2, float eps,
                            float weight_decay, int64_t step, float max_grad_norm, float* norm_out,
                            void* workspace /* >= 16 B, zeroed */, licv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer entry points (what a host-side plugin of the reference would call): inputs and
 * outputs live in HOST memory.  Pinned / registered buffers are read and written by the kernels
 * directly over the host link (zero-copy: both directions busy inside one kernel, nothing staged
 * in HBM); pageable buffers are staged through device scratch owned by the session with
 * cudaMemcpyAsync.  Calls on one session are pipelined on internal streams; licv_host_sync waits
 * for all of them.  One session is driven by one host thread at a time.
 * ------------------------------------------------------------------------------------------ */
typedef struct licv_host_session licv_host_session;
int licv_host_session_create(licv_host_session** out, int64_t scratch_bytes_per_slot, int n_slots);
int licv_host_session_destroy(licv_host_session* s);
int licv_host_sync(licv_host_session* s);
void* licv_host_alloc_pinned(int64_t bytes);
void licv_host_free_pinned(void* p);

int licv_inject_fwd_host(licv_host_session* s, const void* h, const float* shift, void* out,
                         int64_t n_tokens, int d, int h_dtype, int out_dtype, unsigned round_flags);
int licv_inject_bwd_host(licv_host_session* s, const void* h, const void* g, const float* shift,
                         void* dh, float* d_shift /* [d], overwritten */, int64_t n_tokens, int d,
                         int h_dtype, int g_dtype, unsigned round_flags);
/* Forward that keeps its hidden states on the device for the matching backward (what autograd's
 * "saved for backward" is on a GPU): the backward then moves only g in and dh out.  `key` names
 * the pair (e.g. the layer index); a second save under a live key replaces it; the saved copy
 * is released by the backward that uses it.  licv_inject_bwd_host_saved returns
 * LICV_ERR_BAD_ARGUMENT when no forward was saved under `key`. */
int licv_inject_fwd_host_save(licv_host_session* s, int64_t key, const void* h, const float* shift,
                              void* out, int64_t n_tokens, int d, int h_dtype, int out_dtype,
                              unsigned round_flags);
int licv_inject_bwd_host_saved(licv_host_session* s, int64_t key, const void* g, const float* shift,
                               void* dh, float* d_shift /* [d], overwritten */, int64_t n_tokens,
                               int d, int h_dtype, int g_dtype, unsigned round_flags);
int licv_kd_loss_fwd_bwd_host(licv_host_session* s, const void* stu, void* dstu, const void* tea,
                              const int32_t* kl_tea_row, const int64_t* ce_label, int64_t n_kl,
                              int64_t n_ce, float temperature, float kl_eps, float hard_loss_weight,
                              int only_hard_loss, float grad_scale, float* out_losses /* host [3] */,
                              int64_t n_rows, int64_t n_tea_rows, int vocab, int dtype,
                              unsigned round_flags);

#ifdef __cplusplus
}
#endif
#endif /* LICV_B200_H_ */
