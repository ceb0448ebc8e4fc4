// torch.ops.licv.* - the hot path registered as torch operators from C++ (TORCH_LIBRARY).
//
// north_star / SURVEY 8(b): "a thin C-ABI torch extension registered as custom autograd ops that
// replace the baukit TraceDict hooks" (icv_src/icv_model/icv_intervention.py:88-113).  This file is
// that shim and nothing else: every operator checks its tensors, allocates its outputs with
// PyTorch, and calls ONE entry point of include/licv_b200.h on the current CUDA stream of the
// tensors' device.  No kernel lives here (the .so built from this file links liblicv_b200.so), and
// there is no CPU implementation: the operators are registered for the CUDA and Meta keys only, so
// a CPU tensor fails in the dispatcher.
//
//   licv::inject(h, shift, out_dtype, round_flags) -> out          icv_intervention.py:61-86
//   licv::inject_bwd(h, g, shift, round_flags) -> (dh, d_shift)    its closed-form backward
//   licv::kd_loss(stu, tea, kl_tea_row, ce_label, counts, n_kl, n_ce, temperature, kl_eps,
//                 hard_loss_weight, only_hard_loss, round_flags) -> (total, kl, ce, d total/d stu)
//                                                                   icv_module.py:94-134, functional
//   licv::get_mask(input_ids, mask_length, pad_token_id) -> mask   icv_module.py:136-148
//
// Autograd formulas are C++ too (torch::autograd::Function at the Autograd key, calling the
// operators back through the dispatcher so that a trace sees them), i.e. the hook body runs
// without a Python frame or a ctypes call in either direction.  Threading: forward
// on the caller's thread, backward on autograd's device thread - every call takes the device guard
// and the current stream itself and keeps no state except the loss workspace cache below.
#include <ATen/ATen.h>
#include <ATen/core/dispatch/Dispatcher.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/csrc/autograd/custom_function.h>
#include <torch/library.h>

#include <map>
#include <mutex>
#include <tuple>
#include <utility>

#include "licv_b200.h"

namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

int dtype_code(c10::ScalarType t) {
    switch (t) {
        case c10::ScalarType::Float: return LICV_F32;
        case c10::ScalarType::BFloat16: return LICV_BF16;
        case c10::ScalarType::Half: return LICV_F16;
        default: TORCH_CHECK(false, "licv: unsupported dtype ", t, " (bf16, fp16, fp32 only)");
    }
}

void check_rc(int rc, const char* what) {
    TORCH_CHECK(rc == 0, what, ": ", licv_status_string(rc), " (", rc, ")");
}

licv_stream_t stream_of(const Tensor& t) {
    return reinterpret_cast<licv_stream_t>(c10::cuda::getCurrentCUDAStream(t.device().index()).stream());
}

const void* ptr_or_null(const std::optional<Tensor>& t) { return t.has_value() ? t->const_data_ptr() : nullptr; }

// loss-kernel workspace per (device, stream), grown on demand; its 16-byte header stays zero
// between calls (the kernel restores it)
Tensor kd_workspace(const Tensor& like, int64_t n_rows) {
    static std::mutex mu;
    static std::map<std::pair<int, void*>, Tensor> cache;
    const int64_t need = std::max<int64_t>(licv_kd_loss_workspace_bytes(n_rows), 4096);
    const auto key = std::make_pair((int)like.device().index(), (void*)stream_of(like));
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it == cache.end() || it->second.numel() < need) {
        if (cache.size() > 64) cache.clear();
        Tensor ws = at::zeros({need}, like.options().dtype(at::kByte));
        cache[key] = ws;
        return ws;
    }
    return it->second;
}

// ---------------------------------------------------------------------------------------------
// CUDA implementations
// ---------------------------------------------------------------------------------------------
Tensor inject_cuda(const Tensor& h, const Tensor& shift, c10::ScalarType out_dtype, int64_t round_flags) {
    TORCH_CHECK(h.is_cuda() && shift.is_cuda(), "licv::inject: CUDA tensors only (there is no CPU path)");
    TORCH_CHECK(h.dim() >= 1, "licv::inject: h [..., d]");
    const int64_t d = h.size(-1);
    TORCH_CHECK(shift.scalar_type() == at::kFloat && shift.numel() == d,
                "shift must be an fp32 vector of the hidden size");
    c10::cuda::CUDAGuard guard(h.device());
    const Tensor hc = h.contiguous(), sc = shift.contiguous();
    Tensor out = at::empty(hc.sizes(), hc.options().dtype(out_dtype));
    const int64_t n_tok = d ? hc.numel() / d : 0;
    check_rc(licv_inject_fwd(hc.const_data_ptr(), sc.const_data_ptr<float>(), out.mutable_data_ptr(), n_tok,
                             (int)d, dtype_code(hc.scalar_type()), dtype_code(out_dtype),
                             (unsigned)round_flags, stream_of(hc)),
             "licv_inject_fwd");
    return out;
}

std::tuple<Tensor, Tensor> inject_bwd_cuda(const Tensor& h, const Tensor& g, const Tensor& shift,
                                           int64_t round_flags) {
    TORCH_CHECK(h.is_cuda() && g.is_cuda() && shift.is_cuda(), "licv::inject_bwd: CUDA tensors only");
    const int64_t d = h.size(-1);
    TORCH_CHECK(shift.scalar_type() == at::kFloat && shift.numel() == d,
                "shift must be an fp32 vector of the hidden size");
    c10::cuda::CUDAGuard guard(h.device());
    const Tensor hc = h.contiguous(), gc = g.contiguous(), sc = shift.contiguous();
    Tensor dh = at::empty_like(hc);
    Tensor d_shift = at::zeros(sc.sizes(), sc.options());
    const int64_t n_tok = d ? hc.numel() / d : 0;
    check_rc(licv_inject_bwd(hc.const_data_ptr(), gc.const_data_ptr(), sc.const_data_ptr<float>(),
                             dh.mutable_data_ptr(), d_shift.mutable_data_ptr<float>(), n_tok, (int)d,
                             dtype_code(hc.scalar_type()), dtype_code(gc.scalar_type()),
                             (unsigned)round_flags, stream_of(hc)),
             "licv_inject_bwd");
    return {dh, d_shift};
}

std::tuple<Tensor, Tensor, Tensor, Tensor> kd_loss_cuda(
    const Tensor& stu, const std::optional<Tensor>& tea, const std::optional<Tensor>& kl_tea_row,
    const std::optional<Tensor>& ce_label, const std::optional<Tensor>& counts, int64_t n_kl, int64_t n_ce,
    double temperature, double kl_eps, double hard_loss_weight, bool only_hard_loss, int64_t round_flags) {
    TORCH_CHECK(stu.is_cuda(), "licv::kd_loss: CUDA tensors only (there is no CPU path)");
    TORCH_CHECK(stu.dim() == 2 && stu.stride(1) == 1, "stu must be [R,V] with a contiguous last dimension");
    const int64_t R = stu.size(0), V = stu.size(1);
    if (tea.has_value())
        TORCH_CHECK(tea->is_cuda() && tea->dim() == 2 && tea->stride(1) == 1 && tea->size(1) == V &&
                        tea->scalar_type() == stu.scalar_type(),
                    "tea must be [Rt,V] of the student's dtype with a contiguous last dim");
    if (kl_tea_row.has_value())
        TORCH_CHECK(kl_tea_row->is_cuda() && kl_tea_row->scalar_type() == at::kInt && kl_tea_row->is_contiguous(),
                    "kl_tea_row: contiguous int32 on the device");
    if (ce_label.has_value())
        TORCH_CHECK(ce_label->is_cuda() && ce_label->scalar_type() == at::kLong && ce_label->is_contiguous(),
                    "ce_label: contiguous int64 on the device");
    if (counts.has_value())
        TORCH_CHECK(counts->is_cuda() && counts->scalar_type() == at::kInt && counts->is_contiguous(),
                    "counts: contiguous int32 on the device");
    c10::cuda::CUDAGuard guard(stu.device());
    Tensor losses = at::empty({3}, stu.options().dtype(at::kFloat));
    // functional form: the gradient goes to a new tensor (a traced graph must not see its input's
    // storage change under it); the in-place form is ops.kd_loss
    Tensor dstu = at::empty_strided(stu.sizes(), stu.strides(), stu.options());
    Tensor ws = kd_workspace(stu, R);
    const int64_t tea_stride = (tea.has_value() && tea->size(0) > 0) ? tea->stride(0) : V;
    check_rc(licv_kd_loss_fwd_bwd(stu.const_data_ptr(), dstu.mutable_data_ptr(), ptr_or_null(tea),
                                  static_cast<const int32_t*>(ptr_or_null(kl_tea_row)),
                                  static_cast<const int64_t*>(ptr_or_null(ce_label)),
                                  static_cast<const int32_t*>(ptr_or_null(counts)), n_kl, n_ce,
                                  (float)temperature, (float)kl_eps, (float)hard_loss_weight,
                                  only_hard_loss ? 1 : 0, 1.0f, losses.mutable_data_ptr<float>(),
                                  ws.mutable_data_ptr(), R, (int)V, R > 0 ? stu.stride(0) : V, tea_stride,
                                  dtype_code(stu.scalar_type()), (unsigned)round_flags, stream_of(stu)),
             "licv_kd_loss_fwd_bwd");
    return {losses[2].clone(), losses[0].clone(), losses[1].clone(), dstu};
}

Tensor get_mask_cuda(const Tensor& input_ids, const Tensor& mask_length, int64_t pad_token_id) {
    TORCH_CHECK(input_ids.is_cuda() && mask_length.is_cuda(), "licv::get_mask: CUDA tensors only");
    TORCH_CHECK(input_ids.dim() == 2, "input_ids [B,T]");
    c10::cuda::CUDAGuard guard(input_ids.device());
    const Tensor ids = input_ids.contiguous().to(at::kLong), ml = mask_length.contiguous().to(at::kLong);
    Tensor mask = at::empty(ids.sizes(), ids.options().dtype(at::kBool));
    check_rc(licv_get_mask(ids.const_data_ptr<int64_t>(), ml.const_data_ptr<int64_t>(), pad_token_id,
                           (int)ids.size(0), (int)ids.size(1), static_cast<uint8_t*>(mask.mutable_data_ptr()),
                           stream_of(ids)),
             "licv_get_mask");
    return mask;
}

// ---------------------------------------------------------------------------------------------
// Meta (fake) implementations: shapes and dtypes only, what a trace needs
// ---------------------------------------------------------------------------------------------
Tensor inject_meta(const Tensor& h, const Tensor& shift, c10::ScalarType out_dtype, int64_t) {
    return at::empty(h.sizes(), h.options().dtype(out_dtype));
}
std::tuple<Tensor, Tensor> inject_bwd_meta(const Tensor& h, const Tensor&, const Tensor& shift, int64_t) {
    return {at::empty(h.sizes(), h.options()), at::empty(shift.sizes(), shift.options().dtype(at::kFloat))};
}
std::tuple<Tensor, Tensor, Tensor, Tensor> kd_loss_meta(const Tensor& stu, const std::optional<Tensor>&,
                                                        const std::optional<Tensor>&, const std::optional<Tensor>&,
                                                        const std::optional<Tensor>&, int64_t, int64_t, double,
                                                        double, double, bool, int64_t) {
    auto scalar = [&] { return at::empty({}, stu.options().dtype(at::kFloat)); };
    return {scalar(), scalar(), scalar(), at::empty_strided(stu.sizes(), stu.strides(), stu.options())};
}
Tensor get_mask_meta(const Tensor& input_ids, const Tensor&, int64_t) {
    return at::empty(input_ids.sizes(), input_ids.options().dtype(at::kBool));
}

// ---------------------------------------------------------------------------------------------
// Autograd formulas
// ---------------------------------------------------------------------------------------------
Tensor call_inject(const Tensor& h, const Tensor& shift, c10::ScalarType out_dtype, int64_t flags) {
    static auto op = c10::Dispatcher::singleton()
                         .findSchemaOrThrow("licv::inject", "")
                         .typed<Tensor(const Tensor&, const Tensor&, c10::ScalarType, int64_t)>();
    return op.call(h, shift, out_dtype, flags);
}
std::tuple<Tensor, Tensor> call_inject_bwd(const Tensor& h, const Tensor& g, const Tensor& shift, int64_t flags) {
    static auto op = c10::Dispatcher::singleton()
                         .findSchemaOrThrow("licv::inject_bwd", "")
                         .typed<std::tuple<Tensor, Tensor>(const Tensor&, const Tensor&, const Tensor&, int64_t)>();
    return op.call(h, g, shift, flags);
}

// h, shift -> out.  Saves h only (the reference's autograd keeps several fp32 [B,T,d] intermediates
// per layer plus baukit's clone).
struct InjectFn : public torch::autograd::Function<InjectFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& h, const Tensor& shift, c10::ScalarType out_dtype,
                          int64_t flags) {
        at::AutoDispatchBelowADInplaceOrView below;
        ctx->save_for_backward({h, shift});
        ctx->saved_data["flags"] = flags;
        return call_inject(h, shift, out_dtype, flags);
    }
    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        auto [dh, d_shift] = call_inject_bwd(saved[0], grads[0].contiguous(), saved[1],
                                             ctx->saved_data["flags"].toInt());
        return {dh, d_shift, Tensor(), Tensor()};
    }
};
Tensor inject_autograd(const Tensor& h, const Tensor& shift, c10::ScalarType out_dtype, int64_t flags) {
    return InjectFn::apply(h, shift, out_dtype, flags);
}

using KdOut = std::tuple<Tensor, Tensor, Tensor, Tensor>;
struct KdLossFn : public torch::autograd::Function<KdLossFn> {
    static variable_list forward(AutogradContext* ctx, const Tensor& stu, const std::optional<Tensor>& tea,
                                 const std::optional<Tensor>& kl_tea_row, const std::optional<Tensor>& ce_label,
                                 const std::optional<Tensor>& counts, int64_t n_kl, int64_t n_ce, double temperature,
                                 double kl_eps, double hard_loss_weight, bool only_hard_loss, int64_t flags) {
        at::AutoDispatchBelowADInplaceOrView below;
        static auto op = c10::Dispatcher::singleton()
                             .findSchemaOrThrow("licv::kd_loss", "")
                             .typed<KdOut(const Tensor&, const std::optional<Tensor>&, const std::optional<Tensor>&,
                                          const std::optional<Tensor>&, const std::optional<Tensor>&, int64_t,
                                          int64_t, double, double, double, bool, int64_t)>();
        auto [total, kl, ce, dstu] = op.call(stu, tea, kl_tea_row, ce_label, counts, n_kl, n_ce, temperature,
                                             kl_eps, hard_loss_weight, only_hard_loss, flags);
        ctx->save_for_backward({dstu});
        ctx->mark_non_differentiable({kl, ce, dstu});
        return {total, kl, ce, dstu};
    }
    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const Tensor dstu = ctx->get_saved_variables()[0];
        variable_list out(12);
        out[0] = dstu * grads[0].to(dstu.scalar_type());
        return out;
    }
};
KdOut kd_loss_autograd(const Tensor& stu, const std::optional<Tensor>& tea, const std::optional<Tensor>& kl_tea_row,
                       const std::optional<Tensor>& ce_label, const std::optional<Tensor>& counts, int64_t n_kl,
                       int64_t n_ce, double temperature, double kl_eps, double hard_loss_weight,
                       bool only_hard_loss, int64_t flags) {
    auto r = KdLossFn::apply(stu, tea, kl_tea_row, ce_label, counts, n_kl, n_ce, temperature, kl_eps,
                             hard_loss_weight, only_hard_loss, flags);
    return {r[0], r[1], r[2], r[3]};
}

}  // namespace

TORCH_LIBRARY(licv, m) {
    m.def("inject(Tensor h, Tensor shift, ScalarType out_dtype, int round_flags) -> Tensor");
    m.def("inject_bwd(Tensor h, Tensor g, Tensor shift, int round_flags) -> (Tensor, Tensor)");
    m.def(
        "kd_loss(Tensor stu, Tensor? tea, Tensor? kl_tea_row, Tensor? ce_label, Tensor? counts, int n_kl, "
        "int n_ce, float temperature, float kl_eps, float hard_loss_weight, bool only_hard_loss, "
        "int round_flags) -> (Tensor, Tensor, Tensor, Tensor)");
    m.def("get_mask(Tensor input_ids, Tensor mask_length, int pad_token_id) -> Tensor");
}

TORCH_LIBRARY_IMPL(licv, CUDA, m) {
    m.impl("inject", &inject_cuda);
    m.impl("inject_bwd", &inject_bwd_cuda);
    m.impl("kd_loss", &kd_loss_cuda);
    m.impl("get_mask", &get_mask_cuda);
}

TORCH_LIBRARY_IMPL(licv, Meta, m) {
    m.impl("inject", &inject_meta);
    m.impl("inject_bwd", &inject_bwd_meta);
    m.impl("kd_loss", &kd_loss_meta);
    m.impl("get_mask", &get_mask_meta);
}

TORCH_LIBRARY_IMPL(licv, Autograd, m) {
    m.impl("inject", &inject_autograd);
    m.impl("kd_loss", &kd_loss_autograd);
}
