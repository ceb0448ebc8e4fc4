// C-ABI entry points of the injection kernels (argument checks, row plan, dtype dispatch).
#include "licv_inject_impl.cuh"

using namespace licv;
using namespace licv::inject;

extern "C" int licv_inject_fwd(const void* h, const float* shift, void* out, int64_t n_tokens, int d,
                               int h_dtype, int out_dtype, unsigned round_flags,
                               licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n_tokens < 0) return LICV_ERR_BAD_ARGUMENT;
    int nvec = 0;
    if (int rc = check_row(d, h_dtype, &nvec)) return rc;
    if (out_dtype != h_dtype && out_dtype != LICV_F32) return LICV_ERR_BAD_DTYPE;
    if (n_tokens == 0) return LICV_OK;
    if (!h || !shift || !out) return LICV_ERR_NULL_POINTER;
    if (!aligned16(h) || !aligned16(shift) || !aligned16(out)) return LICV_ERR_MISALIGNED;
    if (h == out) return LICV_ERR_BAD_ARGUMENT;
    Launch L;
    L.n_tok = n_tokens;
    L.nvec = nvec;
    L.flags = h_dtype == LICV_F32 ? 0u : round_flags;
    L.stream = reinterpret_cast<cudaStream_t>(stream);
    if (!plan_row(nvec, 8, &L.plan)) return LICV_ERR_BAD_DIM;
    Args a{};
    a.h = static_cast<const uint4*>(h);
    a.shift = shift;
    a.out = static_cast<uint4*>(out);
    a.n_tok = n_tokens;
    a.nvec = nvec;
    a.flags = L.flags;
    switch (h_dtype) {
        case LICV_F32: return run_fwd<LICV_F32>(a, L, out_dtype);
        case LICV_BF16: return run_fwd<LICV_BF16>(a, L, out_dtype);
        default: return run_fwd<LICV_F16>(a, L, out_dtype);
    }
}

extern "C" int licv_inject_bwd(const void* h, const void* g, const float* shift, void* dh,
                               float* d_shift, int64_t n_tokens, int d, int h_dtype, int g_dtype,
                               unsigned round_flags, licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n_tokens < 0) return LICV_ERR_BAD_ARGUMENT;
    int nvec = 0;
    if (int rc = check_row(d, h_dtype, &nvec)) return rc;
    if (g_dtype != h_dtype && g_dtype != LICV_F32) return LICV_ERR_BAD_DTYPE;
    if (n_tokens == 0) return LICV_OK;
    if (!h || !g || !shift || !d_shift) return LICV_ERR_NULL_POINTER;
    if (!aligned16(h) || !aligned16(g) || !aligned16(shift) || !aligned16(d_shift) ||
        (dh && !aligned16(dh)))
        return LICV_ERR_MISALIGNED;
    if (dh == h || (dh == g && g_dtype != h_dtype)) return LICV_ERR_BAD_ARGUMENT;
    Launch L;
    L.n_tok = n_tokens;
    L.nvec = nvec;
    L.flags = h_dtype == LICV_F32 ? 0u : round_flags;
    L.stream = reinterpret_cast<cudaStream_t>(stream);
    if (!plan_row(nvec, 8, &L.plan)) return LICV_ERR_BAD_DIM;
    Args a{};
    a.h = static_cast<const uint4*>(h);
    a.g = static_cast<const uint4*>(g);
    a.shift = shift;
    a.out = static_cast<uint4*>(dh);
    a.d_shift = d_shift;
    a.n_tok = n_tokens;
    a.nvec = nvec;
    a.flags = L.flags;
    switch (h_dtype) {
        case LICV_F32: return run_bwd<LICV_F32>(a, L, g_dtype);
        case LICV_BF16: return run_bwd<LICV_BF16>(a, L, g_dtype);
        default: return run_bwd<LICV_F16>(a, L, g_dtype);
    }
}

// ---------------------------------------------------------------------------------------------
// spread backward (see include/licv_b200.h)
// ---------------------------------------------------------------------------------------------
static int elem_bytes(int dtype) { return dtype == LICV_F32 ? 4 : 2; }

extern "C" int licv_inject_bwd_rows(int64_t n_tokens, int d, int h_dtype, int g_dtype) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n_tokens < 0) return LICV_ERR_BAD_ARGUMENT;
    int nvec = 0;
    if (int rc = check_row(d, h_dtype, &nvec)) return rc;
    if (g_dtype != h_dtype && g_dtype != LICV_F32) return LICV_ERR_BAD_DTYPE;
    if (n_tokens == 0 || !pipe_row(nvec)) return 1;
    return spread_rows_for(pipe_partial_ctas(n_tokens, nvec, elem_bytes(h_dtype), elem_bytes(g_dtype)));
}

extern "C" int licv_inject_bwd_spread(const void* h, const void* g, const float* shift, void* dh,
                                      float* rows, int n_rows, int64_t n_tokens, int d, int h_dtype,
                                      int g_dtype, unsigned round_flags, licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n_tokens < 0) return LICV_ERR_BAD_ARGUMENT;
    int nvec = 0;
    if (int rc = check_row(d, h_dtype, &nvec)) return rc;
    if (g_dtype != h_dtype && g_dtype != LICV_F32) return LICV_ERR_BAD_DTYPE;
    if (n_rows < 1 || n_rows > 64 || (n_rows & (n_rows - 1)) != 0) return LICV_ERR_BAD_ARGUMENT;
    if (!rows) return LICV_ERR_NULL_POINTER;
    if (!aligned16(rows)) return LICV_ERR_MISALIGNED;
    if (n_tokens == 0 || !pipe_row(nvec))   // rows that the TMA ring does not take: replica 0
        return licv_inject_bwd(h, g, shift, dh, rows, n_tokens, d, h_dtype, g_dtype, round_flags, stream);
    if (!h || !g || !shift) return LICV_ERR_NULL_POINTER;
    if (!aligned16(h) || !aligned16(g) || !aligned16(shift) || (dh && !aligned16(dh)))
        return LICV_ERR_MISALIGNED;
    if (dh == h || (dh == g && g_dtype != h_dtype)) return LICV_ERR_BAD_ARGUMENT;
    Launch L;
    L.n_tok = n_tokens;
    L.nvec = nvec;
    L.flags = h_dtype == LICV_F32 ? 0u : round_flags;
    L.stream = reinterpret_cast<cudaStream_t>(stream);
    L.partial_rows = pipe_partial_ctas(n_tokens, nvec, elem_bytes(h_dtype), elem_bytes(g_dtype));
    if (!plan_row(nvec, 8, &L.plan)) return LICV_ERR_BAD_DIM;
    Args a{};
    a.h = static_cast<const uint4*>(h);
    a.g = static_cast<const uint4*>(g);
    a.shift = shift;
    a.out = static_cast<uint4*>(dh);
    a.d_shift = nullptr;
    a.rows = rows;
    a.row_mask = n_rows - 1;
    a.n_tok = n_tokens;
    a.nvec = nvec;
    a.flags = L.flags;
    switch (h_dtype) {
        case LICV_F32: return run_bwd<LICV_F32>(a, L, g_dtype);
        case LICV_BF16: return run_bwd<LICV_BF16>(a, L, g_dtype);
        default: return run_bwd<LICV_F16>(a, L, g_dtype);
    }
}
