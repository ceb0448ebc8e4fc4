// Residual-stream injection of the learnable in-context vector, forward and backward (sm_100a).
//
//   forward   out = (h + s) / ||h + s||_2 * ||h||_2          per token, no eps
//   backward  y^ = y/||y||, r = ||h||/||y||, c = y^.g
//             g_y = r (g - y^ c)      dh = g_y + c h/||h||      ds = sum_tokens g_y
//
// Replaces intervention_function (reference icv_src/icv_model/icv_intervention.py:61-86: five
// eager kernels forward, ~10 plus saved fp32 intermediates backward) with ONE pass over HBM per
// direction: forward reads h and writes out (2 e n d bytes), backward reads h and g and writes dh
// (3 e n d bytes); the shift vector and the d_shift accumulators live in registers.
//
// Mapping: a CTA owns whole rows.  Thread t owns the 16-byte vectors {t, t + blockDim, ...} of a
// row (VPT of them), so a warp reads 512 contiguous bytes per instruction; its slice of the shift
// and of d_shift stays in registers for the CTA's lifetime.  Each iteration stages TB tokens:
// all TB*VPT 128-bit loads are issued before the first use, the per-token dot products are
// reduced by warp shuffle then across warps through a double-buffered shared-memory slab (one
// __syncthreads per iteration), and the results are written back with 128-bit stores.  The grid
// is sized to the number of CTAs that are resident at once (SMs x occupancy) and strides over
// token groups.  d_shift is summed in fp32 registers over all the CTA's tokens and leaves the CTA
// as one REDG.F32x4 per four columns.
//
// Neither kernel is a dense contraction: no tensor cores, the bound is HBM bandwidth.
#include <cstdlib>

#include "licv_common.cuh"

namespace licv {
namespace {

constexpr int kMaxThreads = 512;
constexpr int kMaxWarps = kMaxThreads / kWarp;

// Sum NV values over the CTA.  `slab` holds kMaxWarps*NV floats and must not be the slab used by
// the previous call (double buffering makes one barrier per call sufficient).
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* slab) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    const int nw = blockDim.x >> 5;
    if (nw == 1) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) slab[warp * NV + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.f;
    for (int w = 0; w < nw; ++w) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] += slab[w * NV + i];
    }
}

// the thread's slice of the shift vector, rounded to the hidden dtype when the reference's shift
// tensor is itself low precision
template <int HDT, int VPT>
__device__ __forceinline__ void load_shift(const float* __restrict__ shift, int nvec, unsigned flags,
                                           float (&s)[VPT][Fmt<HDT>::kPerVec]) {
    constexpr int EPV = Fmt<HDT>::kPerVec;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int j = threadIdx.x + k * blockDim.x;
#pragma unroll
        for (int e = 0; e < EPV; e += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < nvec) v = *reinterpret_cast<const float4*>(shift + (int64_t)j * EPV + e);
            s[k][e] = v.x; s[k][e + 1] = v.y; s[k][e + 2] = v.z; s[k][e + 3] = v.w;
        }
        if (flags & LICV_ROUND_Y) {
#pragma unroll
            for (int e = 0; e < EPV; ++e) s[k][e] = Fmt<HDT>::round(s[k][e]);
        }
    }
}

template <int HDT, int ODT, int VPT, int TB>
__global__ void __launch_bounds__(kMaxThreads)
inject_fwd_kernel(const uint4* __restrict__ h, const float* __restrict__ shift,
                  uint4* __restrict__ out, int64_t n_tok, int nvec, unsigned flags) {
    constexpr int EPV = Fmt<HDT>::kPerVec;
    constexpr int OV = (Fmt<ODT>::kBytes > Fmt<HDT>::kBytes) ? 2 : 1;  // out vectors per in vector
    constexpr int OPV = Fmt<ODT>::kPerVec;
    __shared__ __align__(16) float slab[2][kMaxWarps * 2 * TB];

    float s[VPT][EPV];
    load_shift<HDT, VPT>(shift, nvec, flags, s);
    const bool ry = flags & LICV_ROUND_Y;

    int it = 0;
    for (int64_t t0 = (int64_t)blockIdx.x * TB; t0 < n_tok; t0 += (int64_t)gridDim.x * TB, ++it) {
        uint4 hv[TB][VPT];
#pragma unroll
        for (int b = 0; b < TB; ++b) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = threadIdx.x + k * blockDim.x;
                hv[b][k] = make_uint4(0u, 0u, 0u, 0u);
                if (t0 + b < n_tok && j < nvec) hv[b][k] = ld_stream(h + (t0 + b) * nvec + j);
            }
        }
        float acc[2 * TB];
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            float hh = 0.f, yy = 0.f;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                float x[EPV];
                unpack<HDT>(hv[b][k], x);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    float y = x[e] + s[k][e];
                    if (ry) y = Fmt<HDT>::round(y);
                    hh = fmaf(x[e], x[e], hh);
                    yy = fmaf(y, y, yy);
                }
            }
            acc[2 * b] = hh;
            acc[2 * b + 1] = yy;
        }
        block_sum<2 * TB>(acc, slab[it & 1]);
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            if (t0 + b >= n_tok) break;
            float nh = sqrtf(acc[2 * b]);
            float ny = sqrtf(acc[2 * b + 1]);
            if (flags & LICV_ROUND_NH) nh = Fmt<HDT>::round(nh);
            if (flags & LICV_ROUND_NY) ny = Fmt<HDT>::round(ny);
            const float inv_ny = 1.0f / ny;  // ny == 0 -> inf -> 0*inf = NaN, like the reference
            const bool rt = flags & LICV_ROUND_T;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = threadIdx.x + k * blockDim.x;
                if (j >= nvec) continue;
                float x[EPV];
                unpack<HDT>(hv[b][k], x);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    float y = x[e] + s[k][e];
                    if (ry) y = Fmt<HDT>::round(y);
                    float t = y * inv_ny;
                    if (rt) t = Fmt<HDT>::round(t);
                    x[e] = t * nh;
                }
                uint4* dst = out + ((t0 + b) * nvec + j) * OV;
#pragma unroll
                for (int o = 0; o < OV; ++o) st_vec(dst + o, pack<ODT>(x + o * OPV));
            }
        }
    }
}

template <int HDT, int GDT, int VPT, int TB>
__global__ void __launch_bounds__(kMaxThreads)
inject_bwd_kernel(const uint4* __restrict__ h, const uint4* g, const float* __restrict__ shift,
                  uint4* dh, float* __restrict__ d_shift, int64_t n_tok, int nvec, unsigned flags) {
    constexpr int EPV = Fmt<HDT>::kPerVec;
    constexpr int GV = (Fmt<GDT>::kBytes > Fmt<HDT>::kBytes) ? 2 : 1;  // g vectors per h vector
    constexpr int GPV = Fmt<GDT>::kPerVec;
    __shared__ __align__(16) float slab[2][kMaxWarps * 3 * TB];

    float s[VPT][EPV];
    load_shift<HDT, VPT>(shift, nvec, flags, s);
    const bool ry = flags & LICV_ROUND_Y;

    float ds[VPT][EPV];
#pragma unroll
    for (int k = 0; k < VPT; ++k)
#pragma unroll
        for (int e = 0; e < EPV; ++e) ds[k][e] = 0.f;

    int it = 0;
    for (int64_t t0 = (int64_t)blockIdx.x * TB; t0 < n_tok; t0 += (int64_t)gridDim.x * TB, ++it) {
        uint4 hv[TB][VPT];
        uint4 gv[TB][VPT][GV];
#pragma unroll
        for (int b = 0; b < TB; ++b) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = threadIdx.x + k * blockDim.x;
                const bool live = (t0 + b < n_tok) && (j < nvec);
                hv[b][k] = make_uint4(0u, 0u, 0u, 0u);
                if (live) hv[b][k] = ld_stream(h + (t0 + b) * nvec + j);
#pragma unroll
                for (int o = 0; o < GV; ++o) {
                    gv[b][k][o] = make_uint4(0u, 0u, 0u, 0u);
                    // g may be overwritten by dh (same thread, after this read): coherent load
                    if (live) gv[b][k][o] = ld_plain(g + ((t0 + b) * nvec + j) * GV + o);
                }
            }
        }
        float acc[3 * TB];
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            float hh = 0.f, yy = 0.f, yg = 0.f;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                float x[EPV], gg[EPV];
                unpack<HDT>(hv[b][k], x);
#pragma unroll
                for (int o = 0; o < GV; ++o) unpack<GDT>(gv[b][k][o], gg + o * GPV);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    float y = x[e] + s[k][e];
                    if (ry) y = Fmt<HDT>::round(y);
                    hh = fmaf(x[e], x[e], hh);
                    yy = fmaf(y, y, yy);
                    yg = fmaf(y, gg[e], yg);
                }
            }
            acc[3 * b] = hh;
            acc[3 * b + 1] = yy;
            acc[3 * b + 2] = yg;
        }
        block_sum<3 * TB>(acc, slab[it & 1]);
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            if (t0 + b >= n_tok) break;
            float nh = sqrtf(acc[3 * b]);
            float ny = sqrtf(acc[3 * b + 1]);
            if (flags & LICV_ROUND_NH) nh = Fmt<HDT>::round(nh);
            if (flags & LICV_ROUND_NY) ny = Fmt<HDT>::round(ny);
            const float inv_ny = 1.0f / ny;
            const float r = nh * inv_ny;               // ||h|| / ||y||
            const float c = acc[3 * b + 2] * inv_ny;   // y^ . g
            const float ky = r * c * inv_ny;           // g_y = r g - ky y
            const float kh = c / nh;                   // dh  = g_y + kh h
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = threadIdx.x + k * blockDim.x;
                if (j >= nvec) continue;
                float x[EPV], gg[EPV];
                unpack<HDT>(hv[b][k], x);
#pragma unroll
                for (int o = 0; o < GV; ++o) unpack<GDT>(gv[b][k][o], gg + o * GPV);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    float y = x[e] + s[k][e];
                    if (ry) y = Fmt<HDT>::round(y);
                    const float gy = fmaf(r, gg[e], -ky * y);
                    ds[k][e] += gy;
                    x[e] = fmaf(kh, x[e], gy);
                }
                if (dh != nullptr) st_vec(dh + (t0 + b) * nvec + j, pack<HDT>(x));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int j = threadIdx.x + k * blockDim.x;
        if (j >= nvec) continue;
#pragma unroll
        for (int e = 0; e < EPV; e += 4)
            red_add_v4(d_shift + (int64_t)j * EPV + e, ds[k][e], ds[k][e + 1], ds[k][e + 2],
                       ds[k][e + 3]);
    }
}

// ---------------------------------------------------------------------------------------------
// launch configuration
// ---------------------------------------------------------------------------------------------
struct RowPlan {
    int threads = 0;
    int vpt = 0;
};

int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

// threads per CTA and vectors per thread for a row of `nvec` 16-byte vectors: one vector per
// thread up to 256 threads, then two (d = 4096 bf16 -> 256 x 2), then 512 threads x 2 or x 4
bool plan_row(int nvec, RowPlan* p) {
    static const int forced = env_int("LICV_INJECT_THREADS", 0);  // tuning knob
    int threads = 32;
    if (forced > 0) {
        threads = forced;
    } else {
        while (threads < 256 && threads < nvec) threads *= 2;
        if ((nvec + threads - 1) / threads > 2) threads = kMaxThreads;
    }
    int vpt = (nvec + threads - 1) / threads;
    if (vpt == 3) vpt = 4;
    if (vpt < 1 || vpt > 4 || threads % 32 != 0 || threads > kMaxThreads) return false;
    p->threads = threads;
    p->vpt = vpt;
    return true;
}

template <typename K>
int resident_ctas(K kernel, int threads) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess ||
        per_sm < 1)
        per_sm = 1;
    return per_sm * device_info().sm_count;
}

template <int HDT, int ODT, int VPT, int TB>
int launch_fwd(const void* h, const float* shift, void* out, int64_t n_tok, int nvec, int threads,
               unsigned flags, cudaStream_t st) {
    auto kern = inject_fwd_kernel<HDT, ODT, VPT, TB>;
    static const int resident[2] = {resident_ctas(kern, 256), resident_ctas(kern, kMaxThreads)};
    const int cap = resident[threads > 256];
    const int64_t groups = (n_tok + TB - 1) / TB;
    const int grid = (int)(groups < cap ? groups : cap);
    kern<<<grid, threads, 0, st>>>(static_cast<const uint4*>(h), shift, static_cast<uint4*>(out),
                                   n_tok, nvec, flags);
    return (int)cudaGetLastError();
}

template <int HDT, int GDT, int VPT, int TB>
int launch_bwd(const void* h, const void* g, const float* shift, void* dh, float* ds, int64_t n_tok,
               int nvec, int threads, unsigned flags, cudaStream_t st) {
    auto kern = inject_bwd_kernel<HDT, GDT, VPT, TB>;
    static const int resident[2] = {resident_ctas(kern, 256), resident_ctas(kern, kMaxThreads)};
    const int cap = resident[threads > 256];
    const int64_t groups = (n_tok + TB - 1) / TB;
    const int grid = (int)(groups < cap ? groups : cap);
    kern<<<grid, threads, 0, st>>>(static_cast<const uint4*>(h), static_cast<const uint4*>(g), shift,
                                   static_cast<uint4*>(dh), ds, n_tok, nvec, flags);
    return (int)cudaGetLastError();
}

// tokens staged per iteration: fill the machine first (TB = 1 while there are fewer token groups
// than resident CTAs), then deepen the per-thread load batch to TB*VPT = 8 vectors (4 for the
// backward, which stages two tensors)
template <int HDT, int ODT, int VPT>
int dispatch_fwd_tb(const void* h, const float* s, void* out, int64_t n, int nvec, int threads,
                    unsigned flags, cudaStream_t st) {
    constexpr int TBmax = 8 / VPT;
    const int64_t fill = (int64_t)device_info().sm_count * 4;
    if (TBmax > 1 && n >= fill * TBmax)
        return launch_fwd<HDT, ODT, VPT, TBmax>(h, s, out, n, nvec, threads, flags, st);
    return launch_fwd<HDT, ODT, VPT, 1>(h, s, out, n, nvec, threads, flags, st);
}

template <int HDT, int GDT, int VPT>
int dispatch_bwd_tb(const void* h, const void* g, const float* s, void* dh, float* ds, int64_t n,
                    int nvec, int threads, unsigned flags, cudaStream_t st) {
    constexpr int GV = (Fmt<GDT>::kBytes > Fmt<HDT>::kBytes) ? 2 : 1;
    // staged vectors per thread = TB * VPT * (1 + GV): 16 for the bf16/fp16 d=4096 shape (fits
    // the 128-register budget without spilling), 8 or fewer elsewhere
    constexpr int kFit = 8 / (VPT * (1 + GV));
    constexpr int TBmax = (HDT != LICV_F32 && VPT == 2 && GV == 1) ? 4
                          : (kFit >= 4 ? 4 : (kFit >= 2 ? 2 : 1));
    const int64_t fill = (int64_t)device_info().sm_count * 2;
    if (TBmax > 1 && n >= fill * TBmax)
        return launch_bwd<HDT, GDT, VPT, TBmax>(h, g, s, dh, ds, n, nvec, threads, flags, st);
    return launch_bwd<HDT, GDT, VPT, 1>(h, g, s, dh, ds, n, nvec, threads, flags, st);
}

template <int HDT, int ODT>
int dispatch_fwd_vpt(const void* h, const float* s, void* out, int64_t n, int nvec,
                     const RowPlan& p, unsigned flags, cudaStream_t st) {
    switch (p.vpt) {
        case 1: return dispatch_fwd_tb<HDT, ODT, 1>(h, s, out, n, nvec, p.threads, flags, st);
        case 2: return dispatch_fwd_tb<HDT, ODT, 2>(h, s, out, n, nvec, p.threads, flags, st);
        default: return dispatch_fwd_tb<HDT, ODT, 4>(h, s, out, n, nvec, p.threads, flags, st);
    }
}

template <int HDT, int GDT>
int dispatch_bwd_vpt(const void* h, const void* g, const float* s, void* dh, float* ds, int64_t n,
                     int nvec, const RowPlan& p, unsigned flags, cudaStream_t st) {
    switch (p.vpt) {
        case 1: return dispatch_bwd_tb<HDT, GDT, 1>(h, g, s, dh, ds, n, nvec, p.threads, flags, st);
        case 2: return dispatch_bwd_tb<HDT, GDT, 2>(h, g, s, dh, ds, n, nvec, p.threads, flags, st);
        default: return dispatch_bwd_tb<HDT, GDT, 4>(h, g, s, dh, ds, n, nvec, p.threads, flags, st);
    }
}

int check_row(int d, int dtype, int* nvec) {
    if (dtype != LICV_F32 && dtype != LICV_BF16 && dtype != LICV_F16) return LICV_ERR_BAD_DTYPE;
    const int per = dtype == LICV_F32 ? 4 : 8;
    if (d <= 0 || d % per != 0) return LICV_ERR_BAD_DIM;
    *nvec = d / per;
    return LICV_OK;
}

}  // namespace
}  // namespace licv

using namespace licv;

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
    RowPlan p;
    if (!plan_row(nvec, &p)) return LICV_ERR_BAD_DIM;
    if (h_dtype == LICV_F32) round_flags = 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (h_dtype == LICV_F32)
        return dispatch_fwd_vpt<LICV_F32, LICV_F32>(h, shift, out, n_tokens, nvec, p, round_flags, st);
    if (h_dtype == LICV_BF16) {
        if (out_dtype == LICV_F32)
            return dispatch_fwd_vpt<LICV_BF16, LICV_F32>(h, shift, out, n_tokens, nvec, p, round_flags, st);
        return dispatch_fwd_vpt<LICV_BF16, LICV_BF16>(h, shift, out, n_tokens, nvec, p, round_flags, st);
    }
    if (out_dtype == LICV_F32)
        return dispatch_fwd_vpt<LICV_F16, LICV_F32>(h, shift, out, n_tokens, nvec, p, round_flags, st);
    return dispatch_fwd_vpt<LICV_F16, LICV_F16>(h, shift, out, n_tokens, nvec, p, round_flags, st);
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
    RowPlan p;
    if (!plan_row(nvec, &p)) return LICV_ERR_BAD_DIM;
    if (h_dtype == LICV_F32) round_flags = 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (h_dtype == LICV_F32)
        return dispatch_bwd_vpt<LICV_F32, LICV_F32>(h, g, shift, dh, d_shift, n_tokens, nvec, p, round_flags, st);
    if (h_dtype == LICV_BF16) {
        if (g_dtype == LICV_F32)
            return dispatch_bwd_vpt<LICV_BF16, LICV_F32>(h, g, shift, dh, d_shift, n_tokens, nvec, p, round_flags, st);
        return dispatch_bwd_vpt<LICV_BF16, LICV_BF16>(h, g, shift, dh, d_shift, n_tokens, nvec, p, round_flags, st);
    }
    if (g_dtype == LICV_F32)
        return dispatch_bwd_vpt<LICV_F16, LICV_F32>(h, g, shift, dh, d_shift, n_tokens, nvec, p, round_flags, st);
    return dispatch_bwd_vpt<LICV_F16, LICV_F16>(h, g, shift, dh, d_shift, n_tokens, nvec, p, round_flags, st);
}
