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
// Mapping.  A ROW GROUP of GT threads (a whole number of warps, 128 for d = 4096 bf16) owns whole
// rows: thread t of the group owns the 16-byte vectors {t, t + GT, ...} of a row (VPT of them,
// up to 32 elements), so a warp reads 512 contiguous bytes per instruction, and the thread's slice
// of the shift and of d_shift stays in registers for the kernel's lifetime.  A CTA holds G
// independent row groups, each synchronising on its own named barrier, and strides over tokens
// group by group; the grid is sized to the CTAs that are resident at once.  Per iteration a group
// stages TB tokens: every 128-bit load is issued before the first use; the per-token dot products
// are reduced with a multi-value butterfly (P values in log2(P) + log2(32/P) shuffle rounds, not
// 5 P), then across the group's warps through a double-buffered shared-memory slab (one barrier
// per iteration); results leave through 128-bit stores.  d_shift is summed in fp32 registers over
// all the group's tokens, then across the CTA's groups in shared memory, and leaves the CTA as one
// REDG.F32x4 per four columns.
//
// The places where the reference's eager chain rounds for bf16/fp16 hidden states are a
// compile-time mode for the per-element ones (RND: 0 none, 1 `h + s` stored in low precision,
// 2 also `y / ||y||`), and run-time flags for the two per-token norms.
//
// Neither kernel is a dense contraction: no tensor cores, the bound is HBM bandwidth.
#pragma once

#include <cstdlib>

#include "licv_common.cuh"

namespace licv {
namespace inject {

constexpr int kCtaThreads = 256;   // upper bound on threads per CTA (launch bound)
constexpr int kMaxWarps = kCtaThreads / kWarp;

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void group_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Sum P (a power of two <= 16) values per lane across the warp.  On return v[0] of lane L is the
// total of value index warp_value_index<P>(L); 31 shuffles in all for P = 32 values would be the
// limit case, for P = 4 it is 2 + 1 + 3 = 6 instead of 20.
template <int P>
__device__ __forceinline__ void warp_multi_sum(float (&v)[P]) {
    const int lane = threadIdx.x & 31;
    int o = 16;
#pragma unroll
    for (int n = P; n > 1; n >>= 1, o >>= 1) {
        const bool upper = lane & o;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
#pragma unroll
    for (; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}
template <int P>
__device__ __forceinline__ int warp_value_index(int lane) {
    int idx = 0, o = 16;
#pragma unroll
    for (int n = P; n > 1; n >>= 1, o >>= 1) idx = (idx << 1) | ((lane & o) ? 1 : 0);
    return idx;
}

// Sum P values over a row group.  `slab` (P * kMaxWarps floats, 16-byte aligned) must not be the
// slab of the group's previous call (double buffering makes one barrier per call sufficient).
template <int P>
__device__ __forceinline__ void group_sum(float (&v)[P], float* slab, int tg, int gt, int bar_id) {
    warp_multi_sum<P>(v);
    const int nw = gt >> 5;
    const int lane = tg & 31, warp = tg >> 5;
    if (nw == 1) {
        // single-warp group: gather the P totals back into every lane
        const float mine = v[0];
#pragma unroll
        for (int i = 0; i < P; ++i) {
            // the lane whose value index is i and whose low bits are zero
            int src = 0, o = 16;
#pragma unroll
            for (int n = P, b = i; n > 1; n >>= 1, o >>= 1) src |= ((b & (n >> 1)) ? o : 0);
            v[i] = __shfl_sync(0xffffffffu, mine, src);
        }
        return;
    }
    if ((lane & (32 / P - 1)) == 0) slab[warp * P + warp_value_index<P>(lane)] = v[0];
    group_barrier(bar_id, gt);
#pragma unroll
    for (int i = 0; i < P; ++i) v[i] = 0.f;
    for (int w = 0; w < nw; ++w) {
        if constexpr (P % 4 == 0) {
#pragma unroll
            for (int i = 0; i < P; i += 4) {
                const float4 q = *reinterpret_cast<const float4*>(slab + w * P + i);
                v[i] += q.x; v[i + 1] += q.y; v[i + 2] += q.z; v[i + 3] += q.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < P; ++i) v[i] += slab[w * P + i];
        }
    }
}

__host__ __device__ constexpr int pow2_at_least(int n) { return n <= 1 ? 1 : (n <= 2 ? 2 : (n <= 4 ? 4 : (n <= 8 ? 8 : 16))); }

// the thread's slice of the shift vector (rounded to the hidden dtype when the reference's shift
// tensor is itself low precision)
template <int HDT, int VPT, int RND>
__device__ __forceinline__ void load_shift(const float* __restrict__ shift, int nvec, int tg, int gt,
                                           float (&s)[VPT][Fmt<HDT>::kPerVec]) {
    constexpr int EPV = Fmt<HDT>::kPerVec;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int j = tg + k * gt;
#pragma unroll
        for (int e = 0; e < EPV; e += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < nvec) v = *reinterpret_cast<const float4*>(shift + (int64_t)j * EPV + e);
            s[k][e] = v.x; s[k][e + 1] = v.y; s[k][e + 2] = v.z; s[k][e + 3] = v.w;
        }
        if (RND >= 1) {
#pragma unroll
            for (int e = 0; e < EPV; ++e) s[k][e] = Fmt<HDT>::round(s[k][e]);
        }
    }
}

struct Args {
    const uint4* h;
    const uint4* g;       // backward only
    const float* shift;
    uint4* out;           // forward: out; backward: dh (may be null)
    float* d_shift;       // backward only: [d], accumulated (atomics)
    float* rows;          // backward only: if set, CTA b adds its d_shift into rows[b & row_mask][d]
    int row_mask;         //   (n_rows - 1, n_rows a power of two): fewer CTAs per atomic address
    int64_t n_tok;
    int nvec;             // 16-byte vectors per row of h
    int gt;               // threads per row group
    unsigned flags;       // LICV_ROUND_NH / LICV_ROUND_NY (per-token, run time)
};

template <int HDT, int ODT, int VPT, int TB, int RND>
__global__ void __launch_bounds__(kCtaThreads, 2) fwd_kernel(Args a) {
    constexpr int EPV = Fmt<HDT>::kPerVec;
    constexpr int OV = (Fmt<ODT>::kBytes > Fmt<HDT>::kBytes) ? 2 : 1;  // out vectors per in vector
    constexpr int OPV = Fmt<ODT>::kPerVec;
    constexpr int P = pow2_at_least(2 * TB);
    __shared__ __align__(16) float slab[2][kMaxWarps * P];

    const int gt = a.gt, nvec = a.nvec;
    const int grp = threadIdx.x / gt, tg = threadIdx.x - grp * gt;
    const int groups_per_cta = blockDim.x / gt;
    float* my_slab0 = slab[0] + (grp * (gt >> 5)) * P;
    float* my_slab1 = slab[1] + (grp * (gt >> 5)) * P;

    // hint this thread's first loads into L2 while the previous kernel drains (see prefetch_l2):
    // one lane per 128-byte line
    if ((tg & 7) == 0) {
        const int64_t t0 = ((int64_t)blockIdx.x * groups_per_cta + grp) * TB;
#pragma unroll
        for (int b = 0; b < TB; ++b) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = tg + k * gt;
                if (t0 + b < a.n_tok && j < nvec) prefetch_l2(a.h + (t0 + b) * nvec + j);
            }
        }
    }
    // no early pdl_launch_dependents() here: the forward leaves a free CTA slot per SM, and a
    // successor that moves in early slows this kernel down more than it gains (measured 3.5 vs
    // 3.1 us per 256-token launch); the trigger is implicit at completion
    pdl_wait();   // the previous kernel's results (h, the shift, ...) are visible from here on
    float s[VPT][EPV];
    load_shift<HDT, VPT, RND>(a.shift, nvec, tg, gt, s);

    const int64_t stride = (int64_t)gridDim.x * groups_per_cta * TB;
    int it = 0;
    for (int64_t t0 = ((int64_t)blockIdx.x * groups_per_cta + grp) * TB; t0 < a.n_tok;
         t0 += stride, ++it) {
        uint4 hv[TB][VPT];
#pragma unroll
        for (int b = 0; b < TB; ++b) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = tg + k * gt;
                hv[b][k] = make_uint4(0u, 0u, 0u, 0u);
                if (t0 + b < a.n_tok && j < nvec) hv[b][k] = ld_stream(a.h + (t0 + b) * nvec + j);
            }
        }
        float acc[P];
#pragma unroll
        for (int i = 0; i < P; ++i) acc[i] = 0.f;
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
                    if (RND >= 1) y = Fmt<HDT>::round(y);
                    hh = fmaf(x[e], x[e], hh);
                    yy = fmaf(y, y, yy);
                    if (RND >= 1) x[e] = y;
                }
                // low-precision y is exactly representable in the hidden dtype: keep it packed
                // in place of h, the second phase needs nothing else
                if (RND >= 1) hv[b][k] = pack<HDT>(x);
            }
            acc[2 * b] = hh;
            acc[2 * b + 1] = yy;
        }
        group_sum<P>(acc, (it & 1) ? my_slab1 : my_slab0, tg, gt, grp + 1);
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            if (t0 + b >= a.n_tok) break;
            float nh = sqrt_approx(acc[2 * b]);
            float ny = sqrt_approx(acc[2 * b + 1]);
            if (a.flags & LICV_ROUND_NH) nh = Fmt<HDT>::round(nh);
            if (a.flags & LICV_ROUND_NY) ny = Fmt<HDT>::round(ny);
            const float inv_ny = rcp_approx(ny);  // ny == 0 -> inf -> 0*inf = NaN, like the reference
            const float scale = inv_ny * nh;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = tg + k * gt;
                if (j >= nvec) continue;
                float x[EPV];
                unpack<HDT>(hv[b][k], x);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    if (RND == 0) {
                        x[e] = (x[e] + s[k][e]) * scale;
                    } else if (RND == 1) {
                        x[e] = x[e] * scale;
                    } else {
                        x[e] = Fmt<HDT>::round(x[e] * inv_ny) * nh;
                    }
                }
                uint4* dst = a.out + ((t0 + b) * nvec + j) * OV;
#pragma unroll
                for (int o = 0; o < OV; ++o) st_vec(dst + o, pack<ODT>(x + o * OPV));
            }
        }
    }
}

template <int HDT, int GDT, int VPT, int TB, int RND>
__global__ void __launch_bounds__(kCtaThreads, 2) bwd_kernel(Args a) {
    constexpr int EPV = Fmt<HDT>::kPerVec;
    constexpr int GV = (Fmt<GDT>::kBytes > Fmt<HDT>::kBytes) ? 2 : 1;  // g vectors per h vector
    constexpr int GPV = Fmt<GDT>::kPerVec;
    constexpr int P = pow2_at_least(3 * TB);
    __shared__ __align__(16) float slab[2][kMaxWarps * P];
    extern __shared__ __align__(16) float ds_smem[];   // [groups_per_cta - 1][nvec * EPV] when G > 1

    const int gt = a.gt, nvec = a.nvec;
    const int grp = threadIdx.x / gt, tg = threadIdx.x - grp * gt;
    const int groups_per_cta = blockDim.x / gt;
    float* my_slab0 = slab[0] + (grp * (gt >> 5)) * P;
    float* my_slab1 = slab[1] + (grp * (gt >> 5)) * P;

    pdl_launch_dependents();
    pdl_wait();   // the previous kernel's results (h, the shift, ...) are visible from here on
    float s[VPT][EPV];
    load_shift<HDT, VPT, RND>(a.shift, nvec, tg, gt, s);

    float ds[VPT][EPV];
#pragma unroll
    for (int k = 0; k < VPT; ++k)
#pragma unroll
        for (int e = 0; e < EPV; ++e) ds[k][e] = 0.f;

    const int64_t stride = (int64_t)gridDim.x * groups_per_cta * TB;
    int it = 0;
    for (int64_t t0 = ((int64_t)blockIdx.x * groups_per_cta + grp) * TB; t0 < a.n_tok;
         t0 += stride, ++it) {
        uint4 hv[TB][VPT];
        uint4 gv[TB][VPT][GV];
#pragma unroll
        for (int b = 0; b < TB; ++b) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = tg + k * gt;
                const bool live = (t0 + b < a.n_tok) && (j < nvec);
                hv[b][k] = make_uint4(0u, 0u, 0u, 0u);
                if (live) hv[b][k] = ld_stream(a.h + (t0 + b) * nvec + j);
#pragma unroll
                for (int o = 0; o < GV; ++o) {
                    gv[b][k][o] = make_uint4(0u, 0u, 0u, 0u);
                    // g may be overwritten by dh (same thread, after this read): coherent load
                    if (live) gv[b][k][o] = ld_plain(a.g + ((t0 + b) * nvec + j) * GV + o);
                }
            }
        }
        float acc[P];
#pragma unroll
        for (int i = 0; i < P; ++i) acc[i] = 0.f;
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
                    if (RND >= 1) y = Fmt<HDT>::round(y);
                    hh = fmaf(x[e], x[e], hh);
                    yy = fmaf(y, y, yy);
                    yg = fmaf(y, gg[e], yg);
                }
            }
            acc[3 * b] = hh;
            acc[3 * b + 1] = yy;
            acc[3 * b + 2] = yg;
        }
        group_sum<P>(acc, (it & 1) ? my_slab1 : my_slab0, tg, gt, grp + 1);
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            if (t0 + b >= a.n_tok) break;
            float nh = sqrt_approx(acc[3 * b]);
            float ny = sqrt_approx(acc[3 * b + 1]);
            if (a.flags & LICV_ROUND_NH) nh = Fmt<HDT>::round(nh);
            if (a.flags & LICV_ROUND_NY) ny = Fmt<HDT>::round(ny);
            const float inv_ny = rcp_approx(ny);
            const float r = nh * inv_ny;               // ||h|| / ||y||
            const float c = acc[3 * b + 2] * inv_ny;   // y^ . g
            const float ky = -r * c * inv_ny;          // g_y = r g + ky y
            const float kh = c * rcp_approx(nh);       // dh  = g_y + kh h
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = tg + k * gt;
                if (j >= nvec) continue;
                float x[EPV], gg[EPV];
                unpack<HDT>(hv[b][k], x);
#pragma unroll
                for (int o = 0; o < GV; ++o) unpack<GDT>(gv[b][k][o], gg + o * GPV);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    float y = x[e] + s[k][e];
                    if (RND >= 1) y = Fmt<HDT>::round(y);
                    const float gy = fmaf(r, gg[e], ky * y);
                    ds[k][e] += gy;
                    x[e] = fmaf(kh, x[e], gy);
                }
                if (a.out != nullptr) st_vec(a.out + (t0 + b) * nvec + j, pack<HDT>(x));
            }
        }
    }

    // d_shift: groups 1.. hand their partial sums to group 0 through shared memory, group 0
    // issues one REDG.F32x4 per four columns
    if (groups_per_cta > 1) {
        const int row_floats = nvec * EPV;
        if (grp > 0) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = tg + k * gt;
                if (j >= nvec) continue;
#pragma unroll
                for (int e = 0; e < EPV; e += 4)
                    *reinterpret_cast<float4*>(ds_smem + (grp - 1) * row_floats + j * EPV + e) =
                        make_float4(ds[k][e], ds[k][e + 1], ds[k][e + 2], ds[k][e + 3]);
            }
        }
        __syncthreads();
        if (grp > 0) return;
        for (int q = 0; q < groups_per_cta - 1; ++q) {
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = tg + k * gt;
                if (j >= nvec) continue;
#pragma unroll
                for (int e = 0; e < EPV; e += 4) {
                    const float4 v =
                        *reinterpret_cast<const float4*>(ds_smem + q * row_floats + j * EPV + e);
                    ds[k][e] += v.x; ds[k][e + 1] += v.y; ds[k][e + 2] += v.z; ds[k][e + 3] += v.w;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int j = tg + k * gt;
        if (j >= nvec) continue;
#pragma unroll
        for (int e = 0; e < EPV; e += 4)
            red_add_v4(a.d_shift + (int64_t)j * EPV + e, ds[k][e], ds[k][e + 1], ds[k][e + 2],
                       ds[k][e + 3]);
    }
}

// ---------------------------------------------------------------------------------------------
// launch configuration
// ---------------------------------------------------------------------------------------------
struct RowPlan {
    int gt = 0;    // threads per row group
    int vpt = 0;   // vectors per thread (template value: 1, 2, 4 or 8)
};

inline int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

// 4 vectors per thread (8 once a row no longer fits 256 threads x 4), whole warps, at most 256
// threads per row group
inline bool plan_row(int nvec, int max_vpt, RowPlan* p) {
    static const int forced_gt = env_int("LICV_INJECT_GROUP_THREADS", 0);  // tuning knob
    int gt = ((nvec + 3) / 4 + 31) / 32 * 32;
    if (gt > kCtaThreads) gt = kCtaThreads;
    if (forced_gt > 0) gt = forced_gt;
    if (gt > kCtaThreads || gt % 32 != 0) return false;
    int need = (nvec + gt - 1) / gt, vpt = 1;
    while (vpt < need) vpt *= 2;
    if (vpt > max_vpt) return false;
    p->gt = gt;
    p->vpt = vpt;
    return true;
}

template <typename K>
int resident_ctas_per_sm(K kernel, int threads, size_t smem) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) !=
            cudaSuccess || per_sm < 1)
        per_sm = 1;
    return per_sm;
}

inline int check_row(int d, int dtype, int* nvec) {
    if (dtype != LICV_F32 && dtype != LICV_BF16 && dtype != LICV_F16) return LICV_ERR_BAD_DTYPE;
    const int per = dtype == LICV_F32 ? 4 : 8;
    if (d <= 0 || d % per != 0) return LICV_ERR_BAD_DIM;
    *nvec = d / per;
    return LICV_OK;
}

inline int rnd_mode(unsigned flags) {
    if (!(flags & LICV_ROUND_Y)) return 0;
    return (flags & LICV_ROUND_T) ? 2 : 1;
}

}  // namespace inject
}  // namespace licv
