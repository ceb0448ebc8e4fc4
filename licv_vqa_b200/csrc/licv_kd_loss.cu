// L-ICV distillation loss: KL(teacher || student) + hard_loss_weight * CE, forward and backward.
//
//   q = softmax(stu/T), p = softmax(tea/T)
//   kl_n = sum_v p (ln(p+eps) - ln(q+eps))          loss_kl = T^2/N sum_n kl_n
//   w = p q/(q+eps), W_n = sum_v w                  d stu[n,j] = (T/N)(q_j W_n - w_j)
//   ce_n = lse(stu) - stu[label]                    d stu[n,j] += (lambda/M)(softmax(stu)_j - [j=label])
//
// Replaces VQAICVModule.calculate_kl_divergence (reference icv_src/icv_module.py:121-134, ~12
// eager kernels forward + ~15 backward over fp32 [N,V] copies), the two boolean-mask gathers
// (:108-111, rows are addressed through an index list instead of being copied), the HF-internal
// shifted cross-entropy consumed at :94-98,115-117 and the combine at :100-101,107-119.
//
// This file holds the GENERIC kernel: one CTA per student row, any vocabulary size, any dtype,
// any row alignment.  A row is pulled from HBM once; the later sweeps (the eps in the logarithms
// makes the gradient need W_n, i.e. a full sweep, before the first gradient element can be
// written) re-read it through L2 where it is still resident (a row is 64-128 KB, L2 is 126 MB).
// HBM traffic is therefore the algorithmic 3 e V bytes per KL row (2 e V for a CE-only row),
// and dstu may alias stu.  licv_kd_loss_cluster.cu holds the cluster kernel (every exponential
// evaluated once, rows cached on chip) that the dispatcher prefers when a row fits a cluster.
#include "licv_common.cuh"
#include "licv_kd_loss.cuh"

namespace licv {
namespace {

constexpr int kThreads = 512;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int NV>
__device__ __forceinline__ void cta_sum(float (&v)[NV], float* slab) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();  // slab reuse
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
template <int NV>
__device__ __forceinline__ void cta_max(float (&v)[NV], float* slab) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_max(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) slab[warp * NV + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = -INFINITY;
    for (int w = 0; w < nw; ++w) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = fmaxf(v[i], slab[w * NV + i]);
    }
}

// Visit every element j of a row pair: f(j, x_j, t_j) (t_j = 0 when there is no teacher row).
// Rows whose starts share the same 16-byte phase go through 128-bit loads with a scalar head and
// tail; otherwise element by element (coalesced 2- or 4-byte accesses).
template <int DT, bool HAS_T, typename F>
__device__ __forceinline__ void visit_row(const char* xrow, const char* trow, int V, bool vec_ok,
                                          F&& f) {
    constexpr int EPV = Fmt<DT>::kPerVec;
    constexpr int EB = Fmt<DT>::kBytes;
    int head = 0, nvec = 0;
    if (vec_ok) {
        head = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(xrow) & 15u)) & 15u) / EB);
        if (head > V) head = V;
        nvec = (V - head) / EPV;
    }
    const int body_end = head + nvec * EPV;
    // scalar head + tail (fewer than 2*EPV elements in the vector case, the whole row otherwise)
    const int n_scalar = head + (V - body_end);
    for (int i = threadIdx.x; i < n_scalar; i += blockDim.x) {
        const int j = i < head ? i : body_end + (i - head);
        const float x = load_elem<DT>(xrow, j);
        const float t = HAS_T ? load_elem<DT>(trow, j) : 0.f;
        f(j, x, t);
    }
    const uint4* xv = reinterpret_cast<const uint4*>(xrow + (size_t)head * EB);
    const uint4* tv = reinterpret_cast<const uint4*>(trow + (size_t)head * EB);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float x[EPV], t[EPV];
        unpack<DT>(ld_plain(xv + i), x);
        if (HAS_T) {
            unpack<DT>(ld_plain(tv + i), t);
        }
#pragma unroll
        for (int e = 0; e < EPV; ++e) f(head + i * EPV + e, x[e], HAS_T ? t[e] : 0.f);
    }
}

struct RowConst {
    float inv_t;      // 1/T
    float c2;         // log2(e) (tempered logits are formed explicitly when rounding) or log2(e)/T
    bool round_tempered;
};

template <int DT>
__device__ __forceinline__ float tempered(float x, const RowConst& k) {
    // z = logits / T as the reference stores it (icv_module.py:122-123)
    float z = x * k.inv_t;
    if (k.round_tempered) z = Fmt<DT>::round(z);
    return z;
}

// WANT_T: also d loss / d temperature (learnable_t, icv_module.py:49-52).  Both in-place divides by
// T (icv_module.py:122-123) and the T^2 factor (:133) depend on it; with z = logits / T
//   dKL_n/dT = -(1/T) [ sum_j (q_j W - w_j) zs_j + sum_j p_j (a_j - abar) zt_j ],
//   a = ln(p+eps) - ln(q+eps) + p/(p+eps),  abar = sum_v p_v a_v = KL_n + sum_v p_v^2/(p_v+eps)
// and d loss/dT = 2 T mean_n KL_n + T^2 mean_n dKL_n/dT.  Two more row sums in sweep 3 and one more
// sweep; the row's dKL_n/dT goes to row_loss[2 n_rows + r], the total to out_losses[3].
template <int DT, bool WANT_T>
__global__ void __launch_bounds__(kThreads)
kd_loss_generic_kernel(KdArgs a) {
    __shared__ float slab[(kThreads / 32) * 5];
    __shared__ int s_last;
    constexpr int EB = Fmt<DT>::kBytes;

    const float T = a.temperature;
    RowConst k;
    k.inv_t = 1.0f / T;
    k.round_tempered = (a.round_flags & LICV_ROUND_TEMPERED) && DT != LICV_F32 && T != 1.0f;
    k.c2 = kLog2e;
    const bool t_is_one = (T == 1.0f);

    const int64_t n_kl = a.counts ? (int64_t)a.counts[0] : a.n_kl;
    const int64_t n_ce = a.counts ? (int64_t)a.counts[1] : a.n_ce;
    const bool use_kl = !a.only_hard_loss;
    const bool use_ce = a.ce_label != nullptr;
    const float kl_w = use_kl ? a.grad_scale * T / (float)n_kl : 0.f;   // times (q W - w)
    const float ce_w = use_ce ? a.grad_scale * (a.only_hard_loss ? 1.0f : a.hard_loss_weight) /
                                    (float)n_ce
                              : 0.f;
    float* row_kl = a.row_loss;
    float* row_ce = a.row_loss + a.n_rows;
    float* row_dt = a.row_loss + 2 * a.n_rows;   // WANT_T only

    for (int64_t r = blockIdx.x; r < a.n_rows; r += gridDim.x) {
        const int64_t tr = !use_kl ? -1 : (a.kl_tea_row ? (int64_t)a.kl_tea_row[r] : r);
        const int64_t lab = use_ce ? a.ce_label[r] : -100;
        const bool has_kl = tr >= 0;
        const bool has_ce = lab != -100;
        const char* xrow = static_cast<const char*>(a.stu) + (size_t)r * a.stu_stride * EB;
        char* grow = a.dstu ? static_cast<char*>(a.dstu) + (size_t)r * a.stu_stride * EB : nullptr;
        const char* trow = has_kl ? static_cast<const char*>(a.tea) + (size_t)tr * a.tea_stride * EB
                                  : xrow;
        const uint32_t px = (uint32_t)(reinterpret_cast<uintptr_t>(xrow) & 15u);
        const bool vec_ok = (!has_kl || (reinterpret_cast<uintptr_t>(trow) & 15u) == px) &&
                            (!grow || (reinterpret_cast<uintptr_t>(grow) & 15u) == px);

        if (!has_kl && !has_ce) {
            // neither loss touches this row: its gradient is zero
            if (grow) {
                for (int j = threadIdx.x; j < a.vocab; j += blockDim.x) store_elem<DT>(grow, j, 0.f);
            }
            if (threadIdx.x == 0) {
                row_kl[r] = 0.f;
                row_ce[r] = 0.f;
                if (WANT_T) row_dt[r] = 0.f;
            }
            continue;
        }

        // ---- sweep 1 (HBM): maxima of the tempered student / teacher logits ------------------
        float mx[2] = {-INFINITY, -INFINITY};
        if (has_kl) {
            visit_row<DT, true>(xrow, trow, a.vocab, vec_ok, [&](int, float x, float t) {
                mx[0] = fmaxf(mx[0], x);
                mx[1] = fmaxf(mx[1], t);
            });
        } else {
            visit_row<DT, false>(xrow, trow, a.vocab, vec_ok,
                                 [&](int, float x, float) { mx[0] = fmaxf(mx[0], x); });
        }
        cta_max<2>(mx, slab);
        // the label logit, fetched before the gradient sweep may overwrite the row in place
        const bool lab_ok = has_ce && lab >= 0 && lab < a.vocab;
        float x_lab = 0.f;
        if (threadIdx.x == 0 && lab_ok) x_lab = load_elem<DT>(xrow, lab);
        const float x_max = mx[0];                       // raw student maximum (CE)
        const float zs_max = tempered<DT>(mx[0], k);     // monotone: max of tempered = tempered max
        const float zt_max = tempered<DT>(mx[1], k);

        // ---- sweep 2 (L2): partition sums ----------------------------------------------------
        float sm[3] = {0.f, 0.f, 0.f};  // sum exp(zs - max), sum exp(zt - max), sum exp(x - xmax)
        if (has_kl) {
            visit_row<DT, true>(xrow, trow, a.vocab, vec_ok, [&](int, float x, float t) {
                sm[0] += ex2((tempered<DT>(x, k) - zs_max) * k.c2);
                sm[1] += ex2((tempered<DT>(t, k) - zt_max) * k.c2);
                if (has_ce && !t_is_one) sm[2] += ex2((x - x_max) * kLog2e);
            });
        } else {
            visit_row<DT, false>(xrow, trow, a.vocab, vec_ok, [&](int, float x, float) {
                sm[2] += ex2((x - x_max) * kLog2e);
            });
        }
        cta_sum<3>(sm, slab);
        if (has_kl && has_ce && t_is_one) sm[2] = sm[0];
        const float inv_ls = 1.0f / sm[0];
        const float inv_lt = 1.0f / sm[1];
        const float inv_lce = 1.0f / sm[2];

        // ---- sweep 3 (L2): KL value and W_n --------------------------------------------------
        float kw[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // KL/ln2, W; WANT_T: sum p^2/(p+eps), sum p a zt, sum p zt
        if (has_kl) {
            visit_row<DT, true>(xrow, trow, a.vocab, vec_ok, [&](int, float x, float t) {
                const float q = ex2((tempered<DT>(x, k) - zs_max) * k.c2) * inv_ls;
                const float zt = tempered<DT>(t, k);
                const float p = ex2((zt - zt_max) * k.c2) * inv_lt;
                const float rq = __frcp_rn(q + a.kl_eps);
                // ln(p+eps) - ln(q+eps) = ln((p+eps)/(q+eps))
                const float lr = lg2((p + a.kl_eps) * rq);
                kw[0] = fmaf(p, lr, kw[0]);
                kw[1] = fmaf(p * q, rq, kw[1]);
                if (WANT_T) {
                    const float ppe = p * __frcp_rn(p + a.kl_eps);
                    kw[2] = fmaf(p, ppe, kw[2]);
                    kw[3] = fmaf(p * fmaf(lr, kLn2, ppe), zt, kw[3]);
                    kw[4] = fmaf(p, zt, kw[4]);
                }
            });
            cta_sum<5>(kw, slab);
        }
        const float W = kw[1];
        if (WANT_T) {
            // one more sweep, before the gradient may overwrite the row: sum_j (q_j W - w_j) zs_j
            float s1[1] = {0.f};
            if (has_kl) {
                visit_row<DT, true>(xrow, trow, a.vocab, vec_ok, [&](int, float x, float t) {
                    const float zs = tempered<DT>(x, k);
                    const float q = ex2((zs - zs_max) * k.c2) * inv_ls;
                    const float p = ex2((tempered<DT>(t, k) - zt_max) * k.c2) * inv_lt;
                    const float w = p * q * __frcp_rn(q + a.kl_eps);
                    s1[0] = fmaf(fmaf(q, W, -w), zs, s1[0]);
                });
                cta_sum<1>(s1, slab);
            }
            if (threadIdx.x == 0) {
                const float abar = kw[0] * kLn2 + kw[2];
                row_dt[r] = has_kl ? -(s1[0] + kw[3] - abar * kw[4]) / T : 0.f;
            }
        }

        // ---- sweep 4 (L2 read, HBM write): gradient ------------------------------------------
        if (grow) {
            const float ce_on = has_ce ? ce_w : 0.f;
            auto grad = [&](int j, float x, float t) -> float {
                float gsum = 0.f;
                float q = 0.f;
                if (has_kl) {
                    q = ex2((tempered<DT>(x, k) - zs_max) * k.c2) * inv_ls;
                    const float p = ex2((tempered<DT>(t, k) - zt_max) * k.c2) * inv_lt;
                    const float w = p * q * __frcp_rn(q + a.kl_eps);
                    gsum = kl_w * fmaf(q, W, -w);
                }
                if (has_ce) {
                    const float sx = (has_kl && t_is_one) ? q : ex2((x - x_max) * kLog2e) * inv_lce;
                    gsum = fmaf(ce_on, sx - (j == (int)lab ? 1.0f : 0.f), gsum);
                }
                return gsum;
            };
            // same traversal as visit_row, writing as it goes (dstu may alias stu: each element
            // is read and written by the same thread)
            constexpr int EPV = Fmt<DT>::kPerVec;
            int head = 0, nvec = 0;
            if (vec_ok) {
                head = (int)(((16u - px) & 15u) / EB);
                if (head > a.vocab) head = a.vocab;
                nvec = (a.vocab - head) / EPV;
            }
            const int body_end = head + nvec * EPV;
            const int n_scalar = head + (a.vocab - body_end);
            for (int i = threadIdx.x; i < n_scalar; i += blockDim.x) {
                const int j = i < head ? i : body_end + (i - head);
                const float x = load_elem<DT>(xrow, j);
                const float t = has_kl ? load_elem<DT>(trow, j) : 0.f;
                store_elem<DT>(grow, j, grad(j, x, t));
            }
            const uint4* xv = reinterpret_cast<const uint4*>(xrow + (size_t)head * EB);
            const uint4* tv = reinterpret_cast<const uint4*>(trow + (size_t)head * EB);
            uint4* gv = reinterpret_cast<uint4*>(grow + (size_t)head * EB);
            for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
                float x[EPV], t[EPV];
                unpack<DT>(ld_plain(xv + i), x);
                if (has_kl) unpack<DT>(ld_plain(tv + i), t);
#pragma unroll
                for (int e = 0; e < EPV; ++e)
                    x[e] = grad(head + i * EPV + e, x[e], has_kl ? t[e] : 0.f);
                st_vec(gv + i, pack<DT>(x));
            }
        }

        if (threadIdx.x == 0) {
            row_kl[r] = has_kl ? kw[0] * kLn2 : 0.f;
            float ce = 0.f;
            if (has_ce)  // an out-of-range label is an error in torch; poison the loss instead
                ce = lab_ok ? x_max + logf(sm[2]) - x_lab : __int_as_float(0x7fc00000);
            row_ce[r] = ce;
        }
    }

    // ---- last CTA reduces the per-row losses (fixed order: deterministic) ---------------------
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(a.counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        float tot[3] = {0.f, 0.f, 0.f};
        for (int64_t r = threadIdx.x; r < a.n_rows; r += blockDim.x) {
            tot[0] += __ldcg(row_kl + r);
            tot[1] += __ldcg(row_ce + r);
            if (WANT_T) tot[2] += __ldcg(row_dt + r);
        }
        cta_sum<3>(tot, slab);
        if (threadIdx.x == 0) {
            const float kl = use_kl ? tot[0] * T * T / (float)n_kl : 0.f;
            const float ce = use_ce ? tot[1] / (float)n_ce : 0.f;
            a.out_losses[0] = kl;
            a.out_losses[1] = ce;
            a.out_losses[2] = a.only_hard_loss ? ce : (use_ce ? fmaf(a.hard_loss_weight, ce, kl) : kl);
            if (WANT_T)   // d(T^2 mean KL)/dT; zero when the KL term is switched off
                a.out_losses[3] = use_kl ? (2.0f * T * tot[0] + T * T * tot[2]) / (float)n_kl : 0.f;
            *a.counter = 0u;  // leave the workspace ready for the next call
        }
    }
}

template <int DT>
__global__ void scale_inplace_kernel(void* x, int64_t n, const float* scale) {
    const float s = *scale;
    if (s == 1.0f) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        store_elem<DT>(x, i, load_elem<DT>(x, i) * s);
}

}  // namespace

int launch_kd_generic(const KdArgs& a, int dtype, cudaStream_t st, bool want_dtemp) {
    const int cap = device_info().sm_count * 2;
    int grid = (int)(a.n_rows < cap ? a.n_rows : cap);
    if (grid < 1) grid = 1;  // no rows: the finalising CTA still reports mean-of-empty = NaN
    if (want_dtemp) {
        switch (dtype) {
            case LICV_F32: kd_loss_generic_kernel<LICV_F32, true><<<grid, kThreads, 0, st>>>(a); break;
            case LICV_BF16: kd_loss_generic_kernel<LICV_BF16, true><<<grid, kThreads, 0, st>>>(a); break;
            default: kd_loss_generic_kernel<LICV_F16, true><<<grid, kThreads, 0, st>>>(a); break;
        }
        return (int)cudaGetLastError();
    }
    switch (dtype) {
        case LICV_F32: kd_loss_generic_kernel<LICV_F32, false><<<grid, kThreads, 0, st>>>(a); break;
        case LICV_BF16: kd_loss_generic_kernel<LICV_BF16, false><<<grid, kThreads, 0, st>>>(a); break;
        default: kd_loss_generic_kernel<LICV_F16, false><<<grid, kThreads, 0, st>>>(a); break;
    }
    return (int)cudaGetLastError();
}

}  // namespace licv

using namespace licv;

namespace {
// The stream kernel needs a few rows per CTA to reach its steady state (its first row pays for the
// TMA ring fill and a CTA-wide look at the first chunk): measured against the tensor-memory
// kernel, 32 002 bf16 logits, it wins from ~256 KL+CE rows / ~1000 CE-only rows on (profiles/
// r2_kd_small_rows.txt).  Below 4 rows per SM the tensor-memory kernel runs where it can.
int kd_pick(int vocab, int dtype, float temperature, bool kl_and_ce, int64_t n_rows, int* C, int* NV, int* NT) {
    const bool tmem_ok = kd_tmem_plan(vocab, dtype, temperature, kl_and_ce);
    const int sms = device_info().status == LICV_OK ? device_info().sm_count : 148;
    if (kd_stream_plan(vocab, dtype) && (!tmem_ok || n_rows >= (int64_t)4 * sms || kd_stream_forced()))
        return LICV_KD_KERNEL_STREAM;
    if (tmem_ok) return LICV_KD_KERNEL_TMEM;
    if (kd_cluster_plan(vocab, dtype, temperature, kl_and_ce, C, NV, NT)) return LICV_KD_KERNEL_CLUSTER;
    return LICV_KD_KERNEL_GENERIC;
}
}  // namespace

extern "C" int licv_kd_loss_plan(int vocab, int dtype, float temperature, int kl_and_ce, int64_t n_rows) {
    if (dtype != LICV_F32 && dtype != LICV_BF16 && dtype != LICV_F16) return LICV_ERR_BAD_DTYPE;
    if (vocab <= 0 || n_rows < 0) return LICV_ERR_BAD_ARGUMENT;
    int C = 0, NV = 0, NT = 0;
    return kd_pick(vocab, dtype, temperature, kl_and_ce != 0, n_rows, &C, &NV, &NT);
}

extern "C" int64_t licv_kd_loss_workspace_bytes(int64_t n_rows) {
    if (n_rows < 0) n_rows = 0;
    return 16 + ((2 * n_rows * (int64_t)sizeof(float) + 15) / 16) * 16;
}

static int kd_loss_entry(const void* stu, void* dstu, const void* tea,
                         const int32_t* kl_tea_row, const int64_t* ce_label,
                         const int32_t* counts, int64_t n_kl, int64_t n_ce,
                         float temperature, float kl_eps, float hard_loss_weight,
                         int only_hard_loss, float grad_scale, float* out_losses,
                         void* workspace, int64_t n_rows, int vocab, int64_t stu_stride,
                         int64_t tea_stride, int dtype, unsigned round_flags,
                         licv_stream_t stream, bool want_dtemp) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (dtype != LICV_F32 && dtype != LICV_BF16 && dtype != LICV_F16) return LICV_ERR_BAD_DTYPE;
    if (n_rows < 0 || vocab <= 0 || stu_stride < vocab || (tea && tea_stride < vocab))
        return LICV_ERR_BAD_ARGUMENT;
    if (!out_losses || !workspace) return LICV_ERR_NULL_POINTER;
    if (n_rows > 0 && !stu) return LICV_ERR_NULL_POINTER;
    if (!only_hard_loss && n_rows > 0 && !tea) return LICV_ERR_NULL_POINTER;
    if (only_hard_loss && !ce_label) return LICV_ERR_BAD_ARGUMENT;
    if (!aligned16(workspace)) return LICV_ERR_MISALIGNED;
    if (!(temperature > 0.f)) return LICV_ERR_BAD_ARGUMENT;

    KdArgs a;
    a.stu = stu; a.dstu = dstu; a.tea = tea;
    a.kl_tea_row = kl_tea_row;
    a.ce_label = ce_label;   // NULL disables the CE term (hard_loss_weight == 0 in the reference)
    a.counts = counts;
    a.n_kl = n_kl; a.n_ce = n_ce;
    a.temperature = temperature; a.kl_eps = kl_eps; a.hard_loss_weight = hard_loss_weight;
    a.only_hard_loss = only_hard_loss != 0;
    a.grad_scale = grad_scale;
    a.out_losses = out_losses;
    a.counter = static_cast<unsigned*>(workspace);
    a.row_loss = reinterpret_cast<float*>(static_cast<char*>(workspace) + 16);
    a.n_rows = n_rows; a.vocab = vocab; a.stu_stride = stu_stride; a.tea_stride = tea_stride;
    a.round_flags = round_flags;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (want_dtemp) return launch_kd_generic(a, dtype, st, true);   // the generic kernel holds the logits it needs
    int C = 0, NV = 0, NT = 0;
    switch (kd_pick(vocab, dtype, temperature, ce_label != nullptr && !only_hard_loss, n_rows, &C, &NV, &NT)) {
        case LICV_KD_KERNEL_STREAM: return launch_kd_stream(a, dtype, st);
        case LICV_KD_KERNEL_TMEM: return launch_kd_tmem(a, dtype, st);
        case LICV_KD_KERNEL_CLUSTER: return launch_kd_cluster(a, dtype, C, NV, NT, st);
        default: return launch_kd_generic(a, dtype, st, false);
    }
}

extern "C" int licv_kd_loss_fwd_bwd(const void* stu, void* dstu, const void* tea,
                                    const int32_t* kl_tea_row, const int64_t* ce_label,
                                    const int32_t* counts, int64_t n_kl, int64_t n_ce,
                                    float temperature, float kl_eps, float hard_loss_weight,
                                    int only_hard_loss, float grad_scale, float* out_losses,
                                    void* workspace, int64_t n_rows, int vocab, int64_t stu_stride,
                                    int64_t tea_stride, int dtype, unsigned round_flags,
                                    licv_stream_t stream) {
    return kd_loss_entry(stu, dstu, tea, kl_tea_row, ce_label, counts, n_kl, n_ce, temperature, kl_eps,
                         hard_loss_weight, only_hard_loss, grad_scale, out_losses, workspace, n_rows,
                         vocab, stu_stride, tea_stride, dtype, round_flags, stream, false);
}

extern "C" int64_t licv_kd_loss_dtemp_workspace_bytes(int64_t n_rows) {
    if (n_rows < 0) n_rows = 0;
    return 16 + ((3 * n_rows * (int64_t)sizeof(float) + 15) / 16) * 16;
}

extern "C" int licv_kd_loss_fwd_bwd_dtemp(const void* stu, void* dstu, const void* tea,
                                          const int32_t* kl_tea_row, const int64_t* ce_label,
                                          const int32_t* counts, int64_t n_kl, int64_t n_ce,
                                          float temperature, float kl_eps, float hard_loss_weight,
                                          int only_hard_loss, float grad_scale, float* out_losses4,
                                          void* workspace, int64_t n_rows, int vocab,
                                          int64_t stu_stride, int64_t tea_stride, int dtype,
                                          unsigned round_flags, licv_stream_t stream) {
    return kd_loss_entry(stu, dstu, tea, kl_tea_row, ce_label, counts, n_kl, n_ce, temperature, kl_eps,
                         hard_loss_weight, only_hard_loss, grad_scale, out_losses4, workspace, n_rows,
                         vocab, stu_stride, tea_stride, dtype, round_flags, stream, true);
}

extern "C" int licv_scale_inplace(void* x, int64_t n, const float* scale, int dtype,
                                  licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n < 0) return LICV_ERR_BAD_ARGUMENT;
    if (n == 0) return LICV_OK;
    if (!x || !scale) return LICV_ERR_NULL_POINTER;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int threads = 256;
    int64_t blocks = (n + threads * 8 - 1) / (threads * 8);
    const int cap = device_info().sm_count * 8;
    const int grid = (int)(blocks < cap ? blocks : cap);
    switch (dtype) {
        case LICV_F32: scale_inplace_kernel<LICV_F32><<<grid, threads, 0, st>>>(x, n, scale); break;
        case LICV_BF16: scale_inplace_kernel<LICV_BF16><<<grid, threads, 0, st>>>(x, n, scale); break;
        case LICV_F16: scale_inplace_kernel<LICV_F16><<<grid, threads, 0, st>>>(x, n, scale); break;
        default: return LICV_ERR_BAD_DTYPE;
    }
    return (int)cudaGetLastError();
}
