// Small kernels and host plumbing around the two hot kernels: device query, the encoder product
// (a1+a2) and its backward, get_mask (a6), row pairing / label preparation (a6+a7+a9), and the
// fused clip + AdamW step on the flat ICV parameter buffer (f2).
#include <mutex>

#include <cooperative_groups.h>
#include <cstdlib>

#include "licv_common.cuh"

namespace licv {

thread_local int tl_grid_cap = 0;

const DeviceInfo& device_info() {
    // one entry per device ordinal; a process here drives one GPU, so a small fixed table will do
    static DeviceInfo table[64];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        static DeviceInfo bad;
        cudaGetLastError();
        return bad;
    }
    DeviceInfo& info = table[dev];
    if (info.sm_count == 0) {
        std::lock_guard<std::mutex> lock(mu);
        if (info.sm_count == 0) {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
                info.cc_major = prop.major;
                info.cc_minor = prop.minor;
                info.status = (prop.major == 10) ? LICV_OK : LICV_ERR_NO_DEVICE;
                info.sm_count = prop.multiProcessorCount;
            } else {
                cudaGetLastError();
                info.sm_count = -1;
            }
        }
    }
    return info;
}

namespace {

__device__ __forceinline__ float sigmoidf(float a) { return 1.0f / (1.0f + __expf(-a)); }

__global__ void icv_scale_kernel(const float* __restrict__ alpha, const float* __restrict__ vec,
                                 float* __restrict__ icv, int n_layers, int d, int use_sigmoid) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t n4 = (int64_t)n_layers * d / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(i * 4 / d);
        float a = alpha[l];
        if (use_sigmoid) a = sigmoidf(a);
        float4 v = reinterpret_cast<const float4*>(vec)[i];
        v.x *= a; v.y *= a; v.z *= a; v.w *= a;
        reinterpret_cast<float4*>(icv)[i] = v;
    }
}

// one CTA per layer: d_vec = alpha_eff * d_icv, d_alpha = (d_icv . vec) * dsigmoid
__global__ void __launch_bounds__(1024)
icv_scale_bwd_kernel(const float* __restrict__ alpha, const float* __restrict__ vec,
                     const float* __restrict__ d_icv, float* __restrict__ d_vec,
                     float* __restrict__ d_alpha, int d, int use_sigmoid) {
    pdl_launch_dependents();
    const int l = blockIdx.x;
    // the parameters are last step's: hint this thread's share into L2 ahead of the wait
    if ((threadIdx.x & 7) == 0 && (int)threadIdx.x < d / 4) prefetch_l2(vec + (int64_t)l * d + threadIdx.x * 4);
    pdl_wait();
    __shared__ float slab[32];
    float a = alpha[l];
    float da = 1.0f;
    if (use_sigmoid) {
        a = sigmoidf(a);
        da = a * (1.0f - a);
    }
    const float4* g4 = reinterpret_cast<const float4*>(d_icv + (int64_t)l * d);
    const float4* v4 = reinterpret_cast<const float4*>(vec + (int64_t)l * d);
    float4* o4 = reinterpret_cast<float4*>(d_vec + (int64_t)l * d);
    float dot = 0.f;
    for (int i = threadIdx.x; i < d / 4; i += blockDim.x) {
        const float4 g = g4[i];
        const float4 v = v4[i];
        dot = fmaf(g.x, v.x, dot); dot = fmaf(g.y, v.y, dot);
        dot = fmaf(g.z, v.z, dot); dot = fmaf(g.w, v.w, dot);
        o4[i] = make_float4(a * g.x, a * g.y, a * g.z, a * g.w);
    }
    if (d_alpha == nullptr) return;
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) slab[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += slab[w];
        d_alpha[l] = s * da;
    }
}

// The tail of a backward pass in ONE launch (reduce_rows + icv_scale_bwd + the optimizer's sum of
// squares).  d_icv = sum of the replicas the spread backward launches left (zero-filled again if
// `clear`), d_vec (+)= alpha_eff * d_icv, d_alpha (+)= (d_icv . vec) * dsigmoid, partial[l] =
// |prescale * (d_vec row, d_alpha)|^2 of the STORED values, summed in a fixed order (shuffle
// butterfly, the warps in index order, then the CTAs of a layer in index order).
struct FinishArgs {
    float* rows; int n_rows; int64_t layer_stride;
    const float* alpha; const float* vec;
    float* d_icv; float* d_vec; float* d_alpha; float* partial;
    float prescale; int d, use_sigmoid, accumulate, clear;
};

// columns [lo, hi) (in float4 units) of layer l, one float4 column per thread and sweep:
// -> this CTA's (d_icv . vec, sum of squares of the stored d_vec * prescale) in thread 0
__device__ __forceinline__ float2 finish_columns(const FinishArgs& a, int l, int lo, int hi, float alpha_eff,
                                                 float (*slab)[32]) {
    const int d4 = a.d / 4;
    const float4* v4 = reinterpret_cast<const float4*>(a.vec + (int64_t)l * a.d);
    float4* o4 = reinterpret_cast<float4*>(a.d_vec + (int64_t)l * a.d);
    float dot = 0.f, ss = 0.f;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        float4* src = reinterpret_cast<float4*>(a.rows + l * a.layer_stride) + i;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        // the first 16 replicas (all of them for the shapes licv_inject_bwd_rows answers) in ONE
        // batch of loads: the launch is a chain of L2 round trips, not bandwidth
        int p = 0;
        {
            float4 r[16];
#pragma unroll
            for (int u = 0; u < 16; ++u)
                r[u] = u < a.n_rows ? __ldcg(src + (int64_t)u * d4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                g.x += r[u].x; g.y += r[u].y; g.z += r[u].z; g.w += r[u].w;
            }
            p = a.n_rows < 16 ? a.n_rows : 16;
        }
        for (; p < a.n_rows; ++p) {
            const float4 r = __ldcg(src + (int64_t)p * d4);
            g.x += r.x; g.y += r.y; g.z += r.z; g.w += r.w;
        }
        if (a.clear) {   // the replicas are ready for the next backward pass
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = 0; q < a.n_rows; ++q) src[(int64_t)q * d4] = z;
        }
        if (a.d_icv) reinterpret_cast<float4*>(a.d_icv + (int64_t)l * a.d)[i] = g;
        const float4 v = v4[i];
        dot = fmaf(g.x, v.x, dot); dot = fmaf(g.y, v.y, dot);
        dot = fmaf(g.z, v.z, dot); dot = fmaf(g.w, v.w, dot);
        float4 o = make_float4(alpha_eff * g.x, alpha_eff * g.y, alpha_eff * g.z, alpha_eff * g.w);
        if (a.accumulate) {
            const float4 old = o4[i];
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        o4[i] = o;
        const float s0 = o.x * a.prescale, s1 = o.y * a.prescale, s2 = o.z * a.prescale, s3 = o.w * a.prescale;
        ss = fmaf(s0, s0, ss); ss = fmaf(s1, s1, ss); ss = fmaf(s2, s2, ss); ss = fmaf(s3, s3, ss);
    }
    dot = warp_sum(dot);
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) {
        slab[0][threadIdx.x >> 5] = dot;
        slab[1][threadIdx.x >> 5] = ss;
    }
    __syncthreads();
    float sd = 0.f, sq = 0.f;
    if (threadIdx.x == 0) {
        for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) {
            sd += slab[0][w];
            sq += slab[1][w];
        }
    }
    return make_float2(sd, sq);
}

// thread 0 of the layer's (first) CTA: d_alpha and the layer's squared norm
__device__ __forceinline__ void finish_layer(const FinishArgs& a, int l, float sd, float sq, float dsig) {
    if (a.d_alpha) {
        float x = sd * dsig;
        if (a.accumulate) x += a.d_alpha[l];
        a.d_alpha[l] = x;
        x *= a.prescale;
        sq = fmaf(x, x, sq);
    }
    if (a.partial) a.partial[l] = sq;
}

__device__ __forceinline__ float2 alpha_eff_of(const FinishArgs& a, int l) {   // (alpha_eff, dsigmoid)
    float al = a.alpha[l];
    if (!a.use_sigmoid) return make_float2(al, 1.0f);
    al = sigmoidf(al);
    return make_float2(al, al * (1.0f - al));
}

// one CTA per layer (narrow layers)
__global__ void __launch_bounds__(1024, 1) icv_grad_finish_kernel(FinishArgs a) {
    pdl_launch_dependents();
    const int l = blockIdx.x;
    const int d4 = a.d / 4;
    // the parameters are last step's: hint this thread's share into L2 ahead of the wait
    if ((threadIdx.x & 7) == 0 && (int)threadIdx.x < d4) prefetch_l2(a.vec + (int64_t)l * a.d + threadIdx.x * 4);
    pdl_wait();
    __shared__ float slab[2][32];
    const float2 ae = alpha_eff_of(a, l);
    const float2 s = finish_columns(a, l, 0, d4, ae.x, slab);
    if (threadIdx.x == 0) finish_layer(a, l, s.x, s.y, ae.y);
}

// wide layers: a thread-block cluster of kFinishCluster CTAs per layer, CTA c takes the c-th share
// of the columns (one SM per layer could not pull its 16 replicas x 16 KB out of L2 and zero them
// fast enough: 10 us for 32 layers), and the CTAs' (dot, sum of squares) meet in CTA 0's shared
// memory (DSMEM stores, one cluster barrier), added in CTA order.
constexpr int kFinishCluster = 4;
__global__ void __launch_bounds__(256) icv_grad_finish_cluster_kernel(FinishArgs a) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    pdl_launch_dependents();
    const int l = blockIdx.x / kFinishCluster, c = blockIdx.x % kFinishCluster;
    const int d4 = a.d / 4;
    const int per = (d4 + kFinishCluster - 1) / kFinishCluster;
    const int lo = c * per, hi = lo + per < d4 ? lo + per : d4;
    if ((threadIdx.x & 7) == 0 && lo + (int)threadIdx.x < hi)
        prefetch_l2(a.vec + (int64_t)l * a.d + (lo + threadIdx.x) * 4);
    pdl_wait();
    __shared__ float slab[2][32];
    __shared__ float met[2][kFinishCluster];   // CTA 0's: every CTA's (dot, sum of squares)
    const float2 ae = alpha_eff_of(a, l);
    const float2 s = finish_columns(a, l, lo, hi, ae.x, slab);
    if (threadIdx.x == 0) {
        float* dst = cluster.map_shared_rank(&met[0][0], 0);
        dst[c] = s.x;
        dst[kFinishCluster + c] = s.y;
    }
    cluster.sync();
    if (c == 0 && threadIdx.x == 0) {
        float sd = 0.f, sq = 0.f;
        for (int k = 0; k < kFinishCluster; ++k) {
            sd += met[0][k];
            sq += met[1][k];
        }
        finish_layer(a, l, sd, sq, ae.y);
    }
}

// out[l][c] (+)= sum_p rows[l][p][c]: one thread per (layer, 4 columns), the rows are L2-resident
// (the backward launches of this step wrote them)
__global__ void __launch_bounds__(256)
reduce_rows_kernel(float* __restrict__ rows, float* __restrict__ out, int n_rows,
                   int64_t layer_stride, int d, int64_t total4, int accumulate, int clear) {
    pdl_launch_dependents();
    pdl_wait();
    const int d4 = d / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = i / d4;
        const int c = (int)(i - l * d4);
        float4* src = reinterpret_cast<float4*>(rows + l * layer_stride) + c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int p = 0;
        for (; p + 8 <= n_rows; p += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (int64_t)(p + u) * d4);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
            }
        }
        for (; p < n_rows; ++p) {
            const float4 v = __ldcg(src + (int64_t)p * d4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (clear) {   // the replicas are ready for the next backward pass
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = 0; q < n_rows; ++q) src[(int64_t)q * d4] = z;
        }
        float4* dst = reinterpret_cast<float4*>(out + l * d) + c;
        if (accumulate) {
            const float4 o = *dst;
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        *dst = acc;
    }
}

__global__ void get_mask_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ len,
                                int64_t pad, int batch, int seq, uint8_t* __restrict__ mask) {
    pdl_launch_dependents();
    pdl_wait();
    const int n = batch * seq;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = i / seq, t = i - b * seq;
        mask[i] = (uint8_t)(((int64_t)t >= len[b]) && (ids[i] != pad));
    }
}

// exclusive rank of this thread's flag among the CTA's flags; `base` carries across tiles
__device__ __forceinline__ int cta_rank(bool flag, int* warp_tot, int* tile_total) {
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
        const int c = warp_tot[w];
        if (w < warp) off += c;
        tot += c;
    }
    *tile_total = tot;
    return off + __popc(bal & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(1024)
kd_prepare_rows_kernel(const int64_t* __restrict__ s_ids, const int64_t* __restrict__ s_len,
                       const int64_t* __restrict__ s_att, const int64_t* __restrict__ t_ids,
                       const int64_t* __restrict__ t_len, int64_t pad, int64_t image_tok,
                       int ce_variant, int batch, int Tq, int Tt, int32_t* __restrict__ kl_tea_row,
                       int64_t* __restrict__ ce_label, int32_t* __restrict__ counts,
                       int32_t* __restrict__ tea_sel) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ int32_t stu_by_rank[];  // [batch*Tq]: flat student row of the k-th KL row
    __shared__ int warp_tot[32];
    __shared__ int s_m;
    const int ns = batch * Tq, nt = batch * Tt;
    if (threadIdx.x == 0) s_m = 0;

    // The launch is a chain of global round trips, not work: everything this thread needs for its
    // FIRST student row, teacher row and label is requested here, together (one round trip instead
    // of three behind each other's barriers); rows past the first blockDim.x of a list take the loops.
    const int i_first = threadIdx.x;
    int64_t f_sid = pad, f_slen = 0, f_tid = pad, f_tlen = 0, f_next = -100, f_att = 1;
    if (i_first < ns) {
        f_sid = s_ids[i_first];
        f_slen = s_len[i_first / Tq];
        if (ce_label != nullptr && (i_first % Tq) + 1 < Tq) {
            f_next = s_ids[i_first + 1];
            if (s_att != nullptr) f_att = s_att[i_first + 1];
        }
    }
    if (i_first < nt) {
        f_tid = t_ids[i_first];
        f_tlen = t_len[i_first / Tt];
    }

    // student rows in row-major order -> rank
    int base = 0;
    for (int i0 = 0; i0 < ns; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        bool sel = false;
        if (i < ns) {
            const int b = i / Tq, t = i - b * Tq;
            sel = i0 == 0 ? (((int64_t)t >= f_slen) && (f_sid != pad))
                          : (((int64_t)t >= s_len[b]) && (s_ids[i] != pad));
            kl_tea_row[i] = -1;
        }
        int tot;
        const int rk = cta_rank(sel, warp_tot, &tot);
        if (sel) stu_by_rank[base + rk] = i;
        base += tot;
    }
    const int n_stu = base;
    __syncthreads();

    // teacher rows in row-major order: the k-th one pairs with the k-th student row
    base = 0;
    for (int i0 = 0; i0 < nt; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        bool sel = false;
        if (i < nt) {
            const int b = i / Tt, t = i - b * Tt;
            sel = i0 == 0 ? (((int64_t)t >= f_tlen) && (f_tid != pad))
                          : (((int64_t)t >= t_len[b]) && (t_ids[i] != pad));
        }
        int tot;
        const int rk = cta_rank(sel, warp_tot, &tot);
        if (sel && base + rk < n_stu) {
            // compact form: the student row points at the RANK of its teacher row, and tea_sel
            // lists the selected teacher rows in that order (for a gather before lm_head)
            kl_tea_row[stu_by_rank[base + rk]] = tea_sel ? base + rk : i;
            if (tea_sel) tea_sel[base + rk] = i;
        }
        base += tot;
    }
    const int n_tea = base;
    if (tea_sel) {   // entries past the last pair: a valid row, so that a fixed-size gather is safe
        for (int i = (n_tea < n_stu ? n_tea : n_stu) + threadIdx.x; i < ns; i += blockDim.x) tea_sel[i] = 0;
    }

    // next-token labels for labels = input_ids
    int m_local = 0;
    if (ce_label != nullptr) {
        for (int i = threadIdx.x; i < ns; i += blockDim.x) {
            const int b = i / Tq, t = i - b * Tq;
            int64_t lab = -100;
            if (t + 1 < Tq) {
                const bool first = i == i_first;
                lab = first ? f_next : s_ids[i + 1];
                if (ce_variant != 2) {
                    if (s_att != nullptr && (first ? f_att : s_att[i + 1]) == 0) lab = -100;
                    if (ce_variant == 1 && lab == image_tok) lab = -100;
                }
            }
            ce_label[i] = lab;
            m_local += (lab != -100);
        }
        m_local = (int)warp_sum((float)m_local);  // exact: counts stay far below 2^24
        if ((threadIdx.x & 31) == 0 && m_local) atomicAdd(&s_m, m_local);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        counts[0] = n_stu;
        counts[1] = s_m;
        counts[2] = n_tea;
        counts[3] = 0;
    }
}

// ---- optimizer ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, int64_t n, float prescale, float* __restrict__ acc) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float slab[8];
    float s = 0.f;
    // 128-bit loads where the buffer allows: one round trip per thread for the 131 104 floats of
    // the ICV parameters instead of four dependent ones (the launch is latency-bound)
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = (reinterpret_cast<uintptr_t>(g) & 15u) == 0 ? (n >> 2) : 0;
    for (int64_t i = gtid; i < n4; i += gstride) {
        const float4 x = reinterpret_cast<const float4*>(g)[i];
        const float a0 = x.x * prescale, a1 = x.y * prescale, a2 = x.z * prescale, a3 = x.w * prescale;
        s = fmaf(a0, a0, s);
        s = fmaf(a1, a1, s);
        s = fmaf(a2, a2, s);
        s = fmaf(a3, a3, s);
    }
    for (int64_t i = n4 * 4 + gtid; i < n; i += gstride) {
        const float x = g[i] * prescale;
        s = fmaf(x, x, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) slab[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += slab[w];
        atomicAdd(acc, t);
    }
}

struct AdamArgs {
    float* p; const float* g; float* m; float* v;
    int64_t n_vec, n_alpha;
    float lr_vec, lr_alpha, beta1, beta2, eps, wd, prescale, max_norm;
    float bc1, bc2_sqrt;   // 1 - beta1^t, sqrt(1 - beta2^t)
    float* norm_out;
    float* acc;            // workspace[0]: sum of squares
    unsigned* ticket;      // workspace[1]
    const float* partials; // or: n_partials sums of squares to be added in index order
    int n_partials;
    const unsigned* skip;  // if set and non-zero: the gradient is not trustworthy, leave everything as it is
    int vec4;              // 128-bit accesses allowed (alignment, n_vec % 4 == 0)
};

__global__ void __launch_bounds__(256) adamw_kernel(AdamArgs a) {
    pdl_launch_dependents();
    // parameters and moments are last step's: hint them into L2 while the norm is still being summed
    if (a.vec4 && (threadIdx.x & 7) == 0) {
        const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i4 * 4 < a.n_vec + a.n_alpha) {
            prefetch_l2(a.p + i4 * 4);
            prefetch_l2(a.m + i4 * 4);
            prefetch_l2(a.v + i4 * 4);
        }
    }
    pdl_wait();
    // a failed gradient exchange (a peer never delivered) must not move the parameters
    if (a.skip != nullptr && *reinterpret_cast<const volatile unsigned*>(a.skip) != 0u) return;
    // this thread's first vector of gradient / parameter / moments is requested BEFORE the norm is
    // summed: the launch is a chain of L2 round trips, and these four do not depend on the norm
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = a.vec4 ? ((a.n_vec + a.n_alpha) >> 2) : 0;
    const bool first = gtid < n4;
    float4 g4f = make_float4(0.f, 0.f, 0.f, 0.f), p4f = g4f, m4f = g4f, v4f = g4f;
    if (first) {
        g4f = reinterpret_cast<const float4*>(a.g)[gtid];
        p4f = reinterpret_cast<float4*>(a.p)[gtid];
        m4f = reinterpret_cast<float4*>(a.m)[gtid];
        v4f = reinterpret_cast<float4*>(a.v)[gtid];
    }
    // every CTA reads the norm before taking a ticket; the last ticket holder clears the workspace
    float coef = a.prescale;
    float sumsq;
    if (a.partials != nullptr) {
        // fixed-order sum: the same bits in every CTA and on every rank.  A fixed TREE, not one
        // thread walking the ~260 partials (each load waited for the one before: ~15 us of every
        // multi-GPU step): thread t takes partials t, t + 256, ..., then the shuffle butterfly and
        // the eight warp sums in index order - the same association everywhere
        __shared__ float s_part[8];
        float t = 0.f;
        for (int i = threadIdx.x; i < a.n_partials; i += 256) t += a.partials[i];
        t = warp_sum(t);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = t;
        __syncthreads();
        sumsq = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sumsq += s_part[w];
    } else {
        sumsq = *reinterpret_cast<volatile float*>(a.acc);
    }
    const float norm = sqrtf(sumsq);
    if (a.max_norm > 0.f) {
        const float c = a.max_norm / (norm + 1e-6f);   // torch.nn.utils.clip_grad_norm_
        coef *= fminf(c, 1.0f);
    }
    const int64_t n = a.n_vec + a.n_alpha;
    auto update = [&](float lr, float graw, float& p, float& m, float& v) {
        const float g = graw * coef;
        p = p * (1.0f - lr * a.wd);
        m = a.beta1 * m + (1.0f - a.beta1) * g;
        v = a.beta2 * v + (1.0f - a.beta2) * g * g;
        const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
        p -= (lr / a.bc1) * (m / denom);
    };
    // 128-bit accesses (a.vec4: 16-byte aligned buffers, n_vec a multiple of 4 so that no vector
    // straddles the two learning rates): one round trip per thread instead of four dependent ones
    if (first) {
        const float lr = gtid * 4 < a.n_vec ? a.lr_vec : a.lr_alpha;
        update(lr, g4f.x, p4f.x, m4f.x, v4f.x);
        update(lr, g4f.y, p4f.y, m4f.y, v4f.y);
        update(lr, g4f.z, p4f.z, m4f.z, v4f.z);
        update(lr, g4f.w, p4f.w, m4f.w, v4f.w);
        reinterpret_cast<float4*>(a.p)[gtid] = p4f;
        reinterpret_cast<float4*>(a.m)[gtid] = m4f;
        reinterpret_cast<float4*>(a.v)[gtid] = v4f;
    }
    for (int64_t i4 = gtid + gstride; i4 < n4; i4 += gstride) {
        const float lr = i4 * 4 < a.n_vec ? a.lr_vec : a.lr_alpha;
        const float4 g4 = reinterpret_cast<const float4*>(a.g)[i4];
        float4 p4 = reinterpret_cast<float4*>(a.p)[i4];
        float4 m4 = reinterpret_cast<float4*>(a.m)[i4];
        float4 v4 = reinterpret_cast<float4*>(a.v)[i4];
        update(lr, g4.x, p4.x, m4.x, v4.x);
        update(lr, g4.y, p4.y, m4.y, v4.y);
        update(lr, g4.z, p4.z, m4.z, v4.z);
        update(lr, g4.w, p4.w, m4.w, v4.w);
        reinterpret_cast<float4*>(a.p)[i4] = p4;
        reinterpret_cast<float4*>(a.m)[i4] = m4;
        reinterpret_cast<float4*>(a.v)[i4] = v4;
    }
    for (int64_t i = n4 * 4 + gtid; i < n; i += gstride) {
        const float lr = i < a.n_vec ? a.lr_vec : a.lr_alpha;
        float p = a.p[i], m = a.m[i], v = a.v[i];
        update(lr, a.g[i], p, m, v);
        a.p[i] = p; a.m[i] = m; a.v[i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0 && a.norm_out) *a.norm_out = norm;
        __threadfence();
        if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) {
            *a.acc = 0.f;
            *a.ticket = 0u;
        }
    }
}

}  // namespace
}  // namespace licv

using namespace licv;

extern "C" const char* licv_status_string(int status) {
    switch (status) {
        case LICV_OK: return "ok";
        case LICV_ERR_NULL_POINTER: return "null pointer";
        case LICV_ERR_BAD_DTYPE: return "unsupported dtype (bf16, fp16, fp32 only)";
        case LICV_ERR_BAD_DIM:
            return "unsupported hidden size (rows must be whole 16-byte vectors, at most 32 KB)";
        case LICV_ERR_MISALIGNED: return "pointer not 16-byte aligned";
        case LICV_ERR_BAD_ARGUMENT: return "bad argument";
        case LICV_ERR_WORKSPACE: return "workspace too small";
        case LICV_ERR_NO_DEVICE: return "no sm_100 (B200) device: liblicv_b200 has no other path";
        default: break;
    }
    if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
    return "unknown status";
}

extern "C" int licv_abi_version(void) { return LICV_ABI_VERSION; }

extern "C" int licv_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    const DeviceInfo& d = device_info();
    if (sm_count) *sm_count = d.sm_count;
    if (cc_major) *cc_major = d.cc_major;
    if (cc_minor) *cc_minor = d.cc_minor;
    return d.status;
}

extern "C" int licv_icv_scale(const float* alpha_raw, const float* vec, float* icv, int n_layers,
                              int d, int use_sigmoid, licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n_layers < 0 || d <= 0 || d % 4 != 0) return LICV_ERR_BAD_DIM;
    if (n_layers == 0) return LICV_OK;
    if (!alpha_raw || !vec || !icv) return LICV_ERR_NULL_POINTER;
    if (!aligned16(vec) || !aligned16(icv)) return LICV_ERR_MISALIGNED;
    const int64_t n4 = (int64_t)n_layers * d / 4;
    const int grid = (int)((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184);
    return launch_pdl(icv_scale_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                      alpha_raw, vec, icv, n_layers, d, use_sigmoid);
}

extern "C" int licv_icv_scale_bwd(const float* alpha_raw, const float* vec, const float* d_icv,
                                  float* d_vec, float* d_alpha_raw, int n_layers, int d,
                                  int use_sigmoid, licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n_layers < 0 || d <= 0 || d % 4 != 0) return LICV_ERR_BAD_DIM;
    if (n_layers == 0) return LICV_OK;
    if (!alpha_raw || !vec || !d_icv || !d_vec) return LICV_ERR_NULL_POINTER;
    if (!aligned16(vec) || !aligned16(d_icv) || !aligned16(d_vec)) return LICV_ERR_MISALIGNED;
    // one 128-bit vector per thread where the row allows it (one round trip instead of four)
    int threads = 256;
    while (threads < 1024 && threads * 4 < d) threads *= 2;
    return launch_pdl(icv_scale_bwd_kernel, dim3(n_layers), dim3(threads), 0,
                      reinterpret_cast<cudaStream_t>(stream), alpha_raw, vec, d_icv, d_vec, d_alpha_raw,
                      d, use_sigmoid);
}

extern "C" int licv_reduce_rows(float* rows, float* out, int n_layers, int n_rows,
                                int64_t layer_stride, int d, int accumulate, int clear,
                                licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n_layers < 0 || n_rows < 0 || d <= 0 || d % 4 != 0 || layer_stride % 4 != 0 ||
        layer_stride < (int64_t)n_rows * d)
        return LICV_ERR_BAD_DIM;
    if (n_layers == 0) return LICV_OK;
    if (!rows || !out) return LICV_ERR_NULL_POINTER;
    if (!aligned16(rows) || !aligned16(out)) return LICV_ERR_MISALIGNED;
    const int64_t total4 = (int64_t)n_layers * (d / 4);
    int64_t grid = (total4 + 255) / 256;
    const int64_t cap = (int64_t)device_info().sm_count * 8;
    if (grid > cap) grid = cap;
    return launch_pdl(reduce_rows_kernel, dim3((unsigned)grid), dim3(256), 0,
                      reinterpret_cast<cudaStream_t>(stream), rows, out, n_rows, layer_stride, d, total4,
                      accumulate, clear);
}

extern "C" int licv_icv_grad_finish(float* rows, int n_rows, int64_t layer_stride,
                                    const float* alpha_raw, const float* vec, float* d_icv,
                                    float* d_vec, float* d_alpha_raw, float* norm_partials,
                                    float grad_prescale, int n_layers, int d, int use_sigmoid,
                                    int accumulate, int clear, licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n_layers < 0 || n_rows < 0 || d <= 0 || d % 4 != 0 || layer_stride % 4 != 0 ||
        layer_stride < (int64_t)n_rows * d)
        return LICV_ERR_BAD_DIM;
    if (n_layers == 0) return LICV_OK;
    if (!alpha_raw || !vec || !d_vec || (n_rows > 0 && !rows)) return LICV_ERR_NULL_POINTER;
    if (!aligned16(rows) || !aligned16(vec) || !aligned16(d_vec) || !aligned16(d_icv))
        return LICV_ERR_MISALIGNED;
    FinishArgs a;
    a.rows = rows; a.n_rows = n_rows; a.layer_stride = layer_stride;
    a.alpha = alpha_raw; a.vec = vec;
    a.d_icv = d_icv; a.d_vec = d_vec; a.d_alpha = d_alpha_raw; a.partial = norm_partials;
    a.prescale = grad_prescale; a.d = d; a.use_sigmoid = use_sigmoid;
    a.accumulate = accumulate; a.clear = clear;
    static const bool cluster_form = [] {
        const char* v = std::getenv("LICV_FINISH_CLUSTER");   // 0: one CTA per layer at any width (A/B)
        return !(v && v[0] == '0');
    }();
    if (d >= 2048 && cluster_form) {   // wide layers: a cluster of CTAs per layer
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_layers * kFinishCluster);
        cfg.blockDim = dim3(256);
        cfg.stream = reinterpret_cast<cudaStream_t>(stream);
        cudaLaunchAttribute attr[2];
        cfg.attrs = attr;
        cfg.numAttrs = launch_attrs(attr, kFinishCluster);
        return (int)cudaLaunchKernelEx(&cfg, icv_grad_finish_cluster_kernel, a);
    }
    int threads = 256;
    while (threads < 1024 && threads * 4 < d) threads *= 2;
    return launch_pdl(icv_grad_finish_kernel, dim3(n_layers), dim3(threads), 0,
                      reinterpret_cast<cudaStream_t>(stream), a);
}

extern "C" int licv_get_mask(const int64_t* input_ids, const int64_t* mask_length,
                             int64_t pad_token_id, int batch, int seq_len, uint8_t* mask,
                             licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (batch < 0 || seq_len < 0) return LICV_ERR_BAD_ARGUMENT;
    if (batch == 0 || seq_len == 0) return LICV_OK;
    if (!input_ids || !mask_length || !mask) return LICV_ERR_NULL_POINTER;
    const int n = batch * seq_len;
    const int grid = (n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184;
    return launch_pdl(get_mask_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                      input_ids, mask_length, pad_token_id, batch, seq_len, mask);
}

extern "C" int licv_kd_select_rows(const int64_t* stu_ids, const int64_t* stu_mask_length,
                                   const int64_t* stu_attention_mask, const int64_t* tea_ids,
                                   const int64_t* tea_mask_length, int64_t pad_token_id,
                                   int64_t image_token_id, int ce_variant, int batch, int stu_len,
                                   int tea_len, int32_t* kl_tea_row, int64_t* ce_label,
                                   int32_t* counts, int32_t* tea_sel, licv_stream_t stream);

extern "C" int licv_kd_prepare_rows(const int64_t* stu_ids, const int64_t* stu_mask_length,
                                    const int64_t* stu_attention_mask, const int64_t* tea_ids,
                                    const int64_t* tea_mask_length, int64_t pad_token_id,
                                    int64_t image_token_id, int ce_variant, int batch, int stu_len,
                                    int tea_len, int32_t* kl_tea_row, int64_t* ce_label,
                                    int32_t* counts, licv_stream_t stream) {
    return licv_kd_select_rows(stu_ids, stu_mask_length, stu_attention_mask, tea_ids,
                               tea_mask_length, pad_token_id, image_token_id, ce_variant, batch,
                               stu_len, tea_len, kl_tea_row, ce_label, counts, nullptr, stream);
}

extern "C" int licv_kd_select_rows(const int64_t* stu_ids, const int64_t* stu_mask_length,
                                   const int64_t* stu_attention_mask, const int64_t* tea_ids,
                                   const int64_t* tea_mask_length, int64_t pad_token_id,
                                   int64_t image_token_id, int ce_variant, int batch, int stu_len,
                                   int tea_len, int32_t* kl_tea_row, int64_t* ce_label,
                                   int32_t* counts, int32_t* tea_sel, licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (batch <= 0 || stu_len <= 0 || tea_len <= 0 || ce_variant < 0 || ce_variant > 2)
        return LICV_ERR_BAD_ARGUMENT;
    if (!stu_ids || !stu_mask_length || !tea_ids || !tea_mask_length || !kl_tea_row || !counts)
        return LICV_ERR_NULL_POINTER;
    const size_t smem = (size_t)batch * stu_len * sizeof(int32_t);
    if (smem > 200 * 1024) return LICV_ERR_BAD_ARGUMENT;  // > 51200 student positions
    if (smem > 48 * 1024) {
        static std::once_flag once;
        std::call_once(once, [] {
            cudaFuncSetAttribute(kd_prepare_rows_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        });
    }
    return launch_pdl(kd_prepare_rows_kernel, dim3(1), dim3(1024), smem,
                      reinterpret_cast<cudaStream_t>(stream), stu_ids, stu_mask_length, stu_attention_mask, tea_ids, tea_mask_length, pad_token_id, image_token_id, ce_variant, batch, stu_len, tea_len, kl_tea_row, ce_label, counts, tea_sel);
}

namespace licv {
// AdamW kernel alone; the squared gradient norm (of grad * grad_prescale) is already in workspace[0]
int launch_adamw_after_norm(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                            int64_t n_vec, int64_t n_alpha, float lr_vec, float lr_alpha, float beta1,
                            float beta2, float eps, float weight_decay, int64_t step,
                            float grad_prescale, float max_grad_norm, float* norm_out,
                            void* workspace, const float* norm_partials, int n_partials,
                            const unsigned* skip_if_set, cudaStream_t st) {
    const int64_t n = n_vec + n_alpha;
    AdamArgs a;
    a.partials = norm_partials;
    a.n_partials = n_partials;
    a.skip = skip_if_set;
    a.p = param; a.g = grad; a.m = exp_avg; a.v = exp_avg_sq;
    a.n_vec = n_vec; a.n_alpha = n_alpha;
    a.lr_vec = lr_vec; a.lr_alpha = lr_alpha; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    a.wd = weight_decay; a.prescale = grad_prescale; a.max_norm = max_grad_norm;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.norm_out = norm_out;
    a.acc = static_cast<float*>(workspace);
    a.ticket = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + 4);
    a.vec4 = (n_vec % 4 == 0) &&
             ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
               reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u) == 0;
    const int grid = (int)((n + 1023) / 1024 < 148 ? (n + 1023) / 1024 : 148);
    return launch_pdl(adamw_kernel, dim3(grid), dim3(256), 0, st, a);
}
}  // namespace licv

extern "C" int licv_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                               int64_t n_vec, int64_t n_alpha, float lr_vec, float lr_alpha,
                               float beta1, float beta2, float eps, float weight_decay, int64_t step,
                               float grad_prescale, float max_grad_norm, float* norm_out,
                               void* workspace, licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n_vec < 0 || n_alpha < 0 || step < 1) return LICV_ERR_BAD_ARGUMENT;
    const int64_t n = n_vec + n_alpha;
    if (n == 0) return LICV_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !workspace) return LICV_ERR_NULL_POINTER;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = (int)((n + 1023) / 1024 < 148 ? (n + 1023) / 1024 : 148);
    if (int rc = launch_pdl(sumsq_kernel, dim3(grid), dim3(256), 0, st, grad, n, grad_prescale,
                            static_cast<float*>(workspace)))
        return rc;
    return launch_adamw_after_norm(param, grad, exp_avg, exp_avg_sq, n_vec, n_alpha, lr_vec, lr_alpha,
                                   beta1, beta2, eps, weight_decay, step, grad_prescale,
                                   max_grad_norm, norm_out, workspace, nullptr, 0, nullptr, st);
}

extern "C" int licv_adamw_step_partials(float* param, const float* grad, float* exp_avg,
                                        float* exp_avg_sq, int64_t n_vec, int64_t n_alpha,
                                        float lr_vec, float lr_alpha, float beta1, float beta2,
                                        float eps, float weight_decay, int64_t step,
                                        float grad_prescale, float max_grad_norm, float* norm_out,
                                        void* workspace, const float* norm_partials, int n_partials,
                                        licv_stream_t stream) {
    if (device_info().status != LICV_OK) return device_info().status;
    if (n_vec < 0 || n_alpha < 0 || step < 1 || n_partials < 1) return LICV_ERR_BAD_ARGUMENT;
    if (n_vec + n_alpha == 0) return LICV_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !workspace || !norm_partials)
        return LICV_ERR_NULL_POINTER;
    return launch_adamw_after_norm(param, grad, exp_avg, exp_avg_sq, n_vec, n_alpha, lr_vec, lr_alpha,
                                   beta1, beta2, eps, weight_decay, step, grad_prescale,
                                   max_grad_norm, norm_out, workspace, norm_partials, n_partials,
                                   nullptr, reinterpret_cast<cudaStream_t>(stream));
}
