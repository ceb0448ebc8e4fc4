// L-ICV distillation loss, cluster kernel: ONE pass over HBM, every transcendental evaluated once.
//
// Same arithmetic as the generic kernel in licv_kd_loss.cu (reference icv_src/icv_module.py:
// 121-134 for the KL, the HF-internal shifted CE consumed at :94-98,115-117, the combine at
// :100-101,107-119); what changes is where a row lives between its three sweeps.
//
// The kl_eps inside the logarithms makes d(stu) depend on W_n = sum_v p q/(q+eps), so a row needs
// (1) its softmax partition sums, (2) the KL value and W_n, (3) the gradient: three sweeps with
// two row-wide reductions between them.  The generic kernel re-reads the 16-bit logits from L2
// for every sweep and re-evaluates both exponentials each time (~9 MUFU per element: the SFU
// pipe, not HBM, bounds it at ~20 % of the roofline).  Here a row pair is spread over a
// thread-block CLUSTER (C CTAs x NT threads, C x NV x NT 16-byte vectors >= one row; V = 32002
// 16-bit: 8 CTAs x 128 threads x 4 vectors, six CTAs of different rows per SM):
//
//   sweep B  registers (raw logits, prefetched during the previous row) -> e = 2^(u - m_thread)
//            for student and teacher, written as fp32 to a thread-private shared-memory cache
//            (each thread only ever reads back its own slots: no barrier, no bank conflicts);
//            the softmax sums are carried as online-softmax (max, sum) pairs, so ONE reduction
//            yields max and partition sum of both rows;
//   sweep C  cache -> p, q, the KL terms and w = p q/(q+eps); kl_w * w overwrites the teacher
//            slot;  second reduction: (KL_n, W_n);
//   sweep D  cache -> gradient = e_s A - kl_w w - ce_w [j = label], packed and stored.
//
// 4 MUFU per element (2 ex2, 1 rcp, 1 lg2) instead of 9, HBM traffic = the algorithmic 3 e V
// bytes per KL row (2 e V for a CE-only row, e V zero-fill for a row in neither loss).  The two
// reductions cross the cluster through distributed shared memory: every warp sends its 16-byte
// partial to every CTA with st.async, counted on the receiver's mbarrier (no CTA barrier, no
// cluster barrier, no release fence on the path).  CTAs of different clusters share an SM, so one
// row's reductions and loads overlap another's arithmetic; the next row's logits are requested
// while the first reduction's partials travel.
//
// Rows may start on any element boundary (V = 32002 / 32003): the student row is walked in
// 16-byte-aligned vectors, partial first/last vectors go element by element, the teacher row is
// read with the widest loads its relative phase allows.
//
// Not a dense contraction: no tensor cores; bounds are HBM, then the SFU and FMA pipes.
#include <cstdlib>
#include <type_traits>

#include "licv_common.cuh"
#include "licv_kd_loss.cuh"
#include "licv_kd_rows.cuh"

namespace licv {
namespace {


struct Exchange {
    float4* slots;      // [C * W] partials in THIS CTA's shared memory (W = warps per CTA)
    uint64_t* bar;      // completes when C * W * 16 bytes have arrived
    uint32_t parity;
};

// this warp's partial -> slot (rank, warp) of every CTA of the cluster
template <int W>
__device__ __forceinline__ void exchange_send(const Exchange& x, float4 v, uint32_t rank, uint32_t C) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) mbar_expect_tx(x.bar, C * W * 16u);
    if ((uint32_t)lane < C) st_async_f4(x.slots + rank * W + warp, x.bar, (uint32_t)lane, v);
}
template <int W>
__device__ __forceinline__ MZ2 exchange_recv_mz(Exchange& x, uint32_t C) {
    const int lane = threadIdx.x & 31;
    mbar_wait(x.bar, x.parity);
    x.parity ^= 1u;
    const int n = (int)C * W;
    MZ2 v{kNoMaxI, 0.f, kNoMaxI, 0.f};
    if (lane < n) {
        const float4 p = x.slots[lane];
        v = MZ2{__float_as_int(p.x), p.y, __float_as_int(p.z), p.w};
    }
    if (lane + 32 < n) {
        const float4 p = x.slots[lane + 32];
        v = mz_join(v, MZ2{__float_as_int(p.x), p.y, __float_as_int(p.z), p.w});
    }
    return mz_warp(v);
}
template <int W>
__device__ __forceinline__ float2 exchange_recv_sum(Exchange& x, uint32_t C) {
    const int lane = threadIdx.x & 31;
    mbar_wait(x.bar, x.parity);
    x.parity ^= 1u;
    const int n = (int)C * W;
    float2 v = make_float2(0.f, 0.f);
    if (lane < n) {
        const float4 p = x.slots[lane];
        v = make_float2(p.x, p.y);
    }
    if (lane + 32 < n) {
        const float4 p = x.slots[lane + 32];
        v.x += p.x;
        v.y += p.y;
    }
    v.x = warp_sum(v.x);
    v.y = warp_sum(v.y);
    return v;
}

#ifdef LICV_TRACE
// debug build only: phase timestamps (clock64) of one thread of the first CTAs, 8 per row
__device__ long long g_trace[64 * 64 * 8];
#define LICV_TP(slot)                                                                      \
    do {                                                                                   \
        if (tid == LICV_TRACE_TID && blockIdx.x < 64 && trace_row < 64)                    \
            g_trace[(blockIdx.x * 64 + trace_row) * 8 + (slot)] = clock64();               \
    } while (0)
#ifndef LICV_TRACE_TID
#define LICV_TRACE_TID 0
#endif
#else
#define LICV_TP(slot) do { } while (0)
#endif


template <int DT, int NV, int NT>
__global__ void __launch_bounds__(NT, (NV <= 4 ? 768 / NT : 1)) kd_loss_cluster_kernel(KdArgs a) {
    constexpr int kT = NT;
    constexpr int kWarps = NT / 32;
    constexpr int EPV = Fmt<DT>::kPerVec;
    constexpr int EB = Fmt<DT>::kBytes;
    constexpr int Q = EPV / 4;                      // float4 slots per vector
    constexpr int kStep = kT * EPV;                 // elements between a thread's vectors
    extern __shared__ __align__(16) float4 cache[];  // [2][NV * Q][kT]: e_s then e_t / kl_w * w
    __shared__ __align__(16) float4 slots[3][8 * kMaxWarps];   // reduction 1 (two row parities), reduction 2
    __shared__ __align__(8) uint64_t xbar[3];   // reduction 1 (two row parities), reduction 2
    __shared__ float s_tot[2 * kWarps];
    __shared__ int s_last;

    float4* const cs = cache;
    float4* const ct = cache + NV * Q * kT;
    const int tid = threadIdx.x;
    const uint32_t C = cluster_num_ctas();
    const uint32_t rank = cluster_cta_rank();
    const int64_t n_clusters = gridDim.x / C;
    const int64_t cluster_id = blockIdx.x / C;
    if (tid == 0) {
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        mbar_init(&xbar[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_launch_dependents();
    cluster_sync_all();   // every peer's barriers exist before the first message
    pdl_wait();           // the previous kernel's logits / row lists are visible from here on
    // Reduction 1 has two (slots, mbarrier) sets used alternately by row parity: on CE-only rows
    // nothing separates one row's reduction 1 from the next row's, so a fast peer may send its
    // next partial while this CTA has not yet received (or a slow warp here still reads) all the
    // current ones; it cannot get two rows ahead, because it needs every warp's next partial
    // first, and a warp sends that only after it is done with the current set.
    Exchange ex1a{slots[0], &xbar[0], 0u}, ex1b{slots[1], &xbar[1], 0u}, ex2x{slots[2], &xbar[2], 0u};
    uint32_t row_par = 0;

    const float T = a.temperature;
    const float inv_t = 1.0f / T;
    const bool round_tempered = (a.round_flags & LICV_ROUND_TEMPERED) && DT != LICV_F32 && T != 1.0f;
    const int64_t n_kl = a.counts ? (int64_t)a.counts[0] : a.n_kl;
    const int64_t n_ce = a.counts ? (int64_t)a.counts[1] : a.n_ce;
    const bool use_kl = !a.only_hard_loss;
    const bool use_ce = a.ce_label != nullptr;
    const float kl_w = use_kl ? a.grad_scale * T / (float)n_kl : 0.f;
    const float ce_w = use_ce ? a.grad_scale * (a.only_hard_loss ? 1.0f : a.hard_loss_weight) /
                                    (float)n_ce
                              : 0.f;
    const float eps = a.kl_eps;
    float* row_kl = a.row_loss;
    float* row_ce = a.row_loss + a.n_rows;
    const int V = a.vocab;

    // A row is described by (r, tr, lab): teacher row (-1 = no KL) and label (kLabNone = no CE);
    // everything else is recomputed where it is needed instead of being carried in registers.
    auto fetch_tr = [&](int64_t r) -> int {
        if (!use_kl || r >= a.n_rows) return -1;
        return a.kl_tea_row ? a.kl_tea_row[r] : (int)r;
    };
    auto fetch_lab = [&](int64_t r) -> int64_t {     // raw: converted where it is first needed
        if (!use_ce || r >= a.n_rows) return kLabNone;
        return a.ce_label[r];
    };
    auto narrow_lab = [](int64_t l) -> int {
        return (l < -100 || l > 0x7fffffff) ? kLabBad : (int)l;
    };
    auto x_row = [&](int64_t r) { return static_cast<const char*>(a.stu) + (size_t)r * a.stu_stride * EB; };
    auto t_row = [&](int tr) { return static_cast<const char*>(a.tea) + (size_t)tr * a.tea_stride * EB; };
    auto g_row = [&](int64_t r) { return static_cast<char*>(a.dstu) + (size_t)r * a.stu_stride * EB; };
    auto phase16 = [](const void* p) { return (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u); };
    // first element of this thread's first vector of row r (the others follow every kStep)
    auto j_first = [&](const char* xr) -> int {
        return (((int)rank * NV) * kT + tid) * EPV - (int)(phase16(xr) / EB);
    };
    // every vector of every lane of this warp lies wholly inside the row: no per-vector checks.
    // Decided per warp, like the loads: a lane that took the slow store path on its own made its
    // warp run both paths for all NV vectors and arrive late at the row's next barrier.
    auto all_full = [&](int j0) -> bool { return warp_inside_row<DT>(j0, kStep, NV, V, tid & 31); };

    uint4 xs[NV], xt[NV];
    auto load_raw = [&](int64_t r, int tr, int lab) {
        if (tr < 0 && lab == kLabNone) return;
        const char* xr = x_row(r);
        const int j0 = j_first(xr);
        load_row_vecs<DT, NV>(xs, xr, j0, kStep, V, 16, tid & 31);
        if (tr >= 0) {
            const char* tp = t_row(tr);
            const uint32_t dph = (phase16(tp) - phase16(xr)) & 15u;
            load_row_vecs<DT, NV>(xt, tp, j0, kStep, V, dph == 0 ? 16 : (int)(dph & (0u - dph)), tid & 31);
        }
    };
    auto store_grad = [&](char* gr_row, bool g_vec, int j0, bool full, int k, const float* gr) {
        const int jwk = j0 - (tid & 31) * EPV + k * kStep;          // lane 0's element of vector k
        if (g_vec && (full || (jwk >= 0 && jwk + 32 * EPV <= V)))
            st_vec(reinterpret_cast<uint4*>(gr_row + ((int64_t)j0 + (int64_t)k * kStep) * EB), pack<DT>(gr));
        else
            store_row_vec<DT>(gr_row, j0 + k * kStep, V, g_vec, gr);
    };
    // the thread's integer maximum of u = logit * c (>= every u, so every e below is <= 1)
    auto thread_max = [&](const uint4 (&raw)[NV], float it_row, bool rnd) -> int {
        RawMax<DT> mx;
#pragma unroll
        for (int k = 0; k < NV; ++k) mx.add(raw[k]);
        float m = mx.get() * it_row;                           // monotone: max, then temper
        if (rnd) m = Fmt<DT>::round(m);
        m = fminf(fmaxf(ceilf(m * kLog2e), -1.0e6f), 1.0e6f);  // -inf (empty) -> -1e6
        return (int)m;
    };
    // sweep B for one row: e = 2^(u - m) into the cache, returns the thread's sum
    auto sweep_b = [&](const uint4 (&raw)[NV], float m, float c_row, float it_row, bool rnd,
                       float4* dst) -> float {
        float z = 0.f;
        if (rnd) {   // the tempered logit is stored in the logits' dtype (icv_module.py:122-123)
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                float x[EPV];
                unpack<DT>(raw[k], x);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    x[e] = ex2(fmaf(Fmt<DT>::round(x[e] * it_row), c_row, -m));
                    z += x[e];
                }
#pragma unroll
                for (int h = 0; h < Q; ++h)
                    dst[(k * Q + h) * kT + tid] =
                        make_float4(x[4 * h], x[4 * h + 1], x[4 * h + 2], x[4 * h + 3]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                float x[EPV];
                unpack<DT>(raw[k], x);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    x[e] = ex2(fmaf(x[e], c_row, -m));
                    z += x[e];
                }
#pragma unroll
                for (int h = 0; h < Q; ++h)
                    dst[(k * Q + h) * kT + tid] =
                        make_float4(x[4 * h], x[4 * h + 1], x[4 * h + 2], x[4 * h + 3]);
            }
        }
        return z;
    };

    // software pipeline over this cluster's rows: the raw logits of row `r` are in registers,
    // (tr_n, lab_n) describe the row after it
    int64_t r = cluster_id;
    int tr = fetch_tr(r), lab = narrow_lab(fetch_lab(r));
    if (r < a.n_rows) load_raw(r, tr, lab);
    int tr_n = fetch_tr(r + n_clusters);
    int64_t lab_n = fetch_lab(r + n_clusters);
#ifdef LICV_TRACE
    int trace_row = -1;
#endif
    while (r < a.n_rows) {
        const int64_t rn = r + n_clusters;
#ifdef LICV_TRACE
        ++trace_row;
#endif
        LICV_TP(0);
        const bool has_kl = tr >= 0, has_ce = lab != kLabNone;
        if (!has_kl && !has_ce) {
            // neither loss touches this row: its gradient is zero
            if (a.dstu) {
                float z[EPV];
#pragma unroll
                for (int e = 0; e < EPV; ++e) z[e] = 0.f;
                const char* xr = x_row(r);
                char* gp = g_row(r);
                const int j0 = j_first(xr);
                const bool full = all_full(j0), g_vec = phase16(gp) == phase16(xr);
#pragma unroll
                for (int k = 0; k < NV; ++k) store_grad(gp, g_vec, j0, full, k, z);
            }
            if (rank == 0 && tid == 0) { row_kl[r] = 0.f; row_ce[r] = 0.f; }
        } else {
            // KL rows work on the tempered logits, CE-only rows on the raw ones
            const float it_row = has_kl ? inv_t : 1.0f;
            const bool rnd_row = has_kl && round_tempered;
            const float c_row = rnd_row ? kLog2e : kLog2e * it_row;
            // the label logit, before the gradient may overwrite the row in place
            const bool lab_ok = has_ce && lab >= 0 && lab < V;
            float x_lab = 0.f;
            if (rank == 0 && tid == 0 && lab_ok) x_lab = load_elem<DT>(x_row(r), lab);

            // ---- sweep B: exponentials relative to the thread's own maxima -> cache -----------
            MZ2 mine{kNoMaxI, 0.f, kNoMaxI, 0.f};
            mask_row_vecs<DT, NV>(xs, j_first(x_row(r)), kStep, V, tid & 31);
            if (has_kl) mask_row_vecs<DT, NV>(xt, j_first(x_row(r)), kStep, V, tid & 31);
            mine.ms = thread_max(xs, it_row, rnd_row);
            mine.zs = sweep_b(xs, (float)mine.ms, c_row, it_row, rnd_row, cs);
            if (has_kl) {
                mine.mt = thread_max(xt, it_row, rnd_row);
                mine.zt = sweep_b(xt, (float)mine.mt, c_row, it_row, rnd_row, ct);
            }
            LICV_TP(1);
            // ---- reduction 1: maxima and partition sums of both rows.  The partial is sent
            //      first; the request for the next row's logits (the raw registers are free now)
            //      is issued while the partials travel, and lands during the sweeps below. -------
            Exchange& ex1 = row_par ? ex1b : ex1a;
            row_par ^= 1u;
            {
                const MZ2 w = mz_warp(mine);
                exchange_send<kWarps>(ex1, make_float4(__int_as_float(w.ms), w.zs, __int_as_float(w.mt), w.zt),
                              rank, C);
            }
            const int lab_next = narrow_lab(lab_n);
            if (rn < a.n_rows) load_raw(rn, tr_n, lab_next);
            const int tr_nn = fetch_tr(rn + n_clusters);
            const int64_t lab_nn = fetch_lab(rn + n_clusters);
            LICV_TP(2);
            const MZ2 tot = exchange_recv_mz<kWarps>(ex1, C);
            LICV_TP(3);
            const float fs = pow2i(mine.ms - tot.ms) * rcp(tot.zs);      // q = e_s * fs

            // ---- sweep C: KL terms and W_n ------------------------------------------------------
            float W = 0.f, kl_row = 0.f;
            if (has_kl) {
                const float ft = pow2i(mine.mt - tot.mt) * rcp(tot.zt);  // p = e_t * ft
                float klp = 0.f, wp = 0.f;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
#pragma unroll
                    for (int h = 0; h < Q; ++h) {
                        const float4 es = cs[(k * Q + h) * kT + tid];
                        const float4 et = ct[(k * Q + h) * kT + tid];
                        const float ea[4] = {es.x, es.y, es.z, es.w};
                        const float eb[4] = {et.x, et.y, et.z, et.w};
                        float wk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float q = ea[e] * fs;
                            const float p = eb[e] * ft;
                            const float rq = rcp(q + eps);
                            // ln(p+eps) - ln(q+eps) = ln((p+eps)/(q+eps))
                            klp = fmaf(p, lg2((p + eps) * rq), klp);
                            const float w = p * q * rq;
                            wp += w;
                            wk[e] = w * kl_w;
                        }
                        if (a.dstu) ct[(k * Q + h) * kT + tid] = make_float4(wk[0], wk[1], wk[2], wk[3]);
                    }
                }
                LICV_TP(4);
                // ---- reduction 2: KL_n and W_n --------------------------------------------------
                exchange_send<kWarps>(ex2x, make_float4(warp_sum(klp), warp_sum(wp), 0.f, 0.f), rank, C);
                const float2 r2 = exchange_recv_sum<kWarps>(ex2x, C);
                LICV_TP(5);
                kl_row = r2.x * kLn2;
                W = r2.y;
            }

            // ---- sweep D: gradient ----------------------------------------------------------------
            if (a.dstu) {
                const float ce_on = has_ce ? ce_w : 0.f;
                const float A = fs * fmaf(kl_w, W, ce_on);    // has_kl false: W = 0 -> fs * ce_w
                const char* xr = x_row(r);
                char* gp = g_row(r);
                const int j0 = j_first(xr);
                const bool full = all_full(j0), g_vec = phase16(gp) == phase16(xr);
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    float gr[EPV];
#pragma unroll
                    for (int h = 0; h < Q; ++h) {
                        const float4 es = cs[(k * Q + h) * kT + tid];
                        float4 wk = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (has_kl) wk = ct[(k * Q + h) * kT + tid];
                        gr[4 * h] = fmaf(es.x, A, -wk.x);
                        gr[4 * h + 1] = fmaf(es.y, A, -wk.y);
                        gr[4 * h + 2] = fmaf(es.z, A, -wk.z);
                        gr[4 * h + 3] = fmaf(es.w, A, -wk.w);
                    }
                    if (has_ce) {
                        const unsigned rel = (unsigned)(lab - (j0 + k * kStep));
                        if (rel < (unsigned)EPV && lab >= 0) {
#pragma unroll
                            for (int e = 0; e < EPV; ++e)
                                if (e == (int)rel) gr[e] -= ce_on;
                        }
                    }
                    store_grad(gp, g_vec, j0, full, k, gr);
                }
            }
            LICV_TP(6);
            if (rank == 0 && tid == 0) {
                row_kl[r] = kl_row;
                float ce = 0.f;
                if (has_ce)  // an out-of-range label is an error in torch; poison the loss instead
                    ce = lab_ok ? ((float)tot.ms + lg2(tot.zs)) * kLn2 - x_lab : __int_as_float(0x7fc00000);
                row_ce[r] = ce;
            }
            r = rn;
            tr = tr_n; lab = lab_next;
            tr_n = tr_nn; lab_n = lab_nn;
            continue;
        }
        // (row in neither loss) advance: nothing was prefetched for the next row yet
        r = rn;
        tr = tr_n; lab = narrow_lab(lab_n);
        if (r < a.n_rows) load_raw(r, tr, lab);
        tr_n = fetch_tr(r + n_clusters); lab_n = fetch_lab(r + n_clusters);
    }

    cluster_sync_all();   // no CTA leaves while a peer may still send to it
    // ---- the last CTA to finish reduces the per-row losses (fixed order: deterministic) --------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned done_ctas = atomicAdd(a.counter, 1u);
        s_last = (done_ctas == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        float tk = 0.f, tc = 0.f;
        for (int64_t i = tid; i < a.n_rows; i += kT) {
            tk += __ldcg(row_kl + i);
            tc += __ldcg(row_ce + i);
        }
        tk = warp_sum(tk);
        tc = warp_sum(tc);
        if ((tid & 31) == 0) { s_tot[tid >> 5] = tk; s_tot[kWarps + (tid >> 5)] = tc; }
        __syncthreads();
        if (tid == 0) {
            tk = 0.f; tc = 0.f;
            for (int w = 0; w < kWarps; ++w) { tk += s_tot[w]; tc += s_tot[kWarps + w]; }
            const float kl = use_kl ? tk * T * T / (float)n_kl : 0.f;
            const float ce = use_ce ? tc / (float)n_ce : 0.f;
            a.out_losses[0] = kl;
            a.out_losses[1] = ce;
            a.out_losses[2] = a.only_hard_loss ? ce : (use_ce ? fmaf(a.hard_loss_weight, ce, kl) : kl);
            *a.counter = 0u;  // leave the workspace ready for the next call
        }
    }
}

// =================================================================================================
// Variant: ONE CTA per SM owns a whole row pair; no cross-SM synchronisation at all.
//
// A row pair of 32k 16-bit logits needs 256 KB of fp32 cache (e_s and e_t / kl_w w), more than
// the 227 KB of shared memory - but an SM also has 256 KB of TENSOR MEMORY (512 columns x 128
// lanes x 32 bit), readable and writable from registers with tcgen05.ld / tcgen05.st (SASS LDTM /
// STTM).  Here e_s lives in shared memory (128 KB, thread-private float4 slots) and e_t / w in
// TMEM (256 columns = 128 KB: thread t of warp w owns 64 columns of lane 32 (w % 4) + t), so the
// three sweeps of a row run inside one 512-thread CTA and the two row-wide reductions are plain
// CTA barriers among 16 warps of the same SM - the cluster kernel above loses ~45 % of its time
// waiting for the slowest of 32 warps spread over 8 SMs.  The raw logits of the next row are
// prefetched into registers right after sweep B.  16-bit logits, V <= 32760.
// =================================================================================================

constexpr int kMaxOrdered = 1024;   // launches of up to this many rows are worked in cost order

template <int DT, int kMT, int kMNV>
__global__ void __launch_bounds__(kMT, 1) kd_loss_tmem_kernel(KdArgs a) {
    constexpr int EPV = Fmt<DT>::kPerVec;   // 8
    constexpr int EB = Fmt<DT>::kBytes;     // 2
    constexpr int NV = kMNV;
    constexpr int kMW = kMT / 32;
    constexpr int kMCols = tmem_cols(kMT, kMNV);
    constexpr int kStep = kMT * EPV;
    constexpr int kMeta = 64;
    static_assert(EPV == 8, "16-bit logits only");
    extern __shared__ __align__(16) float4 cs[];           // [NV * 2][kMT]: e_s, fp32
    __shared__ __align__(16) float4 part[2][kMW];
    __shared__ uint32_t s_tmem;
    __shared__ float s_tot[2 * kMW];
    __shared__ int s_last;
    __shared__ int s_tr[kMeta], s_lab[kMeta];               // (teacher row, label) of the next rows
    // Work order (see build_order below): rows sorted by cost class, dealt to the CTAs in snake order
    constexpr int kOrdRounds = (kMaxOrdered + kMT - 1) / kMT;
    __shared__ uint16_t s_order[kMaxOrdered];
    __shared__ int s_ocnt[3 * kOrdRounds * kMW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&s_tmem)),
                     "n"(kMCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    pdl_launch_dependents();
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // this thread's 64 columns: lane quarter of the warp, column group of the warp
    const uint32_t tcol = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * (NV * 8));
    pdl_wait();

    const float T = a.temperature;
    const float inv_t = 1.0f / T;
    const bool round_tempered = (a.round_flags & LICV_ROUND_TEMPERED) && T != 1.0f;
    const int64_t n_kl = a.counts ? (int64_t)a.counts[0] : a.n_kl;
    const int64_t n_ce = a.counts ? (int64_t)a.counts[1] : a.n_ce;
    const bool use_kl = !a.only_hard_loss;
    const bool use_ce = a.ce_label != nullptr;
    const float kl_w = use_kl ? a.grad_scale * T / (float)n_kl : 0.f;
    const float ce_w = use_ce ? a.grad_scale * (a.only_hard_loss ? 1.0f : a.hard_loss_weight) /
                                    (float)n_ce
                              : 0.f;
    const float eps = a.kl_eps;
    float* row_kl = a.row_loss;
    float* row_ce = a.row_loss + a.n_rows;
    const int V = a.vocab;

    auto fetch_tr = [&](int64_t r) -> int {
        if (!use_kl || r >= a.n_rows) return -1;
        return a.kl_tea_row ? a.kl_tea_row[r] : (int)r;
    };
    auto fetch_lab = [&](int64_t r) -> int {
        if (!use_ce || r >= a.n_rows) return kLabNone;
        const int64_t l = a.ce_label[r];
        return (l < -100 || l > 0x7fffffff) ? kLabBad : (int)l;
    };
    auto x_row = [&](int64_t r) { return static_cast<const char*>(a.stu) + (size_t)r * a.stu_stride * EB; };
    auto t_row = [&](int tr) { return static_cast<const char*>(a.tea) + (size_t)tr * a.tea_stride * EB; };
    auto g_row = [&](int64_t r) { return static_cast<char*>(a.dstu) + (size_t)r * a.stu_stride * EB; };
    auto phase16 = [](const void* p) { return (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15u); };
    auto j_first = [&](const char* xr) -> int { return tid * EPV - (int)(phase16(xr) / EB); };
    // every vector of every lane of this warp lies wholly inside the row: no per-vector checks.
    // Decided per warp, like the loads: a lane that took the slow store path on its own made its
    // warp run both paths for all NV vectors and arrive late at the row's next barrier.
    auto all_full = [&](int j0) -> bool { return warp_inside_row<DT>(j0, kStep, NV, V, tid & 31); };

    uint4 xs[NV], xt[NV];
    auto load_raw = [&](int64_t r, int tr, int lab) {
        if (tr < 0 && lab == kLabNone) return;
        const char* xr = x_row(r);
        const int j0 = j_first(xr);
        load_row_vecs<DT, NV>(xs, xr, j0, kStep, V, 16, lane);
        if (tr >= 0) {
            const char* tp = t_row(tr);
            const uint32_t dph = (phase16(tp) - phase16(xr)) & 15u;
            load_row_vecs<DT, NV>(xt, tp, j0, kStep, V, dph == 0 ? 16 : (int)(dph & (0u - dph)), lane);
        }
    };
    auto thread_max = [&](const uint4 (&raw)[NV], float it_row, bool rnd) -> int {
        RawMax<DT> mx;
#pragma unroll
        for (int k = 0; k < NV; ++k) mx.add(raw[k]);
        float m = mx.get() * it_row;
        if (rnd) m = Fmt<DT>::round(m);
        m = fminf(fmaxf(ceilf(m * kLog2e), -1.0e6f), 1.0e6f);
        return (int)m;
    };
    // CTA-wide reductions: warp partials through shared memory, one barrier each
    auto cta_mz = [&](MZ2 v, float4* slab) -> MZ2 {
        const MZ2 w = mz_warp(v);
        if (lane == 0) slab[warp] = make_float4(__int_as_float(w.ms), w.zs, __int_as_float(w.mt), w.zt);
        __syncthreads();
        MZ2 c{kNoMaxI, 0.f, kNoMaxI, 0.f};
        if (lane < kMW) {
            const float4 p = slab[lane];
            c = MZ2{__float_as_int(p.x), p.y, __float_as_int(p.z), p.w};
        }
        return mz_warp(c);
    };
    auto cta_sum2 = [&](float x, float y, float4* slab) -> float2 {
        x = warp_sum(x);
        y = warp_sum(y);
        if (lane == 0) slab[warp] = make_float4(x, y, 0.f, 0.f);
        __syncthreads();
        float sx = 0.f, sy = 0.f;
        if (lane < kMW) {
            const float4 p = slab[lane];
            sx = p.x;
            sy = p.y;
        }
        return make_float2(warp_sum(sx), warp_sum(sy));
    };

    const int64_t stride = gridDim.x;
    // Work order.  A KL row costs about two CE-only rows (two cached streams, three sweeps) and a
    // row in neither loss only a zero fill, and this kernel runs the launches with a few rows per
    // SM (the training step: 256 rows on 148 SMs), where dealing rows r, r + G, ... to CTA r gave
    // some SMs a KL row AND a second row while the launch waited for them.  Every CTA therefore
    // sorts the rows by class (KL, CE only, neither; stable, from the row lists alone - integer
    // work, the same in every CTA) and the sorted list is dealt in snake order (pass 0: CTA b takes
    // item b, pass 1: item 2G-1-b, ...), the longest-first rule for identical machines: the SMs
    // that got the dear rows get the cheap second rows, or none.  Results do not depend on which
    // CTA works a row (per-row values, reduced in row order at the end).
    const bool ordered = a.n_rows > stride && a.n_rows <= kMaxOrdered;
    if (ordered) {
        int cls[kOrdRounds], before[kOrdRounds];
#pragma unroll
        for (int q = 0; q < kOrdRounds; ++q) {
            const int64_t t = (int64_t)q * kMT + tid;
            int c = 3;
            if (t < a.n_rows) c = fetch_tr(t) >= 0 ? 0 : (fetch_lab(t) != kLabNone ? 1 : 2);
            cls[q] = c;
            before[q] = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const unsigned b = __ballot_sync(0xffffffffu, c == k);
                if (c == k) before[q] = __popc(b & ((1u << lane) - 1u));
                if (lane == 0) s_ocnt[(k * kOrdRounds + q) * kMW + warp] = __popc(b);
            }
        }
        __syncthreads();
        if (warp == 0) {   // exclusive scan of the 3 * rounds * warps counts, class-major
            constexpr int kN = 3 * kOrdRounds * kMW, kPer = (kN + 31) / 32;
            int v[kPer], sum = 0;
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int idx = lane * kPer + i;
                v[i] = idx < kN ? s_ocnt[idx] : 0;
                sum += v[i];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            int run = incl - sum;
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
                const int idx = lane * kPer + i;
                if (idx < kN) s_ocnt[idx] = run;
                run += v[i];
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kOrdRounds; ++q)
            if (cls[q] < 3)
                s_order[s_ocnt[(cls[q] * kOrdRounds + q) * kMW + warp] + before[q]] = (uint16_t)(q * kMT + tid);
        // (the barrier inside fill_meta publishes s_order)
    }
    // the row this CTA works as its `ord`-th, or n_rows when it has no such row
    auto row_of = [&](int ord) -> int64_t {
        const int64_t b = blockIdx.x;
        const int64_t pos = (int64_t)ord * stride + ((ordered && (ord & 1)) ? stride - 1 - b : b);
        if (pos >= a.n_rows) return a.n_rows;
        return ordered ? (int64_t)s_order[pos] : pos;
    };
    // Row descriptors come from shared memory, refilled kMeta rows at a time: a label load issued
    // behind the next row's prefetch and narrowed at once stalled every warp for a full HBM round
    // trip per row (6 % of the kernel in one IADD3), and carrying the raw 64-bit label a row
    // instead costs two registers this kernel does not have.
    auto fill_meta = [&](int ord0) {
        __syncthreads();
        if (tid < kMeta) {
            const int64_t rr = row_of(ord0 + tid);
            s_tr[tid] = fetch_tr(rr);
            s_lab[tid] = fetch_lab(rr);
        }
        __syncthreads();
    };
    int it = 0, meta0 = 0;                                  // this CTA's row ordinal, first cached ordinal
    fill_meta(0);
    int64_t r = row_of(0);
    int tr = s_tr[0], lab = s_lab[0];
    if (r < a.n_rows) load_raw(r, tr, lab);
    int64_t rn;
    for (; r < a.n_rows; r = rn, ++it) {
        rn = row_of(it + 1);
        if (it + 1 - meta0 == kMeta) {
            fill_meta(it + 1);
            meta0 = it + 1;
        }
        const bool has_kl = tr >= 0, has_ce = lab != kLabNone;
        const char* xr = x_row(r);
        const int j0 = j_first(xr);
        const bool full = all_full(j0);
        char* gp = a.dstu ? g_row(r) : nullptr;
        const bool g_vec = gp && phase16(gp) == phase16(xr);
        auto store_grad = [&](int k, const float* gr) {
            const int jwk = j0 - lane * EPV + k * kStep;            // lane 0's element of vector k
            if (g_vec && (full || (jwk >= 0 && jwk + 32 * EPV <= V)))
                st_vec(reinterpret_cast<uint4*>(gp + ((int64_t)j0 + (int64_t)k * kStep) * EB), pack<DT>(gr));
            else
                store_row_vec<DT>(gp, j0 + k * kStep, V, g_vec, gr);
        };
        if (!has_kl && !has_ce) {
            if (gp) {
                float z[EPV];
#pragma unroll
                for (int e = 0; e < EPV; ++e) z[e] = 0.f;
#pragma unroll
                for (int k = 0; k < NV; ++k) store_grad(k, z);
            }
            if (tid == 0) { row_kl[r] = 0.f; row_ce[r] = 0.f; }
        } else {
            const float it_row = has_kl ? inv_t : 1.0f;
            const bool rnd_row = has_kl && round_tempered;
            const float c_row = rnd_row ? kLog2e : kLog2e * it_row;
            const bool lab_ok = has_ce && lab >= 0 && lab < V;
            float x_lab = 0.f;
            if (tid == 0 && lab_ok) x_lab = load_elem<DT>(xr, lab);

            // ---- sweep B: e_s -> shared memory, e_t -> tensor memory --------------------------
            MZ2 mine{kNoMaxI, 0.f, kNoMaxI, 0.f};
            mask_row_vecs<DT, NV>(xs, j0, kStep, V, lane);
            if (has_kl) mask_row_vecs<DT, NV>(xt, j0, kStep, V, lane);
            mine.ms = thread_max(xs, it_row, rnd_row);
            {
                const float m = (float)mine.ms;
                auto body = [&](auto rnd_tag) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        float x[EPV];
                        unpack<DT>(xs[k], x);
#pragma unroll
                        for (int e = 0; e < EPV; ++e) {
                            float u = x[e];
                            if (decltype(rnd_tag)::value) u = Fmt<DT>::round(u * it_row);
                            x[e] = ex2(fmaf(u, c_row, -m));
                            mine.zs += x[e];
                        }
                        cs[(k * 2) * kMT + tid] = make_float4(x[0], x[1], x[2], x[3]);
                        cs[(k * 2 + 1) * kMT + tid] = make_float4(x[4], x[5], x[6], x[7]);
                    }
                };
                if (rnd_row) body(std::true_type{}); else body(std::false_type{});
            }
            if (has_kl) {
                mine.mt = thread_max(xt, it_row, rnd_row);
                const float m = (float)mine.mt;
                auto body = [&](auto rnd_tag) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        float x[EPV];
                        unpack<DT>(xt[k], x);
#pragma unroll
                        for (int e = 0; e < EPV; ++e) {
                            float u = x[e];
                            if (decltype(rnd_tag)::value) u = Fmt<DT>::round(u * it_row);
                            x[e] = ex2(fmaf(u, c_row, -m));
                            mine.zt += x[e];
                        }
                        tmem_st8(tcol + k * 8, x);
                    }
                };
                if (rnd_row) body(std::true_type{}); else body(std::false_type{});
                tmem_wait_st();
            }
            // ---- reduction 1 (one CTA barrier) ------------------------------------------------
            const MZ2 tot = cta_mz(mine, part[0]);
            const float fs = pow2i(mine.ms - tot.ms) * rcp(tot.zs);
            // the raw registers are free: request the next row (after the reduction, so that
            // nothing the reduction needs has to live across this burst of loads)
            const int tr_n = s_tr[it + 1 - meta0], lab_n = s_lab[it + 1 - meta0];
            if (rn < a.n_rows) load_raw(rn, tr_n, lab_n);

            // ---- sweep C ----------------------------------------------------------------------
            float W = 0.f, kl_row = 0.f;
            if (has_kl) {
                const float ft = pow2i(mine.mt - tot.mt) * rcp(tot.zt);
                float klp = 0.f, wp = 0.f;
                // e_t comes back from tensor memory two vectors at a time, one load ahead of the
                // arithmetic
                uint32_t tb[2][8];
                tmem_ld8_issue(tcol, tb[0]);
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    tmem_ld_wait(tb[k & 1]);
                    if (k + 1 < NV) tmem_ld8_issue(tcol + (k + 1) * 8, tb[(k + 1) & 1]);
                    float wk[8];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 es = cs[(k * 2 + h) * kMT + tid];
                        const float ea[4] = {es.x, es.y, es.z, es.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float q = ea[e] * fs;
                            const float p = __uint_as_float(tb[k & 1][4 * h + e]) * ft;
                            const float rq = rcp(q + eps);
                            klp = fmaf(p, lg2((p + eps) * rq), klp);
                            const float w = p * q * rq;
                            wp += w;
                            wk[4 * h + e] = w * kl_w;
                        }
                    }
                    if (gp) tmem_st8(tcol + k * 8, wk);
                }
                tmem_wait_st();
                const float2 r2 = cta_sum2(klp, wp, part[1]);
                kl_row = r2.x * kLn2;
                W = r2.y;
            }

            // ---- sweep D ----------------------------------------------------------------------
            if (gp) {
                const float ce_on = has_ce ? ce_w : 0.f;
                const float A = fs * fmaf(kl_w, W, ce_on);
                uint32_t tb[2][8];
                if (has_kl) tmem_ld8_issue(tcol, tb[0]);
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    if (has_kl) {
                        tmem_ld_wait(tb[k & 1]);
                        if (k + 1 < NV) tmem_ld8_issue(tcol + (k + 1) * 8, tb[(k + 1) & 1]);
                    }
                    const float4 e0 = cs[(k * 2) * kMT + tid], e1 = cs[(k * 2 + 1) * kMT + tid];
                    const float ea[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                    float gr[EPV];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float wk = has_kl ? __uint_as_float(tb[k & 1][e]) : 0.f;
                        gr[e] = fmaf(ea[e], A, -wk);
                    }
                    if (has_ce) {
                        const unsigned rel = (unsigned)(lab - (j0 + k * kStep));
                        if (rel < (unsigned)EPV && lab >= 0) {
#pragma unroll
                            for (int e = 0; e < EPV; ++e)
                                if (e == (int)rel) gr[e] -= ce_on;
                        }
                    }
                    store_grad(k, gr);
                }
            }
            if (tid == 0) {
                row_kl[r] = kl_row;
                float ce = 0.f;
                if (has_ce)
                    ce = lab_ok ? ((float)tot.ms + lg2(tot.zs)) * kLn2 - x_lab : __int_as_float(0x7fc00000);
                row_ce[r] = ce;
            }
            tr = tr_n; lab = lab_n;
            // no barrier here: the caches are thread-private and each of part[0] / part[1] is
            // rewritten only after the OTHER reduction's barrier, which every reader has passed
            continue;
        }
        tr = s_tr[it + 1 - meta0]; lab = s_lab[it + 1 - meta0];
        if (rn < a.n_rows) load_raw(rn, tr, lab);
    }

    // TMEM is released by the warp that allocated it, after every warp is done with it
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(kMCols)
                     : "memory");
    }
    // ---- the last CTA to finish reduces the per-row losses (fixed order: deterministic) --------
    if (tid == 0) {
        __threadfence();
        const unsigned done_ctas = atomicAdd(a.counter, 1u);
        s_last = (done_ctas == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        float tk = 0.f, tc = 0.f;
        for (int64_t i = tid; i < a.n_rows; i += kMT) {
            tk += __ldcg(row_kl + i);
            tc += __ldcg(row_ce + i);
        }
        tk = warp_sum(tk);
        tc = warp_sum(tc);
        if (lane == 0) { s_tot[warp] = tk; s_tot[kMW + warp] = tc; }
        __syncthreads();
        if (tid == 0) {
            tk = 0.f; tc = 0.f;
            for (int w = 0; w < kMW; ++w) { tk += s_tot[w]; tc += s_tot[kMW + w]; }
            const float kl = use_kl ? tk * T * T / (float)n_kl : 0.f;
            const float ce = use_ce ? tc / (float)n_ce : 0.f;
            a.out_losses[0] = kl;
            a.out_losses[1] = ce;
            a.out_losses[2] = a.only_hard_loss ? ce : (use_ce ? fmaf(a.hard_loss_weight, ce, kl) : kl);
            *a.counter = 0u;
        }
    }
}

inline int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

template <int DT, int NV, int NT>
int launch_cluster(const KdArgs& a, int C, cudaStream_t st) {
    auto kern = kd_loss_cluster_kernel<DT, NV, NT>;
    constexpr size_t smem = (size_t)2 * NV * Fmt<DT>::kPerVec * NT * sizeof(float);
    if (C > 8 || C < 1) return LICV_ERR_BAD_ARGUMENT;
    static PerDevice<int> raised;
    if (smem > 48 * 1024) {
        const int e = raised.get([&] {
            return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        });
        if (e != 0) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;   // st.async / mapa are cluster instructions: C = 1 is launched as a cluster too
    cfg.numAttrs = launch_attrs(attr, C);
    // clusters resident at once (per instantiation, cluster size and device)
    static PerDevice<int64_t> cap_by_c[9];
    const int64_t cap_c = cap_by_c[C].get([&]() -> int64_t {
        int n = 0;
        cfg.gridDim = dim3(C * device_info().sm_count);
        if (C > 1 && cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) return n;
        cudaGetLastError();
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem) != cudaSuccess ||
            per_sm < 1)
            per_sm = 1;
        const int64_t c = (int64_t)per_sm * device_info().sm_count / C;
        return c < 1 ? 1 : c;
    });
    int64_t clusters = a.n_rows < cap_c ? a.n_rows : cap_c;
    if (clusters < 1) clusters = 1;   // no rows: the finalising CTA still reports mean-of-empty
    cfg.gridDim = dim3((unsigned)(clusters * C));
    return (int)cudaLaunchKernelEx(&cfg, kern, a);
}

template <int DT>
int dispatch_nv(const KdArgs& a, int C, int NV, int NT, cudaStream_t st) {
    if (NT == 128) {
        switch (NV) {
            case 4: return launch_cluster<DT, 4, 128>(a, C, st);
            default: return LICV_ERR_BAD_ARGUMENT;
        }
    }
    switch (NV) {
        case 1: return launch_cluster<DT, 1, 256>(a, C, st);
        case 2: return launch_cluster<DT, 2, 256>(a, C, st);
        case 4: return launch_cluster<DT, 4, 256>(a, C, st);
        default: return launch_cluster<DT, 8, 256>(a, C, st);
    }
}

template <int DT, int MT, int MNV>
int launch_tmem(const KdArgs& a, cudaStream_t st) {
    auto kern = kd_loss_tmem_kernel<DT, MT, MNV>;
    constexpr size_t smem = (size_t)MNV * 2 * MT * sizeof(float4);   // e_s: 128-144 KB
    static PerDevice<int> raised;
    const int e = raised.get([&] {
        return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    });
    if (e != 0) return e;
    int64_t grid = a.n_rows < device_info().sm_count ? a.n_rows : device_info().sm_count;
    if (grid < 1) grid = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(MT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = launch_attrs(attr, 0);
    return (int)cudaLaunchKernelEx(&cfg, kern, a);
}
template <int DT>
int launch_tmem_shape(const KdArgs& a, cudaStream_t st) {
    static const int threads = env_int("LICV_KD_TMEM_THREADS", 512);
    if (threads == 1024) return launch_tmem<DT, 1024, 4>(a, st);
    if (threads == 768) return launch_tmem<DT, 768, 6>(a, st);
    return launch_tmem<DT, 512, 8>(a, st);
}

}  // namespace

bool kd_tmem_plan(int vocab, int dtype, float temperature, bool kl_and_ce) {
    static const int on = env_int("LICV_KD_TMEM", 1);   // 0: always the cluster kernel
    if (!on || dtype == LICV_F32) return false;
    if (kl_and_ce && temperature != 1.0f) return false;
    return (int64_t)(vocab + 7) / 8 + 1 <= 4096;    // vectors of 8 elements a CTA holds
}
int launch_kd_tmem(const KdArgs& a, int dtype, cudaStream_t st) {
    return dtype == LICV_BF16 ? launch_tmem_shape<LICV_BF16>(a, st) : launch_tmem_shape<LICV_F16>(a, st);
}

namespace {
}  // namespace

#ifdef LICV_TRACE
extern "C" int licv_debug_read_trace(long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * n);
}
#endif

bool kd_cluster_plan(int vocab, int dtype, float temperature, bool kl_and_ce, int* C, int* NV,
                     int* NT) {
    static const int off = env_int("LICV_KD_NO_CLUSTER", 0);
    static const int small_ctas = env_int("LICV_KD_THREADS", 128);   // 128: six 4-warp CTAs per SM
    if (off) return false;
    // a KL + CE row at T != 1 needs a third exponential stream (CE works on the raw logits):
    // left to the generic kernel
    if (kl_and_ce && temperature != 1.0f) return false;
    const int epv = dtype == LICV_F32 ? 4 : 8;
    const int64_t need = ((int64_t)vocab + epv - 1) / epv + 1;   // + 1: a row may straddle
    *NT = 256;
    if (small_ctas == 128 && (int64_t)8 * 4 * 128 >= need && (int64_t)4 * 4 * 128 < need) {
        *C = 8; *NV = 4; *NT = 128;    // rows of 2049..4096 vectors: 8 CTAs x 128 threads x 4
        return true;
    }
    for (int nv = 1; nv <= 4; nv *= 2)
        if ((int64_t)nv * kMaxT >= need) { *C = 1; *NV = nv; return true; }
    for (int c = 2; c <= 8; c *= 2)
        if ((int64_t)c * 4 * kMaxT >= need) { *C = c; *NV = 4; return true; }
    if ((int64_t)8 * 8 * kMaxT >= need) { *C = 8; *NV = 8; return true; }
    return false;
}

int launch_kd_cluster(const KdArgs& a, int dtype, int C, int NV, int NT, cudaStream_t st) {
    switch (dtype) {
        case LICV_F32: return dispatch_nv<LICV_F32>(a, C, NV, NT, st);
        case LICV_BF16: return dispatch_nv<LICV_BF16>(a, C, NV, NT, st);
        default: return dispatch_nv<LICV_F16>(a, C, NV, NT, st);
    }
}

}  // namespace licv
