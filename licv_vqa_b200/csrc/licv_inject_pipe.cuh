// Injection backward for wide rows, staged through shared memory by the TMA engine (sm_100a).
//
// The register-staged kernel of licv_inject.cuh has to hold, per thread, its slice of the shift
// vector, its d_shift accumulators AND every 128-bit load it wants in flight; for the backward
// (two input streams) that caps the bytes in flight per SM well below what HBM3e needs.  Here the
// loads are taken out of the register file: one elected thread issues `cp.async.bulk` (1-D TMA
// bulk copies, SASS UBLKCP) of whole token rows of h and g into a ring of shared-memory stages,
// several stages ahead of the arithmetic, each stage completing on its own mbarrier.  The 256
// threads of the CTA form one row group (thread t owns the 16-byte vectors t, t + 256, ... of
// every row); per stage they sweep their slices out of shared memory (conflict-free LDS.128) to
// reduce the three dot products of each token, sweep them a second time for dh and d_shift, and
// write dh with 128-bit stores - no token data is held in registers across the reduction.  One
// CTA barrier per stage; the stage of the previous iteration is handed back to the TMA engine
// right after it.
//
//   HBM traffic: h and g read once, dh written once (3 e d bytes per token), d_shift leaves each
//   CTA once as REDG.F32x4.
//
// Used when a row is a whole number of 256-vector sweeps (d = 2048 k for bf16/fp16, 1024 k for
// fp32) up to 4 sweeps; everything else takes the register-staged kernel.
#pragma once

#include "licv_inject.cuh"

namespace licv {
namespace inject {

constexpr int kPipeThreads = 256;
constexpr int kPipeMaxStages = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LICV_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LICV_DONE;\n"
        "bra LICV_WAIT;\n"
        "LICV_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// streaming data: evict-first in L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(const void* p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "r"(smem_u32(p)));
    return r;
}


// ---- thread-block cluster helpers (distributed shared memory) --------------------------------
__device__ __forceinline__ uint32_t cluster_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_num_ctas() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(addr));
    return v;
}

#ifdef LICV_TRACE
// debug build only: per CTA {entry, after griddepcontrol.wait, first stage landed, first reduction
// done, loop done, d_shift out} in ns (globaltimer), of the LAST launch in this translation unit
static __device__ unsigned long long g_pipe_trace[512 * 8];
#define LICV_PIPE_STAMP(i)                                                                     \
    do {                                                                                       \
        if (threadIdx.x == 0 && blockIdx.x < 512) {                                            \
            unsigned long long t_;                                                             \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                            \
            g_pipe_trace[blockIdx.x * 8 + (i)] = t_;                                           \
        }                                                                                      \
    } while (0)
#else
#define LICV_PIPE_STAMP(i) do { } while (0)
#endif

struct PipeArgs {
    Args a;
    int n_stages;
    int64_t n_batches;   // ceil(n_tok / TB)
};

template <int HDT, int GDT, int VPT, int TB, int RND>
__global__ void __launch_bounds__(kPipeThreads, 3) bwd_pipe_kernel(PipeArgs pa) {
    constexpr int EPV = Fmt<HDT>::kPerVec;
    constexpr int GV = (Fmt<GDT>::kBytes > Fmt<HDT>::kBytes) ? 2 : 1;
    constexpr int GPV = Fmt<GDT>::kPerVec;
    constexpr int P = pow2_at_least(3 * TB);
    constexpr int kWarps = kPipeThreads / kWarp;
    constexpr int kHRow = VPT * kPipeThreads * 16;        // bytes of one h row
    constexpr int kGRow = kHRow * GV;                     // bytes of one g row
    constexpr int kStage = TB * (kHRow + kGRow);

    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(16) float slab[2][kWarps * P];
    __shared__ __align__(8) uint64_t full[kPipeMaxStages];

    const Args& a = pa.a;
    const int tid = threadIdx.x;
    const int S = pa.n_stages;
    const int64_t G = gridDim.x;
    const unsigned char* hb = reinterpret_cast<const unsigned char*>(a.h);
    const unsigned char* gb = reinterpret_cast<const unsigned char*>(a.g);

    LICV_PIPE_STAMP(0);
    if (tid == 32) {
        // hint the first stages' rows into L2 while the previous kernel drains (see prefetch_l2)
        for (int it = 0; it < S; ++it) {
            const int64_t batch = (int64_t)blockIdx.x + it * G;
            if (batch >= pa.n_batches) break;
            const int64_t t0 = batch * TB;
            const int64_t left = a.n_tok - t0;
            const uint32_t ntok = (uint32_t)(left < TB ? left : TB);
            prefetch_l2_bulk(hb + t0 * kHRow, ntok * kHRow);
            prefetch_l2_bulk(gb + t0 * kGRow, ntok * kGRow);
        }
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
    }
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();   // everything above ran while the previous kernel drained; global memory from here
    LICV_PIPE_STAMP(1);

    uint64_t policy = 0;
    // issue the copies of this CTA's `it`-th batch into stage it % S
    auto produce = [&](int64_t it) {
        const int64_t batch = (int64_t)blockIdx.x + it * G;
        if (batch >= pa.n_batches) return;
        const int s = (int)(it % S);
        const int64_t t0 = batch * TB;
        const int64_t left = a.n_tok - t0;
        const uint32_t ntok = (uint32_t)(left < TB ? left : TB);
        unsigned char* dst = ring + (size_t)s * kStage;
        mbar_expect_tx(&full[s], ntok * (uint32_t)(kHRow + kGRow));
        bulk_g2s(dst, hb + t0 * kHRow, ntok * kHRow, &full[s], policy);
        bulk_g2s(dst + TB * kHRow, gb + t0 * kGRow, ntok * kGRow, &full[s], policy);
    };
    if (tid == 0) {
        policy = l2_evict_first_policy();
        for (int it = 0; it < S; ++it) produce(it);
    }

    float s[VPT][EPV];
    load_shift<HDT, VPT, RND>(a.shift, VPT * kPipeThreads, tid, kPipeThreads, s);
    float ds[VPT][EPV];
#pragma unroll
    for (int k = 0; k < VPT; ++k)
#pragma unroll
        for (int e = 0; e < EPV; ++e) ds[k][e] = 0.f;

    int64_t it = 0;
    for (int64_t batch = blockIdx.x; batch < pa.n_batches; batch += G, ++it) {
        const int stage = (int)(it % S);
        const uint32_t parity = (uint32_t)((it / S) & 1);
        const int64_t t0 = batch * TB;
        const unsigned char* sh = ring + (size_t)stage * kStage;
        const unsigned char* sg = sh + TB * kHRow;
        mbar_wait(&full[stage], parity);
        if (it == 0) LICV_PIPE_STAMP(2);

        float acc[P];
#pragma unroll
        for (int i = 0; i < P; ++i) acc[i] = 0.f;
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            float hh = 0.f, yy = 0.f, yg = 0.f;
#pragma unroll
            for (int k = 0; k < VPT; ++k) {
                const int j = tid + k * kPipeThreads;
                float x[EPV], gg[EPV];
                unpack<HDT>(lds128(sh + b * kHRow + j * 16), x);
#pragma unroll
                for (int o = 0; o < GV; ++o)
                    unpack<GDT>(lds128(sg + b * kGRow + (j * GV + o) * 16), gg + o * GPV);
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
        // cross-warp sums; every thread reaching the barrier inside has finished the PREVIOUS
        // iteration, whose stage can therefore be refilled (this one is read again below)
        {
            float* my = slab[it & 1];
            warp_multi_sum<P>(acc);
            const int lane = tid & 31, warp = tid >> 5;
            if ((lane & (32 / P - 1)) == 0) my[warp * P + warp_value_index<P>(lane)] = acc[0];
            __syncthreads();
            if (it == 0) LICV_PIPE_STAMP(3);
            if (tid == 0 && it > 0) produce(it - 1 + S);
#pragma unroll
            for (int i = 0; i < P; ++i) acc[i] = 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                if constexpr (P % 4 == 0) {
#pragma unroll
                    for (int i = 0; i < P; i += 4) {
                        const float4 q = *reinterpret_cast<const float4*>(my + w * P + i);
                        acc[i] += q.x; acc[i + 1] += q.y; acc[i + 2] += q.z; acc[i + 3] += q.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < P; ++i) acc[i] += my[w * P + i];
                }
            }
        }
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
                const int j = tid + k * kPipeThreads;
                float x[EPV], gg[EPV];
                unpack<HDT>(lds128(sh + b * kHRow + j * 16), x);
#pragma unroll
                for (int o = 0; o < GV; ++o)
                    unpack<GDT>(lds128(sg + b * kGRow + (j * GV + o) * 16), gg + o * GPV);
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    float y = x[e] + s[k][e];
                    if (RND >= 1) y = Fmt<HDT>::round(y);
                    const float gy = fmaf(r, gg[e], ky * y);
                    ds[k][e] += gy;
                    x[e] = fmaf(kh, x[e], gy);
                }
                if (a.out != nullptr)
                    st_vec(a.out + (t0 + b) * (VPT * kPipeThreads) + j, pack<HDT>(x));
            }
        }
    }

    // d_shift.  The L2 atomic units retire about one fp32 add per clock per slice and a row of
    // d_shift lives in few slices, so what costs is the NUMBER of CTAs that add: the CTAs of a
    // thread-block cluster first sum their rows through distributed shared memory (CTA r of C
    // sums columns [r d/C, (r+1) d/C) of all C rows), then each issues REDG.F32x4 for its
    // columns only: C times fewer atomics.
    // Spread mode: d_shift leaves as REDG.F32x4 into one of n_rows replicas of the [d] vector
    // (replica = CTA index mod n_rows), added up later by licv_reduce_rows.  The L2 atomic units
    // serialise per address (~27 clocks per contending warp): 256 CTAs on one [d] vector cost
    // 3.5 us, which is what the cluster pre-reduction below buys back at the price of two cluster
    // barriers; with <= 16 CTAs per replica the atomics are free and the CTA leaves at once.
    LICV_PIPE_STAMP(4);
    if (a.rows != nullptr) {
        float* row = a.rows + (int64_t)((int)blockIdx.x & a.row_mask) * (VPT * kPipeThreads * EPV);
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int j = tid + k * kPipeThreads;
#pragma unroll
            for (int e = 0; e < EPV; e += 4)
                red_add_v4(row + (int64_t)j * EPV + e, ds[k][e], ds[k][e + 1], ds[k][e + 2], ds[k][e + 3]);
        }
        LICV_PIPE_STAMP(5);
        return;
    }
    const uint32_t C = cluster_num_ctas();
    if (C == 1) {
#pragma unroll
        for (int k = 0; k < VPT; ++k) {
            const int j = tid + k * kPipeThreads;
#pragma unroll
            for (int e = 0; e < EPV; e += 4)
                red_add_v4(a.d_shift + (int64_t)j * EPV + e, ds[k][e], ds[k][e + 1], ds[k][e + 2],
                           ds[k][e + 3]);
        }
        return;
    }
    constexpr int kRowF4 = VPT * kPipeThreads * EPV / 4;    // float4s in one d_shift row
    float4* row = reinterpret_cast<float4*>(ring);           // every stage has been consumed
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int j = tid + k * kPipeThreads;
#pragma unroll
        for (int e = 0; e < EPV; e += 4)
            row[(j * EPV + e) / 4] = make_float4(ds[k][e], ds[k][e + 1], ds[k][e + 2], ds[k][e + 3]);
    }
    cluster_sync_all();
    const uint32_t rank = cluster_cta_rank();
    const int per = kRowF4 / (int)C;   // C divides kRowF4 (C is 2, 4 or 8)
    for (int i = tid; i < per; i += kPipeThreads) {
        const int q = (int)rank * per + i;
        float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
        for (uint32_t r = 0; r < C; ++r) {
            const float4 v = ld_dsmem_f4(dsmem_addr(row + q, r));
            tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
        }
        red_add_v4(a.d_shift + (int64_t)q * 4, tot.x, tot.y, tot.z, tot.w);
    }
    cluster_sync_all();   // no CTA may exit while a peer still reads its row
}

}  // namespace inject
}  // namespace licv
