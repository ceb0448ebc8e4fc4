// Helpers shared by the distillation-loss kernels (cluster / tensor-memory / stream variants):
// approximate transcendentals, row walking on arbitrary element phases, online-softmax (m, z)
// pairs with integer maxima, tensor-memory register transfers.  Reference arithmetic:
// icv_src/icv_module.py:121-134 (KL), the HF shifted CE consumed at :94-98,115-117.
#pragma once

#include "licv_common.cuh"

namespace licv {
namespace {

constexpr int kMaxT = 256;            // threads per CTA: 256, or 128 (more, smaller CTAs per SM)
constexpr int kMaxWarps = kMaxT / 32;
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
__device__ __forceinline__ float rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
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
// mbarrier in this CTA's shared memory, completed by bytes that peers write with st.async
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// try_wait with a suspend-time hint: a waiting warp is parked by the hardware instead of
// re-issuing the probe (spinning warps outrank working ones in the issue arbiter)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LICV_KD_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra LICV_KD_DONE;\n"
        "bra LICV_KD_WAIT;\n"
        "LICV_KD_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(1000000u)
        : "memory");
}
// 16 bytes into slot `local` of CTA `rank`, counted on that CTA's mbarrier `bar`: a one-way
// message, no fence and no round trip (a release at cluster scope would wait for this warp's
// earlier gradient stores to drain)
__device__ __forceinline__ void st_async_f4(const void* local, const uint64_t* bar, uint32_t rank,
                                            float4 v) {
    uint32_t addr, mbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(mbar) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile(
        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];" ::
            "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(mbar)
        : "memory");
}
// -inf in the storage format, as a 16-byte vector (masks elements outside the row)
template <int DT> __device__ __forceinline__ uint32_t neg_inf_word();
template <> __device__ __forceinline__ uint32_t neg_inf_word<LICV_F32>() { return 0xff800000u; }
template <> __device__ __forceinline__ uint32_t neg_inf_word<LICV_BF16>() { return 0xff80ff80u; }
template <> __device__ __forceinline__ uint32_t neg_inf_word<LICV_F16>() { return 0xfc00fc00u; }

template <int DT>
__device__ __forceinline__ uint32_t load_bits(const char* row, int64_t j) {
    if (Fmt<DT>::kBytes == 4) return reinterpret_cast<const uint32_t*>(row)[j];
    return reinterpret_cast<const uint16_t*>(row)[j];
}

// 16 bytes at p, which is aligned to `align` bytes (16, 8, 4 or 2)
__device__ __forceinline__ uint4 load_vec_any(const char* p, int align) {
    if (align >= 16) return ld_stream(reinterpret_cast<const uint4*>(p));
    if (align >= 8) {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(p) + 1);
        return make_uint4(a.x, a.y, b.x, b.y);
    }
    if (align >= 4) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
        return make_uint4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
    }
    const uint16_t* q = reinterpret_cast<const uint16_t*>(p);
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = (uint32_t)__ldg(q + 2 * i) | ((uint32_t)__ldg(q + 2 * i + 1) << 16);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// elements of the vector that starts at row element j0 and lie outside [0, V) become -inf
template <int DT>
__device__ __forceinline__ uint4 mask_vec(uint4 v, int j0, int V) {
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
    const uint32_t ninf = neg_inf_word<DT>();
    if constexpr (Fmt<DT>::kBytes == 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if ((unsigned)(j0 + e) >= (unsigned)V) w[e] = ninf;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool lo = (unsigned)(j0 + 2 * i) < (unsigned)V, hi = (unsigned)(j0 + 2 * i + 1) < (unsigned)V;
            w[i] = ((lo ? w[i] : ninf) & 0xffffu) | ((hi ? w[i] : ninf) & 0xffff0000u);
        }
    }
    return v;
}

// This thread's NV vectors of one row: elements j0 + k * kStep .. + EPV - 1, k < NV.
// (row + j0 * EB) is aligned to `align` bytes.  load_row_vecs only ISSUES the loads;
// mask_row_vecs, called where the registers are first consumed, turns the elements outside [0, V)
// into -inf.
//
// Two rules, both measured (8192 x 32002 bf16 rows: 630 us without them, 540 us with):
//  * control flow is decided per WARP (from lane 0's element index), never per thread: a warp
//    whose lanes take different load paths into the same destination registers stalls at the
//    second path until the first path's loads have landed (write-after-write);
//  * nothing here may read a loaded register: the loads are the next row's prefetch, issued in
//    the middle of the current row, and the warp at a row end would wait a full HBM round trip
//    for them - every other warp of the row then waits for it at the next barrier.
//
// A warp that straddles a row end therefore loads whole vectors from clamped addresses.  With
// align == 16 such a vector is the 16-byte granule that holds the row's first or last element,
// so the bytes of it that lie outside the row are read (and discarded by the mask) too: they
// belong to the neighbouring row or, for the first / last row of a buffer, to the same 16-byte
// granule of the allocation - never to another page.  With align < 16 (a teacher row whose
// 16-byte phase differs from the student row's) the boundary vector is read element by element
// from clamped indices, nothing outside the row is touched, and that one path does consume its
// loads early.
template <int DT>
__device__ __forceinline__ bool warp_inside_row(int j0, int kStep, int nv, int V, int lane) {
    const int jw = j0 - lane * Fmt<DT>::kPerVec;           // lane 0's first element
    return jw >= 0 && jw + (nv - 1) * kStep + 32 * Fmt<DT>::kPerVec <= V;
}

template <int DT, int NV>
__device__ __forceinline__ void load_row_vecs(uint4 (&out)[NV], const char* row, int j0, int kStep, int V,
                                              int align, int lane) {
    constexpr int EPV = Fmt<DT>::kPerVec;
    constexpr int EB = Fmt<DT>::kBytes;
    const uint32_t ninf = neg_inf_word<DT>();
    const char* p0 = row + (int64_t)j0 * EB;
    if (warp_inside_row<DT>(j0, kStep, NV, V, lane)) {
#pragma unroll
        for (int k = 0; k < NV; ++k) out[k] = load_vec_any(p0 + (size_t)k * kStep * EB, align);
        return;
    }
    const int jw = j0 - lane * EPV;
    // first and last vector that overlap the row (same 16-byte phase as j0)
    const int jmin = -((-j0) & (EPV - 1));
    const int jmax = jmin + ((V - 1 - jmin) & ~(EPV - 1));
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int jk = j0 + k * kStep, jwk = jw + k * kStep;
        if (jwk >= 0 && jwk + 32 * EPV <= V) {
            out[k] = load_vec_any(p0 + (size_t)k * kStep * EB, align);
        } else if (jwk >= V || jwk + 32 * EPV <= 0) {
            out[k] = make_uint4(ninf, ninf, ninf, ninf);
        } else if (align >= 16) {
            const int jc = min(max(jk, jmin), jmax);
            out[k] = ld_stream(reinterpret_cast<const uint4*>(row + (int64_t)jc * EB));
        } else {
            uint32_t bits[EPV];
#pragma unroll
            for (int e = 0; e < EPV; ++e) bits[e] = load_bits<DT>(row, min(max(jk + e, 0), V - 1));
            if constexpr (EB == 4) {
                out[k] = make_uint4(bits[0], bits[1], bits[2], bits[3]);
            } else {
                out[k] = make_uint4(bits[0] | (bits[1] << 16), bits[2] | (bits[3] << 16),
                                    bits[4] | (bits[5] << 16), bits[6] | (bits[7] << 16));
            }
        }
    }
}

template <int DT, int NV>
__device__ __forceinline__ void mask_row_vecs(uint4 (&v)[NV], int j0, int kStep, int V, int lane) {
    if (warp_inside_row<DT>(j0, kStep, NV, V, lane)) return;
    const int jw = j0 - lane * Fmt<DT>::kPerVec;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int jwk = jw + k * kStep;
        if (jwk < 0 || jwk + 32 * Fmt<DT>::kPerVec > V) v[k] = mask_vec<DT>(v[k], j0 + k * kStep, V);
    }
}

template <int DT>
__device__ __forceinline__ void store_row_vec(char* row, int j0, int V, bool vec_ok, const float* f) {
    constexpr int EPV = Fmt<DT>::kPerVec;
    constexpr int EB = Fmt<DT>::kBytes;
    if (j0 >= V || j0 + EPV <= 0) return;
    if (vec_ok && j0 >= 0 && j0 + EPV <= V) {
        st_vec(reinterpret_cast<uint4*>(row + (int64_t)j0 * EB), pack<DT>(f));
        return;
    }
#pragma unroll
    for (int e = 0; e < EPV; ++e) {
        const int j = j0 + e;
        if (j >= 0 && j < V) store_elem<DT>(row, j, f[e]);
    }
}

// running maximum over the raw storage words of a vector
template <int DT> struct RawMax;
template <> struct RawMax<LICV_F32> {
    float m = -INFINITY;
    __device__ __forceinline__ void add(const uint4& v) {
        m = fmaxf(fmaxf(m, __uint_as_float(v.x)), fmaxf(__uint_as_float(v.y), __uint_as_float(v.z)));
        m = fmaxf(m, __uint_as_float(v.w));
    }
    __device__ __forceinline__ float get() const { return m; }
};
template <> struct RawMax<LICV_BF16> {
    __nv_bfloat162 m;
    __device__ __forceinline__ RawMax() {
        const uint32_t w = 0xff80ff80u;
        m = *reinterpret_cast<const __nv_bfloat162*>(&w);
    }
    __device__ __forceinline__ void add(const uint4& v) {
        const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
        m = __hmax2(__hmax2(m, p[0]), __hmax2(p[1], __hmax2(p[2], p[3])));
    }
    __device__ __forceinline__ float get() const {
        return fmaxf(__bfloat162float(m.x), __bfloat162float(m.y));
    }
};
template <> struct RawMax<LICV_F16> {
    __half2 m;
    __device__ __forceinline__ RawMax() {
        const uint32_t w = 0xfc00fc00u;
        m = *reinterpret_cast<const __half2*>(&w);
    }
    __device__ __forceinline__ void add(const uint4& v) {
        const __half2* p = reinterpret_cast<const __half2*>(&v);
        m = __hmax2(__hmax2(m, p[0]), __hmax2(p[1], __hmax2(p[2], p[3])));
    }
    __device__ __forceinline__ float get() const {
        return fmaxf(__half2float(m.x), __half2float(m.y));
    }
};

// ---------------------------------------------------------------------------------------------
// Cluster-wide reductions without a CTA barrier and without a serial section: every WARP sends
// its partial (16 bytes) straight to every CTA of the cluster (st.async, counted on the
// receiver's mbarrier); the receiver's warps each read the C x 8 partials (one or two per lane)
// and finish with a shuffle reduction.  A row's softmax statistics travel as online-softmax
// pairs (m, z): sum_j 2^(u_j) = z 2^m with m an INTEGER, so rescaling to a common maximum is an
// exponent-field operation on the ALU pipe - the joins cost no MUFU.
// ---------------------------------------------------------------------------------------------
constexpr int kNoMaxI = -(1 << 20);   // "no element": any real maximum wins, scale factor 0

// 2^k for an integer k <= 0 (0 below the normal range)
__device__ __forceinline__ float pow2i(int k) {
    const int e = k + 127;
    return __int_as_float((e > 0 ? e : 0) << 23);
}
struct MZ2 {
    int ms;
    float zs;
    int mt;
    float zt;
};
__device__ __forceinline__ MZ2 mz_join(const MZ2& a, const MZ2& b) {
    MZ2 r;
    r.ms = a.ms > b.ms ? a.ms : b.ms;
    r.mt = a.mt > b.mt ? a.mt : b.mt;
    r.zs = fmaf(a.zs, pow2i(a.ms - r.ms), b.zs * pow2i(b.ms - r.ms));
    r.zt = fmaf(a.zt, pow2i(a.mt - r.mt), b.zt * pow2i(b.mt - r.mt));
    return r;
}
// warp-wide join: integer maxima with one REDUX each, every lane rescales its own sum once to
// the common maximum, then two butterfly sums
__device__ __forceinline__ MZ2 mz_warp(MZ2 v) {
    const int ms = __reduce_max_sync(0xffffffffu, v.ms);
    const int mt = __reduce_max_sync(0xffffffffu, v.mt);
    float zs = v.zs * pow2i(v.ms - ms);
    float zt = v.zt * pow2i(v.mt - mt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        zs += __shfl_xor_sync(0xffffffffu, zs, o);
        zt += __shfl_xor_sync(0xffffffffu, zt, o);
    }
    return MZ2{ms, zs, mt, zt};
}

// (threads, 16-byte vectors per thread and row): 512 x 8, 768 x 6 or 1024 x 4; warp w owns TMEM lane
// quarter w % 4 and the column group w / 4 (8 NV columns wide)
__host__ __device__ constexpr int tmem_cols(int threads, int nv) {
    return (threads / 128) * nv * 8 <= 256 ? 256 : 512;     // allocations are powers of two
}

// 8 consecutive columns of this thread's TMEM lane <-> 8 registers.  tcgen05.ld is asynchronous:
// the registers are valid only after tcgen05.wait::ld, so the wait is written as an asm that
// "modifies" them - the compiler cannot move a use above it.
__device__ __forceinline__ void tmem_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
                   "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                   "+r"(r[7])
                 :
                 : "memory");
}

constexpr int kLabNone = -100;            // not a CE row
constexpr int kLabBad = 0x7fffffff;       // a label outside int32: out of range for any V

}  // namespace
}  // namespace licv
