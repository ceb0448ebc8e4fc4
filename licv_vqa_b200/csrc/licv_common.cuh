// Shared device helpers for the L-ICV hot-path kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <mutex>

#include "licv_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "liblicv_b200 is written for sm_100a (B200) only"
#endif

namespace licv {

constexpr int kWarp = 32;

// ---------------------------------------------------------------------------------------------
// storage formats
// ---------------------------------------------------------------------------------------------
template <int DT> struct Fmt;
template <> struct Fmt<LICV_F32> {
    static constexpr int kBytes = 4;
    static constexpr int kPerVec = 4;  // elements per 16-byte vector
    __device__ static __forceinline__ float round(float x) { return x; }
};
template <> struct Fmt<LICV_BF16> {
    static constexpr int kBytes = 2;
    static constexpr int kPerVec = 8;
    __device__ static __forceinline__ float round(float x) {
        return __bfloat162float(__float2bfloat16_rn(x));
    }
};
template <> struct Fmt<LICV_F16> {
    static constexpr int kBytes = 2;
    static constexpr int kPerVec = 8;
    __device__ static __forceinline__ float round(float x) {
        return __half2float(__float2half_rn(x));
    }
};

// unpack one 16-byte vector into fp32 lanes
template <int DT> __device__ __forceinline__ void unpack(const uint4& v, float* f);
template <> __device__ __forceinline__ void unpack<LICV_F32>(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ void unpack<LICV_BF16>(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <> __device__ __forceinline__ void unpack<LICV_F16>(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 p = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        f[2 * i] = p.x;
        f[2 * i + 1] = p.y;
    }
}

// pack fp32 lanes into one 16-byte vector (round-to-nearest-even)
template <int DT> __device__ __forceinline__ uint4 pack(const float* f);
template <> __device__ __forceinline__ uint4 pack<LICV_F32>(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
}
template <> __device__ __forceinline__ uint4 pack<LICV_BF16>(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&p);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ uint4 pack<LICV_F16>(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 p = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&p);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// scalar element load/store
template <int DT> __device__ __forceinline__ float load_elem(const void* base, int64_t i);
template <> __device__ __forceinline__ float load_elem<LICV_F32>(const void* b, int64_t i) {
    return static_cast<const float*>(b)[i];
}
template <> __device__ __forceinline__ float load_elem<LICV_BF16>(const void* b, int64_t i) {
    return __bfloat162float(static_cast<const __nv_bfloat16*>(b)[i]);
}
template <> __device__ __forceinline__ float load_elem<LICV_F16>(const void* b, int64_t i) {
    return __half2float(static_cast<const __half*>(b)[i]);
}
template <int DT> __device__ __forceinline__ void store_elem(void* base, int64_t i, float x);
template <> __device__ __forceinline__ void store_elem<LICV_F32>(void* b, int64_t i, float x) {
    static_cast<float*>(b)[i] = x;
}
template <> __device__ __forceinline__ void store_elem<LICV_BF16>(void* b, int64_t i, float x) {
    static_cast<__nv_bfloat16*>(b)[i] = __float2bfloat16_rn(x);
}
template <> __device__ __forceinline__ void store_elem<LICV_F16>(void* b, int64_t i, float x) {
    static_cast<__half*>(b)[i] = __float2half_rn(x);
}

// ---------------------------------------------------------------------------------------------
// 128-bit global memory access
// ---------------------------------------------------------------------------------------------
// streaming read of data that is touched once: read-only path, do not allocate in L1
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// plain coherent read (used when the destination may alias the source)
__device__ __forceinline__ uint4 ld_plain(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ void st_vec(uint4* p, const uint4& v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
// fp32x4 reduction into global memory (one REDG.F32x4, no return value)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// device properties, cached per device
struct DeviceInfo {
    int sm_count = 0;
    int cc_major = 0;
    int cc_minor = 0;
    int status = LICV_ERR_NO_DEVICE;
};
const DeviceInfo& device_info();

// Launch constants that depend on the device (occupancy, "this function's shared-memory limit was
// raised"): computed once PER DEVICE, thread-safely - the entry points are re-entrant, a process may
// drive several GPUs, and autograd calls the backward ops from its own per-device threads.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return (d < 0 || d >= kMaxDevices) ? 0 : d;
}
template <typename T>
class PerDevice {
    T v_[kMaxDevices] = {};
    std::once_flag once_[kMaxDevices];

public:
    template <typename F>
    const T& get(F&& make) {
        const int d = current_device();
        std::call_once(once_[d], [&] { v_[d] = make(); });
        return v_[d];
    }
};

// Upper bound on the CTAs of the next injection launches of this thread (0 = none).  The
// host-buffer entry points set it: when the operands live in host memory the link, not the SMs,
// is the bottleneck, and a SMALL persistent grid lets each CTA's loads of token t+1 overlap its
// stores of token t, i.e. keeps both directions of the link busy inside one kernel.
extern thread_local int tl_grid_cap;
struct GridCapScope {
    int saved;
    explicit GridCapScope(int cap) : saved(tl_grid_cap) { tl_grid_cap = cap; }
    ~GridCapScope() { tl_grid_cap = saved; }
};

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch (PDL)
// ---------------------------------------------------------------------------------------------
// The hot kernels are launched with the programmatic-stream-serialization attribute: their CTAs
// may be scheduled, and run their prologue (barrier init, index arithmetic), while the previous
// kernel in the stream is still draining; `pdl_wait()` is the point after which the previous
// kernel's memory operations are complete and visible - it must precede the first access to
// global memory.  `pdl_launch_dependents()` lets the NEXT kernel begin its own prologue.  At the
// training shape (2-6 MB per launch) the gaps between launches are a third of the step.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// L2 prefetch hints.  They may run BEFORE pdl_wait(): a prefetch observes no value (L2 is the point
// of coherence - a line the previous kernel writes afterwards is simply updated in place), so at
// the latency-bound launch sizes the HBM round trip of a kernel's first loads overlaps the tail of
// its predecessor instead of starting behind griddepcontrol.wait.
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {   // bytes % 16 == 0
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
inline bool pdl_enabled() {
    static const bool on = [] {
        const char* v = std::getenv("LICV_PDL");
        return !(v && v[0] == '0');
    }();
    return on;
}
// fills `attr` (room for 2) for a launch with an optional cluster dimension; returns the count
inline int launch_attrs(cudaLaunchAttribute* attr, int cluster) {
    int n = 0;
    if (cluster > 0) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    return n;
}

// <<<grid, block, smem, stream>>> with the PDL attribute; the kernel must call pdl_wait() before
// its first global-memory access
template <typename... KArgs, typename... Args>
inline int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                      Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = launch_attrs(attr, 0);
    return (int)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace licv
