// On-chip throughput probes for the loss kernel's budget (B200, one 512-thread CTA per SM):
// MUFU, tensor-memory loads / stores (tcgen05.ld / st 32x32b.x8), shared-memory 128-bit loads,
// packed fp32x2 FMAs, and MUFU + FFMA2 issued together.  Prints bytes or ops per clock per SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench tools/ubench.cu && tools/bin/ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int kT = 512;
constexpr int kIters = 2048;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(kT, 1) probe(float* out, long long* clocks) {
    extern __shared__ __align__(16) float4 sm[];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tcol = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    for (int i = tid; i < 8 * kT; i += kT) sm[i] = make_float4(i * 1e-3f, 1.f, 2.f, 3.f);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = tid * 1e-4f + e;
    uint32_t r[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    // fill tensor memory so that loads return defined data
    for (int k = 0; k < 8; ++k)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tcol + k * 8),
                     "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    float2 acc[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        if (MODE == 0 || MODE == 5) {          // 8 MUFU.EX2
#pragma unroll
            for (int e = 0; e < 8; ++e) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[e]));
        }
        if (MODE == 1) {                       // LDTM x8 (two in flight)
            uint32_t a[8], b[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7])
                         : "r"(tcol + (it & 7) * 8) : "memory");
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7])
                         : "r"(tcol + ((it + 1) & 7) * 8) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] ^= a[e] + b[e];
        }
        if (MODE == 2) {                       // STTM x8 x2
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tcol + (it & 7) * 8),
                         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tcol + ((it + 1) & 7) * 8),
                         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        if (MODE == 3) {                       // 4 x LDS.128
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float4 q;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                             : "r"(smem_u32(sm + ((it + k) & 7) * kT + tid)));
                v[2 * k] += q.x + q.z;
                v[2 * k + 1] += q.y + q.w;
            }
        }
        if (MODE == 4 || MODE == 5) {          // 16 FFMA2 (32 fp32 FMAs)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    acc[h] = __ffma2_rn(acc[h], make_float2(1.0001f, 0.9999f), make_float2(v[h], v[h + 4]));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[e] + (float)r[e];
#pragma unroll
    for (int h = 0; h < 4; ++h) s += acc[h].x + acc[h].y;
    out[blockIdx.x * kT + tid] = s;
    if (tid == 0) clocks[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(s_tmem) : "memory");
}

template <int MODE>
int run(const char* name, double units_per_iter_per_thread, const char* unit) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float* out;
    long long* clk;
    CK(cudaMalloc(&out, sizeof(float) * sms * kT));
    CK(cudaMalloc(&clk, sizeof(long long) * sms));
    const size_t smem = 8 * kT * sizeof(float4);
    CK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; ++rep) probe<MODE><<<sms, kT, smem>>>(out, clk);
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)h[i];
    mean /= sms;
    printf("{\"probe\": \"%s\", \"clk_per_iter\": %.2f, \"per_clk_per_sm\": %.2f, \"unit\": \"%s\"}\n", name,
           mean / kIters, units_per_iter_per_thread * kT * kIters / mean, unit);
    cudaFree(out);
    cudaFree(clk);
    return 0;
}

int main() {
    if (run<0>("mufu_ex2", 8, "MUFU")) return 1;
    if (run<1>("tmem_ld_x8", 64, "bytes")) return 1;
    if (run<2>("tmem_st_x8", 64, "bytes")) return 1;
    if (run<3>("lds128", 64, "bytes")) return 1;
    if (run<4>("ffma2", 32, "fp32 FMA")) return 1;
    if (run<5>("mufu+ffma2 (8 MUFU + 32 FMA per iter)", 8, "MUFU")) return 1;
    return 0;
}
