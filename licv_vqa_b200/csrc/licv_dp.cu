// Data-parallel exchange of the ICV gradient fused with the optimizer step, over NVLink peer
// memory (one process per GPU, all on one NVSwitch node).
//
// The reference gets this from Lightning DDP: a bucketed NCCL all-reduce of icv_encoder.*.grad,
// a second collective for the logged scalars, clip_grad_norm_, AdamW (icv_src/icv_module.py:
// 171-209, config/trainer/ddp.yaml:5,7).  The message is tiny - L*d + L (+ 4 scalars) fp32 =
// 0.5 MB for idefics shapes - so the step is pure latency: a library all-reduce costs ~20-35 us
// of launch + protocol against < 1 us of wire time.  Here every rank keeps its flat gradient in
// a cudaMalloc'd region that all peers map (CUDA IPC); ONE kernel
//
//   phase 0  copies the local gradient into this rank's exchange slot and, when the last CTA is
//            done, publishes the step number into every peer's flag array (fence.sys, then relaxed stores),
//   phase 1  waits until every peer has published the same step (ld.acquire.sys, bounded),
//   phase 2  reads all ranks' slots over NVLink (plain 128-bit loads on mapped peer pointers),
//            sums them in RANK ORDER (every rank gets bit-identical sums), writes the sum back
//            as the gradient and accumulates its squared norm for the clip,
//
// and the existing AdamW kernel follows on the same stream.  Two slots alternate by step parity,
// so a slow peer still reading step k never races a fast peer writing step k + 1; the step
// counter lives in device memory, so the launch can be replayed from a CUDA graph.
#include <cmath>
#include <cstring>
#include <new>

#include "licv_common.cuh"

namespace licv {
int launch_adamw_after_norm(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                            int64_t n_vec, int64_t n_alpha, float lr_vec, float lr_alpha, float beta1,
                            float beta2, float eps, float weight_decay, int64_t step,
                            float grad_prescale, float max_grad_norm, float* norm_out,
                            void* workspace, const float* norm_partials, int n_partials,
                            cudaStream_t st);
}

constexpr int kDpMaxWorld = 16;
constexpr int kDpCtas = 128;   // all co-resident (no shared memory, 256 threads): one float4 per thread
                               // for idefics shapes, so phase 2 is ONE NVLink round trip
constexpr int kDpThreads = 256;

struct licv_dp_comm {
    int rank = 0, world = 1;
    int64_t n = 0;                 // floats per slot
    char* region[kDpMaxWorld] = {};   // region[rank] is local, the others are IPC mappings
    bool mapped[kDpMaxWorld] = {};
};

namespace {

struct Layout {
    int64_t slot_bytes, flags_off, ctl_off, partial_off, total;
};
Layout layout_of(int64_t n) {
    Layout L;
    L.slot_bytes = ((n * 4 + 255) / 256) * 256;
    L.flags_off = 2 * L.slot_bytes;
    L.ctl_off = L.flags_off + kDpMaxWorld * 8;
    L.partial_off = L.ctl_off + 64;
    L.total = L.partial_off + kDpCtas * 4;
    return L;
}

struct DpArgs {
    const char* peer[kDpMaxWorld];   // every rank's region
    char* local;
    float* grad;          // in: local gradient [n]; out: sum over ranks
    int64_t n, n_norm;    // floats exchanged, floats that enter the norm (the parameters)
    int64_t slot_bytes, flags_off, ctl_off;
    int rank, world;
    float prescale;       // 1 / world for the norm
    float* partial;       // [gridDim.x] per-CTA sums of squares (summed in fixed order later)
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() {
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}

__global__ void __launch_bounds__(kDpThreads) dp_exchange_kernel(DpArgs a) {
    __shared__ float slab[kDpThreads / 32];
    licv::pdl_launch_dependents();
    licv::pdl_wait();
    // control block of this rank: {step, ticket0, ticket1, error}
    unsigned long long* step_ctr = reinterpret_cast<unsigned long long*>(a.local + a.ctl_off);
    unsigned* ticket0 = reinterpret_cast<unsigned*>(a.local + a.ctl_off + 8);
    unsigned* ticket1 = ticket0 + 1;
    unsigned* error = ticket0 + 2;
    const unsigned long long step = *step_ctr + 1;
    const int64_t slot_off = (int64_t)(step & 1ull) * a.slot_bytes;
    const int64_t n4 = a.n / 4;   // a.n is padded to a multiple of 4 by the host side
    const int tid = threadIdx.x;

    // ---- phase 0: my gradient -> my slot; the last CTA tells every peer -----------------------
    {
        const float4* src = reinterpret_cast<const float4*>(a.grad);
        float4* dst = reinterpret_cast<float4*>(a.local + slot_off);
        for (int64_t i = (int64_t)blockIdx.x * kDpThreads + tid; i < n4; i += (int64_t)gridDim.x * kDpThreads)
            dst[i] = src[i];
    }
    // one system-scope fence per CTA (after the CTA barrier it covers every thread's stores), a
    // ticket, and the last CTA publishes: 128 fences instead of 32768
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        if (atomicAdd(ticket0, 1u) == gridDim.x - 1) {
            *ticket0 = 0u;
            // one fence orders every CTA's slot stores (each already fenced, observed through the
            // ticket) before the flags; the flag stores themselves are then relaxed and leave
            // back to back - a release store per peer would pay the fence `world - 1` times
            __threadfence_system();
            for (int p = 0; p < a.world; ++p) {
                if (p == a.rank) continue;
                unsigned long long* flag = reinterpret_cast<unsigned long long*>(
                    const_cast<char*>(a.peer[p]) + a.flags_off) + a.rank;
                st_relaxed_sys(flag, step);
            }
        }
    }
    // ---- phase 1: every peer has published this step (bounded: a dead peer must not hang us) --
    if (tid < a.world && tid != a.rank) {
        const unsigned long long* flag =
            reinterpret_cast<const unsigned long long*>(a.local + a.flags_off) + tid;
        const long long t0 = clock64();
        while (ld_relaxed_sys(flag) < step) {      // cheap probes, one acquire fence at the end
            if (clock64() - t0 > (1ll << 33)) {   // ~4 s
                *error = 1u;
                break;
            }
        }
        fence_acq_rel_sys();
    }
    __syncthreads();
    // ---- phase 2: sum all ranks' slots in rank order ------------------------------------------
    float sq = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * kDpThreads + tid; i < n4; i += (int64_t)gridDim.x * kDpThreads) {
        // every peer's value is requested before any is used: ONE NVLink round trip, not `world`
        // of them (a loop with a run-time trip count serialises load -> add -> load)
        float4 v[kDpMaxWorld];
#pragma unroll
        for (int r = 0; r < kDpMaxWorld; ++r)
            if (r < a.world) v[r] = reinterpret_cast<const float4*>(a.peer[r] + slot_off)[i];
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kDpMaxWorld; ++r)
            if (r < a.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
        reinterpret_cast<float4*>(a.grad)[i] = s;
        const float e[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i * 4 + k < a.n_norm) {
                const float x = e[k] * a.prescale;
                sq = fmaf(x, x, sq);
            }
        }
    }
    sq = licv::warp_sum(sq);
    if ((tid & 31) == 0) slab[tid >> 5] = sq;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < kDpThreads / 32; ++w) t += slab[w];
        a.partial[blockIdx.x] = t;     // no atomics: the norm must be bit-identical on every rank
        __threadfence();
        if (atomicAdd(ticket1, 1u) == gridDim.x - 1) {   // the last CTA closes the step
            *ticket1 = 0u;
            *step_ctr = step;
        }
    }
}

}  // namespace

extern "C" int64_t licv_dp_region_bytes(int64_t n_floats) {
    if (n_floats < 0) return 0;
    return layout_of((n_floats + 3) / 4 * 4).total;
}

extern "C" int licv_dp_region_alloc(int64_t n_floats, void** region, void* ipc_handle_64) {
    if (!region || !ipc_handle_64) return LICV_ERR_NULL_POINTER;
    if (n_floats <= 0) return LICV_ERR_BAD_ARGUMENT;
    if (licv::device_info().status != LICV_OK) return licv::device_info().status;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const int64_t bytes = licv_dp_region_bytes(n_floats);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return (int)e;
    }
    std::memcpy(ipc_handle_64, &h, 64);
    *region = p;
    return LICV_OK;
}

extern "C" int licv_dp_comm_create(licv_dp_comm** out, int rank, int world, void* region,
                                   const void* all_handles, int64_t n_floats) {
    if (!out || !region) return LICV_ERR_NULL_POINTER;
    if (world < 1 || world > kDpMaxWorld || rank < 0 || rank >= world || n_floats <= 0)
        return LICV_ERR_BAD_ARGUMENT;
    if (world > 1 && !all_handles) return LICV_ERR_NULL_POINTER;
    auto* c = new (std::nothrow) licv_dp_comm();
    if (!c) return LICV_ERR_BAD_ARGUMENT;
    c->rank = rank;
    c->world = world;
    c->n = (n_floats + 3) / 4 * 4;
    c->region[rank] = static_cast<char*>(region);
    for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(all_handles) + (size_t)p * 64, 64);
        void* ptr = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < world; ++q)
                if (c->mapped[q]) cudaIpcCloseMemHandle(c->region[q]);
            delete c;
            return (int)e;
        }
        c->region[p] = static_cast<char*>(ptr);
        c->mapped[p] = true;
    }
    *out = c;
    return LICV_OK;
}

extern "C" int licv_dp_comm_destroy(licv_dp_comm* c) {
    if (!c) return LICV_OK;
    cudaDeviceSynchronize();
    for (int p = 0; p < c->world; ++p)
        if (c->mapped[p]) cudaIpcCloseMemHandle(c->region[p]);
    if (c->region[c->rank]) cudaFree(c->region[c->rank]);
    delete c;
    return LICV_OK;
}

extern "C" int licv_dp_comm_error(licv_dp_comm* c) {
    if (!c) return LICV_ERR_NULL_POINTER;
    unsigned err = 0;
    const Layout L = layout_of(c->n);
    if (cudaMemcpy(&err, c->region[c->rank] + L.ctl_off + 16, 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return (int)cudaGetLastError();
    return (int)err;
}

extern "C" int licv_dp_allreduce_adamw(licv_dp_comm* c, float* param, float* grad, float* exp_avg,
                                       float* exp_avg_sq, int64_t n_vec, int64_t n_alpha,
                                       int64_t n_extra, float lr_vec, float lr_alpha, float beta1,
                                       float beta2, float eps, float weight_decay, int64_t step,
                                       float max_grad_norm, float* norm_out, void* workspace,
                                       licv_stream_t stream) {
    if (!c) return LICV_ERR_NULL_POINTER;
    if (licv::device_info().status != LICV_OK) return licv::device_info().status;
    if (n_vec < 0 || n_alpha < 0 || n_extra < 0 || step < 1) return LICV_ERR_BAD_ARGUMENT;
    const int64_t n = n_vec + n_alpha + n_extra;
    if ((n + 3) / 4 * 4 != c->n) return LICV_ERR_BAD_ARGUMENT;   // the buffer the comm was made for
    if (!param || !grad || !exp_avg || !exp_avg_sq || !workspace) return LICV_ERR_NULL_POINTER;
    if (!licv::aligned16(grad)) return LICV_ERR_MISALIGNED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const Layout L = layout_of(c->n);
    DpArgs a;
    for (int p = 0; p < kDpMaxWorld; ++p) a.peer[p] = p < c->world ? c->region[p] : nullptr;
    a.local = c->region[c->rank];
    a.grad = grad;
    a.n = c->n;            // grad must have room for the padding (n rounded up to 4 floats)
    a.n_norm = n_vec + n_alpha;
    a.slot_bytes = L.slot_bytes;
    a.flags_off = L.flags_off;
    a.ctl_off = L.ctl_off;
    a.rank = c->rank;
    a.world = c->world;
    a.prescale = 1.0f / (float)c->world;
    a.partial = reinterpret_cast<float*>(c->region[c->rank] + L.partial_off);
    if (int rc = licv::launch_pdl(dp_exchange_kernel, dim3(kDpCtas), dim3(kDpThreads), 0, st, a)) return rc;
    return licv::launch_adamw_after_norm(param, grad, exp_avg, exp_avg_sq, n_vec, n_alpha, lr_vec,
                                         lr_alpha, beta1, beta2, eps, weight_decay, step,
                                         1.0f / (float)c->world, max_grad_norm, norm_out, workspace,
                                         a.partial, kDpCtas, st);
}
