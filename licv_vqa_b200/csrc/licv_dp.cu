// Data-parallel exchange of the ICV gradient fused with the optimizer step, over NVLink peer
// memory (one process per GPU, all on one NVSwitch node).
//
// The reference gets this from Lightning DDP: a bucketed NCCL all-reduce of icv_encoder.*.grad,
// a second collective for the logged scalars, clip_grad_norm_, AdamW (icv_src/icv_module.py:
// 171-209, config/trainer/ddp.yaml:5,7).  The message is tiny - L*d + L (+ 4 scalars) fp32 =
// 0.5 MB for idefics shapes - so the step is pure latency: a library all-reduce costs ~20-35 us
// of launch + protocol against < 1 us of wire time.  Here every rank owns a cudaMalloc'd region
// that all peers map (CUDA IPC), and ONE kernel does the whole exchange with no flag, fence,
// ticket or barrier on the way:
//
//   push    each thread packs two gradient floats, each with the step number, into one 16-byte
//           packet {f0, tag, f1, tag} and stores it (st.v2.b64) into the slot [step parity][this
//           rank] of EVERY peer's region (posted writes over NVLink, one per peer, back to back);
//   gather  the same thread reads the packets the peers pushed at the same index into its own
//           region - local memory - until each carries this step's tag, sums the values in
//           RANK ORDER (every rank gets bit-identical sums), writes the sum back as the
//           gradient and accumulates its squared norm for the clip,
//
// and the existing AdamW kernel follows on the same stream.  Every 8-byte half of a packet carries
// its own copy of the tag next to the float it guards - NCCL's LL protocol: the PTX memory model
// treats a vector access as one scalar access per element, an aligned 8-byte element is
// single-copy atomic, and the reader accepts a packet only when BOTH halves show this step's tag
// - so a half is either absent or complete and the latency of the exchange is ONE one-way trip
// instead of copy -> system fence -> flag -> flag seen -> remote read round trip (measured, two
// ranks, CTA 0: 1.4 + 3.5 + 6.5 + 4.3 us for those four).  Two slots per source alternate by step
// parity: a peer can only push step k + 2 after it has finished step k + 1, which needed this
// rank's step-k+1 packets, which are pushed after this rank finished reading step k.  The step
// counter lives in device memory, so the launch can be replayed from a CUDA graph.
//
// Two forms of the kernel share packets, slots, tags and error handling.  The all-to-all form is
// the one described above.  The OWNER form (dp_exchange_owner_kernel, the default: see dp_algo)
// gives every packet an owning rank - contributions go to the owner only, the owner pushes the
// rank-order sum to everybody - which moves 2 (world - 1) / world of the buffer per GPU instead
// of (world - 1) x it, for one more one-way trip; it measured faster at 2 and at 8 ranks.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "licv_common.cuh"

namespace licv {
int launch_adamw_after_norm(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                            int64_t n_vec, int64_t n_alpha, float lr_vec, float lr_alpha, float beta1,
                            float beta2, float eps, float weight_decay, int64_t step,
                            float grad_prescale, float max_grad_norm, float* norm_out,
                            void* workspace, const float* norm_partials, int n_partials,
                            const unsigned* skip_if_set, cudaStream_t st);
}

constexpr int kDpMaxWorld = 16;
constexpr int kDpThreads = 256;
constexpr int kDpMaxCtas = 512;   // one packet per thread for idefics shapes (171 CTAs), co-resident

struct licv_dp_comm {
    int rank = 0, world = 1;
    int64_t n = 0;                 // floats per slot
    char* region[kDpMaxWorld] = {};   // region[rank] is local, the others are IPC mappings
    bool mapped[kDpMaxWorld] = {};
};

namespace {

// region: [parity 2][source rank kDpMaxWorld][packets * 16 B] | control block | per-CTA partials
struct Layout {
    int64_t n_packets, slot_bytes, ctl_off, partial_off, total;
};
Layout layout_of(int64_t n) {
    Layout L;
    L.n_packets = (n + 1) / 2;
    L.slot_bytes = ((L.n_packets * 16 + 255) / 256) * 256;
    L.ctl_off = 2 * kDpMaxWorld * L.slot_bytes;
    L.partial_off = L.ctl_off + 64;
    L.total = L.partial_off + kDpMaxCtas * 4;
    return L;
}
int grid_of(const Layout& L) {
    const int64_t want = (L.n_packets + kDpThreads - 1) / kDpThreads;
    const int64_t cap = 2 * (int64_t)licv::device_info().sm_count;
    int64_t g = want < cap ? want : cap;
    if (g > kDpMaxCtas) g = kDpMaxCtas;
    return (int)(g < 1 ? 1 : g);
}

struct DpArgs {
    char* region[kDpMaxWorld];   // every rank's region (region[rank] is local)
    float* grad;          // in: local gradient [n]; out: sum over ranks
    int64_t n, n_norm;    // floats exchanged, floats that enter the norm (the parameters)
    int64_t n_packets, slot_bytes, ctl_off;
    int rank, world;
    float prescale;       // 1 / world for the norm
    float* partial;       // [gridDim.x] per-CTA sums of squares (summed in fixed order later)
    int64_t slice;        // owner form: packets per owning rank (owner of packet i = i / slice)
};

// one packet = two 8-byte elements {float bits | tag << 32}: each element is one scalar access
__device__ __forceinline__ void st_packet(void* p, unsigned long long lo, unsigned long long hi) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void ld_packet(const void* p, unsigned long long& lo, unsigned long long& hi) {
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long half_packet(float f, unsigned tag) {
    return (unsigned long long)__float_as_uint(f) | ((unsigned long long)tag << 32);
}
__device__ __forceinline__ bool packet_ok(unsigned long long lo, unsigned long long hi, unsigned tag) {
    return (unsigned)(lo >> 32) == tag && (unsigned)(hi >> 32) == tag;
}

#ifdef LICV_TRACE
// debug build: globaltimer at the phase boundaries of CTA 0, one record of 8 per launch (ring of 256)
__device__ long long g_dp_trace[256 * 8];
__device__ unsigned g_dp_trace_n;
__device__ __forceinline__ long long dp_now() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define DP_TP(k)                                                                   \
    do {                                                                           \
        if (blockIdx.x == 0 && threadIdx.x == 0) g_dp_trace[(trace_slot & 255) * 8 + (k)] = dp_now(); \
    } while (0)
#else
#define DP_TP(k) do { } while (0)
#endif

__global__ void __launch_bounds__(kDpThreads) dp_exchange_kernel(DpArgs a) {
    __shared__ float slab[kDpThreads / 32];
#ifdef LICV_TRACE
    const unsigned trace_slot = g_dp_trace_n;
    DP_TP(0);
#endif
    licv::pdl_launch_dependents();
    licv::pdl_wait();
    DP_TP(1);
    // control block of this rank: {step, ticket, -, error}
    char* local = a.region[a.rank];
    unsigned long long* step_ctr = reinterpret_cast<unsigned long long*>(local + a.ctl_off);
    unsigned* ticket = reinterpret_cast<unsigned*>(local + a.ctl_off + 8);
    unsigned* error = ticket + 2;
    const unsigned long long step = *step_ctr + 1;
    const unsigned tag = (unsigned)step;          // the regions start zeroed and steps count from 1
    const int64_t parity_off = (int64_t)(step & 1ull) * kDpMaxWorld * a.slot_bytes;
    const int64_t my_slot = parity_off + (int64_t)a.rank * a.slot_bytes;
    const int tid = threadIdx.x;

    float sq = 0.f;
    bool dead = false;        // a peer never delivered: this rank must not apply the step
    for (int64_t i = (int64_t)blockIdx.x * kDpThreads + tid; i < a.n_packets;
         i += (int64_t)gridDim.x * kDpThreads) {
        float f[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) f[k] = i * 2 + k < a.n ? a.grad[i * 2 + k] : 0.f;
        // ---- push: one posted 16-byte write per peer ------------------------------------------
        const unsigned long long plo = half_packet(f[0], tag), phi = half_packet(f[1], tag);
#pragma unroll
        for (int p = 0; p < kDpMaxWorld; ++p)
            if (p < a.world && p != a.rank) st_packet(a.region[p] + my_slot + i * 16, plo, phi);
        DP_TP(2);
        // ---- gather: the peers' packets at the same index, from LOCAL memory ------------------
        unsigned long long vlo[kDpMaxWorld], vhi[kDpMaxWorld];
#pragma unroll
        for (int p = 0; p < kDpMaxWorld; ++p)
            if (p < a.world && p != a.rank)
                ld_packet(local + parity_off + (int64_t)p * a.slot_bytes + i * 16, vlo[p], vhi[p]);
#pragma unroll
        for (int p = 0; p < kDpMaxWorld; ++p) {
            if (p < a.world && p != a.rank && !packet_ok(vlo[p], vhi[p], tag)) {
                const long long t0 = clock64();
                do {                                   // bounded: a dead peer must not hang us
                    ld_packet(local + parity_off + (int64_t)p * a.slot_bytes + i * 16, vlo[p], vhi[p]);
                    if (clock64() - t0 > (1ll << 33)) {   // ~4 s
                        *error = 1u;
                        dead = true;
                        break;
                    }
                } while (!packet_ok(vlo[p], vhi[p], tag));
            }
        }
        DP_TP(3);
        if (dead) continue;    // keep the local gradient; the optimizer kernel is skipped (error flag)
        float s[2] = {0.f, 0.f};
#pragma unroll
        for (int p = 0; p < kDpMaxWorld; ++p) {
            if (p < a.world) {
                if (p == a.rank) {
                    s[0] += f[0]; s[1] += f[1];
                } else {
                    s[0] += __uint_as_float((unsigned)vlo[p]);
                    s[1] += __uint_as_float((unsigned)vhi[p]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (i * 2 + k < a.n) a.grad[i * 2 + k] = s[k];
            if (i * 2 + k < a.n_norm) {
                const float x = s[k] * a.prescale;
                sq = fmaf(x, x, sq);
            }
        }
    }
    DP_TP(4);
    sq = licv::warp_sum(sq);
    if ((tid & 31) == 0) slab[tid >> 5] = sq;
    __syncthreads();
    if (tid == 0) {
#ifdef LICV_TRACE
        if (blockIdx.x == 0) { DP_TP(5); g_dp_trace_n = trace_slot + 1; }
#endif
        float t = 0.f;
        for (int w = 0; w < kDpThreads / 32; ++w) t += slab[w];
        a.partial[blockIdx.x] = t;     // no atomics: the norm must be bit-identical on every rank
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {   // the last CTA closes the step
            *ticket = 0u;
            *step_ctr = step;
        }
    }
}

// The OWNER form of the same exchange, for many ranks.  The all-to-all form above moves
// (world - 1) x the whole packet buffer out of and into every GPU (7.3 MB each way at world = 8:
// bandwidth, ~13 us of every step); here packet i has an owning rank (i / slice): every rank pushes
// its packet to the OWNER only, the owner adds the world's packets in rank order and pushes the sum
// to every peer - 2 (world - 1) / world of the buffer per GPU and direction (1.8 MB at world = 8)
// for one more one-way trip.  The sums land in the owner's source slot at the owner's slice, the
// contributions in the sender's source slot at the receiver's slice: positions never collide, and
// the same two parity slots serve (a rank two steps ahead implies every owner finished the step in
// between, which implies every rank finished this one).  Same packets, tags, bounded waits, error
// word and rank-order sums: every replica ends with the same bits as with the all-to-all form.
__global__ void __launch_bounds__(kDpThreads) dp_exchange_owner_kernel(DpArgs a) {
    __shared__ float slab[kDpThreads / 32];
    licv::pdl_launch_dependents();
    licv::pdl_wait();
    char* local = a.region[a.rank];
    unsigned long long* step_ctr = reinterpret_cast<unsigned long long*>(local + a.ctl_off);
    unsigned* ticket = reinterpret_cast<unsigned*>(local + a.ctl_off + 8);
    unsigned* error = ticket + 2;
    const unsigned long long step = *step_ctr + 1;
    const unsigned tag = (unsigned)step;
    const int64_t parity_off = (int64_t)(step & 1ull) * kDpMaxWorld * a.slot_bytes;
    const int64_t my_slot = parity_off + (int64_t)a.rank * a.slot_bytes;
    const int tid = threadIdx.x;

    // bounded wait for one packet of this step (false: the sender never delivered)
    auto await = [&](const char* p, unsigned long long& lo, unsigned long long& hi) -> bool {
        ld_packet(p, lo, hi);
        if (packet_ok(lo, hi, tag)) return true;
        const long long t0 = clock64();
        do {
            ld_packet(p, lo, hi);
            if (clock64() - t0 > (1ll << 33)) {   // ~4 s
                *error = 1u;
                return false;
            }
        } while (!packet_ok(lo, hi, tag));
        return true;
    };

    float sq = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * kDpThreads + tid; i < a.n_packets;
         i += (int64_t)gridDim.x * kDpThreads) {
        float f[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) f[k] = i * 2 + k < a.n ? a.grad[i * 2 + k] : 0.f;
        const int own = (int)(i / a.slice);
        float s[2] = {0.f, 0.f};
        if (own != a.rank) {
            // contribute to the owner, then take the owner's sum from LOCAL memory
            char* owner_region = a.region[0];
#pragma unroll
            for (int p = 1; p < kDpMaxWorld; ++p)
                if (p == own) owner_region = a.region[p];
            st_packet(owner_region + my_slot + i * 16, half_packet(f[0], tag), half_packet(f[1], tag));
            unsigned long long lo, hi;
            if (!await(local + parity_off + (int64_t)own * a.slot_bytes + i * 16, lo, hi)) continue;
            s[0] = __uint_as_float((unsigned)lo);
            s[1] = __uint_as_float((unsigned)hi);
        } else {
            unsigned long long vlo[kDpMaxWorld], vhi[kDpMaxWorld];
#pragma unroll
            for (int p = 0; p < kDpMaxWorld; ++p)
                if (p < a.world && p != a.rank)
                    ld_packet(local + parity_off + (int64_t)p * a.slot_bytes + i * 16, vlo[p], vhi[p]);
            bool dead = false;
#pragma unroll
            for (int p = 0; p < kDpMaxWorld; ++p)
                if (p < a.world && p != a.rank && !packet_ok(vlo[p], vhi[p], tag) && !dead)
                    dead = !await(local + parity_off + (int64_t)p * a.slot_bytes + i * 16, vlo[p], vhi[p]);
            if (dead) continue;    // nobody gets this sum: every rank raises its error word
#pragma unroll
            for (int p = 0; p < kDpMaxWorld; ++p) {
                if (p < a.world) {
                    if (p == a.rank) {
                        s[0] += f[0]; s[1] += f[1];
                    } else {
                        s[0] += __uint_as_float((unsigned)vlo[p]);
                        s[1] += __uint_as_float((unsigned)vhi[p]);
                    }
                }
            }
            const unsigned long long plo = half_packet(s[0], tag), phi = half_packet(s[1], tag);
#pragma unroll
            for (int p = 0; p < kDpMaxWorld; ++p)
                if (p < a.world && p != a.rank) st_packet(a.region[p] + my_slot + i * 16, plo, phi);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (i * 2 + k < a.n) a.grad[i * 2 + k] = s[k];
            if (i * 2 + k < a.n_norm) {
                const float x = s[k] * a.prescale;
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
        a.partial[blockIdx.x] = t;
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *ticket = 0u;
            *step_ctr = step;
        }
    }
}

// which form: LICV_DP_ALGO = all | owner | auto (auto: the owner form from kDpOwnerFromWorld ranks).
// Measured under the driver's command (20 steps): N = 2 0.2914 (owner) vs 0.2929 ms (all), N = 8
// 0.2968 vs 0.3062 ms - the posted 16-byte writes, not the extra trip, are what the exchange costs.
constexpr int kDpOwnerFromWorld = 2;
int dp_algo() {
    static const int algo = [] {
        const char* v = std::getenv("LICV_DP_ALGO");
        if (v && std::strcmp(v, "all") == 0) return 1;
        if (v && std::strcmp(v, "owner") == 0) return 2;
        return 0;
    }();
    return algo;
}

}  // namespace

#ifdef LICV_TRACE
extern "C" int licv_debug_read_dp_trace(long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_dp_trace, sizeof(long long) * n);
}
#endif

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
    // peers write into this region as soon as they hold the handle: the zero tags must be there
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
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

extern "C" int licv_dp_region_free(void* region) {
    if (!region) return LICV_OK;
    return (int)cudaFree(region);
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

extern "C" int licv_dp_comm_reset_error(licv_dp_comm* c) {
    if (!c) return LICV_ERR_NULL_POINTER;
    const Layout L = layout_of(c->n);
    if (cudaMemset(c->region[c->rank] + L.ctl_off + 16, 0, 4) != cudaSuccess) return (int)cudaGetLastError();
    return LICV_OK;
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
    for (int p = 0; p < kDpMaxWorld; ++p) a.region[p] = p < c->world ? c->region[p] : nullptr;
    a.grad = grad;
    a.n = c->n;            // grad must have room for the padding (n rounded up to 4 floats)
    a.n_norm = n_vec + n_alpha;
    a.n_packets = L.n_packets;
    a.slot_bytes = L.slot_bytes;
    a.ctl_off = L.ctl_off;
    a.rank = c->rank;
    a.world = c->world;
    a.prescale = 1.0f / (float)c->world;
    a.partial = reinterpret_cast<float*>(c->region[c->rank] + L.partial_off);
    a.slice = (L.n_packets + c->world - 1) / c->world;
    const int grid = grid_of(L);
    const int algo = dp_algo();
    const bool owner_form = c->world > 1 && L.n_packets >= c->world &&
                            (algo == 2 || (algo == 0 && c->world >= kDpOwnerFromWorld));
    if (int rc = owner_form
                     ? licv::launch_pdl(dp_exchange_owner_kernel, dim3(grid), dim3(kDpThreads), 0, st, a)
                     : licv::launch_pdl(dp_exchange_kernel, dim3(grid), dim3(kDpThreads), 0, st, a))
        return rc;
    return licv::launch_adamw_after_norm(param, grad, exp_avg, exp_avg_sq, n_vec, n_alpha, lr_vec,
                                         lr_alpha, beta1, beta2, eps, weight_decay, step,
                                         1.0f / (float)c->world, max_grad_norm, norm_out, workspace,
                                         a.partial, grid,
                                         reinterpret_cast<const unsigned*>(c->region[c->rank] + L.ctl_off + 16),
                                         st);
}
