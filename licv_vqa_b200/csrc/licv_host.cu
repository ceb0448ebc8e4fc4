// Host-buffer entry points: the calls a host-side plugin makes when its tensors live in host
// memory.  A session owns n_slots (device scratch, stream) pairs; call k runs entirely on slot
// k % n_slots.
//
// Two data paths:
//  * ZERO-COPY (pinned / registered host buffers, the normal case): the kernels are launched
//    directly on the device view of the host buffers - hidden states and logits are read over
//    PCIe by the kernel's own 128-bit loads / TMA bulk copies and results are written straight
//    back to host memory, so both directions of the link are busy at once inside ONE kernel and
//    nothing is staged in HBM.  Only the few-KB operands that many CTAs re-read (shift vector,
//    row lists) and the atomically accumulated d_shift go through device scratch.
//  * STAGED (pageable host memory, or LICV_HOST_ZERO_COPY=0): cudaMemcpyAsync into the slot's
//    scratch, kernel, cudaMemcpyAsync back; the copy-in of one call overlaps the kernel of the
//    previous one and the copy-out of the one before.
#include <cstdlib>
#include <new>
#include <unordered_map>
#include <vector>

#include "licv_common.cuh"

struct licv_host_session {
    struct Slot {
        char* scratch = nullptr;
        cudaStream_t stream = nullptr;
    };
    // hidden states kept on the device between a forward and its backward ("saved for backward")
    struct Saved {
        void* dev = nullptr;
        int64_t bytes = 0;          // capacity
        cudaEvent_t ready = nullptr;   // recorded after the last use (write by fwd / read by bwd)
    };
    std::vector<Slot> slots;
    int64_t bytes = 0;
    uint64_t next = 0;
    std::unordered_map<int64_t, Saved> saved;   // key -> buffer holding that forward's h
    std::vector<Saved> free_list;               // released buffers, reused by size
};

namespace {

// bump allocator over one slot's scratch, 256-byte granules
struct Carver {
    char* base;
    int64_t cap;
    int64_t off = 0;
    bool ok = true;
    void* take(int64_t bytes) {
        const int64_t a = (off + 255) & ~int64_t(255);
        if (a + bytes > cap) {
            ok = false;
            return nullptr;
        }
        off = a + bytes;
        return base + a;
    }
};

inline int64_t esize(int dtype) { return dtype == LICV_F32 ? 4 : 2; }

int host_grid_cap() {
    static const int cap = [] {
        const char* v = std::getenv("LICV_HOST_GRID_CAP");
        return v ? std::atoi(v) : 16;
    }();
    return cap;
}

bool zero_copy_enabled() {
    static const bool on = [] {
        const char* v = std::getenv("LICV_HOST_ZERO_COPY");
        return !(v && v[0] == '0');
    }();
    return on;
}

// device-side address of a host buffer the GPU can reach directly (pinned, registered, managed or
// device memory); nullptr for pageable memory
template <typename T>
T* device_view(T* p) {
    if (!p || !zero_copy_enabled()) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice ||
        at.type == cudaMemoryTypeManaged)
        return static_cast<T*>(const_cast<void*>(static_cast<const void*>(at.devicePointer)));
    return nullptr;
}

licv_host_session::Slot* acquire(licv_host_session* s) {
    auto* slot = &s->slots[s->next++ % s->slots.size()];
    // the slot's previous call (n_slots calls ago) must have drained before its scratch is reused
    cudaStreamSynchronize(slot->stream);
    return slot;
}

#define LICV_CUDA(x)                                   \
    do {                                               \
        cudaError_t e__ = (x);                         \
        if (e__ != cudaSuccess) return (int)e__;       \
    } while (0)

}  // namespace

extern "C" int licv_host_session_create(licv_host_session** out, int64_t scratch_bytes_per_slot,
                                        int n_slots) {
    if (!out) return LICV_ERR_NULL_POINTER;
    if (scratch_bytes_per_slot <= 0 || n_slots < 1 || n_slots > 16) return LICV_ERR_BAD_ARGUMENT;
    if (licv::device_info().status != LICV_OK) return licv::device_info().status;
    auto* s = new (std::nothrow) licv_host_session();
    if (!s) return LICV_ERR_BAD_ARGUMENT;
    s->bytes = scratch_bytes_per_slot;
    s->slots.resize(n_slots);
    for (auto& slot : s->slots) {
        cudaError_t e = cudaMalloc(&slot.scratch, (size_t)scratch_bytes_per_slot);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&slot.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            licv_host_session_destroy(s);
            return (int)e;
        }
    }
    *out = s;
    return LICV_OK;
}

extern "C" int licv_host_session_destroy(licv_host_session* s) {
    if (!s) return LICV_OK;
    for (auto& slot : s->slots) {
        if (slot.stream) {
            cudaStreamSynchronize(slot.stream);
            cudaStreamDestroy(slot.stream);
        }
        if (slot.scratch) cudaFree(slot.scratch);
    }
    for (auto& kv : s->saved) {
        if (kv.second.ready) cudaEventDestroy(kv.second.ready);
        if (kv.second.dev) cudaFree(kv.second.dev);
    }
    for (auto& b : s->free_list) {
        if (b.ready) cudaEventDestroy(b.ready);
        if (b.dev) cudaFree(b.dev);
    }
    delete s;
    return LICV_OK;
}

extern "C" int licv_host_sync(licv_host_session* s) {
    if (!s) return LICV_ERR_NULL_POINTER;
    for (auto& slot : s->slots) LICV_CUDA(cudaStreamSynchronize(slot.stream));
    return LICV_OK;
}

extern "C" void* licv_host_alloc_pinned(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

extern "C" void licv_host_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

extern "C" int licv_inject_fwd_host(licv_host_session* s, const void* h, const float* shift,
                                    void* out, int64_t n_tokens, int d, int h_dtype, int out_dtype,
                                    unsigned round_flags) {
    if (!s) return LICV_ERR_NULL_POINTER;
    if (n_tokens < 0 || d <= 0) return LICV_ERR_BAD_ARGUMENT;
    if (n_tokens == 0) return LICV_OK;
    if (!h || !shift || !out) return LICV_ERR_NULL_POINTER;
    auto* slot = acquire(s);
    Carver c{slot->scratch, s->bytes};
    const int64_t hb = n_tokens * d * esize(h_dtype), ob = n_tokens * d * esize(out_dtype);
    {
        const void* zh = device_view(h);
        void* zout = device_view(out);
        if (zh && zout && licv::aligned16(zh) && licv::aligned16(zout)) {   // zero-copy: h is read and out is written over the link by the kernel
            float* ds = static_cast<float*>(c.take((int64_t)d * 4));
            if (!c.ok) return LICV_ERR_WORKSPACE;
            LICV_CUDA(cudaMemcpyAsync(ds, shift, (size_t)d * 4, cudaMemcpyHostToDevice, slot->stream));
            licv::GridCapScope few_ctas(host_grid_cap());
            return licv_inject_fwd(zh, ds, zout, n_tokens, d, h_dtype, out_dtype, round_flags,
                                   reinterpret_cast<licv_stream_t>(slot->stream));
        }
    }
    void* dh = c.take(hb);
    float* ds = static_cast<float*>(c.take((int64_t)d * 4));
    void* dout = c.take(ob);
    if (!c.ok) return LICV_ERR_WORKSPACE;
    LICV_CUDA(cudaMemcpyAsync(dh, h, (size_t)hb, cudaMemcpyHostToDevice, slot->stream));
    LICV_CUDA(cudaMemcpyAsync(ds, shift, (size_t)d * 4, cudaMemcpyHostToDevice, slot->stream));
    if (int rc = licv_inject_fwd(dh, ds, dout, n_tokens, d, h_dtype, out_dtype, round_flags,
                                 reinterpret_cast<licv_stream_t>(slot->stream)))
        return rc;
    LICV_CUDA(cudaMemcpyAsync(out, dout, (size_t)ob, cudaMemcpyDeviceToHost, slot->stream));
    return LICV_OK;
}

extern "C" int licv_inject_bwd_host(licv_host_session* s, const void* h, const void* g,
                                    const float* shift, void* dh_out, float* d_shift,
                                    int64_t n_tokens, int d, int h_dtype, int g_dtype,
                                    unsigned round_flags) {
    if (!s) return LICV_ERR_NULL_POINTER;
    if (n_tokens < 0 || d <= 0) return LICV_ERR_BAD_ARGUMENT;
    if (!d_shift) return LICV_ERR_NULL_POINTER;
    if (n_tokens > 0 && (!h || !g || !shift)) return LICV_ERR_NULL_POINTER;
    auto* slot = acquire(s);
    Carver c{slot->scratch, s->bytes};
    const int64_t hb = n_tokens * d * esize(h_dtype), gb = n_tokens * d * esize(g_dtype);
    if (n_tokens > 0) {
        const void* zh = device_view(h);
        const void* zg = device_view(g);
        void* zdh = dh_out ? device_view(dh_out) : nullptr;
        if (zh && zg && (!dh_out || zdh) && licv::aligned16(zh) && licv::aligned16(zg) &&
            licv::aligned16(zdh)) {
            float* ds = static_cast<float*>(c.take((int64_t)d * 4));
            float* dds = static_cast<float*>(c.take((int64_t)d * 4));
            if (!c.ok) return LICV_ERR_WORKSPACE;
            LICV_CUDA(cudaMemsetAsync(dds, 0, (size_t)d * 4, slot->stream));
            LICV_CUDA(cudaMemcpyAsync(ds, shift, (size_t)d * 4, cudaMemcpyHostToDevice, slot->stream));
            licv::GridCapScope few_ctas(host_grid_cap());
            if (int rc = licv_inject_bwd(zh, zg, ds, zdh, dds, n_tokens, d, h_dtype, g_dtype,
                                         round_flags, reinterpret_cast<licv_stream_t>(slot->stream)))
                return rc;
            LICV_CUDA(cudaMemcpyAsync(d_shift, dds, (size_t)d * 4, cudaMemcpyDeviceToHost, slot->stream));
            return LICV_OK;
        }
    }
    void* dh = c.take(hb);
    void* dg = c.take(gb);
    float* ds = static_cast<float*>(c.take((int64_t)d * 4));
    float* dds = static_cast<float*>(c.take((int64_t)d * 4));
    void* ddh = dh_out ? c.take(hb) : nullptr;
    if (!c.ok) return LICV_ERR_WORKSPACE;
    LICV_CUDA(cudaMemsetAsync(dds, 0, (size_t)d * 4, slot->stream));
    if (n_tokens > 0) {
        LICV_CUDA(cudaMemcpyAsync(dh, h, (size_t)hb, cudaMemcpyHostToDevice, slot->stream));
        LICV_CUDA(cudaMemcpyAsync(dg, g, (size_t)gb, cudaMemcpyHostToDevice, slot->stream));
        LICV_CUDA(cudaMemcpyAsync(ds, shift, (size_t)d * 4, cudaMemcpyHostToDevice, slot->stream));
        if (int rc = licv_inject_bwd(dh, dg, ds, ddh, dds, n_tokens, d, h_dtype, g_dtype,
                                     round_flags, reinterpret_cast<licv_stream_t>(slot->stream)))
            return rc;
        if (dh_out)
            LICV_CUDA(cudaMemcpyAsync(dh_out, ddh, (size_t)hb, cudaMemcpyDeviceToHost, slot->stream));
    }
    LICV_CUDA(cudaMemcpyAsync(d_shift, dds, (size_t)d * 4, cudaMemcpyDeviceToHost, slot->stream));
    return LICV_OK;
}

// ---- forward that keeps h on the device for its backward -------------------------------------
extern "C" int licv_inject_fwd_host_save(licv_host_session* s, int64_t key, const void* h,
                                         const float* shift, void* out, int64_t n_tokens, int d,
                                         int h_dtype, int out_dtype, unsigned round_flags) {
    if (!s) return LICV_ERR_NULL_POINTER;
    if (n_tokens <= 0 || d <= 0) return LICV_ERR_BAD_ARGUMENT;
    if (!h || !shift || !out) return LICV_ERR_NULL_POINTER;
    const int64_t hb = n_tokens * d * esize(h_dtype), ob = n_tokens * d * esize(out_dtype);
    // a buffer for this key: the key's previous one, else a released one that is large enough
    licv_host_session::Saved buf;
    auto it = s->saved.find(key);
    if (it != s->saved.end()) {
        buf = it->second;
        s->saved.erase(it);
    }
    if (buf.bytes < hb) {
        if (buf.dev) s->free_list.push_back(buf);
        buf = licv_host_session::Saved();
        for (size_t i = 0; i < s->free_list.size(); ++i) {
            if (s->free_list[i].bytes >= hb) {
                buf = s->free_list[i];
                s->free_list.erase(s->free_list.begin() + i);
                break;
            }
        }
    }
    if (!buf.dev) {
        LICV_CUDA(cudaMalloc(&buf.dev, (size_t)hb));
        buf.bytes = hb;
        LICV_CUDA(cudaEventCreateWithFlags(&buf.ready, cudaEventDisableTiming));
        LICV_CUDA(cudaEventRecord(buf.ready, nullptr));
    }
    auto* slot = acquire(s);
    Carver c{slot->scratch, s->bytes};
    float* ds = static_cast<float*>(c.take((int64_t)d * 4));
    void* zout = device_view(out);
    void* dout = (zout && licv::aligned16(zout)) ? zout : c.take(ob);
    if (!c.ok) {
        s->free_list.push_back(buf);
        return LICV_ERR_WORKSPACE;
    }
    cudaStream_t st = slot->stream;
    int rc = (int)cudaStreamWaitEvent(st, buf.ready, 0);   // the buffer's previous user is done
    if (!rc) rc = (int)cudaMemcpyAsync(buf.dev, h, (size_t)hb, cudaMemcpyHostToDevice, st);
    if (!rc) rc = (int)cudaMemcpyAsync(ds, shift, (size_t)d * 4, cudaMemcpyHostToDevice, st);
    if (!rc) {
        licv::GridCapScope few_ctas(dout == zout ? host_grid_cap() : 0);
        rc = licv_inject_fwd(buf.dev, ds, dout, n_tokens, d, h_dtype, out_dtype, round_flags,
                             reinterpret_cast<licv_stream_t>(st));
    }
    if (!rc && dout != zout) rc = (int)cudaMemcpyAsync(out, dout, (size_t)ob, cudaMemcpyDeviceToHost, st);
    if (!rc) rc = (int)cudaEventRecord(buf.ready, st);
    if (rc) {
        s->free_list.push_back(buf);
        return rc;
    }
    s->saved[key] = buf;
    return LICV_OK;
}

extern "C" int licv_inject_bwd_host_saved(licv_host_session* s, int64_t key, const void* g,
                                          const float* shift, void* dh_out, float* d_shift,
                                          int64_t n_tokens, int d, int h_dtype, int g_dtype,
                                          unsigned round_flags) {
    if (!s) return LICV_ERR_NULL_POINTER;
    if (n_tokens <= 0 || d <= 0) return LICV_ERR_BAD_ARGUMENT;
    if (!g || !shift || !d_shift) return LICV_ERR_NULL_POINTER;
    auto it = s->saved.find(key);
    const int64_t hb = n_tokens * d * esize(h_dtype), gb = n_tokens * d * esize(g_dtype);
    if (it == s->saved.end() || it->second.bytes < hb) return LICV_ERR_BAD_ARGUMENT;   // no such forward
    licv_host_session::Saved buf = it->second;
    auto* slot = acquire(s);
    Carver c{slot->scratch, s->bytes};
    float* ds = static_cast<float*>(c.take((int64_t)d * 4));
    float* dds = static_cast<float*>(c.take((int64_t)d * 4));
    const void* zg = device_view(g);
    void* zdh = dh_out ? device_view(dh_out) : nullptr;
    const bool zero_copy = zg && licv::aligned16(zg) && (!dh_out || (zdh && licv::aligned16(zdh)));
    void* dg = zero_copy ? const_cast<void*>(zg) : c.take(gb);
    void* ddh = !dh_out ? nullptr : (zero_copy ? zdh : c.take(hb));
    if (!c.ok) return LICV_ERR_WORKSPACE;
    cudaStream_t st = slot->stream;
    LICV_CUDA(cudaStreamWaitEvent(st, buf.ready, 0));        // the forward's copy of h has landed
    LICV_CUDA(cudaMemsetAsync(dds, 0, (size_t)d * 4, st));
    LICV_CUDA(cudaMemcpyAsync(ds, shift, (size_t)d * 4, cudaMemcpyHostToDevice, st));
    if (!zero_copy) LICV_CUDA(cudaMemcpyAsync(dg, g, (size_t)gb, cudaMemcpyHostToDevice, st));
    {
        licv::GridCapScope few_ctas(zero_copy ? host_grid_cap() : 0);
        if (int rc = licv_inject_bwd(buf.dev, dg, ds, ddh, dds, n_tokens, d, h_dtype, g_dtype,
                                     round_flags, reinterpret_cast<licv_stream_t>(st)))
            return rc;
    }
    if (dh_out && !zero_copy)
        LICV_CUDA(cudaMemcpyAsync(dh_out, ddh, (size_t)hb, cudaMemcpyDeviceToHost, st));
    LICV_CUDA(cudaMemcpyAsync(d_shift, dds, (size_t)d * 4, cudaMemcpyDeviceToHost, st));
    // the saved h is consumed: its buffer may be reused once this backward has run
    LICV_CUDA(cudaEventRecord(buf.ready, st));
    s->saved.erase(it);
    s->free_list.push_back(buf);
    return LICV_OK;
}

extern "C" int licv_kd_loss_fwd_bwd_host(licv_host_session* s, const void* stu, void* dstu,
                                         const void* tea, const int32_t* kl_tea_row,
                                         const int64_t* ce_label, int64_t n_kl, int64_t n_ce,
                                         float temperature, float kl_eps, float hard_loss_weight,
                                         int only_hard_loss, float grad_scale, float* out_losses,
                                         int64_t n_rows, int64_t n_tea_rows, int vocab, int dtype,
                                         unsigned round_flags) {
    if (!s) return LICV_ERR_NULL_POINTER;
    if (n_rows < 0 || n_tea_rows < 0 || vocab <= 0) return LICV_ERR_BAD_ARGUMENT;
    if (!out_losses) return LICV_ERR_NULL_POINTER;
    if (n_rows > 0 && !stu) return LICV_ERR_NULL_POINTER;
    auto* slot = acquire(s);
    Carver c{slot->scratch, s->bytes};
    const int64_t sb = n_rows * vocab * esize(dtype), tb = n_tea_rows * vocab * esize(dtype);
    if (n_rows > 0) {
        const void* zs = device_view(stu);
        const void* zt = tea ? device_view(tea) : nullptr;
        void* zd = dstu ? device_view(dstu) : nullptr;
        if (zs && (!tea || zt) && (!dstu || zd)) {
            int32_t* d_ktr = kl_tea_row ? static_cast<int32_t*>(c.take(n_rows * 4 + 4)) : nullptr;
            int64_t* d_lab = ce_label ? static_cast<int64_t*>(c.take(n_rows * 8 + 8)) : nullptr;
            float* d_loss = static_cast<float*>(c.take(16));
            void* d_ws = c.take(licv_kd_loss_workspace_bytes(n_rows));
            if (!c.ok) return LICV_ERR_WORKSPACE;
            cudaStream_t st = slot->stream;
            LICV_CUDA(cudaMemsetAsync(d_ws, 0, 16, st));
            if (d_ktr) LICV_CUDA(cudaMemcpyAsync(d_ktr, kl_tea_row, (size_t)n_rows * 4, cudaMemcpyHostToDevice, st));
            if (d_lab) LICV_CUDA(cudaMemcpyAsync(d_lab, ce_label, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st));
            if (int rc = licv_kd_loss_fwd_bwd(zs, zd, zt, d_ktr, d_lab, nullptr, n_kl, n_ce, temperature,
                                              kl_eps, hard_loss_weight, only_hard_loss, grad_scale,
                                              d_loss, d_ws, n_rows, vocab, vocab, vocab, dtype,
                                              round_flags, reinterpret_cast<licv_stream_t>(st)))
                return rc;
            LICV_CUDA(cudaMemcpyAsync(out_losses, d_loss, 12, cudaMemcpyDeviceToHost, st));
            return LICV_OK;
        }
    }
    // device rows are padded to a multiple of 8 elements so that every row starts 16-byte aligned
    void* d_stu = c.take(sb > 0 ? sb : 16);
    void* d_tea = (tea && tb > 0) ? c.take(tb) : nullptr;
    int32_t* d_ktr = kl_tea_row ? static_cast<int32_t*>(c.take(n_rows * 4 + 4)) : nullptr;
    int64_t* d_lab = ce_label ? static_cast<int64_t*>(c.take(n_rows * 8 + 8)) : nullptr;
    float* d_loss = static_cast<float*>(c.take(16));
    const int64_t wsb = licv_kd_loss_workspace_bytes(n_rows);
    void* d_ws = c.take(wsb);
    if (!c.ok) return LICV_ERR_WORKSPACE;
    cudaStream_t st = slot->stream;
    LICV_CUDA(cudaMemsetAsync(d_ws, 0, 16, st));
    if (sb > 0) LICV_CUDA(cudaMemcpyAsync(d_stu, stu, (size_t)sb, cudaMemcpyHostToDevice, st));
    if (d_tea) LICV_CUDA(cudaMemcpyAsync(d_tea, tea, (size_t)tb, cudaMemcpyHostToDevice, st));
    if (d_ktr) LICV_CUDA(cudaMemcpyAsync(d_ktr, kl_tea_row, (size_t)n_rows * 4, cudaMemcpyHostToDevice, st));
    if (d_lab) LICV_CUDA(cudaMemcpyAsync(d_lab, ce_label, (size_t)n_rows * 8, cudaMemcpyHostToDevice, st));
    if (int rc = licv_kd_loss_fwd_bwd(d_stu, dstu ? d_stu : nullptr, d_tea, d_ktr, d_lab, nullptr,
                                      n_kl, n_ce, temperature, kl_eps, hard_loss_weight,
                                      only_hard_loss, grad_scale, d_loss, d_ws, n_rows, vocab, vocab,
                                      vocab, dtype, round_flags,
                                      reinterpret_cast<licv_stream_t>(st)))
        return rc;
    if (dstu && sb > 0) LICV_CUDA(cudaMemcpyAsync(dstu, d_stu, (size_t)sb, cudaMemcpyDeviceToHost, st));
    LICV_CUDA(cudaMemcpyAsync(out_losses, d_loss, 12, cudaMemcpyDeviceToHost, st));
    return LICV_OK;
}
