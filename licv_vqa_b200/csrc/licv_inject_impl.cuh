// Dispatch of the injection kernels for one hidden-state dtype (instantiated per dtype in
// licv_inject_fwd_*.cu / licv_inject_bwd_*.cu so the translation units compile in parallel).
#pragma once
#include "licv_inject.cuh"
#include "licv_inject_pipe.cuh"

namespace licv {
namespace inject {

struct Launch {
    int64_t n_tok;
    int nvec;
    RowPlan plan;
    unsigned flags;
    cudaStream_t stream;
    int partial_rows = 0;   // > 0: backward in spread mode on exactly this many CTAs
};

// Spread backward: CTAs of a launch and the replica count that keeps <= 16 CTAs on one atomic
// address.  Token batches as in dispatch_bwd_pipe_tb: one token per CTA until the machine is full
// (tokens inside a CTA are processed one after the other, ~0.9 us each: at 256 tokens 256 CTAs x 1
// token take 4.7 us, 128 x 2 5.0 us, 64 x 4 7.0 us, the cluster reduction 5.6 us), at most two CTAs
// per SM.
inline int pipe_partial_tb(int64_t n_tok, int nvec, int h_bytes, int g_bytes) {
    const int vpt = nvec / kPipeThreads;
    const int gv = g_bytes > h_bytes ? 2 : 1;
    const int fit = 8 / (vpt * (1 + gv));
    const int tb_big = fit >= 4 ? 4 : (fit >= 2 ? 2 : 1);
    const int64_t fill = (int64_t)device_info().sm_count * 2;
    return (tb_big > 1 && n_tok >= fill * tb_big) ? tb_big : 1;
}
inline int pipe_partial_ctas(int64_t n_tok, int nvec, int h_bytes, int g_bytes) {
    const int tb = pipe_partial_tb(n_tok, nvec, h_bytes, g_bytes);
    int64_t ctas = (n_tok + tb - 1) / tb;
    const int64_t cap = (int64_t)device_info().sm_count * 2;
    if (ctas > cap) ctas = cap;
    return ctas < 1 ? 1 : (int)ctas;
}
inline int spread_rows_for(int ctas) {
    int r = 1;
    while (r < 16 && r * 16 < ctas) r *= 2;
    return r;
}

template <int HDT, int ODT, int VPT, int TB, int RND>
int launch_fwd(const Args& a0, const Launch& L) {
    auto kern = fwd_kernel<HDT, ODT, VPT, TB, RND>;
    const int groups = kCtaThreads / L.plan.gt > 0 ? kCtaThreads / L.plan.gt : 1;
    const int threads = groups * L.plan.gt;
    static PerDevice<int> per_sm_dev;
    const int per_sm = per_sm_dev.get([&] { return resident_ctas_per_sm(kern, kCtaThreads, 0); });
    int64_t cap = (int64_t)per_sm * device_info().sm_count;
    if (tl_grid_cap > 0 && tl_grid_cap < cap) cap = tl_grid_cap;
    const int64_t need = ((L.n_tok + TB - 1) / TB + groups - 1) / groups;
    Args a = a0;
    a.gt = L.plan.gt;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(need < cap ? need : cap));
    cfg.blockDim = dim3(threads);
    cfg.stream = L.stream;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = launch_attrs(attr, 0);
    return (int)cudaLaunchKernelEx(&cfg, kern, a);
}

template <int HDT, int GDT, int VPT, int TB, int RND>
int launch_bwd(const Args& a0, const Launch& L) {
    auto kern = bwd_kernel<HDT, GDT, VPT, TB, RND>;
    static const int min_tok = env_int("LICV_BWD_MIN_TOKENS_PER_GROUP", 2);  // fewer, fatter CTAs
    const int groups = kCtaThreads / L.plan.gt > 0 ? kCtaThreads / L.plan.gt : 1;
    const int threads = groups * L.plan.gt;
    // one fp32 row (nvec * EPV floats) per extra group
    const size_t smem = (size_t)(groups - 1) * L.nvec * Fmt<HDT>::kPerVec * sizeof(float);
    if (smem > 48 * 1024) {
        static PerDevice<int> raised;
        raised.get([&] {
            return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        });
    }
    static PerDevice<int> per_sm_dev;
    const int per_sm = per_sm_dev.get([&] { return resident_ctas_per_sm(kern, kCtaThreads, 16 * 1024); });
    int64_t cap = (int64_t)per_sm * device_info().sm_count;
    if (tl_grid_cap > 0 && tl_grid_cap < cap) cap = tl_grid_cap;
    const int tok_per_group = TB > min_tok ? TB : min_tok;
    const int64_t need = ((L.n_tok + tok_per_group - 1) / tok_per_group + groups - 1) / groups;
    Args a = a0;
    a.gt = L.plan.gt;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(need < cap ? need : cap));
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = L.stream;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = launch_attrs(attr, 0);
    return (int)cudaLaunchKernelEx(&cfg, kern, a);
}

// tokens staged per iteration: one while there are too few tokens to fill the machine, else
// enough for 8 vector loads in flight per thread
template <int HDT, int ODT, int VPT, int RND>
int dispatch_fwd_tb(const Args& a, const Launch& L) {
    constexpr int TBbig = 8 / VPT > 0 ? 8 / VPT : 1;
    const int groups = kCtaThreads / L.plan.gt > 0 ? kCtaThreads / L.plan.gt : 1;
    const int64_t fill = (int64_t)device_info().sm_count * 2 * groups;
    if (TBbig > 1 && L.n_tok >= fill * TBbig) return launch_fwd<HDT, ODT, VPT, TBbig, RND>(a, L);
    return launch_fwd<HDT, ODT, VPT, 1, RND>(a, L);
}

template <int HDT, int GDT, int VPT, int RND>
int dispatch_bwd_tb(const Args& a, const Launch& L) {
    constexpr int GV = (Fmt<GDT>::kBytes > Fmt<HDT>::kBytes) ? 2 : 1;
    constexpr int kFit = 8 / (VPT * (1 + GV));
    constexpr int TBbig = kFit >= 4 ? 4 : (kFit >= 2 ? 2 : 1);
    const int groups = kCtaThreads / L.plan.gt > 0 ? kCtaThreads / L.plan.gt : 1;
    const int64_t fill = (int64_t)device_info().sm_count * 2 * groups;
    if (TBbig > 1 && L.n_tok >= fill * TBbig) return launch_bwd<HDT, GDT, VPT, TBbig, RND>(a, L);
    return launch_bwd<HDT, GDT, VPT, 1, RND>(a, L);
}


// ---- TMA-staged backward (licv_inject_pipe.cuh): rows of 1, 2 or 4 sweeps of 256 vectors ------
inline bool pipe_row(int nvec) {
    static const int off = env_int("LICV_INJECT_NO_PIPE", 0);
    return !off && nvec % kPipeThreads == 0 && nvec / kPipeThreads <= 4 &&
           (nvec / kPipeThreads) != 3;
}

template <int HDT, int GDT, int VPT, int TB, int RND>
int launch_bwd_pipe(const Args& a, const Launch& L) {
    auto kern = bwd_pipe_kernel<HDT, GDT, VPT, TB, RND>;
    constexpr int GV = (Fmt<GDT>::kBytes > Fmt<HDT>::kBytes) ? 2 : 1;
    constexpr int kStage = TB * VPT * kPipeThreads * 16 * (1 + GV);
    static const int budget = env_int("LICV_PIPE_SMEM_KB", 108) * 1024;   // two CTAs per SM
    int stages = budget / kStage;
    if (stages > kPipeMaxStages) stages = kPipeMaxStages;
    if (stages < 2) return LICV_ERR_BAD_DIM;
    const size_t smem = (size_t)stages * kStage;
    static PerDevice<int> raised;   // smem is a constant of the instantiation
    {
        const int e = raised.get([&] {
            return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        });
        if (e != 0) return e;
    }
    PipeArgs pa;
    pa.a = a;
    pa.n_stages = stages;
    pa.n_batches = (L.n_tok + TB - 1) / TB;
    static const int per_sm = env_int("LICV_PIPE_CTAS_PER_SM", 2);
    static const int cluster = env_int("LICV_PIPE_CLUSTER", 4);   // 1, 2, 4 or 8
    // a cluster launch and its two cluster barriers cost ~1 us: only worth it once enough CTAs
    // would otherwise queue on the same d_shift addresses
    const int C = (L.partial_rows == 0 && (cluster == 2 || cluster == 4 || cluster == 8) && pa.n_batches > 32)
                      ? cluster
                      : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kPipeThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = L.stream;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = launch_attrs(attr, C > 1 ? C : 0);
    // resident CTAs: whole clusters that fit the device at once (per cluster size and device)
    static PerDevice<int64_t> cap_by_c[9];
    const int64_t cap = cap_by_c[C].get([&]() -> int64_t {
        int n_clusters = 0;
        int64_t c;
        cfg.gridDim = dim3(C * device_info().sm_count);
        if (C > 1 && cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg) == cudaSuccess &&
            n_clusters > 0) {
            c = (int64_t)n_clusters * C;
        } else {
            cudaGetLastError();
            c = (int64_t)per_sm * device_info().sm_count / C * C;
        }
        const int64_t want = (int64_t)per_sm * device_info().sm_count;
        if (c > want) c = want / C * C;
        return c;
    });
    int64_t lim = cap;
    if (tl_grid_cap > 0 && tl_grid_cap < lim) lim = tl_grid_cap;
    int64_t grid = pa.n_batches < lim ? pa.n_batches : lim;
    grid = (grid + C - 1) / C * C;
    if (L.partial_rows > 0) grid = L.partial_rows;   // every row of the caller's block is written
    cfg.gridDim = dim3((unsigned)grid);
    return (int)cudaLaunchKernelEx(&cfg, kern, pa);
}

template <int HDT, int GDT, int VPT, int RND>
int dispatch_bwd_pipe_tb(const Args& a, const Launch& L) {
    constexpr int GV = (Fmt<GDT>::kBytes > Fmt<HDT>::kBytes) ? 2 : 1;
    constexpr int kFit = 8 / (VPT * (1 + GV));
    constexpr int TBbig = kFit >= 4 ? 4 : (kFit >= 2 ? 2 : 1);
    static const int forced_tb = env_int("LICV_PIPE_TB", 0);   // tuning knob: 1 = one token per stage
    const int64_t fill = (int64_t)device_info().sm_count * 2;
    // (spread mode ignores the tuning knob: pipe_partial_tb mirrors the plain rule)
    const bool big = TBbig > 1 && (L.partial_rows > 0
                                       ? L.n_tok >= fill * TBbig
                                       : (forced_tb != 1 && (forced_tb > 1 || L.n_tok >= fill * TBbig)));
    if (big) return launch_bwd_pipe<HDT, GDT, VPT, TBbig, RND>(a, L);
    return launch_bwd_pipe<HDT, GDT, VPT, 1, RND>(a, L);
}

template <int HDT, int GDT, int RND>
int dispatch_bwd_pipe(const Args& a, const Launch& L) {
    switch (L.nvec / kPipeThreads) {
        case 1: return dispatch_bwd_pipe_tb<HDT, GDT, 1, RND>(a, L);
        case 2: return dispatch_bwd_pipe_tb<HDT, GDT, 2, RND>(a, L);
        default: return dispatch_bwd_pipe_tb<HDT, GDT, 4, RND>(a, L);
    }
}

template <int HDT, int ODT, int RND>
int dispatch_fwd_vpt(const Args& a, const Launch& L) {
    switch (L.plan.vpt) {
        case 1: return dispatch_fwd_tb<HDT, ODT, 1, RND>(a, L);
        case 2: return dispatch_fwd_tb<HDT, ODT, 2, RND>(a, L);
        case 4: return dispatch_fwd_tb<HDT, ODT, 4, RND>(a, L);
        default: return dispatch_fwd_tb<HDT, ODT, 8, RND>(a, L);  // rows of 16-32 KB
    }
}

template <int HDT, int GDT, int RND>
int dispatch_bwd_vpt(const Args& a, const Launch& L) {
    switch (L.plan.vpt) {
        case 1: return dispatch_bwd_tb<HDT, GDT, 1, RND>(a, L);
        case 2: return dispatch_bwd_tb<HDT, GDT, 2, RND>(a, L);
        case 4: return dispatch_bwd_tb<HDT, GDT, 4, RND>(a, L);
        default: return dispatch_bwd_tb<HDT, GDT, 8, RND>(a, L);  // rows of 16-32 KB
    }
}

template <int HDT, int ODT>
int dispatch_fwd_rnd(const Args& a, const Launch& L) {
    if constexpr (HDT == LICV_F32) {
        return dispatch_fwd_vpt<HDT, ODT, 0>(a, L);
    } else {
        switch (rnd_mode(L.flags)) {
            case 0: return dispatch_fwd_vpt<HDT, ODT, 0>(a, L);
            case 1: return dispatch_fwd_vpt<HDT, ODT, 1>(a, L);
            default: return dispatch_fwd_vpt<HDT, ODT, 2>(a, L);
        }
    }
}

template <int HDT, int GDT>
int dispatch_bwd_rnd(const Args& a, const Launch& L) {
    if (pipe_row(L.nvec)) {
        if constexpr (HDT == LICV_F32) {
            return dispatch_bwd_pipe<HDT, GDT, 0>(a, L);
        } else {
            return rnd_mode(L.flags) == 0 ? dispatch_bwd_pipe<HDT, GDT, 0>(a, L)
                                          : dispatch_bwd_pipe<HDT, GDT, 1>(a, L);
        }
    }
    if constexpr (HDT == LICV_F32) {
        return dispatch_bwd_vpt<HDT, GDT, 0>(a, L);
    } else {
        // the backward only needs to know whether y was stored in low precision
        return rnd_mode(L.flags) == 0 ? dispatch_bwd_vpt<HDT, GDT, 0>(a, L)
                                      : dispatch_bwd_vpt<HDT, GDT, 1>(a, L);
    }
}

// entry points per hidden dtype (defined in the per-dtype translation units)
template <int HDT> int run_fwd(const Args& a, const Launch& L, int out_dtype);
template <int HDT> int run_bwd(const Args& a, const Launch& L, int g_dtype);

#define LICV_DEFINE_RUN_FWD(HDT)                                                   \
    template <> int run_fwd<HDT>(const Args& a, const Launch& L, int out_dtype) {  \
        if (out_dtype == LICV_F32) return dispatch_fwd_rnd<HDT, LICV_F32>(a, L);   \
        return dispatch_fwd_rnd<HDT, HDT>(a, L);                                   \
    }
#define LICV_DEFINE_RUN_BWD(HDT)                                                   \
    template <> int run_bwd<HDT>(const Args& a, const Launch& L, int g_dtype) {    \
        if (g_dtype == LICV_F32) return dispatch_bwd_rnd<HDT, LICV_F32>(a, L);     \
        return dispatch_bwd_rnd<HDT, HDT>(a, L);                                   \
    }

}  // namespace inject
}  // namespace licv
