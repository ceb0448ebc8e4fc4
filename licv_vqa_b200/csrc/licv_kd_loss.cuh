// Argument block shared by the distillation-loss kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace licv {

struct KdArgs {
    const void* stu;
    void* dstu;              // may alias stu; nullptr = loss only
    const void* tea;
    const int32_t* kl_tea_row;  // nullptr = identity pairing
    const int64_t* ce_label;    // nullptr = no CE term
    const int32_t* counts;      // device {N, M} or nullptr
    int64_t n_kl, n_ce;         // host N, M (used when counts == nullptr)
    float temperature, kl_eps, hard_loss_weight, grad_scale;
    bool only_hard_loss;
    float* out_losses;          // {kl, ce, total}
    unsigned* counter;          // zero on entry, zero on exit
    float* row_loss;            // [2 * n_rows]: per-row KL sums, per-row CE values
    int64_t n_rows;
    int vocab;
    int64_t stu_stride, tea_stride;  // elements
    unsigned round_flags;
};

int launch_kd_generic(const KdArgs& a, int dtype, cudaStream_t st, bool want_dtemp = false);

// cluster kernel (licv_kd_loss_cluster.cu): does a row of `vocab` elements fit a cluster, and how
bool kd_cluster_plan(int vocab, int dtype, float temperature, bool kl_and_ce, int* C, int* NV,
                     int* NT);
// one CTA per SM, the row cached in shared memory + tensor memory: 16-bit logits, V <= 32760
// (LICV_KD_TMEM=0 disables it)
bool kd_tmem_plan(int vocab, int dtype, float temperature, bool kl_and_ce);
int launch_kd_tmem(const KdArgs& a, int dtype, cudaStream_t st);
// stream kernel (licv_kd_loss_stream.cu): 16-bit logits, 16 377 <= V <= 32 752; TMA-staged rows,
// sweep D of a row fused with sweep B of the next (LICV_KD_STREAM=0 disables it)
bool kd_stream_plan(int vocab, int dtype);
bool kd_stream_forced();   // LICV_KD_STREAM=2 / licv_debug_set_kd_stream(2): also for a handful of rows
int launch_kd_stream(const KdArgs& a, int dtype, cudaStream_t st);
int launch_kd_cluster(const KdArgs& a, int dtype, int C, int NV, int NT, cudaStream_t st);

}  // namespace licv
