// Shared declarations of the reward-head kernels (head.cu) used by the engine.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mjv {

struct HeadJob {      // y[B,N] = act(x[B,K] · W[N,K]^T + bias), bf16
    const void* x;
    const void* w;
    const void* bias;  // or null
    void* y;
    int N, K;
    int ldx, ldy;
    int relu;
};

struct HeadFinalizeArgs {
    int B;
    int num_objectives, num_aspects;
    float temperature;
    const __nv_bfloat16* raw_rewards;  // [B, ld_raw] regression_layer output
    int ld_raw;
    const __nv_bfloat16* transform;    // [NO, NO] reward_transform_matrix
    const __nv_bfloat16* aspect_logits;    // [B, ld_al]
    int ld_al;
    const __nv_bfloat16* criteria_logits;  // [B, ld_cl]
    int ld_cl;
    const __nv_bfloat16* aspect_logit_scale;    // [1]
    const __nv_bfloat16* criteria_logit_scale;  // [1]
    const int* group_offsets;  // [NA+1] offsets into group_items (dict order of aspect2criteria)
    const int* group_items;    // [NO] criterion index at each concat position
    float* scores;             // [B, ld_scores] = rewards(NO) | aspect_scores(NA) | score(1)
    int ld_scores;
    float* extras;             // [B, ld_extras] = criteria_gating_output(NO) | aspect_gating_output(NA) |
    int ld_extras;             //                  aspect_weights(NO) | weighted_scores(1)
};

struct HeadJobs {  // passed to the kernel by value: no device-side job table to upload
    HeadJob j[3];
    int n;
};

int skinny_linear(const HeadJobs& jobs, int B, cudaStream_t st);
int head_finalize(const HeadFinalizeArgs& a, cudaStream_t st);
int gating_softmax(const void* logits, int ld, int n, float temperature, const void* logit_scale, void* out, int ldo, int B,
                   cudaStream_t st);

}  // namespace mjv
