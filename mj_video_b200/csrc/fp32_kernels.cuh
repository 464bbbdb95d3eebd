// fp32 verification mode (north_star: "1e-4 in fp32"): plain fp32 CUDA kernels, one per reference op, used by
// engine_fp32.cuh. No tensor cores, no fusion beyond a GEMM epilogue: this mode exists to check the bf16 product path and
// the reference against each other at fp32 accuracy, not to be fast (one 2B row ~ 1 s).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mjv {
namespace f32 {

enum { EPI_NONE = 0, EPI_GELU = 1, EPI_RELU = 2, EPI_RESID = 3, EPI_RESID_SCALE = 4 };

// C[M,N] = epi(A[M,K] · W[N,K]^T + bias[N]);  EPI_RESID: res + x;  EPI_RESID_SCALE: res + x * scale[N]   (K % 4 == 0)
int sgemm_tn(const float* A, long long lda, const float* W, long long ldw, float* C, long long ldc, long long M, int N,
             int K, const float* bias, int epi, const float* res, long long ldr, const float* scale, cudaStream_t st);
// F.layer_norm over the last dim (modeling_intern_vit.py:291,293; modeling_internvl_chat.py:136)
int layernorm(const float* x, const float* w, const float* b, float* out, long long rows, int C, float eps, cudaStream_t st);
// InternLM2RMSNorm (modeling_internlm2.py:138-143); row_idx (optional) gathers source rows
int rmsnorm(const float* x, const float* w, float* out, long long rows, int C, float eps, const int* row_idx,
            cudaStream_t st);
// [T,3,img,img] (bf16 or fp32) -> [T*g*g, 3*ps*ps] rows in (c, ky, kx) order = nn.Conv2d weight layout (modeling_intern_vit.py:164)
int patchify(const void* px, int px_dtype, float* out, int T, int img, int ps, cudaStream_t st);
// hidden[t,0,:] = cls + pos[0]; hidden[t,1+p,:] = conv[t*P+p,:] + pos[1+p]   (modeling_intern_vit.py:166-172)
int vit_assemble(const float* conv, const float* cls, const float* pos, float* hidden, int T, int P, int C, cudaStream_t st);
// pixel_shuffle(scale 0.5) of the non-CLS tokens: [T,1+g*g,C] -> [T,(g/2)^2,4C] (modeling_internvl_chat.py:228-242)
int pixel_shuffle(const float* hidden, float* out, int T, int g, int C, int ps_v1, cudaStream_t st);
// out[r,:] = table[ids[r],:] for rows whose id != skip_id (modeling_internvl_chat.py:163)
int embed(const int* ids, const float* table, float* out, long long rows, int C, int skip_id, cudaStream_t st);
// dst[row_map[i],:] = src[i,:]  (splice of the projector rows, modeling_internvl_chat.py:176-177)
int scatter_rows(const float* src, const int* row_map, float* dst, long long n, int C, cudaStream_t st);
int gather_rows(const float* src, const int* idx, float* dst, long long n, int C, cudaStream_t st);
// rotate-half RoPE in place on `heads` heads of hd columns starting at col0 + (h / group) * kv_stride + (h % group) * hd
int rope(float* x, long long ld, long long rows, int heads, int group, int kv_stride, int col0, int hd,
         const float* cos_t, const float* sin_t, int table_ld, const int* row_pos, cudaStream_t st);
// out[row, h*hd + d] = softmax(scale * q·k^T (+ causal mask)) · v ; sequences from cu_seqlens (or fixed seq_len)
struct AttnF32 {
    const float *q, *k, *v;
    long long ldq, ldk, ldv;
    int q_col0, k_col0, v_col0;       // first column of head 0
    int q_group_stride, kv_stride;    // q head h: q_col0 + (h / group) * q_group_stride + (h % group) * hd; kv head g: col0 + g * kv_stride
    int heads, group, hd;
    float scale;
    int causal;
    const int* cu_seqlens;            // [nseq + 1] or nullptr
    int nseq, seq_len;
    float* out;
    long long ldo;
};
int attention(const AttnF32& a, cudaStream_t st);
// silu(gate) * up, element-wise (modeling_internlm2.py:262)
int silu_mul(const float* gate, const float* up, float* out, long long n, cudaStream_t st);
// any supported dtype -> fp32 copy
int to_f32(const void* src, int dtype, float* dst, long long n, cudaStream_t st);

// MoE head tail in fp32 (moe_reward.py:246-297); one warp per sample
struct HeadF32 {
    const float *rewards_raw;      // [B, NO] = regression_layer(h_last)
    const float *transform;        // [NO, NO] reward_transform_matrix
    const float *aspect_logits;    // [B, NA]
    const float *criteria_logits;  // [B, NO]
    const float *aspect_logit_scale, *criteria_logit_scale;  // [1]
    const int *group_offsets, *group_items;
    float temperature;
    int NO, NA, B;
    float *scores, *extras;        // [B, NO+NA+1], [B, NO+NA+NO+1]
};
int head_finalize(const HeadF32& a, cudaStream_t st);
// GatingNetwork.forward tail (moe_reward.py:33-36): softmax(logits / T, dim=1) * logit_scale[0]; n <= 32
int gating_softmax(const float* logits, int n, float temperature, const float* logit_scale, float* out, int B, cudaStream_t st);

}  // namespace f32
}  // namespace mjv
