// Internal launch functions of the kernel families, used by the engine (engine.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/mjv_b200.h"

namespace mjv {
int gemm_bf16(const mjv_gemm_desc& d, cudaStream_t stream);
int attention(const mjv_attn_desc& d, cudaStream_t stream);
int layernorm(const void* x, const void* w, const void* b, void* out, long long rows, int C, float eps, cudaStream_t st);
int pixel_shuffle_layernorm(const void* vit_hidden, const void* w, const void* b, void* out, int tiles, int seq,
                            int c_vit, float eps, int ps_v1, cudaStream_t st);
int rmsnorm(const void* x, const void* w, void* out, long long rows, int C, float eps, const int* row_idx,
            cudaStream_t st);
int embed_tokens(const int* ids, const void* table, void* out, long long rows, int C, int img_ctx_id, int vocab,
                 cudaStream_t st);
int gather_rows(const void* src, const int* idx, void* dst, int rows, int C, cudaStream_t st);
int im2col_patches(const void* px, int px_dtype, void* out, int T, int img, int ps, int Kpad, cudaStream_t st);
int cls_rows(const void* cls, const void* pos, void* hidden, int T, int S, int C, cudaStream_t st);
int preprocess_frames(const void* frames, int F, int H, int W, int image_size, int max_num, int use_thumbnail,
                      void* out, int out_dtype, void* scratch, size_t scratch_bytes, int* tiles_per_frame,
                      cudaStream_t st);
}  // namespace mjv
