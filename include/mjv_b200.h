/* mjv_b200.h — C ABI of the B200-native MJ-VIDEO reward-scoring forward pass.
 *
 * The reference (aiming-lab/MJ-Video) has no FFI: its hot path is three Python callables
 *   load_video                         scripts/data_processor/data.py:158-179
 *   prepare_chat_input                 scripts/model/internvl2/modeling_internvl_chat.py:36-89
 *   InternVLChatRewardModeling.forward scripts/model/moe_reward.py:183-297
 * The Python package `mj_video_b200` keeps those names/signatures and calls THIS library through ctypes.
 * Each entry point below names the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; device pointers are raw `void*`; `stream` is a cudaStream_t passed as `void*`.
 *   - the library never synchronises the device and never allocates caller-visible memory; scratch comes from a
 *     caller-provided workspace (mjv_workspace_bytes). Weights are copied/re-packed into library-owned memory.
 *   - return 0 on success, a negative MJV_ERR_* otherwise; mjv_last_error() gives the (thread-local) message.
 *   - bf16 tensors are row-major with the innermost dimension contiguous.
 *   - a handle is not thread-safe; handles on different devices are independent.
 */
#ifndef MJV_B200_H_
#define MJV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MJV_ABI_VERSION 1

enum {
    MJV_OK = 0,
    MJV_ERR_ARG = -1,     /* bad argument / unsupported shape */
    MJV_ERR_CUDA = -2,    /* CUDA runtime / driver failure     */
    MJV_ERR_STATE = -3,   /* handle not ready (missing weights) */
    MJV_ERR_KEY = -4,     /* unknown state-dict key             */
    MJV_ERR_NOMEM = -5,   /* workspace too small                */
    MJV_ERR_NCCL = -6
};

enum { MJV_DTYPE_F32 = 0, MJV_DTYPE_BF16 = 1, MJV_DTYPE_F16 = 2 };

int mjv_abi_version(void);
const char* mjv_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * Primitive: tcgen05 GEMM with fused epilogue.  D[M,N] = A[M,K] · W[N,K]^T.
 * Replaces every nn.Linear / nn.Conv2d call site of the path (SURVEY §2.4 K1,K4,K7,K8,K10,K13,K16,K17):
 *   modeling_intern_vit.py:164,230,242,260-262; modeling_internvl_chat.py:135-140;
 *   modeling_internlm2.py:262,461,502.
 * Exposed so that each fused epilogue is separately testable against the oracle.
 * ---------------------------------------------------------------------------------------------------------- */
enum {
    MJV_EPI_STORE = 0,
    MJV_EPI_GELU = 1,
    MJV_EPI_RESID = 2,
    MJV_EPI_SWIGLU = 3,
    MJV_EPI_QKV_LLM = 4,
    MJV_EPI_QKV_VIT = 5,
    MJV_EPI_PATCH = 6,
    MJV_EPI_SCATTER = 7
};

typedef struct mjv_gemm_desc {
    const void* a;        /* [M, K] bf16, row stride lda (elements, multiple of 8) */
    const void* w;        /* [N, K] bf16, row stride ldw                             */
    void* out;            /* bf16 output (meaning depends on epilogue)               */
    int64_t lda, ldw, ldo;
    int32_t M, N, K;
    int32_t epilogue;     /* MJV_EPI_*                                               */
    int32_t block_n;      /* 0 = auto, else 128 or 256                               */
    const void* bias;     /* [N] bf16 or NULL                                        */
    const void* residual; /* [M, ldr] bf16 (RESID)                                   */
    int64_t ldr;
    const void* scale;    /* [N] bf16 layer-scale or NULL (RESID)                    */
    void* out_k;          /* QKV_LLM: K [M, n_k_cols]                                */
    void* out_vt;         /* QKV_*: transposed V                                     */
    const float* rope_cos;/* QKV_LLM: [max_pos, 64]                                  */
    const float* rope_sin;
    const int32_t* row_pos; /* QKV_LLM: position id of each row                      */
    const int32_t* row_map; /* SCATTER: destination row of each m                    */
    const int32_t* row_vt;  /* QKV_LLM: V^T column of each row (NULL = row index)    */
    const void* pos_emb;  /* PATCH: [seq+1, N] bf16                                  */
    int32_t n_q_cols, n_k_cols;
    int64_t vt_stride_head;
    int32_t vt_ld;
    int32_t seq;
    int32_t heads;
} mjv_gemm_desc;

int mjv_gemm_bf16(const mjv_gemm_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Primitive: fused softmax attention (tcgen05 QK^T and PV, online softmax, O accumulated in TMEM).
 * Replaces flash_attn_varlen_qkvpacked_func(causal=False) (modeling_intern_vit.py:72-75; eager :210-227) and
 * flash_attn_func / flash_attn_varlen_func(causal=True) with GQA (modeling_internlm2.py:498-561; eager :393-410).
 *   batched mode (cu_seqlens == NULL): nseq sequences of seq_len rows; q,k rows = nseq*seq_len;
 *       V^T planes = nseq*heads, each [head_dim][vt_ld] holding seq_len valid columns.
 *   varlen  mode (cu_seqlens != NULL): packed rows; sequence s owns rows [cu[s], cu[s+1]); V^T planes = kv_heads,
 *       each [head_dim][vt_ld] over the packed rows; `work` lists the (sequence, 128-row query block) items.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct mjv_attn_desc {
    const void* q;        /* bf16 [rows, ldq]; head h occupies columns q_col0 + h*head_dim ...            */
    const void* k;        /* bf16 [rows, ldk]; kv head g occupies columns k_col0 + g*head_dim ...         */
    const void* vt;       /* bf16 V transposed                                                            */
    void* out;            /* bf16 [rows, ldo]; head h occupies columns h*head_dim ...                     */
    int64_t ldq, ldk, ldo;
    int64_t vt_ld, vt_plane_stride;
    int64_t total_rows;
    int32_t head_dim;     /* 64 or 128                                                                    */
    int32_t heads, kv_heads;
    int32_t q_col0, k_col0;
    int32_t causal;
    int32_t nseq, seq_len;
    const int32_t* cu_seqlens; /* device [nseq+1] or NULL                                                 */
    const int32_t* work;       /* device int32[2*n_work] (seq, q_block) pairs or NULL                     */
    const int32_t* vt_starts;  /* varlen: device [nseq] first V^T column of each sequence (multiple of 8: TMA
                                  needs a 16-byte aligned innermost coordinate) or NULL (= cu_seqlens)      */
    int64_t vt_cols;           /* varlen: number of V^T columns in use (<= vt_ld); 0 = total_rows         */
    int32_t n_work;
    float scale;          /* softmax scale (1/sqrt(head_dim))                                             */
} mjv_attn_desc;

int mjv_attention(const mjv_attn_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MJV_B200_H_ */
