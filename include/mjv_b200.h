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
/* sizeof() of the ABI structs as this library was compiled (0 mjv_gemm_desc, 1 mjv_attn_desc, 2 mjv_config, 3 mjv_batch,
 * 4 mjv_outputs; -1 otherwise): lets a binding check its own struct mirrors before the first compute call. */
int mjv_abi_struct_size(int which);
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
    int32_t cta_group;    /* 0 = auto (CTA pairs when block_n = 256), 1 or 2         */
    int32_t v_head_dim;   /* QKV_LLM: columns per V head in the N layout and rows per V^T plane (0 = 128); a multiple of
                             32 — V needs no rotary slot, so narrow heads (e.g. 96) are packed                          */
    int32_t qk_head_slot; /* QKV_LLM: columns per q / k head in N. 0 or 128: rotate-half pairs sit 64 columns apart inside a
                             128-wide slot. 96: native head_dim-96 layout, the pair (d, d+48) is stored in ADJACENT columns
                             (2d, 2d+1) - q.k is invariant under the common permutation - so RoPE works chunk by chunk and
                             a head never has to align with the 128/256-column tile                                   */
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
    double flops_hint;    /* varlen: algorithmic FLOPs of this launch for the profiler (0 = unknown)      */
    /* Heads narrower than the 64/128 tile (e.g. 96): q/k stay in head_dim-wide slots, but
     *   v_head_dim   rows per V^T plane (0 = head_dim); rows beyond it are zero-filled by TMA,
     *   out_head_dim columns per head written to `out` (0 = head_dim; a multiple of 32), head h at column h*out_head_dim. */
    int32_t v_head_dim, out_head_dim;
    /* qk_head_stride: columns from one q (or k) head to the next in their tensors (0 = head_dim). With 96 and head_dim 128
     * the 128-wide tile is loaded from column h*96 and only the first 96 columns enter Q.K^T (6 of the 8 k-steps): the
     * remaining columns (the next head, or TMA zero fill past the last one) are never read by the MMA. Varlen causal only. */
    int32_t qk_head_stride;
} mjv_attn_desc;

int mjv_attention(const mjv_attn_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Row-wise primitives (test hooks; the engine calls the same kernels internally)
 *   mjv_layernorm               nn.LayerNorm            modeling_intern_vit.py:291,293
 *   mjv_pixel_shuffle_layernorm pixel_shuffle + mlp1[0]  modeling_internvl_chat.py:228-262,136 (ps_v1 != 0: ps_version
 *                               'v1', the transposed token grid the reference produces when it skips the final permute)
 *   mjv_rmsnorm                 InternLM2RMSNorm        modeling_internlm2.py:138-143 (row_idx gathers source rows)
 *   mjv_embed_tokens            tok_embeddings lookup   modeling_internvl_chat.py:163 (IMG_CONTEXT rows skipped)
 *   mjv_im2col_patches          nn.Conv2d(3,C,14,14) operand layout  modeling_intern_vit.py:164
 * ---------------------------------------------------------------------------------------------------------- */
int mjv_layernorm(const void* x, const void* w, const void* b, void* out, int64_t rows, int32_t C, float eps,
                  void* stream);
int mjv_pixel_shuffle_layernorm(const void* vit_hidden, const void* w, const void* b, void* out, int32_t tiles,
                                int32_t seq, int32_t c_vit, float eps, int32_t ps_v1, void* stream);
int mjv_rmsnorm(const void* x, const void* w, void* out, int64_t rows, int32_t C, float eps, const int32_t* row_idx,
                void* stream);
int mjv_embed_tokens(const int32_t* ids, const void* table, void* out, int64_t rows, int32_t C, int32_t img_ctx_id,
                     int32_t vocab, void* stream);
int mjv_im2col_patches(const void* px, int32_t px_dtype, void* out, int32_t T, int32_t img, int32_t ps, int32_t Kpad,
                       void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Frame path: uint8 [n_frames, H, W, 3] frames (device) -> normalised CHW tiles [n_frames*tiles_per_frame,3,S,S].
 * Replaces the per-frame body of load_video (data_processor/data.py:158-179): dynamic_preprocess (data.py:81-117,
 * PIL bicubic resize + row-major crops + thumbnail) and build_transform (data.py:56-64). Bit-exact with Pillow.
 * `out` NULL = query tiles_per_frame only. A resize runs in ONE fused kernel (horizontal taps into shared memory, vertical
 * taps + crop + normalisation in the same CTA) and needs no scratch. `scratch` (uint8, >= n_frames*H*max(S*gi,S)*3 bytes)
 * is used only by the two-pass fallback (a block that does not fit in shared memory, or MJV_RESIZE_TWO_PASS=1 in the
 * environment): without it that case returns MJV_ERR_NOMEM, and the caller retries with the buffer.
 * mjv_choose_grid / mjv_bicubic_coeffs are pure host helpers (usable without a GPU).
 * ---------------------------------------------------------------------------------------------------------- */
int mjv_preprocess_frames(const void* frames_u8, int32_t n_frames, int32_t height, int32_t width, int32_t image_size,
                          int32_t max_num, int32_t use_thumbnail, void* out, int32_t out_dtype, void* scratch,
                          int64_t scratch_bytes, int32_t* tiles_per_frame, void* stream);
int mjv_choose_grid(int32_t width, int32_t height, int32_t min_num, int32_t max_num, int32_t image_size, int32_t* gi,
                    int32_t* gj);
int mjv_bicubic_coeffs(int32_t in_size, int32_t out_size, int32_t* ksize, int32_t* bounds, int32_t* kk,
                       int64_t kk_capacity);

/* ------------------------------------------------------------------------------------------------------------
 * Model handle: one per (device, model). Replaces InternVLChatRewardModeling.__init__ / load_state_dict
 * (moe_reward.py:138-181; eval_genai_mjvideo.py:91-115) and .forward (moe_reward.py:183-297).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct mjv_handle mjv_handle;

enum { MJV_LLM_INTERNLM2 = 0, MJV_LLM_LLAMA = 1 };

typedef struct mjv_config {
    int32_t image_size, patch_size;
    int32_t vit_hidden, vit_heads, vit_mlp, vit_layers; /* vit_layers: encoder layers actually APPLIED (select_layer) */
    float vit_eps;
    int32_t llm_arch;
    int32_t llm_hidden, llm_inter, llm_layers, llm_heads, llm_kv_heads, vocab;
    float rms_eps;
    int32_t max_pos;
    float mlp1_eps;
    int32_t num_objectives, num_aspects, gating_hidden, gating_n_hidden;
    float gating_temperature;
    int32_t group_offsets[33]; /* aspect2criteria in dict order: items of aspect a are                      */
    int32_t group_items[32];   /*   group_items[group_offsets[a] .. group_offsets[a+1])                     */
    int32_t ps_version;        /* pixel-shuffle layout: 0 or 2 = 'v2', 1 = 'v1' (modeling_internvl_chat.py:237-241)   */
    int32_t precision;         /* 0 = bf16 product path (tcgen05 kernels, fp32 accumulation);
                                  1 = fp32 VERIFICATION mode: every weight and activation in fp32, one plain fp32 CUDA
                                  kernel per reference op (csrc/engine_fp32.cuh) - what a float32
                                  InternVLChatRewardModeling computes (moe_reward.py:183-297), gate 1e-4 abs. In this mode
                                  pixel_values may be fp32 or bf16, mjv_outputs.hidden_state / prompt_embedding are
                                  fp32 [B, D], mjv_gating_forward takes and returns fp32, and the bf16-only primitives
                                  and stage entry points return MJV_ERR_STATE.                                          */
} mjv_config;

int mjv_create(const mjv_config* cfg, mjv_handle** out);
void mjv_destroy(mjv_handle* h);
/* Copy + re-pack one tensor of the reference state dict (SURVEY §A.2 key names; also "rope.cos"/"rope.sin"
 * [max_pos, head_dim/2] fp32 tables built by the host shim exactly as InternLM2RotaryEmbedding does).
 * `data` may be host or device memory. Unused keys (e.g. language_model.output.weight) are accepted and dropped.
 * Llama-style separate q/k/v and gate/up keys, InternLM2's fused wqkv, and Phi-3's fused `self_attn.qkv_proj.weight`
 * (rows q|k|v) / `mlp.gate_up_proj.weight` (rows gate|up) are all accepted. */
int mjv_load_weight(mjv_handle* h, const char* key, const void* data, const int64_t* shape, int32_t ndim,
                    int32_t dtype);
/* Check that every tensor the forward needs is present; returns MJV_ERR_STATE and names the first missing key. */
int mjv_finalize_weights(mjv_handle* h);
int64_t mjv_workspace_bytes(const mjv_handle* h, int32_t n_rows, int32_t n_tiles, int64_t n_tokens);

/* One batch, already indexed by the host shim from the CPU copy of input_ids (no device->host sync on the path):
 * rows are PACKED (padding removed): sequence b owns packed token rows [cu_seqlens[b], cu_seqlens[b+1]). */
typedef struct mjv_batch {
    int32_t n_rows;            /* B samples                                                                  */
    int32_t n_tiles;           /* T tiles in pixel_values                                                     */
    int32_t n_img_tokens;      /* number of <IMG_CONTEXT> tokens that receive a projector row (<= T*tokens)   */
    int32_t n_work;            /* entries of `work`                                                           */
    int32_t img_context_token_id; /* rows holding this id are written by the projector, not the embedding table */
    int64_t n_tokens;          /* M = sum of real sequence lengths                                            */
    int64_t vt_cols;           /* columns of the transposed-V buffer (sum of 8-aligned lengths)               */
    double attn_flops;         /* algorithmic FLOPs of ONE causal attention launch: sum_b 4*H*hd*n_b(n_b+1)/2 (profiler only) */
    const void* pixel_values;  /* device [T,3,S,S]                                                            */
    int32_t pixel_dtype;       /* MJV_DTYPE_BF16 or MJV_DTYPE_F32                                             */
    const int32_t* token_ids;  /* device [M] packed token ids                                                 */
    const int32_t* row_pos;    /* device [M] position inside the sequence                                     */
    const int32_t* row_vt;     /* device [M] transposed-V column                                              */
    const int32_t* img_rows;   /* device [n_img_tokens] packed row of the i-th <IMG_CONTEXT> token (row-major)*/
    const int32_t* cu_seqlens; /* device [B+1]                                                                */
    const int32_t* vt_starts;  /* device [B]                                                                  */
    const int32_t* work;       /* device [2*n_work] (sequence, 128-row query block), heavy first              */
    const int32_t* gather_rows;/* device [2B]: packed rows of the last real token (moe_reward.py:226-228) then
                                  of the gating token (find_token_for_gating, moe_reward.py:50-57)            */
    const int32_t* work_last;  /* device [2*n_work_last]: the (sequence, query block) items that contain a gathered
                                  row — the only attention blocks the LAST decoder layer needs (NULL = all)    */
    int32_t n_work_last;
    int32_t max_seq_len;       /* longest sequence of the batch: must not exceed the rows of the loaded RoPE tables
                                  (max_position_embeddings, or more after the host re-based them: dynamic NTK)   */
} mjv_batch;

typedef struct mjv_outputs {
    float* scores;   /* device [B, 34]: rewards(28) | aspect_scores(5) | score(1)  (NO | NA | 1 in general)   */
    float* extras;   /* device [B, NO+NA+NO+1]: criteria_gating_output | aspect_gating_output |
                        aspect_weights | weighted_scores                                                      */
    void* hidden_state;      /* device bf16 [B, D] (moe_reward.py:236-237)                                    */
    void* prompt_embedding;  /* device bf16 [B, D] (moe_reward.py:243)                                        */
} mjv_outputs;

/* stage entry points (each separately testable / profilable); all enqueue on `stream` and return immediately */
int mjv_vit_forward(mjv_handle* h, const void* pixel_values, int32_t pixel_dtype, int32_t n_tiles, void* vit_hidden,
                    void* workspace, int64_t workspace_bytes, void* stream);
int mjv_project_and_splice(mjv_handle* h, const void* vit_hidden, const mjv_batch* batch, void* embeds,
                           void* workspace, int64_t workspace_bytes, void* stream);
int mjv_llm_prefill(mjv_handle* h, void* embeds_inout, const mjv_batch* batch, void* gathered_2b,
                    void* workspace, int64_t workspace_bytes, void* stream);
int mjv_reward_head(mjv_handle* h, const void* gathered_2b, int32_t n_rows, const mjv_outputs* out, void* workspace,
                    int64_t workspace_bytes, void* stream);
/* GatingNetwork as a standalone operator (moe_reward.py:16-42): which = 0 aspect_gating, 1 criteria_gating;
 * x [n_rows, llm_hidden] bf16 -> out [n_rows, num_aspects | num_objectives] bf16: softmax(MLP(x) / T) * logit_scale
 * (GatingNetwork.forward, :29-36) or the raw logits when with_softmax = 0 (forward_wo_softmax, :38-42).
 * workspace: >= mjv_workspace_bytes(h, n_rows, 0, 0). */
int mjv_gating_forward(mjv_handle* h, int32_t which, const void* x, int32_t n_rows, int32_t with_softmax, void* out,
                       void* workspace, int64_t workspace_bytes, void* stream);
/* the whole forward: ViT -> projector -> splice -> LLM prefill -> head */
int mjv_score_forward(mjv_handle* h, const mjv_batch* batch, const mjv_outputs* out, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU: the ONE collective of the path. The reference issues no collective at all (its process group of
 * size 1 exists only for torch.distributed.get_rank(): modeling_internvl_chat.py:172, eval_genai_mjvideo.py:24); sharding
 * preference pairs over ranks adds an equal-count all-gather of the fp32 score table [rows_per_rank, width]
 * (width = num_objectives + num_aspects + 1; a rank with fewer rows pads). NCCL is resolved at run time (dlopen).
 *   mjv_comm_unique_id : rank 0 fills a 128-byte id which the caller distributes (torch.distributed store, MPI, ...)
 *   mjv_comm_create    : every rank, with its CUDA device current; returns an opaque communicator (an ncclComm_t)
 *   mjv_allgather_scores: enqueued on `stream`; out = [world * rows_per_rank, width], rank-major
 * A caller that already owns an ncclComm_t may pass it to mjv_allgather_scores directly.
 * ---------------------------------------------------------------------------------------------------------- */
int mjv_comm_unique_id(void* id_128_bytes);
int mjv_comm_create(const void* id_128_bytes, int32_t rank, int32_t world, void** comm_out);
int mjv_comm_destroy(void* comm);
int mjv_allgather_scores(void* comm, const float* local_rows, int64_t rows_per_rank, int32_t width, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Launch accounting and per-launch timing (used by bench.py; see csrc/launch.cuh for the tag list).
 * mjv_profile_enable(mask): bracket every launch whose tag bit is set with two CUDA events on the launching stream.
 * mjv_profile_read: elapsed ms + algorithmic work (FLOPs or bytes) per recorded launch; clears the records.
 * ---------------------------------------------------------------------------------------------------------- */
long long mjv_launch_count(void);
void mjv_reset_launch_count(void);
void mjv_profile_enable(uint32_t tag_mask);
int mjv_profile_read(int32_t* tags, float* ms, double* work, int32_t capacity);
const char* mjv_kernel_tag_name(int32_t tag);
int mjv_kernel_tag_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MJV_B200_H_ */
