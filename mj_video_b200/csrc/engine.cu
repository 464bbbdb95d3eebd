// Model handle, weight ingestion/re-packing and the forward orchestration (ViT -> projector/splice -> LLM prefill ->
// reward head). Host-side C++; every device op is one of this library's own kernels. Mirrors
// InternVLChatRewardModeling.forward (scripts/model/moe_reward.py:183-297) and what it calls:
//   InternVLChatModel.forward / extract_feature   modeling_internvl_chat.py:146-262
//   InternVisionModel / EncoderLayer              modeling_intern_vit.py:162-174, 283-295
//   InternLM2Model / DecoderLayer                 modeling_internlm2.py:854-984, 651-671
// Deliberately NOT computed: the LM-head logits (modeling_internlm2.py:1081-1082) — their result is never read by the
// reward path — and the final norm of rows other than the 2 per sample the head gathers (moe_reward.py:236-243).
#include <math.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "head.cuh"
#include "kernels.cuh"
#include "launch.cuh"

namespace mjv {

struct Tensor {
    void* ptr = nullptr;
    long long rows = 0, cols = 0, ld = 0;
    bool is_float = false;
    long long filled_rows = 0;
    long long expect_rows = -1;  // source rows that make the tensor complete (-1: all of `rows`)
};

struct VitLayerW {
    const void *ls1, *ls2, *qkv_w, *qkv_b, *proj_w, *proj_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b, *n1_w, *n1_b, *n2_w, *n2_b;
};
struct LlmLayerW {
    const void *wqkv, *wo, *w13, *w2, *attn_norm, *ffn_norm;
};
struct GateW {
    std::vector<const void*> w, b;
    const void* logit_scale;
};

}  // namespace mjv

struct mjv_handle {
    mjv_config cfg;
    std::unordered_map<std::string, mjv::Tensor> weights;
    std::unordered_map<std::string, mjv::Tensor> weights32;  // precision == 1: fp32 copies in the reference's own layouts
    bool finalized = false;
    int kpad = 0;
    // resolved pointers
    const void *cls = nullptr, *pos = nullptr, *patch_w = nullptr, *patch_b = nullptr;
    std::vector<mjv::VitLayerW> vit;
    const void *mlp1_ln_w = nullptr, *mlp1_ln_b = nullptr, *mlp1_1_w = nullptr, *mlp1_1_b = nullptr, *mlp1_3_w = nullptr,
               *mlp1_3_b = nullptr;
    const void* tok_emb = nullptr;
    std::vector<mjv::LlmLayerW> llm;
    const void* final_norm = nullptr;
    const float *rope_cos = nullptr, *rope_sin = nullptr;
    long long rope_rows = 0;
    const void *regression = nullptr, *transform = nullptr;
    mjv::GateW aspect, criteria;
    int *group_offsets_dev = nullptr, *group_items_dev = nullptr;
};

namespace mjv {
namespace {

inline long long round_up(long long x, long long a) { return (x + a - 1) / a * a; }

struct Arena {
    char* base;
    size_t cap, off;
    Arena(void* p, size_t bytes) : base(static_cast<char*>(p)), cap(bytes), off(0) {}
    void* alloc(size_t bytes) {
        size_t a = (off + 1023) & ~size_t(1023);
        if (a + bytes > cap) return nullptr;
        off = a + bytes;
        return base + a;
    }
};
#define MJV_ALLOC(var, type, arena, bytes)                                                                   \
    type var = static_cast<type>((arena).alloc(bytes));                                                      \
    if (var == nullptr)                                                                                      \
        return set_error(MJV_ERR_NOMEM, "workspace too small: need %zu more bytes at %s:%d", (size_t)(bytes), \
                         __FILE__, __LINE__)

// fp32 verification mode (engine_fp32.cuh, included further down once Dims / al exist)
int load_weight_f32(mjv_handle* h, const std::string& key, const void* data, long long rows, long long cols, int dtype);
int finalize_f32(mjv_handle* h);

// ------------------------------------------------------------------------------------------------ weight ingestion
__global__ void convert_rows_kernel(const void* __restrict__ src, int src_dtype, long long rows, long long cols,
                                    __nv_bfloat16* __restrict__ dst, long long dst_ld,
                                    const int* __restrict__ dst_row_map, const int* __restrict__ dst_col_map) {
    const long long total = rows * cols;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / cols, c = i - r * cols;
        float v;
        if (src_dtype == MJV_DTYPE_F32) v = static_cast<const float*>(src)[i];
        else if (src_dtype == MJV_DTYPE_BF16) v = __bfloat162float(static_cast<const __nv_bfloat16*>(src)[i]);
        else v = __half2float(static_cast<const __half*>(src)[i]);
        const long long dr = dst_row_map != nullptr ? dst_row_map[r] : r;
        const long long dc = dst_col_map != nullptr ? dst_col_map[c] : c;
        dst[dr * dst_ld + dc] = __float2bfloat16_rn(v);
    }
}

size_t dtype_size(int dt) { return dt == MJV_DTYPE_F32 ? 4 : 2; }

// LLM heads narrower than 128 (e.g. 96 for the Phi-3-mini-shaped backbone of BASELINE configs[4]) are stored in a
// 128-wide slot so that the head_dim-128 RoPE epilogue and attention kernel serve them unchanged: the two rotary
// halves go to columns [0, hd/2) and [64, 64 + hd/2) (rotate-half pairs column c with c + 64), V keeps its order, and
// every pad column comes from an all-zero weight row, i.e. is exactly 0 in q, k, v and the attention output. The
// permutation is applied to q and k alike, so q.k is unchanged; `wo` gets matching zero columns.
constexpr int kHeadSlot = 128;
// head_dim 96 (Phi-3-mini shapes, BASELINE configs[4]) gets a NATIVE 96-wide slot when the q and k column counts stay
// multiples of the 128-column epilogue group (heads and kv heads multiples of 4): the rotary pair (d, d + 48) is stored in
// adjacent columns (2d, 2d + 1), the QKV GEMM has no pad columns (N 11264 -> 9216 at 32 + 32 + 32 heads) and attention
// feeds only 6 of the 8 k-steps of its 128-wide Q / K tiles to the tensor core (mjv_attn_desc.qk_head_stride).
inline int qk_head_slot(const mjv_config& c) {
    const int hd = c.llm_hidden / c.llm_heads;
    return (hd == 96 && c.llm_heads % 4 == 0 && c.llm_kv_heads % 4 == 0) ? 96 : kHeadSlot;
}
inline int qk_slot_dim(int d, int hd, int hs) {
    if (hs == 96) return d < hd / 2 ? 2 * d : 2 * (d - hd / 2) + 1;
    return d < hd / 2 ? d : 64 + (d - hd / 2);
}
// V needs no rotary slot and the attention output no padding: when head_dim is a multiple of 32 (e.g. 96) V heads are
// packed in the QKV GEMM's N (if the total stays a multiple of the 128-column tile) and the attention kernel writes only
// head_dim columns per head, so `wo` keeps its original K.
inline int v_head_width(const mjv_config& c) {
    const int hd = c.llm_hidden / c.llm_heads;
    if (hd % 32 != 0) return kHeadSlot;
    const long long qk = static_cast<long long>(c.llm_heads + c.llm_kv_heads) * qk_head_slot(c);
    const long long packed = qk + static_cast<long long>(c.llm_kv_heads) * hd;
    const long long padded = qk + static_cast<long long>(c.llm_kv_heads) * kHeadSlot;
    // keep the 256-column CTA-pair tile whenever the padded layout would have it
    if (packed % 256 == 0) return hd;
    if (padded % 256 != 0 && packed % 128 == 0) return hd;
    return kHeadSlot;
}
inline int o_head_width(const mjv_config& c) {
    const int hd = c.llm_hidden / c.llm_heads;
    return hd % 32 == 0 ? hd : kHeadSlot;
}

bool ends_with(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}
void replace_all(std::string& s, const std::string& a, const std::string& b) {
    size_t p = 0;
    while ((p = s.find(a, p)) != std::string::npos) {
        s.replace(p, a.size(), b);
        p += b.size();
    }
}

// HF-Llama key names -> the InternLM2 names used internally (SURVEY §A.2 last paragraph)
std::string canonical_key(std::string k) {
    replace_all(k, ".embed_tokens.", ".tok_embeddings.");
    replace_all(k, ".self_attn.o_proj.", ".attention.wo.");
    replace_all(k, ".self_attn.q_proj.", ".attention.wq.");
    replace_all(k, ".self_attn.k_proj.", ".attention.wk.");
    replace_all(k, ".self_attn.v_proj.", ".attention.wv.");
    replace_all(k, ".mlp.gate_proj.", ".feed_forward.w1.");
    replace_all(k, ".mlp.up_proj.", ".feed_forward.w3.");
    replace_all(k, ".mlp.down_proj.", ".feed_forward.w2.");
    replace_all(k, ".input_layernorm.", ".attention_norm.");
    replace_all(k, ".post_attention_layernorm.", ".ffn_norm.");
    replace_all(k, "language_model.lm_head.", "language_model.output.");
    return k;
}

int ensure_tensor(mjv_handle* h, const std::string& key, long long rows, long long cols, long long ld, bool is_float,
                  Tensor** out) {
    auto it = h->weights.find(key);
    if (it == h->weights.end()) {
        Tensor t;
        t.rows = rows;
        t.cols = cols;
        t.ld = ld;
        t.is_float = is_float;
        const size_t bytes = static_cast<size_t>(rows) * ld * (is_float ? 4 : 2);
        MJV_CUDA_OK(cudaMalloc(&t.ptr, bytes));
        MJV_CUDA_OK(cudaMemset(t.ptr, 0, bytes));
        it = h->weights.emplace(key, t).first;
    } else {
        MJV_CHECK(it->second.rows == rows && it->second.cols == cols, MJV_ERR_ARG,
                  "weight %s: shape mismatch with an earlier part", key.c_str());
    }
    *out = &it->second;
    return 0;
}

int load_weight(mjv_handle* h, const char* key_c, const void* data, const int64_t* shape, int ndim, int dtype) {
    MJV_CHECK(h && key_c && data && shape && ndim >= 1 && ndim <= 4, MJV_ERR_ARG, "mjv_load_weight: bad arguments");
    MJV_CHECK(dtype == MJV_DTYPE_F32 || dtype == MJV_DTYPE_BF16 || dtype == MJV_DTYPE_F16, MJV_ERR_ARG,
              "mjv_load_weight: unsupported dtype %d", dtype);
    const mjv_config& c = h->cfg;
    std::string key = canonical_key(key_c);
    long long rows = shape[0], cols = 1;
    for (int i = 1; i < ndim; ++i) cols *= shape[i];
    if (ndim == 1) {
        cols = rows;
        rows = 1;
    }
    if (ends_with(key, "language_model.output.weight")) return 0;  // LM head: never used by the reward path
    // Phi-3 checkpoints (InternVL2-4B; reference README.md:18) ship q|k|v and gate|up as single matrices
    // (transformers Phi3Attention.qkv_proj: rows [q heads | k heads | v heads]; Phi3MLP.gate_up_proj: rows [gate | up], chunk(2)).
    // Row ranges of a row-major matrix are contiguous: hand them on under the separate Llama-style names.
    {
        const std::string raw(key_c);
        const size_t esz = dtype_size(dtype);
        const char* bytes = static_cast<const char*>(data);
        const auto sub = [&](const std::string& from, const std::string& to, long long r0, long long nr) {
            std::string k2 = raw;
            replace_all(k2, from, to);
            int64_t shp[2] = {nr, cols};
            return load_weight(h, k2.c_str(), bytes + static_cast<size_t>(r0) * cols * esz, shp, 2, dtype);
        };
        if (ends_with(raw, ".self_attn.qkv_proj.weight")) {
            const long long hd_ = c.llm_hidden / c.llm_heads, nq = c.llm_heads * hd_, nk = c.llm_kv_heads * hd_;
            MJV_CHECK(ndim == 2 && rows == nq + 2 * nk, MJV_ERR_ARG, "%s: %lld rows, expected %lld", key_c, rows, nq + 2 * nk);
            MJV_PROPAGATE(sub(".qkv_proj.", ".q_proj.", 0, nq));
            MJV_PROPAGATE(sub(".qkv_proj.", ".k_proj.", nq, nk));
            return sub(".qkv_proj.", ".v_proj.", nq + nk, nk);
        }
        if (ends_with(raw, ".mlp.gate_up_proj.weight")) {
            MJV_CHECK(ndim == 2 && rows == 2LL * c.llm_inter, MJV_ERR_ARG, "%s: %lld rows, expected %lld", key_c, rows,
                      2LL * c.llm_inter);
            MJV_PROPAGATE(sub(".gate_up_proj.", ".gate_proj.", 0, c.llm_inter));
            return sub(".gate_up_proj.", ".up_proj.", c.llm_inter, c.llm_inter);
        }
    }
    const long long n_elem = rows * cols;
    h->finalized = false;
    if (c.precision == 1) return load_weight_f32(h, key, data, rows, cols, dtype);  // verification mode: no re-packing

    // rope tables stay fp32
    if (key == "rope.cos" || key == "rope.sin") {
        MJV_CHECK(dtype == MJV_DTYPE_F32, MJV_ERR_ARG, "%s must be float32", key.c_str());
        // the table may be re-loaded with more rows (dynamic-NTK re-basing for a batch longer than max_pos): replace it
        auto old = h->weights.find(key);
        if (old != h->weights.end() && (old->second.rows != rows || old->second.cols != cols)) {
            MJV_CUDA_OK(cudaDeviceSynchronize());  // launches that still read the old table
            cudaFree(old->second.ptr);
            h->weights.erase(old);
        }
        Tensor* t;
        MJV_PROPAGATE(ensure_tensor(h, key, rows, cols, cols, true, &t));
        MJV_CUDA_OK(cudaMemcpy(t->ptr, data, n_elem * 4, cudaMemcpyDefault));
        t->filled_rows = rows;
        return 0;
    }

    const int hd = c.llm_hidden / c.llm_heads;
    const int hs = qk_head_slot(c);
    const int hv = v_head_width(c), ho = o_head_width(c);
    const long long v_base = static_cast<long long>(c.llm_heads + c.llm_kv_heads) * hs;  // first V row of the fused weight
    const long long qkv_rows = v_base + static_cast<long long>(c.llm_kv_heads) * hv;
    std::string dst_key = key;
    long long dst_rows = rows, dst_ld = cols;
    std::vector<int> row_map, col_map;
    if (ends_with(key, "patch_embedding.weight")) {
        dst_ld = h->kpad;
    } else if (ends_with(key, "attention.wqkv.weight")) {
        // InternLM2 layout: per kv group [q x groups, k, v] x head_dim (modeling_internlm2.py:463-473)
        const int g = c.llm_heads / c.llm_kv_heads;
        MJV_CHECK(rows == static_cast<long long>(c.llm_heads + 2 * c.llm_kv_heads) * hd, MJV_ERR_ARG,
                  "%s: unexpected rows %lld", key.c_str(), rows);
        dst_rows = qkv_rows;
        row_map.resize(rows);
        for (int kv = 0; kv < c.llm_kv_heads; ++kv)
            for (int s = 0; s < g + 2; ++s)
                for (int d = 0; d < hd; ++d) {
                    const int src = (kv * (g + 2) + s) * hd + d;
                    int dst;
                    if (s < g) dst = (kv * g + s) * hs + qk_slot_dim(d, hd, hs);
                    else if (s == g) dst = (c.llm_heads + kv) * hs + qk_slot_dim(d, hd, hs);
                    else dst = static_cast<int>(v_base) + kv * hv + d;
                    row_map[src] = dst;
                }
    } else if (ends_with(key, "attention.wq.weight") || ends_with(key, "attention.wk.weight") ||
               ends_with(key, "attention.wv.weight")) {
        const char which = key[key.size() - 8];  // 'q' / 'k' / 'v'
        dst_key = key.substr(0, key.size() - 9) + "wqkv.weight";
        dst_rows = qkv_rows;
        const int n_heads = which == 'q' ? c.llm_heads : c.llm_kv_heads;
        MJV_CHECK(rows == static_cast<long long>(n_heads) * hd, MJV_ERR_ARG, "%s: unexpected rows %lld", key.c_str(), rows);
        const int base = which == 'q' ? 0 : (which == 'k' ? c.llm_heads : c.llm_heads + c.llm_kv_heads);
        row_map.resize(rows);
        for (long long r = 0; r < rows; ++r) {
            const int head = static_cast<int>(r / hd), d = static_cast<int>(r % hd);
            row_map[r] = which == 'v' ? static_cast<int>(v_base) + head * hv + d : (base + head) * hs + qk_slot_dim(d, hd, hs);
        }
    } else if (ends_with(key, "attention.wo.weight")) {
        MJV_CHECK(cols == static_cast<long long>(c.llm_heads) * hd, MJV_ERR_ARG, "%s: unexpected cols %lld", key.c_str(), cols);
        dst_ld = static_cast<long long>(c.llm_heads) * ho;
        if (hd != ho) {
            col_map.resize(cols);
            for (long long cc = 0; cc < cols; ++cc) col_map[cc] = static_cast<int>((cc / hd) * ho + cc % hd);
        }
    } else if (ends_with(key, "feed_forward.w1.weight") || ends_with(key, "feed_forward.w3.weight")) {
        // gate/up interleaved in 32-row chunks so one GEMM tile holds both halves of SwiGLU
        const bool is_up = key[key.size() - 8] == '3';
        MJV_CHECK(rows % 32 == 0, MJV_ERR_ARG, "%s: rows %lld not a multiple of 32", key.c_str(), rows);
        dst_key = key.substr(0, key.size() - 9) + "w13.weight";
        dst_rows = 2 * rows;
        row_map.resize(rows);
        for (long long r = 0; r < rows; ++r) row_map[r] = static_cast<int>((r / 32) * 64 + (is_up ? 32 : 0) + r % 32);
    }

    Tensor* t;
    MJV_PROPAGATE(ensure_tensor(h, dst_key, dst_rows, cols, dst_ld, false, &t));
    if (ends_with(dst_key, "attention.wqkv.weight"))
        t->expect_rows = static_cast<long long>(c.llm_heads + 2 * c.llm_kv_heads) * hd;
    void* staging = nullptr;
    MJV_CUDA_OK(cudaMalloc(&staging, n_elem * dtype_size(dtype)));
    int *map_dev = nullptr, *cmap_dev = nullptr;
    cudaError_t e = cudaMemcpy(staging, data, n_elem * dtype_size(dtype), cudaMemcpyDefault);
    if (e == cudaSuccess && !row_map.empty()) {
        e = cudaMalloc(&map_dev, row_map.size() * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpy(map_dev, row_map.data(), row_map.size() * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && !col_map.empty()) {
        e = cudaMalloc(&cmap_dev, col_map.size() * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpy(cmap_dev, col_map.data(), col_map.size() * sizeof(int), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        convert_rows_kernel<<<sm_count() * 4, 256>>>(staging, dtype, rows, cols, static_cast<__nv_bfloat16*>(t->ptr),
                                                     t->ld, map_dev, cmap_dev);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();  // load time only; the compute path never synchronises
    }
    cudaFree(staging);
    if (map_dev) cudaFree(map_dev);
    if (cmap_dev) cudaFree(cmap_dev);
    if (e != cudaSuccess) return set_error(MJV_ERR_CUDA, "mjv_load_weight(%s): %s", key_c, cudaGetErrorString(e));
    t->filled_rows += rows;
    return 0;
}

int need(mjv_handle* h, const std::string& key, const void** out, long long rows = -1) {
    auto it = h->weights.find(key);
    if (it == h->weights.end()) return set_error(MJV_ERR_STATE, "missing weight: %s", key.c_str());
    const long long want = it->second.expect_rows >= 0 ? it->second.expect_rows : it->second.rows;
    if (it->second.filled_rows < want)
        return set_error(MJV_ERR_STATE, "weight %s only partially loaded (%lld of %lld rows)", key.c_str(),
                         it->second.filled_rows, want);
    if (rows >= 0 && it->second.rows * it->second.cols != rows)
        return set_error(MJV_ERR_STATE, "weight %s has %lld elements, expected %lld", key.c_str(),
                         it->second.rows * it->second.cols, rows);
    *out = it->second.ptr;
    return 0;
}

int finalize(mjv_handle* h) {
    const mjv_config& c = h->cfg;
    if (c.precision == 1) return finalize_f32(h);
    const std::string v = "model.vision_model.";
    MJV_PROPAGATE(need(h, v + "embeddings.class_embedding", &h->cls));
    MJV_PROPAGATE(need(h, v + "embeddings.position_embedding", &h->pos));
    MJV_PROPAGATE(need(h, v + "embeddings.patch_embedding.weight", &h->patch_w));
    MJV_PROPAGATE(need(h, v + "embeddings.patch_embedding.bias", &h->patch_b));
    h->vit.resize(c.vit_layers);
    for (int i = 0; i < c.vit_layers; ++i) {
        const std::string p = v + "encoder.layers." + std::to_string(i) + ".";
        VitLayerW& w = h->vit[i];
        MJV_PROPAGATE(need(h, p + "ls1", &w.ls1));
        MJV_PROPAGATE(need(h, p + "ls2", &w.ls2));
        MJV_PROPAGATE(need(h, p + "attn.qkv.weight", &w.qkv_w));
        MJV_PROPAGATE(need(h, p + "attn.qkv.bias", &w.qkv_b));
        MJV_PROPAGATE(need(h, p + "attn.proj.weight", &w.proj_w));
        MJV_PROPAGATE(need(h, p + "attn.proj.bias", &w.proj_b));
        MJV_PROPAGATE(need(h, p + "mlp.fc1.weight", &w.fc1_w));
        MJV_PROPAGATE(need(h, p + "mlp.fc1.bias", &w.fc1_b));
        MJV_PROPAGATE(need(h, p + "mlp.fc2.weight", &w.fc2_w));
        MJV_PROPAGATE(need(h, p + "mlp.fc2.bias", &w.fc2_b));
        MJV_PROPAGATE(need(h, p + "norm1.weight", &w.n1_w));
        MJV_PROPAGATE(need(h, p + "norm1.bias", &w.n1_b));
        MJV_PROPAGATE(need(h, p + "norm2.weight", &w.n2_w));
        MJV_PROPAGATE(need(h, p + "norm2.bias", &w.n2_b));
    }
    MJV_PROPAGATE(need(h, "model.mlp1.0.weight", &h->mlp1_ln_w));
    MJV_PROPAGATE(need(h, "model.mlp1.0.bias", &h->mlp1_ln_b));
    MJV_PROPAGATE(need(h, "model.mlp1.1.weight", &h->mlp1_1_w));
    MJV_PROPAGATE(need(h, "model.mlp1.1.bias", &h->mlp1_1_b));
    MJV_PROPAGATE(need(h, "model.mlp1.3.weight", &h->mlp1_3_w));
    MJV_PROPAGATE(need(h, "model.mlp1.3.bias", &h->mlp1_3_b));
    const std::string lm = "model.language_model.model.";
    MJV_PROPAGATE(need(h, lm + "tok_embeddings.weight", &h->tok_emb));
    h->llm.resize(c.llm_layers);
    for (int i = 0; i < c.llm_layers; ++i) {
        const std::string p = lm + "layers." + std::to_string(i) + ".";
        LlmLayerW& w = h->llm[i];
        MJV_PROPAGATE(need(h, p + "attention.wqkv.weight", &w.wqkv));
        MJV_PROPAGATE(need(h, p + "attention.wo.weight", &w.wo));
        MJV_PROPAGATE(need(h, p + "feed_forward.w13.weight", &w.w13));
        MJV_PROPAGATE(need(h, p + "feed_forward.w2.weight", &w.w2));
        MJV_PROPAGATE(need(h, p + "attention_norm.weight", &w.attn_norm));
        MJV_PROPAGATE(need(h, p + "ffn_norm.weight", &w.ffn_norm));
    }
    MJV_PROPAGATE(need(h, lm + "norm.weight", &h->final_norm));
    const void* tmp;
    // [max_pos, 64]: one column per rotary pair of the 128-wide head slot (pad pairs: cos 1, sin 0)
    // (rows >= max_pos: the host re-loads a longer, re-based table when a batch exceeds max_position_embeddings)
    MJV_PROPAGATE(need(h, "rope.cos", &tmp));
    h->rope_cos = static_cast<const float*>(tmp);
    MJV_PROPAGATE(need(h, "rope.sin", &tmp));
    h->rope_sin = static_cast<const float*>(tmp);
    {
        const Tensor &tc = h->weights.at("rope.cos"), &ts = h->weights.at("rope.sin");
        MJV_CHECK(tc.cols == kHeadSlot / 2 && ts.cols == kHeadSlot / 2 && tc.rows == ts.rows && tc.rows >= c.max_pos,
                  MJV_ERR_STATE, "rope tables must be [>= %d, %d] float32", c.max_pos, kHeadSlot / 2);
        h->rope_rows = tc.rows;
    }
    MJV_PROPAGATE(need(h, "regression_layer.weight", &h->regression));
    MJV_PROPAGATE(need(h, "reward_transform_matrix", &h->transform));
    for (int net = 0; net < 2; ++net) {
        GateW& g = net == 0 ? h->aspect : h->criteria;
        const std::string p = net == 0 ? "aspect_gating." : "criteria_gating.";
        g.w.resize(c.gating_n_hidden + 1);
        g.b.resize(c.gating_n_hidden + 1);
        for (int l = 0; l <= c.gating_n_hidden; ++l) {
            MJV_PROPAGATE(need(h, p + "layers." + std::to_string(l) + ".weight", &g.w[l]));
            MJV_PROPAGATE(need(h, p + "layers." + std::to_string(l) + ".bias", &g.b[l]));
        }
        MJV_PROPAGATE(need(h, p + "logit_scale", &g.logit_scale));
    }
    h->finalized = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ sizes
struct Dims {
    int S, P, C, Fv, heads, Kpad, Spad, tok_per_tile;
    int D, F, Hq, Hkv, hd, hs;  // hd: model head_dim; hs: stored head slot of q / k (128)
    int hv, ho;                 // stored width of a V head (N of the QKV GEMM, rows of a V^T plane) and of an output head
};
Dims dims_of(const mjv_handle* h) {
    const mjv_config& c = h->cfg;
    Dims d;
    const int g = c.image_size / c.patch_size;
    d.P = g * g;
    d.S = d.P + 1;
    d.C = c.vit_hidden;
    d.Fv = c.vit_mlp;
    d.heads = c.vit_heads;
    d.Kpad = h->kpad;
    d.Spad = static_cast<int>(round_up(d.S, 8));
    d.tok_per_tile = d.P / 4;
    d.D = c.llm_hidden;
    d.F = c.llm_inter;
    d.Hq = c.llm_heads;
    d.Hkv = c.llm_kv_heads;
    d.hd = c.llm_hidden / c.llm_heads;
    d.hs = qk_head_slot(c);
    d.hv = v_head_width(c);
    d.ho = o_head_width(c);
    return d;
}
inline size_t al(size_t bytes) { return (bytes + 1023) & ~size_t(1023); }

}  // namespace
}  // namespace mjv
#include "engine_fp32.cuh"
namespace mjv {
namespace {

size_t vit_temp_bytes(const Dims& d, long long T) {
    const size_t Mv = static_cast<size_t>(T) * d.S;
    size_t mid = Mv * d.Fv;
    const size_t patches = static_cast<size_t>(T) * d.P * d.Kpad;
    if (patches > mid) mid = patches;
    return al(Mv * d.C * 2) + al(Mv * 2 * d.C * 2) + al(static_cast<size_t>(T) * d.heads * 64 * d.Spad * 2) +
           al(Mv * d.C * 2) + al(mid * 2) + 8192;
}
size_t proj_temp_bytes(const Dims& d, long long T) {
    const size_t Mp = static_cast<size_t>(T) * d.tok_per_tile;
    return al(Mp * 4 * d.C * 2) + al(Mp * d.D * 2) + 4096;
}
size_t llm_temp_bytes(const Dims& d, long long M, long long vt_cols, long long B) {
    const size_t m = static_cast<size_t>(M);
    const size_t g2 = static_cast<size_t>(2 * B);
    return al(m * d.D * 2) + al(m * d.Hq * d.hs * 2) + al(m * d.Hkv * d.hs * 2) +
           al(static_cast<size_t>(d.Hkv) * d.hs * vt_cols * 2) + al(m * d.Hq * d.hs * 2) + al(m * d.F * 2) +
           al(g2 * d.Hq * d.hs * 2) + 2 * al(g2 * d.D * 2) + al(g2 * d.F * 2) + 16384;
}
size_t head_temp_bytes(const mjv_config& c, long long B) {
    return 4 * al(static_cast<size_t>(B) * c.gating_hidden * 2) + 3 * al(static_cast<size_t>(B) * 32 * 2) + 8192;
}

// ------------------------------------------------------------------------------------------------ stages
mjv_gemm_desc gemm_base(const void* a, long long lda, const void* w, long long ldw, void* out, long long ldo, int M,
                        int N, int K, int epi) {
    mjv_gemm_desc g;
    memset(&g, 0, sizeof(g));
    g.a = a;
    g.lda = lda;
    g.w = w;
    g.ldw = ldw;
    g.out = out;
    g.ldo = ldo;
    g.M = M;
    g.N = N;
    g.K = K;
    g.epilogue = epi;
    return g;
}

int vit_forward(mjv_handle* h, const void* px, int px_dtype, int T, void* vit_hidden, Arena& ar, cudaStream_t st) {
    const Dims d = dims_of(h);
    const mjv_config& c = h->cfg;
    MJV_CHECK(d.C / d.heads == 64, MJV_ERR_ARG, "ViT head_dim %d unsupported (64 only)", d.C / d.heads);
    const long long Mv = static_cast<long long>(T) * d.S;
    MJV_CHECK(Mv < (1LL << 31) / 4, MJV_ERR_ARG, "too many ViT rows (%lld); chunk the tiles", Mv);
    size_t mid_elems = static_cast<size_t>(Mv) * d.Fv;
    const size_t patch_elems = static_cast<size_t>(T) * d.P * d.Kpad;
    if (patch_elems > mid_elems) mid_elems = patch_elems;
    MJV_ALLOC(ln, __nv_bfloat16*, ar, static_cast<size_t>(Mv) * d.C * 2);
    MJV_ALLOC(qk, __nv_bfloat16*, ar, static_cast<size_t>(Mv) * 2 * d.C * 2);
    MJV_ALLOC(vt, __nv_bfloat16*, ar, static_cast<size_t>(T) * d.heads * 64 * d.Spad * 2);
    MJV_ALLOC(att, __nv_bfloat16*, ar, static_cast<size_t>(Mv) * d.C * 2);
    MJV_ALLOC(mid, __nv_bfloat16*, ar, mid_elems * 2);
    __nv_bfloat16* hid = static_cast<__nv_bfloat16*>(vit_hidden);

    // patch embedding (+bias +pos-emb) and the CLS rows (InternVisionEmbeddings.forward, modeling_intern_vit.py:162-174)
    MJV_PROPAGATE(im2col_patches(px, px_dtype, mid, T, c.image_size, c.patch_size, d.Kpad, st));
    {
        mjv_gemm_desc g = gemm_base(mid, d.Kpad, h->patch_w, d.Kpad, hid, d.C, T * d.P, d.C, d.Kpad, MJV_EPI_PATCH);
        g.bias = h->patch_b;
        g.pos_emb = h->pos;
        g.seq = d.P;
        MJV_PROPAGATE(gemm_bf16(g, st));
    }
    MJV_PROPAGATE(cls_rows(h->cls, h->pos, hid, T, d.S, d.C, st));

    for (int i = 0; i < c.vit_layers; ++i) {
        const VitLayerW& w = h->vit[i];
        MJV_PROPAGATE(layernorm(hid, w.n1_w, w.n1_b, ln, Mv, d.C, c.vit_eps, st));
        {
            mjv_gemm_desc g = gemm_base(ln, d.C, w.qkv_w, d.C, qk, 2 * d.C, (int)Mv, 3 * d.C, d.C, MJV_EPI_QKV_VIT);
            g.bias = w.qkv_b;
            g.out_vt = vt;
            g.n_q_cols = d.C;
            g.n_k_cols = d.C;
            g.vt_stride_head = 64LL * d.Spad;
            g.vt_ld = d.Spad;
            g.seq = d.S;
            g.heads = d.heads;
            MJV_PROPAGATE(gemm_bf16(g, st));
        }
        {
            mjv_attn_desc a;
            memset(&a, 0, sizeof(a));
            a.q = qk;
            a.k = qk;
            a.vt = vt;
            a.out = att;
            a.ldq = a.ldk = 2 * d.C;
            a.ldo = d.C;
            a.vt_ld = d.Spad;
            a.vt_plane_stride = 64LL * d.Spad;
            a.total_rows = Mv;
            a.head_dim = 64;
            a.heads = a.kv_heads = d.heads;
            a.q_col0 = 0;
            a.k_col0 = d.C;
            a.causal = 0;
            a.nseq = T;
            a.seq_len = d.S;
            a.scale = 0.125f;  // head_dim ** -0.5 (modeling_intern_vit.py:207)
            MJV_PROPAGATE(attention(a, st));
        }
        {
            mjv_gemm_desc g = gemm_base(att, d.C, w.proj_w, d.C, hid, d.C, (int)Mv, d.C, d.C, MJV_EPI_RESID);
            g.bias = w.proj_b;
            g.scale = w.ls1;
            g.residual = hid;
            g.ldr = d.C;
            MJV_PROPAGATE(gemm_bf16(g, st));
        }
        MJV_PROPAGATE(layernorm(hid, w.n2_w, w.n2_b, ln, Mv, d.C, c.vit_eps, st));
        {
            mjv_gemm_desc g = gemm_base(ln, d.C, w.fc1_w, d.C, mid, d.Fv, (int)Mv, d.Fv, d.C, MJV_EPI_GELU);
            g.bias = w.fc1_b;
            MJV_PROPAGATE(gemm_bf16(g, st));
        }
        {
            mjv_gemm_desc g = gemm_base(mid, d.Fv, w.fc2_w, d.Fv, hid, d.C, (int)Mv, d.C, d.Fv, MJV_EPI_RESID);
            g.bias = w.fc2_b;
            g.scale = w.ls2;
            g.residual = hid;
            g.ldr = d.C;
            MJV_PROPAGATE(gemm_bf16(g, st));
        }
    }
    return 0;
}

int project_and_splice(mjv_handle* h, const void* vit_hidden, const mjv_batch& b, void* embeds, Arena& ar,
                       cudaStream_t st) {
    const Dims d = dims_of(h);
    const mjv_config& c = h->cfg;
    const int T = b.n_tiles;
    const long long Mp = static_cast<long long>(T) * d.tok_per_tile;
    MJV_CHECK(b.n_img_tokens <= Mp, MJV_ERR_ARG, "%d image tokens but only %lld projector rows", b.n_img_tokens, Mp);
    // text rows: embedding lookup; image rows are written by the projector's scatter epilogue (the splice)
    MJV_PROPAGATE(embed_tokens(b.token_ids, h->tok_emb, embeds, b.n_tokens, d.D, b.img_context_token_id, c.vocab, st));
    if (T == 0 || b.n_img_tokens == 0) return 0;
    MJV_ALLOC(pn, __nv_bfloat16*, ar, static_cast<size_t>(Mp) * 4 * d.C * 2);
    MJV_ALLOC(p1, __nv_bfloat16*, ar, static_cast<size_t>(Mp) * d.D * 2);
    MJV_PROPAGATE(pixel_shuffle_layernorm(vit_hidden, h->mlp1_ln_w, h->mlp1_ln_b, pn, T, d.S, d.C, c.mlp1_eps, c.ps_version == 1, st));
    {
        mjv_gemm_desc g = gemm_base(pn, 4 * d.C, h->mlp1_1_w, 4 * d.C, p1, d.D, (int)Mp, d.D, 4 * d.C, MJV_EPI_GELU);
        g.bias = h->mlp1_1_b;
        MJV_PROPAGATE(gemm_bf16(g, st));
    }
    {
        // only the first n_img_tokens rows are spliced (the reference truncates on a count mismatch:
        // modeling_internvl_chat.py:178-186)
        mjv_gemm_desc g = gemm_base(p1, d.D, h->mlp1_3_w, d.D, embeds, d.D, b.n_img_tokens, d.D, d.D, MJV_EPI_SCATTER);
        g.bias = h->mlp1_3_b;
        g.row_map = b.img_rows;
        MJV_PROPAGATE(gemm_bf16(g, st));
    }
    return 0;
}

int llm_prefill(mjv_handle* h, void* embeds, const mjv_batch& b, void* gathered, Arena& ar, cudaStream_t st) {
    const Dims d = dims_of(h);
    const mjv_config& c = h->cfg;
    MJV_CHECK(d.hd % 2 == 0 && d.hd >= 2 && d.hd <= d.hs, MJV_ERR_ARG, "LLM head_dim %d unsupported (even, <= 128)", d.hd);
    const long long M = b.n_tokens;
    MJV_CHECK(M > 0 && M < (1LL << 31) / 16, MJV_ERR_ARG, "bad token count %lld", M);
    MJV_CHECK(b.max_seq_len > 0 && b.max_seq_len <= h->rope_rows, MJV_ERR_ARG,
              "longest sequence (%d tokens) exceeds the loaded RoPE table (%lld rows)", b.max_seq_len, h->rope_rows);
    const int nq = d.Hq * d.hs, nk = d.Hkv * d.hs;   // q / k columns (128-wide rotary slots, or native 96)
    const int nv = d.Hkv * d.hv;                      // V columns of the QKV GEMM (packed heads)
    const int no = d.Hq * d.ho;                       // attention output columns = K of `wo`
    __nv_bfloat16* hcur = static_cast<__nv_bfloat16*>(embeds);
    MJV_ALLOC(xn, __nv_bfloat16*, ar, static_cast<size_t>(M) * d.D * 2);
    MJV_ALLOC(q, __nv_bfloat16*, ar, static_cast<size_t>(M) * nq * 2);
    MJV_ALLOC(k, __nv_bfloat16*, ar, static_cast<size_t>(M) * nk * 2);
    const size_t vt_bytes = static_cast<size_t>(d.Hkv) * d.hv * b.vt_cols * 2;
    MJV_ALLOC(vt, __nv_bfloat16*, ar, vt_bytes);
    MJV_ALLOC(att, __nv_bfloat16*, ar, static_cast<size_t>(M) * no * 2);
    MJV_ALLOC(mid, __nv_bfloat16*, ar, static_cast<size_t>(M) * d.F * 2);
    // alignment-pad columns of V^T are never written: keep them finite (they are multiplied by exact zeros)
    MJV_CUDA_OK(cudaMemsetAsync(vt, 0, vt_bytes, st));

    // compact buffers of the last layer: only the 2 rows per sample the head reads go through wo / MLP / final norm
    const int G2 = 2 * b.n_rows;
    const bool shortcut = b.work_last != nullptr && b.n_work_last > 0;
    MJV_ALLOC(att_c, __nv_bfloat16*, ar, static_cast<size_t>(G2) * no * 2);
    MJV_ALLOC(h_c, __nv_bfloat16*, ar, static_cast<size_t>(G2) * d.D * 2);
    MJV_ALLOC(xn_c, __nv_bfloat16*, ar, static_cast<size_t>(G2) * d.D * 2);
    MJV_ALLOC(mid_c, __nv_bfloat16*, ar, static_cast<size_t>(G2) * d.F * 2);

    for (int i = 0; i < c.llm_layers; ++i) {
        const LlmLayerW& w = h->llm[i];
        const bool last = shortcut && (i == c.llm_layers - 1);
        MJV_PROPAGATE(rmsnorm(hcur, w.attn_norm, xn, M, d.D, c.rms_eps, nullptr, st));
        {
            mjv_gemm_desc g = gemm_base(xn, d.D, w.wqkv, d.D, q, nq, (int)M, nq + nk + nv, d.D, MJV_EPI_QKV_LLM);
            g.v_head_dim = d.hv;
            g.qk_head_slot = d.hs;
            g.out_k = k;
            g.out_vt = vt;
            g.rope_cos = h->rope_cos;
            g.rope_sin = h->rope_sin;
            g.row_pos = b.row_pos;
            g.row_vt = b.row_vt;
            g.n_q_cols = nq;
            g.n_k_cols = nk;
            g.vt_stride_head = static_cast<long long>(d.hv) * b.vt_cols;
            g.vt_ld = static_cast<int>(b.vt_cols);
            MJV_PROPAGATE(gemm_bf16(g, st));
        }
        {
            mjv_attn_desc a;
            memset(&a, 0, sizeof(a));
            a.q = q;
            a.k = k;
            a.vt = vt;
            a.out = att;
            a.ldq = nq;
            a.ldk = nk;
            a.ldo = no;
            a.out_head_dim = d.ho;
            a.v_head_dim = d.hv;
            a.vt_ld = b.vt_cols;
            a.vt_plane_stride = static_cast<long long>(d.hv) * b.vt_cols;
            a.total_rows = M;
            a.head_dim = kHeadSlot;          // the kernel's tile width
            a.qk_head_stride = d.hs;         // 128, or 96: heads packed, 6 of 8 k-steps
            a.heads = d.Hq;
            a.kv_heads = d.Hkv;
            a.causal = 1;
            a.nseq = b.n_rows;
            a.cu_seqlens = b.cu_seqlens;
            // the last layer only needs the query blocks that hold a gathered row (its K/V still cover every token)
            a.work = last ? b.work_last : b.work;
            a.n_work = last ? b.n_work_last : b.n_work;
            a.vt_starts = b.vt_starts;
            a.vt_cols = b.vt_cols;
            a.flops_hint = last ? 0.0 : b.attn_flops;
            a.scale = 1.0f / sqrtf(static_cast<float>(d.hd));
            MJV_PROPAGATE(attention(a, st));
        }
        __nv_bfloat16* hres = hcur;   // residual stream of the rows processed from here on
        const __nv_bfloat16* att_in = att;
        __nv_bfloat16* xn_in = xn;
        __nv_bfloat16* mid_in = mid;
        int rows = (int)M;
        if (last) {
            MJV_PROPAGATE(gather_rows(att, b.gather_rows, att_c, G2, no, st));
            MJV_PROPAGATE(gather_rows(hcur, b.gather_rows, h_c, G2, d.D, st));
            hres = h_c;
            att_in = att_c;
            xn_in = xn_c;
            mid_in = mid_c;
            rows = G2;
        }
        {
            mjv_gemm_desc g = gemm_base(att_in, no, w.wo, no, hres, d.D, rows, d.D, no, MJV_EPI_RESID);
            g.residual = hres;
            g.ldr = d.D;
            MJV_PROPAGATE(gemm_bf16(g, st));
        }
        MJV_PROPAGATE(rmsnorm(hres, w.ffn_norm, xn_in, rows, d.D, c.rms_eps, nullptr, st));
        {
            mjv_gemm_desc g = gemm_base(xn_in, d.D, w.w13, d.D, mid_in, d.F, rows, 2 * d.F, d.D, MJV_EPI_SWIGLU);
            MJV_PROPAGATE(gemm_bf16(g, st));
        }
        {
            mjv_gemm_desc g = gemm_base(mid_in, d.F, w.w2, d.F, hres, d.D, rows, d.D, d.F, MJV_EPI_RESID);
            g.residual = hres;
            g.ldr = d.D;
            MJV_PROPAGATE(gemm_bf16(g, st));
        }
    }
    // final norm only on the 2 rows per sample the head reads (hidden_states[-1][b, idx]: moe_reward.py:211,236-243)
    if (shortcut) {
        MJV_PROPAGATE(rmsnorm(h_c, h->final_norm, gathered, G2, d.D, c.rms_eps, nullptr, st));
    } else {
        MJV_PROPAGATE(rmsnorm(hcur, h->final_norm, gathered, G2, d.D, c.rms_eps, b.gather_rows, st));
    }
    return 0;
}

int reward_head(mjv_handle* h, const void* gathered, int B, const mjv_outputs& out, Arena& ar, cudaStream_t st) {
    const mjv_config& c = h->cfg;
    const int D = c.llm_hidden, GH = c.gating_hidden, NO = c.num_objectives, NA = c.num_aspects;
    MJV_CHECK(out.scores != nullptr && out.extras != nullptr, MJV_ERR_ARG, "reward_head: scores/extras are required");
    const __nv_bfloat16* x_last = static_cast<const __nv_bfloat16*>(gathered);
    const __nv_bfloat16* x_gate = x_last + static_cast<long long>(B) * D;
    __nv_bfloat16* abuf[2];
    __nv_bfloat16* cbuf[2];
    for (int i = 0; i < 2; ++i) {
        MJV_ALLOC(a_, __nv_bfloat16*, ar, static_cast<size_t>(B) * GH * 2);
        MJV_ALLOC(c_, __nv_bfloat16*, ar, static_cast<size_t>(B) * GH * 2);
        abuf[i] = a_;
        cbuf[i] = c_;
    }
    MJV_ALLOC(raw, __nv_bfloat16*, ar, static_cast<size_t>(B) * 32 * 2);
    MJV_ALLOC(alog, __nv_bfloat16*, ar, static_cast<size_t>(B) * 32 * 2);
    MJV_ALLOC(clog, __nv_bfloat16*, ar, static_cast<size_t>(B) * 32 * 2);

    const int nh = c.gating_n_hidden;
    const __nv_bfloat16* ain = x_gate;
    const __nv_bfloat16* cin = x_gate;
    int in_dim = D;
    for (int l = 0; l <= nh; ++l) {
        const bool last = (l == nh);
        HeadJobs jobs;
        memset(&jobs, 0, sizeof(jobs));
        HeadJob& ja = jobs.j[0];
        ja.x = ain; ja.w = h->aspect.w[l]; ja.bias = h->aspect.b[l];
        ja.y = last ? alog : abuf[l & 1];
        ja.N = last ? NA : GH; ja.K = in_dim; ja.ldx = in_dim; ja.ldy = last ? 32 : GH; ja.relu = last ? 0 : 1;
        HeadJob& jc = jobs.j[1];
        jc.x = cin; jc.w = h->criteria.w[l]; jc.bias = h->criteria.b[l];
        jc.y = last ? clog : cbuf[l & 1];
        jc.N = last ? NO : GH; jc.K = in_dim; jc.ldx = in_dim; jc.ldy = last ? 32 : GH; jc.relu = last ? 0 : 1;
        jobs.n = 2;
        if (l == 0) {  // regression_layer (no bias) rides along with the first gating layer
            HeadJob& jr = jobs.j[2];
            jr.x = x_last; jr.w = h->regression; jr.bias = nullptr; jr.y = raw;
            jr.N = NO; jr.K = D; jr.ldx = D; jr.ldy = 32; jr.relu = 0;
            jobs.n = 3;
        }
        MJV_PROPAGATE(skinny_linear(jobs, B, st));
        ain = abuf[l & 1];
        cin = cbuf[l & 1];
        in_dim = GH;
    }
    HeadFinalizeArgs f;
    memset(&f, 0, sizeof(f));
    f.B = B;
    f.num_objectives = NO;
    f.num_aspects = NA;
    f.temperature = c.gating_temperature;
    f.raw_rewards = raw; f.ld_raw = 32;
    f.transform = static_cast<const __nv_bfloat16*>(h->transform);
    f.aspect_logits = alog; f.ld_al = 32;
    f.criteria_logits = clog; f.ld_cl = 32;
    f.aspect_logit_scale = static_cast<const __nv_bfloat16*>(h->aspect.logit_scale);
    f.criteria_logit_scale = static_cast<const __nv_bfloat16*>(h->criteria.logit_scale);
    f.group_offsets = h->group_offsets_dev;
    f.group_items = h->group_items_dev;
    f.scores = out.scores; f.ld_scores = NO + NA + 1;
    f.extras = out.extras; f.ld_extras = NO + NA + NO + 1;
    MJV_PROPAGATE(head_finalize(f, st));
    if (out.hidden_state)
        MJV_CUDA_OK(cudaMemcpyAsync(out.hidden_state, x_last, static_cast<size_t>(B) * D * 2, cudaMemcpyDeviceToDevice, st));
    if (out.prompt_embedding)
        MJV_CUDA_OK(cudaMemcpyAsync(out.prompt_embedding, x_gate, static_cast<size_t>(B) * D * 2, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// GatingNetwork as a standalone operator (moe_reward.py:16-42): n_hidden x (Linear + ReLU), Linear, then either the raw
// logits (forward_wo_softmax) or softmax(x / T) * logit_scale (forward). x: [B, llm_hidden] bf16.
int gating_forward(mjv_handle* h, int which, const void* x, int B, int with_softmax, void* out, Arena& ar, cudaStream_t st) {
    const mjv_config& c = h->cfg;
    MJV_CHECK(which == 0 || which == 1, MJV_ERR_ARG, "gating_forward: network %d (0 = aspect, 1 = criteria)", which);
    const GateW& g = which == 0 ? h->aspect : h->criteria;
    const int n_out = which == 0 ? c.num_aspects : c.num_objectives;
    const int GH = c.gating_hidden, nh = c.gating_n_hidden;
    __nv_bfloat16* buf[2];
    for (int i = 0; i < 2; ++i) {
        MJV_ALLOC(b_, __nv_bfloat16*, ar, static_cast<size_t>(B) * GH * 2);
        buf[i] = b_;
    }
    MJV_ALLOC(logits, __nv_bfloat16*, ar, static_cast<size_t>(B) * 32 * 2);
    const void* in = x;
    int in_dim = c.llm_hidden;
    for (int l = 0; l <= nh; ++l) {
        const bool last = (l == nh);
        HeadJobs jobs;
        memset(&jobs, 0, sizeof(jobs));
        HeadJob& j = jobs.j[0];
        j.x = in; j.w = g.w[l]; j.bias = g.b[l];
        j.y = last ? (with_softmax ? static_cast<void*>(logits) : out) : static_cast<void*>(buf[l & 1]);
        j.N = last ? n_out : GH; j.K = in_dim; j.ldx = in_dim;
        j.ldy = last ? (with_softmax ? 32 : n_out) : GH;
        j.relu = last ? 0 : 1;
        jobs.n = 1;
        MJV_PROPAGATE(skinny_linear(jobs, B, st));
        in = buf[l & 1];
        in_dim = GH;
    }
    if (with_softmax)
        MJV_PROPAGATE(gating_softmax(logits, 32, n_out, c.gating_temperature, g.logit_scale, out, n_out, B, st));
    return 0;
}

int check_ready(const mjv_handle* h) {
    MJV_CHECK(h != nullptr, MJV_ERR_ARG, "null handle");
    MJV_CHECK(h->finalized, MJV_ERR_STATE, "weights not finalized: call mjv_finalize_weights first");
    return 0;
}

}  // namespace
}  // namespace mjv

// ================================================================================================ C ABI
using namespace mjv;

extern "C" int mjv_create(const mjv_config* cfg, mjv_handle** out) {
    MJV_CHECK(cfg && out, MJV_ERR_ARG, "mjv_create: null argument");
    MJV_CHECK(cfg->image_size % cfg->patch_size == 0, MJV_ERR_ARG, "image_size must be a multiple of patch_size");
    MJV_CHECK(cfg->vit_hidden % 64 == 0 && cfg->llm_hidden % 64 == 0, MJV_ERR_ARG, "hidden sizes must be multiples of 64");
    MJV_CHECK(cfg->num_objectives >= 1 && cfg->num_objectives <= 32 && cfg->num_aspects >= 1 && cfg->num_aspects <= 32,
              MJV_ERR_ARG, "head supports 1..32 criteria and aspects");
    MJV_CHECK(cfg->gating_n_hidden >= 1, MJV_ERR_ARG, "gating_n_hidden must be >= 1");
    int dev_count = 0;
    MJV_CHECK(cudaGetDeviceCount(&dev_count) == cudaSuccess && dev_count > 0, MJV_ERR_CUDA,
              "no CUDA device: mj_video_b200 has no CPU fallback");
    mjv_handle* h = new mjv_handle();
    h->cfg = *cfg;
    h->kpad = static_cast<int>(round_up(3LL * cfg->patch_size * cfg->patch_size, 64));
    cudaError_t e = cudaMalloc(&h->group_offsets_dev, 33 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&h->group_items_dev, 32 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(h->group_offsets_dev, cfg->group_offsets, 33 * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->group_items_dev, cfg->group_items, 32 * sizeof(int), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        delete h;
        return set_error(MJV_ERR_CUDA, "mjv_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return 0;
}

extern "C" void mjv_destroy(mjv_handle* h) {
    if (!h) return;
    for (auto& kv : h->weights) cudaFree(kv.second.ptr);
    for (auto& kv : h->weights32) cudaFree(kv.second.ptr);
    cudaFree(h->group_offsets_dev);
    cudaFree(h->group_items_dev);
    delete h;
}

extern "C" int mjv_load_weight(mjv_handle* h, const char* key, const void* data, const int64_t* shape, int32_t ndim,
                               int32_t dtype) {
    return load_weight(h, key, data, shape, ndim, dtype);
}

extern "C" int mjv_finalize_weights(mjv_handle* h) {
    MJV_CHECK(h != nullptr, MJV_ERR_ARG, "null handle");
    return finalize(h);
}

extern "C" int64_t mjv_workspace_bytes(const mjv_handle* h, int32_t n_rows, int32_t n_tiles, int64_t n_tokens) {
    if (!h) return -1;
    if (h->cfg.precision == 1) return static_cast<int64_t>(workspace_bytes_f32(h, n_rows, n_tiles, n_tokens));
    const Dims d = dims_of(h);
    const long long vt_cols = n_tokens + 8LL * n_rows;
    size_t persistent = al(static_cast<size_t>(n_tiles) * d.S * d.C * 2) + al(static_cast<size_t>(n_tokens) * d.D * 2) +
                        al(2ULL * n_rows * d.D * 2) + 8192;
    size_t temp = vit_temp_bytes(d, n_tiles);
    size_t t2 = proj_temp_bytes(d, n_tiles);
    if (t2 > temp) temp = t2;
    t2 = llm_temp_bytes(d, n_tokens, vt_cols, n_rows);
    if (t2 > temp) temp = t2;
    t2 = head_temp_bytes(h->cfg, n_rows);
    if (t2 > temp) temp = t2;
    return static_cast<int64_t>(persistent + temp + 65536);
}

extern "C" int mjv_gating_forward(mjv_handle* h, int32_t which, const void* x, int32_t n_rows, int32_t with_softmax,
                                  void* out, void* workspace, int64_t workspace_bytes, void* stream) {
    MJV_PROPAGATE(check_ready(h));
    MJV_CHECK(x != nullptr && out != nullptr && n_rows >= 0, MJV_ERR_ARG, "mjv_gating_forward: bad arguments");
    Arena ar(workspace, static_cast<size_t>(workspace_bytes));
    if (h->cfg.precision == 1) {  // fp32 in, fp32 out
        const int n_out = which == 0 ? h->cfg.num_aspects : h->cfg.num_objectives;
        const char* name = which == 0 ? "aspect_gating" : "criteria_gating";
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (!with_softmax) return gating_logits_f32(h, name, static_cast<const float*>(x), n_rows, n_out, static_cast<float*>(out), ar, st);
        MJV_ALLOC(logits, float*, ar, static_cast<size_t>(n_rows) * 64 * 4 + 16);
        MJV_PROPAGATE(gating_logits_f32(h, name, static_cast<const float*>(x), n_rows, n_out, logits, ar, st));
        const float* ls;
        MJV_PROPAGATE(need32(h, std::string(name) + ".logit_scale", &ls));
        return f32::gating_softmax(logits, n_out, h->cfg.gating_temperature, ls, static_cast<float*>(out), n_rows, st);
    }
    return gating_forward(h, which, x, n_rows, with_softmax, out, ar, static_cast<cudaStream_t>(stream));
}

extern "C" int mjv_vit_forward(mjv_handle* h, const void* pixel_values, int32_t pixel_dtype, int32_t n_tiles,
                               void* vit_hidden, void* workspace, int64_t workspace_bytes, void* stream) {
    MJV_PROPAGATE(check_ready(h));
    MJV_CHECK(h->cfg.precision == 0, MJV_ERR_STATE, "stage entry points are bf16-only: a handle in fp32 verification mode runs through mjv_score_forward");
    if (n_tiles == 0) return 0;
    Arena ar(workspace, static_cast<size_t>(workspace_bytes));
    return vit_forward(h, pixel_values, pixel_dtype, n_tiles, vit_hidden, ar, static_cast<cudaStream_t>(stream));
}

extern "C" int mjv_project_and_splice(mjv_handle* h, const void* vit_hidden, const mjv_batch* batch, void* embeds,
                                      void* workspace, int64_t workspace_bytes, void* stream) {
    MJV_PROPAGATE(check_ready(h));
    MJV_CHECK(h->cfg.precision == 0, MJV_ERR_STATE, "stage entry points are bf16-only: a handle in fp32 verification mode runs through mjv_score_forward");
    MJV_CHECK(batch != nullptr, MJV_ERR_ARG, "null batch");
    Arena ar(workspace, static_cast<size_t>(workspace_bytes));
    return project_and_splice(h, vit_hidden, *batch, embeds, ar, static_cast<cudaStream_t>(stream));
}

extern "C" int mjv_llm_prefill(mjv_handle* h, void* embeds_inout, const mjv_batch* batch, void* gathered_2b,
                               void* workspace, int64_t workspace_bytes, void* stream) {
    MJV_PROPAGATE(check_ready(h));
    MJV_CHECK(h->cfg.precision == 0, MJV_ERR_STATE, "stage entry points are bf16-only: a handle in fp32 verification mode runs through mjv_score_forward");
    MJV_CHECK(batch != nullptr, MJV_ERR_ARG, "null batch");
    Arena ar(workspace, static_cast<size_t>(workspace_bytes));
    return llm_prefill(h, embeds_inout, *batch, gathered_2b, ar, static_cast<cudaStream_t>(stream));
}

extern "C" int mjv_reward_head(mjv_handle* h, const void* gathered_2b, int32_t n_rows, const mjv_outputs* out,
                               void* workspace, int64_t workspace_bytes, void* stream) {
    MJV_PROPAGATE(check_ready(h));
    MJV_CHECK(h->cfg.precision == 0, MJV_ERR_STATE, "stage entry points are bf16-only: a handle in fp32 verification mode runs through mjv_score_forward");
    MJV_CHECK(out != nullptr, MJV_ERR_ARG, "null outputs");
    Arena ar(workspace, static_cast<size_t>(workspace_bytes));
    return reward_head(h, gathered_2b, n_rows, *out, ar, static_cast<cudaStream_t>(stream));
}

extern "C" int mjv_score_forward(mjv_handle* h, const mjv_batch* batch, const mjv_outputs* out, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
    MJV_PROPAGATE(check_ready(h));
    MJV_CHECK(batch && out && workspace, MJV_ERR_ARG, "mjv_score_forward: null argument");
    if (batch->n_rows == 0) return 0;
    const Dims d = dims_of(h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Arena ar(workspace, static_cast<size_t>(workspace_bytes));
    if (h->cfg.precision == 1) return score_forward_f32(h, *batch, *out, ar, st);
    MJV_ALLOC(vit_hidden, __nv_bfloat16*, ar, static_cast<size_t>(batch->n_tiles) * d.S * d.C * 2 + 16);
    MJV_ALLOC(embeds, __nv_bfloat16*, ar, static_cast<size_t>(batch->n_tokens) * d.D * 2);
    MJV_ALLOC(gathered, __nv_bfloat16*, ar, 2ULL * batch->n_rows * d.D * 2);
    const size_t mark = ar.off;
    if (batch->n_tiles > 0) {
        MJV_PROPAGATE(vit_forward(h, batch->pixel_values, batch->pixel_dtype, batch->n_tiles, vit_hidden, ar, st));
    }
    ar.off = mark;
    MJV_PROPAGATE(project_and_splice(h, vit_hidden, *batch, embeds, ar, st));
    ar.off = mark;
    MJV_PROPAGATE(llm_prefill(h, embeds, *batch, gathered, ar, st));
    ar.off = mark;
    MJV_PROPAGATE(reward_head(h, gathered, batch->n_rows, *out, ar, st));
    return 0;
}
