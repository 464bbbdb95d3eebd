// fp32 verification mode of the forward (mjv_config.precision == 1): the SAME path as engine.cu - ViT -> projector / splice
// -> LLM prefill -> MoE head - with every activation and weight in fp32 and one plain kernel per reference op
// (fp32_kernels.cuh), in the reference's op order:
//   InternVLChatRewardModeling.forward   scripts/model/moe_reward.py:183-297 run with a float32 module (SURVEY §A.7:
//   the fp32 CPU run is the numerical oracle; gate 1e-4 abs on rewards / aspect_scores / score)
// Weights are kept in the reference's own layouts (no re-packing), keyed by the canonical state-dict names.
// Included by engine.cu after mjv_handle; not a separate translation unit.
#pragma once
#include "fp32_kernels.cuh"

namespace mjv {
namespace {

int load_weight_f32(mjv_handle* h, const std::string& key, const void* data, long long rows, long long cols, int dtype) {
    const long long n = rows * cols;
    auto it = h->weights32.find(key);
    if (it != h->weights32.end() && it->second.rows * it->second.cols != n) {
        MJV_CUDA_OK(cudaDeviceSynchronize());
        cudaFree(it->second.ptr);
        h->weights32.erase(it);
        it = h->weights32.end();
    }
    if (it == h->weights32.end()) {
        Tensor t;
        t.rows = rows;
        t.cols = cols;
        t.ld = cols;
        t.is_float = true;
        MJV_CUDA_OK(cudaMalloc(&t.ptr, static_cast<size_t>(n) * 4));
        it = h->weights32.emplace(key, t).first;
    }
    if (dtype == MJV_DTYPE_F32) {
        MJV_CUDA_OK(cudaMemcpy(it->second.ptr, data, static_cast<size_t>(n) * 4, cudaMemcpyDefault));
    } else {
        void* staging = nullptr;
        MJV_CUDA_OK(cudaMalloc(&staging, static_cast<size_t>(n) * 2));
        cudaError_t e = cudaMemcpy(staging, data, static_cast<size_t>(n) * 2, cudaMemcpyDefault);
        int rc = 0;
        if (e == cudaSuccess) rc = f32::to_f32(staging, dtype, static_cast<float*>(it->second.ptr), n, 0);
        if (e == cudaSuccess && rc == 0) e = cudaDeviceSynchronize();
        cudaFree(staging);
        if (rc != 0) return rc;
        if (e != cudaSuccess) return set_error(MJV_ERR_CUDA, "mjv_load_weight(%s): %s", key.c_str(), cudaGetErrorString(e));
    }
    it->second.filled_rows = rows;
    return 0;
}

int need32(const mjv_handle* h, const std::string& key, const float** out, long long elems = -1) {
    auto it = h->weights32.find(key);
    if (it == h->weights32.end()) return set_error(MJV_ERR_STATE, "missing weight: %s", key.c_str());
    if (elems >= 0 && it->second.rows * it->second.cols != elems)
        return set_error(MJV_ERR_STATE, "weight %s has %lld elements, expected %lld", key.c_str(),
                         it->second.rows * it->second.cols, elems);
    *out = static_cast<const float*>(it->second.ptr);
    return 0;
}

int finalize_f32(mjv_handle* h) {
    const mjv_config& c = h->cfg;
    const float* p;
    const std::string v = "model.vision_model.";
    MJV_PROPAGATE(need32(h, v + "embeddings.class_embedding", &p, c.vit_hidden));
    MJV_PROPAGATE(need32(h, v + "embeddings.patch_embedding.weight", &p, 3LL * c.patch_size * c.patch_size * c.vit_hidden));
    for (int i = 0; i < c.vit_layers; ++i) {
        const std::string q = v + "encoder.layers." + std::to_string(i) + ".";
        for (const char* s : {"ls1", "ls2", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
                              "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias", "norm1.weight", "norm1.bias",
                              "norm2.weight", "norm2.bias"})
            MJV_PROPAGATE(need32(h, q + s, &p));
    }
    for (const char* s : {"model.mlp1.0.weight", "model.mlp1.0.bias", "model.mlp1.1.weight", "model.mlp1.1.bias",
                          "model.mlp1.3.weight", "model.mlp1.3.bias", "model.language_model.model.tok_embeddings.weight",
                          "model.language_model.model.norm.weight", "regression_layer.weight", "reward_transform_matrix",
                          "aspect_gating.logit_scale", "criteria_gating.logit_scale", "rope.cos", "rope.sin"})
        MJV_PROPAGATE(need32(h, s, &p));
    for (int i = 0; i < c.llm_layers; ++i) {
        const std::string q = "model.language_model.model.layers." + std::to_string(i) + ".";
        if (c.llm_arch == 0) {
            MJV_PROPAGATE(need32(h, q + "attention.wqkv.weight", &p));
        } else {
            for (const char* s : {"attention.wq.weight", "attention.wk.weight", "attention.wv.weight"}) MJV_PROPAGATE(need32(h, q + s, &p));
        }
        for (const char* s : {"attention.wo.weight", "feed_forward.w1.weight", "feed_forward.w3.weight", "feed_forward.w2.weight",
                              "attention_norm.weight", "ffn_norm.weight"})
            MJV_PROPAGATE(need32(h, q + s, &p));
    }
    for (const char* g : {"aspect_gating", "criteria_gating"})
        for (int i = 0; i <= c.gating_n_hidden; ++i) {
            MJV_PROPAGATE(need32(h, std::string(g) + ".layers." + std::to_string(i) + ".weight", &p));
            MJV_PROPAGATE(need32(h, std::string(g) + ".layers." + std::to_string(i) + ".bias", &p));
        }
    auto rc = h->weights32.find("rope.cos");
    h->rope_rows = rc->second.rows;
    h->finalized = true;
    return 0;
}

size_t workspace_bytes_f32(const mjv_handle* h, long long B, long long T, long long M) {
    const Dims d = dims_of(h);
    const mjv_config& c = h->cfg;
    const size_t Mv = static_cast<size_t>(T) * d.S, Mp = static_cast<size_t>(T) * d.tok_per_tile, m = static_cast<size_t>(M);
    const size_t kp = 3ULL * c.patch_size * c.patch_size;
    const size_t persistent = al(Mv * d.C * 4) + al(m * d.D * 4) + al(2ULL * B * d.D * 4);
    const size_t vit = al(static_cast<size_t>(T) * d.P * kp * 4) + al(static_cast<size_t>(T) * d.P * d.C * 4) + al(Mv * d.C * 4) +
                       al(Mv * 3 * d.C * 4) + al(Mv * d.C * 4) + al(Mv * d.Fv * 4);
    const size_t proj = 2 * al(Mp * 4 * d.C * 4) + 2 * al(Mp * d.D * 4);
    const size_t qkv = static_cast<size_t>(d.Hq + 2 * d.Hkv) * d.hd;
    const size_t llm = al(m * d.D * 4) + 3 * al(m * qkv * 4) + al(m * d.Hq * d.hd * 4) + 3 * al(m * d.F * 4);
    const size_t head = 4 * al(static_cast<size_t>(B) * c.gating_hidden * 4) + 4 * al(static_cast<size_t>(B) * 64 * 4);
    size_t temp = vit;
    if (proj > temp) temp = proj;
    if (llm > temp) temp = llm;
    if (head > temp) temp = head;
    return persistent + temp + 65536;
}

// GatingNetwork body (moe_reward.py:27-42): n_hidden x (Linear + ReLU), then Linear -> logits [B, n_out]
int gating_logits_f32(mjv_handle* h, const char* name, const float* x, int B, int n_out, float* logits, Arena& ar, cudaStream_t st) {
    const mjv_config& c = h->cfg;
    MJV_ALLOC(t0, float*, ar, static_cast<size_t>(B) * c.gating_hidden * 4);
    MJV_ALLOC(t1, float*, ar, static_cast<size_t>(B) * c.gating_hidden * 4);
    const float* cur = x;
    int k = c.llm_hidden;
    for (int i = 0; i <= c.gating_n_hidden; ++i) {
        const float *w, *b;
        MJV_PROPAGATE(need32(h, std::string(name) + ".layers." + std::to_string(i) + ".weight", &w));
        MJV_PROPAGATE(need32(h, std::string(name) + ".layers." + std::to_string(i) + ".bias", &b));
        const bool last = i == c.gating_n_hidden;
        float* dst = last ? logits : (i % 2 == 0 ? t0 : t1);
        const int n = last ? n_out : c.gating_hidden;
        MJV_PROPAGATE(f32::sgemm_tn(cur, k, w, k, dst, n, B, n, k, b, last ? f32::EPI_NONE : f32::EPI_RELU, nullptr, 0, nullptr, st));
        cur = dst;
        k = n;
    }
    return 0;
}

int score_forward_f32(mjv_handle* h, const mjv_batch& b, const mjv_outputs& out, Arena& ar, cudaStream_t st) {
    const mjv_config& c = h->cfg;
    const Dims d = dims_of(h);
    const long long T = b.n_tiles, M = b.n_tokens;
    const int B = b.n_rows;
    const long long Mv = T * d.S;
    const int kp = 3 * c.patch_size * c.patch_size;
    const int g = c.image_size / c.patch_size;
    MJV_ALLOC(hid, float*, ar, static_cast<size_t>(Mv) * d.C * 4 + 16);
    MJV_ALLOC(embeds, float*, ar, static_cast<size_t>(M) * d.D * 4);
    MJV_ALLOC(gathered, float*, ar, 2ULL * B * d.D * 4);
    const size_t mark = ar.off;
    const std::string v = "model.vision_model.";
    const float *w0, *w1, *w2, *w3;

    // ------------------------------------------------------------------ InternViT (modeling_intern_vit.py:162-174, 283-295)
    if (T > 0) {
        MJV_ALLOC(patches, float*, ar, static_cast<size_t>(T) * d.P * kp * 4);
        MJV_ALLOC(conv, float*, ar, static_cast<size_t>(T) * d.P * d.C * 4);
        MJV_ALLOC(ln, float*, ar, static_cast<size_t>(Mv) * d.C * 4);
        MJV_ALLOC(qkv, float*, ar, static_cast<size_t>(Mv) * 3 * d.C * 4);
        MJV_ALLOC(att, float*, ar, static_cast<size_t>(Mv) * d.C * 4);
        MJV_ALLOC(mid, float*, ar, static_cast<size_t>(Mv) * d.Fv * 4);
        MJV_PROPAGATE(f32::patchify(b.pixel_values, b.pixel_dtype, patches, static_cast<int>(T), c.image_size, c.patch_size, st));
        MJV_PROPAGATE(need32(h, v + "embeddings.patch_embedding.weight", &w0));
        MJV_PROPAGATE(need32(h, v + "embeddings.patch_embedding.bias", &w1));
        MJV_PROPAGATE(f32::sgemm_tn(patches, kp, w0, kp, conv, d.C, T * d.P, d.C, kp, w1, f32::EPI_NONE, nullptr, 0, nullptr, st));
        MJV_PROPAGATE(need32(h, v + "embeddings.class_embedding", &w0));
        MJV_PROPAGATE(need32(h, v + "embeddings.position_embedding", &w1, static_cast<long long>(d.S) * d.C));
        MJV_PROPAGATE(f32::vit_assemble(conv, w0, w1, hid, static_cast<int>(T), d.P, d.C, st));
        const int vhd = d.C / d.heads;
        for (int i = 0; i < c.vit_layers; ++i) {
            const std::string p = v + "encoder.layers." + std::to_string(i) + ".";
            MJV_PROPAGATE(need32(h, p + "norm1.weight", &w0));
            MJV_PROPAGATE(need32(h, p + "norm1.bias", &w1));
            MJV_PROPAGATE(f32::layernorm(hid, w0, w1, ln, Mv, d.C, c.vit_eps, st));
            MJV_PROPAGATE(need32(h, p + "attn.qkv.weight", &w0));
            MJV_PROPAGATE(need32(h, p + "attn.qkv.bias", &w1));
            MJV_PROPAGATE(f32::sgemm_tn(ln, d.C, w0, d.C, qkv, 3 * d.C, Mv, 3 * d.C, d.C, w1, f32::EPI_NONE, nullptr, 0, nullptr, st));
            f32::AttnF32 a;
            memset(&a, 0, sizeof(a));
            a.q = a.k = a.v = qkv;
            a.ldq = a.ldk = a.ldv = 3 * d.C;
            a.q_col0 = 0;
            a.k_col0 = d.C;
            a.v_col0 = 2 * d.C;
            a.q_group_stride = a.kv_stride = vhd;
            a.heads = d.heads;
            a.group = 1;
            a.hd = vhd;
            a.scale = 1.0f / sqrtf(static_cast<float>(vhd));
            a.causal = 0;
            a.nseq = static_cast<int>(T);
            a.seq_len = d.S;
            a.out = att;
            a.ldo = d.C;
            MJV_PROPAGATE(f32::attention(a, st));
            MJV_PROPAGATE(need32(h, p + "attn.proj.weight", &w0));
            MJV_PROPAGATE(need32(h, p + "attn.proj.bias", &w1));
            MJV_PROPAGATE(need32(h, p + "ls1", &w2));
            MJV_PROPAGATE(f32::sgemm_tn(att, d.C, w0, d.C, hid, d.C, Mv, d.C, d.C, w1, f32::EPI_RESID_SCALE, hid, d.C, w2, st));
            MJV_PROPAGATE(need32(h, p + "norm2.weight", &w0));
            MJV_PROPAGATE(need32(h, p + "norm2.bias", &w1));
            MJV_PROPAGATE(f32::layernorm(hid, w0, w1, ln, Mv, d.C, c.vit_eps, st));
            MJV_PROPAGATE(need32(h, p + "mlp.fc1.weight", &w0));
            MJV_PROPAGATE(need32(h, p + "mlp.fc1.bias", &w1));
            MJV_PROPAGATE(f32::sgemm_tn(ln, d.C, w0, d.C, mid, d.Fv, Mv, d.Fv, d.C, w1, f32::EPI_GELU, nullptr, 0, nullptr, st));
            MJV_PROPAGATE(need32(h, p + "mlp.fc2.weight", &w0));
            MJV_PROPAGATE(need32(h, p + "mlp.fc2.bias", &w1));
            MJV_PROPAGATE(need32(h, p + "ls2", &w2));
            MJV_PROPAGATE(f32::sgemm_tn(mid, d.Fv, w0, d.Fv, hid, d.C, Mv, d.C, d.Fv, w1, f32::EPI_RESID_SCALE, hid, d.C, w2, st));
        }
    }
    ar.off = mark;

    // ------------------------------------------------------------------ projector + splice (modeling_internvl_chat.py:163-188, 228-262)
    {
        const long long Mp = T * d.tok_per_tile;
        MJV_ALLOC(ps, float*, ar, static_cast<size_t>(Mp) * 4 * d.C * 4 + 16);
        MJV_ALLOC(ln, float*, ar, static_cast<size_t>(Mp) * 4 * d.C * 4 + 16);
        MJV_ALLOC(m1, float*, ar, static_cast<size_t>(Mp) * d.D * 4 + 16);
        MJV_ALLOC(ve, float*, ar, static_cast<size_t>(Mp) * d.D * 4 + 16);
        if (T > 0) {
            MJV_PROPAGATE(f32::pixel_shuffle(hid, ps, static_cast<int>(T), g, d.C, c.ps_version == 1, st));
            MJV_PROPAGATE(need32(h, "model.mlp1.0.weight", &w0));
            MJV_PROPAGATE(need32(h, "model.mlp1.0.bias", &w1));
            MJV_PROPAGATE(f32::layernorm(ps, w0, w1, ln, Mp, 4 * d.C, c.mlp1_eps, st));
            MJV_PROPAGATE(need32(h, "model.mlp1.1.weight", &w0));
            MJV_PROPAGATE(need32(h, "model.mlp1.1.bias", &w1));
            MJV_PROPAGATE(f32::sgemm_tn(ln, 4 * d.C, w0, 4 * d.C, m1, d.D, Mp, d.D, 4 * d.C, w1, f32::EPI_GELU, nullptr, 0, nullptr, st));
            MJV_PROPAGATE(need32(h, "model.mlp1.3.weight", &w0));
            MJV_PROPAGATE(need32(h, "model.mlp1.3.bias", &w1));
            MJV_PROPAGATE(f32::sgemm_tn(m1, d.D, w0, d.D, ve, d.D, Mp, d.D, d.D, w1, f32::EPI_NONE, nullptr, 0, nullptr, st));
        }
        MJV_PROPAGATE(need32(h, "model.language_model.model.tok_embeddings.weight", &w0));
        MJV_PROPAGATE(f32::embed(b.token_ids, w0, embeds, M, d.D, b.img_context_token_id, st));
        MJV_CHECK(b.n_img_tokens <= Mp, MJV_ERR_ARG, "more <IMG_CONTEXT> tokens (%d) than visual embeddings (%lld)", b.n_img_tokens, Mp);
        MJV_PROPAGATE(f32::scatter_rows(ve, b.img_rows, embeds, b.n_img_tokens, d.D, st));
    }
    ar.off = mark;

    // ------------------------------------------------------------------ LLM prefill (modeling_internlm2.py:854-984 / Llama branch)
    {
        MJV_CHECK(b.max_seq_len <= h->rope_rows, MJV_ERR_ARG, "sequence of %d tokens but the RoPE tables hold %lld rows",
                  b.max_seq_len, h->rope_rows);
        const int hd = d.hd, H = d.Hq, KV = d.Hkv, grp = H / KV;
        const int nqkv = (H + 2 * KV) * hd;
        const float *cos_t, *sin_t;
        MJV_PROPAGATE(need32(h, "rope.cos", &cos_t));
        MJV_PROPAGATE(need32(h, "rope.sin", &sin_t));
        MJV_ALLOC(xn, float*, ar, static_cast<size_t>(M) * d.D * 4);
        MJV_ALLOC(qkv, float*, ar, static_cast<size_t>(M) * nqkv * 4);
        MJV_ALLOC(kbuf, float*, ar, static_cast<size_t>(M) * nqkv * 4);   // Llama branch: separate k / v projections
        MJV_ALLOC(vbuf, float*, ar, static_cast<size_t>(M) * nqkv * 4);
        MJV_ALLOC(att, float*, ar, static_cast<size_t>(M) * H * hd * 4);
        MJV_ALLOC(g1, float*, ar, static_cast<size_t>(M) * d.F * 4);
        MJV_ALLOC(up, float*, ar, static_cast<size_t>(M) * d.F * 4);
        MJV_ALLOC(mid, float*, ar, static_cast<size_t>(M) * d.F * 4);
        float* hcur = embeds;
        for (int i = 0; i < c.llm_layers; ++i) {
            const std::string p = "model.language_model.model.layers." + std::to_string(i) + ".";
            MJV_PROPAGATE(need32(h, p + "attention_norm.weight", &w0));
            MJV_PROPAGATE(f32::rmsnorm(hcur, w0, xn, M, d.D, c.rms_eps, nullptr, st));
            f32::AttnF32 a;
            memset(&a, 0, sizeof(a));
            if (c.llm_arch == 0) {
                // wqkv rows: per kv group [q x groups, k, v] x head_dim (modeling_internlm2.py:463-473)
                MJV_PROPAGATE(need32(h, p + "attention.wqkv.weight", &w0, static_cast<long long>(nqkv) * d.D));
                MJV_PROPAGATE(f32::sgemm_tn(xn, d.D, w0, d.D, qkv, nqkv, M, nqkv, d.D, nullptr, f32::EPI_NONE, nullptr, 0, nullptr, st));
                const int stride = (grp + 2) * hd;
                MJV_PROPAGATE(f32::rope(qkv, nqkv, M, H, grp, stride, 0, hd, cos_t, sin_t, 64, b.row_pos, st));
                MJV_PROPAGATE(f32::rope(qkv, nqkv, M, KV, 1, stride, grp * hd, hd, cos_t, sin_t, 64, b.row_pos, st));
                a.q = a.k = a.v = qkv;
                a.ldq = a.ldk = a.ldv = nqkv;
                a.q_col0 = 0;
                a.k_col0 = grp * hd;
                a.v_col0 = (grp + 1) * hd;
                a.q_group_stride = a.kv_stride = stride;
            } else {
                MJV_PROPAGATE(need32(h, p + "attention.wq.weight", &w0, static_cast<long long>(H) * hd * d.D));
                MJV_PROPAGATE(need32(h, p + "attention.wk.weight", &w1, static_cast<long long>(KV) * hd * d.D));
                MJV_PROPAGATE(need32(h, p + "attention.wv.weight", &w2, static_cast<long long>(KV) * hd * d.D));
                MJV_PROPAGATE(f32::sgemm_tn(xn, d.D, w0, d.D, qkv, H * hd, M, H * hd, d.D, nullptr, f32::EPI_NONE, nullptr, 0, nullptr, st));
                MJV_PROPAGATE(f32::sgemm_tn(xn, d.D, w1, d.D, kbuf, KV * hd, M, KV * hd, d.D, nullptr, f32::EPI_NONE, nullptr, 0, nullptr, st));
                MJV_PROPAGATE(f32::sgemm_tn(xn, d.D, w2, d.D, vbuf, KV * hd, M, KV * hd, d.D, nullptr, f32::EPI_NONE, nullptr, 0, nullptr, st));
                MJV_PROPAGATE(f32::rope(qkv, H * hd, M, H, 1, hd, 0, hd, cos_t, sin_t, 64, b.row_pos, st));
                MJV_PROPAGATE(f32::rope(kbuf, KV * hd, M, KV, 1, hd, 0, hd, cos_t, sin_t, 64, b.row_pos, st));
                a.q = qkv;
                a.k = kbuf;
                a.v = vbuf;
                a.ldq = H * hd;
                a.ldk = a.ldv = KV * hd;
                a.q_group_stride = grp * hd;
                a.kv_stride = hd;
            }
            a.heads = H;
            a.group = grp;
            a.hd = hd;
            a.scale = 1.0f / sqrtf(static_cast<float>(hd));
            a.causal = 1;
            a.cu_seqlens = b.cu_seqlens;
            a.nseq = B;
            a.seq_len = b.max_seq_len;
            a.out = att;
            a.ldo = H * hd;
            MJV_PROPAGATE(f32::attention(a, st));
            MJV_PROPAGATE(need32(h, p + "attention.wo.weight", &w0));
            MJV_PROPAGATE(f32::sgemm_tn(att, H * hd, w0, H * hd, hcur, d.D, M, d.D, H * hd, nullptr, f32::EPI_RESID, hcur, d.D, nullptr, st));
            MJV_PROPAGATE(need32(h, p + "ffn_norm.weight", &w0));
            MJV_PROPAGATE(f32::rmsnorm(hcur, w0, xn, M, d.D, c.rms_eps, nullptr, st));
            MJV_PROPAGATE(need32(h, p + "feed_forward.w1.weight", &w1));
            MJV_PROPAGATE(need32(h, p + "feed_forward.w3.weight", &w3));
            MJV_PROPAGATE(need32(h, p + "feed_forward.w2.weight", &w2));
            MJV_PROPAGATE(f32::sgemm_tn(xn, d.D, w1, d.D, g1, d.F, M, d.F, d.D, nullptr, f32::EPI_NONE, nullptr, 0, nullptr, st));
            MJV_PROPAGATE(f32::sgemm_tn(xn, d.D, w3, d.D, up, d.F, M, d.F, d.D, nullptr, f32::EPI_NONE, nullptr, 0, nullptr, st));
            MJV_PROPAGATE(f32::silu_mul(g1, up, mid, M * d.F, st));
            MJV_PROPAGATE(f32::sgemm_tn(mid, d.F, w2, d.F, hcur, d.D, M, d.D, d.F, nullptr, f32::EPI_RESID, hcur, d.D, nullptr, st));
        }
        // final norm on the 2 rows per sample the head reads (row-wise, hence exact): last token, gating token
        MJV_PROPAGATE(need32(h, "model.language_model.model.norm.weight", &w0));
        MJV_PROPAGATE(f32::rmsnorm(hcur, w0, gathered, 2LL * B, d.D, c.rms_eps, b.gather_rows, st));
    }
    ar.off = mark;

    // ------------------------------------------------------------------ MoE reward head (moe_reward.py:236-297)
    {
        const int NO = c.num_objectives, NA = c.num_aspects;
        const float* h_last = gathered;
        const float* h_gate = gathered + static_cast<size_t>(B) * d.D;
        MJV_ALLOC(raw, float*, ar, static_cast<size_t>(B) * 64 * 4);
        MJV_ALLOC(al_, float*, ar, static_cast<size_t>(B) * 64 * 4);
        MJV_ALLOC(cl_, float*, ar, static_cast<size_t>(B) * 64 * 4);
        MJV_PROPAGATE(need32(h, "regression_layer.weight", &w0, static_cast<long long>(NO) * d.D));
        MJV_PROPAGATE(f32::sgemm_tn(h_last, d.D, w0, d.D, raw, NO, B, NO, d.D, nullptr, f32::EPI_NONE, nullptr, 0, nullptr, st));
        MJV_PROPAGATE(gating_logits_f32(h, "aspect_gating", h_gate, B, NA, al_, ar, st));
        MJV_PROPAGATE(gating_logits_f32(h, "criteria_gating", h_gate, B, NO, cl_, ar, st));
        f32::HeadF32 a;
        memset(&a, 0, sizeof(a));
        a.rewards_raw = raw;
        MJV_PROPAGATE(need32(h, "reward_transform_matrix", &a.transform, static_cast<long long>(NO) * NO));
        a.aspect_logits = al_;
        a.criteria_logits = cl_;
        MJV_PROPAGATE(need32(h, "aspect_gating.logit_scale", &a.aspect_logit_scale));
        MJV_PROPAGATE(need32(h, "criteria_gating.logit_scale", &a.criteria_logit_scale));
        a.group_offsets = h->group_offsets_dev;
        a.group_items = h->group_items_dev;
        a.temperature = c.gating_temperature;
        a.NO = NO;
        a.NA = NA;
        a.B = B;
        a.scores = out.scores;
        a.extras = out.extras;
        MJV_PROPAGATE(f32::head_finalize(a, st));
        // hidden_state / prompt_embedding are fp32 [B, D] in this mode
        if (out.hidden_state != nullptr)
            MJV_CUDA_OK(cudaMemcpyAsync(out.hidden_state, h_last, static_cast<size_t>(B) * d.D * 4, cudaMemcpyDeviceToDevice, st));
        if (out.prompt_embedding != nullptr)
            MJV_CUDA_OK(cudaMemcpyAsync(out.prompt_embedding, h_gate, static_cast<size_t>(B) * d.D * 4, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

}  // namespace
}  // namespace mjv
