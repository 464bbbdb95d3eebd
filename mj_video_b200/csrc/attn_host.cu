// Host launcher for the tcgen05 attention kernels (see attn_sm100.cuh).
#include "attn2_sm100.cuh"
#include "common.cuh"
#include <stdlib.h>
#include "launch.cuh"

namespace mjv {

template <int HD, bool CAUSAL>
static int launch_attn_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnArgs& a,
                         dim3 grid, cudaStream_t stream, double work) {
    using L = attn_detail::Smem<HD>;
    auto kfn = attention_kernel<HD, CAUSAL>;
    static PerDeviceOnce once;  // per instantiation and device
    if (once.first()) {
        MJV_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        MJV_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    {
        LaunchScope scope(CAUSAL ? TAG_ATTN_CAUSAL : TAG_ATTN_BIDIR, stream, work);
        kfn<<<grid, attn_detail::kThreads, L::kTotal, stream>>>(tq, tk, tv, a);
    }
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int HD, bool CAUSAL, bool DBG = false, int POLY = 0>
static int launch_attn2_t(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnArgs& a,
                          dim3 grid, cudaStream_t stream, double work) {
    using L = attn2::Smem<HD>;
    auto kfn = attention2_kernel<HD, CAUSAL, DBG, POLY>;
    static PerDeviceOnce once;  // per instantiation and device
    if (once.first()) {
        MJV_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    }
    {
        LaunchScope scope(CAUSAL ? TAG_ATTN_CAUSAL : TAG_ATTN_BIDIR, stream, work);
        kfn<<<grid, attn2::kThreads, L::kTotal, stream>>>(tq, tk, tv, a);
    }
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

// MJV_ATTN_IMPL=1 selects the first-generation kernel (attn_sm100.cuh: one 128-row tile per CTA, 64-column blocks, two
// CTAs per SM) for A/B measurements; the default is attention2_kernel (attn2_sm100.cuh).
// default share of polynomial exponentials (of 8 pairs) per shape; 0 until a measurement says otherwise
constexpr int kPolyVit = 0;
constexpr int kPolyLlm = 0;

static int attn_impl() {
    const char* e = getenv("MJV_ATTN_IMPL");  // read per call so that one process can time both
    return (e != nullptr && e[0] == '1') ? 1 : 2;
}

int attention(const mjv_attn_desc& d, cudaStream_t stream) {
    MJV_CHECK(d.head_dim == 64 || d.head_dim == 128, MJV_ERR_ARG, "attention: head_dim %d unsupported", d.head_dim);
    MJV_CHECK(d.heads > 0 && d.kv_heads > 0 && d.heads % d.kv_heads == 0, MJV_ERR_ARG, "attention: bad head counts");
    const bool varlen = d.cu_seqlens != nullptr;
    if (varlen) MJV_CHECK(d.work != nullptr && d.n_work > 0, MJV_ERR_ARG, "attention: varlen mode needs a work list");
    else MJV_CHECK(d.nseq > 0 && d.seq_len > 0, MJV_ERR_ARG, "attention: batched mode needs nseq/seq_len");
    const int hd = d.head_dim;
    const int impl = attn_impl();
    const int bkv = impl == 2 ? attn2::BKV : attn_detail::BKV;

    const int v_hd = d.v_head_dim > 0 ? d.v_head_dim : hd;
    const int out_hd = d.out_head_dim > 0 ? d.out_head_dim : hd;
    MJV_CHECK(v_hd <= hd && out_hd <= hd && out_hd % 32 == 0, MJV_ERR_ARG,
              "attention: v_head_dim %d / out_head_dim %d must fit the %d-wide tile (out a multiple of 32)", v_hd, out_hd, hd);
    const int qk_stride = d.qk_head_stride > 0 ? d.qk_head_stride : hd;
    MJV_CHECK(qk_stride == hd || (impl == 2 && varlen && hd == 128 && qk_stride == 96), MJV_ERR_ARG,
              "attention: qk_head_stride %d unsupported (head_dim, or 96 under the 128-wide causal varlen tile)", qk_stride);
    CUtensorMap tq, tk, tv;
    if (varlen) {
        uint64_t qd[3] = {(uint64_t)(d.q_col0 + d.heads * qk_stride), (uint64_t)d.total_rows, 1};
        uint64_t qs[3] = {1, (uint64_t)d.ldq, (uint64_t)d.ldq * (uint64_t)d.total_rows};
        uint32_t qb[3] = {64, 128, 1};
        MJV_PROPAGATE(make_tmap_bf16(&tq, d.q, 3, qd, qs, qb));
        uint64_t kd[3] = {(uint64_t)(d.k_col0 + d.kv_heads * qk_stride), (uint64_t)d.total_rows, 1};
        uint64_t ks[3] = {1, (uint64_t)d.ldk, (uint64_t)d.ldk * (uint64_t)d.total_rows};
        uint32_t kb[3] = {64, (uint32_t)bkv, 1};
        MJV_PROPAGATE(make_tmap_bf16(&tk, d.k, 3, kd, ks, kb));
        // rows beyond v_hd do not exist in the plane: the box still spans hd rows and TMA zero-fills the rest
        uint64_t vd[3] = {(uint64_t)(d.vt_cols > 0 ? d.vt_cols : d.total_rows), (uint64_t)v_hd, (uint64_t)d.kv_heads};
        uint64_t vs[3] = {1, (uint64_t)d.vt_ld, (uint64_t)d.vt_plane_stride};
        uint32_t vb[3] = {64, (uint32_t)hd, 1};
        MJV_PROPAGATE(make_tmap_bf16(&tv, d.vt, 3, vd, vs, vb));
    } else {
        uint64_t qd[3] = {(uint64_t)(d.q_col0 + d.heads * hd), (uint64_t)d.seq_len, (uint64_t)d.nseq};
        uint64_t qs[3] = {1, (uint64_t)d.ldq, (uint64_t)d.ldq * (uint64_t)d.seq_len};
        uint32_t qb[3] = {64, 128, 1};
        MJV_PROPAGATE(make_tmap_bf16(&tq, d.q, 3, qd, qs, qb));
        uint64_t kd[3] = {(uint64_t)(d.k_col0 + d.kv_heads * hd), (uint64_t)d.seq_len, (uint64_t)d.nseq};
        uint64_t ks[3] = {1, (uint64_t)d.ldk, (uint64_t)d.ldk * (uint64_t)d.seq_len};
        uint32_t kb[3] = {64, (uint32_t)bkv, 1};
        MJV_PROPAGATE(make_tmap_bf16(&tk, d.k, 3, kd, ks, kb));
        uint64_t vd[3] = {(uint64_t)d.seq_len, (uint64_t)v_hd, (uint64_t)d.nseq * (uint64_t)d.heads};
        uint64_t vs[3] = {1, (uint64_t)d.vt_ld, (uint64_t)d.vt_plane_stride};
        uint32_t vb[3] = {64, (uint32_t)hd, 1};
        MJV_PROPAGATE(make_tmap_bf16(&tv, d.vt, 3, vd, vs, vb));
        MJV_CHECK(d.heads == d.kv_heads, MJV_ERR_ARG, "attention: batched mode expects one V plane per query head");
    }

    AttnArgs a;
    memset(&a, 0, sizeof(a));
    a.seq_len = d.seq_len;
    a.cu_seqlens = d.cu_seqlens;
    a.work = reinterpret_cast<const int2*>(d.work);
    a.vt_starts = d.vt_starts;
    a.heads = d.heads;
    a.kv_group = d.heads / d.kv_heads;
    a.q_col0 = d.q_col0;
    a.k_col0 = d.k_col0;
    a.out = static_cast<__nv_bfloat16*>(d.out);
    a.ldo = (int)d.ldo;
    a.scale_log2 = d.scale * 1.4426950408889634f;
    a.batched = varlen ? 0 : 1;
    a.out_hd = out_hd;
    a.qk_stride = qk_stride;
    a.qk_ksteps = qk_stride / 16;
    a.pv_n = (v_hd + 15) / 16 * 16;

    // algorithmic FLOPs: 4*heads*hd*S^2 per sequence (bidirectional); causal work is supplied by the caller (flops_hint)
    double work = d.flops_hint;
    if (!varlen) work = 4.0 * d.heads * hd * (double)d.seq_len * d.seq_len * d.nseq;
    const int n_qblocks = varlen ? 0 : (d.seq_len + 127) / 128;
    if (impl == 2) {
        // persistent grid: one CTA per SM, each item = two query tiles sharing their K/V stream
        a.pair_heads = (varlen && a.kv_group % 2 == 0) ? 1 : 0;
        if (const char* tp = getenv("MJV_ATTN_TRACE"))  // tools/attn_trace.py: device pointer (hex) of the event buffer
            a.trace = reinterpret_cast<long long*>(strtoull(tp, nullptr, 16));
        // single trailing key on the CUDA cores; needs more key blocks than score buffers (see the q_full wait in the kernel)
        a.kv_tail = (!varlen && !d.causal && hd == 64 && v_hd == hd && d.seq_len % 128 == 1 && d.seq_len > 4 * 128) ? 1 : 0;
        a.k_ptr = static_cast<const __nv_bfloat16*>(d.k);
        a.ldk = d.ldk;
        a.vt_ptr = static_cast<const __nv_bfloat16*>(d.vt);
        a.vt_ld = d.vt_ld;
        a.vt_plane_stride = d.vt_plane_stride;
        a.pair_tail = (!varlen && (n_qblocks & 1) && d.heads % 2 == 0 && d.heads == d.kv_heads) ? 1 : 0;
        const long long n2 = varlen ? (long long)d.n_work * (a.pair_heads ? d.heads / 2 : d.heads)
                             : a.pair_tail ? (long long)d.nseq * (d.heads * (n_qblocks / 2) + d.heads / 2)
                                           : (long long)d.nseq * d.heads * ((n_qblocks + 1) / 2);
        MJV_CHECK(n2 > 0 && n2 < (1LL << 31), MJV_ERR_ARG, "attention: bad item count %lld", n2);
        a.n_items = (int)n2;
        a.n_qblocks = n_qblocks;
        dim3 grid2((unsigned)(n2 < sm_count() ? n2 : sm_count()), 1, 1);
        if (a.trace != nullptr) {  // tools: the debug instantiations exist for the two bench shapes only
            if (hd == 64 && !d.causal) return launch_attn2_t<64, false, true>(tq, tk, tv, a, grid2, stream, work);
            if (hd == 128 && d.causal) return launch_attn2_t<128, true, true>(tq, tk, tv, a, grid2, stream, work);
        }
        // exponentials on the FMA pipe: POLY of 8 pairs (tools/attn_bench.py sweeps MJV_ATTN_POLY; defaults below)
        int poly = (hd == 64 && !d.causal) ? kPolyVit : ((hd == 128 && d.causal) ? kPolyLlm : 0);
        if (const char* pe = getenv("MJV_ATTN_POLY")) poly = atoi(pe);
        if (hd == 64 && !d.causal) {
            if (poly == 2) return launch_attn2_t<64, false, false, 2>(tq, tk, tv, a, grid2, stream, work);
            if (poly == 3) return launch_attn2_t<64, false, false, 3>(tq, tk, tv, a, grid2, stream, work);
            if (poly == 4) return launch_attn2_t<64, false, false, 4>(tq, tk, tv, a, grid2, stream, work);
        }
        if (hd == 128 && d.causal) {
            if (poly == 2) return launch_attn2_t<128, true, false, 2>(tq, tk, tv, a, grid2, stream, work);
            if (poly == 3) return launch_attn2_t<128, true, false, 3>(tq, tk, tv, a, grid2, stream, work);
            if (poly == 4) return launch_attn2_t<128, true, false, 4>(tq, tk, tv, a, grid2, stream, work);
        }
        if (hd == 64) {
            if (d.causal) return launch_attn2_t<64, true>(tq, tk, tv, a, grid2, stream, work);
            return launch_attn2_t<64, false>(tq, tk, tv, a, grid2, stream, work);
        }
        if (d.causal) return launch_attn2_t<128, true>(tq, tk, tv, a, grid2, stream, work);
        return launch_attn2_t<128, false>(tq, tk, tv, a, grid2, stream, work);
    }
    // first generation: 2 CTAs per SM (registers / smem / TMEM allow exactly two), static round-robin over the items
    if (varlen) {
        // its work entries are plain (seq, q_block) pairs: the packed second block of attention2's MHA lists is not understood
        MJV_CHECK(d.heads % d.kv_heads == 0, MJV_ERR_ARG, "attention: bad head counts");
    }
    const long long n_items = varlen ? (long long)d.n_work * d.heads : (long long)d.nseq * d.heads * n_qblocks;
    MJV_CHECK(n_items > 0 && n_items < (1LL << 31), MJV_ERR_ARG, "attention: bad item count %lld", n_items);
    a.n_items = (int)n_items;
    a.n_qblocks = n_qblocks;
    const int max_ctas = 2 * sm_count();
    dim3 grid((unsigned)(n_items < max_ctas ? n_items : max_ctas), 1, 1);
    if (hd == 64) {
        if (d.causal) return launch_attn_t<64, true>(tq, tk, tv, a, grid, stream, work);
        return launch_attn_t<64, false>(tq, tk, tv, a, grid, stream, work);
    }
    if (d.causal) return launch_attn_t<128, true>(tq, tk, tv, a, grid, stream, work);
    return launch_attn_t<128, false>(tq, tk, tv, a, grid, stream, work);
}

}  // namespace mjv

extern "C" int mjv_attention(const mjv_attn_desc* desc, void* stream) {
    if (desc == nullptr) return mjv::set_error(MJV_ERR_ARG, "mjv_attention: null desc");
    return mjv::attention(*desc, static_cast<cudaStream_t>(stream));
}
