// Host launcher for the tcgen05 GEMM family (see gemm_sm100.cuh).
#include <stdlib.h>

#include "common.cuh"
#include "gemm_sm100.cuh"
#include "launch.cuh"

namespace mjv {

template <int BN, int EPI, int STAGES, int CG>
static int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                         const CUtensorMap& tmO2, const GemmArgs& args, cudaStream_t stream) {
    using L = gemm_detail::SmemLayout<BN, STAGES, CG>;
    auto kfn = gemm_bf16_tn_kernel<BN, EPI, STAGES, CG>;
    static PerDeviceOnce once;  // per instantiation and device
    if (once.first()) {
        MJV_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    }
    const int tiles = args.num_m_blocks * args.num_n_blocks;
    const int max_groups = sm_count() / CG;
    const int groups = tiles < max_groups ? tiles : max_groups;
    {
        LaunchScope scope(TAG_GEMM_STORE + EPI, stream, 2.0 * args.M * args.N * args.K);
        if constexpr (CG == 1) {
            kfn<<<groups, gemm_detail::kThreads, L::kTotal, stream>>>(tmA, tmB, tmO, tmO2, args);
        } else {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3(groups * CG);
            cfg.blockDim = dim3(gemm_detail::kThreads);
            cfg.dynamicSmemBytes = L::kTotal;
            cfg.stream = stream;
            cudaLaunchAttribute attr;
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = CG;
            attr.val.clusterDim.y = 1;
            attr.val.clusterDim.z = 1;
            cfg.attrs = &attr;
            cfg.numAttrs = 1;
            MJV_CUDA_OK(cudaLaunchKernelEx(&cfg, kfn, tmA, tmB, tmO, tmO2, args));
        }
    }
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int BN, int STAGES, int CG>
static int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                        const CUtensorMap& tmO2, const GemmArgs& a, cudaStream_t s) {
    switch (epi) {
        case EPI_STORE: return launch_gemm_t<BN, EPI_STORE, STAGES, CG>(tmA, tmB, tmO, tmO2, a, s);
        case EPI_GELU: return launch_gemm_t<BN, EPI_GELU, STAGES, CG>(tmA, tmB, tmO, tmO2, a, s);
        case EPI_RESID: return launch_gemm_t<BN, EPI_RESID, STAGES, CG>(tmA, tmB, tmO, tmO2, a, s);
        case EPI_SWIGLU: return launch_gemm_t<BN, EPI_SWIGLU, STAGES, CG>(tmA, tmB, tmO, tmO2, a, s);
        case EPI_QKV_LLM: return launch_gemm_t<BN, EPI_QKV_LLM, STAGES, CG>(tmA, tmB, tmO, tmO2, a, s);
        case EPI_QKV_VIT: return launch_gemm_t<BN, EPI_QKV_VIT, STAGES, CG>(tmA, tmB, tmO, tmO2, a, s);
        case EPI_PATCH: return launch_gemm_t<BN, EPI_PATCH, STAGES, CG>(tmA, tmB, tmO, tmO2, a, s);
        case EPI_SCATTER: return launch_gemm_t<BN, EPI_SCATTER, STAGES, CG>(tmA, tmB, tmO, tmO2, a, s);
        default: return set_error(MJV_ERR_ARG, "unknown GEMM epilogue %d", epi);
    }
}

int gemm_bf16(const mjv_gemm_desc& d, cudaStream_t stream) {
    MJV_CHECK(d.M > 0 && d.N > 0 && d.K > 0, MJV_ERR_ARG, "gemm: empty problem M=%d N=%d K=%d", d.M, d.N, d.K);
    MJV_CHECK(d.lda % 8 == 0 && d.ldw % 8 == 0 && d.K % 8 == 0, MJV_ERR_ARG,
              "gemm: lda/ldw/K must be multiples of 8 (lda=%lld ldw=%lld K=%d)", (long long)d.lda, (long long)d.ldw,
              d.K);
    int bn = d.block_n;
    if (bn == 0) bn = (d.N % 256 == 0) ? 256 : 128;
    MJV_CHECK((bn == 128 || bn == 256) && d.N % bn == 0, MJV_ERR_ARG, "gemm: N=%d not a multiple of block_n=%d", d.N,
              bn);
    if (d.epilogue == EPI_QKV_LLM) {
        MJV_CHECK(d.n_q_cols % 128 == 0 && d.n_k_cols % 128 == 0 && d.rope_cos && d.rope_sin && d.row_pos &&
                      d.out_k && d.out_vt,
                  MJV_ERR_ARG, "gemm: QKV_LLM needs head_dim 128 and rope tables");
        const int v_hd = d.v_head_dim > 0 ? d.v_head_dim : 128;
        MJV_CHECK(v_hd % 32 == 0 && v_hd <= 128 && (d.N - d.n_q_cols - d.n_k_cols) % v_hd == 0, MJV_ERR_ARG,
                  "gemm: QKV_LLM v_head_dim %d must be a multiple of 32 dividing the V columns", v_hd);
    }
    if (d.epilogue == EPI_QKV_VIT)
        MJV_CHECK(d.bias && d.out_vt && d.seq > 0 && d.heads > 0, MJV_ERR_ARG, "gemm: QKV_VIT needs bias/out_vt");
    if (d.epilogue == EPI_RESID) MJV_CHECK(d.residual != nullptr, MJV_ERR_ARG, "gemm: RESID needs residual");
    if (d.epilogue == EPI_PATCH) MJV_CHECK(d.pos_emb && d.seq > 0, MJV_ERR_ARG, "gemm: PATCH needs pos_emb/seq");
    if (d.epilogue == EPI_SCATTER) MJV_CHECK(d.row_map != nullptr, MJV_ERR_ARG, "gemm: SCATTER needs row_map");

    // CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles) whenever the tile width allows it
    int cg = d.cta_group;
    if (cg == 0) {
        static const int env_cg = getenv("MJV_GEMM_CG") ? atoi(getenv("MJV_GEMM_CG")) : 0;
        // measured inside the (power-capped) bench step on B200: pairs win on every shape of the path, +5..15%
        cg = env_cg != 0 ? env_cg : 2;
    }
    if (bn != 256 || d.M <= 128) cg = 1;
    MJV_CHECK(cg == 1 || cg == 2, MJV_ERR_ARG, "gemm: cta_group must be 1 or 2");
    CUtensorMap tmA, tmB;
    MJV_PROPAGATE(make_tmap_2d(&tmA, d.a, (uint64_t)d.M, (uint64_t)d.K, (uint64_t)d.lda, 128));
    MJV_PROPAGATE(make_tmap_2d(&tmB, d.w, (uint64_t)d.N, (uint64_t)d.K, (uint64_t)d.ldw, (uint32_t)(bn / cg)));

    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.M = d.M;
    a.N = d.N;
    a.K = d.K;
    a.num_m_blocks = (d.M + 128 * cg - 1) / (128 * cg);
    a.num_n_blocks = d.N / bn;
    a.out = static_cast<__nv_bfloat16*>(d.out);
    a.ldo = (int)d.ldo;
    a.bias = static_cast<const __nv_bfloat16*>(d.bias);
    a.residual = static_cast<const __nv_bfloat16*>(d.residual);
    a.ldr = (int)d.ldr;
    a.scale = static_cast<const __nv_bfloat16*>(d.scale);
    a.out_k = static_cast<__nv_bfloat16*>(d.out_k);
    a.out_vt = static_cast<__nv_bfloat16*>(d.out_vt);
    a.rope_cos = d.rope_cos;
    a.rope_sin = d.rope_sin;
    a.row_pos = d.row_pos;
    a.row_map = d.row_map;
    a.row_vt = d.row_vt;
    a.pos = static_cast<const __nv_bfloat16*>(d.pos_emb);
    a.n_q_cols = d.n_q_cols;
    a.n_k_cols = d.n_k_cols;
    a.vt_stride_head = d.vt_stride_head;
    a.vt_ld = d.vt_ld;
    a.seq = d.seq;
    a.heads = d.heads;
    a.v_hd = d.v_head_dim > 0 ? d.v_head_dim : 128;
    a.qk_slot = d.qk_head_slot > 0 ? d.qk_head_slot : 128;

    // output tensor maps for the TMA-store epilogue (BN == 256): [32 rows x 64 columns] boxes, 128B swizzle
    CUtensorMap tmO, tmO2;
    memset(&tmO, 0, sizeof(tmO));
    memset(&tmO2, 0, sizeof(tmO2));
    if (bn == 256 && d.epilogue != EPI_SCATTER) {
        uint64_t rows = (uint64_t)d.M, cols = (uint64_t)d.N;
        if (d.epilogue == EPI_SWIGLU) cols = (uint64_t)d.N / 2;
        if (d.epilogue == EPI_PATCH) rows = (uint64_t)d.M + (uint64_t)d.M / (uint64_t)d.seq;
        if (d.epilogue == EPI_QKV_VIT) cols = (uint64_t)(d.n_q_cols + d.n_k_cols);
        if (d.epilogue == EPI_QKV_LLM) cols = (uint64_t)d.n_q_cols;
        MJV_CHECK(d.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0, MJV_ERR_ARG,
                  "gemm: output must be 16-byte aligned with ldo a multiple of 8");
        MJV_PROPAGATE(make_tmap_2d(&tmO, d.out, rows, cols, (uint64_t)d.ldo, 32));
        if (d.epilogue == EPI_QKV_LLM)
            MJV_PROPAGATE(make_tmap_2d(&tmO2, d.out_k, rows, (uint64_t)d.n_k_cols, (uint64_t)d.n_k_cols, 32));
        if (d.epilogue == EPI_PATCH) MJV_CHECK(d.seq % 32 == 0, MJV_ERR_ARG, "gemm: PATCH needs seq %% 32 == 0");
        if (d.epilogue == EPI_RESID) {
            MJV_CHECK(d.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(d.residual) & 15) == 0, MJV_ERR_ARG,
                      "gemm: residual must be 16-byte aligned with ldr a multiple of 8");
            MJV_PROPAGATE(make_tmap_2d(&tmO2, d.residual, rows, cols, (uint64_t)d.ldr, 32));
        }
    }
    if (bn == 256 && cg == 2) return dispatch_epi<256, 6, 2>(d.epilogue, tmA, tmB, tmO, tmO2, a, stream);
    if (bn == 256) return dispatch_epi<256, 4, 1>(d.epilogue, tmA, tmB, tmO, tmO2, a, stream);
    return dispatch_epi<128, 6, 1>(d.epilogue, tmA, tmB, tmO, tmO2, a, stream);
}

}  // namespace mjv

extern "C" int mjv_gemm_bf16(const mjv_gemm_desc* desc, void* stream) {
    if (desc == nullptr) return mjv::set_error(MJV_ERR_ARG, "mjv_gemm_bf16: null desc");
    return mjv::gemm_bf16(*desc, static_cast<cudaStream_t>(stream));
}
