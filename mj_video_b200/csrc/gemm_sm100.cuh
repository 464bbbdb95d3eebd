// Persistent, warp-specialised bf16 GEMM for sm_100a:  D[M,N] = A[M,K] · W[N,K]^T  (+ fused epilogue)
//
//   * A (activations) and W (nn.Linear weight, [out,in]) are both K-major, loaded by TMA into 128B-swizzled
//     shared-memory stages; tcgen05.mma (cta_group::1, M=128, N=BN, K=16) accumulates fp32 in TMEM.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warp 2 = TMEM allocator,
//     warps 4..11 = epilogue (TMEM -> registers -> fused math -> global): two warps per TMEM lane quarter, each
//     taking half of the tile's columns, so every SM sub-partition has two epilogue warps to hide latency.
//   * two TMEM accumulator buffers (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//   * static persistent schedule: tile = blockIdx.x + i * gridDim.x, n-fastest so co-resident CTAs share A rows
//     and the whole W stays L2-resident.
//
// Epilogues cover the reference's call sites: K1 patch-embed, K4 ViT qkv, K7/K8 ViT proj/fc1/fc2,
// K10 mlp1, K13+K14 LLM wqkv+RoPE, K16 wo, K17 SwiGLU. They mirror the reference's bf16 rounding points
//  so the result tracks the reference's own bf16 forward, not just the fp32 ideal.
#pragma once
#include "ptx.cuh"

namespace mjv {

enum GemmEpilogue : int {
    EPI_STORE = 0,    // out = bf16(acc + bias)
    EPI_GELU = 1,     // out = bf16(gelu_erf(bf16(acc + bias)))
    EPI_RESID = 2,    // out = bf16(res + bf16(bf16(acc + bias) * scale))        (bias/scale optional)
    EPI_SWIGLU = 3,   // W rows interleaved [32 gate | 32 up]; out[:, j] = bf16(silu(g) * u), N_out = N/2
    EPI_QKV_LLM = 4,  // columns [q heads | k heads | v heads] x 128; RoPE on q,k; V written transposed
    EPI_QKV_VIT = 5,  // columns [q | k | v] x (heads x 64); +bias; q,k -> qk[M,2*C]; V written transposed per tile
    EPI_PATCH = 6,    // patch-embed: out row = m + m/P + 1, + bias + pos_emb[m%P + 1]
    EPI_SCATTER = 7,  // out row = row_map[m]; out = bf16(acc + bias)
};

struct GemmArgs {
    int M, N, K;
    int num_m_blocks, num_n_blocks;
    __nv_bfloat16* out;
    int ldo;
    const __nv_bfloat16* bias;      // [N] or null
    const __nv_bfloat16* residual;  // [M, ldr] or null
    int ldr;
    const __nv_bfloat16* scale;  // [N] layer-scale or null
    // specialised epilogues
    __nv_bfloat16* out_k;      // QKV_LLM: K [M, n_kv*128]
    __nv_bfloat16* out_vt;     // QKV_*: V^T
    const float* rope_cos;     // [max_pos, 64]
    const float* rope_sin;     // [max_pos, 64]
    const int* row_pos;        // QKV_LLM: position id per packed token row
    const int* row_map;        // SCATTER: destination row per m
    const int* row_vt;         // QKV_LLM: V^T column per packed token row (sequence starts 8-aligned for TMA)
    const __nv_bfloat16* pos;  // PATCH: position embedding [P+1, N]
    int n_q_cols;              // QKV: number of q columns
    int n_k_cols;              // QKV: number of k columns
    long long vt_stride_head;  // QKV: elements between consecutive (tile,)head planes of V^T ( = hd * vt_ld )
    int vt_ld;                 // QKV: row stride of V^T (padded token count)
    int seq;                   // QKV_VIT: tokens per tile (1025); PATCH: patches per tile (1024)
    int heads;                 // QKV_VIT: heads
    int v_hd;                  // QKV_LLM: columns per V head in N and rows per V^T plane (multiple of 32, <= 128)
    int qk_slot;               // QKV_LLM: 128 (rotate-half partner 64 columns away) or 96 (partner in the adjacent column)
};

namespace gemm_detail {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;

// Exact-erf GELU on a pair of values with packed fp32x2 math and ONE MUFU per element (the chip is power-capped in
// these GEMMs: every epilogue instruction costs main-loop clock).
// erf by Abramowitz-Stegun 7.1.28: 1 - erf(z) = (1 + a1 z + ... + a6 z^6)^-16 =: r16, |err| <= 3e-7 (far below bf16
// resolution). gelu(x) = 0.5 x (1 + erf(x/sqrt2)) = max(x, 0) - 0.5 |x| r16(|x|/sqrt2), which also keeps the relative
// accuracy of the negative tail.
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
    const uint64_t ax2 = ptx::pack_f32x2(fabsf(x0), fabsf(x1));
    const uint64_t z2 = ptx::fmul2(ax2, ptx::pack_f32x2(0.70710678f, 0.70710678f));
    uint64_t d2 = ptx::ffma2(ptx::pack_f32x2(0.0000430638f, 0.0000430638f), z2, ptx::pack_f32x2(0.0002765672f, 0.0002765672f));
    d2 = ptx::ffma2(d2, z2, ptx::pack_f32x2(0.0001520143f, 0.0001520143f));
    d2 = ptx::ffma2(d2, z2, ptx::pack_f32x2(0.0092705272f, 0.0092705272f));
    d2 = ptx::ffma2(d2, z2, ptx::pack_f32x2(0.0422820123f, 0.0422820123f));
    d2 = ptx::ffma2(d2, z2, ptx::pack_f32x2(0.0705230784f, 0.0705230784f));
    d2 = ptx::ffma2(d2, z2, ptx::pack_f32x2(1.0f, 1.0f));
    float d0, d1;
    ptx::unpack_f32x2(d2, d0, d1);
    uint64_t r2 = ptx::pack_f32x2(__fdividef(1.0f, d0), __fdividef(1.0f, d1));
    r2 = ptx::fmul2(r2, r2);
    r2 = ptx::fmul2(r2, r2);
    r2 = ptx::fmul2(r2, r2);
    r2 = ptx::fmul2(r2, r2);  // r^16 = 1 - erf(|z|)
    const uint64_t h2 = ptx::fmul2(ax2, ptx::pack_f32x2(-0.5f, -0.5f));
    const uint64_t o2 = ptx::ffma2(h2, r2, ptx::pack_f32x2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
    ptx::unpack_f32x2(o2, x0, x1);
}
// silu(g) * u for two pairs with packed math and ONE MUFU per gate value: x * sigmoid(x) = h + h * tanh(h), h = x / 2
// (tanh.approx: relative error 2^-11, a quarter of a bf16 ulp). Applied to the fp32 accumulators and rounded once; the
// reference rounds gate, up and silu(gate) to bf16 on the way (three separate tensors there), a difference of at most
// ~1.5 bf16 ulp of the product in exchange for 7 fewer epilogue instructions per output on a power-capped GEMM.
__device__ __forceinline__ uint32_t swiglu_pack(float g0, float g1, float u0, float u1) {
    const uint64_t h2 = ptx::fmul2(ptx::pack_f32x2(g0, g1), ptx::pack_f32x2(0.5f, 0.5f));
    float h0, h1;
    ptx::unpack_f32x2(h2, h0, h1);
    float t0, t1;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
    const uint64_t s2 = ptx::ffma2(h2, ptx::pack_f32x2(t0, t1), h2);  // silu(g)
    float p0, p1;
    ptx::unpack_f32x2(ptx::fmul2(s2, ptx::pack_f32x2(u0, u1)), p0, p1);
    return ptx::pack_bf16x2(p0, p1);
}

__device__ __forceinline__ void load32_bf16(const __nv_bfloat16* p, uint32_t* u) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 t = __ldg(q + i);
        u[4 * i + 0] = t.x;
        u[4 * i + 1] = t.y;
        u[4 * i + 2] = t.z;
        u[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void store32_bf16(__nv_bfloat16* p, const uint32_t* u) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = make_uint4(u[4 * i + 0], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
}

constexpr int kStagingBytesPerWarp = 32 * 128;  // [32 rows][64 bf16] swizzled tile for one TMA store

template <int BN, int STAGES, int CG = 1>
struct SmemLayout {
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = (BN / CG) * BK * 2;  // cta_group::2: each CTA of the pair stages half of the B tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr bool kTmaStore = (BN == 256);  // 64-column output boxes per epilogue warp need BN/2 >= 128
    static constexpr int kStagingOffset = STAGES * kStageBytes;
    static constexpr int kBarOffset = kStagingOffset + (kTmaStore ? kEpiWarps * kStagingBytesPerWarp : 0);
    static constexpr int kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

}  // namespace gemm_detail

// CG = 1: one CTA per 128 x BN tile. CG = 2: a CTA pair (cluster of 2) per 256 x BN tile with tcgen05 cta_group::2 —
// each CTA holds 128 rows of A, half of the B tile and the accumulator of its 128 rows; the leader CTA issues the MMAs.
template <int BN, int EPI, int STAGES, int CG>
__global__ void __launch_bounds__(gemm_detail::kThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
                    const GemmArgs args) {
    using namespace gemm_detail;
    using L = SmemLayout<BN, STAGES, CG>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * L::kABytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* res_bar = tmem_empty + 2;  // [kEpiWarps] residual-tile TMA loads, one barrier per epilogue warp
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kEpiWarps);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = args.num_m_blocks * args.num_n_blocks;  // tiles of (128*CG) x BN
    const int nkb = (args.K + BK - 1) / BK;
    const int cta_rank = CG == 2 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
    const int tile0 = blockIdx.x / CG;       // first tile of this CTA (pair)
    const int tile_step = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], CG);  // one producer arrival per CTA of the pair (leader's barrier is used)
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tmem_full[a], 1);
            ptx::mbar_init(&tmem_empty[a], kEpiWarps * CG);  // one arrive per epilogue warp (of both CTAs)
        }
        for (int w = 0; w < kEpiWarps; ++w) ptx::mbar_init(&res_bar[w], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (CG == 2) {
            ptx::tmem_alloc_cg2(tmem_slot, 2 * BN);
            ptx::tmem_relinquish_cg2();
        } else {
            ptx::tmem_alloc(tmem_slot, 2 * BN);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync();  // peer barriers are initialised before any remote arrive / TMA
    else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                const int m0 = (tile / args.num_n_blocks) * (BM * CG) + cta_rank * BM;
                const int n0 = (tile % args.num_n_blocks) * BN + cta_rank * (BN / CG);
                for (int kb = 0; kb < nkb; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if constexpr (CG == 2) {
                        // both CTAs' bytes complete on the leader's barrier
                        if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
                        else ptx::mbar_arrive_leader(&full_bar[stage]);
                        ptx::tma_load_2d_cg2(sA + stage * L::kABytes, &tmA, &full_bar[stage], kb * BK, m0, ptx::kEvictNormal);
                        ptx::tma_load_2d_cg2(sB + stage * L::kBBytes, &tmB, &full_bar[stage], kb * BK, n0, ptx::kEvictLast);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
                    ptx::tma_load_2d(sA + stage * L::kABytes, &tmA, &full_bar[stage], kb * BK, m0, ptx::kEvictNormal);
                    ptx::tma_load_2d(sB + stage * L::kBBytes, &tmB, &full_bar[stage], kb * BK, n0, ptx::kEvictLast);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1 && cta_rank == 0) {
        // ===================================================== MMA issuer (leader CTA only when CG == 2)
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BM * CG, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t adesc = ptx::make_desc_sw128(ptx::smem_u32(sA + stage * L::kABytes));
                    const uint64_t bdesc = ptx::make_desc_sw128(ptx::smem_u32(sB + stage * L::kBBytes));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if constexpr (CG == 2)
                            ptx::mma_bf16_ss_cg2(d_tmem, ptx::desc_advance_k16(adesc, k), ptx::desc_advance_k16(bdesc, k),
                                                 idesc, (kb | k) != 0 ? 1u : 0u);
                        else
                            ptx::mma_bf16_ss(d_tmem, ptx::desc_advance_k16(adesc, k), ptx::desc_advance_k16(bdesc, k),
                                             idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if constexpr (CG == 2) {
                        ptx::mma_commit_cg2(&empty_bar[stage]);  // frees the smem slot in BOTH CTAs
                        if (kb == nkb - 1) ptx::mma_commit_cg2(&tmem_full[acc]);
                    } else {
                        ptx::mma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
                        if (kb == nkb - 1) ptx::mma_commit(&tmem_full[acc]);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue (warp%4 selects the TMEM lane quarter,
        // (warp-4)/4 the column half). With BN == 256 results leave through a per-warp swizzled staging tile and a TMA
        // bulk store ([32 rows x 64 columns] boxes); per-thread row-strided global stores cost up to 40% of the kernel.
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        constexpr int HB = BN / 2;
        constexpr bool kTma = L::kTmaStore;
        const int c_begin = half * HB;
        const int c_end = c_begin + HB;
        uint8_t* stg = smem + L::kStagingOffset + (warp - 4) * kStagingBytesPerWarp;
        auto stg_wait = [&]() {  // the previous bulk store has finished reading the staging tile
            if (lane == 0) ptx::bulk_wait_read0();
            __syncwarp();
        };
        auto stg_put = [&](int hc, const uint32_t* o) {  // 32 columns (16 packed pairs) of this thread's row
            uint8_t* rowp = stg + lane * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g)
                *reinterpret_cast<uint4*>(rowp + (((hc * 4 + g) ^ (lane & 7)) * 16)) =
                    make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
        };
        auto stg_store = [&](const CUtensorMap* tm, int col, int row) {
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                ptx::tma_store_2d(tm, stg, col, row);
                ptx::bulk_commit();
            }
        };
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t res_phase = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            const int m0 = (tile / args.num_n_blocks) * (BM * CG) + cta_rank * BM;
            const int n0 = (tile % args.num_n_blocks) * BN;
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < args.M;
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t t_base = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);

            if constexpr (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_RESID || EPI == EPI_PATCH ||
                          EPI == EPI_SCATTER) {
                long long orow = row;
                int orow0 = m0 + q * 32;  // first output row of this warp's 32-row group (TMA store)
                int prow = 0;
                if constexpr (EPI == EPI_PATCH) {
                    orow = row + row / args.seq + 1;
                    orow0 = m0 + q * 32 + (m0 + q * 32) / args.seq + 1;  // 32-row groups never straddle two images
                    prow = row % args.seq + 1;
                }
                if constexpr (EPI == EPI_SCATTER) orow = row_ok ? args.row_map[row] : 0;
                constexpr bool tma_out = kTma && EPI != EPI_SCATTER;
                constexpr bool tma_res = tma_out && EPI == EPI_RESID;  // residual tile arrives by TMA into the staging tile
#pragma unroll 1
                for (int c = c_begin; c < c_end; c += 32) {
                    const int hc = ((c - c_begin) >> 5) & 1;
                    if (tma_out && hc == 0) {
                        stg_wait();
                        if (tma_res && lane == 0) {
                            ptx::mbar_expect_tx(&res_bar[warp - 4], kStagingBytesPerWarp);
                            ptx::tma_load_2d(stg, &tmO2, &res_bar[warp - 4], n0 + c, orow0, ptx::kEvictFirst);
                        }
                    }
                    uint32_t v[32];
                    ptx::tmem_ld_x32(t_base + c, v);
                    ptx::tmem_wait_ld();
                    const int n = n0 + c;
                    uint32_t o[16];
                    uint32_t b[16];
                    if (args.bias != nullptr) {
                        load32_bf16(args.bias + n, b);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) b[i] = 0u;
                    }
                    uint32_t r[16], s[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = s[i] = 0u;
                    if constexpr (tma_res) {
                        if (args.scale != nullptr) load32_bf16(args.scale + n, s);
                        if (hc == 0) {
                            ptx::mbar_wait(&res_bar[warp - 4], res_phase);
                            res_phase ^= 1;
                        }
                        const uint8_t* rowp = stg + lane * 128;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const uint4 t4 = *reinterpret_cast<const uint4*>(rowp + (((hc * 4 + g) ^ (lane & 7)) * 16));
                            r[4 * g] = t4.x;
                            r[4 * g + 1] = t4.y;
                            r[4 * g + 2] = t4.z;
                            r[4 * g + 3] = t4.w;
                        }
                    }
                    if (row_ok) {
                        if constexpr (EPI == EPI_RESID && !tma_res) {
                            load32_bf16(args.residual + static_cast<long long>(row) * args.ldr + n, r);
                            if (args.scale != nullptr) load32_bf16(args.scale + n, s);
                        }
                        if constexpr (EPI == EPI_PATCH) load32_bf16(args.pos + static_cast<long long>(prow) * args.N + n, r);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float x0, x1;
                        ptx::unpack_f32x2(ptx::fadd2(ptx::pack_f32x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])),
                                                     ptx::pack_f32x2(ptx::bf16_lo(b[i]), ptx::bf16_hi(b[i]))),
                                          x0, x1);
                        // GELU is applied to the fp32 pre-activation (the reference rounds fc1's output to bf16 first;
                        // the difference is below one bf16 ulp of the result). STORE/PATCH round once in the final pack.
                        if constexpr (EPI == EPI_GELU) gelu_erf2(x0, x1);
                        if constexpr (EPI == EPI_RESID) {
                            x0 = ptx::bf16_round(x0);
                            x1 = ptx::bf16_round(x1);
                            if (args.scale != nullptr) {
                                x0 = ptx::bf16_round(x0 * ptx::bf16_lo(s[i]));
                                x1 = ptx::bf16_round(x1 * ptx::bf16_hi(s[i]));
                            }
                            x0 += ptx::bf16_lo(r[i]);
                            x1 += ptx::bf16_hi(r[i]);
                        }
                        if constexpr (EPI == EPI_PATCH) {
                            x0 = ptx::bf16_round(x0) + ptx::bf16_lo(r[i]);
                            x1 = ptx::bf16_round(x1) + ptx::bf16_hi(r[i]);
                        }
                        o[i] = ptx::pack_bf16x2(x0, x1);
                    }
                    if constexpr (tma_out) {
                        stg_put(hc, o);
                        if (hc == 1) stg_store(&tmO, n - 32, orow0);
                    } else {
                        if (row_ok) store32_bf16(args.out + orow * args.ldo + n, o);
                    }
                }
            } else if constexpr (EPI == EPI_SWIGLU) {
#pragma unroll 1
                for (int c = c_begin; c < c_end; c += 64) {
                    const int hc = ((c - c_begin) >> 6) & 1;
                    if (kTma && hc == 0) stg_wait();
                    uint32_t g[32], u[32];
                    ptx::tmem_ld_x32(t_base + c, g);
                    ptx::tmem_ld_x32(t_base + c + 32, u);
                    ptx::tmem_wait_ld();
                    uint32_t o[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        o[i] = swiglu_pack(__uint_as_float(g[2 * i]), __uint_as_float(g[2 * i + 1]),
                                           __uint_as_float(u[2 * i]), __uint_as_float(u[2 * i + 1]));
                    if constexpr (kTma) {
                        stg_put(hc, o);
                        if (hc == 1) stg_store(&tmO, (n0 + c_begin) >> 1, m0 + q * 32);
                    } else {
                        if (row_ok) store32_bf16(args.out + static_cast<long long>(row) * args.ldo + ((n0 + c) >> 1), o);
                    }
                }
            } else if constexpr (EPI == EPI_QKV_LLM) {
                // head_dim 128: rotate-half pairs column j with j+64, i.e. 32-column chunk c with chunk c+2.
                const int pos = row_ok ? args.row_pos[row] : 0;
                // BN == 256: this warp owns head `half` of the tile; BN == 128: the two halves share the single head
                constexpr int kHeadsPerTile = BN / 128;
#pragma unroll 1
                for (int hh = 0; hh < kHeadsPerTile; ++hh) {
                    if (kTma && hh != half) continue;
                    const int h0 = hh * 128;
                    const int n = n0 + h0;  // first column of this head
                    if (n < args.n_q_cols + args.n_k_cols) {
                        const bool is_q = n < args.n_q_cols;
                        const int ncol = is_q ? n : n - args.n_q_cols;
                        __nv_bfloat16* dst =
                            is_q ? args.out + static_cast<long long>(row) * args.ldo + n
                                 : args.out_k + static_cast<long long>(row) * args.n_k_cols + (n - args.n_q_cols);
                        const CUtensorMap* tm = is_q ? &tmO : &tmO2;
                        if (args.qk_slot == 96) {
                            // native head_dim 96: heads are three 32-column chunks wide and the rotary pair (d, d + 48) sits
                            // in adjacent columns (2d, 2d + 1), so every chunk rotates on its own: chunk g of the q (or k)
                            // columns belongs to head g / 3 and holds pairs (g % 3) * 16 ... + 15 of that head
#pragma unroll 1
                            for (int cp = 0; cp < 2; ++cp) {
                                if (!kTma && cp != half) continue;   // BN == 128: each half does one 64-column box
                                if (kTma) stg_wait();
#pragma unroll 1
                                for (int cc = 0; cc < 2; ++cc) {
                                    const int c = cp * 64 + cc * 32;
                                    uint32_t x[32];
                                    ptx::tmem_ld_x32(t_base + h0 + c, x);
                                    ptx::tmem_wait_ld();
                                    uint32_t o[16];
#pragma unroll
                                    for (int i = 0; i < 16; ++i) o[i] = 0u;
                                    if (row_ok) {
                                        const int p0 = (((ncol + c) >> 5) % 3) * 16;
                                        const float4* cs = reinterpret_cast<const float4*>(args.rope_cos + static_cast<long long>(pos) * 64 + p0);
                                        const float4* sn = reinterpret_cast<const float4*>(args.rope_sin + static_cast<long long>(pos) * 64 + p0);
#pragma unroll
                                        for (int i = 0; i < 4; ++i) {
                                            const float4 cv = __ldg(cs + i);
                                            const float4 sv = __ldg(sn + i);
                                            const float cc4[4] = {cv.x, cv.y, cv.z, cv.w};
                                            const float ss4[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
                                            for (int j = 0; j < 4; ++j) {
                                                const float p1 = ptx::bf16_round(__uint_as_float(x[8 * i + 2 * j]));
                                                const float p2 = ptx::bf16_round(__uint_as_float(x[8 * i + 2 * j + 1]));
                                                // q*cos + rotate_half(q)*sin, each product and the sum rounded to bf16
                                                const float a = ptx::bf16_round(p1 * cc4[j]) + ptx::bf16_round(-p2 * ss4[j]);
                                                const float b = ptx::bf16_round(p2 * cc4[j]) + ptx::bf16_round(p1 * ss4[j]);
                                                o[4 * i + j] = ptx::pack_bf16x2(a, b);
                                            }
                                        }
                                    }
                                    if constexpr (kTma) {
                                        stg_put(cc, o);
                                    } else {
                                        if (row_ok) store32_bf16(dst + c, o);
                                    }
                                }
                                if constexpr (kTma) stg_store(tm, ncol + cp * 64, m0 + q * 32);
                            }
                            continue;
                        }
                        uint32_t hi_keep[2][16];  // rotated second halves (columns 64..127), stored after the first box
#pragma unroll
                        for (int cp = 0; cp < 2; ++cp) {
                            if (!kTma && cp != half) continue;  // BN == 128: each half does one rope pair
                            const int c = cp * 32;
                            if (kTma && cp == 0) stg_wait();
                            uint32_t x1[32], x2[32];
                            ptx::tmem_ld_x32(t_base + h0 + c, x1);
                            ptx::tmem_ld_x32(t_base + h0 + c + 64, x2);
                            ptx::tmem_wait_ld();
                            uint32_t o1[16], o2[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) o1[i] = o2[i] = 0u;
                            if (row_ok) {
                                const float4* cs = reinterpret_cast<const float4*>(args.rope_cos + static_cast<long long>(pos) * 64 + c);
                                const float4* sn = reinterpret_cast<const float4*>(args.rope_sin + static_cast<long long>(pos) * 64 + c);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 cv = __ldg(cs + i);
                                    const float4 sv = __ldg(sn + i);
                                    const float cc[4] = {cv.x, cv.y, cv.z, cv.w};
                                    const float ss[4] = {sv.x, sv.y, sv.z, sv.w};
                                    float a[4], b[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const float p1 = ptx::bf16_round(__uint_as_float(x1[4 * i + j]));
                                        const float p2 = ptx::bf16_round(__uint_as_float(x2[4 * i + j]));
                                        // q*cos + rotate_half(q)*sin, each product and the sum rounded to bf16
                                        a[j] = ptx::bf16_round(p1 * cc[j]) + ptx::bf16_round(-p2 * ss[j]);
                                        b[j] = ptx::bf16_round(p2 * cc[j]) + ptx::bf16_round(p1 * ss[j]);
                                    }
                                    o1[2 * i] = ptx::pack_bf16x2(a[0], a[1]);
                                    o1[2 * i + 1] = ptx::pack_bf16x2(a[2], a[3]);
                                    o2[2 * i] = ptx::pack_bf16x2(b[0], b[1]);
                                    o2[2 * i + 1] = ptx::pack_bf16x2(b[2], b[3]);
                                }
                            }
                            if constexpr (kTma) {
                                stg_put(cp, o1);
#pragma unroll
                                for (int i = 0; i < 16; ++i) hi_keep[cp][i] = o2[i];
                            } else {
                                if (row_ok) {
                                    store32_bf16(dst + c, o1);
                                    store32_bf16(dst + c + 64, o2);
                                }
                            }
                        }
                        if constexpr (kTma) {
                            stg_store(tm, ncol, m0 + q * 32);       // columns 0..63 of the head
                            stg_wait();
                            stg_put(0, hi_keep[0]);
                            stg_put(1, hi_keep[1]);
                            stg_store(tm, ncol + 64, m0 + q * 32);  // columns 64..127
                        }
                    } else {
                        // V columns: transposed store, V^T[head][d][token]. V heads are v_hd (a multiple of 32) columns wide
                        // in N — narrower than the 128-wide rotary slots of q and k when the model's head_dim is — so a
                        // 32-column chunk never straddles two heads.
                        const int vcol0 = n - args.n_q_cols - args.n_k_cols;  // first V column of this 128-column group
                        const int vtc = (row_ok && args.row_vt != nullptr) ? args.row_vt[row] : row;
                        const int cb = kTma ? 0 : half * 64;
                        const int ce = kTma ? 128 : half * 64 + 64;
#pragma unroll 1
                        for (int c = cb; c < ce; c += 32) {
                            if (n + c >= args.N) break;  // the last group may be partial when v_hd < 128
                            uint32_t v[32];
                            ptx::tmem_ld_x32(t_base + h0 + c, v);
                            ptx::tmem_wait_ld();
                            const int vhead = (vcol0 + c) / args.v_hd;
                            const int d0 = (vcol0 + c) - vhead * args.v_hd;
                            __nv_bfloat16* dst = args.out_vt + static_cast<long long>(vhead) * args.vt_stride_head +
                                                 static_cast<long long>(d0) * args.vt_ld + vtc;
                            if (row_ok) {
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    dst[static_cast<long long>(i) * args.vt_ld] = __float2bfloat16_rn(__uint_as_float(v[i]));
                            }
                        }
                    }
                }
            } else if constexpr (EPI == EPI_QKV_VIT) {
                const int t = row / args.seq;
                const int s = row - t * args.seq;
                const bool qk_tile = n0 < args.n_q_cols + args.n_k_cols;  // tiles never straddle the k|v boundary
#pragma unroll 1
                for (int c = c_begin; c < c_end; c += 32) {
                    const int hc = ((c - c_begin) >> 5) & 1;
                    if (kTma && qk_tile && hc == 0) stg_wait();
                    uint32_t v[32];
                    ptx::tmem_ld_x32(t_base + c, v);
                    ptx::tmem_wait_ld();
                    const int n = n0 + c;
                    uint32_t b[16];
                    load32_bf16(args.bias + n, b);
                    if (n < args.n_q_cols + args.n_k_cols) {
                        uint32_t o[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            o[i] = ptx::pack_bf16x2(__uint_as_float(v[2 * i]) + ptx::bf16_lo(b[i]),
                                                    __uint_as_float(v[2 * i + 1]) + ptx::bf16_hi(b[i]));
                        if constexpr (kTma) {
                            stg_put(hc, o);
                            if (hc == 1) stg_store(&tmO, n - 32, m0 + q * 32);
                        } else {
                            if (row_ok) store32_bf16(args.out + static_cast<long long>(row) * args.ldo + n, o);
                        }
                    } else if (row_ok) {
                        const int vcol = n - args.n_q_cols - args.n_k_cols;  // head*64 + d
                        const int head = vcol >> 6;
                        const int d0 = vcol & 63;
                        __nv_bfloat16* dst = args.out_vt +
                                             (static_cast<long long>(t) * args.heads + head) * args.vt_stride_head +
                                             static_cast<long long>(d0) * args.vt_ld + s;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            dst[static_cast<long long>(2 * i) * args.vt_ld] =
                                __float2bfloat16_rn(__uint_as_float(v[2 * i]) + ptx::bf16_lo(b[i]));
                            dst[static_cast<long long>(2 * i + 1) * args.vt_ld] =
                                __float2bfloat16_rn(__uint_as_float(v[2 * i + 1]) + ptx::bf16_hi(b[i]));
                        }
                    }
                }
            }

            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 2) ptx::mbar_arrive_leader(&tmem_empty[acc]);
                else ptx::mbar_arrive(&tmem_empty[acc]);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (kTma && lane == 0) ptx::bulk_wait0();  // all bulk stores of this warp have landed before the CTA exits
    }

    ptx::tc_fence_before();
    if constexpr (CG == 2) ptx::cluster_sync();  // no CTA of the pair may exit while the other can still signal it
    else __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        if constexpr (CG == 2) ptx::tmem_dealloc_cg2(tmem_base, 2 * BN);
        else ptx::tmem_dealloc(tmem_base, 2 * BN);
    }
}

}  // namespace mjv
