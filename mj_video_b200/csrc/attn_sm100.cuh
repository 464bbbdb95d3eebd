// Fused softmax attention on tcgen05 / TMEM / TMA for sm_100a.
//
// Replaces K6  flash_attn_varlen_qkvpacked_func(causal=False)  modeling_intern_vit.py:72-75 / _naive_attn :210-227
//                        K15 flash_attn_func / flash_attn_varlen_func(causal=True) modeling_internlm2.py:498-561 / eager :393-410
//
// One CTA = one (sequence, head, 128-query block); KV blocks of 64. Software-pipelined over KV blocks j:
//   MMA thread : S_{j+1} = Q·K_{j+1}ᵀ is issued BEFORE waiting for P_j (S is double-buffered in TMEM)  -> s_full[b]
//   softmax    : 128 threads, one query row each (TMEM lane == row): ONE tcgen05.ld pass brings the 64 scores of the
//                row into registers, row max, p = exp2(s·c − m) with packed fp32x2 math; P (bf16) is written BACK INTO
//                TMEM over the S buffer it came from (tcgen05.st), one barrier arrival per warp        -> p_full
//   MMA thread : O += P_j·V_j (A = P from TMEM, B = Vᵀ tile from smem, accumulator stays in TMEM)     -> pv_done, kv_empty
//                (tcgen05.mma executes in issue order, so S_{j+2} cannot overwrite P_j before P_j·V_j has read it)
// so the softmax of block j+1 starts as soon as block j's P is handed over, while P_j·V_j runs on the tensor pipe.
// The output accumulator never leaves TMEM until the end; the running max is updated lazily (only when it grows by more
// than 2^8), in which case O is rescaled in place with tcgen05.ld/st after waiting for the P·V in flight (pv_done).
// Two CTAs per SM (<=112 KB smem, 256 TMEM columns each).
//
// V is consumed TRANSPOSED ([head][d][token], written that way by the QKV GEMM epilogue) so that both MMAs use the same
// K-major operand layout as the GEMM kernel.
#pragma once
#include "ptx.cuh"

namespace mjv {

struct AttnArgs {
    int seq_len;            // fixed sequence length when cu_seqlens == nullptr (ViT: 1025)
    const int* cu_seqlens;  // [nseq+1] packed row offsets (LLM) or nullptr
    const int2* work;       // [(seq, q_block)] schedule (heavy blocks first) or nullptr -> (blockIdx.z, blockIdx.x)
    const int* vt_starts;   // varlen: first V^T column of each sequence (8-aligned) or nullptr (= cu_seqlens)
    int heads;              // query heads
    int kv_group;           // query heads per kv head
    int q_col0, k_col0;     // first q / k column inside their tensor maps
    __nv_bfloat16* out;     // [rows, ldo]
    int ldo;
    float scale_log2;       // softmax scale * log2(e)
    int batched;            // 1: maps carry a per-sequence batch dim and rows are sequence-relative (ViT)
};

namespace attn_detail {
constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int kThreads = 256;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain: p stays below 2^8

template <int HD>
struct Smem {
    static constexpr int kStages = HD == 64 ? 3 : 2;
    static constexpr int kQBytes = BQ * HD * 2;
    static constexpr int kKBytes = BKV * HD * 2;
    static constexpr int kVBytes = HD * BKV * 2;
    static constexpr int kQOff = 0;
    static constexpr int kKOff = kQOff + kQBytes;
    static constexpr int kVOff = kKOff + kStages * kKBytes;
    static constexpr int kBarOff = kVOff + kStages * kVBytes;
    static constexpr int kTotal = kBarOff + 256;
};
}  // namespace attn_detail

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(attn_detail::kThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnArgs args) {
    using namespace attn_detail;
    using L = Smem<HD>;
    constexpr int ST = L::kStages;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem + L::kQOff;
    uint8_t* sK = smem + L::kKOff;
    uint8_t* sV = smem + L::kVOff;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;         // [ST]
    uint64_t* v_full = bars + 1 + ST;    // [ST]
    uint64_t* kv_empty = bars + 1 + 2 * ST;  // [ST]
    uint64_t* s_full = bars + 1 + 3 * ST;    // [2]
    uint64_t* p_full = s_full + 2;
    uint64_t* pv_done = s_full + 3;
    uint64_t* o_done = s_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- work item
    int seq, qblk;
    if (args.work != nullptr) {
        const int2 w = args.work[blockIdx.x];
        seq = w.x;
        qblk = w.y;
    } else {
        seq = blockIdx.z;
        qblk = blockIdx.x;
    }
    const int h = blockIdx.y;
    int len, row_base;
    if (args.cu_seqlens != nullptr) {
        row_base = args.cu_seqlens[seq];
        len = args.cu_seqlens[seq + 1] - row_base;
    } else {
        len = args.seq_len;
        row_base = seq * len;
    }
    const int q0 = qblk * BQ;
    const int kv_end = CAUSAL ? min(len, q0 + BQ) : len;
    const int n_blocks = (kv_end + BKV - 1) / BKV;
    const int tma_row0 = args.batched ? 0 : row_base;
    const int tma_batch = args.batched ? seq : 0;
    const int kvh = h / args.kv_group;
    const int v_plane = args.batched ? seq * args.heads + h : kvh;
    const int v_col0 = args.batched ? 0 : (args.vt_starts != nullptr ? args.vt_starts[seq] : row_base);

    if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) {
        printf("[mjv] attention: dynamic smem base not 1024-aligned\n");
        __trap();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        for (int s = 0; s < ST; ++s) {
            ptx::mbar_init(&k_full[s], 1);
            ptx::mbar_init(&v_full[s], 1);
            ptx::mbar_init(&kv_empty[s], 1);
        }
        ptx::mbar_init(&s_full[0], 1);
        ptx::mbar_init(&s_full[1], 1);
        ptx::mbar_init(p_full, 4);  // one arrival per softmax warp
        ptx::mbar_init(pv_done, 1);
        ptx::mbar_init(o_done, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, 256);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_O = tmem_base + 128;  // HD columns; S buffers at columns 0 and 64

    if (warp == 0) {
        // ===================================================== TMA producer
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(q_full, L::kQBytes);
#pragma unroll
            for (int kd = 0; kd < HD / 64; ++kd)
                ptx::tma_load_3d(sQ + kd * (BQ * 128), &tmQ, q_full, args.q_col0 + h * HD + kd * 64, tma_row0 + q0,
                                 tma_batch, ptx::kEvictFirst);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < n_blocks; ++j) {
                ptx::mbar_wait(&kv_empty[st], ph ^ 1);
                ptx::mbar_expect_tx(&k_full[st], L::kKBytes);
#pragma unroll
                for (int kd = 0; kd < HD / 64; ++kd)
                    ptx::tma_load_3d(sK + st * L::kKBytes + kd * (BKV * 128), &tmK, &k_full[st],
                                     args.k_col0 + kvh * HD + kd * 64, tma_row0 + j * BKV, tma_batch, ptx::kEvictLast);
                ptx::mbar_expect_tx(&v_full[st], L::kVBytes);
                ptx::tma_load_3d(sV + st * L::kVBytes, &tmV, &v_full[st], v_col0 + j * BKV, 0, v_plane,
                                 ptx::kEvictLast);
                if (++st == ST) {
                    st = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BQ, BKV);
            constexpr uint32_t idesc_o = ptx::make_idesc_bf16(BQ, HD);
            const uint32_t q_addr = ptx::smem_u32(sQ);
            auto issue_s = [&](int jj, int stage) {
                const uint32_t k_addr = ptx::smem_u32(sK + stage * L::kKBytes);
                const uint32_t d = tmem_base + (jj & 1) * BKV;
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks) {
                    const uint64_t a = ptx::make_desc_sw128(q_addr + (ks >> 2) * (BQ * 128) + (ks & 3) * 32);
                    const uint64_t b = ptx::make_desc_sw128(k_addr + (ks >> 2) * (BKV * 128) + (ks & 3) * 32);
                    ptx::mma_bf16_ss(d, a, b, idesc_s, ks != 0 ? 1u : 0u);
                }
                ptx::mma_commit(&s_full[jj & 1]);
            };
            ptx::mbar_wait(q_full, 0);
            ptx::mbar_wait(&k_full[0], 0);
            ptx::tc_fence_after();
            issue_s(0, 0);
            int st = 0;          // stage of block j
            uint32_t ph = 0;
            for (int j = 0; j < n_blocks; ++j) {
                int st1 = st + 1;
                uint32_t ph1 = ph;
                if (st1 == ST) {
                    st1 = 0;
                    ph1 ^= 1;
                }
                if (j + 1 < n_blocks) {
                    // S buffer (j+1)&1 was last read by softmax_{j-1}, which arrived on p_full_{j-1} (waited below)
                    ptx::mbar_wait(&k_full[st1], ph1);
                    ptx::tc_fence_after();
                    issue_s(j + 1, st1);
                }
                ptx::mbar_wait(p_full, j & 1);
                ptx::mbar_wait(&v_full[st], ph);
                ptx::tc_fence_after();
                const uint32_t v_addr = ptx::smem_u32(sV + st * L::kVBytes);
                const int valid = min(BKV, kv_end - j * BKV);
                const int ksteps = (valid + 15) >> 4;
                const uint32_t p_tmem = tmem_base + (j & 1) * BKV;  // P_j lives in the first 32 columns of S buffer j&1
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t b = ptx::make_desc_sw128(v_addr + ks * 32);
                    ptx::mma_bf16_ts(tmem_O, p_tmem + ks * 8, b, idesc_o, (j | ks) != 0 ? 1u : 0u);
                }
                ptx::mma_commit(&kv_empty[st]);
                ptx::mma_commit(pv_done);
                st = st1;
                ph = ph1;
            }
            // pv_done alternates phases every block, so a parity wait on it can alias when a warp races through its last
            // (masked) steps; the final hand-over gets its own single-phase barrier
            ptx::mma_commit(o_done);
        }
    } else if (warp >= 4) {
        // ===================================================== softmax / output (TMEM lane == query row)
        const int qw = warp & 3;
        const int r = qw * 32 + lane;
        const int qrow = q0 + r;                          // position inside the sequence
        const bool warp_has_rows = (q0 + qw * 32) < len;  // warp-uniform
        const uint32_t lane_sel = static_cast<uint32_t>(qw * 32) << 16;
        float m_run = -INFINITY;
        float l_run = 0.0f;

        for (int j = 0; j < n_blocks; ++j) {
            ptx::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            ptx::tc_fence_after();
            const uint32_t tmem_S = tmem_base + (j & 1) * BKV + lane_sel;
            const int kv0 = j * BKV;
            const int valid = min(BKV, kv_end - kv0);  // columns of this block that exist (block level)
            int lim = valid;                           // per-row column limit (exclusive, block-relative)
            if (CAUSAL) lim = min(valid, qrow - kv0 + 1);
            const bool warp_masked = CAUSAL && (kv0 > q0 + qw * 32 + 31);  // every row of this warp sees nothing here
            // every column of this block is visible to every row of this warp -> no per-element masking
            const bool full_block = (valid == BKV) && (!CAUSAL || (kv0 + BKV - 1 <= q0 + qw * 32));
            uint32_t v[64];  // scores of this row; v[0..31] is overwritten in place by the packed bf16 probabilities
            bool write_p = false;

            if (warp_has_rows && !warp_masked) {
                ptx::tmem_ld_x32(tmem_S, v);
                ptx::tmem_ld_x32(tmem_S + 32, v + 32);
                ptx::tmem_wait_ld();
                // ---- block row max
                float bm;
                if (full_block) {
                    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])));
                        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v[2 * i + 2]), __uint_as_float(v[2 * i + 3])));
                    }
                    bm = fmaxf(mx0, mx1) * args.scale_log2;  // scale > 0
                } else {
                    bm = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        bm = fmaxf(bm, (i < lim) ? __uint_as_float(v[i]) * args.scale_log2 : -INFINITY);
                }
                // ---- lazy rescale of the TMEM accumulator
                float alpha = 1.0f;
                bool need = false;
                if (j == 0) {
                    m_run = bm;
                } else if (bm > m_run + kRescaleThreshold) {
                    alpha = ptx::ex2_approx(m_run - bm);
                    m_run = bm;
                    need = true;
                }
                if (__any_sync(0xffffffffu, need)) {
                    ptx::mbar_wait(pv_done, (j - 1) & 1);  // P·V_{j-1} must have landed in O
                    ptx::tc_fence_after();
                    l_run *= alpha;
#pragma unroll 1
                    for (int c = 0; c < HD / 32; ++c) {
                        uint32_t o[32];
                        ptx::tmem_ld_x32(tmem_O + lane_sel + c * 32, o);
                        ptx::tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        ptx::tmem_st_x32(tmem_O + lane_sel + c * 32, o);
                    }
                    ptx::tmem_wait_st();
                }
                // ---- p = exp2(s*c - m), row sum, bf16 pack
                if (full_block) {
                    const uint64_t sc2 = ptx::pack_f32x2(args.scale_log2, args.scale_log2);
                    const uint64_t nm2 = ptx::pack_f32x2(-m_run, -m_run);
                    uint64_t ps2 = ptx::pack_f32x2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const uint64_t t2 = ptx::ffma2(
                            ptx::pack_f32x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
                        float t0, t1;
                        ptx::unpack_f32x2(t2, t0, t1);
                        const float p0 = ptx::ex2_approx(t0);
                        const float p1 = ptx::ex2_approx(t1);
                        ps2 = ptx::fadd2(ps2, ptx::pack_f32x2(p0, p1));
                        v[i] = ptx::pack_bf16x2(p0, p1);
                    }
                    float s0, s1;
                    ptx::unpack_f32x2(ps2, s0, s1);
                    l_run += s0 + s1;
                } else {
                    float psum = 0.0f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float p0 = ptx::ex2_approx(fmaf(__uint_as_float(v[2 * i]), args.scale_log2, -m_run));
                        float p1 = ptx::ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), args.scale_log2, -m_run));
                        p0 = (2 * i < lim) ? p0 : 0.0f;
                        p1 = (2 * i + 1 < lim) ? p1 : 0.0f;
                        psum += p0 + p1;
                        v[i] = ptx::pack_bf16x2(p0, p1);
                    }
                    l_run += psum;
                }
                write_p = true;
            } else if (warp_has_rows) {
                // rows exist but see no column of this block: P rows must be exact zeros
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0u;
                write_p = true;
            }
            // P_j (bf16 pairs, 32 columns) overwrites the scores it was computed from
            if (write_p) {
                ptx::tmem_st_x32(tmem_S, v);
                ptx::tmem_wait_st();
            }
            ptx::tc_fence_before();
            __syncwarp();
            // A warp with little to do (no valid rows / causally masked) must not lap the others: its arrival for step j
            // may only be counted once phase j-1 of p_full has completed. Checked here, after the work, so that busy
            // warps (for which the phase completed long ago) never stall on it and do not run in lock-step.
            if (lane == 0) {
                if (j > 0) ptx::mbar_wait(p_full, (j - 1) & 1);
                ptx::mbar_arrive(p_full);
            }
            __syncwarp();
        }

        // ---- epilogue: O / l -> bf16
        ptx::mbar_wait(o_done, 0);
        ptx::tc_fence_after();
        if (warp_has_rows) {
            const float inv = 1.0f / l_run;
            __nv_bfloat16* dst = args.out + static_cast<long long>(row_base + qrow) * args.ldo + h * HD;
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t o[32];
                ptx::tmem_ld_x32(tmem_O + lane_sel + c * 32, o);
                ptx::tmem_wait_ld();
                if (qrow < len) {
                    uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        d4[g] = make_uint4(
                            ptx::pack_bf16x2(__uint_as_float(o[8 * g + 0]) * inv, __uint_as_float(o[8 * g + 1]) * inv),
                            ptx::pack_bf16x2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv),
                            ptx::pack_bf16x2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv),
                            ptx::pack_bf16x2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv));
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace mjv
