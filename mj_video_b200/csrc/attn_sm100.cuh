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
//   MMA thread : O += P_j·V_j (A = P from TMEM, B = Vᵀ tile from smem, accumulator stays in TMEM)     -> v_empty
//                (tcgen05.mma executes in issue order, so S_{j+2} cannot overwrite P_j before P_j·V_j has read it)
// so the softmax of block j+1 starts as soon as block j's P is handed over, while P_j·V_j runs on the tensor pipe.
// The output accumulator never leaves TMEM until the end; the running max is updated lazily (only when it grows by more
// than 2^8), in which case O is rescaled in place with tcgen05.ld/st after waiting for the P·V in flight (the v_empty
// barrier of its V tile, which the same tcgen05.commit signals).
// PERSISTENT: the grid is 2 CTAs per SM; every CTA walks a static list of (sequence, head, query-block) items. The TMA and
// MMA roles run ahead into the next item (Q double-buffered for head_dim 64, K/V ring and S ring shared across items) and
// the output accumulator is handed over with o_done / o_empty, so barrier init, TMEM allocation, descriptor prefetch and
// the Q/K load latency are paid once per CTA instead of once per item.
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
    const int2* work;       // varlen: [(seq, q_block)] schedule (heavy blocks first); nullptr in batched mode
    int n_items;            // number of (sequence, head, q-block) items
    int n_qblocks;          // batched mode: q-blocks per sequence
    const int* vt_starts;   // varlen: first V^T column of each sequence (8-aligned) or nullptr (= cu_seqlens)
    int heads;              // query heads
    int kv_group;           // query heads per kv head
    int q_col0, k_col0;     // first q / k column inside their tensor maps
    __nv_bfloat16* out;     // [rows, ldo]
    int ldo;
    float scale_log2;       // softmax scale * log2(e)
    int batched;            // 1: maps carry a per-sequence batch dim and rows are sequence-relative (ViT)
    int out_hd;             // columns per head written to `out` (<= HD, multiple of 16): narrow heads are stored packed
    int pair_heads;         // attention2 (attn2_sm100.cuh), varlen: the two query tiles of an item are the heads 2p, 2p+1
    int pair_tail;          // attention2, batched, odd q-block count (ViT: 1025 = 8 x 128 + 1): the last q-blocks of heads 2p, 2p+1
                            // share an item (each tile with its OWN K/V stream) instead of running as two one-tile items
    int qk_stride;          // attention2: columns from one q / k head to the next in their tensors (= HD, or 96 for packed
                            // head_dim-96 heads under the 128-wide tile)
    int qk_ksteps;          // attention2: 16-column k-steps of Q.K^T that hold real columns (HD / 16, or 6 of 8 for head_dim 96)
    int pv_n;               // attention2: N of the P.V MMA = V^T rows that exist, rounded up to 16 (<= HD)
    int kv_tail;            // attention2, batched bidirectional head_dim 64, seq_len % 128 == 1 (ViT: 1025): the last key is folded
                            // in on the CUDA cores (softmax warps: q.k_last; epilogue: + p_last * v_last) instead of costing
                            // every item a ninth 128-column step with one valid column
    const __nv_bfloat16* k_ptr;    // raw K / V^T pointers for that path (same tensors the maps describe)
    long long ldk;
    const __nv_bfloat16* vt_ptr;
    long long vt_ld, vt_plane_stride;
    long long* trace;       // attention2, tools only: CTA 0 records (clock64, tag<<32|step) pairs per role here (or nullptr)
};

namespace attn_detail {
constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int kThreads = 256;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain: p stays below 2^8

template <int HD>
struct Smem {
    static constexpr int kKStages = HD == 64 ? 4 : 3;  // K ring (runs ahead with the S stream)
    static constexpr int kVStages = HD == 64 ? 4 : 2;  // V ring
    static constexpr int kSBufs = HD == 64 ? 3 : 2;    // S/P buffers in TMEM (3 x 64 + 64 = 2 x 64 + 128 = 256 columns)
    static constexpr int kQBufs = HD == 64 ? 2 : 1;    // Q tiles in smem
    static constexpr int kQBytes = BQ * HD * 2;
    static constexpr int kKBytes = BKV * HD * 2;
    static constexpr int kVBytes = HD * BKV * 2;
    static constexpr int kQOff = 0;
    static constexpr int kKOff = kQOff + kQBufs * kQBytes;
    static constexpr int kVOff = kKOff + kKStages * kKBytes;
    static constexpr int kBarOff = kVOff + kVStages * kVBytes;
    static constexpr int kTotal = kBarOff + 256;
};

struct Item {
    int seq, h, q0, len, row_base, kv_end, n_blocks, tma_row0, tma_batch, kvh, v_plane, v_col0;
};

template <bool CAUSAL>
__device__ __forceinline__ Item decode_item(const AttnArgs& args, int it) {
    Item t;
    int qblk;
    if (args.work != nullptr) {
        const int2 w = args.work[it / args.heads];
        t.seq = w.x;
        qblk = w.y;
        t.h = it % args.heads;
    } else {
        qblk = it % args.n_qblocks;
        const int sh = it / args.n_qblocks;
        t.h = sh % args.heads;
        t.seq = sh / args.heads;
    }
    if (args.cu_seqlens != nullptr) {
        t.row_base = args.cu_seqlens[t.seq];
        t.len = args.cu_seqlens[t.seq + 1] - t.row_base;
    } else {
        t.len = args.seq_len;
        t.row_base = t.seq * t.len;
    }
    t.q0 = qblk * BQ;
    t.kv_end = CAUSAL ? min(t.len, t.q0 + BQ) : t.len;
    t.n_blocks = (t.kv_end + BKV - 1) / BKV;
    t.tma_row0 = args.batched ? 0 : t.row_base;
    t.tma_batch = args.batched ? t.seq : 0;
    t.kvh = t.h / args.kv_group;
    t.v_plane = args.batched ? t.seq * args.heads + t.h : t.kvh;
    t.v_col0 = args.batched ? 0 : (args.vt_starts != nullptr ? args.vt_starts[t.seq] : t.row_base);
    return t;
}
}  // namespace attn_detail

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(attn_detail::kThreads, 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnArgs args) {
    using namespace attn_detail;
    using L = Smem<HD>;
    constexpr int KS = L::kKStages;
    constexpr int VS = L::kVStages;
    constexpr int NSB = L::kSBufs;
    constexpr int NQ = L::kQBufs;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem + L::kQOff;
    uint8_t* sK = smem + L::kKOff;
    uint8_t* sV = smem + L::kVOff;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* q_full = bars;                 // [2]
    uint64_t* q_empty = bars + 2;            // [2]
    uint64_t* k_full = bars + 4;             // [KS]
    uint64_t* k_empty = k_full + KS;         // [KS]  K tile consumed by its S MMA
    uint64_t* v_full = k_empty + KS;         // [VS]
    uint64_t* v_empty = v_full + VS;         // [VS]  V tile consumed by its P·V MMA
    uint64_t* s_full = v_empty + VS;         // [3]
    uint64_t* p_full = s_full + 3;           // [3], one per S/P buffer
    uint64_t* o_done = p_full + 3;           // accumulator of an item complete (MMA -> softmax)
    uint64_t* o_empty = o_done + 1;          // accumulator read out (softmax -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_items = args.n_items;

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
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&q_full[b], 1);
            ptx::mbar_init(&q_empty[b], 1);
        }
        for (int b = 0; b < 3; ++b) ptx::mbar_init(&p_full[b], 4);  // one arrival per softmax warp
        for (int s = 0; s < KS; ++s) {
            ptx::mbar_init(&k_full[s], 1);
            ptx::mbar_init(&k_empty[s], 1);
        }
        for (int s = 0; s < VS; ++s) {
            ptx::mbar_init(&v_full[s], 1);
            ptx::mbar_init(&v_empty[s], 1);
        }
        for (int b = 0; b < 3; ++b) ptx::mbar_init(&s_full[b], 1);
        ptx::mbar_init(o_done, 1);
        ptx::mbar_init(o_empty, 4);
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
    const uint32_t tmem_O = tmem_base + NSB * BKV;  // HD columns; S/P buffers at columns 0, 64(, 128)

    if (warp == 0) {
        // ===================================================== TMA producer
        if (ptx::elect_one()) {
            int kst = 0, vst = 0;
            uint32_t kph = 0, vph = 0;
            int qi = 0;  // items handled by this CTA so far
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++qi) {
                const Item t = decode_item<CAUSAL>(args, it);
                const int qb = qi % NQ;
                ptx::mbar_wait(&q_empty[qb], ((qi / NQ) & 1) ^ 1);
                ptx::mbar_expect_tx(&q_full[qb], L::kQBytes);
#pragma unroll
                for (int kd = 0; kd < HD / 64; ++kd)
                    ptx::tma_load_3d(sQ + qb * L::kQBytes + kd * (BQ * 128), &tmQ, &q_full[qb],
                                     args.q_col0 + t.h * HD + kd * 64, t.tma_row0 + t.q0, t.tma_batch, ptx::kEvictFirst);
                for (int j = 0; j < t.n_blocks; ++j) {
                    ptx::mbar_wait(&k_empty[kst], kph ^ 1);
                    ptx::mbar_expect_tx(&k_full[kst], L::kKBytes);
#pragma unroll
                    for (int kd = 0; kd < HD / 64; ++kd)
                        ptx::tma_load_3d(sK + kst * L::kKBytes + kd * (BKV * 128), &tmK, &k_full[kst],
                                         args.k_col0 + t.kvh * HD + kd * 64, t.tma_row0 + j * BKV, t.tma_batch,
                                         ptx::kEvictLast);
                    if (++kst == KS) {
                        kst = 0;
                        kph ^= 1;
                    }
                    ptx::mbar_wait(&v_empty[vst], vph ^ 1);
                    ptx::mbar_expect_tx(&v_full[vst], L::kVBytes);
                    ptx::tma_load_3d(sV + vst * L::kVBytes, &tmV, &v_full[vst], t.v_col0 + j * BKV, 0, t.v_plane,
                                     ptx::kEvictLast);
                    if (++vst == VS) {
                        vst = 0;
                        vph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BQ, BKV);
            constexpr uint32_t idesc_o = ptx::make_idesc_bf16(BQ, HD);
            // ---- S stream: runs ahead of the P·V stream by up to NSB-1 blocks, across item boundaries
            int s_it = blockIdx.x;   // item of the next S block
            int s_qi = 0;            // its local index
            int s_j = 0;             // block inside that item
            int s_nblk = s_it < n_items ? decode_item<CAUSAL>(args, s_it).n_blocks : 0;
            int s_st = 0;            // K/V ring slot of the next S block
            uint32_t s_ph = 0;
            int gs = 0;              // S blocks issued so far
            auto issue_next_s = [&]() {
                const int qb = s_qi % NQ;
                if (s_j == 0) ptx::mbar_wait(&q_full[qb], (s_qi / NQ) & 1);
                ptx::mbar_wait(&k_full[s_st], s_ph);
                ptx::tc_fence_after();
                const uint32_t q_addr = ptx::smem_u32(sQ + qb * L::kQBytes);
                const uint32_t k_addr = ptx::smem_u32(sK + s_st * L::kKBytes);
                const uint32_t d = tmem_base + (gs % NSB) * BKV;
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks) {
                    const uint64_t a = ptx::make_desc_sw128(q_addr + (ks >> 2) * (BQ * 128) + (ks & 3) * 32);
                    const uint64_t b = ptx::make_desc_sw128(k_addr + (ks >> 2) * (BKV * 128) + (ks & 3) * 32);
                    ptx::mma_bf16_ss(d, a, b, idesc_s, ks != 0 ? 1u : 0u);
                }
                ptx::mma_commit(&s_full[gs % NSB]);
                ptx::mma_commit(&k_empty[s_st]);  // K tile free once this S retires
                ++gs;
                if (++s_st == KS) {
                    s_st = 0;
                    s_ph ^= 1;
                }
                if (++s_j == s_nblk) {  // last S of the item: its Q tile can be replaced once these MMAs retire
                    ptx::mma_commit(&q_empty[qb]);
                    s_it += gridDim.x;
                    ++s_qi;
                    s_j = 0;
                    s_nblk = s_it < n_items ? decode_item<CAUSAL>(args, s_it).n_blocks : 0;
                }
            };
            // ---- P·V stream
            int st = 0;
            uint32_t ph = 0;
            int g = 0;   // steps consumed so far (global over items)
            int qi = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++qi) {
                const Item t = decode_item<CAUSAL>(args, it);
                for (int j = 0; j < t.n_blocks; ++j, ++g) {
                    // S_{g+NSB-1} reuses the buffer of P_{g-1}: P·V_{g-1} has been issued and tcgen05.mma executes in issue
                    // order; the softmax warps finished reading S_{g-1} before p_full_{g-1} completed
                    while (gs < g + NSB && s_it < n_items) issue_next_s();
                    ptx::mbar_wait(&p_full[g % NSB], (g / NSB) & 1);
                    ptx::mbar_wait(&v_full[st], ph);
                    if (j == 0 && qi > 0) ptx::mbar_wait(o_empty, (qi - 1) & 1);  // previous item's O has been read out
                    ptx::tc_fence_after();
                    const uint32_t v_addr = ptx::smem_u32(sV + st * L::kVBytes);
                    const int valid = min(BKV, t.kv_end - j * BKV);
                    const int ksteps = (valid + 15) >> 4;
                    const uint32_t p_tmem = tmem_base + (g % NSB) * BKV;  // P_g lives in the first 32 columns of its S buffer
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t b = ptx::make_desc_sw128(v_addr + ks * 32);
                        ptx::mma_bf16_ts(tmem_O, p_tmem + ks * 8, b, idesc_o, (j | ks) != 0 ? 1u : 0u);
                    }
                    ptx::mma_commit(&v_empty[st]);
                    if (++st == VS) {
                        st = 0;
                        ph ^= 1;
                    }
                }
                ptx::mma_commit(o_done);
            }
        }
    } else if (warp >= 4) {
        // ===================================================== softmax / output (TMEM lane == query row)
        const int qw = warp & 3;
        const int r = qw * 32 + lane;
        const uint32_t lane_sel = static_cast<uint32_t>(qw * 32) << 16;
        int g = 0;
        int qi = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++qi) {
            const Item t = decode_item<CAUSAL>(args, it);
            const int q0 = t.q0, len = t.len, kv_end = t.kv_end;
            const int qrow = q0 + r;                          // position inside the sequence
            const bool warp_has_rows = (q0 + qw * 32) < len;  // warp-uniform
            float m_run = -INFINITY;
            float l_run = 0.0f;

            for (int j = 0; j < t.n_blocks; ++j, ++g) {
                ptx::mbar_wait(&s_full[g % NSB], (g / NSB) & 1);
                ptx::tc_fence_after();
                const uint32_t tmem_S = tmem_base + (g % NSB) * BKV + lane_sel;
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
                        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;  // 4 independent chains
#pragma unroll
                        for (int i = 0; i < 64; i += 8) {
                            mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
                            mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
                            mx2 = fmaxf(mx2, fmaxf(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])));
                            mx3 = fmaxf(mx3, fmaxf(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
                        }
                        bm = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * args.scale_log2;  // scale > 0
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
                        // P·V_{g-1} must have landed in O: its commit completes the v_empty barrier of the V tile it read
                        // (ring slot (g-1) % VS, phase (g-1) / VS). That slot's NEXT completion belongs to P·V_{g-1+VS}, which
                        // cannot be issued before this warp hands over P_g (VS >= 2), so the parity wait cannot alias.
                        ptx::mbar_wait(&v_empty[(g - 1) % VS], ((g - 1) / VS) & 1);
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
                        uint64_t ps2[4];  // independent partial sums: no 32-deep dependent add chain
#pragma unroll
                        for (int a = 0; a < 4; ++a) ps2[a] = ptx::pack_f32x2(0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint64_t t2 = ptx::ffma2(
                                ptx::pack_f32x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
                            float t0, t1;
                            ptx::unpack_f32x2(t2, t0, t1);
                            const float p0 = ptx::ex2_approx(t0);
                            const float p1 = ptx::ex2_approx(t1);
                            ps2[i & 3] = ptx::fadd2(ps2[i & 3], ptx::pack_f32x2(p0, p1));
                            v[i] = ptx::pack_bf16x2(p0, p1);
                        }
                        float s0, s1, s2, s3;
                        ptx::unpack_f32x2(ptx::fadd2(ps2[0], ps2[1]), s0, s1);
                        ptx::unpack_f32x2(ptx::fadd2(ps2[2], ps2[3]), s2, s3);
                        l_run += (s0 + s1) + (s2 + s3);
                    } else {
                        // columns at or beyond wlim are invisible to every row of this warp (block tail / above the causal
                        // diagonal): the branch is warp-uniform, so those pairs cost no exponentials
                        const int wlim = CAUSAL ? min(valid, q0 + qw * 32 + 32 - kv0) : valid;
                        float psum = 0.0f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float p0 = 0.0f, p1 = 0.0f;
                            if (2 * i < wlim) {
                                p0 = ptx::ex2_approx(fmaf(__uint_as_float(v[2 * i]), args.scale_log2, -m_run));
                                p1 = ptx::ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), args.scale_log2, -m_run));
                                p0 = (2 * i < lim) ? p0 : 0.0f;
                                p1 = (2 * i + 1 < lim) ? p1 : 0.0f;
                                psum += p0 + p1;
                            }
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
                // P_g (bf16 pairs, 32 columns) overwrites the scores it was computed from
                if (write_p) {
                    ptx::tmem_st_x32(tmem_S, v);
                    ptx::tmem_wait_st();
                }
                ptx::tc_fence_before();
                __syncwarp();
                // One p_full barrier per S/P buffer: a warp with little to do (no valid rows / causally masked) can run at
                // most NSB-1 steps ahead of the slowest warp (S_{g+NSB-1} is only issued after p_full_{g-1} completed), so its
                // early arrivals land on the other barriers and can never complete a phase prematurely.
                if (lane == 0) ptx::mbar_arrive(&p_full[g % NSB]);
            }

            // ---- item epilogue: O / l -> bf16
            ptx::mbar_wait(o_done, qi & 1);
            ptx::tc_fence_after();
            if (warp_has_rows) {
                const float inv = 1.0f / l_run;
                __nv_bfloat16* dst = args.out + static_cast<long long>(t.row_base + qrow) * args.ldo + t.h * args.out_hd;
#pragma unroll 1
                for (int c = 0; c < HD / 32; ++c) {
                    if (c * 32 >= args.out_hd) break;  // columns beyond the model's head_dim are exact zeros: not stored
                    uint32_t o[32];
                    ptx::tmem_ld_x32(tmem_O + lane_sel + c * 32, o);
                    ptx::tmem_wait_ld();
                    if (qrow < len) {
                        uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {
                            d4[gq] = make_uint4(
                                ptx::pack_bf16x2(__uint_as_float(o[8 * gq + 0]) * inv, __uint_as_float(o[8 * gq + 1]) * inv),
                                ptx::pack_bf16x2(__uint_as_float(o[8 * gq + 2]) * inv, __uint_as_float(o[8 * gq + 3]) * inv),
                                ptx::pack_bf16x2(__uint_as_float(o[8 * gq + 4]) * inv, __uint_as_float(o[8 * gq + 5]) * inv),
                                ptx::pack_bf16x2(__uint_as_float(o[8 * gq + 6]) * inv, __uint_as_float(o[8 * gq + 7]) * inv));
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(o_empty);  // the accumulator may be overwritten by the next item
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
