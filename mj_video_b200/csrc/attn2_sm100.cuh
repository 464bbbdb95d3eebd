// Fused softmax attention, second generation: TWO 128-row query tiles per CTA, 128-column KV blocks, one CTA per SM.
//
// Replaces K6  flash_attn_varlen_qkvpacked_func(causal=False)  modeling_intern_vit.py:72-75 / _naive_attn :210-227
//                        K15 flash_attn_func / flash_attn_varlen_func(causal=True) modeling_internlm2.py:498-561 / eager :393-410
//
// Why a second kernel (profiles/r01_ncu_full_attn_vit_v4.csv, source page): in attn_sm100.cuh a softmax warp spent only
// 38 % of its time in the exponential loop; the rest was per-64-column-step overhead (barrier waits, spill reloads under
// the 128-register cap of two CTAs per SM, instruction-cache misses of the doubly unrolled masked path, the item epilogue).
// Here every step covers 128 columns, the two softmax warpgroups own 192 registers per thread (setmaxnreg; epilogue 72,
// producer / issuers 56), the masked
// path is a small pre-pass in front of the one exponential loop, and the O / l -> bf16 epilogue runs on its own
// warpgroup while the softmax warps already work on the next item.
//
// Roles (512 threads, 16 warps; TMEM lane quarter of a warp = warp % 4):
//   warps 0-3   softmax of query tile 0 (thread = one query row, 128 scores per step)
//   warps 4-7   softmax of query tile 1
//   warps 8-11  epilogue: O_t / l -> bf16 -> global, then hands the accumulator back (o_empty)
//   warp  12    TMA producer (Q tiles, K ring, V^T ring)
//   warp  13    P·V issuer (O_t += P·V with P read from TMEM)
//   warp  14    S issuer (S = Q·K^T into a ring of kNSB score buffers)
// The two query tiles of an item share every K / V^T tile (one smem read stream for both):
//   * bidirectional batched mode (ViT): consecutive 128-row blocks of one (sequence, head);
//   * causal varlen mode with an even GQA group (InternLM2, 16 q / 8 kv heads): the two query heads of one kv head at the
//     same query block - identical masks, shared K/V;
//   * causal varlen mode otherwise (MHA, e.g. the Llama / Phi-3 shaped branch): two query blocks of one head, packed by the
//     host in one work entry (low 16 bits: first block, high 16 bits: second block + 1, or 0 for none).
// One exception shares nothing: with an odd number of query blocks and an even head count (ViT: 1025 rows = 8 blocks + 1
// row) the last blocks of heads 2p and 2p + 1 form a "dual" item whose tiles read their own K / V (a stage per step).
// Sequence lengths of 128 n + 1 in that mode also fold the last key in on the CUDA cores (AttnArgs.kv_tail) instead of
// paying a block with one valid column. head_dim 128 keeps P in shared memory (one tile per query tile, SS-form P.V) and
// hands the score buffer back as soon as it has been read (s_free); head_dim 64 writes P over S in TMEM (TS-form P.V).
// Heads packed 96 columns apart run under the 128-wide tile with 6 of its 8 k-steps (AttnArgs.qk_stride / qk_ksteps).
// Steps are numbered globally (g) in the order (block j, tile t) over the active tiles; step g uses score buffer
// g % kNSB; S_{g+kNSB} (same buffer) is issued once pv_done signals that P_g·V has read P_g. The running maximum is updated lazily (only when it grows by more than 2^8), the rare rescale of
// O_t happens in the softmax warps after waiting for the previous P·V of that tile (pv_done).
#pragma once
#include "attn_sm100.cuh"

#ifndef MJV_TRACE_EPI
#define MJV_TRACE_EPI 0   // tools: record the epilogue (tags 20-23) instead of the S issuer in trace role 3
#endif
#ifndef MJV_ATTN_PSMEM
#define MJV_ATTN_PSMEM 1
#endif
namespace mjv {
namespace attn2 {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int kThreads = 512;
constexpr float kRescaleThreshold = 8.0f;  // log2 domain: p stays below 2^8

constexpr int kTraceEvents = 2048;  // per role (tools/attn_trace.py)

template <int HD>
struct Smem {
    // score buffers in TMEM: head_dim 64 leaves room for three (3 x 128 + 2 x 64 = 512 columns), so S of a tile's NEXT block
    // is computed while its softmax still works on the current one; head_dim 128 fills TMEM with two (2 x 128 + 2 x 128)
    static constexpr bool kPSmem = HD == 128 && MJV_ATTN_PSMEM;   // P through shared memory (see below)
    static constexpr int kNSB = HD == 64 ? 3 : 2;
    static constexpr int kQSets = HD == 64 ? 2 : 1;    // item-level double buffering of the Q tile pair
    static constexpr int kQBufs = 2 * kQSets;
    static constexpr int kKStages = HD == 64 ? 3 : (kPSmem ? 1 : 2);
    static constexpr int kVStages = HD == 64 ? 3 : 2;
    static constexpr int kQBytes = BQ * HD * 2;
    static constexpr int kKBytes = BKV * HD * 2;
    static constexpr int kVBytes = HD * BKV * 2;
    // head_dim 128: TMEM is full (2 x 128 S + 2 x 128 O), so with P written over S the next S of a tile could only be issued
    // after P·V of the current one (each warpgroup waited ~1400 of its ~3800-cycle step). There P goes through shared memory
    // instead (one tile per query tile, SS-form P·V) and the score buffer is handed back as soon as the row is in registers.
    // The 64 KB come from K: one stage instead of two - S is now issued a whole step ahead, which hides the K load.
    static constexpr int kPBytes = kPSmem ? 2 * BQ * BKV * 2 : 0;   // one P tile per query tile
    static constexpr int kQOff = 0;
    static constexpr int kKOff = kQOff + kQBufs * kQBytes;
    static constexpr int kVOff = kKOff + kKStages * kKBytes;
    static constexpr int kPOff = kVOff + kVStages * kVBytes;
    static constexpr int kLOff = kPOff + kPBytes;                // row sums handed to the epilogue: [2][128] fp32
    // trailing-key path (head_dim 64 only): p of the last key per row [2][128] fp32, its V row per tile [2][64] fp32, and
    // the key's scaled score per row [2][128] (parked in shared memory between the first step and the end of the item:
    // one more live register in the softmax loop cost 370 cycles per step)
    static constexpr int kTailOff = kLOff + 2 * BQ * 4;
    static constexpr int kTailBytes = HD == 64 ? 2 * BQ * 4 + 2 * HD * 4 + 2 * BQ * 4 : 0;   // + the key's score per row
    static constexpr int kBarOff = kTailOff + kTailBytes;
    static constexpr int kTotal = kBarOff + 512;
};

// two-element fields are only ever indexed with compile-time constants (a runtime index would put them in local memory);
// sel2() picks by a runtime tile number
struct Item {
    int row_base, len, tma_row0, tma_batch, kvh, v_plane, v_col0, nblk_max;
    int dual;   // the two tiles read different K/V (heads h, h+1): every step has its own K and V stage
    int h[2], q0[2], kv_end[2], nblk[2];
};
__device__ __forceinline__ int sel2(int t, const int* a) { return t != 0 ? a[1] : a[0]; }
__device__ __forceinline__ uint32_t sel2(int t, const uint32_t* a) { return t != 0 ? a[1] : a[0]; }

template <bool CAUSAL>
__device__ __forceinline__ Item decode_item(const AttnArgs& a, int it) {
    Item t;
    t.dual = 0;
    int seq, qb0, qb1;
    bool has1;
    if (a.work != nullptr) {
        if (a.pair_heads) {
            const int per = a.heads >> 1;
            const int2 w = a.work[it / per];
            const int hp = it % per;
            seq = w.x;
            qb0 = qb1 = w.y & 0xffff;
            has1 = true;
            t.h[0] = 2 * hp;
            t.h[1] = 2 * hp + 1;
        } else {
            const int2 w = a.work[it / a.heads];
            seq = w.x;
            qb0 = w.y & 0xffff;
            const int hi = (w.y >> 16) & 0xffff;
            has1 = hi != 0;
            qb1 = hi - 1;
            t.h[0] = t.h[1] = it % a.heads;
        }
    } else if (a.pair_tail) {
        // per sequence: heads x (n_qblocks / 2) two-block items, then heads / 2 items pairing the odd last q-block of two heads
        const int np = a.n_qblocks >> 1;
        const int full = a.heads * np;
        const int per_seq = full + (a.heads >> 1);
        seq = it / per_seq;
        const int r = it - seq * per_seq;
        if (r < full) {
            t.h[0] = t.h[1] = r / np;
            qb0 = 2 * (r % np);
            qb1 = qb0 + 1;
        } else {
            t.h[0] = 2 * (r - full);
            t.h[1] = t.h[0] + 1;
            qb0 = qb1 = a.n_qblocks - 1;
            t.dual = 1;
        }
        has1 = true;
    } else {
        const int npairs = (a.n_qblocks + 1) >> 1;
        const int pr = it % npairs;
        const int sh = it / npairs;
        t.h[0] = t.h[1] = sh % a.heads;
        seq = sh / a.heads;
        qb0 = 2 * pr;
        qb1 = qb0 + 1;
        has1 = qb1 < a.n_qblocks;
    }
    if (a.cu_seqlens != nullptr) {
        t.row_base = a.cu_seqlens[seq];
        t.len = a.cu_seqlens[seq + 1] - t.row_base;
    } else {
        t.len = a.seq_len;
        t.row_base = seq * t.len;
    }
    t.q0[0] = qb0 * BQ;
    t.q0[1] = qb1 * BQ;
    t.kv_end[0] = CAUSAL ? min(t.len, t.q0[0] + BQ) : t.len - a.kv_tail;
    t.kv_end[1] = CAUSAL ? min(t.len, t.q0[1] + BQ) : t.len - a.kv_tail;
    t.nblk[0] = (t.kv_end[0] + BKV - 1) / BKV;
    t.nblk[1] = has1 ? (t.kv_end[1] + BKV - 1) / BKV : 0;
    t.nblk_max = max(t.nblk[0], t.nblk[1]);
    t.tma_row0 = a.batched ? 0 : t.row_base;
    t.tma_batch = a.batched ? seq : 0;
    t.kvh = t.h[0] / a.kv_group;
    t.v_plane = a.batched ? seq * a.heads + t.h[0] : t.kvh;
    t.v_col0 = a.batched ? 0 : (a.vt_starts != nullptr ? a.vt_starts[seq] : t.row_base);
    return t;
}

// what the MMA issuer needs of an item (it tracks two items at once: the S stream runs ahead of the P·V stream)
struct ItemLite {
    int nblk[2], kv_end[2], nblk_max, dual;
};
template <bool CAUSAL>
__device__ __forceinline__ ItemLite decode_lite(const AttnArgs& a, int it) {
    const Item t = decode_item<CAUSAL>(a, it);
    ItemLite l;
    l.nblk[0] = t.nblk[0];
    l.nblk[1] = t.nblk[1];
    l.kv_end[0] = t.kv_end[0];
    l.kv_end[1] = t.kv_end[1];
    l.nblk_max = t.nblk_max;
    l.dual = t.dual;
    return l;
}

// ---- softmax building blocks (C = 32-column chunk index, compile time; v = the 128 scores of this thread's row)
template <int C>
__device__ __forceinline__ void mask_chunk(uint32_t* v, int lim) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
        if (C * 32 + i >= lim) v[C * 32 + i] = 0xff800000u;  // -inf
}
template <int C>
__device__ __forceinline__ void max_chunk(const uint32_t* v, float* mx) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        mx[0] = fmaxf(mx[0], fmaxf(__uint_as_float(v[C * 32 + i]), __uint_as_float(v[C * 32 + i + 1])));
        mx[1] = fmaxf(mx[1], fmaxf(__uint_as_float(v[C * 32 + i + 2]), __uint_as_float(v[C * 32 + i + 3])));
        mx[2] = fmaxf(mx[2], fmaxf(__uint_as_float(v[C * 32 + i + 4]), __uint_as_float(v[C * 32 + i + 5])));
        mx[3] = fmaxf(mx[3], fmaxf(__uint_as_float(v[C * 32 + i + 6]), __uint_as_float(v[C * 32 + i + 7])));
    }
}
// p = exp2(s*c - m) for the 32 scores of chunk C; packed bf16 pairs go to pk[16]; row-sum partials in ps2[4].
// Software-pipelined by hand: the consumers (row sum, bf16 pack) of pair i are placed kExpDepth pairs behind its two
// MUFU.EX2. ptxas otherwise consumes each pair right after the next pair's MUFUs are issued (16 cycles later), which is
// shorter than the MUFU latency: a warp alone on its scheduler then stalls 8 cycles per pair (24 instead of 16 cycles).
constexpr int kExpDepth = 3;

// 2^x for a packed pair WITHOUT the MUFU (the FlashAttention-4 trick): MUFU.EX2 issues one warp instruction per 8 cycles
// per scheduler, and at head_dim 64 the exponentials, not the tensor cores, bound the kernel; the FMA pipe has slack.
// Cody-Waite: n = round(x) through the 1.5*2^23 magic add (n lands in the low mantissa bits), f = x - n in [-0.5, 0.5],
// 2^f by a degree-3 minimax polynomial (max relative error 7.5e-5, far below the bf16 rounding of P), then n is added
// to the exponent field. Inputs are clamped at -125 (2^-125 ~ 0); masked scores (-inf) never take this path.
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1) {
    float x0, x1;
    ptx::unpack_f32x2(x2, x0, x1);
    const uint64_t xc = ptx::pack_f32x2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
    const uint64_t t = ptx::fadd2(xc, ptx::pack_f32x2(12582912.0f, 12582912.0f));
    const uint64_t n = ptx::fadd2(t, ptx::pack_f32x2(-12582912.0f, -12582912.0f));
    const uint64_t f = ptx::ffma2(n, ptx::pack_f32x2(-1.0f, -1.0f), xc);
    uint64_t p = ptx::ffma2(f, ptx::pack_f32x2(0.05517132f, 0.05517132f), ptx::pack_f32x2(0.24261054f, 0.24261054f));
    p = ptx::ffma2(p, f, ptx::pack_f32x2(0.69326099f, 0.69326099f));
    p = ptx::ffma2(p, f, ptx::pack_f32x2(0.99992811f, 0.99992811f));
    float pa, pb, ta, tb;
    ptx::unpack_f32x2(p, pa, pb);
    ptx::unpack_f32x2(t, ta, tb);
    p0 = __int_as_float(__float_as_int(pa) + (__float_as_int(ta) << 23));
    p1 = __int_as_float(__float_as_int(pb) + (__float_as_int(tb) << 23));
}

// POLY of every 8 pairs take the polynomial, the others the MUFU
template <int C, int POLY = 0>
__device__ __forceinline__ void exp_chunk(const uint32_t* v, uint32_t* pk, uint64_t sc2, uint64_t nm2, uint64_t* ps2) {
    float p[32];
#pragma unroll
    for (int i = 0; i < 16 + kExpDepth; ++i) {
        if (i < 16) {
            const uint64_t t2 = ptx::ffma2(
                ptx::pack_f32x2(__uint_as_float(v[C * 32 + 2 * i]), __uint_as_float(v[C * 32 + 2 * i + 1])), sc2, nm2);
            if ((i & 7) < POLY) {
                exp2_poly2(t2, p[2 * i], p[2 * i + 1]);
            } else {
                float t0, t1;
                ptx::unpack_f32x2(t2, t0, t1);
                p[2 * i] = ptx::ex2_approx(t0);
                p[2 * i + 1] = ptx::ex2_approx(t1);
            }
        }
        if (i >= kExpDepth) {
            const int k = i - kExpDepth;
            ps2[k & 3] = ptx::fadd2(ps2[k & 3], ptx::pack_f32x2(p[2 * k], p[2 * k + 1]));
            pk[k] = ptx::pack_bf16x2(p[2 * k], p[2 * k + 1]);
        }
    }
}

// 32 keys of one row (exp_chunk's 16 bf16 pairs = four 16-byte chunks) into the shared P tile (SW128, K-major)
template <int C>
__device__ __forceinline__ void st_p_smem(uint32_t prow, uint32_t rx, const uint32_t* pk) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t c = (C & 1) * 4 + q;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (C >> 1) * 16384 + ((c ^ rx) << 4)),
                     "r"(pk[4 * q]), "r"(pk[4 * q + 1]), "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                     : "memory");
    }
}

// tools only: event k of a role -> trace[role*2*kTraceEvents + 2k] = clock64, [.. + 1] = tag << 32 | step
__device__ __forceinline__ void trace_ev(long long* trace, int role, int& n, int tag, int step) {
    if (n < kTraceEvents) {
        long long* p = trace + (static_cast<long long>(role) * kTraceEvents + n) * 2;
        p[0] = clock64();
        p[1] = (static_cast<long long>(tag) << 32) | static_cast<unsigned>(step);
        ++n;
    }
}

}  // namespace attn2

// DBG (tools only, selected by the launcher when MJV_ATTN_TRACE is set): every barrier wait carries a watchdog that records
// which wait wedged (role/barrier code) into the host-pinned trace buffer before trapping.
// POLY: of every 8 pairs of exponentials in a fully visible block, POLY are evaluated on the FMA pipe (exp2_poly2).
template <int HD, bool CAUSAL, bool DBG = false, int POLY = 0>
__global__ void __launch_bounds__(attn2::kThreads, 1)
attention2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnArgs args) {
    using namespace attn2;
    using L = Smem<HD>;
    constexpr int KS = L::kKStages;
    constexpr int VS = L::kVStages;
    constexpr int QSETS = L::kQSets;
    constexpr int kNSB = L::kNSB;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem + L::kQOff;
    uint8_t* sK = smem + L::kKOff;
    uint8_t* sV = smem + L::kVOff;
    float* sL = reinterpret_cast<float*>(smem + L::kLOff);
    constexpr bool KVT = !CAUSAL && HD == 64;   // the trailing-key path exists in this instantiation (args.kv_tail turns it on)
    [[maybe_unused]] float* sPT = reinterpret_cast<float*>(smem + L::kTailOff);        // [2][BQ]
    [[maybe_unused]] float* sVT = reinterpret_cast<float*>(smem + L::kTailOff) + 2 * BQ;  // [2][HD]
    [[maybe_unused]] float* sMT = reinterpret_cast<float*>(smem + L::kTailOff) + 2 * BQ + 2 * HD;  // [2][BQ]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* q_full = bars;                 // [4]
    uint64_t* q_empty = bars + 4;            // [4]  Q tile consumed by the last S MMA of its item
    uint64_t* k_full = bars + 8;             // [KS]
    uint64_t* k_empty = k_full + KS;         // [KS] K tile consumed by the S MMAs of both query tiles
    uint64_t* v_full = k_empty + KS;         // [VS]
    uint64_t* v_empty = v_full + VS;         // [VS] V tile consumed by the P·V MMAs of both query tiles
    // s_full is kept PER QUERY TILE (slot = that tile's step count % kNSB), not per score buffer: an mbarrier wait only knows
    // a phase parity, so a waiter must see every completion of the barrier it waits on. With per-buffer barriers the
    // warpgroup of a tile that sits out an item (single-tile item) had to watch foreign steps without gating them, and a
    // slow warp was lapped by two phases near the end of a launch (wedged pipeline). Per tile, every completion is one of
    // the waiter's own steps, and slot reuse (kNSB tile steps later) is gated by its own P of the earlier step.
    uint64_t* s_full = v_empty + VS;         // [2][kNSB] scores ready (S issuer -> softmax of that tile)
    uint64_t* p_full = s_full + 2 * kNSB;    // [kNSB] probabilities written (softmax -> P·V issuer), 4 warp arrivals
    uint64_t* pv_done = p_full + kNSB;       // [kNSB] P·V of the step that used buffer b (only waited on by the rare rescale)
    uint64_t* o_done = pv_done + kNSB;       // [2] accumulator of tile t complete (MMA -> epilogue)
    uint64_t* o_empty = o_done + 2;          // [2] accumulator read out (epilogue -> MMA / softmax), 4 warp arrivals
    uint64_t* l_ready = o_empty + 2;         // [2] row sums in smem (softmax -> epilogue), 4 warp arrivals
    // P-in-smem mode (head_dim 128) only:
    uint64_t* s_free = l_ready + 2;          // [kNSB] scores of that buffer are in registers (softmax -> S issuer), 4 warp arrivals
    uint64_t* pvd = s_free + kNSB;           // [2] every P·V of tile t: O_t updated and the tile's P buffer read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pvd + 2);
    constexpr bool PSMEM = L::kPSmem;
    uint8_t* sP = smem + L::kPOff;

    // shuffled from lane 0 so that the compiler knows it is warp-uniform (role indices, TMEM addresses -> uniform registers)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int n_items = args.n_items;
    // wait_wd: waits that always carry the watchdog; wait_sp: bare spins (watchdog + record only in the DBG build)
    long long* const dbg_rec = DBG ? args.trace + 4 * 2 * kTraceEvents + 8 * (threadIdx.x >> 5) : nullptr;
    long long dbg_aux = 0;  // DBG: item << 32 | step counter of the role, reported when a wait wedges
    auto wait_wd = [&](uint64_t* bar, uint32_t parity, int code) {
        if constexpr (DBG) ptx::mbar_wait_dbg(bar, parity, dbg_rec, code, dbg_aux);
        else ptx::mbar_wait_nocall(bar, parity);
    };
    auto wait_sp = [&](uint64_t* bar, uint32_t parity, int code) {
        if constexpr (DBG) ptx::mbar_wait_dbg(bar, parity, dbg_rec, code, dbg_aux);
        else ptx::mbar_wait_spin(bar, parity);
    };

    if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need a 1024-byte aligned base
    if (warp == 12 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
    }
    if (warp == 13 && lane == 0) {
        for (int b = 0; b < 4; ++b) {
            ptx::mbar_init(&q_full[b], 1);
            ptx::mbar_init(&q_empty[b], 1);
        }
        for (int s = 0; s < KS; ++s) {
            ptx::mbar_init(&k_full[s], 1);
            ptx::mbar_init(&k_empty[s], 1);
        }
        for (int s = 0; s < VS; ++s) {
            ptx::mbar_init(&v_full[s], 1);
            ptx::mbar_init(&v_empty[s], 1);
        }
        for (int b = 0; b < kNSB; ++b) {
            ptx::mbar_init(&s_full[b], 1);
            ptx::mbar_init(&s_full[kNSB + b], 1);
            ptx::mbar_init(&p_full[b], 4);
            ptx::mbar_init(&pv_done[b], 1);
            ptx::mbar_init(&s_free[b], 4);
        }
        for (int t = 0; t < 2; ++t) {
            ptx::mbar_init(&o_done[t], 1);
            ptx::mbar_init(&pvd[t], 1);
            ptx::mbar_init(&o_empty[t], 4);
            ptx::mbar_init(&l_ready[t], 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 15) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_O0 = tmem_base + kNSB * BKV;  // O_t at columns kNSB*128 + t*HD

    const int wg = warp >> 2;
    if (wg < 2) {
        // ===================================================== softmax of query tile `wg` (TMEM lane == query row)
        ptx::setmaxnreg_inc<192>();
        const int tt = wg;
        const int qw = warp & 3;
        const int r = qw * 32 + lane;
        const uint32_t lane_sel = static_cast<uint32_t>(qw * 32) << 16;
        const uint32_t tmem_O = tmem_O0 + tt * HD + lane_sel;
        int sb = 0;           // score buffer of the next step (ring over ALL active steps of both tiles)
        uint32_t sph = 0;     // ... and the parity of that ring pass (p_full / pv_done phases)
        int my_slot = 0;      // s_full slot of this tile's next step (ring over its OWN steps)
        uint32_t my_ph = 0;
        uint32_t n_mine = 0;  // steps of this tile so far (tracing only)
        int prev_sb = 0;      // ring position of this tile's previous step: its P·V signals pv_done[prev_sb] in phase prev_sph
        uint32_t prev_sph = 0;
        uint32_t n_use = 0;   // items in which this tile was active so far
        const float scale_log2 = args.scale_log2;
        const bool tracing = DBG && args.trace != nullptr && blockIdx.x == 0 && qw == 0 && lane == 0;
        int n_tr = 0;

        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const Item t = decode_item<CAUSAL>(args, it);
            const int q0 = sel2(tt, t.q0), len = t.len, kv_end = sel2(tt, t.kv_end);
            const int my_nblk = sel2(tt, t.nblk);
            const int other_nblk = sel2(tt ^ 1, t.nblk);
            const int qrow = q0 + r;
            const bool warp_has_rows = (q0 + qw * 32) < len;  // warp-uniform
            float m_run = -INFINITY;
            float l_run = 0.0f;
            if constexpr (KVT) {
                float m_tail = -INFINITY;   // score of the trailing key, scaled to the log2 domain
                if (args.kv_tail != 0 && my_nblk > 0) {
                    // s_last = q_row . k_last on the CUDA cores: the row from the swizzled Q tile (TMA has landed: q_full), the
                    // key row straight from global (one address per warp). Products of bf16 pairs are exact in fp32, as on
                    // the tensor core. The tile cannot be reloaded before this read: an item has more key blocks than there
                    // are score buffers, so its last S needs a P of this item first.
                    const int buf = (n_use % QSETS) * 2 + tt;
                    wait_wd(&q_full[buf], (n_use / QSETS) & 1, 7);
                    if (warp_has_rows) {
                        const uint32_t qaddr = ptx::smem_u32(sQ) + buf * L::kQBytes + (r >> 3) * 1024 + (r & 7) * 128;
                        const uint4* kt = reinterpret_cast<const uint4*>(
                            args.k_ptr + static_cast<long long>(t.row_base + len - 1) * args.ldk + args.k_col0 +
                            (t.kvh + (t.dual ? tt : 0)) * args.qk_stride);
                        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            uint32_t q4[4];
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(q4[0]), "=r"(q4[1]), "=r"(q4[2]), "=r"(q4[3])
                                         : "r"(qaddr + ((c ^ (r & 7)) << 4)));
                            const uint4 k4 = __ldg(kt + c);
                            const uint32_t kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                acc0 = fmaf(__uint_as_float(q4[i] << 16), __uint_as_float(kk[i] << 16), acc0);
                                acc1 = fmaf(__uint_as_float(q4[i] & 0xffff0000u), __uint_as_float(kk[i] & 0xffff0000u), acc1);
                            }
                        }
                        m_tail = (acc0 + acc1) * scale_log2;
                    }
                }
                sMT[tt * BQ + r] = m_tail;   // private to this thread: no synchronisation
            }

            for (int j = 0; j < t.nblk_max; ++j) {
                if constexpr (DBG) dbg_aux = (static_cast<long long>(it) << 32) | (static_cast<long long>(j) << 16) | (sb << 8) | sph;
                if (tt == 1 && j < other_nblk) {  // tile 0's step of this block comes first in the ring
                    if (++sb == kNSB) { sb = 0; sph ^= 1; }
                }
                if (j < my_nblk) {
                    if (tracing) trace_ev(args.trace, tt, n_tr, 0, (int)n_mine);
                    wait_wd(&s_full[tt * kNSB + my_slot], my_ph, 1);
                    ptx::tc_fence_after();
                    if (++my_slot == kNSB) { my_slot = 0; my_ph ^= 1; }
                    if (tracing) trace_ev(args.trace, tt, n_tr, 1, (int)n_mine);
                    const uint32_t tmem_S = tmem_base + sb * BKV + lane_sel;
                    // P-in-smem mode, once the scores are in registers (every warp of the tile, rows or not): hand the score
                    // buffer back to the S issuer and wait for this tile's previous P·V - O_t may then be rescaled, the
                    // tile's P buffer overwritten, and p_full[tt] (one phase per step of the tile) arrived on again.
                    auto psmem_handoff = [&]() {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&s_free[sb]);
                        if (n_mine > 0) wait_wd(&pvd[tt], (n_mine - 1) & 1, 2);
                        ptx::tc_fence_after();
                    };
                    const int kv0 = j * BKV;
                    const int valid = min(BKV, kv_end - kv0);                    // columns of this block that exist
                    const int lim = CAUSAL ? min(valid, qrow - kv0 + 1) : valid;  // per-row visible columns
                    // columns at or beyond wlim are invisible to every row of this warp (warp-uniform)
                    const int wlim = CAUSAL ? min(valid, q0 + qw * 32 + 32 - kv0) : valid;
                    const bool full = (valid == BKV) && (!CAUSAL || (kv0 + BKV - 1 <= q0 + qw * 32));
                    const int nch = (wlim + 31) >> 5;                  // 32-column chunks with visible columns
                    const int pch = (((valid + 15) >> 4) + 1) >> 1;     // chunks the P·V MMA reads (16-column k-steps)
                    if (warp_has_rows) {
                        uint32_t v[128];
                        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                        if (full) {
                            // fast path, straight-line: the second half of the row is still in flight while the maximum of
                            // the first half is taken
                            ptx::tmem_ld_x32(tmem_S, v);
                            ptx::tmem_ld_x32(tmem_S + 32, v + 32);
                            ptx::tmem_wait_ld();
                            ptx::tmem_ld_x32(tmem_S + 64, v + 64);
                            ptx::tmem_ld_x32(tmem_S + 96, v + 96);
                            max_chunk<0>(v, mx);
                            max_chunk<1>(v, mx);
                            ptx::tmem_wait_ld();
                            max_chunk<2>(v, mx);
                            max_chunk<3>(v, mx);
                        } else {
                            ptx::tmem_ld_x32(tmem_S, v);
                            if (nch > 1) ptx::tmem_ld_x32(tmem_S + 32, v + 32);
                            if (nch > 2) ptx::tmem_ld_x32(tmem_S + 64, v + 64);
                            if (nch > 3) ptx::tmem_ld_x32(tmem_S + 96, v + 96);
                            ptx::tmem_wait_ld();
                            mask_chunk<0>(v, lim);
                            if (nch > 1) mask_chunk<1>(v, lim);
                            if (nch > 2) mask_chunk<2>(v, lim);
                            if (nch > 3) mask_chunk<3>(v, lim);
                            max_chunk<0>(v, mx);
                            if (nch > 1) max_chunk<1>(v, mx);
                            if (nch > 2) max_chunk<2>(v, mx);
                            if (nch > 3) max_chunk<3>(v, mx);
                        }
                        const float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale_log2;  // scale > 0
                        if constexpr (PSMEM) psmem_handoff();
                        // ---- lazy rescale of the TMEM accumulator
                        float alpha = 1.0f;
                        bool need = false;
                        if (j == 0) {
                            m_run = bm;
                            if constexpr (KVT) {   // the trailing key is the first "block" of the online softmax (p = 2^(m_tail - m))
                                const float m_tail = sMT[tt * BQ + r];
                                m_run = fmaxf(bm, m_tail);
                                l_run = ptx::ex2_approx(m_tail - m_run);
                            }
                        } else if (bm > m_run + kRescaleThreshold) {
                            alpha = ptx::ex2_approx(m_run - bm);
                            m_run = bm;
                            need = true;
                        }
                        if (__any_sync(0xffffffffu, need)) {
                            // The previous P·V of this tile must have landed in O_t. pv_done is kept per score buffer so that
                            // the parity wait is exact: that buffer's previous use completed before S of the previous step
                            // was issued, and its next use needs a P that does not exist yet - the count is n-1 or n.
                            if constexpr (!PSMEM) wait_wd(&pv_done[prev_sb], prev_sph, 2);
                            ptx::tc_fence_after();
                            l_run *= alpha;
#pragma unroll 1
                            for (int c = 0; c < HD / 8; ++c) {
                                uint32_t o[8];
                                ptx::tmem_ld_x8(tmem_O + c * 8, o);
                                ptx::tmem_wait_ld();
#pragma unroll
                                for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                                ptx::tmem_st_x8(tmem_O + c * 8, o);
                            }
                            ptx::tmem_wait_st();
                        }
                        // ---- p = exp2(s*c - m), row sum, bf16 pack; P (bf16 pairs) overwrites the first 64 columns of S
                        if (tracing) trace_ev(args.trace, tt, n_tr, 2, (int)n_mine);
                        const uint64_t sc2 = ptx::pack_f32x2(scale_log2, scale_log2);
                        const uint64_t nm2 = ptx::pack_f32x2(-m_run, -m_run);
                        uint64_t ps2[4];
#pragma unroll
                        for (int a = 0; a < 4; ++a) ps2[a] = ptx::pack_f32x2(0.f, 0.f);
                        uint32_t pk[16];
                        uint32_t zero16[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) zero16[i] = 0u;
                        if constexpr (PSMEM) {
                            // K-major SW128 tile [128 rows][2 atoms of 64 keys]: 16-byte chunk c of row r sits at c ^ (r & 7)
                            const uint32_t prow = ptx::smem_u32(sP) + tt * (BQ * BKV * 2) + (r >> 3) * 1024 + (r & 7) * 128;
                            const uint32_t rx = r & 7;
                            if (full) {
                                uint32_t pk1[16], pk2[16], pk3[16];
                                exp_chunk<0, POLY>(v, pk, sc2, nm2, ps2);
                                st_p_smem<0>(prow, rx, pk);
                                exp_chunk<1, POLY>(v, pk1, sc2, nm2, ps2);
                                st_p_smem<1>(prow, rx, pk1);
                                exp_chunk<2, POLY>(v, pk2, sc2, nm2, ps2);
                                st_p_smem<2>(prow, rx, pk2);
                                exp_chunk<3, POLY>(v, pk3, sc2, nm2, ps2);
                                st_p_smem<3>(prow, rx, pk3);
                            } else {
                                exp_chunk<0>(v, pk, sc2, nm2, ps2);
                                st_p_smem<0>(prow, rx, pk);
                                if (nch > 1) {
                                    exp_chunk<1>(v, pk, sc2, nm2, ps2);
                                    st_p_smem<1>(prow, rx, pk);
                                } else if (pch > 1) {
                                    st_p_smem<1>(prow, rx, zero16);
                                }
                                if (nch > 2) {
                                    exp_chunk<2>(v, pk, sc2, nm2, ps2);
                                    st_p_smem<2>(prow, rx, pk);
                                } else if (pch > 2) {
                                    st_p_smem<2>(prow, rx, zero16);
                                }
                                if (nch > 3) {
                                    exp_chunk<3>(v, pk, sc2, nm2, ps2);
                                    st_p_smem<3>(prow, rx, pk);
                                } else if (pch > 3) {
                                    st_p_smem<3>(prow, rx, zero16);
                                }
                            }
                            ptx::fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
                        } else if (full) {
                            // one basic block for the whole row: the tail of a chunk (its last pairs wait for the MUFU
                            // latency, then the store) overlaps the exponentials of the next chunk
                            uint32_t pk1[16], pk2[16], pk3[16];
                            exp_chunk<0, POLY>(v, pk, sc2, nm2, ps2);
                            ptx::tmem_st_x16(tmem_S, pk);
                            exp_chunk<1, POLY>(v, pk1, sc2, nm2, ps2);
                            ptx::tmem_st_x16(tmem_S + 16, pk1);
                            exp_chunk<2, POLY>(v, pk2, sc2, nm2, ps2);
                            ptx::tmem_st_x16(tmem_S + 32, pk2);
                            exp_chunk<3, POLY>(v, pk3, sc2, nm2, ps2);
                            ptx::tmem_st_x16(tmem_S + 48, pk3);
                        } else {
                            exp_chunk<0>(v, pk, sc2, nm2, ps2);
                            ptx::tmem_st_x16(tmem_S, pk);
                            if (nch > 1) {
                                exp_chunk<1>(v, pk, sc2, nm2, ps2);
                                ptx::tmem_st_x16(tmem_S + 16, pk);
                            } else if (pch > 1) {
                                ptx::tmem_st_x16(tmem_S + 16, zero16);
                            }
                            if (nch > 2) {
                                exp_chunk<2>(v, pk, sc2, nm2, ps2);
                                ptx::tmem_st_x16(tmem_S + 32, pk);
                            } else if (pch > 2) {
                                ptx::tmem_st_x16(tmem_S + 32, zero16);
                            }
                            if (nch > 3) {
                                exp_chunk<3>(v, pk, sc2, nm2, ps2);
                                ptx::tmem_st_x16(tmem_S + 48, pk);
                            } else if (pch > 3) {
                                ptx::tmem_st_x16(tmem_S + 48, zero16);
                            }
                        }
                        float s0, s1, s2, s3;
                        ptx::unpack_f32x2(ptx::fadd2(ps2[0], ps2[1]), s0, s1);
                        ptx::unpack_f32x2(ptx::fadd2(ps2[2], ps2[3]), s2, s3);
                        l_run += (s0 + s1) + (s2 + s3);
                        if (tracing) trace_ev(args.trace, tt, n_tr, 3, (int)n_mine);
                        if constexpr (!PSMEM) ptx::tmem_wait_st();
                        if (tracing) trace_ev(args.trace, tt, n_tr, 4, (int)n_mine);
                    } else {
                        if constexpr (PSMEM) psmem_handoff();
                    }
                    ptx::tc_fence_before();
                    __syncwarp();
                    // A warp without rows still waits for s_full above: S of the step that reuses this buffer (and this p_full
                    // barrier) is only issued after the P·V of this step, i.e. after all four arrivals of this phase.
                    if (lane == 0) ptx::mbar_arrive(&p_full[PSMEM ? tt : sb]);
                    ++n_mine;
                    prev_sb = sb;
                    prev_sph = sph;
                    if (++sb == kNSB) { sb = 0; sph ^= 1; }
                }
                if (tt == 0 && j < other_nblk) {  // tile 1's step of this block follows in the ring
                    if (++sb == kNSB) { sb = 0; sph ^= 1; }
                }
            }
            if (my_nblk > 0) {
                // hand the row sums to the epilogue warpgroup (which must be done with the previous item's)
                if (n_use > 0) wait_wd(&o_empty[tt], (n_use - 1) & 1, 3);
                sL[tt * BQ + r] = l_run;
                if constexpr (KVT) sPT[tt * BQ + r] = ptx::ex2_approx(sMT[tt * BQ + r] - m_run);   // 0 when the path is off (-inf)
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&l_ready[tt]);
                ++n_use;
            }
        }
    } else if (wg == 2) {
        // ===================================================== epilogue: O_t / l -> bf16 -> global
        ptx::setmaxnreg_dec<72>();
        const int qw = warp & 3;
        const int r = qw * 32 + lane;
        const uint32_t lane_sel = static_cast<uint32_t>(qw * 32) << 16;
        uint32_t n_use[2] = {0, 0};
        [[maybe_unused]] const bool tracing = DBG && MJV_TRACE_EPI && args.trace != nullptr && blockIdx.x == 0 && warp == 8 && lane == 0;
        [[maybe_unused]] int n_tr = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const Item t = decode_item<CAUSAL>(args, it);
            if constexpr (KVT) {
                if (args.kv_tail != 0) {
                    // V row of the trailing key, one element per thread (tile r / 64, dim r % 64), shared through sVT
                    const int tv = r >> 6, dv = r & 63;
                    float vt = 0.f;
                    if (sel2(tv, t.nblk) > 0)
                        vt = __bfloat162float(args.vt_ptr[static_cast<long long>(t.v_plane + (t.dual ? tv : 0)) * args.vt_plane_stride +
                                                          static_cast<long long>(dv) * args.vt_ld + t.v_col0 + t.len - 1]);
                    asm volatile("bar.sync 1, 128;" ::: "memory");   // the previous item's reads of sVT are done
                    sVT[tv * HD + dv] = vt;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
            }
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
                if (t.nblk[tt] == 0) continue;
                if constexpr (DBG) dbg_aux = (static_cast<long long>(it) << 32) | (tt << 16) | n_use[tt];
                if (MJV_TRACE_EPI && tracing) trace_ev(args.trace, 3, n_tr, 20, (int)(n_use[tt] * 2 + tt));
                wait_wd(&l_ready[tt], n_use[tt] & 1, 4);
                if (MJV_TRACE_EPI && tracing) trace_ev(args.trace, 3, n_tr, 21, (int)(n_use[tt] * 2 + tt));
                wait_wd(&o_done[tt], n_use[tt] & 1, 5);
                if (MJV_TRACE_EPI && tracing) trace_ev(args.trace, 3, n_tr, 22, (int)(n_use[tt] * 2 + tt));
                ptx::tc_fence_after();
                const int qrow = t.q0[tt] + r;
                const bool warp_has_rows = (t.q0[tt] + qw * 32) < t.len;
                if (warp_has_rows) {
                    const float inv = 1.0f / sL[tt * BQ + r];
                    [[maybe_unused]] float p_tail = 0.f;
                    if constexpr (KVT) p_tail = sPT[tt * BQ + r];
                    // One thread owns one row in TMEM: a store instruction writes 16 bytes of 32 different rows. A variant that
                    // transposed 4 x 4 quads inside lane groups (8 rows x 64 contiguous bytes per instruction) measured no
                    // faster: the drain (2500-4500 cycles per tile, tools/attn_trace.py with -DMJV_TRACE_EPI=1) is bound by
                    // the issue slots this warpgroup gets beside the two softmax warpgroups, not by the LSU.
                    __nv_bfloat16* dst =
                        args.out + static_cast<long long>(t.row_base + qrow) * args.ldo + t.h[tt] * args.out_hd;
                    const uint32_t tmem_O = tmem_O0 + tt * HD + lane_sel;
#pragma unroll 1
                    for (int c = 0; c < HD / 32; ++c) {
                        if (c * 32 >= args.out_hd) break;  // columns beyond the model's head_dim are exact zeros: not stored
                        uint32_t o[32];
                        ptx::tmem_ld_x32(tmem_O + c * 32, o);
                        ptx::tmem_wait_ld();
                        if constexpr (KVT) {
                            if (args.kv_tail != 0) {   // + p_last * v_last (fp32, before the normalisation)
                                const float4* v4 = reinterpret_cast<const float4*>(sVT + tt * HD + c * 32);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 vv = v4[i];
                                    o[4 * i + 0] = __float_as_uint(fmaf(p_tail, vv.x, __uint_as_float(o[4 * i + 0])));
                                    o[4 * i + 1] = __float_as_uint(fmaf(p_tail, vv.y, __uint_as_float(o[4 * i + 1])));
                                    o[4 * i + 2] = __float_as_uint(fmaf(p_tail, vv.z, __uint_as_float(o[4 * i + 2])));
                                    o[4 * i + 3] = __float_as_uint(fmaf(p_tail, vv.w, __uint_as_float(o[4 * i + 3])));
                                }
                            }
                        }
                        if (qrow < t.len) {
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
                if (MJV_TRACE_EPI && tracing) trace_ev(args.trace, 3, n_tr, 23, (int)(n_use[tt] * 2 + tt));
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&o_empty[tt]);
                ++n_use[tt];
            }
        }
    } else {
        ptx::setmaxnreg_dec<56>();
        if (warp == 12) {
            // ===================================================== TMA producer
            if (ptx::elect_one()) {
                int kst = 0, vst = 0;
                uint32_t kph = 0, vph = 0;
                uint32_t n_use[2] = {0, 0};
                for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                    const Item t = decode_item<CAUSAL>(args, it);
#pragma unroll
                    for (int tt = 0; tt < 2; ++tt) {
                        if (t.nblk[tt] == 0) continue;
                        const int buf = (n_use[tt] % QSETS) * 2 + tt;
                        wait_sp(&q_empty[buf], ((n_use[tt] / QSETS) & 1) ^ 1, 10);
                        ptx::mbar_expect_tx(&q_full[buf], L::kQBytes);
#pragma unroll
                        for (int kd = 0; kd < HD / 64; ++kd)
                            ptx::tma_load_3d(sQ + buf * L::kQBytes + kd * (BQ * 128), &tmQ, &q_full[buf],
                                             args.q_col0 + t.h[tt] * args.qk_stride + kd * 64, t.tma_row0 + t.q0[tt], t.tma_batch,
                                             ptx::kEvictFirst);
                        ++n_use[tt];
                    }
                    for (int j = 0; j < t.nblk_max; ++j) {
                        if constexpr (DBG) dbg_aux = (static_cast<long long>(it) << 32) | (static_cast<long long>(j) << 16) | (vst << 8) | vph;
                        // a dual item streams K/V of head h for tile 0, then of head h + 1 for tile 1, block by block
                        for (int s = 0; s <= t.dual; ++s) {
                            wait_sp(&k_empty[kst], kph ^ 1, 11);
                            ptx::mbar_expect_tx(&k_full[kst], L::kKBytes);
#pragma unroll
                            for (int kd = 0; kd < HD / 64; ++kd)
                                ptx::tma_load_3d(sK + kst * L::kKBytes + kd * (BKV * 128), &tmK, &k_full[kst],
                                                 args.k_col0 + (t.kvh + s) * args.qk_stride + kd * 64, t.tma_row0 + j * BKV, t.tma_batch,
                                                 ptx::kEvictLast);
                            if (++kst == KS) { kst = 0; kph ^= 1; }
                            if constexpr (KS == 1) {
                                // single K stage: the next block's load can only be issued when S of this one has retired -
                                // pull it into L2 now, so that load costs an L2 hit (~700 cycles) and not an HBM round trip
                                // (~2200, tools/attn_trace.py: score waits of 1250 cycles per step whenever K came from HBM)
                                if (j + 1 < t.nblk_max && !t.dual) {
#pragma unroll
                                    for (int kd = 0; kd < HD / 64; ++kd)
                                        ptx::tma_prefetch_3d(&tmK, args.k_col0 + t.kvh * args.qk_stride + kd * 64,
                                                             t.tma_row0 + (j + 1) * BKV, t.tma_batch);
                                }
                            }
                            wait_sp(&v_empty[vst], vph ^ 1, 12);
                            ptx::mbar_expect_tx(&v_full[vst], L::kVBytes);
#pragma unroll
                            for (int half = 0; half < 2; ++half)
                                ptx::tma_load_3d(sV + vst * L::kVBytes + half * (HD * 128), &tmV, &v_full[vst],
                                                 t.v_col0 + j * BKV + half * 64, 0, t.v_plane + s, ptx::kEvictLast);
                            if (++vst == VS) { vst = 0; vph ^= 1; }
                        }
                    }
                }
            }
        } else if (warp == 13) {
            // ===================================================== P·V issuer
            // Two issuing warps (this one and the S issuer below) instead of one: a single issuer is blocked while the tensor
            // pipe drains its queue, so with one warp the 4 S MMAs of step g+kNSB sat between P_g·V and the wait for P_{g+1}
            // (1100 cycles per step, tools/attn_trace.py) and the issuer, not the softmax, set the pace. Both warps run
            // their bookkeeping and barrier waits warp-uniformly (descriptors stay in uniform registers); only the elected
            // lane issues tcgen05.mma / tcgen05.commit. MMAs of different threads are not ordered against each other, so
            // the reuse of a score buffer is explicit: the S issuer waits for pv_done[b] of the step that last read it.
            const bool leader = ptx::elect_one();
            // N of P.V = rows of the V^T plane that exist (a multiple of 16; 96 for head_dim 96: the zero-filled rows are skipped)
            const uint32_t idesc_o = ptx::make_idesc_bf16(BQ, args.pv_n);
            const uint64_t v_desc0 = ptx::make_desc_sw128(ptx::smem_u32(sV));
            [[maybe_unused]] const uint64_t p_desc0 = ptx::make_desc_sw128(ptx::smem_u32(sP));
            int vst = 0;
            uint32_t vph = 0;
            int g = 0;
            int p_sb = 0;             // score buffer holding P of the next P·V step
            uint32_t p_ph = 0;
            uint32_t o_use[2] = {0, 0};
            [[maybe_unused]] uint32_t p_use[2] = {0, 0};   // P-in-smem mode: steps of each tile so far (p_full[tt] phases)
            const bool tracing = DBG && args.trace != nullptr && blockIdx.x == 0 && leader;
            int n_tr = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const ItemLite t = decode_lite<CAUSAL>(args, it);
                for (int j = 0; j < t.nblk_max; ++j) {
#pragma unroll 1
                    for (int tt = 0; tt < 2; ++tt) {  // one copy of the code: the issuer's instruction footprint matters
                        if (!(j < sel2(tt, t.nblk))) continue;
                        if constexpr (DBG) dbg_aux = (static_cast<long long>(it) << 32) | (static_cast<long long>(g) << 8) | (p_sb << 4) | p_ph;
                        const uint64_t vb0 = v_desc0 + static_cast<uint64_t>((vst * L::kVBytes) >> 4);
                        const bool first_of_block = t.dual || (tt == 0) || !(j < t.nblk[0]);
                        const bool last_of_block = t.dual || (tt == 1) || !(j < t.nblk[1]);
                        const bool last_of_tile = j == sel2(tt, t.nblk) - 1;
                        // waits that are normally satisfied long before P arrives come first: off the critical path
                        if (first_of_block) wait_sp(&v_full[vst], vph, 20);
                        const uint32_t ou = sel2(tt, o_use);
                        // the epilogue has read out the previous item's accumulator of this tile
                        if (j == 0 && ou > 0) wait_sp(&o_empty[tt], (ou - 1) & 1, 21);
                        const int valid = min(BKV, sel2(tt, t.kv_end) - j * BKV);
                        const int ksteps = (valid + 15) >> 4;
                        const uint32_t p_tmem = tmem_base + p_sb * BKV;  // P_g: bf16 pairs in the first 64 columns
                        const uint32_t o_tmem = tmem_O0 + tt * HD;
                        if constexpr (PSMEM) wait_sp(&p_full[tt], sel2(tt, p_use) & 1, 22);
                        else wait_sp(&p_full[p_sb], p_ph, 22);
                        ptx::tc_fence_after();
                        if (tracing) trace_ev(args.trace, 2, n_tr, 10, g);
                        if (leader) {
#pragma unroll
                            for (int ks = 0; ks < BKV / 16; ++ks) {
                                if (ks < ksteps) {
                                    const uint64_t b = vb0 + static_cast<uint64_t>(((ks >> 2) * (HD * 128) + (ks & 3) * 32) >> 4);
                                    if constexpr (PSMEM) {
                                        const uint64_t a = p_desc0 + static_cast<uint64_t>((tt * (BQ * BKV * 2) + (ks >> 2) * (BQ * 128) + (ks & 3) * 32) >> 4);
                                        ptx::mma_bf16_ss(o_tmem, a, b, idesc_o, (j | ks) != 0 ? 1u : 0u);
                                    } else {
                                        ptx::mma_bf16_ts(o_tmem, p_tmem + ks * 8, b, idesc_o, (j | ks) != 0 ? 1u : 0u);
                                    }
                                }
                            }
                            if constexpr (PSMEM) ptx::mma_commit(&pvd[tt]);
                            else ptx::mma_commit(&pv_done[p_sb]);
                            if (last_of_tile) ptx::mma_commit(&o_done[tt]);
                            if (last_of_block) ptx::mma_commit(&v_empty[vst]);
                        }
                        __syncwarp();
                        if (tracing) trace_ev(args.trace, 2, n_tr, 11, g);
                        if (last_of_tile) {
                            if (tt != 0) ++o_use[1];
                            else ++o_use[0];
                        }
                        if (last_of_block) {
                            if (++vst == VS) { vst = 0; vph ^= 1; }
                        }
                        ++g;
                        if constexpr (PSMEM) {
                            if (tt != 0) ++p_use[1];
                            else ++p_use[0];
                        }
                        if (++p_sb == kNSB) { p_sb = 0; p_ph ^= 1; }
                    }
                }
            }
        } else if (warp == 14) {
            // ===================================================== S issuer: S_g = Q_t · K_j^T into score buffer g % kNSB
            const bool leader = ptx::elect_one();
            constexpr uint32_t idesc_s = ptx::make_idesc_bf16(BQ, BKV);
            const uint64_t q_desc0 = ptx::make_desc_sw128(ptx::smem_u32(sQ));
            const uint64_t k_desc0 = ptx::make_desc_sw128(ptx::smem_u32(sK));
            const int qk_ksteps = args.qk_ksteps;
            int kst = 0;
            uint32_t kph = 0;
            uint32_t q_use[2] = {0, 0};
            int t_slot[2] = {0, 0};   // s_full slot of each tile's next step
            int sb = 0;               // score buffer of the next S step
            uint32_t sph = 0;         // parity of its CURRENT use; its previous use completed pv_done[sb] in phase sph ^ 1
            bool wrapped = false;     // every buffer has been used once
            const bool tracing = DBG && !MJV_TRACE_EPI && args.trace != nullptr && blockIdx.x == 0 && leader;
            int n_tr = 0;
            int gs = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const ItemLite t = decode_lite<CAUSAL>(args, it);
                for (int j = 0; j < t.nblk_max; ++j) {
#pragma unroll 1
                    for (int tt = 0; tt < 2; ++tt) {
                        if (!(j < sel2(tt, t.nblk))) continue;
                        const uint64_t b0 = k_desc0 + static_cast<uint64_t>((kst * L::kKBytes) >> 4);
                        const bool first_of_block = t.dual || (tt == 0) || !(j < t.nblk[0]);
                        const bool last_of_block = t.dual || (tt == 1) || !(j < t.nblk[1]);
                        const bool last_of_tile = j == sel2(tt, t.nblk) - 1;
                        const uint32_t use = sel2(tt, q_use);
                        const int buf = (use % QSETS) * 2 + tt;
                        if constexpr (DBG) dbg_aux = (static_cast<long long>(it) << 32) | (static_cast<long long>(gs) << 8) | (sb << 4) | sph;
                        if (j == 0) wait_sp(&q_full[buf], (use / QSETS) & 1, 30);
                        if (first_of_block) wait_sp(&k_full[kst], kph, 31);
                        // the P·V that read this buffer's previous contents (P of step gs - kNSB) has completed
                        // (P-in-smem mode: its scores have been read into registers - P·V no longer touches the buffer)
                        if (wrapped) wait_sp(PSMEM ? &s_free[sb] : &pv_done[sb], sph ^ 1, 32);
                        ptx::tc_fence_after();
                        const uint64_t a0 = q_desc0 + static_cast<uint64_t>((buf * L::kQBytes) >> 4);
                        const uint32_t d = tmem_base + sb * BKV;
                        if (leader) {
#pragma unroll
                            for (int ks = 0; ks < HD / 16; ++ks) {
                                // packed head_dim-96 heads: columns 96..127 of the tile belong to the next head, not read
                                if (ks < qk_ksteps) {
                                    const uint64_t a = a0 + static_cast<uint64_t>(((ks >> 2) * (BQ * 128) + (ks & 3) * 32) >> 4);
                                    const uint64_t b = b0 + static_cast<uint64_t>(((ks >> 2) * (BKV * 128) + (ks & 3) * 32) >> 4);
                                    ptx::mma_bf16_ss(d, a, b, idesc_s, ks != 0 ? 1u : 0u);
                                }
                            }
                            ptx::mma_commit(&s_full[tt * kNSB + sel2(tt, t_slot)]);
                            if (last_of_block) ptx::mma_commit(&k_empty[kst]);  // K tile free once the S MMAs of both tiles retire
                            if (last_of_tile) ptx::mma_commit(&q_empty[buf]);   // last S of this tile in the item
                        }
                        __syncwarp();
                        if (tt != 0) {
                            if (++t_slot[1] == kNSB) t_slot[1] = 0;
                        } else {
                            if (++t_slot[0] == kNSB) t_slot[0] = 0;
                        }
                        if (tracing) trace_ev(args.trace, 3, n_tr, 12, gs);
                        if (last_of_block) {
                            if (++kst == KS) { kst = 0; kph ^= 1; }
                        }
                        if (last_of_tile) {
                            if (tt != 0) ++q_use[1];
                            else ++q_use[0];
                        }
                        ++gs;
                        if (++sb == kNSB) { sb = 0; sph ^= 1; wrapped = true; }
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 15) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace mjv
