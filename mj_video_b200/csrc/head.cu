// MoE reward head (scripts/model/moe_reward.py:16-42, 236-297): two gating MLPs (Linear+ReLU x n_hidden, Linear),
// the bias-free per-criterion regression, temperature softmax over aspects, per-aspect-group softmax over criteria,
// aspect2criteria weighting and the final score.
//
// Skinny GEMMs (B rows x {1024,28,5} outputs): weights are streamed from HBM exactly once per launch with 16-byte
// loads; activations are staged in shared memory in K-chunks; each warp owns an [8 rows x 4 outputs] register tile with
// K split across lanes and finishes with warp-shuffle reductions. Layers of the two gating nets (and the regression
// layer) that do not depend on each other share a launch. A last one-warp-per-row kernel does the softmaxes and the
// weighted sums with warp shuffles. bf16 rounding points mirror the reference's bf16 path.
#include "common.cuh"
#include "head.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace mjv {
namespace {

constexpr int kRowsPerWarp = 8;
constexpr int kOutsPerWarp = 4;
constexpr int kRowGroups = 4;                           // 32 rows per block
constexpr int kOutGroups = 2;                           // 8 outputs per block
constexpr int kWarps = kRowGroups * kOutGroups;         // 8 warps
constexpr int kChunkK = 512;                            // staged K-chunk
constexpr int kRowsPerBlock = kRowsPerWarp * kRowGroups;
constexpr int kOutsPerBlock = kOutsPerWarp * kOutGroups;

__global__ void __launch_bounds__(kWarps * 32)
skinny_linear_kernel(const HeadJobs jobs, int B) {
    const HeadJob* jobs_arg = jobs.j;
    const int n_jobs = jobs.n;
    __shared__ __align__(16) __nv_bfloat16 sx[kRowsPerBlock][kChunkK];
    // block -> (job, output block, row block)
    __shared__ HeadJob job;
    int blk = blockIdx.x;
    int jsel = -1, oblk = 0, rblk = 0;
    const int row_blocks = (B + kRowsPerBlock - 1) / kRowsPerBlock;
    for (int j = 0; j < n_jobs; ++j) {
        const int nb = ((jobs_arg[j].N + kOutsPerBlock - 1) / kOutsPerBlock) * row_blocks;
        if (blk < nb) {
            jsel = j;
            oblk = blk / row_blocks;
            rblk = blk % row_blocks;
            break;
        }
        blk -= nb;
    }
    if (jsel < 0) return;
    if (threadIdx.x == 0) job = jobs_arg[jsel];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rg = warp % kRowGroups, og = warp / kRowGroups;
    const int row0 = rblk * kRowsPerBlock;
    const int n0 = oblk * kOutsPerBlock + og * kOutsPerWarp;
    const int K = job.K;
    const __nv_bfloat16* X = static_cast<const __nv_bfloat16*>(job.x);
    const __nv_bfloat16* W = static_cast<const __nv_bfloat16*>(job.w);

    float acc[kOutsPerWarp][kRowsPerWarp];
#pragma unroll
    for (int o = 0; o < kOutsPerWarp; ++o)
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) acc[o][r] = 0.f;

    for (int k0 = 0; k0 < K; k0 += kChunkK) {
        const int kc = min(kChunkK, K - k0);
        // stage x[row0:row0+32, k0:k0+kc] (zero-filled beyond B / K)
        for (int i = threadIdx.x; i < kRowsPerBlock * (kChunkK / 8); i += blockDim.x) {
            const int r = i / (kChunkK / 8), c8 = i % (kChunkK / 8);
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (row0 + r < B && c8 * 8 < kc)
                v = *reinterpret_cast<const uint4*>(X + static_cast<long long>(row0 + r) * job.ldx + k0 + c8 * 8);
            *reinterpret_cast<uint4*>(&sx[r][c8 * 8]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int sub = 0; sub < kChunkK / 256; ++sub) {
            const int kk = sub * 256 + lane * 8;
            if (kk < kc) {
                float wv[kOutsPerWarp][8];
#pragma unroll
                for (int o = 0; o < kOutsPerWarp; ++o) {
                    uint4 u = make_uint4(0u, 0u, 0u, 0u);
                    if (n0 + o < job.N) u = __ldg(reinterpret_cast<const uint4*>(W + static_cast<long long>(n0 + o) * K + k0 + kk));
                    wv[o][0] = ptx::bf16_lo(u.x); wv[o][1] = ptx::bf16_hi(u.x);
                    wv[o][2] = ptx::bf16_lo(u.y); wv[o][3] = ptx::bf16_hi(u.y);
                    wv[o][4] = ptx::bf16_lo(u.z); wv[o][5] = ptx::bf16_hi(u.z);
                    wv[o][6] = ptx::bf16_lo(u.w); wv[o][7] = ptx::bf16_hi(u.w);
                }
#pragma unroll
                for (int r = 0; r < kRowsPerWarp; ++r) {
                    const uint4 u = *reinterpret_cast<const uint4*>(&sx[rg * kRowsPerWarp + r][kk]);
                    const float xv[8] = {ptx::bf16_lo(u.x), ptx::bf16_hi(u.x), ptx::bf16_lo(u.y), ptx::bf16_hi(u.y),
                                         ptx::bf16_lo(u.z), ptx::bf16_hi(u.z), ptx::bf16_lo(u.w), ptx::bf16_hi(u.w)};
#pragma unroll
                    for (int o = 0; o < kOutsPerWarp; ++o)
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[o][r] = fmaf(xv[e], wv[o][e], acc[o][r]);
                }
            }
        }
        __syncthreads();
    }
    // warp-shuffle reduction over the K split, then bias / ReLU / bf16 store
#pragma unroll
    for (int o = 0; o < kOutsPerWarp; ++o)
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            float v = acc[o][r];
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            acc[o][r] = v;
        }
    if (lane < kOutsPerWarp * kRowsPerWarp) {
        const int o = lane / kRowsPerWarp, r = lane % kRowsPerWarp;
        const int n = n0 + o, row = row0 + rg * kRowsPerWarp + r;
        if (n < job.N && row < B) {
            float v = 0.f;
#pragma unroll
            for (int oo = 0; oo < kOutsPerWarp; ++oo)
#pragma unroll
                for (int rr = 0; rr < kRowsPerWarp; ++rr)
                    if (oo == o && rr == r) v = acc[oo][rr];
            if (job.bias != nullptr) v += __bfloat162float(static_cast<const __nv_bfloat16*>(job.bias)[n]);
            v = ptx::bf16_round(v);
            if (job.relu) v = fmaxf(v, 0.f);
            static_cast<__nv_bfloat16*>(job.y)[static_cast<long long>(row) * job.ldy + n] = __float2bfloat16_rn(v);
        }
    }
}

// One warp per sample: reward transform, aspect softmax, per-group criteria softmax, weighted sums, score.
__global__ void __launch_bounds__(128)
head_finalize_kernel(const HeadFinalizeArgs a) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.B) return;
    const int lane = threadIdx.x & 31;
    const int NO = a.num_objectives, NA = a.num_aspects;
    const float invT = 1.0f / a.temperature;  // reference divides; x / T in bf16
    // rewards @ reward_transform_matrix (moe_reward.py:240), bf16 matmul
    float rw = 0.f;
    if (lane < NO) {
        float s = 0.f;
        for (int i = 0; i < NO; ++i)
            s = fmaf(__bfloat162float(a.raw_rewards[row * a.ld_raw + i]), __bfloat162float(a.transform[i * NO + lane]), s);
        rw = ptx::bf16_round(s);
        a.scores[row * a.ld_scores + lane] = rw;
    }
    (void)invT;
    // aspect gate: softmax(x / T) * logit_scale (GatingNetwork.forward, moe_reward.py:29-35)
    const float lsa = __bfloat162float(a.aspect_logit_scale[0]);
    const float lsc = __bfloat162float(a.criteria_logit_scale[0]);
    float al = -INFINITY;
    if (lane < NA) al = ptx::bf16_round(__bfloat162float(a.aspect_logits[row * a.ld_al + lane]) / a.temperature);
    float mx = al;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    float e = lane < NA ? expf(al - mx) : 0.f;
    float sum = e;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    const float gate = lane < NA ? ptx::bf16_round(ptx::bf16_round(e / sum) * lsa) : 0.f;
    if (lane < NA) a.extras[row * a.ld_extras + NO + lane] = gate;  // aspect_gating_output
    // criteria logits (forward_wo_softmax) are an output as they are
    float cl = 0.f;
    if (lane < NO) {
        cl = __bfloat162float(a.criteria_logits[row * a.ld_cl + lane]);
        a.extras[row * a.ld_extras + lane] = cl;  // criteria_gating_output
    }
    // per-aspect groups (dict order). lane p handles concat position p -> criterion item[p] of group grp[p]
    int crit = 0, grp = -1;
    if (lane < NO) {
        crit = a.group_items[lane];
        for (int g = 0; g < NA; ++g)
            if (lane >= a.group_offsets[g] && lane < a.group_offsets[g + 1]) grp = g;
    }
    const float cl_sh = __shfl_sync(0xffffffffu, cl, crit);  // all 32 lanes take part in the shuffle
    const float myl = lane < NO ? ptx::bf16_round(cl_sh / a.temperature) : -INFINITY;
    const float myr = __shfl_sync(0xffffffffu, rw, crit);
    float score = 0.f, last_weighted = 0.f;
    for (int g = 0; g < NA; ++g) {
        const bool in = (grp == g);
        float m = in ? myl : -INFINITY;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
        const float ee = in ? expf(myl - m) : 0.f;
        float ss = ee;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
        const float w = in ? ptx::bf16_round(ptx::bf16_round(ee / ss) * lsc) : 0.f;
        if (in) a.extras[row * a.ld_extras + NO + NA + lane] = w;  // aspect_weights (concat order)
        float prod = in ? ptx::bf16_round(myr * w) : 0.f;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) prod += __shfl_xor_sync(0xffffffffu, prod, s);
        const float weighted = ptx::bf16_round(prod);  // (aspect_rewards * aspect_weight).sum(-1) in bf16
        const float g_gate = __shfl_sync(0xffffffffu, gate, g);
        if (lane == 0) a.scores[row * a.ld_scores + NO + g] = weighted;  // aspect_scores (fp32 container)
        score = fmaf(weighted, g_gate, score);
        last_weighted = weighted;
    }
    if (lane == 0) {
        a.scores[row * a.ld_scores + NO + NA] = score;
        a.extras[row * a.ld_extras + NO + NA + NO] = last_weighted;  // weighted_scores (leaked loop variable)
    }
}

// GatingNetwork.forward tail (moe_reward.py:33-36): softmax(x / T, dim=1) * logit_scale[0], one warp per row, n <= 32,
// with the reference's bf16 rounding points (x / T, the softmax output and the product are bf16 tensors there).
__global__ void __launch_bounds__(128)
gating_softmax_kernel(const __nv_bfloat16* __restrict__ logits, int ld, int n, float temperature,
                      const __nv_bfloat16* __restrict__ logit_scale, __nv_bfloat16* __restrict__ out, int ldo, int B) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    const int lane = threadIdx.x & 31;
    float x = -INFINITY;
    if (lane < n) x = ptx::bf16_round(__bfloat162float(logits[row * ld + lane]) / temperature);
    float mx = x;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    const float e = lane < n ? expf(x - mx) : 0.f;
    float sum = e;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    if (lane < n)
        out[row * ldo + lane] = __float2bfloat16_rn(ptx::bf16_round(e / sum) * __bfloat162float(logit_scale[0]));
}

}  // namespace

int gating_softmax(const void* logits, int ld, int n, float temperature, const void* logit_scale, void* out, int ldo, int B,
                   cudaStream_t st) {
    if (B == 0) return 0;
    MJV_CHECK(n >= 1 && n <= 32, MJV_ERR_ARG, "gating softmax: 1..32 outputs supported, got %d", n);
    {
        LaunchScope scope(TAG_HEAD_FINALIZE, st, 0.0);
        gating_softmax_kernel<<<(B + 3) / 4, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(logits), ld, n, temperature,
                                                            static_cast<const __nv_bfloat16*>(logit_scale),
                                                            static_cast<__nv_bfloat16*>(out), ldo, B);
    }
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int skinny_linear(const HeadJobs& jobs, int B, cudaStream_t st) {
    const HeadJob* jobs_host = jobs.j;
    const int n_jobs = jobs.n;
    if (B == 0 || n_jobs == 0) return 0;
    int blocks = 0;
    const int row_blocks = (B + kRowsPerBlock - 1) / kRowsPerBlock;
    for (int j = 0; j < n_jobs; ++j) {
        MJV_CHECK(jobs_host[j].K % 8 == 0 && jobs_host[j].ldx % 8 == 0, MJV_ERR_ARG, "head: K/ldx must be multiples of 8");
        blocks += ((jobs_host[j].N + kOutsPerBlock - 1) / kOutsPerBlock) * row_blocks;
    }
    double bytes = 0;
    for (int j = 0; j < n_jobs; ++j) bytes += 2.0 * jobs_host[j].N * jobs_host[j].K + 2.0 * B * (jobs_host[j].K + jobs_host[j].N);
    {
        LaunchScope scope(TAG_HEAD_LINEAR, st, bytes);
        skinny_linear_kernel<<<blocks, kWarps * 32, 0, st>>>(jobs, B);
    }
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int head_finalize(const HeadFinalizeArgs& a, cudaStream_t st) {
    if (a.B == 0) return 0;
    MJV_CHECK(a.num_objectives <= 32 && a.num_aspects <= 32, MJV_ERR_ARG, "head: at most 32 criteria / aspects");
    {
        LaunchScope scope(TAG_HEAD_FINALIZE, st, 0.0);
        head_finalize_kernel<<<(a.B + 3) / 4, 128, 0, st>>>(a);
    }
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace mjv
