// fp32 verification-mode kernels (see fp32_kernels.cuh). Plain SIMT fp32, IEEE sqrt/div, expf/erff from libdevice.
#include "fp32_kernels.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace mjv {
namespace f32 {
namespace {

// ------------------------------------------------------------------------------------------------ GEMM
constexpr int BM = 128, BN = 128, BK = 16;

template <int EPI>
__global__ void __launch_bounds__(256)
sgemm_tn_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ W, long long ldw,
                float* __restrict__ C, long long ldc, long long M, int N, int K, const float* __restrict__ bias,
                const float* __restrict__ res, long long ldr, const float* __restrict__ scale) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = static_cast<long long>(blockIdx.y) * BM;
    const int n0 = blockIdx.x * BN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 2;          // 0..63
    const int lk = (tid & 3) * 4;     // 0,4,8,12
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = lr + half * 64;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vw = va;
            if (m0 + r < M && k0 + lk < K) va = *reinterpret_cast<const float4*>(A + (m0 + r) * lda + k0 + lk);
            if (n0 + r < N && k0 + lk < K) vw = *reinterpret_cast<const float4*>(W + static_cast<long long>(n0 + r) * ldw + k0 + lk);
            As[lk + 0][r] = va.x; As[lk + 1][r] = va.y; As[lk + 2][r] = va.z; As[lk + 3][r] = va.w;
            Ws[lk + 0][r] = vw.x; Ws[lk + 1][r] = vw.y; Ws[lk + 2][r] = vw.z; Ws[lk + 3][r] = vw.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Ws[k][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Ws[k][tx * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long row = m0 + ty * 8 + i;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + tx * 8 + j;
            if (col >= N) continue;
            float x = acc[i][j];
            if (bias != nullptr) x += bias[col];
            if (EPI == EPI_GELU) x = 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));  // F.gelu (exact)
            if (EPI == EPI_RELU) x = fmaxf(x, 0.f);
            if (EPI == EPI_RESID) x = res[row * ldr + col] + x;
            if (EPI == EPI_RESID_SCALE) x = res[row * ldr + col] + x * scale[col];
            C[row * ldc + col] = x;
        }
    }
}

// ------------------------------------------------------------------------------------------------ norms
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                 float* __restrict__ out, long long rows, int C, float eps) {
    __shared__ float red[8];
    const long long r = blockIdx.x;
    const float* xr = x + r * C;
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += xr[c];
    const float mean = block_sum(s, red) / C;
    float q = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float d = xr[c] - mean;
        q += d * d;
    }
    const float var = block_sum(q, red) / C;
    const float rstd = 1.0f / sqrtf(var + eps);
    for (int c = threadIdx.x; c < C; c += blockDim.x) out[r * C + c] = (xr[c] - mean) * rstd * w[c] + b[c];
}

__global__ void __launch_bounds__(256)
rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out, long long rows, int C,
               float eps, const int* __restrict__ row_idx) {
    __shared__ float red[8];
    const long long r = blockIdx.x;
    const float* xr = x + (row_idx != nullptr ? static_cast<long long>(row_idx[r]) : r) * C;
    float q = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) q += xr[c] * xr[c];
    const float rstd = 1.0f / sqrtf(block_sum(q, red) / C + eps);
    for (int c = threadIdx.x; c < C; c += blockDim.x) out[r * C + c] = w[c] * (xr[c] * rstd);
}

// ------------------------------------------------------------------------------------------------ data movement
template <typename TIn>
__global__ void patchify_kernel(const TIn* __restrict__ px, float* __restrict__ out, int T, int img, int ps) {
    const int g = img / ps;
    const int kdim = 3 * ps * ps;
    const long long total = static_cast<long long>(T) * g * g * kdim;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int e = static_cast<int>(i % kdim);
        const long long p = i / kdim;
        const int px_ = static_cast<int>(p % g), py = static_cast<int>((p / g) % g);
        const long long t = p / (g * g);
        const int c = e / (ps * ps), ky = (e / ps) % ps, kx = e % ps;
        out[i] = static_cast<float>(px[((t * 3 + c) * img + py * ps + ky) * img + px_ * ps + kx]);
    }
}

__global__ void vit_assemble_kernel(const float* __restrict__ conv, const float* __restrict__ cls, const float* __restrict__ pos,
                                    float* __restrict__ hidden, int T, int P, int C) {
    const long long total = static_cast<long long>(T) * (P + 1) * C;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const int s = static_cast<int>((i / C) % (P + 1));
        const long long t = i / (static_cast<long long>(C) * (P + 1));
        const float v = s == 0 ? cls[c] : conv[(t * P + (s - 1)) * C + c];
        hidden[i] = v + pos[static_cast<long long>(s) * C + c];
    }
}

// out[t, q, c4] <- hidden[t, 1 + a*g + b, c] with the index algebra of pixel_shuffle (scale 0.5):
//   v2: q = a2*(g/2) + b2;  v1: q = b2*(g/2) + a2;  a = 2*a2 + c4 / 2C;  c2 = c4 % 2C;  b = 2*b2 + c2 / C;  c = c2 % C
__global__ void pixel_shuffle_kernel(const float* __restrict__ hidden, float* __restrict__ out, int T, int g, int C, int ps_v1) {
    const int g2 = g / 2;
    const long long total = static_cast<long long>(T) * g2 * g2 * 4 * C;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c4 = static_cast<int>(i % (4 * C));
        const int q = static_cast<int>((i / (4 * C)) % (g2 * g2));
        const long long t = i / (static_cast<long long>(4 * C) * g2 * g2);
        const int a2 = ps_v1 ? q % g2 : q / g2;
        const int b2 = ps_v1 ? q / g2 : q % g2;
        const int a = 2 * a2 + c4 / (2 * C);
        const int c2 = c4 % (2 * C);
        const int b = 2 * b2 + c2 / C;
        const int c = c2 % C;
        out[i] = hidden[(t * (g * g + 1) + 1 + a * g + b) * C + c];
    }
}

__global__ void embed_kernel(const int* __restrict__ ids, const float* __restrict__ table, float* __restrict__ out,
                             long long rows, int C, int skip_id) {
    const long long total = rows * C;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / C;
        const int id = ids[r];
        if (id == skip_id) continue;
        out[i] = table[static_cast<long long>(id) * C + (i - r * C)];
    }
}

template <bool SCATTER>
__global__ void move_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, float* __restrict__ dst,
                                 long long n, int C) {
    const long long total = n * C;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / C;
        const int c = static_cast<int>(i - r * C);
        if (SCATTER) dst[static_cast<long long>(idx[r]) * C + c] = src[i];
        else dst[i] = src[static_cast<long long>(idx[r]) * C + c];
    }
}

// q*cos + rotate_half(q)*sin with separately rounded products, as the reference's tensor expression evaluates it
__global__ void rope_kernel(float* __restrict__ x, long long ld, long long rows, int heads, int group, int kv_stride, int col0,
                            int hd, const float* __restrict__ cos_t, const float* __restrict__ sin_t, int table_ld,
                            const int* __restrict__ row_pos) {
    const int half = hd / 2;
    const long long total = rows * heads * half;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int d = static_cast<int>(i % half);
        const int h = static_cast<int>((i / half) % heads);
        const long long r = i / (static_cast<long long>(half) * heads);
        float* p = x + r * ld + col0 + (h / group) * kv_stride + (h % group) * hd;
        const int pos = row_pos[r];
        const float c = cos_t[static_cast<long long>(pos) * table_ld + d], s = sin_t[static_cast<long long>(pos) * table_ld + d];
        const float x1 = p[d], x2 = p[d + half];
        p[d] = __fadd_rn(__fmul_rn(x1, c), __fmul_rn(-x2, s));
        p[d + half] = __fadd_rn(__fmul_rn(x2, c), __fmul_rn(x1, s));
    }
}

__global__ void silu_mul_kernel(const float* __restrict__ gate, const float* __restrict__ up, float* __restrict__ out, long long n) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float g = gate[i];
        out[i] = (g / (1.0f + expf(-g))) * up[i];
    }
}

__global__ void to_f32_kernel(const void* __restrict__ src, int dtype, float* __restrict__ dst, long long n) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float v;
        if (dtype == MJV_DTYPE_F32) v = static_cast<const float*>(src)[i];
        else if (dtype == MJV_DTYPE_BF16) v = __bfloat162float(static_cast<const __nv_bfloat16*>(src)[i]);
        else v = __half2float(static_cast<const __half*>(src)[i]);
        dst[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------ attention
// One warp per query row; lane l owns the keys j = l, l + 32, ... with its own running (max, sum, acc[HD]); the 32
// partial softmaxes are merged at the end. Scores are scale * (q . k), masked keys are simply not visited (the reference
// adds finfo.min there, whose exp underflows to exactly 0).
template <int HD>
__global__ void __launch_bounds__(128)
attention_kernel(const AttnF32 a) {
    __shared__ float qs[4][HD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seq = blockIdx.z, h = blockIdx.y;
    long long row_base;
    int len;
    if (a.cu_seqlens != nullptr) {
        row_base = a.cu_seqlens[seq];
        len = a.cu_seqlens[seq + 1] - a.cu_seqlens[seq];
    } else {
        len = a.seq_len;
        row_base = static_cast<long long>(seq) * len;
    }
    const int qi = blockIdx.x * 4 + warp;
    const bool active = qi < len;
    const int kvh = h / a.group;
    const int qcol = a.q_col0 + kvh * a.q_group_stride + (h % a.group) * HD;
    const int kcol = a.k_col0 + kvh * a.kv_stride;
    const int vcol = a.v_col0 + kvh * a.kv_stride;
    if (active)
        for (int d = lane; d < HD; d += 32) qs[warp][d] = a.q[(row_base + qi) * a.ldq + qcol + d];
    __syncwarp();
    if (!active) return;
    const int kv_end = a.causal ? qi + 1 : len;
    float m = -INFINITY, l = 0.f;
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    for (int j = lane; j < kv_end; j += 32) {
        const float* kr = a.k + (row_base + j) * a.ldk + kcol;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 kv = *reinterpret_cast<const float4*>(kr + d);
            s = fmaf(qs[warp][d], kv.x, s);
            s = fmaf(qs[warp][d + 1], kv.y, s);
            s = fmaf(qs[warp][d + 2], kv.z, s);
            s = fmaf(qs[warp][d + 3], kv.w, s);
        }
        s *= a.scale;
        const float mn = fmaxf(m, s);
        const float corr = expf(m - mn);   // exp(-inf) = 0 on the first key
        const float p = expf(s - mn);
        l = l * corr + p;
        const float* vr = a.v + (row_base + j) * a.ldv + vcol;
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
            const float4 vv = *reinterpret_cast<const float4*>(vr + d);
            acc[d] = fmaf(p, vv.x, acc[d] * corr);
            acc[d + 1] = fmaf(p, vv.y, acc[d + 1] * corr);
            acc[d + 2] = fmaf(p, vv.z, acc[d + 2] * corr);
            acc[d + 3] = fmaf(p, vv.w, acc[d + 3] * corr);
        }
        m = mn;
    }
    // merge the lanes
    float M = m;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, s));
    const float w = (m == -INFINITY) ? 0.f : expf(m - M);
    float L = l * w;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) L += __shfl_xor_sync(0xffffffffu, L, s);
    float* orow = a.out + (row_base + qi) * a.ldo + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
        float v = acc[d] * w;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == (d & 31)) orow[d] = v / L;
    }
}

// ------------------------------------------------------------------------------------------------ head
__global__ void __launch_bounds__(128) head_finalize_kernel(const HeadF32 a) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.B) return;
    const int lane = threadIdx.x & 31;
    const int NO = a.NO, NA = a.NA;
    const int ld_s = NO + NA + 1, ld_e = NO + NA + NO + 1;
    // rewards = regression(h_last) @ reward_transform_matrix (moe_reward.py:246-247)
    float rw = 0.f;
    if (lane < NO)
        for (int k = 0; k < NO; ++k) rw = fmaf(a.rewards_raw[row * NO + k], a.transform[k * NO + lane], rw);
    if (lane < NO) a.scores[row * ld_s + lane] = rw;
    // aspect gate: softmax(logits / T) * logit_scale[0] (GatingNetwork.forward, moe_reward.py:33-36)
    float al = lane < NA ? a.aspect_logits[row * NA + lane] / a.temperature : -INFINITY;
    float mx = al;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    const float e = lane < NA ? expf(al - mx) : 0.f;
    float sum = e;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    const float gate = lane < NA ? (e / sum) * a.aspect_logit_scale[0] : 0.f;
    if (lane < NA) a.extras[row * ld_e + NO + lane] = gate;
    float cl = 0.f;
    if (lane < NO) {
        cl = a.criteria_logits[row * NO + lane];
        a.extras[row * ld_e + lane] = cl;
    }
    int crit = 0, grp = -1;
    if (lane < NO) {
        crit = a.group_items[lane];
        for (int g = 0; g < NA; ++g)
            if (lane >= a.group_offsets[g] && lane < a.group_offsets[g + 1]) grp = g;
    }
    const float cl_sh = __shfl_sync(0xffffffffu, cl, crit);
    const float myl = lane < NO ? cl_sh / a.temperature : -INFINITY;
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
        const float w = in ? (ee / ss) * a.criteria_logit_scale[0] : 0.f;
        if (in) a.extras[row * ld_e + NO + NA + lane] = w;
        float prod = in ? myr * w : 0.f;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) prod += __shfl_xor_sync(0xffffffffu, prod, s);
        const float g_gate = __shfl_sync(0xffffffffu, gate, g);
        if (lane == 0) a.scores[row * ld_s + NO + g] = prod;
        score += prod * g_gate;
        last_weighted = prod;
    }
    if (lane == 0) {
        a.scores[row * ld_s + NO + NA] = score;
        a.extras[row * ld_e + NO + NA + NO] = last_weighted;
    }
}

__global__ void __launch_bounds__(128)
gating_softmax_kernel(const float* __restrict__ logits, int n, float temperature, const float* __restrict__ logit_scale,
                      float* __restrict__ out, int B) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B) return;
    const int lane = threadIdx.x & 31;
    const float x = lane < n ? logits[row * n + lane] / temperature : -INFINITY;
    float mx = x;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    const float e = lane < n ? expf(x - mx) : 0.f;
    float sum = e;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    if (lane < n) out[row * n + lane] = (e / sum) * logit_scale[0];
}

inline int grid_for(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = 148LL * 16;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int sgemm_tn(const float* A, long long lda, const float* W, long long ldw, float* C, long long ldc, long long M, int N,
             int K, const float* bias, int epi, const float* res, long long ldr, const float* scale, cudaStream_t st) {
    if (M == 0 || N == 0) return 0;
    MJV_CHECK(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0, MJV_ERR_ARG, "fp32 gemm: K/lda/ldw must be multiples of 4 (K=%d)", K);
    dim3 grid((N + BN - 1) / BN, static_cast<unsigned>((M + BM - 1) / BM));
    MJV_CHECK(grid.y <= 65535, MJV_ERR_ARG, "fp32 gemm: M=%lld too large for the verification mode", M);
#define MJV_SG(E) sgemm_tn_kernel<E><<<grid, 256, 0, st>>>(A, lda, W, ldw, C, ldc, M, N, K, bias, res, ldr, scale)
    switch (epi) {
        case EPI_NONE: MJV_SG(EPI_NONE); break;
        case EPI_GELU: MJV_SG(EPI_GELU); break;
        case EPI_RELU: MJV_SG(EPI_RELU); break;
        case EPI_RESID: MJV_SG(EPI_RESID); break;
        case EPI_RESID_SCALE: MJV_SG(EPI_RESID_SCALE); break;
        default: return set_error(MJV_ERR_ARG, "fp32 gemm: unknown epilogue %d", epi);
    }
#undef MJV_SG
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int layernorm(const float* x, const float* w, const float* b, float* out, long long rows, int C, float eps, cudaStream_t st) {
    if (rows == 0) return 0;
    layernorm_kernel<<<static_cast<unsigned>(rows), 256, 0, st>>>(x, w, b, out, rows, C, eps);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int rmsnorm(const float* x, const float* w, float* out, long long rows, int C, float eps, const int* row_idx, cudaStream_t st) {
    if (rows == 0) return 0;
    rmsnorm_kernel<<<static_cast<unsigned>(rows), 256, 0, st>>>(x, w, out, rows, C, eps, row_idx);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int patchify(const void* px, int px_dtype, float* out, int T, int img, int ps, cudaStream_t st) {
    if (T == 0) return 0;
    const long long total = static_cast<long long>(T) * (img / ps) * (img / ps) * 3 * ps * ps;
    if (px_dtype == MJV_DTYPE_F32) patchify_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)px, out, T, img, ps);
    else if (px_dtype == MJV_DTYPE_BF16)
        patchify_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>((const __nv_bfloat16*)px, out, T, img, ps);
    else return set_error(MJV_ERR_ARG, "fp32 mode: pixel dtype %d unsupported", px_dtype);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int vit_assemble(const float* conv, const float* cls, const float* pos, float* hidden, int T, int P, int C, cudaStream_t st) {
    if (T == 0) return 0;
    vit_assemble_kernel<<<grid_for(static_cast<long long>(T) * (P + 1) * C), 256, 0, st>>>(conv, cls, pos, hidden, T, P, C);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int pixel_shuffle(const float* hidden, float* out, int T, int g, int C, int ps_v1, cudaStream_t st) {
    if (T == 0) return 0;
    pixel_shuffle_kernel<<<grid_for(static_cast<long long>(T) * g * g * C), 256, 0, st>>>(hidden, out, T, g, C, ps_v1);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int embed(const int* ids, const float* table, float* out, long long rows, int C, int skip_id, cudaStream_t st) {
    if (rows == 0) return 0;
    embed_kernel<<<grid_for(rows * C), 256, 0, st>>>(ids, table, out, rows, C, skip_id);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int scatter_rows(const float* src, const int* row_map, float* dst, long long n, int C, cudaStream_t st) {
    if (n == 0) return 0;
    move_rows_kernel<true><<<grid_for(n * C), 256, 0, st>>>(src, row_map, dst, n, C);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gather_rows(const float* src, const int* idx, float* dst, long long n, int C, cudaStream_t st) {
    if (n == 0) return 0;
    move_rows_kernel<false><<<grid_for(n * C), 256, 0, st>>>(src, idx, dst, n, C);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int rope(float* x, long long ld, long long rows, int heads, int group, int kv_stride, int col0, int hd, const float* cos_t,
         const float* sin_t, int table_ld, const int* row_pos, cudaStream_t st) {
    if (rows == 0) return 0;
    rope_kernel<<<grid_for(rows * heads * (hd / 2)), 256, 0, st>>>(x, ld, rows, heads, group, kv_stride, col0, hd, cos_t, sin_t,
                                                                   table_ld, row_pos);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int attention(const AttnF32& a, cudaStream_t st) {
    if (a.nseq == 0) return 0;
    MJV_CHECK(a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 && a.k_col0 % 4 == 0 && a.v_col0 % 4 == 0 && a.kv_stride % 4 == 0,
              MJV_ERR_ARG, "fp32 attention: strides must be multiples of 4");
    dim3 grid((a.seq_len + 3) / 4, a.heads, a.nseq);   // seq_len = longest sequence in varlen mode
    if (a.hd == 64) attention_kernel<64><<<grid, 128, 0, st>>>(a);
    else if (a.hd == 96) attention_kernel<96><<<grid, 128, 0, st>>>(a);
    else if (a.hd == 128) attention_kernel<128><<<grid, 128, 0, st>>>(a);
    else return set_error(MJV_ERR_ARG, "fp32 attention: head_dim %d unsupported (64, 96, 128)", a.hd);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int silu_mul(const float* gate, const float* up, float* out, long long n, cudaStream_t st) {
    if (n == 0) return 0;
    silu_mul_kernel<<<grid_for(n), 256, 0, st>>>(gate, up, out, n);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int to_f32(const void* src, int dtype, float* dst, long long n, cudaStream_t st) {
    if (n == 0) return 0;
    to_f32_kernel<<<grid_for(n), 256, 0, st>>>(src, dtype, dst, n);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int head_finalize(const HeadF32& a, cudaStream_t st) {
    if (a.B == 0) return 0;
    MJV_CHECK(a.NO <= 32 && a.NA <= 32, MJV_ERR_ARG, "head: at most 32 criteria / aspects");
    head_finalize_kernel<<<(a.B + 3) / 4, 128, 0, st>>>(a);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gating_softmax(const float* logits, int n, float temperature, const float* logit_scale, float* out, int B, cudaStream_t st) {
    if (B == 0) return 0;
    MJV_CHECK(n >= 1 && n <= 32, MJV_ERR_ARG, "gating softmax: 1..32 outputs supported, got %d", n);
    gating_softmax_kernel<<<(B + 3) / 4, 128, 0, st>>>(logits, n, temperature, logit_scale, out, B);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace f32
}  // namespace mjv
