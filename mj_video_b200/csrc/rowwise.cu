// HBM-bound row-wise kernels: LayerNorm, pixel-shuffle+LayerNorm gather, RMSNorm (+row gather), token-embedding gather,
// im2col for the patch-embed GEMM and the CLS rows. One warp per row, 16-byte vector accesses, fp32 statistics.
// Rounding points mirror the reference's bf16 eager path.
#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace mjv {
namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    f[0] = ptx::bf16_lo(u.x); f[1] = ptx::bf16_hi(u.x);
    f[2] = ptx::bf16_lo(u.y); f[3] = ptx::bf16_hi(u.y);
    f[4] = ptx::bf16_lo(u.z); f[5] = ptx::bf16_hi(u.z);
    f[6] = ptx::bf16_lo(u.w); f[7] = ptx::bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    return make_uint4(ptx::pack_bf16x2(f[0], f[1]), ptx::pack_bf16x2(f[2], f[3]), ptx::pack_bf16x2(f[4], f[5]),
                      ptx::pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// out[r,:] = bf16((x - mean) * rstd * w + b) over C columns (nn.LayerNorm: modeling_intern_vit.py:291,293; mlp1.0
// modeling_internvl_chat.py:136). GATHER: the row is assembled from 4 source rows of C/4 columns (pixel_shuffle v2,
// modeling_internvl_chat.py:228-262): out row (t, i, j), column b*C/2 + a*C/4 + c  <-  x[t, 1 + (2i+b)*G + (2j+a), c].
template <int NCHUNK, bool GATHER>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                 const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out, long long rows, int C, float eps,
                 long long ldx, int grid_g, int seq) {
    const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float v[NCHUNK][8];
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c = ch * 256 + lane * 8;
        if (c < C) {
            const __nv_bfloat16* src;
            if (GATHER) {
                const int half = grid_g / 2;
                const int per_tile = half * half;
                const long long t = row / per_tile;
                const int ij = static_cast<int>(row - t * per_tile);
                const int i = ij / half, j = ij - i * half;
                const int cq = C / 4;
                const int blk = c / cq;  // b*2 + a
                const int cc = c - blk * cq;
                const int bb = blk >> 1, aa = blk & 1;
                src = x + (t * seq + 1 + (2 * i + bb) * grid_g + (2 * j + aa)) * ldx + cc;
            } else {
                src = x + row * ldx + c;
            }
            const uint4 u = *reinterpret_cast<const uint4*>(src);
            unpack8(u, v[ch]);
#pragma unroll
            for (int e = 0; e < 8; ++e) sum += v[ch][e];
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[ch][e] = 0.f;
        }
    }
    const float mean = warp_sum(sum) / static_cast<float>(C);
    float sq = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c = ch * 256 + lane * 8;
        if (c < C) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float dlt = v[ch][e] - mean;
                sq += dlt * dlt;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(C) + eps);
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c = ch * 256 + lane * 8;
        if (c < C) {
            float wf[8], bf[8], o[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(w + c)), wf);
            unpack8(__ldg(reinterpret_cast<const uint4*>(b + c)), bf);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (v[ch][e] - mean) * rstd * wf[e] + bf[e];
            *reinterpret_cast<uint4*>(out + row * C + c) = pack8(o);
        }
    }
}

// ------------------------------------------------------------------------------------------------ RMSNorm
// InternLM2RMSNorm (modeling_internlm2.py:138-143): y = bf16(x * rsqrt(mean(x^2) + eps)); out = bf16(w * y).
// row_idx (optional) gathers source rows, so the final norm runs only on the 2 rows per sample the head reads.
template <int NCHUNK>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
               __nv_bfloat16* __restrict__ out, long long rows, int C, float eps, const int* __restrict__ row_idx) {
    const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const long long srow = row_idx != nullptr ? row_idx[row] : row;
    float v[NCHUNK][8];
    float sq = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c = ch * 256 + lane * 8;
        if (c < C) {
            unpack8(*reinterpret_cast<const uint4*>(x + srow * C + c), v[ch]);
#pragma unroll
            for (int e = 0; e < 8; ++e) sq += v[ch][e] * v[ch][e];
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(C) + eps);
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c = ch * 256 + lane * 8;
        if (c < C) {
            float wf[8], o[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(w + c)), wf);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = wf[e] * ptx::bf16_round(v[ch][e] * rstd);
            *reinterpret_cast<uint4*>(out + row * C + c) = pack8(o);
        }
    }
}

// ------------------------------------------------------------------------------------------------ token embeddings
// tok_embeddings(input_ids) for the text rows (modeling_internvl_chat.py:163); rows holding <IMG_CONTEXT> are skipped,
// they are written by the projector GEMM's scatter epilogue (the splice of :176-188).
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
embed_kernel(const int* __restrict__ ids, const __nv_bfloat16* __restrict__ table, __nv_bfloat16* __restrict__ out,
             long long rows, int C, int img_ctx_id, int vocab) {
    const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int id = ids[row];
    if (id == img_ctx_id || id < 0 || id >= vocab) return;
    const int lane = threadIdx.x & 31;
    const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<long long>(id) * C);
    uint4* dst = reinterpret_cast<uint4*>(out + row * C);
    for (int c = lane; c < C / 8; c += 32) dst[c] = __ldg(src + c);
}

// dst[i,:] = src[idx[i],:] (bf16 rows, C multiple of 8)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
gather_rows_kernel(const __nv_bfloat16* __restrict__ src, const int* __restrict__ idx, __nv_bfloat16* __restrict__ dst,
                   int rows, int C) {
    const int row = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<long long>(idx[row]) * C);
    uint4* d = reinterpret_cast<uint4*>(dst + static_cast<long long>(row) * C);
    for (int c = lane; c < C / 8; c += 32) d[c] = s[c];
}

// ------------------------------------------------------------------------------------------------ patch im2col
// pixel_values [T,3,H,W] (bf16 or fp32) -> A[T*P, Kpad] bf16 with k = c*ps*ps + ky*ps + kx (the flattening of
// nn.Conv2d's weight, modeling_intern_vit.py:164), zero padded to Kpad so rows are 16-byte aligned for TMA.
template <typename TIn>
__global__ void __launch_bounds__(256)
im2col_kernel(const TIn* __restrict__ px, __nv_bfloat16* __restrict__ out, int T, int img, int ps, int Kpad) {
    // one thread = 8 consecutive k of one patch row of A (one 16-byte store); 32-bit index arithmetic throughout
    const int g = img / ps;
    const int K = 3 * ps * ps;
    const int kvec = Kpad / 8;
    const unsigned total = static_cast<unsigned>(T) * g * g * kvec;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned p = i / kvec;          // patch index over all tiles
        const int k0 = (i - p * kvec) * 8;
        const unsigned tp = p / g;
        const int pj = p - tp * g;
        const unsigned t = tp / g;
        const int pi = tp - t * g;
        const TIn* base = px + (static_cast<size_t>(t) * 3 * img + pi * ps) * img + pj * ps;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = k0 + e;
            float val = 0.f;
            if (k < K) {
                const int c = k / (ps * ps);
                const int rem = k - c * ps * ps;
                const int ky = rem / ps, kx = rem - ky * ps;
                val = static_cast<float>(base[(static_cast<size_t>(c) * img + ky) * img + kx]);
            }
            f[e] = val;
        }
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(p) * Kpad + k0) = pack8(f);
    }
}

// CLS rows: hidden[t*S + 0, :] = bf16(cls + pos[0]) (modeling_intern_vit.py:166-173)
__global__ void cls_rows_kernel(const __nv_bfloat16* __restrict__ cls, const __nv_bfloat16* __restrict__ pos,
                                __nv_bfloat16* __restrict__ hidden, int T, int S, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * C) return;
    const int t = i / C, c = i - t * C;
    hidden[static_cast<long long>(t) * S * C + c] = __float2bfloat16_rn(__bfloat162float(cls[c]) + __bfloat162float(pos[c]));
}

template <bool GATHER>
int launch_ln(const __nv_bfloat16* x, const __nv_bfloat16* w, const __nv_bfloat16* b, __nv_bfloat16* out,
              long long rows, int C, float eps, long long ldx, int grid_g, int seq, cudaStream_t st) {
    MJV_CHECK(C % 8 == 0 && C <= 4096, MJV_ERR_ARG, "layernorm: C=%d unsupported", C);
    const int nchunk = (C + 255) / 256;
    const unsigned grid = static_cast<unsigned>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const int threads = kWarpsPerBlock * 32;
    LaunchScope scope(GATHER ? TAG_PIXSHUF_LN : TAG_LAYERNORM, st, 4.0 * rows * C);
#define MJV_LN(NC) layernorm_kernel<NC, GATHER><<<grid, threads, 0, st>>>(x, w, b, out, rows, C, eps, ldx, grid_g, seq)
    if (nchunk <= 1) MJV_LN(1);
    else if (nchunk <= 2) MJV_LN(2);
    else if (nchunk <= 4) MJV_LN(4);
    else if (nchunk <= 8) MJV_LN(8);
    else MJV_LN(16);
#undef MJV_LN
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

int layernorm(const void* x, const void* w, const void* b, void* out, long long rows, int C, float eps,
              cudaStream_t st) {
    if (rows == 0) return 0;
    return launch_ln<false>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (const __nv_bfloat16*)b,
                            (__nv_bfloat16*)out, rows, C, eps, C, 0, 0, st);
}

int pixel_shuffle_layernorm(const void* vit_hidden, const void* w, const void* b, void* out, int tiles, int seq,
                            int c_vit, float eps, cudaStream_t st) {
    if (tiles == 0) return 0;
    int g = 1;
    while (g * g < seq - 1) ++g;
    MJV_CHECK(g * g == seq - 1 && g % 2 == 0, MJV_ERR_ARG, "pixel_shuffle: seq=%d is not 1 + (even grid)^2", seq);
    MJV_CHECK((c_vit % 8) == 0, MJV_ERR_ARG, "pixel_shuffle: channels %d", c_vit);
    const long long rows = static_cast<long long>(tiles) * (g / 2) * (g / 2);
    return launch_ln<true>((const __nv_bfloat16*)vit_hidden, (const __nv_bfloat16*)w, (const __nv_bfloat16*)b,
                           (__nv_bfloat16*)out, rows, 4 * c_vit, eps, c_vit, g, seq, st);
}

int rmsnorm(const void* x, const void* w, void* out, long long rows, int C, float eps, const int* row_idx,
            cudaStream_t st) {
    if (rows == 0) return 0;
    MJV_CHECK(C % 8 == 0 && C <= 4096, MJV_ERR_ARG, "rmsnorm: C=%d unsupported", C);
    const int nchunk = (C + 255) / 256;
    const unsigned grid = static_cast<unsigned>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const int threads = kWarpsPerBlock * 32;
    LaunchScope scope(TAG_RMSNORM, st, 4.0 * rows * C);
#define MJV_RMS(NC) \
    rmsnorm_kernel<NC><<<grid, threads, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (__nv_bfloat16*)out, rows, C, eps, row_idx)
    if (nchunk <= 1) MJV_RMS(1);
    else if (nchunk <= 2) MJV_RMS(2);
    else if (nchunk <= 4) MJV_RMS(4);
    else if (nchunk <= 8) MJV_RMS(8);
    else MJV_RMS(16);
#undef MJV_RMS
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int gather_rows(const void* src, const int* idx, void* dst, int rows, int C, cudaStream_t st) {
    if (rows == 0) return 0;
    MJV_CHECK(C % 8 == 0, MJV_ERR_ARG, "gather_rows: C=%d", C);
    {
        LaunchScope scope(TAG_EMBED, st, 4.0 * rows * C);
        gather_rows_kernel<<<(rows + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, st>>>(
            (const __nv_bfloat16*)src, idx, (__nv_bfloat16*)dst, rows, C);
    }
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int embed_tokens(const int* ids, const void* table, void* out, long long rows, int C, int img_ctx_id, int vocab,
                 cudaStream_t st) {
    if (rows == 0) return 0;
    MJV_CHECK(C % 8 == 0, MJV_ERR_ARG, "embed: C=%d", C);
    const unsigned grid = static_cast<unsigned>((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
    {
        LaunchScope scope(TAG_EMBED, st, 4.0 * rows * C);
        embed_kernel<<<grid, kWarpsPerBlock * 32, 0, st>>>(ids, (const __nv_bfloat16*)table, (__nv_bfloat16*)out, rows, C,
                                                            img_ctx_id, vocab);
    }
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int im2col_patches(const void* px, int px_dtype, void* out, int T, int img, int ps, int Kpad, cudaStream_t st) {
    if (T == 0) return 0;
    MJV_CHECK(img % ps == 0 && Kpad >= 3 * ps * ps && Kpad % 8 == 0, MJV_ERR_ARG, "im2col: bad geometry");
    const int grid = sm_count() * 8;
    const int g_ = img / ps;
    LaunchScope scope(TAG_IM2COL, st, (double)T * (3.0 * img * img * (px_dtype == MJV_DTYPE_F32 ? 4 : 2) + 2.0 * g_ * g_ * Kpad));
    if (px_dtype == MJV_DTYPE_BF16)
        im2col_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)px, (__nv_bfloat16*)out, T, img, ps, Kpad);
    else if (px_dtype == MJV_DTYPE_F32)
        im2col_kernel<float><<<grid, 256, 0, st>>>((const float*)px, (__nv_bfloat16*)out, T, img, ps, Kpad);
    else
        return set_error(MJV_ERR_ARG, "im2col: unsupported pixel dtype %d", px_dtype);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

int cls_rows(const void* cls, const void* pos, void* hidden, int T, int S, int C, cudaStream_t st) {
    if (T == 0) return 0;
    {
        LaunchScope scope(TAG_CLS_ROWS, st, 2.0 * T * C);
        cls_rows_kernel<<<(T * C + 255) / 256, 256, 0, st>>>((const __nv_bfloat16*)cls, (const __nv_bfloat16*)pos,
                                                             (__nv_bfloat16*)hidden, T, S, C);
    }
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace mjv

// ---------------------------------------------------------------------------------------------- C ABI (test hooks)
extern "C" int mjv_layernorm(const void* x, const void* w, const void* b, void* out, int64_t rows, int32_t C,
                             float eps, void* stream) {
    return mjv::layernorm(x, w, b, out, rows, C, eps, static_cast<cudaStream_t>(stream));
}
extern "C" int mjv_pixel_shuffle_layernorm(const void* vit_hidden, const void* w, const void* b, void* out,
                                           int32_t tiles, int32_t seq, int32_t c_vit, float eps, void* stream) {
    return mjv::pixel_shuffle_layernorm(vit_hidden, w, b, out, tiles, seq, c_vit, eps,
                                        static_cast<cudaStream_t>(stream));
}
extern "C" int mjv_rmsnorm(const void* x, const void* w, void* out, int64_t rows, int32_t C, float eps,
                           const int32_t* row_idx, void* stream) {
    return mjv::rmsnorm(x, w, out, rows, C, eps, row_idx, static_cast<cudaStream_t>(stream));
}
extern "C" int mjv_embed_tokens(const int32_t* ids, const void* table, void* out, int64_t rows, int32_t C,
                                int32_t img_ctx_id, int32_t vocab, void* stream) {
    return mjv::embed_tokens(ids, table, out, rows, C, img_ctx_id, vocab, static_cast<cudaStream_t>(stream));
}
extern "C" int mjv_im2col_patches(const void* px, int32_t px_dtype, void* out, int32_t T, int32_t img, int32_t ps,
                                  int32_t Kpad, void* stream) {
    return mjv::im2col_patches(px, px_dtype, out, T, img, ps, Kpad, static_cast<cudaStream_t>(stream));
}
