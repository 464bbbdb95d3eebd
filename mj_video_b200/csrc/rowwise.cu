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

// Sum over the W warps that share a row (W = 1: plain warp reduction). Every thread of the block must call it.
template <int W>
__device__ __forceinline__ float row_sum(float v, float* red, int warp_in_block) {
    v = warp_sum(v);
    if (W == 1) return v;
    // the W warps of one row synchronise among themselves (named barrier 1 + row-in-block), not with the other rows of the
    // block: a block-wide __syncthreads made every row wait for the slowest of the 4-8 rows four times per LayerNorm
    // (only for W >= 4: with two warps per row the named barrier measured 3-7 % SLOWER than __syncthreads on RMSNorm)
    const int bar_id = 1 + warp_in_block / W;
    if (W >= 4) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(W * 32) : "memory");  // previous use of `red` is over
    else __syncthreads();
    if ((threadIdx.x & 31) == 0) red[warp_in_block] = v;
    if (W >= 4) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(W * 32) : "memory");
    else __syncthreads();
    float t = 0.f;
    const int base = warp_in_block / W * W;
#pragma unroll
    for (int i = 0; i < W; ++i) t += red[base + i];
    return t;
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// out[r,:] = bf16((x - mean) * rstd * w + b) over C columns (nn.LayerNorm: modeling_intern_vit.py:291,293; mlp1.0
// modeling_internvl_chat.py:136). GATHER: the row is assembled from 4 source rows of C/4 columns (pixel_shuffle v2,
// modeling_internvl_chat.py:228-262): out row (t, i, j), column b*C/2 + a*C/4 + c  <-  x[t, 1 + (2i+b)*G + (2j+a), c].
// W warps share a row (wide rows keep <= 4 16-byte chunks per thread, so many blocks stay resident); the row is held
// packed (bf16) in registers between the passes.
template <int NCHUNK, int W, bool GATHER>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                 const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out, long long rows, int C, float eps,
                 long long ldx, int grid_g, int seq) {
    __shared__ float red[kWarpsPerBlock];
    const int wib = threadIdx.x >> 5;
    const long long row = static_cast<long long>(blockIdx.x) * (kWarpsPerBlock / W) + wib / W;
    const int wsub = wib % W;
    const bool active = row < rows;
    const int lane = threadIdx.x & 31;
    uint4 u[NCHUNK];
    long long g_base = 0;
    int cq = 1;
    if (GATHER && active) {
        const int half = grid_g / 2;
        const int per_tile = half * half;
        const long long t = row / per_tile;
        const int ij = static_cast<int>(row - t * per_tile);
        int i = ij / half, j = ij - i * half;
        if (seq < 0) {  // ps_version 'v1': the output grid is the transpose of 'v2' (no final permute in the reference)
            const int tmp = i;
            i = j;
            j = tmp;
        }
        cq = C / 4;
        const int seq_abs = seq < 0 ? -seq : seq;
        g_base = t * seq_abs + 1 + static_cast<long long>(2 * i) * grid_g + 2 * j;
    }
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c = (ch * W + wsub) * 256 + lane * 8;
        u[ch] = make_uint4(0u, 0u, 0u, 0u);
        if (active && c < C) {
            const __nv_bfloat16* src;
            if (GATHER) {
                const int blk = c / cq;  // b*2 + a
                src = x + (g_base + (blk >> 1) * grid_g + (blk & 1)) * ldx + (c - blk * cq);
            } else {
                src = x + row * ldx + c;
            }
            u[ch] = *reinterpret_cast<const uint4*>(src);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        float f[8];
        unpack8(u[ch], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += f[e];  // chunks beyond C hold zeros
    }
    const float mean = row_sum<W>(sum, red, wib) / static_cast<float>(C);
    float sq = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        if ((ch * W + wsub) * 256 + lane * 8 < C) {
            float f[8];
            unpack8(u[ch], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float dlt = f[e] - mean;
                sq += dlt * dlt;
            }
        }
    }
    const float rstd = rsqrtf(row_sum<W>(sq, red, wib) / static_cast<float>(C) + eps);
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c = (ch * W + wsub) * 256 + lane * 8;
        if (active && c < C) {
            float f[8], wf[8], bf[8], o[8];
            unpack8(u[ch], f);
            unpack8(__ldg(reinterpret_cast<const uint4*>(w + c)), wf);
            unpack8(__ldg(reinterpret_cast<const uint4*>(b + c)), bf);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = (f[e] - mean) * rstd * wf[e] + bf[e];
            *reinterpret_cast<uint4*>(out + row * C + c) = pack8(o);
        }
    }
}

// ------------------------------------------------------------------------------------------------ RMSNorm
// InternLM2RMSNorm (modeling_internlm2.py:138-143): y = bf16(x * rsqrt(mean(x^2) + eps)); out = bf16(w * y).
// row_idx (optional) gathers source rows, so the final norm runs only on the 2 rows per sample the head reads.
template <int NCHUNK, int W>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
               __nv_bfloat16* __restrict__ out, long long rows, int C, float eps, const int* __restrict__ row_idx) {
    __shared__ float red[kWarpsPerBlock];
    const int wib = threadIdx.x >> 5;
    const long long row = static_cast<long long>(blockIdx.x) * (kWarpsPerBlock / W) + wib / W;
    const int wsub = wib % W;
    const bool active = row < rows;
    const int lane = threadIdx.x & 31;
    const long long srow = active ? (row_idx != nullptr ? row_idx[row] : row) : 0;
    uint4 u[NCHUNK];  // the row stays packed (bf16) in registers between the two passes
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c = (ch * W + wsub) * 256 + lane * 8;
        u[ch] = (active && c < C) ? *reinterpret_cast<const uint4*>(x + srow * C + c) : make_uint4(0u, 0u, 0u, 0u);
    }
    float sq = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        float f[8];
        unpack8(u[ch], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) sq += f[e] * f[e];
    }
    const float rstd = rsqrtf(row_sum<W>(sq, red, wib) / static_cast<float>(C) + eps);
#pragma unroll
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const int c = (ch * W + wsub) * 256 + lane * 8;
        if (active && c < C) {
            float f[8], wf[8], o[8];
            unpack8(u[ch], f);
            unpack8(__ldg(reinterpret_cast<const uint4*>(w + c)), wf);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = wf[e] * ptx::bf16_round(f[e] * rstd);
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
// One block = one row of patches of one tile. The 3 x ps image rows it covers are read with coalesced 16-byte loads and
// scattered into shared memory ALREADY in im2col order ([patch][k], 4-byte stores of the even-aligned element pairs; the
// tile's row stride is padded so that the patches do not alias the same banks); every thread then moves 8 consecutive k of
// one patch with one 16-byte shared load and one 16-byte global store. (The first version staged the image rows as they
// are and gathered 2-byte elements: 8 bank-conflicted loads per thread kept the LSU busier than HBM, 4.2 TB/s.)
template <typename TIn>
__global__ void __launch_bounds__(512)
im2col_kernel(const TIn* __restrict__ px, __nv_bfloat16* __restrict__ out, int T, int img, int ps, int Kpad, int kstride) {
    extern __shared__ __align__(16) unsigned char im2col_smem[];
    __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(im2col_smem);  // [g][kstride], kstride >= Kpad
    const int g = img / ps;
    const int K = 3 * ps * ps;
    const int kvec = Kpad / 8;
    const int t = blockIdx.x / g, pi = blockIdx.x - t * g;
    // zero the pad columns K .. Kpad of every patch (K and Kpad are even)
    {
        const int padw = (Kpad - K) >> 1;
        uint32_t* t32 = reinterpret_cast<uint32_t*>(tile);
        for (int i = threadIdx.x; i < g * padw; i += blockDim.x) {
            const int pj = i / padw, w = i - pj * padw;
            t32[(pj * kstride + K) / 2 + w] = 0u;
        }
    }
    constexpr int VEC = 16 / sizeof(TIn);  // elements per 16-byte load (8 or 4: even, so pairs never straddle a patch)
    const int vec_per_row = img / VEC;
    const int n_vec = 3 * ps * vec_per_row;
    for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
        const int r = i / vec_per_row, v = i - r * vec_per_row;  // r = c*ps + ky
        const int c = r / ps, ky = r - c * ps;
        const TIn* src = px + ((static_cast<size_t>(t) * 3 + c) * img + pi * ps + ky) * img + v * VEC;
        const uint4 raw = *reinterpret_cast<const uint4*>(src);
        uint32_t pairs[VEC / 2];
        if (sizeof(TIn) == 2) {
            pairs[0] = raw.x;
            pairs[1] = raw.y;
            if (VEC == 8) {
                pairs[VEC / 2 - 2] = raw.z;
                pairs[VEC / 2 - 1] = raw.w;
            }
        } else {
            const float* f = reinterpret_cast<const float*>(&raw);
            pairs[0] = ptx::pack_bf16x2(f[0], f[1]);
            pairs[1] = ptx::pack_bf16x2(f[2], f[3]);
        }
        const int x0 = v * VEC;
#pragma unroll
        for (int q = 0; q < VEC / 2; ++q) {
            const int x = x0 + 2 * q;            // even; ps is even, so x and x + 1 lie in the same patch
            const int pj = x / ps, kx = x - pj * ps;
            *reinterpret_cast<uint32_t*>(tile + pj * kstride + r * ps + kx) = pairs[q];
        }
    }
    __syncthreads();
    const int kv = threadIdx.x % kvec;
    for (int pj = threadIdx.x / kvec; pj < g; pj += blockDim.x / kvec) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile + pj * kstride + kv * 8);
        const size_t p = (static_cast<size_t>(t) * g + pi) * g + pj;
        *reinterpret_cast<uint4*>(out + p * Kpad + kv * 8) = v;
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

// rows up to 1024 columns: one warp per row; wider rows: 2 or 4 warps per row with <= 4 chunks per thread
template <bool GATHER>
int launch_ln(const __nv_bfloat16* x, const __nv_bfloat16* w, const __nv_bfloat16* b, __nv_bfloat16* out,
              long long rows, int C, float eps, long long ldx, int grid_g, int seq, cudaStream_t st) {
    MJV_CHECK(C % 8 == 0 && C <= 4096, MJV_ERR_ARG, "layernorm: C=%d unsupported", C);
    const int nchunk = (C + 255) / 256;
    // must match the W of the instantiation chosen below. (Two warps x 8 chunks for the 4096-wide projector LayerNorm
    // measured slower than four warps x 4 chunks, alone (4.10 vs 4.52 TB/s) and in the step; eight warps x 2 chunks: 4.19.)
    const int wpr = nchunk <= 4 ? 1 : (nchunk <= 8 ? 2 : 4);
    const int rows_per_block = kWarpsPerBlock / wpr;
    const unsigned grid = static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block);
    const int threads = kWarpsPerBlock * 32;
    LaunchScope scope(GATHER ? TAG_PIXSHUF_LN : TAG_LAYERNORM, st, 4.0 * rows * C);
#define MJV_LN(NC, W) layernorm_kernel<NC, W, GATHER><<<grid, threads, 0, st>>>(x, w, b, out, rows, C, eps, ldx, grid_g, seq)
    if (nchunk <= 1) MJV_LN(1, 1);
    else if (nchunk <= 2) MJV_LN(2, 1);
    else if (nchunk <= 4) MJV_LN(4, 1);
    else if (nchunk <= 8) MJV_LN(4, 2);
    else MJV_LN(4, 4);
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
                            int c_vit, float eps, int ps_v1, cudaStream_t st) {
    if (tiles == 0) return 0;
    int g = 1;
    while (g * g < seq - 1) ++g;
    MJV_CHECK(g * g == seq - 1 && g % 2 == 0, MJV_ERR_ARG, "pixel_shuffle: seq=%d is not 1 + (even grid)^2", seq);
    MJV_CHECK((c_vit % 8) == 0, MJV_ERR_ARG, "pixel_shuffle: channels %d", c_vit);
    const long long rows = static_cast<long long>(tiles) * (g / 2) * (g / 2);
    return launch_ln<true>((const __nv_bfloat16*)vit_hidden, (const __nv_bfloat16*)w, (const __nv_bfloat16*)b,
                           (__nv_bfloat16*)out, rows, 4 * c_vit, eps, c_vit, g, ps_v1 ? -seq : seq, st);
}

int rmsnorm(const void* x, const void* w, void* out, long long rows, int C, float eps, const int* row_idx,
            cudaStream_t st) {
    if (rows == 0) return 0;
    MJV_CHECK(C % 8 == 0 && C <= 4096, MJV_ERR_ARG, "rmsnorm: C=%d unsupported", C);
    const int nchunk = (C + 255) / 256;
    // two warps per row up to C = 4096 with 4-8 16-byte loads in flight per lane: at C = 3072 (Phi-3 / InternVL2-4B) the
    // four-warp split left 3 loads per lane in flight and ran at 2.47 TB/s (profiles/r01)
    const int wpr = nchunk <= 4 ? 1 : 2;   // must match the W of the instantiation chosen below
    const int rows_per_block = kWarpsPerBlock / wpr;
    const unsigned grid = static_cast<unsigned>((rows + rows_per_block - 1) / rows_per_block);
    const int threads = kWarpsPerBlock * 32;
    LaunchScope scope(TAG_RMSNORM, st, 4.0 * rows * C);
#define MJV_RMS(NC, W) \
    rmsnorm_kernel<NC, W><<<grid, threads, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (__nv_bfloat16*)out, rows, C, eps, row_idx)
    if (nchunk <= 1) MJV_RMS(1, 1);
    else if (nchunk <= 2) MJV_RMS(2, 1);
    else if (nchunk <= 4) MJV_RMS(4, 1);
    else if (nchunk <= 8) MJV_RMS(4, 2);
    else if (nchunk <= 12) MJV_RMS(6, 2);
    else MJV_RMS(8, 2);
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
    MJV_CHECK(img % ps == 0 && Kpad >= 3 * ps * ps && Kpad % 8 == 0 && img % 8 == 0, MJV_ERR_ARG, "im2col: bad geometry");
    MJV_CHECK((reinterpret_cast<uintptr_t>(px) & 15) == 0, MJV_ERR_ARG, "im2col: pixel_values must be 16-byte aligned");
    const int g_ = img / ps;
    const int grid = T * g_;
    MJV_CHECK(ps % 2 == 0, MJV_ERR_ARG, "im2col: patch size %d must be even", ps);
    const int kstride = Kpad + 8;   // +16 bytes: a patch row keeps its 16-byte alignment and successive patches shift by 4 banks
    const size_t smem = static_cast<size_t>(g_) * kstride * sizeof(__nv_bfloat16);
    MJV_CHECK(smem <= 200 * 1024, MJV_ERR_ARG, "im2col: one row of patches does not fit in shared memory");
    const int kvec = Kpad / 8;
    MJV_CHECK(kvec <= 512, MJV_ERR_ARG, "im2col: Kpad %d too large", Kpad);
    const int threads = kvec * (kvec >= 384 ? 1 : 384 / kvec);  // a multiple of kvec, ~384 threads
    static PerDeviceOnce once;
    if (once.first()) {
        MJV_CUDA_OK(cudaFuncSetAttribute(im2col_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        MJV_CUDA_OK(cudaFuncSetAttribute(im2col_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    LaunchScope scope(TAG_IM2COL, st, (double)T * (3.0 * img * img * (px_dtype == MJV_DTYPE_F32 ? 4 : 2) + 2.0 * g_ * g_ * Kpad));
    if (px_dtype == MJV_DTYPE_BF16)
        im2col_kernel<__nv_bfloat16><<<grid, threads, smem, st>>>((const __nv_bfloat16*)px, (__nv_bfloat16*)out, T, img, ps, Kpad, kstride);
    else if (px_dtype == MJV_DTYPE_F32)
        im2col_kernel<float><<<grid, threads, smem, st>>>((const float*)px, (__nv_bfloat16*)out, T, img, ps, Kpad, kstride);
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
                                           int32_t tiles, int32_t seq, int32_t c_vit, float eps, int32_t ps_v1,
                                           void* stream) {
    return mjv::pixel_shuffle_layernorm(vit_hidden, w, b, out, tiles, seq, c_vit, eps, ps_v1,
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
