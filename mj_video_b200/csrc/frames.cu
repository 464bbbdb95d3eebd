// Frame path on the GPU: uint8 HWC frames -> (Pillow-exact bicubic resize to the tile grid, row-major 448² crops,
// optional thumbnail) -> ToTensor/Normalize -> CHW bf16/fp32 tiles.
// Replaces load_video's per-frame work (scripts/data_processor/data.py:158-179): Image.resize (data.py:101,115),
// crop (data.py:103-112), build_transform (data.py:56-64). Integer arithmetic is Pillow's (22-bit fixed-point taps,
// horizontal pass into a uint8 temporary, then vertical), so pixels are bit-identical; normalisation is a 3x256 LUT built
// on the host with the same fp32 div/sub/div sequence torchvision performs.
#include <math.h>

#include <algorithm>

#include <map>
#include <tuple>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace mjv {
namespace {

struct NormLut {
    float v[3 * 256];
};

NormLut make_lut() {
    // torchvision: ToTensor -> x.to(float32).div(255); Normalize -> x.sub_(mean).div_(std), mean/std as fp32 tensors
    const float mean[3] = {0.485f, 0.456f, 0.406f};  // data.py:29
    const float stdv[3] = {0.229f, 0.224f, 0.225f};  // data.py:30
    NormLut l;
    for (int c = 0; c < 3; ++c)
        for (int u = 0; u < 256; ++u) {
            volatile float x = static_cast<float>(u) / 255.0f;
            volatile float y = x - mean[c];
            volatile float z = y / stdv[c];
            l.v[c * 256 + u] = z;
        }
    return l;
}

template <typename TOut>
__device__ __forceinline__ TOut cvt_out(float x);
template <>
__device__ __forceinline__ float cvt_out<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// Identity geometry (frame already S x S, one tile): pure layout change + LUT. One thread = 16 pixels of one row.
template <typename TOut>
__global__ void __launch_bounds__(256)
normalize_tiles_kernel(const uint8_t* __restrict__ in, TOut* __restrict__ out, long long n_tiles, int S,
                       const NormLut lut_arg) {
    __shared__ float lut[768];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = lut_arg.v[i];
    __syncthreads();
    const int groups_per_row = S / 16;
    const long long total = n_tiles * S * groups_per_row;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    constexpr int U = 4;  // groups per thread and trip: 12 independent 16-byte loads in flight before the first use
    for (long long g0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g0 < total; g0 += U * stride) {
        uint4 raw[U][3];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const long long g = g0 + k * stride;
            if (g < total) {
                const uint4* src = reinterpret_cast<const uint4*>(in + g * 48);   // 16 pixels x 3 bytes, row-major HWC
                raw[k][0] = __ldg(src);
                raw[k][1] = __ldg(src + 1);
                raw[k][2] = __ldg(src + 2);
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const long long g = g0 + k * stride;
            if (g >= total) break;
            const int gx = static_cast<int>(g % groups_per_row);
            const long long ty = g / groups_per_row;
            const int y = static_cast<int>(ty % S);
            const long long t = ty / S;
            const uint8_t* b = reinterpret_cast<const uint8_t*>(raw[k]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                __align__(16) TOut vals[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) vals[i] = cvt_out<TOut>(lut[c * 256 + b[i * 3 + c]]);
                TOut* dst = out + ((t * 3 + c) * S + y) * S + gx * 16;
                const uint4* v4 = reinterpret_cast<const uint4*>(vals);
#pragma unroll
                for (int i = 0; i < static_cast<int>(sizeof(TOut) * 16 / 16); ++i) reinterpret_cast<uint4*>(dst)[i] = v4[i];
            }
        }
    }
}

// Pillow resampling coefficients for one axis (Resample.c precompute_coeffs + normalize_coeffs_8bpc, bicubic a=-0.5)
struct AxisCoeffs {
    int in_size, out_size, ksize;
    std::vector<int> bounds;  // [out][2] = xmin, count
    std::vector<int> kk;      // [out][ksize]
};

double bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

AxisCoeffs make_coeffs(int in_size, int out_size) {
    AxisCoeffs c;
    c.in_size = in_size;
    c.out_size = out_size;
    if (in_size == out_size) {  // Pillow skips the pass: identity tap (2^21 + p*2^22) >> 22 == p
        c.ksize = 1;
        c.bounds.resize(2 * out_size);
        c.kk.resize(out_size);
        for (int i = 0; i < out_size; ++i) {
            c.bounds[2 * i] = i;
            c.bounds[2 * i + 1] = 1;
            c.kk[i] = 1 << 22;
        }
        return c;
    }
    const double scale = static_cast<double>(in_size) / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    c.ksize = static_cast<int>(ceil(support)) * 2 + 1;
    c.bounds.resize(2 * out_size);
    c.kk.assign(static_cast<size_t>(out_size) * c.ksize, 0);
    std::vector<double> w(c.ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            w[x] = bicubic((x + xmin - center + 0.5) * ss);
            ww += w[x];
        }
        for (int x = 0; x < xmax; ++x) {
            double v = w[x];
            if (ww != 0.0) v /= ww;
            c.kk[static_cast<size_t>(xx) * c.ksize + x] =
                v < 0 ? static_cast<int>(-0.5 + v * (1 << 22)) : static_cast<int>(0.5 + v * (1 << 22));
        }
        c.bounds[2 * xx] = xmin;
        c.bounds[2 * xx + 1] = xmax;
    }
    return c;
}

struct DeviceCoeffs {
    int ksize = 0;
    int out_size = 0;
    bool fits_s24 = true;   // every coefficient within (-2^23, 2^23): the fused kernel's three byte planes hold it
    int* bounds = nullptr;
    int* kk = nullptr;
    const std::vector<int>* host_bounds = nullptr;   // [out][2], owned by the cache

    // largest number of input samples any block of `block` consecutive outputs reads (first tap of its first output to
    // the last tap of its last output)
    int max_span(int block) const {
        int best = 0;
        const std::vector<int>& b = *host_bounds;
        for (int o0 = 0; o0 < out_size; o0 += block) {
            const int o1 = std::min(out_size, o0 + block) - 1;
            best = std::max(best, b[2 * o1] + b[2 * o1 + 1] - b[2 * o0]);
        }
        return best;
    }
};

// geometry cache: coefficient tables depend only on (in,out) per axis and are kept per device. Library-owned device memory, never freed.
int get_device_coeffs(int in_size, int out_size, cudaStream_t st, DeviceCoeffs* out) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, DeviceCoeffs> cache;  // (device, in, out): the tables live on one device
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(current_device(), in_size, out_size);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return 0;
    }
    AxisCoeffs c = make_coeffs(in_size, out_size);
    DeviceCoeffs d;
    d.ksize = c.ksize;
    d.out_size = out_size;
    d.host_bounds = new std::vector<int>(c.bounds);   // library-owned like the device tables
    for (int v : c.kk) d.fits_s24 = d.fits_s24 && v > -(1 << 23) && v < (1 << 23);
    MJV_CUDA_OK(cudaMalloc(&d.bounds, c.bounds.size() * sizeof(int)));
    MJV_CUDA_OK(cudaMalloc(&d.kk, c.kk.size() * sizeof(int)));
    MJV_CUDA_OK(cudaMemcpyAsync(d.bounds, c.bounds.data(), c.bounds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    MJV_CUDA_OK(cudaMemcpyAsync(d.kk, c.kk.data(), c.kk.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    MJV_CUDA_OK(cudaStreamSynchronize(st));  // one-time per geometry: the host vectors die at scope exit
    cache[key] = d;
    *out = d;
    return 0;
}

__device__ __forceinline__ uint8_t clip8(int v) { return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// horizontal pass: in [F,H,W,3] -> tmp [F,H,OW,3] (uint8, as Pillow's ImagingResampleHorizontal_8bpc)
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ tmp, int F, int H, int W, int OW,
                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
    const long long total = static_cast<long long>(F) * H * OW;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % OW);
        const long long fy = i / OW;
        const int xmin = bounds[2 * ox], n = bounds[2 * ox + 1];
        const int* k = kk + static_cast<long long>(ox) * ksize;
        const uint8_t* src = in + (fy * W + xmin) * 3;
        int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
        for (int x = 0; x < n; ++x) {
            const int kv = __ldg(k + x);
            s0 += src[x * 3 + 0] * kv;
            s1 += src[x * 3 + 1] * kv;
            s2 += src[x * 3 + 2] * kv;
        }
        uint8_t* dst = tmp + i * 3;
        dst[0] = clip8(s0 >> 22);
        dst[1] = clip8(s1 >> 22);
        dst[2] = clip8(s2 >> 22);
    }
}

// vertical pass fused with the tile crop and normalisation: tmp [F,H,OW,3] -> tiles. Output tile index inside a frame:
// row-major over the (gi x gj) grid (data.py:103-112); this kernel writes `tiles_here` tiles starting at tile_offset
// of each frame's `tiles_per_frame` slots (the thumbnail is a second call with its own geometry).
template <typename TOut>
__global__ void __launch_bounds__(256)
resample_v_tiles_kernel(const uint8_t* __restrict__ tmp, TOut* __restrict__ out, int F, int H, int OW, int OH, int S,
                        int gi, int tiles_per_frame, int tile_offset, const int* __restrict__ bounds,
                        const int* __restrict__ kk, int ksize, const NormLut lut_arg) {
    __shared__ float lut[768];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = lut_arg.v[i];
    __syncthreads();
    const long long total = static_cast<long long>(F) * OH * OW;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int gx = static_cast<int>(i % OW);
        const long long fy = i / OW;
        const int gy = static_cast<int>(fy % OH);
        const long long f = fy / OH;
        const int ymin = bounds[2 * gy], n = bounds[2 * gy + 1];
        const int* k = kk + static_cast<long long>(gy) * ksize;
        const uint8_t* src = tmp + ((f * H + ymin) * OW + gx) * 3;
        int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
        for (int y = 0; y < n; ++y) {
            const int kv = __ldg(k + y);
            const uint8_t* p = src + static_cast<long long>(y) * OW * 3;
            s0 += p[0] * kv;
            s1 += p[1] * kv;
            s2 += p[2] * kv;
        }
        const int tile = (gy / S) * gi + (gx / S);
        const int y = gy % S, x = gx % S;
        TOut* dst = out + ((f * tiles_per_frame + tile_offset + tile) * 3) * S * S + static_cast<long long>(y) * S + x;
        dst[0] = cvt_out<TOut>(lut[clip8(s0 >> 22)]);
        dst[static_cast<long long>(S) * S] = cvt_out<TOut>(lut[256 + clip8(s1 >> 22)]);
        dst[2LL * S * S] = cvt_out<TOut>(lut[512 + clip8(s2 >> 22)]);
    }
}


// Fused resize: one CTA produces a BW x BH block of the resized frame. The horizontal pass runs over just the input rows
// the block's vertical taps read and keeps its uint8 result (Pillow's intermediate image, same rounding) in shared memory;
// the vertical pass, crop, LUT and planar store follow in the same CTA - the [F,H,OW,3] temporary never exists in HBM.
//   stage:   32 input rows per pass, copied with 4-byte cp.async from the aligned word below each segment start;
//   phase 1: lane = row, warp = output column. The taps of one output pixel are 3n contiguous bytes; they are read as
//            words (4 taps = 3 words after a funnel shift) and the bytes picked with PRMT - a byte load per tap
//            and channel made the first version of this kernel shared-memory-issue bound (tools/frames_bench.py: 1.77 ms
//            for 256 x 720p). Coefficients come 4 at a time (one 16-byte broadcast load). Lanes hit 32 different banks
//            (row stride = 4 mod 128 bytes);
//   phase 2: vertical taps on 4 interleaved bytes (one word) per thread;
//   phase 3: LUT + planar stores with lane = output column inside one colour plane (contiguous 2/4-byte elements).
// Work: 3 * (ksize_h * rows_read / BH + ksize_v) integer MACs per output pixel against ~(3*W*H/(OW*OH) + 3*sizeof(TOut)) bytes:
// 1280x720 -> 448x448 is 20 M MACs (+ as many byte extractions) per 4 MB of traffic. The integer pipe, not HBM, bounds
// this kernel for real video sizes (frame path).
struct FusedGeom {
    int F, H, W, OW, OH, S, gi, tiles_per_frame, tile_offset;
    int BW, BH, ksh, ksh4, ksv, ksv4;
    int stage_stride;   // bytes per staged input row (multiple of 4, = 4 mod 128)
    int stage_bytes;    // staging area (also holds the BH x 3*BW output bytes between phases 2 and 3)
    int rmax;           // rows of hbuf
    long long in_bytes;
    const int *bounds_h, *kk_h, *bounds_v, *kk_v;
};
constexpr int kFusedBW = 64;       // output columns per CTA
constexpr int kHbufStride = 196;   // 64 pixels x 3 bytes + 4: 49 words, odd -> conflict-free across rows

__device__ __forceinline__ uint32_t load_word_guarded(const uint8_t* p, const uint8_t* end) {
    if (p + 4 <= end) return __ldg(reinterpret_cast<const uint32_t*>(p));
    uint32_t v = 0;
    for (int b = 0; b < 4; ++b)
        if (p + b < end) v |= static_cast<uint32_t>(p[b]) << (8 * b);
    return v;
}
// Coefficients are 22-bit fixed point (|k| < 2^23): k = k0 + 256 k1 + 65536 k2 with k0, k1 unsigned bytes and k2 a signed
// byte. Four taps of one colour channel are gathered into one register (PRMT) and multiplied plane by plane with dp4a; the
// three partial sums recombine to exactly Pillow's 32-bit accumulator (two's-complement arithmetic modulo 2^32, and the true
// sum fits in 32 bits as it does in Pillow): 15 instructions for 12 multiply-adds instead of 24.
__device__ __forceinline__ int4 split_coeffs(int c0, int c1, int c2, int c3) {
    const uint32_t u0 = static_cast<uint32_t>(c0), u1 = static_cast<uint32_t>(c1), u2 = static_cast<uint32_t>(c2),
                   u3 = static_cast<uint32_t>(c3);
    int4 r;
    r.x = static_cast<int>((u0 & 0xffu) | ((u1 & 0xffu) << 8) | ((u2 & 0xffu) << 16) | ((u3 & 0xffu) << 24));
    r.y = static_cast<int>(((u0 >> 8) & 0xffu) | (((u1 >> 8) & 0xffu) << 8) | (((u2 >> 8) & 0xffu) << 16) | (((u3 >> 8) & 0xffu) << 24));
    r.z = static_cast<int>(((u0 >> 16) & 0xffu) | (((u1 >> 16) & 0xffu) << 8) | (((u2 >> 16) & 0xffu) << 16) | (((u3 >> 16) & 0xffu) << 24));
    r.w = 0;
    return r;
}
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, int b, uint32_t c) { return __dp4a(a, static_cast<uint32_t>(b), c); }
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint8_t finish_taps(uint32_t a0, uint32_t a1, int a2) {
    const uint32_t sum = (1u << 21) + a0 + (a1 << 8) + (static_cast<uint32_t>(a2) << 16);
    return clip8(static_cast<int>(sum) >> 22);
}

template <typename TOut>
__global__ void __launch_bounds__(256, 5)
resample_fused_kernel(const uint8_t* __restrict__ in, TOut* __restrict__ out, const FusedGeom g, const NormLut lut_arg) {
    extern __shared__ __align__(16) uint8_t fsm[];
    float* lut = reinterpret_cast<float*>(fsm);                       // [768]
    int4* kh4 = reinterpret_cast<int4*>(fsm + 3072);                  // [BW][ksh4 / 4]: byte planes of 4 taps, zero padded
    int4* kv4 = kh4 + g.BW * (g.ksh4 >> 2);                           // [BH][ksv4 / 4]
    int* bh = reinterpret_cast<int*>(kv4 + g.BH * (g.ksv4 >> 2));     // [BW][2]
    int* bv = bh + 2 * g.BW;                                          // [BH][2]
    uint8_t* hbuf = reinterpret_cast<uint8_t*>(bv + 2 * g.BH);        // [rmax][kHbufStride]
    uint8_t* stage = hbuf + ((g.rmax * kHbufStride + 15) & ~15);      // [32][stage_stride]

    const int nbx = (g.OW + g.BW - 1) / g.BW, nby = (g.OH + g.BH - 1) / g.BH;
    const int bx = blockIdx.x % nbx;
    const int by = (blockIdx.x / nbx) % nby;
    const long long f = blockIdx.x / (nbx * nby);
    const int gx0 = bx * g.BW, gy0 = by * g.BH;
    const int bw = min(g.BW, g.OW - gx0), bhh = min(g.BH, g.OH - gy0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gh = g.ksh4 >> 2, gv = g.ksv4 >> 2;

    for (int i = tid; i < 768; i += 256) lut[i] = lut_arg.v[i];
    for (int i = tid; i < bw * gh; i += 256) {
        const int ox = i / gh, x4 = (i - ox * gh) * 4;
        const int* src = g.kk_h + static_cast<long long>(gx0 + ox) * g.ksh;
        kh4[i] = split_coeffs(x4 < g.ksh ? src[x4] : 0, x4 + 1 < g.ksh ? src[x4 + 1] : 0, x4 + 2 < g.ksh ? src[x4 + 2] : 0,
                              x4 + 3 < g.ksh ? src[x4 + 3] : 0);
    }
    for (int i = tid; i < bhh * gv; i += 256) {
        const int oy = i / gv, y4 = (i - oy * gv) * 4;
        const int* src = g.kk_v + static_cast<long long>(gy0 + oy) * g.ksv;
        kv4[i] = split_coeffs(y4 < g.ksv ? src[y4] : 0, y4 + 1 < g.ksv ? src[y4 + 1] : 0, y4 + 2 < g.ksv ? src[y4 + 2] : 0,
                              y4 + 3 < g.ksv ? src[y4 + 3] : 0);
    }
    for (int i = tid; i < 2 * bw; i += 256) bh[i] = g.bounds_h[2 * gx0 + i];
    for (int i = tid; i < 2 * bhh; i += 256) bv[i] = g.bounds_v[2 * gy0 + i];
    __syncthreads();
    const int x_lo = bh[0], x_hi = bh[2 * (bw - 1)] + bh[2 * (bw - 1) + 1];
    const int y_lo = bv[0], y_hi = bv[2 * (bhh - 1)] + bv[2 * (bhh - 1) + 1];
    const int seg_words = ((x_hi - x_lo) * 3 + 3) >> 2;
    const int rows = y_hi - y_lo;
    const uint8_t* frame = in + f * g.H * static_cast<long long>(g.W) * 3;
    const uint8_t* in_end = in + g.in_bytes;

    for (int r0 = 0; r0 < rows; r0 += 32) {
        // ---- stage up to 32 input row segments: warp w copies rows w, w+8, ... with 4-byte cp.async from the aligned word at
        // or below the segment start (no registers, no instructions waiting on the loads); the row's 0-3 lead bytes are
        // absorbed by the per-lane funnel shift of the tap loop. Only the last row of the whole input can touch the end of
        // the allocation and takes guarded synchronous loads.
        for (int r = warp; r < 32 && r0 + r < rows; r += 8) {
            const uint8_t* src = frame + (static_cast<long long>(y_lo + r0 + r) * g.W + x_lo) * 3;
            const int lead = static_cast<int>(reinterpret_cast<uintptr_t>(src) & 3);
            const uint8_t* src_al = src - lead;
            uint32_t* dst4 = reinterpret_cast<uint32_t*>(stage + r * g.stage_stride);
            if (src_al + 4 * (seg_words + 1) <= in_end) {
                const uint32_t dst_s = ptx::smem_u32(dst4);
                for (int i = lane; i <= seg_words; i += 32)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_s + 4 * i), "l"(src_al + 4 * i) : "memory");
            } else {
                for (int i = lane; i <= seg_words; i += 32) dst4[i] = load_word_guarded(src_al + 4 * i, in_end);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        // ---- horizontal taps: lane = staged row, warp = output column (w, w+8, ...)
        if (r0 + lane < rows) {
            const uint32_t* row = reinterpret_cast<const uint32_t*>(stage + lane * g.stage_stride);
            uint8_t* hrow = hbuf + (r0 + lane) * kHbufStride;
            const int lead = static_cast<int>(
                reinterpret_cast<uintptr_t>(frame + (static_cast<long long>(y_lo + r0 + lane) * g.W + x_lo) * 3) & 3);
            for (int ox = warp; ox < bw; ox += 8) {
                const int boff = (bh[2 * ox] - x_lo) * 3 + lead, n = bh[2 * ox + 1];   // n is warp-uniform
                const int sh = (boff & 3) * 8;
                const uint32_t* wp = row + (boff >> 2);
                const int4* kp = kh4 + ox * gh;
                uint32_t r0a = 0, r1a = 0, g0a = 0, g1a = 0, b0a = 0, b1a = 0;
                int r2a = 0, g2a = 0, b2a = 0;
                for (int x4 = 0; x4 < n; x4 += 4) {
                    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
                    wp += 3;
                    // 12 stream bytes = 4 taps x RGB
                    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh),
                                   v2 = __funnelshift_r(w2, w3, sh);
                    const int4 k = *kp++;   // taps beyond n are zero
                    const uint32_t r4 = __byte_perm(__byte_perm(v0, v1, 0x0630), v2, 0x5210);   // bytes 0, 3, 6, 9
                    const uint32_t g4 = __byte_perm(__byte_perm(v0, v1, 0x0741), v2, 0x6210);   // bytes 1, 4, 7, 10
                    const uint32_t b4 = __byte_perm(__byte_perm(v0, v1, 0x0052), v2, 0x7410);   // bytes 2, 5, 8, 11
                    r0a = dp4a_uu(r4, k.x, r0a); r1a = dp4a_uu(r4, k.y, r1a); r2a = dp4a_us(r4, k.z, r2a);
                    g0a = dp4a_uu(g4, k.x, g0a); g1a = dp4a_uu(g4, k.y, g1a); g2a = dp4a_us(g4, k.z, g2a);
                    b0a = dp4a_uu(b4, k.x, b0a); b1a = dp4a_uu(b4, k.y, b1a); b2a = dp4a_us(b4, k.z, b2a);
                }
                hrow[ox * 3 + 0] = finish_taps(r0a, r1a, r2a);
                hrow[ox * 3 + 1] = finish_taps(g0a, g1a, g2a);
                hrow[ox * 3 + 2] = finish_taps(b0a, b1a, b2a);
            }
        }
        __syncthreads();
    }
    // ---- vertical taps: one word = 4 interleaved bytes of a horizontal-pass row per thread, four rows per dp4a (the 4 x 4
    // byte transpose costs 8 PRMTs); results parked as bytes
    uint32_t* obuf = reinterpret_cast<uint32_t*>(stage);   // [BH][48] words
    constexpr int kRowWords = 48;                          // 64 pixels x 3 bytes
    constexpr int kHW = kHbufStride / 4;
    for (int item = tid; item < bhh * kRowWords; item += 256) {
        const int oy = item / kRowWords, wi = item - oy * kRowWords;
        const int ymin = bv[2 * oy] - y_lo, n = bv[2 * oy + 1];
        const int4* kq = kv4 + oy * gv;
        const uint32_t* p = reinterpret_cast<const uint32_t*>(hbuf + ymin * kHbufStride) + wi;
        uint32_t a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
        int a2[4] = {0, 0, 0, 0};
        for (int y4 = 0; y4 < n; y4 += 4) {
            const uint32_t w0 = p[0], w1 = p[kHW], w2 = p[2 * kHW], w3 = p[3 * kHW];   // rows beyond n: zero coefficients
            p += 4 * kHW;
            const uint32_t t01 = __byte_perm(w0, w1, 0x5140), t23 = __byte_perm(w2, w3, 0x5140);
            const uint32_t u01 = __byte_perm(w0, w1, 0x7362), u23 = __byte_perm(w2, w3, 0x7362);
            const uint32_t c[4] = {__byte_perm(t01, t23, 0x5410), __byte_perm(t01, t23, 0x7632),
                                   __byte_perm(u01, u23, 0x5410), __byte_perm(u01, u23, 0x7632)};
            const int4 k = *kq++;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                a0[b] = dp4a_uu(c[b], k.x, a0[b]);
                a1[b] = dp4a_uu(c[b], k.y, a1[b]);
                a2[b] = dp4a_us(c[b], k.z, a2[b]);
            }
        }
        obuf[item] = static_cast<uint32_t>(finish_taps(a0[0], a1[0], a2[0])) |
                     (static_cast<uint32_t>(finish_taps(a0[1], a1[1], a2[1])) << 8) |
                     (static_cast<uint32_t>(finish_taps(a0[2], a1[2], a2[2])) << 16) |
                     (static_cast<uint32_t>(finish_taps(a0[3], a1[3], a2[3])) << 24);
    }
    __syncthreads();
    // ---- crop + LUT + planar store: one warp per (output row, colour plane), lane = output column - the row pointer is
    // computed once per warp and the 2 x 32 elements of a plane row are contiguous in HBM
    const uint8_t* ob = reinterpret_cast<const uint8_t*>(obuf);
    // S is a multiple of both block sizes in every geometry the reference produces (448): the block lies in one tile
    const bool one_tile = (g.S % kFusedBW == 0) && (g.S % g.BH == 0);
    const long long tile_base = f * g.tiles_per_frame + g.tile_offset;
    const int tile0 = (gy0 / g.S) * g.gi + (gx0 / g.S), ty0 = gy0 % g.S, tx0 = gx0 % g.S;
    for (int rp = warp; rp < bhh * 3; rp += 8) {
        const int oy = rp / 3, c = rp - oy * 3;
        const uint8_t* orow = ob + oy * (kRowWords * 4) + c;
        const float* lc = lut + c * 256;
        if (one_tile) {
            TOut* dst = out + (((tile_base + tile0) * 3 + c) * g.S + (ty0 + oy)) * g.S + tx0;
#pragma unroll
            for (int h2 = 0; h2 < kFusedBW / 32; ++h2) {
                const int xo = h2 * 32 + lane;
                if (xo < bw) dst[xo] = cvt_out<TOut>(lc[orow[xo * 3]]);
            }
        } else {
            for (int xo = lane; xo < bw; xo += 32) {
                const int gx = gx0 + xo, gy = gy0 + oy;
                const int tile = (gy / g.S) * g.gi + (gx / g.S);
                out[(((tile_base + tile) * 3 + c) * g.S + gy % g.S) * g.S + gx % g.S] = cvt_out<TOut>(lc[orow[xo * 3]]);
            }
        }
    }
}

// Block geometry of the fused kernel for one resize; false when no block height fits in shared memory (the caller then
// falls back to the two passes).
bool plan_fused(int F, int H, int W, int OW, int OH, int S, int gi, int tpf, int tile_offset, const DeviceCoeffs& ch,
                const DeviceCoeffs& cv, FusedGeom* out_g, size_t* out_smem, long long* out_ctas) {
    constexpr size_t kMaxSmem = 200 * 1024;
    if (!ch.fits_s24 || !cv.fits_s24) return false;
    for (int BH = 32; BH >= 8; BH >>= 1) {
        FusedGeom g;
        g.F = F; g.H = H; g.W = W; g.OW = OW; g.OH = OH; g.S = S; g.gi = gi; g.tiles_per_frame = tpf; g.tile_offset = tile_offset;
        g.BW = kFusedBW; g.BH = BH; g.ksh = ch.ksize; g.ksh4 = (ch.ksize + 3) & ~3; g.ksv = cv.ksize; g.ksv4 = (cv.ksize + 3) & ~3;
        g.bounds_h = ch.bounds; g.kk_h = ch.kk; g.bounds_v = cv.bounds; g.kk_v = cv.kk;
        g.in_bytes = (long long)F * H * W * 3;
        const int seg = ch.max_span(g.BW) * 3 + 24;                // + rounding to whole words and the 4-word read-ahead
        g.stage_stride = ((seg + 127) / 128) * 128 + 4;
        g.stage_bytes = std::max(32 * g.stage_stride, BH * 3 * g.BW);
        g.rmax = cv.max_span(BH) + 3;                              // the vertical pass reads rows four at a time
        const size_t smem = 3072 + 4 * (size_t)(g.BW * g.ksh4 + BH * g.ksv4 + 2 * g.BW + 2 * BH) +
                            (((size_t)g.rmax * kHbufStride + 15) & ~(size_t)15) + (size_t)g.stage_bytes;
        const long long ctas = (long long)F * ((OW + g.BW - 1) / g.BW) * ((OH + BH - 1) / BH);
        if (smem > kMaxSmem || ctas <= 0 || ctas >= (1LL << 31)) continue;
        *out_g = g;
        *out_smem = smem;
        *out_ctas = ctas;
        return true;
    }
    return false;
}

template <typename TOut>
int run_fused(const uint8_t* frames, TOut* out, const FusedGeom& g, size_t smem, long long ctas, const NormLut& lut,
              cudaStream_t st) {
    static PerDeviceOnce attr;
    if (attr.first())
        MJV_CUDA_OK(cudaFuncSetAttribute(resample_fused_kernel<TOut>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    resample_fused_kernel<TOut><<<(unsigned)ctas, 256, smem, st>>>(frames, out, g, lut);
    MJV_CUDA_OK(cudaGetLastError());
    return 0;
}

// grid choice of dynamic_preprocess (data.py:66-99), float64 like the reference
void choose_grid(int width, int height, int min_num, int max_num, int image_size, int* gi, int* gj) {
    const double aspect = static_cast<double>(width) / height;
    std::vector<std::pair<int, int>> cands;
    for (int n = min_num; n <= max_num; ++n)
        for (int i = 1; i <= n; ++i)
            for (int j = 1; j <= n; ++j)
                if (i * j <= max_num && i * j >= min_num) {
                    bool seen = false;
                    for (auto& c : cands) seen |= (c.first == i && c.second == j);
                    if (!seen) cands.push_back({i, j});
                }
    // python: sorted(set, key=area) — a stable sort of the set's iteration order. Ties in area only matter through
    // the ratio comparison below, where equal ratio_diff candidates of equal area behave identically.
    std::stable_sort(cands.begin(), cands.end(),
                     [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first * a.second < b.first * b.second; });
    double best_diff = INFINITY;
    int bi = 1, bj = 1;
    const double area = static_cast<double>(width) * height;
    for (auto& c : cands) {
        const double diff = fabs(aspect - static_cast<double>(c.first) / c.second);
        if (diff < best_diff) {
            best_diff = diff;
            bi = c.first;
            bj = c.second;
        } else if (diff == best_diff) {
            if (area > 0.5 * image_size * image_size * c.first * c.second) {
                bi = c.first;
                bj = c.second;
            }
        }
    }
    *gi = bi;
    *gj = bj;
}

template <typename TOut>
int preprocess_t(const uint8_t* frames, int F, int H, int W, int image_size, int max_num, int use_thumbnail, TOut* out,
                 uint8_t* scratch, size_t scratch_bytes, int* tiles_per_frame_out, cudaStream_t st) {
    static const NormLut lut = make_lut();
    int gi, gj;
    choose_grid(W, H, 1, max_num, image_size, &gi, &gj);
    const int blocks = gi * gj;
    const int tpf = blocks + ((use_thumbnail && blocks != 1) ? 1 : 0);
    if (tiles_per_frame_out) *tiles_per_frame_out = tpf;
    if (out == nullptr) return 0;  // query only
    const int S = image_size;
    const int grid = sm_count() * 8;
    if (blocks == 1 && W == S && H == S && S % 16 == 0 && (reinterpret_cast<uintptr_t>(frames) & 15) == 0) {
        {
            LaunchScope scope(TAG_FRAMES, st, (double)F * S * S * 3.0 * (1.0 + sizeof(TOut)));
            normalize_tiles_kernel<TOut><<<grid, 256, 0, st>>>(frames, out, F, S, lut);
        }
        MJV_CUDA_OK(cudaGetLastError());
        return 0;
    }
    const int OW = S * gi, OH = S * gj;
    // MJV_RESIZE_TWO_PASS=1 forces the older two-kernel path (tools / A-B measurements; also the fallback for geometries whose
    // block does not fit in shared memory - it needs the caller's scratch for the [F,H,OW,3] intermediate)
    const char* two = getenv("MJV_RESIZE_TWO_PASS");
    const bool allow_fused = !(two != nullptr && two[0] == '1');
    const auto resize = [&](int ow, int oh, int grid_i, int tile_offset) -> int {
        DeviceCoeffs ch, cv;
        MJV_PROPAGATE(get_device_coeffs(W, ow, st, &ch));
        MJV_PROPAGATE(get_device_coeffs(H, oh, st, &cv));
        FusedGeom fg;
        size_t fsmem = 0;
        long long fctas = 0;
        if (allow_fused && plan_fused(F, H, W, ow, oh, S, grid_i, tpf, tile_offset, ch, cv, &fg, &fsmem, &fctas)) {
            // algorithmic bytes: the frame read once, the tiles written once
            LaunchScope scope(TAG_FRAMES, st, (double)F * (3.0 * H * W + 3.0 * ow * oh * sizeof(TOut)), 1);
            return run_fused<TOut>(frames, out, fg, fsmem, fctas, lut, st);
        }
        const size_t need = static_cast<size_t>(F) * H * ow * 3;
        MJV_CHECK(scratch != nullptr && scratch_bytes >= need, MJV_ERR_NOMEM,
                  "preprocess: scratch of %zu bytes needed, %zu given", need, scratch_bytes);
        {
            LaunchScope scope(TAG_FRAMES, st, (double)F * (3.0 * H * W + 6.0 * H * ow + 3.0 * ow * oh * sizeof(TOut)), 2);
            resample_h_kernel<<<grid, 256, 0, st>>>(frames, scratch, F, H, W, ow, ch.bounds, ch.kk, ch.ksize);
            resample_v_tiles_kernel<TOut><<<grid, 256, 0, st>>>(scratch, out, F, H, ow, oh, S, grid_i, tpf, tile_offset, cv.bounds,
                                                                cv.kk, cv.ksize, lut);
        }
        MJV_CUDA_OK(cudaGetLastError());
        return 0;
    };
    MJV_PROPAGATE(resize(OW, OH, gi, 0));
    if (tpf > blocks) MJV_PROPAGATE(resize(S, S, 1, blocks));  // thumbnail: whole frame -> S x S (data.py:114-116)
    return 0;
}

}  // namespace

int preprocess_frames(const void* frames, int F, int H, int W, int image_size, int max_num, int use_thumbnail,
                      void* out, int out_dtype, void* scratch, size_t scratch_bytes, int* tiles_per_frame,
                      cudaStream_t st) {
    MJV_CHECK(F >= 0 && H > 0 && W > 0 && image_size > 0 && max_num >= 1, MJV_ERR_ARG, "preprocess: bad geometry");
    if (out_dtype == MJV_DTYPE_BF16)
        return preprocess_t<__nv_bfloat16>((const uint8_t*)frames, F, H, W, image_size, max_num, use_thumbnail,
                                           (__nv_bfloat16*)out, (uint8_t*)scratch, scratch_bytes, tiles_per_frame, st);
    if (out_dtype == MJV_DTYPE_F32)
        return preprocess_t<float>((const uint8_t*)frames, F, H, W, image_size, max_num, use_thumbnail, (float*)out,
                                   (uint8_t*)scratch, scratch_bytes, tiles_per_frame, st);
    return set_error(MJV_ERR_ARG, "preprocess: unsupported output dtype %d", out_dtype);
}

}  // namespace mjv

extern "C" int mjv_preprocess_frames(const void* frames_u8, int32_t n_frames, int32_t height, int32_t width,
                                     int32_t image_size, int32_t max_num, int32_t use_thumbnail, void* out,
                                     int32_t out_dtype, void* scratch, int64_t scratch_bytes,
                                     int32_t* tiles_per_frame, void* stream) {
    return mjv::preprocess_frames(frames_u8, n_frames, height, width, image_size, max_num, use_thumbnail, out,
                                  out_dtype, scratch, (size_t)scratch_bytes, tiles_per_frame,
                                  static_cast<cudaStream_t>(stream));
}
// Pure host helpers (no GPU): tile-grid choice and the Pillow coefficient table, exposed for CPU-side parity tests.
extern "C" int mjv_choose_grid(int32_t width, int32_t height, int32_t min_num, int32_t max_num, int32_t image_size,
                               int32_t* gi, int32_t* gj) {
    if (!gi || !gj || width <= 0 || height <= 0) return mjv::set_error(MJV_ERR_ARG, "mjv_choose_grid: bad args");
    mjv::choose_grid(width, height, min_num, max_num, image_size, gi, gj);
    return 0;
}
extern "C" int mjv_bicubic_coeffs(int32_t in_size, int32_t out_size, int32_t* ksize, int32_t* bounds, int32_t* kk,
                                  int64_t kk_capacity) {
    if (in_size <= 0 || out_size <= 0 || !ksize) return mjv::set_error(MJV_ERR_ARG, "mjv_bicubic_coeffs: bad args");
    mjv::AxisCoeffs c = mjv::make_coeffs(in_size, out_size);
    *ksize = c.ksize;
    if (bounds == nullptr || kk == nullptr) return 0;
    if ((int64_t)c.kk.size() > kk_capacity) return mjv::set_error(MJV_ERR_NOMEM, "mjv_bicubic_coeffs: kk too small");
    memcpy(bounds, c.bounds.data(), c.bounds.size() * sizeof(int));
    memcpy(kk, c.kk.data(), c.kk.size() * sizeof(int));
    return 0;
}
