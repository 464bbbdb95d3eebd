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
    int* bounds = nullptr;
    int* kk = nullptr;
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
    const size_t need = static_cast<size_t>(F) * H * (OW > S ? OW : S) * 3;
    MJV_CHECK(scratch != nullptr && scratch_bytes >= need, MJV_ERR_NOMEM,
              "preprocess: scratch of %zu bytes needed, %zu given", need, scratch_bytes);
    DeviceCoeffs ch, cv;
    MJV_PROPAGATE(get_device_coeffs(W, OW, st, &ch));
    MJV_PROPAGATE(get_device_coeffs(H, OH, st, &cv));
    {
        LaunchScope scope(TAG_FRAMES, st, (double)F * (3.0 * H * W + 6.0 * H * OW + 3.0 * OW * OH * sizeof(TOut)), 2);
        resample_h_kernel<<<grid, 256, 0, st>>>(frames, scratch, F, H, W, OW, ch.bounds, ch.kk, ch.ksize);
        resample_v_tiles_kernel<TOut><<<grid, 256, 0, st>>>(scratch, out, F, H, OW, OH, S, gi, tpf, 0, cv.bounds, cv.kk,
                                                            cv.ksize, lut);
    }
    MJV_CUDA_OK(cudaGetLastError());
    if (tpf > blocks) {  // thumbnail: whole frame -> S x S (data.py:114-116)
        DeviceCoeffs th, tv;
        MJV_PROPAGATE(get_device_coeffs(W, S, st, &th));
        MJV_PROPAGATE(get_device_coeffs(H, S, st, &tv));
        {
            LaunchScope scope(TAG_FRAMES, st, (double)F * (3.0 * H * W + 6.0 * H * S + 3.0 * S * S * sizeof(TOut)), 2);
            resample_h_kernel<<<grid, 256, 0, st>>>(frames, scratch, F, H, W, S, th.bounds, th.kk, th.ksize);
            resample_v_tiles_kernel<TOut><<<grid, 256, 0, st>>>(scratch, out, F, H, S, S, S, 1, tpf, blocks, tv.bounds,
                                                                tv.kk, tv.ksize, lut);
        }
        MJV_CUDA_OK(cudaGetLastError());
    }
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
