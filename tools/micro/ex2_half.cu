// Microbenchmark: throughput of packed half-precision exp2 (ex2.approx.ftz.f16x2 / .bf16x2) vs fp32 MUFU.EX2 on sm_100a,
// and of the surrounding conversions.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void k(int iters, long long* out_clk, unsigned* sink, unsigned seed) {
    uint32_t x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = seed * (threadIdx.x + i + 1);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x[i]));
            if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
            if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
            if (MODE == 3) { float a = __uint_as_float(x[i]), b = __uint_as_float(x[(i + 1) & 31]); uint32_t r;
                             asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); x[i] ^= r; }
            if (MODE == 4) { __half2 h = *reinterpret_cast<__half2*>(&x[i]); float2 f = __half22float2(h); x[i] = __float_as_uint(f.x + f.y); }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc ^= x[i];
    if (threadIdx.x == 0) out_clk[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int MODE>
void run(const char* name, int threads) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 2000;
    long long* d; unsigned* s; cudaMalloc(&d, sms * 8); cudaMalloc(&s, 4);
    k<MODE><<<sms, threads>>>(iters, d, s, 3u); cudaDeviceSynchronize();
    k<MODE><<<sms, threads>>>(iters, d, s, 3u);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, d, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += h[i]; avg /= sms;
    const int wps = threads / 128;
    printf("%-40s warps/SMSP %d : %6.2f clk per warp-instr(-group) per SMSP (%s)\n", name, wps, avg / iters / 32 / wps, cudaGetErrorString(e));
    cudaFree(d); cudaFree(s);
}

int main() {
    for (int thr = 128; thr <= 256; thr *= 2) {
        run<0>("ex2.approx.ftz.f32", thr);
        run<1>("ex2.approx.f16x2 (2 results)", thr);
        run<2>("ex2.approx.ftz.bf16x2 (2 results)", thr);
        run<3>("cvt.rn.f16x2.f32 (+xor)", thr);
        run<4>("half2 -> float2 (+fadd)", thr);
    }
    return 0;
}
