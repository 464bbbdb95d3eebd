// Micro-benchmark (tools only): issue rate of MUFU.EX2 on f32 against the packed bf16x2 form, one and two warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(uint32_t* out, long long* clk, int iters) {
    uint32_t a[8];
    for (int i = 0; i < 8; ++i) a[i] = 0x3c003c00u + threadIdx.x + i;   // small bf16 pairs / small floats
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(*(float*)&a[i]) : "f"(*(float*)&a[i]));
            else asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(a[i]) : "r"(a[i]));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
    uint32_t* out; long long* clk;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&clk, 1024);
    for (int warps : {1, 4, 8, 16}) {
        for (int mode = 0; mode < 2; ++mode) {
            const int iters = 2000;
            if (mode == 0) k<0><<<1, warps * 32>>>(out, clk, iters); else k<1><<<1, warps * 32>>>(out, clk, iters);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
            printf("warps/CTA %2d (%.1f per scheduler) %s: %.2f cycles per warp instruction per scheduler\n", warps, warps / 4.0,
                   mode ? "bf16x2" : "f32   ", (double)c / (iters * 8.0) / (warps > 4 ? warps / 4.0 : 1.0));
        }
    }
    return 0;
}
