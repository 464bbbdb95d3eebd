// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM (sm_100a). Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../mj_video_b200/csrc/ptx.cuh"
using namespace mjv;

__device__ __forceinline__ void ld_x64(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr, uint32_t* v) {  // 16 lanes x 256 bit x 8 = 4 KB, 32 regs
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// MODE 5: ld x32, wait, ld x32, wait (8 KB); MODE 6: one ld .x64 instruction + wait (8 KB); MODE 7: 4 x (ld x16) + wait (8 KB)
// MODE 8: 16x256b.x8 twice (both lane halves) + wait (8 KB); MODE 9: ld x32 + 64 dependent FMAs + wait (compute overlap)
// MODE 0: ld x32 x2 + wait per iteration (8 KB per warp-iteration); MODE 1: ld x32 only (4 KB); MODE 2: st x32 + wait
// MODE 3: 4 ld x32 back-to-back then one wait (16 KB); MODE 4: ld x32x2 + st x32 (the softmax pattern)
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(int iters, long long* out_clk, unsigned* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { ptx::tmem_alloc(&slot, 256); ptx::tmem_relinquish(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t v[64]; uint32_t acc = 0;
    for (int i = 0; i < 64; ++i) v[i] = threadIdx.x + i;
    ptx::tmem_st_x32(base, v); ptx::tmem_st_x32(base + 32, v + 32); ptx::tmem_st_x32(base + 64, v); ptx::tmem_st_x32(base + 96, v);
    ptx::tmem_wait_st();
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { ptx::tmem_ld_x32(base, v); ptx::tmem_ld_x32(base + 32, v + 32); ptx::tmem_wait_ld(); acc += v[0] ^ v[63]; }
        if (MODE == 1) { ptx::tmem_ld_x32(base, v); ptx::tmem_wait_ld(); acc += v[0] ^ v[31]; }
        if (MODE == 2) { v[0] = acc + it; ptx::tmem_st_x32(base, v); ptx::tmem_wait_st(); }
        if (MODE == 3) { ptx::tmem_ld_x32(base, v); ptx::tmem_ld_x32(base + 32, v + 32); ptx::tmem_wait_ld(); acc += v[0] ^ v[63];
                         ptx::tmem_ld_x32(base + 64, v); ptx::tmem_ld_x32(base + 96, v + 32); ptx::tmem_wait_ld(); acc += v[0] ^ v[63]; }
        if (MODE == 5) { ptx::tmem_ld_x32(base, v); ptx::tmem_wait_ld(); acc += v[0] ^ v[31]; ptx::tmem_ld_x32(base + 32, v + 32); ptx::tmem_wait_ld(); acc += v[32] ^ v[63]; }
        if (MODE == 6) { ld_x64(base, v); ptx::tmem_wait_ld(); acc += v[0] ^ v[63]; }
        if (MODE == 7) { ptx::tmem_ld_x16(base, v); ptx::tmem_ld_x16(base + 16, v + 16); ptx::tmem_ld_x16(base + 32, v + 32); ptx::tmem_ld_x16(base + 48, v + 48); ptx::tmem_wait_ld(); acc += v[0] ^ v[63]; }
        if (MODE == 8) { ld_16x256b_x8(base, v); ld_16x256b_x8(base + (16u << 16), v + 32); ptx::tmem_wait_ld(); acc += v[0] ^ v[63]; }
        if (MODE == 4) { ptx::tmem_ld_x32(base, v); ptx::tmem_ld_x32(base + 32, v + 32); ptx::tmem_wait_ld(); acc += v[0] ^ v[63];
                         v[1] = acc; ptx::tmem_st_x32(base, v); ptx::tmem_wait_st(); }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out_clk[blockIdx.x] = t1 - t0;
    if (acc == 0xdeadbeef) sink[0] = acc;
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(slot, 256); }
}

template <int MODE>
void run(const char* name, int threads, int ctas_per_sm, double bytes_per_warp_iter) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4000; const int grid = sms * ctas_per_sm;
    long long* d; unsigned* s; cudaMalloc(&d, grid * 8); cudaMalloc(&s, 4);
    k<MODE><<<grid, threads>>>(iters, d, s); cudaDeviceSynchronize();
    k<MODE><<<grid, threads>>>(iters, d, s);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    const double per_sm = bytes_per_warp_iter * (threads / 32) * ctas_per_sm * iters / avg;
    printf("%-28s thr %3d cta/SM %d : %7.1f clk/iter  %7.1f B/clk/SM  (%s)\n", name, threads, ctas_per_sm, avg / iters, per_sm, cudaGetErrorString(e));
    cudaFree(d); cudaFree(s);
}

int main() {
    for (int c = 1; c <= 2; ++c)
        for (int thr = 32; thr <= 256; thr *= 2) {
            if (thr == 64) continue;
            run<0>("ld x64 + wait", thr, c, 8192);
            run<1>("ld x32 + wait", thr, c, 4096);
            run<3>("2x(ld x64 + wait)", thr, c, 16384);
            run<2>("st x32 + wait", thr, c, 4096);
            run<4>("ld x64, st x32", thr, c, 8192 + 4096);
            run<5>("ld x32,wait,ld x32,wait", thr, c, 8192);
            run<6>("ld .x64 single instr", thr, c, 8192);
            run<7>("4 x ld x16 + wait", thr, c, 8192);
            run<8>("2 x 16x256b.x8 + wait", thr, c, 8192);
        }
    return 0;
}
