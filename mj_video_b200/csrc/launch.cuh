// Launch accounting + optional per-launch CUDA-event timing.
//   * every kernel this library launches goes through a LaunchScope, so `mjv_launch_count()` is the library's own
//     count (bench.py reports it as `gpu_launches`);
//   * when profiling is enabled for a tag (mjv_profile_enable), the scope brackets the launch with two events on the
//     launching stream; mjv_profile_read() returns (tag, milliseconds, algorithmic work) per launch. This is how
//     bench.py measures the dominant kernel's average duration live inside the timed region.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace mjv {

enum KernelTag : int {
    TAG_GEMM_STORE = 0, TAG_GEMM_GELU, TAG_GEMM_RESID, TAG_GEMM_SWIGLU, TAG_GEMM_QKV_LLM, TAG_GEMM_QKV_VIT,
    TAG_GEMM_PATCH, TAG_GEMM_SCATTER,
    TAG_ATTN_BIDIR = 8, TAG_ATTN_CAUSAL, TAG_LAYERNORM, TAG_PIXSHUF_LN, TAG_RMSNORM, TAG_EMBED, TAG_IM2COL, TAG_CLS_ROWS,
    TAG_FRAMES, TAG_HEAD_LINEAR, TAG_HEAD_FINALIZE, TAG_COUNT
};

std::atomic<long long>& launch_counter();
bool profile_wants(int tag);
void profile_begin(int tag, cudaStream_t st, double work, int* slot);
void profile_end(int slot, cudaStream_t st);

// work = algorithmic FLOPs (tensor-bound kernels) or bytes (HBM-bound kernels) of this launch, 0 if unknown here
struct LaunchScope {
    int slot = -1;
    cudaStream_t st;
    LaunchScope(int tag, cudaStream_t stream, double work = 0.0, int n_launches = 1) : st(stream) {
        launch_counter().fetch_add(n_launches, std::memory_order_relaxed);
        if (profile_wants(tag)) profile_begin(tag, stream, work, &slot);
    }
    ~LaunchScope() {
        if (slot >= 0) profile_end(slot, st);
    }
};

}  // namespace mjv
