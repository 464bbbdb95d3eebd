// Host-side helpers shared by the C-ABI translation units: error reporting, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mjv_b200.h"

namespace mjv {

// ---------------------------------------------------------------- errors (thread-local message, negative codes)
inline char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
inline int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
#define MJV_CUDA_OK(expr)                                                                                  \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            return ::mjv::set_error(MJV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                    __FILE__, __LINE__);                                                   \
    } while (0)
#define MJV_CHECK(cond, code, ...)                               \
    do {                                                         \
        if (!(cond)) return ::mjv::set_error(code, __VA_ARGS__); \
    } while (0)
#define MJV_PROPAGATE(expr)     \
    do {                        \
        int _r = (expr);        \
        if (_r != 0) return _r; \
    } while (0)

// ---------------------------------------------------------------- TMA descriptors
// libcuda is resolved at run time through the runtime API so the library loads (and exports its symbols) on a
// machine without a driver; any compute entry point then fails loudly.
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_encode_fn() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }
    return fn;
}

// bf16 tensor of up to 3 dims, innermost contiguous. dims/strides innermost-first; strides in ELEMENTS for dims 1..;
// box innermost-first. 128-byte swizzle (box[0] must be 64 bf16). Out-of-bounds elements are zero-filled.
inline int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_elems, const uint32_t* box) {
    PFN_tmapEncodeTiled fn = get_encode_fn();
    MJV_CHECK(fn != nullptr, MJV_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) gstr[i - 1] = strides_elems[i] * 2;
    }
    MJV_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, MJV_ERR_ARG, "TMA base %p not 16-byte aligned", base);
    for (int i = 0; i + 1 < rank; ++i)
        MJV_CHECK((gstr[i] & 15) == 0, MJV_ERR_ARG, "TMA stride %llu not a multiple of 16 bytes",
                  (unsigned long long)gstr[i]);
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                    es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MJV_CHECK(r == CUDA_SUCCESS, MJV_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

inline int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows) {
    uint64_t dims[2] = {cols, rows};
    uint64_t str[2] = {1, ld};
    uint32_t box[2] = {64, box_rows};
    return make_tmap_bf16(out, base, 2, dims, str, box);
}

// Process-wide caches are kept PER DEVICE: Engine(device=...) and load_video(device=...) accept any device, and function
// attributes, SM counts and library-owned device tables all belong to one device.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int sm_count() {
    static int n[kMaxDevices] = {};
    const int dev = current_device();
    if (n[dev] == 0) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    return n[dev];
}
// `static PerDeviceOnce once; if (once.first()) cudaFuncSetAttribute(...)`: true once per device
struct PerDeviceOnce {
    bool done[kMaxDevices] = {};
    bool first() {
        const int dev = current_device();
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};

}  // namespace mjv
