// C-ABI core: version, errors, launch counter, per-launch event profiler.
#include <mutex>
#include <vector>

#include "common.cuh"
#include "launch.cuh"

namespace mjv {

std::atomic<long long>& launch_counter() {
    static std::atomic<long long> c{0};
    return c;
}

namespace {
struct Record {
    int tag;
    double work;
    cudaEvent_t e0, e1;
};
struct Profiler {
    std::mutex mu;
    unsigned mask = 0;
    std::vector<Record> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};
Profiler& prof() {
    static Profiler p;
    return p;
}
const char* kTagNames[TAG_COUNT] = {
    "gemm_store", "gemm_gelu", "gemm_resid", "gemm_swiglu", "gemm_qkv_llm", "gemm_qkv_vit", "gemm_patch",
    "gemm_scatter", "attn_bidir", "attn_causal", "layernorm", "pixshuf_layernorm", "rmsnorm", "embed", "im2col",
    "cls_rows", "frames", "head_linear", "head_finalize"};
}  // namespace

bool profile_wants(int tag) { return (prof().mask >> tag) & 1u; }

void profile_begin(int tag, cudaStream_t st, double work, int* slot) {
    Profiler& p = prof();
    std::lock_guard<std::mutex> lock(p.mu);
    Record r;
    r.tag = tag;
    r.work = work;
    r.e0 = p.get();
    r.e1 = p.get();
    cudaEventRecord(r.e0, st);
    p.recs.push_back(r);
    *slot = static_cast<int>(p.recs.size()) - 1;
}

void profile_end(int slot, cudaStream_t st) {
    Profiler& p = prof();
    std::lock_guard<std::mutex> lock(p.mu);
    if (slot >= 0 && slot < static_cast<int>(p.recs.size())) cudaEventRecord(p.recs[slot].e1, st);
}

}  // namespace mjv

extern "C" int mjv_abi_version(void) { return MJV_ABI_VERSION; }
extern "C" int mjv_abi_struct_size(int which) {
    switch (which) {
        case 0: return static_cast<int>(sizeof(mjv_gemm_desc));
        case 1: return static_cast<int>(sizeof(mjv_attn_desc));
        case 2: return static_cast<int>(sizeof(mjv_config));
        case 3: return static_cast<int>(sizeof(mjv_batch));
        case 4: return static_cast<int>(sizeof(mjv_outputs));
        default: return -1;
    }
}
extern "C" const char* mjv_last_error(void) { return mjv::last_error_buf(); }
extern "C" long long mjv_launch_count(void) { return mjv::launch_counter().load(); }
extern "C" void mjv_reset_launch_count(void) { mjv::launch_counter().store(0); }

extern "C" void mjv_profile_enable(uint32_t tag_mask) { mjv::prof().mask = tag_mask; }

// Waits for the recorded events (the caller has normally synchronised already), fills up to `capacity` entries and
// clears the record list. Returns the number of records that existed.
extern "C" int mjv_profile_read(int32_t* tags, float* ms, double* work, int32_t capacity) {
    mjv::Profiler& p = mjv::prof();
    std::lock_guard<std::mutex> lock(p.mu);
    const int n = static_cast<int>(p.recs.size());
    for (int i = 0; i < n; ++i) {
        mjv::Record& r = p.recs[i];
        float t = 0.f;
        cudaEventSynchronize(r.e1);
        cudaEventElapsedTime(&t, r.e0, r.e1);
        if (i < capacity) {
            if (tags) tags[i] = r.tag;
            if (ms) ms[i] = t;
            if (work) work[i] = r.work;
        }
        p.pool.push_back(r.e0);
        p.pool.push_back(r.e1);
    }
    p.recs.clear();
    return n;
}

extern "C" const char* mjv_kernel_tag_name(int32_t tag) {
    return (tag >= 0 && tag < mjv::TAG_COUNT) ? mjv::kTagNames[tag] : "?";
}
extern "C" int mjv_kernel_tag_count(void) { return mjv::TAG_COUNT; }
