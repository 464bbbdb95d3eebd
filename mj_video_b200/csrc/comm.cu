// The one collective of the path: all-gather of the per-row score table over NCCL (NVLink 5 /
// NVSwitch). The reference never issues a collective (its NCCL group of size 1 only exists so that
// torch.distributed.get_rank() works: modeling_internvl_chat.py:172, eval_genai_mjvideo.py:24); this is the data-parallel
// addition named by BASELINE.json's north_star. Payload is <= 136 B per row, i.e. latency-bound: a plain ncclAllGather on
// the compute stream is the right tool — there is no compute to overlap it with and nothing a custom NVLink kernel could win.
//
// libnccl is resolved with dlopen at first use (the copy PyTorch already loaded is reused), so the library loads and exports
// its symbols on machines without NCCL; the entry points then fail loudly with MJV_ERR_STATE.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace mjv {
namespace {

struct NcclUniqueId {
    char internal[128];
};
typedef void* NcclComm;
typedef int (*PFN_GetUniqueId)(NcclUniqueId*);
typedef int (*PFN_CommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*PFN_CommDestroy)(NcclComm);
typedef int (*PFN_AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef const char* (*PFN_GetErrorString)(int);
constexpr int kNcclFloat32 = 7;  // ncclFloat32 in ncclDataType_t (stable across NCCL 2.x)

struct Nccl {
    void* lib = nullptr;
    PFN_GetUniqueId get_unique_id = nullptr;
    PFN_CommInitRank comm_init_rank = nullptr;
    PFN_CommDestroy comm_destroy = nullptr;
    PFN_AllGather all_gather = nullptr;
    PFN_GetErrorString error_string = nullptr;
    bool ok = false;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("MJV_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* name : names) {
            if (name == nullptr || n.lib != nullptr) continue;
            n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        }
        if (n.lib == nullptr) return;
        n.get_unique_id = reinterpret_cast<PFN_GetUniqueId>(dlsym(n.lib, "ncclGetUniqueId"));
        n.comm_init_rank = reinterpret_cast<PFN_CommInitRank>(dlsym(n.lib, "ncclCommInitRank"));
        n.comm_destroy = reinterpret_cast<PFN_CommDestroy>(dlsym(n.lib, "ncclCommDestroy"));
        n.all_gather = reinterpret_cast<PFN_AllGather>(dlsym(n.lib, "ncclAllGather"));
        n.error_string = reinterpret_cast<PFN_GetErrorString>(dlsym(n.lib, "ncclGetErrorString"));
        n.ok = n.get_unique_id && n.comm_init_rank && n.comm_destroy && n.all_gather;
    });
    return n;
}

int nccl_ready() {
    if (!nccl().ok)
        return set_error(MJV_ERR_STATE, "NCCL not available (dlopen libnccl.so.2 failed; set MJV_NCCL_LIB): %s",
                         dlerror() ? dlerror() : "symbols missing");
    return 0;
}
int nccl_fail(const char* what, int rc) {
    return set_error(MJV_ERR_CUDA, "%s failed: %s", what, nccl().error_string ? nccl().error_string(rc) : "nccl error");
}

}  // namespace
}  // namespace mjv

using namespace mjv;

extern "C" int mjv_comm_unique_id(void* id_128_bytes) {
    MJV_CHECK(id_128_bytes != nullptr, MJV_ERR_ARG, "mjv_comm_unique_id: null buffer");
    MJV_PROPAGATE(nccl_ready());
    NcclUniqueId id;
    const int rc = nccl().get_unique_id(&id);
    if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id_128_bytes, &id, sizeof(id));
    return 0;
}

extern "C" int mjv_comm_create(const void* id_128_bytes, int32_t rank, int32_t world, void** comm_out) {
    MJV_CHECK(id_128_bytes && comm_out && world >= 1 && rank >= 0 && rank < world, MJV_ERR_ARG, "mjv_comm_create: bad arguments");
    MJV_PROPAGATE(nccl_ready());
    NcclUniqueId id;
    memcpy(&id, id_128_bytes, sizeof(id));
    NcclComm c = nullptr;
    const int rc = nccl().comm_init_rank(&c, world, id, rank);  // binds to the calling thread's current CUDA device
    if (rc != 0) return nccl_fail("ncclCommInitRank", rc);
    *comm_out = c;
    return 0;
}

extern "C" int mjv_comm_destroy(void* comm) {
    if (comm == nullptr) return 0;
    MJV_PROPAGATE(nccl_ready());
    const int rc = nccl().comm_destroy(comm);
    return rc == 0 ? 0 : nccl_fail("ncclCommDestroy", rc);
}

extern "C" int mjv_allgather_scores(void* comm, const float* local_rows, int64_t rows_per_rank, int32_t width, float* out,
                                    void* stream) {
    MJV_CHECK(comm && local_rows && out && rows_per_rank > 0 && width > 0, MJV_ERR_ARG, "mjv_allgather_scores: bad arguments");
    MJV_PROPAGATE(nccl_ready());
    const int rc = nccl().all_gather(local_rows, out, static_cast<size_t>(rows_per_rank) * width, kNcclFloat32, comm,
                                     static_cast<cudaStream_t>(stream));
    return rc == 0 ? 0 : nccl_fail("ncclAllGather", rc);
}
