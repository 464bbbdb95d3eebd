// C-ABI core: version, errors, launch counter.
#include "common.cuh"
#include "launch.cuh"

namespace mjv {
std::atomic<long long>& launch_counter() {
    static std::atomic<long long> c{0};
    return c;
}
}  // namespace mjv

extern "C" int mjv_abi_version(void) { return MJV_ABI_VERSION; }
extern "C" const char* mjv_last_error(void) { return mjv::last_error_buf(); }
extern "C" long long mjv_launch_count(void) { return mjv::launch_counter().load(); }
extern "C" void mjv_reset_launch_count(void) { mjv::launch_counter().store(0); }
