// Launch accounting: every kernel this library launches goes through count_launch(), so bench.py can report
// `gpu_launches` from the library's own counter rather than an estimate.
#pragma once
#include <atomic>

namespace mjv {
std::atomic<long long>& launch_counter();
inline void count_launch(int n = 1) { launch_counter().fetch_add(n, std::memory_order_relaxed); }
}  // namespace mjv
