// Library-level entry points of the C ABI (include/comet_b200.h).
#include "comet_common.cuh"

#include <atomic>

namespace comet {
char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace comet

extern "C" int comet_version(void) { return 100; /* 0.1.0, round 1 */ }
extern "C" const char* comet_last_error(void) { return comet::last_error_buf(); }

extern "C" long long comet_launch_count(void) { return comet::g_launches.load(std::memory_order_relaxed); }

#ifndef COMET_HAVE_TC
extern "C" int comet_has_tensor_path(void) { return 0; }
#endif
