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
static std::atomic<int> g_options[COMET_OPT_COUNT] = {{1}, {1}, {0}, {0}, {1}, {3}, {1}, {1}, {1}, {3}};   // specialised paths on; the two measured-slower experiments off
int option(int which) { return (which >= 0 && which < COMET_OPT_COUNT) ? g_options[which].load(std::memory_order_relaxed) : 0; }
}  // namespace comet

extern "C" int comet_set_option(int option, int value) {
  COMET_REQUIRE(option >= 0 && option < COMET_OPT_COUNT, "unknown option %d", option);
  // boolean switches, except COMET_OPT_GEMM_TMA_STORE / COMET_OPT_GEMM_PAIR whose values are bit masks
  comet::g_options[option].store((option == COMET_OPT_GEMM_TMA_STORE || option == COMET_OPT_ATTN_MMA) ? (value & 3) : option == COMET_OPT_GEMM_PAIR ? (value & 7) : (value ? 1 : 0), std::memory_order_relaxed);
  return COMET_OK;
}
extern "C" int comet_get_option(int option) { return comet::option(option); }

extern "C" int comet_version(void) { return 200; /* 0.2.0, round 2 */ }
extern "C" const char* comet_last_error(void) { return comet::last_error_buf(); }

extern "C" long long comet_launch_count(void) { return comet::g_launches.load(std::memory_order_relaxed); }

#ifndef COMET_HAVE_TC
extern "C" int comet_has_tensor_path(void) { return 0; }
#endif
