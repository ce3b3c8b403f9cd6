// Library-level entry points of the C ABI (include/comet_b200.h).
#include "comet_common.cuh"

namespace comet {
char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
}  // namespace comet

extern "C" int comet_version(void) { return 100; /* 0.1.0, round 1 */ }
extern "C" const char* comet_last_error(void) { return comet::last_error_buf(); }

#ifndef COMET_HAVE_TC
extern "C" int comet_has_tensor_path(void) { return 0; }
#endif
