// Shared host/device helpers for the COMET B200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>

#include "../../include/comet_b200.h"

namespace comet {

// ---- error plumbing ---------------------------------------------------
char* last_error_buf();  // thread-local, defined in cabi.cu
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
#define COMET_REQUIRE(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return ::comet::fail(COMET_ERR_INVALID, __VA_ARGS__); \
  } while (0)
#define COMET_CUDA(call)                                                                        \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return ::comet::fail(COMET_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));    \
  } while (0)
void count_launch();     // defined in cabi.cu: every kernel launch of the library goes through launch_status()
inline int launch_status(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(COMET_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  return COMET_OK;
}

// ---- TMA descriptors (host): cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda) ----------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<TensorMapEncodeFn>(ptr);
    else
      cudaGetLastError();
  }
  return fn;
}
// SM count of the CURRENT device if it is compute capability 10.x, else 0 (no device: 0).  Cached per device ordinal:
// a process may drive several GPUs.
inline int device_sm_count_if_sm100() {
  static int cached[64];
  static bool known[64] = {false};
  int dev = 0, major = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 0; }
  if (!known[dev]) {
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = (major == 10 && sms > 0) ? sms : 0;
    known[dev] = true;
  }
  return cached[dev];
}
// Library options (comet_set_option, cabi.cu): explicit A/B switches instead of environment variables on the launch path.
int option(int which);

// ---- pyramid geometry (host) -------------------------------------------
struct Levels {
  int L;
  int H[COMET_MAX_LEVELS];
  int W[COMET_MAX_LEVELS];
  long long off[COMET_MAX_LEVELS];  // element offset of level l inside `pyr` (off[0] unused)
};
inline Levels make_levels(int BS, int C, int H, int W, int L) {
  Levels lv{};
  lv.L = L;
  long long o = 0;
  for (int l = 0; l < L && l < COMET_MAX_LEVELS; ++l) {
    lv.H[l] = H;
    lv.W[l] = W;
    lv.off[l] = (l == 0) ? 0 : o;
    if (l >= 1) o += (long long)BS * C * H * W;
    H /= 2;
    W /= 2;
  }
  return lv;
}

// ---- device helpers ----------------------------------------------------
__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
// Round the four floats of a vector to bf16 precision (round-to-nearest-even, like round_bf16): two packed conversions
// + four bit operations instead of four scalar conversions + four shifts.
__device__ __forceinline__ void round_bf16x4(float4& g) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(g.x, g.y), b = __floats2bfloat162_rn(g.z, g.w);
  const uint32_t ua = *reinterpret_cast<const uint32_t*>(&a), ub = *reinterpret_cast<const uint32_t*>(&b);
  g.x = __uint_as_float(ua << 16);
  g.y = __uint_as_float(ua & 0xffff0000u);
  g.z = __uint_as_float(ub << 16);
  g.w = __uint_as_float(ub & 0xffff0000u);
}

// Per-axis description of a (2r+1)-wide window of bilinear samples whose centre is `p`
// (grid_sample semantics, align_corners=True, restated from ATen GridSampler):
//   grid position g in [0, 2r+2): integer tap  i0 + g  (zeros: masked when outside; border: clamped)
//   window index  i in [0, 2r+1): value = w0(i) * V[g=i] + w1(i) * V[g=i+1]
struct AxisWindow {
  int i0;     // floor(p) - r
  float p;    // centre coordinate
  float f;    // p - floor(p)
  int size;   // map extent along this axis
  int r;
  bool border;
  __device__ __forceinline__ void init(float centre, int size_, int r_, bool border_) {
    size = size_;
    r = r_;
    border = border_;
    // keep float->int conversion defined for wild coordinates; such windows are entirely off the map
    centre = fminf(fmaxf(centre, -1.0e6f), 1.0e6f);
    if (size == 1) centre = 0.f;  // bilinear_sampler scales by 2/max(size-1,1) and grid_sample by (size-1)/2 = 0
    p = centre;
    float fl = floorf(centre);
    f = centre - fl;
    i0 = (int)fl - r;
  }
  // integer tap of grid position g; returns false when the tap contributes zero
  __device__ __forceinline__ bool tap(int g, int& pos) const {
    pos = i0 + g;
    if (size == 1) { pos = 0; return true; }
    if (border) {
      pos = min(max(pos, 0), size - 1);
      return true;
    }
    return pos >= 0 && pos < size;
  }
  __device__ __forceinline__ void weights(int i, float& w0, float& w1) const {
    w0 = 1.f - f;
    w1 = f;
    if (size == 1) { w0 = 1.f; w1 = 0.f; return; }
    if (border) {
      float c = p + (float)(i - r);
      if (c < 0.f || c > (float)(size - 1)) { w0 = 1.f; w1 = 0.f; }
    }
  }
};

}  // namespace comet
