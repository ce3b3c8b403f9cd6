// Feature pyramid: level l = 2x2 / stride-2 average pooling of level l-1 (floor sizes, no padding).
// Replaces the F.avg_pool2d chain of CorrBlock.__init__ / EfficientCorrBlock.__init__
// (comet/models/track_modules/blocks.py:368-374, :438-444).  Run once per tracker call.
#include "comet_common.cuh"

namespace comet {

// One thread per output element; adjacent threads read adjacent input pairs (coalesced), planes = BS*C.
__global__ void __launch_bounds__(256) avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                        long long planes, int Hi, int Wi, int Ho, int Wo) {
  const long long total = planes * Ho * Wo;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % Wo);
    const long long t = idx / Wo;
    const int yo = (int)(t % Ho);
    const long long pl = t / Ho;
    const float* src = in + (pl * Hi + 2 * yo) * Wi + 2 * xo;
    const float a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + Wi), d = __ldg(src + Wi + 1);
    out[idx] = ((a + b) + (c + d)) * 0.25f;
  }
}

}  // namespace comet

using namespace comet;

extern "C" long long comet_pyramid_offset(int BS, int C, int H, int W, int level) {
  if (level < 1 || level >= COMET_MAX_LEVELS) return -1;
  return make_levels(BS, C, H, W, level + 1).off[level];
}

extern "C" long long comet_pyramid_elems(int BS, int C, int H, int W, int L) {
  if (L < 1 || L > COMET_MAX_LEVELS) return -1;
  Levels lv = make_levels(BS, C, H, W, L);
  if (L == 1) return 0;
  return lv.off[L - 1] + (long long)BS * C * lv.H[L - 1] * lv.W[L - 1];
}

extern "C" int comet_pyramid_f32(const float* fmaps, float* pyr, int BS, int C, int H, int W, int L,
                                 comet_stream_t stream) {
  COMET_REQUIRE(BS >= 0 && C >= 1 && H >= 1 && W >= 1, "bad shape");
  COMET_REQUIRE(L >= 1 && L <= COMET_MAX_LEVELS, "num_levels must be in [1, %d] (got %d)", COMET_MAX_LEVELS, L);
  COMET_REQUIRE((H >> (L - 1)) >= 1 && (W >> (L - 1)) >= 1, "map %dx%d too small for %d levels", H, W, L);
  if (L == 1 || BS == 0) return COMET_OK;
  COMET_REQUIRE(fmaps && pyr, "null pointer");
  Levels lv = make_levels(BS, C, H, W, L);
  const long long planes = (long long)BS * C;
  for (int l = 1; l < L; ++l) {
    const float* src = (l == 1) ? fmaps : pyr + lv.off[l - 1];
    const long long total = planes * lv.H[l] * lv.W[l];
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 64) blocks = 148LL * 64;  // grid-stride beyond 64 CTAs per SM
    avgpool2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, pyr + lv.off[l], planes, lv.H[l - 1],
                                                                        lv.W[l - 1], lv.H[l], lv.W[l]);
    int rc = launch_status("avgpool2_kernel");
    if (rc != COMET_OK) return rc;
  }
  return COMET_OK;
}
