// bilinear_sampler / sample_features4d (comet/models/utils.py:874-974) as direct gather kernels.
// Coordinates arrive in pixel units; the reference scales them to [-1,1] and grid_sample scales them
// back -- that float32 round trip is reproduced so that results agree to rounding.
#include "comet_common.cuh"

namespace comet {

__device__ __forceinline__ float ref_pixel(float c, int size, bool align) {
  // utils.py:925-935 then ATen grid_sampler_unnormalize
  if (align) {
    const float scale = (float)(2.0 / (double)max(size - 1, 1));
    const float g = c * scale - 1.f;
    return ((g + 1.f) / 2.f) * (float)(size - 1);
  }
  const float scale = (float)(2.0 / (double)size);
  const float g = c * scale - 1.f;
  return ((g + 1.f) * (float)size - 1.f) / 2.f;
}

struct Tap {
  int i0, i1;
  float w0, w1;
  bool ok0, ok1;
  __device__ __forceinline__ void init(float p, int size, bool border) {
    p = fminf(fmaxf(p, -1.0e6f), 1.0e6f);
    if (border) p = fminf((float)(size - 1), fmaxf(p, 0.f));
    const float fl = floorf(p);
    i0 = (int)fl;
    i1 = i0 + 1;
    w1 = p - fl;
    w0 = 1.f - w1;
    ok0 = i0 >= 0 && i0 < size;
    ok1 = i1 >= 0 && i1 < size;
  }
};

__global__ void __launch_bounds__(256) sampler4d_kernel(const float* __restrict__ in, const float* __restrict__ coords,
                                                         float* __restrict__ out, int B, int C, int H, int W,
                                                         long long HoWo, bool align, bool border) {
  const long long total = (long long)B * C * HoWo;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long o = idx % HoWo;
    const long long t = idx / HoWo;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const float* cp = coords + ((long long)b * HoWo + o) * 2;
    Tap tx, ty;
    tx.init(ref_pixel(__ldg(cp), W, align), W, border);
    ty.init(ref_pixel(__ldg(cp + 1), H, align), H, border);
    const float* img = in + ((long long)b * C + c) * H * W;
    float v = 0.f;
    if (ty.ok0 && tx.ok0) v += __ldg(img + (long long)ty.i0 * W + tx.i0) * (tx.w0 * ty.w0);
    if (ty.ok0 && tx.ok1) v += __ldg(img + (long long)ty.i0 * W + tx.i1) * (tx.w1 * ty.w0);
    if (ty.ok1 && tx.ok0) v += __ldg(img + (long long)ty.i1 * W + tx.i0) * (tx.w0 * ty.w1);
    if (ty.ok1 && tx.ok1) v += __ldg(img + (long long)ty.i1 * W + tx.i1) * (tx.w1 * ty.w1);
    out[idx] = v;
  }
}

__global__ void __launch_bounds__(256) sampler5d_kernel(const float* __restrict__ in, const float* __restrict__ coords,
                                                         float* __restrict__ out, int B, int C, int T, int H, int W,
                                                         long long DHW, bool align, bool border) {
  const long long total = (long long)B * C * DHW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long o = idx % DHW;
    const long long t = idx / DHW;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    const float* cp = coords + ((long long)b * DHW + o) * 3;  // (t, x, y)  utils.py:921-923
    Tap tt, tx, ty;
    tt.init(ref_pixel(__ldg(cp), T, align), T, border);
    tx.init(ref_pixel(__ldg(cp + 1), W, align), W, border);
    ty.init(ref_pixel(__ldg(cp + 2), H, align), H, border);
    const float* vol = in + ((long long)b * C + c) * T * H * W;
    float v = 0.f;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      const bool okt = kt ? tt.ok1 : tt.ok0;
      const int it = kt ? tt.i1 : tt.i0;
      const float wt = kt ? tt.w1 : tt.w0;
#pragma unroll
      for (int ky = 0; ky < 2; ++ky) {
        const bool oky = ky ? ty.ok1 : ty.ok0;
        const int iy = ky ? ty.i1 : ty.i0;
        const float wy = ky ? ty.w1 : ty.w0;
#pragma unroll
        for (int kx = 0; kx < 2; ++kx) {
          const bool okx = kx ? tx.ok1 : tx.ok0;
          const int ix = kx ? tx.i1 : tx.i0;
          const float wx = kx ? tx.w1 : tx.w0;
          if (okt && oky && okx) v += __ldg(vol + ((long long)it * H + iy) * W + ix) * (wx * wy * wt);
        }
      }
    }
    out[idx] = v;
  }
}

// (B,C,H,W) @ (B,R,2) -> (B,R,C); c is the fastest output index so stores coalesce.
__global__ void __launch_bounds__(256) features4d_kernel(const float* __restrict__ in, long long in_sb,
                                                          const float* __restrict__ coords, long long c_sb,
                                                          long long c_sr, float* __restrict__ out, int B, int C, int H,
                                                          int W, int R) {
  const long long total = (long long)B * R * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const long long t = idx / C;
    const int r = (int)(t % R);
    const int b = (int)(t / R);
    const float* cp = coords + b * c_sb + r * c_sr;
    Tap tx, ty;
    tx.init(ref_pixel(__ldg(cp), W, true), W, true);
    ty.init(ref_pixel(__ldg(cp + 1), H, true), H, true);
    const float* img = in + b * in_sb + (long long)c * H * W;
    float v = 0.f;
    if (ty.ok0 && tx.ok0) v += __ldg(img + (long long)ty.i0 * W + tx.i0) * (tx.w0 * ty.w0);
    if (ty.ok0 && tx.ok1) v += __ldg(img + (long long)ty.i0 * W + tx.i1) * (tx.w1 * ty.w0);
    if (ty.ok1 && tx.ok0) v += __ldg(img + (long long)ty.i1 * W + tx.i0) * (tx.w0 * ty.w1);
    if (ty.ok1 && tx.ok1) v += __ldg(img + (long long)ty.i1 * W + tx.i1) * (tx.w1 * ty.w1);
    out[idx] = v;
  }
}

// Channel-last input (B,H,W,C): one warp per (b, r) point, lanes stride over the channels, so each of the four taps is a
// contiguous C-float line (coalesced).  Serves sample_features4d on a channels-last fmaps view and the sampled position
// embedding from the cached channel-last sin/cos table (in_sb = 0: one table shared by the batch).
__global__ void __launch_bounds__(256) features4d_cl_kernel(const float* __restrict__ in, long long in_sb,
                                                             const float* __restrict__ coords, long long c_sb,
                                                             long long c_sr, float* __restrict__ out, int B, int C, int H,
                                                             int W, int R) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long pt = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pt < (long long)B * R; pt += nw) {
    const int r = (int)(pt % R);
    const int b = (int)(pt / R);
    const float* cp = coords + b * c_sb + r * c_sr;
    Tap tx, ty;
    tx.init(ref_pixel(__ldg(cp), W, true), W, true);
    ty.init(ref_pixel(__ldg(cp + 1), H, true), H, true);
    const float* img = in + b * in_sb;
    const float* p00 = img + ((long long)ty.i0 * W + tx.i0) * C;
    const float* p01 = img + ((long long)ty.i0 * W + tx.i1) * C;
    const float* p10 = img + ((long long)ty.i1 * W + tx.i0) * C;
    const float* p11 = img + ((long long)ty.i1 * W + tx.i1) * C;
    const bool k00 = ty.ok0 && tx.ok0, k01 = ty.ok0 && tx.ok1, k10 = ty.ok1 && tx.ok0, k11 = ty.ok1 && tx.ok1;
    const float w00 = tx.w0 * ty.w0, w01 = tx.w1 * ty.w0, w10 = tx.w0 * ty.w1, w11 = tx.w1 * ty.w1;
    float* o = out + pt * C;
    for (int c = lane; c < C; c += 32) {
      float v = 0.f;   // ATen accumulation order: nw, ne, sw, se
      if (k00) v += __ldg(p00 + c) * w00;
      if (k01) v += __ldg(p01 + c) * w01;
      if (k10) v += __ldg(p10 + c) * w10;
      if (k11) v += __ldg(p11 + c) * w11;
      o[c] = v;
    }
  }
}

static inline unsigned grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace comet

using namespace comet;

extern "C" int comet_bilinear_sampler4d_f32(const float* input, const float* coords, float* out, int B, int C, int H,
                                            int W, int Ho, int Wo, int align_corners, int pad_mode,
                                            comet_stream_t stream) {
  COMET_REQUIRE(B >= 0 && C >= 0 && Ho >= 0 && Wo >= 0 && H >= 1 && W >= 1, "bad shape");
  COMET_REQUIRE(pad_mode == COMET_PAD_ZEROS || pad_mode == COMET_PAD_BORDER, "bad pad_mode %d", pad_mode);
  const long long total = (long long)B * C * Ho * Wo;
  if (total == 0) return COMET_OK;
  COMET_REQUIRE(input && coords && out, "null pointer");
  sampler4d_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(input, coords, out, B, C, H, W,
                                                                      (long long)Ho * Wo, align_corners != 0,
                                                                      pad_mode == COMET_PAD_BORDER);
  return launch_status("sampler4d_kernel");
}

extern "C" int comet_bilinear_sampler5d_f32(const float* input, const float* coords, float* out, int B, int C, int T,
                                            int H, int W, int Do, int Ho, int Wo, int align_corners, int pad_mode,
                                            comet_stream_t stream) {
  COMET_REQUIRE(B >= 0 && C >= 0 && Do >= 0 && Ho >= 0 && Wo >= 0 && T >= 1 && H >= 1 && W >= 1, "bad shape");
  COMET_REQUIRE(pad_mode == COMET_PAD_ZEROS || pad_mode == COMET_PAD_BORDER, "bad pad_mode %d", pad_mode);
  const long long total = (long long)B * C * Do * Ho * Wo;
  if (total == 0) return COMET_OK;
  COMET_REQUIRE(input && coords && out, "null pointer");
  sampler5d_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(input, coords, out, B, C, T, H, W,
                                                                      (long long)Do * Ho * Wo, align_corners != 0,
                                                                      pad_mode == COMET_PAD_BORDER);
  return launch_status("sampler5d_kernel");
}

extern "C" int comet_sample_features4d_f32(const float* input, long long in_sb, const float* coords, long long c_sb,
                                           long long c_sr, float* out, int B, int C, int H, int W, int R,
                                           comet_stream_t stream) {
  COMET_REQUIRE(B >= 0 && C >= 0 && R >= 0 && H >= 1 && W >= 1, "bad shape");
  const long long total = (long long)B * R * C;
  if (total == 0) return COMET_OK;
  COMET_REQUIRE(input && coords && out, "null pointer");
  features4d_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(input, in_sb, coords, c_sb, c_sr, out, B, C, H,
                                                                       W, R);
  return launch_status("features4d_kernel");
}

extern "C" int comet_sample_features4d_cl_f32(const float* input, long long in_sb, const float* coords, long long c_sb,
                                              long long c_sr, float* out, int B, int C, int H, int W, int R,
                                              comet_stream_t stream) {
  COMET_REQUIRE(B >= 0 && C >= 0 && R >= 0 && H >= 1 && W >= 1, "bad shape");
  const long long pts = (long long)B * R;
  if (pts == 0 || C == 0) return COMET_OK;
  COMET_REQUIRE(input && coords && out, "null pointer");
  long long blocks = (pts + 7) / 8;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  features4d_cl_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(input, in_sb, coords, c_sb, c_sr, out, B, C, H, W,
                                                                         R);
  return launch_status("features4d_cl_kernel");
}
