// sin/cos encodings of the track tokens and the time axis (comet/models/utils.py:37-101, :724-832).
//   get_2d_embedding                      -> embed2d_kernel        (float32, accurate sinf/cosf: arguments reach ~1e4)
//   get_1d_sincos_pos_embed(_from_grid)   -> sincos1d_kernel       (float64 inside, like the reference's host code)
//   get_2d_sincos_pos_embed               -> sincos2d_kernel
//   sample_features4d(pos_embed, coords0) -> sampled_pos_emb_kernel (table never stored: evaluated at the 4 taps)
#include "comet_common.cuh"

namespace comet {

// omega_k = 1 / 10000^(k / half), float64 (utils.py:50-52)
__device__ __forceinline__ double omega_of(int k, int half) { return 1.0 / pow(10000.0, (double)k / (double)half); }

__global__ void __launch_bounds__(256) embed2d_kernel(const float* __restrict__ xy, float* __restrict__ out,
                                                       long long M, int C, int cat) {
  const int Dout = 2 * C + (cat ? 2 : 0);
  const long long total = M * Dout;
  const float step = 1000.0f / (float)C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int d = (int)(idx % Dout);
    const long long m = idx / Dout;
    float v;
    if (cat && d < 2) {
      v = __ldg(xy + m * 2 + d);
    } else {
      if (cat) d -= 2;
      const int axis = d / C, w = d - axis * C;
      const float arg = __fmul_rn(__ldg(xy + m * 2 + axis), (float)(w & ~1) * step);
      v = (w & 1) ? cosf(arg) : sinf(arg);
    }
    out[idx] = v;
  }
}

__global__ void __launch_bounds__(256) sincos1d_kernel(const float* __restrict__ pos, float* __restrict__ out,
                                                        long long M, int D) {
  const int half = D / 2;
  const long long total = M * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % D);
    const long long m = idx / D;
    const int k = d < half ? d : d - half;
    const double a = (double)__ldg(pos + m) * omega_of(k, half);
    out[idx] = (float)(d < half ? sin(a) : cos(a));
  }
}

// out (D,H,W): channel blocks [sin_x | cos_x | sin_y | cos_y], each D/4 wide (utils.py:740-745, :796-803)
__global__ void __launch_bounds__(256) sincos2d_kernel(float* __restrict__ out, int D, int H, int W) {
  const int quarter = D / 4;
  const long long total = (long long)D * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const long long t = idx / W;
    const int y = (int)(t % H);
    const int d = (int)(t / H);
    const int part = d / quarter, k = d - part * quarter;
    const double a = (double)(part < 2 ? x : y) * omega_of(k, quarter);
    out[idx] = (float)((part & 1) ? cos(a) : sin(a));
  }
}

__device__ __forceinline__ float table_entry(int part, int k, int quarter, int x, int y) {
  const double a = (double)(part < 2 ? x : y) * omega_of(k, quarter);
  return (float)((part & 1) ? cos(a) : sin(a));
}

__global__ void __launch_bounds__(256) sampled_pos_emb_kernel(const float* __restrict__ coords0, long long c_sb,
                                                               long long c_sn, float* __restrict__ out, int B, int N,
                                                               int D, int H, int W) {
  const int quarter = D / 4;
  const long long total = (long long)B * N * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % D);
    const long long t = idx / D;
    const int n = (int)(t % N);
    const int b = (int)(t / N);
    const float* cp = coords0 + b * c_sb + n * c_sn;
    AxisWindow ax, ay;  // r = 0: a single bilinear sample, border padding (sample_features4d)
    ax.init(__ldg(cp), W, 0, true);
    ay.init(__ldg(cp + 1), H, 0, true);
    int x0, x1, y0, y1;
    ax.tap(0, x0); ax.tap(1, x1); ay.tap(0, y0); ay.tap(1, y1);
    float wx0, wx1, wy0, wy1;
    ax.weights(0, wx0, wx1);
    ay.weights(0, wy0, wy1);
    const int part = d / quarter, k = d - part * quarter;
    float v = table_entry(part, k, quarter, x0, y0) * (wx0 * wy0);
    v += table_entry(part, k, quarter, x1, y0) * (wx1 * wy0);
    v += table_entry(part, k, quarter, x0, y1) * (wx0 * wy1);
    v += table_entry(part, k, quarter, x1, y1) * (wx1 * wy1);
    out[idx] = v;
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

extern "C" int comet_embed2d_f32(const float* xy, float* out, long long M, int C, int cat_coords,
                                 comet_stream_t stream) {
  COMET_REQUIRE(M >= 0 && C >= 2 && C % 2 == 0, "get_2d_embedding needs an even C >= 2 (got %d)", C);
  if (M == 0) return COMET_OK;
  COMET_REQUIRE(xy && out, "null pointer");
  embed2d_kernel<<<grid_for(M * (2 * C + 2)), 256, 0, (cudaStream_t)stream>>>(xy, out, M, C, cat_coords != 0);
  return launch_status("embed2d_kernel");
}

extern "C" int comet_sincos1d_from_grid_f32(const float* pos, float* out, long long M, int D, comet_stream_t stream) {
  COMET_REQUIRE(M >= 0 && D >= 2 && D % 2 == 0, "embed_dim must be even (got %d)", D);
  if (M == 0) return COMET_OK;
  COMET_REQUIRE(pos && out, "null pointer");
  sincos1d_kernel<<<grid_for(M * D), 256, 0, (cudaStream_t)stream>>>(pos, out, M, D);
  return launch_status("sincos1d_kernel");
}

extern "C" int comet_sincos2d_f32(float* out, int D, int H, int W, comet_stream_t stream) {
  COMET_REQUIRE(D >= 4 && D % 4 == 0, "embed_dim must be a multiple of 4 (got %d)", D);
  COMET_REQUIRE(H >= 1 && W >= 1, "bad grid size");
  COMET_REQUIRE(out, "null pointer");
  sincos2d_kernel<<<grid_for((long long)D * H * W), 256, 0, (cudaStream_t)stream>>>(out, D, H, W);
  return launch_status("sincos2d_kernel");
}

extern "C" int comet_sampled_pos_emb_f32(const float* coords0, long long c_sb, long long c_sn, float* out, int B,
                                         int N, int D, int H, int W, comet_stream_t stream) {
  COMET_REQUIRE(B >= 0 && N >= 0 && D >= 4 && D % 4 == 0, "embed_dim must be a multiple of 4 (got %d)", D);
  COMET_REQUIRE(H >= 1 && W >= 1, "bad grid size");
  if ((long long)B * N == 0) return COMET_OK;
  COMET_REQUIRE(coords0 && out, "null pointer");
  sampled_pos_emb_kernel<<<grid_for((long long)B * N * D), 256, 0, (cudaStream_t)stream>>>(coords0, c_sb, c_sn, out,
                                                                                            B, N, D, H, W);
  return launch_status("sampled_pos_emb_kernel");
}
