// Materialised correlation volume for one pyramid level (CorrBlock.corr, blocks.py:409-429):
//   vol[bs, n, hw] = (sum_c T[bs, n, c] * F[bs, c, hw]) / sqrt(C)
// Only used to serve the public attribute CorrBlock.corrs_pyramid; the tracker path uses the fused kernels
// and never writes the volume.  Plain shared-memory tiled SIMT GEMM (float32, 64x64 tile, 4x4 per thread).
#include "comet_common.cuh"

namespace comet {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256) corr_volume_kernel(const float* __restrict__ T, long long t_sbs, long long t_sn,
                                                           const float* __restrict__ F, float* __restrict__ V, int N,
                                                           int C, int HW, float sqrt_c, int bf16) {
  __shared__ float As[TK][TM + 1];  // As[k][m]
  __shared__ float Bs[TK][TN];      // Bs[k][n]
  const int bs = blockIdx.z;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const float* Tb = T + bs * t_sbs;
  const float* Fb = F + (long long)bs * C * HW;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < C; k0 += TK) {
    for (int i = threadIdx.x; i < TM * TK; i += 256) {
      const int m = i / TK, k = i - m * TK;
      float v = (m0 + m < N && k0 + k < C) ? __ldg(Tb + (long long)(m0 + m) * t_sn + k0 + k) : 0.f;
      As[k][m] = bf16 ? round_bf16(v) : v;
    }
    for (int i = threadIdx.x; i < TK * TN; i += 256) {
      const int k = i / TN, n = i - k * TN;
      float v = (n0 + n < HW && k0 + k < C) ? __ldg(Fb + (long long)(k0 + k) * HW + n0 + n) : 0.f;
      Bs[k][n] = bf16 ? round_bf16(v) : v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= HW) continue;
      float v = acc[i][j];
      if (bf16) v = round_bf16(v);
      v = __fdiv_rn(v, sqrt_c);
      if (bf16) v = round_bf16(v);
      V[((long long)bs * N + m) * HW + n] = v;
    }
  }
}

}  // namespace comet

using namespace comet;

extern "C" int comet_corr_volume_f32(const float* targets, long long t_sbs, long long t_sn, const float* fmap_level,
                                     float* vol, int BS, int N, int C, int HW, int prec_mode, comet_stream_t stream) {
  COMET_REQUIRE(BS >= 0 && N >= 0 && C >= 1 && HW >= 1, "bad shape");
  COMET_REQUIRE(prec_mode == COMET_PREC_F32 || prec_mode == COMET_PREC_BF16_AUTOCAST, "bad prec_mode %d", prec_mode);
  if ((long long)BS * N == 0) return COMET_OK;
  COMET_REQUIRE(targets && fmap_level && vol, "null pointer");
  COMET_REQUIRE(BS <= 65535, "BS=%d exceeds gridDim.z; split the call", BS);
  dim3 grid((HW + TN - 1) / TN, (N + TM - 1) / TM, BS);
  corr_volume_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(targets, t_sbs, t_sn, fmap_level, vol, N, C, HW,
                                                             sqrtf((float)C), prec_mode == COMET_PREC_BF16_AUTOCAST);
  return launch_status("corr_volume_kernel");
}
