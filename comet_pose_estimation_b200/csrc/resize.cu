// Bilinear resize with align_corners=True -- F.interpolate(x, size, mode="bilinear", align_corners=True) as the patch
// encoder of the fine tracker calls it (ShallowEncoder.forward, comet/models/track_modules/blocks.py:176-190: two
// 8->16 / 4->16 residual up-samplings and the final 16x16 -> 31x31 map that becomes the fine tracker's `fmaps`).
// ATen's CUDA kernel walks batch x channels inside every thread; at the fine tracker's shape (8192 patches x 32
// channels, 31x31 outputs) that is 961 threads doing 262144 iterations each: 192-314 ms per sequence on B200, 90 % of
// refine_track.  Here one thread produces one output element (NCHW) or one 4-channel group (channel-last): HBM-bound.
#include "comet_common.cuh"

namespace comet {

// ATen semantics (UpSample.cuh, area_pixel_compute_source_index with align_corners): src = dst * (in-1)/(out-1);
// i0 = (int)src, i1 = i0 + (i0 < in-1), w1 = src - i0, w0 = 1 - w1; value = wy0*(wx0*v00 + wx1*v01) + wy1*(wx0*v10 + wx1*v11).
struct ResizeTap {
  int i0, i1;
  float w0, w1;
  __device__ __forceinline__ void init(int dst, float scale, int in) {
    const float src = scale * (float)dst;
    i0 = (int)src;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    w1 = src - (float)i0;
    w0 = 1.f - w1;
  }
};

__device__ __forceinline__ float rs_load(const float* p, int i) { return __ldg(p + i); }
__device__ __forceinline__ float rs_load(const __nv_bfloat16* p, int i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void rs_store(float* p, long long i, float v) { p[i] = v; }
__device__ __forceinline__ void rs_store(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }

// T = float, or __nv_bfloat16 (the encoders under torch.autocast: float32 arithmetic on the bf16 values, result rounded
// to bf16 -- the same values as a float32 resize between two casts, without the two cast passes)
template <typename T>
__global__ void __launch_bounds__(256) upsample_nchw_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                            long long planes, int Hi, int Wi, int Ho, int Wo, float sy,
                                                            float sx) {
  const long long total = planes * Ho * Wo;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % Wo);
    const long long t = idx / Wo;
    const int yo = (int)(t % Ho);
    const long long pl = t / Ho;
    ResizeTap ty, tx;
    ty.init(yo, sy, Hi);
    tx.init(xo, sx, Wi);
    const T* p = in + pl * Hi * Wi;
    const float v00 = rs_load(p, ty.i0 * Wi + tx.i0), v01 = rs_load(p, ty.i0 * Wi + tx.i1);
    const float v10 = rs_load(p, ty.i1 * Wi + tx.i0), v11 = rs_load(p, ty.i1 * Wi + tx.i1);
    rs_store(out, idx, ty.w0 * (tx.w0 * v00 + tx.w1 * v01) + ty.w1 * (tx.w0 * v10 + tx.w1 * v11));
  }
}

// channel-last (N, H, W, C), C % 4 == 0: thread <-> (n, yo, xo, 4 channels)
__global__ void __launch_bounds__(256) upsample_cl_kernel(const float4* __restrict__ in, float4* __restrict__ out, long long N,
                                                          int C4, int Hi, int Wi, int Ho, int Wo, float sy, float sx) {
  const long long total = N * Ho * Wo * C4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % C4);
    long long t = idx / C4;
    const int xo = (int)(t % Wo);
    t /= Wo;
    const int yo = (int)(t % Ho);
    const long long n = t / Ho;
    ResizeTap ty, tx;
    ty.init(yo, sy, Hi);
    tx.init(xo, sx, Wi);
    const float4* p = in + n * (long long)Hi * Wi * C4 + c4;
    const float4 v00 = __ldg(p + (ty.i0 * Wi + tx.i0) * C4), v01 = __ldg(p + (ty.i0 * Wi + tx.i1) * C4);
    const float4 v10 = __ldg(p + (ty.i1 * Wi + tx.i0) * C4), v11 = __ldg(p + (ty.i1 * Wi + tx.i1) * C4);
    float4 o;
    o.x = ty.w0 * (tx.w0 * v00.x + tx.w1 * v01.x) + ty.w1 * (tx.w0 * v10.x + tx.w1 * v11.x);
    o.y = ty.w0 * (tx.w0 * v00.y + tx.w1 * v01.y) + ty.w1 * (tx.w0 * v10.y + tx.w1 * v11.y);
    o.z = ty.w0 * (tx.w0 * v00.z + tx.w1 * v01.z) + ty.w1 * (tx.w0 * v10.z + tx.w1 * v11.z);
    o.w = ty.w0 * (tx.w0 * v00.w + tx.w1 * v01.w) + ty.w1 * (tx.w0 * v10.w + tx.w1 * v11.w);
    out[idx] = o;
  }
}

}  // namespace comet

using namespace comet;

extern "C" int comet_upsample_bilinear_ac_f32(const float* in, float* out, long long N, int C, int Hi, int Wi, int Ho,
                                              int Wo, int layout, comet_stream_t stream) {
  COMET_REQUIRE(N >= 0 && C >= 0 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1, "bad shape");
  COMET_REQUIRE(layout == COMET_FMAPS_NCHW || layout == COMET_FMAPS_CHANNEL_LAST, "bad layout %d", layout);
  const long long total = N * C * (long long)Ho * Wo;
  if (total == 0) return COMET_OK;
  COMET_REQUIRE(in && out, "null pointer");
  const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  auto blocks_for = [](long long n) {
    long long b = (n + 255) / 256;
    if (b > 148LL * 64) b = 148LL * 64;
    return (unsigned)(b < 1 ? 1 : b);
  };
  if (layout == COMET_FMAPS_CHANNEL_LAST) {
    COMET_REQUIRE(C % 4 == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0,
                  "channel-last resize needs C %% 4 == 0 and 16-byte aligned buffers");
    upsample_cl_kernel<<<blocks_for(total / 4), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), N, C / 4, Hi, Wi, Ho, Wo, sy, sx);
    return launch_status("upsample_cl_kernel");
  }
  upsample_nchw_kernel<float><<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(in, out, N * C, Hi, Wi, Ho, Wo, sy, sx);
  return launch_status("upsample_nchw_kernel");
}

extern "C" int comet_upsample_bilinear_ac_bf16(const void* in, void* out, long long N, int C, int Hi, int Wi, int Ho, int Wo,
                                               comet_stream_t stream) {
  COMET_REQUIRE(N >= 0 && C >= 0 && Hi >= 1 && Wi >= 1 && Ho >= 1 && Wo >= 1, "bad shape");
  const long long total = N * C * (long long)Ho * Wo;
  if (total == 0) return COMET_OK;
  COMET_REQUIRE(in && out, "null pointer");
  const float sy = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  long long b = (total + 255) / 256;
  if (b > 148LL * 64) b = 148LL * 64;
  upsample_nchw_kernel<__nv_bfloat16><<<(unsigned)b, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), N * C, Hi, Wi, Ho, Wo, sy, sx);
  return launch_status("upsample_nchw_kernel<bf16>");
}

// ---- instance normalisation (+ optional ReLU) of the patch encoder ---------------------------------------------------
// nn.InstanceNorm2d(affine=False, track_running_stats=False, eps) as ShallowEncoder / ResidualBlock use it
// (blocks.py:128-131, modules.py:86-90): y = (x - mean) / sqrt(var + eps) per (sample, channel) plane, biased variance.
// ATen routes it through batch_norm over N*C "channels" (batch_norm_collect_statistics: 8.6 ms per sequence at 8192
// patches, the largest item of the encoder once the resizes are fixed).  Two passes over a plane that stays in L1/L2.
namespace comet {

// NCHW, small planes (the patch encoder's 16x16 .. 4x4 maps): one warp per (n, c) plane of HW contiguous elements
__global__ void __launch_bounds__(256) instance_norm_nchw_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                 long long planes, int HW, float eps, int relu) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long pl = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pl < planes; pl += nw) {
    const float* p = in + pl * HW;
    float s = 0.f;
    for (int i = lane; i < HW; i += 32) s += __ldg(p + i);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)HW;
    float v = 0.f;
    for (int i = lane; i < HW; i += 32) { const float d = __ldg(p + i) - mean; v = fmaf(d, d, v); }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)HW + eps);
    float* q = out + pl * HW;
    for (int i = lane; i < HW; i += 32) {
      float y = (__ldg(p + i) - mean) * rstd;
      q[i] = relu ? fmaxf(y, 0.f) : y;
    }
  }
}

// NCHW, large planes (BasicEncoder: 1024 .. 4096 planes of 128x128 .. 16x16 positions per 16-frame sequence): one CTA per
// plane.  A warp per plane leaves ~7 warps per SM walking 64 KB each -- 3.7 ms per sequence at 512x512 frames, a third of
// the encoder.  The plane is read from HBM once: it is parked in shared memory (float32, up to 128x128 positions) for the
// second and third pass; larger planes re-read global memory (L2).  T = float or __nv_bfloat16 (the encoder under
// torch.autocast: statistics in float32 of the bf16 values, result rounded to bf16 -- what ATen's batch-norm kernel does
// for a bf16 input -- without the two cast passes around a float32 kernel).
constexpr int INORM_SMEM_ELEMS = 16384;
__device__ __forceinline__ float inorm_load(const float* p, long long i) { return __ldg(p + i); }
__device__ __forceinline__ float inorm_load(const __nv_bfloat16* p, long long i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void inorm_store(float* p, long long i, float v) { p[i] = v; }
__device__ __forceinline__ void inorm_store(__nv_bfloat16* p, long long i, float v) { p[i] = __float2bfloat16_rn(v); }

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                 // red[] of the previous reduction has been read
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += red[w];
  return t;
}

template <typename T>
__global__ void __launch_bounds__(256) instance_norm_plane_kernel(const T* __restrict__ in, T* __restrict__ out, long long planes,
                                                                  int HW, float eps, int relu) {
  extern __shared__ __align__(16) float plane_smem[];
  __shared__ float red[8];
  const bool parked = HW <= INORM_SMEM_ELEMS;
  for (long long pl = blockIdx.x; pl < planes; pl += gridDim.x) {
    const T* p = in + pl * HW;
    T* q = out + pl * HW;
    float s = 0.f;
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float x = inorm_load(p, i);
      if (parked) plane_smem[i] = x;
      s += x;
    }
    const float mean = block_sum_256(s, red) / (float)HW;
    float v = 0.f;
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float d = (parked ? plane_smem[i] : inorm_load(p, i)) - mean;
      v = fmaf(d, d, v);
    }
    const float rstd = rsqrtf(block_sum_256(v, red) / (float)HW + eps);
    for (int i = threadIdx.x; i < HW; i += 256) {
      const float y = ((parked ? plane_smem[i] : inorm_load(p, i)) - mean) * rstd;
      inorm_store(q, i, relu ? fmaxf(y, 0.f) : y);
    }
    __syncthreads();               // the parked plane is dead before the next one moves in
  }
}

// channel-last (N, HW, C): one thread per (n, c), consecutive threads <-> consecutive channels (coalesced lines)
__global__ void __launch_bounds__(256) instance_norm_cl_kernel(const float* __restrict__ in, float* __restrict__ out, long long N,
                                                               int C, int HW, float eps, int relu) {
  const long long total = N * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long n = idx / C;
    const int c = (int)(idx - n * C);
    const float* p = in + n * (long long)HW * C + c;
    float s = 0.f;
    for (int i = 0; i < HW; ++i) s += __ldg(p + (long long)i * C);
    const float mean = s / (float)HW;
    float v = 0.f;
    for (int i = 0; i < HW; ++i) { const float d = __ldg(p + (long long)i * C) - mean; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(v / (float)HW + eps);
    float* q = out + n * (long long)HW * C + c;
    for (int i = 0; i < HW; ++i) {
      float y = (__ldg(p + (long long)i * C) - mean) * rstd;
      q[(long long)i * C] = relu ? fmaxf(y, 0.f) : y;
    }
  }
}

}  // namespace comet

namespace comet {
template <typename T>
static int launch_instance_norm_planes(const T* in, T* out, long long planes, int HW, float eps, int relu, comet_stream_t stream) {
  static bool configured[64] = {false};
  int dev = 0;
  COMET_CUDA(cudaGetDevice(&dev));
  const int smem = HW <= INORM_SMEM_ELEMS ? HW * 4 : 0;
  if (dev < 64 && !configured[dev]) {
    COMET_CUDA(cudaFuncSetAttribute(instance_norm_plane_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, INORM_SMEM_ELEMS * 4));
    configured[dev] = true;
  }
  long long blocks = planes < 148LL * 16 ? planes : 148LL * 16;
  instance_norm_plane_kernel<T><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(in, out, planes, HW, eps, relu);
  return launch_status("instance_norm_plane_kernel");
}
}  // namespace comet

extern "C" int comet_instance_norm_f32(const float* in, float* out, long long N, int C, int HW, int layout, int relu,
                                       float eps, comet_stream_t stream) {
  COMET_REQUIRE(N >= 0 && C >= 0 && HW >= 1, "bad shape");
  COMET_REQUIRE(layout == COMET_FMAPS_NCHW || layout == COMET_FMAPS_CHANNEL_LAST, "bad layout %d", layout);
  if (N * C == 0) return COMET_OK;
  COMET_REQUIRE(in && out, "null pointer");
  if (layout == COMET_FMAPS_CHANNEL_LAST) {
    long long blocks = (N * C + 255) / 256;
    if (blocks > 148LL * 64) blocks = 148LL * 64;
    comet::instance_norm_cl_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, N, C, HW, eps, relu);
    return comet::launch_status("instance_norm_cl_kernel");
  }
  if (HW >= 1024) return comet::launch_instance_norm_planes<float>(in, out, N * C, HW, eps, relu, stream);
  long long blocks = (N * C + 7) / 8;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  comet::instance_norm_nchw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, N * C, HW, eps, relu);
  return comet::launch_status("instance_norm_nchw_kernel");
}

extern "C" int comet_instance_norm_bf16(const void* in, void* out, long long N, int C, int HW, int relu, float eps,
                                        comet_stream_t stream) {
  COMET_REQUIRE(N >= 0 && C >= 0 && HW >= 1, "bad shape");
  if (N * C == 0) return COMET_OK;
  COMET_REQUIRE(in && out, "null pointer");
  return comet::launch_instance_norm_planes<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(in),
                                                           reinterpret_cast<__nv_bfloat16*>(out), N * C, HW, eps, relu, stream);
}

// ---- patch gather of refine_track (comet/models/refine_track.py:71-111) ----------------------------------------------
// The reference unfolds the image into all psize x psize windows (a view) and picks N of them per frame with advanced
// indexing, in (b s n) order.  Here one thread writes one output pixel (all C channels) of patch (b, n, s) -- the order
// the fine tracker consumes -- in channel-last memory, straight from the NCHW image.
namespace comet {
__global__ void __launch_bounds__(256) extract_patches_kernel(const float* __restrict__ images, const int* __restrict__ topleft,
                                                              float* __restrict__ out, int B, int S, int N, int C, int H, int W,
                                                              int P) {
  const long long total = (long long)B * N * S * P * P;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(idx % P);
    long long t = idx / P;
    const int py = (int)(t % P);
    t /= P;                                  // patch index in (b, n, s) order
    const int s = (int)(t % S);
    const long long bn = t / S;
    const int n = (int)(bn % N), b = (int)(bn / N);
    const int* tl = topleft + (((long long)b * S + s) * N + n) * 2;   // (B,S,N,2) = (x, y)
    // corners are clamped to the image (memory safety for arbitrary callers; refine_track already hands over
    // clamped corners, for which this is the identity)
    const int x = min(max(__ldg(tl), 0), W - P) + px, y = min(max(__ldg(tl + 1), 0), H - P) + py;
    const float* img = images + ((long long)b * S + s) * C * H * W + (long long)y * W + x;
    float* o = out + idx * C;
    for (int c = 0; c < C; ++c) o[c] = __ldg(img + (long long)c * H * W);
  }
}
}  // namespace comet

extern "C" int comet_extract_patches_f32(const float* images, const int* topleft, float* out, int B, int S, int N, int C,
                                         int H, int W, int P, comet_stream_t stream) {
  COMET_REQUIRE(B >= 0 && S >= 0 && N >= 0 && C >= 1 && P >= 1 && H >= P && W >= P, "bad shape");
  const long long total = (long long)B * N * S * P * P;
  if (total == 0) return COMET_OK;
  COMET_REQUIRE(images && topleft && out, "null pointer");
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  comet::extract_patches_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(images, topleft, out, B, S, N, C, H, W, P);
  return comet::launch_status("extract_patches_kernel");
}
