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

// Channel-last pyramid for small maps (the fine tracker's 31x31 patches; W <= 32).  One CTA per (frame, map):
// warp w streams the channel planes w, w+nw, ... from global memory (row pairs, fully coalesced, all loads of a
// plane in flight), pools 2x2 with shuffles, and parks the results in a shared tile [level][pos][C+1] (the +1
// keeps both the per-channel column accesses and the per-position row accesses bank-conflict free).  Deeper
// levels are pooled from the tile.  After one barrier the CTA writes every level as ONE contiguous channel-last
// block (BS, H_l, W_l, C) with coalesced stores.
constexpr int CL_MAX_ROWS = 16;  // H/2 <= 16
__global__ void __launch_bounds__(256, 4) pyramid_cl_kernel(const float* __restrict__ in, float* __restrict__ pyr,
                                                          int C, int H, int W, int L, Levels lv) {
  extern __shared__ float tile[];  // levels 1..L-1 back to back, each H_l*W_l rows of (C+1) floats
  const long long map = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int CP = C + 1;
  const int H1 = H / 2, W1 = W / 2;
  for (int c = warp; c < C; c += nw) {
    const float* plane = in + (map * C + c) * (long long)H * W;
    // level 1 straight from global memory: lane <-> input column
    float r0[CL_MAX_ROWS], r1[CL_MAX_ROWS];
#pragma unroll
    for (int y = 0; y < CL_MAX_ROWS; ++y) {
      const bool on = y < H1 && lane < W;
      r0[y] = on ? __ldg(plane + (2 * y) * W + lane) : 0.f;
      r1[y] = on ? __ldg(plane + (2 * y + 1) * W + lane) : 0.f;
    }
#pragma unroll
    for (int y = 0; y < CL_MAX_ROWS; ++y) {
      if (y < H1) {
        const int x2 = (2 * lane) & 31;
        const float a = __shfl_sync(0xffffffffu, r0[y], x2), b = __shfl_sync(0xffffffffu, r0[y], x2 + 1);
        const float cc = __shfl_sync(0xffffffffu, r1[y], x2), d = __shfl_sync(0xffffffffu, r1[y], x2 + 1);
        if (lane < W1) tile[(y * W1 + lane) * CP + c] = ((a + b) + (cc + d)) * 0.25f;
      }
    }
    __syncwarp();
    // deeper levels from the tile (this warp's own channel column)
    int Hi = H1, Wi = W1, in_off = 0;
    for (int l = 2; l < L; ++l) {
      const int Ho = Hi / 2, Wo = Wi / 2;
      const int out_off = in_off + Hi * Wi * CP;
      for (int i = lane; i < Ho * Wo; i += 32) {
        const int yo = i / Wo, xo = i - yo * Wo;
        const float* s4 = tile + in_off + ((2 * yo) * Wi + 2 * xo) * CP + c;
        tile[out_off + i * CP + c] = ((s4[0] + s4[CP]) + (s4[Wi * CP] + s4[(Wi + 1) * CP])) * 0.25f;
      }
      __syncwarp();
      in_off = out_off;
      Hi = Ho; Wi = Wo;
    }
  }
  __syncthreads();
  int Hl = H1, Wl = W1, off = 0;
  for (int l = 1; l < L; ++l) {
    float* dst = pyr + lv.off[l] + map * (long long)C * Hl * Wl;
    const int n = Hl * Wl * C;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int pos = i / C, c = i - pos * C;
      dst[i] = tile[off + pos * CP + c];
    }
    off += Hl * Wl * CP;
    Hl /= 2; Wl /= 2;
  }
}

// Specialisation for the fine tracker (31x31 patches, C = 32, L = 3: 31 -> 15 -> 7).  The generic kernel above is
// issue-bound at this shape (ncu: issue slots 80 % busy, DRAM 49 %); here everything is compile-time, each 2x2 pool
// costs one shuffle (vertical add in registers, horizontal add with the neighbour lane), level 2 and the write-out
// use shifts instead of divisions and 128-bit stores.  ~5 k warp instructions per map against a budget of ~30 k at
// the DRAM rate.
__global__ void __launch_bounds__(256, 4) pyramid_cl_fine_kernel(const float* __restrict__ in, float* __restrict__ pyr,
                                                                 long long off1, long long off2) {
  constexpr int W = 31, H1 = 15, H2 = 7, C = 32, CP = 33;
  constexpr int N1 = H1 * H1, N2 = H2 * H2;
  __shared__ float tile[(N1 + N2) * CP];
  const long long map = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool on = lane < W;
#pragma unroll 1
  for (int c = warp; c < C; c += 8) {
    const float* plane = in + (map * C + c) * (long long)(W * W) + (on ? lane : 0);
    float r[2 * H1];
#pragma unroll
    for (int y = 0; y < 2 * H1; ++y) r[y] = __ldg(plane + y * W);   // rows 0..29; lane 31 re-reads column 0 (unused)
#pragma unroll
    for (int y = 0; y < H1; ++y) {
      const float sv = r[2 * y] + r[2 * y + 1];
      const float sn = __shfl_down_sync(0xffffffffu, sv, 1);
      if (!(lane & 1) && lane < 2 * H1) tile[(y * H1 + (lane >> 1)) * CP + c] = (sv + sn) * 0.25f;
    }
  }
  __syncthreads();
  // level 2 from the level-1 tile: thread <-> (position, channel)
  for (int i = threadIdx.x; i < N2 * C; i += 256) {
    const int pos = i >> 5, c = i & 31;
    const int yo = pos / H2, xo = pos - yo * H2;
    const float* s4 = tile + ((2 * yo) * H1 + 2 * xo) * CP + c;
    tile[(N1 + pos) * CP + c] = ((s4[0] + s4[H1 * CP]) + (s4[CP] + s4[(H1 + 1) * CP])) * 0.25f;
  }
  // write-out, channel-last, one float4 per thread (conflict-free: bank = pos + c + k)
  float4* d1 = reinterpret_cast<float4*>(pyr + off1 + map * (long long)(N1 * C));
  for (int t = threadIdx.x; t < N1 * (C / 4); t += 256) {
    const float* sp = tile + (t >> 3) * CP + (t & 7) * 4;
    d1[t] = make_float4(sp[0], sp[1], sp[2], sp[3]);
  }
  __syncthreads();
  float4* d2 = reinterpret_cast<float4*>(pyr + off2 + map * (long long)(N2 * C));
  for (int t = threadIdx.x; t < N2 * (C / 4); t += 256) {
    const float* sp = tile + (N1 + (t >> 3)) * CP + (t & 7) * 4;
    d2[t] = make_float4(sp[0], sp[1], sp[2], sp[3]);
  }
}

// Channel-last INPUT (the producer's native layout, e.g. a cuDNN encoder run in torch.channels_last): level 0 is
// (BS, H, W, C) and never copied; levels 1..L-1 are written channel-last.  One CTA per map, thread <-> (output
// position, 4 channels): four 128-bit loads (one per 2x2 tap, lanes of a position read one contiguous line), one
// 128-bit store; the level stays in shared memory for the next one.  Row / column H-1 of an odd map is never read.
__global__ void __launch_bounds__(256, 2) pyramid_cl_in_kernel(const float4* __restrict__ in, float* __restrict__ pyr,
                                                               int C4, int H, int W, int L, Levels lv) {
  extern __shared__ float4 tile4[];  // levels 1..L-1 back to back, dense (pos, C4)
  const long long map = blockIdx.x;
  const float4* src = in + map * (long long)H * W * C4;
  int Hi = H, Wi = W, in_off = 0, out_off = 0;
  for (int l = 1; l < L; ++l) {
    const int Ho = Hi / 2, Wo = Wi / 2;
    const int items = Ho * Wo * C4;
    float4* dst = reinterpret_cast<float4*>(pyr + lv.off[l]) + map * (long long)items;
    for (int i0 = threadIdx.x; i0 < items; i0 += 4 * blockDim.x) {
      float4 t[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = min(i0 + u * (int)blockDim.x, items - 1);
        const int pos = i / C4, c4 = i - pos * C4;
        const int yo = pos / Wo, xo = pos - yo * Wo;
        const int base = ((2 * yo) * Wi + 2 * xo) * C4 + c4;
        if (l == 1) {
          t[u][0] = __ldg(src + base); t[u][1] = __ldg(src + base + C4);
          t[u][2] = __ldg(src + base + Wi * C4); t[u][3] = __ldg(src + base + (Wi + 1) * C4);
        } else {
          t[u][0] = tile4[in_off + base]; t[u][1] = tile4[in_off + base + C4];
          t[u][2] = tile4[in_off + base + Wi * C4]; t[u][3] = tile4[in_off + base + (Wi + 1) * C4];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * (int)blockDim.x;
        if (i < items) {
          float4 v;
          v.x = ((t[u][0].x + t[u][1].x) + (t[u][2].x + t[u][3].x)) * 0.25f;
          v.y = ((t[u][0].y + t[u][1].y) + (t[u][2].y + t[u][3].y)) * 0.25f;
          v.z = ((t[u][0].z + t[u][1].z) + (t[u][2].z + t[u][3].z)) * 0.25f;
          v.w = ((t[u][0].w + t[u][1].w) + (t[u][2].w + t[u][3].w)) * 0.25f;
          tile4[out_off + i] = v;
          dst[i] = v;
        }
      }
    }
    __syncthreads();
    in_off = out_off;
    out_off += items;
    Hi = Ho; Wi = Wo;
  }
}

// Channel-last input at the fine tracker's shape (31x31, C = 32, L = 3).  The register version of this kernel kept
// only ~110 KB per SM in flight (every in-flight byte of a load lives in a register) and reached 81 % of the DRAM
// rate; here the map streams through shared memory instead: two input rows (2 x 31 positions x 128 B = 7936 B,
// contiguous in a channel-last map) are one bulk async copy (cp.async.bulk -> mbarrier), ten of them in flight per
// CTA, two persistent CTAs per SM.  Threads pool a landed row pair (4 x LDS.128, conflict-free), keep level 1 in shared
// memory for level 2 and write both levels with 128-bit stores.
namespace bulk {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // a protocol bug must fail visibly, not hang the device: trap after 10 s of WALL time (not a poll count)
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 4095u) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 10000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
}  // namespace bulk

constexpr int PF_STAGES = 10;
constexpr int PF_ROWPAIR = 2 * 31 * 8;                 // float4 per row pair
constexpr int PF_N1 = 15 * 15 * 8, PF_N2 = 7 * 7 * 8;  // float4 items per pooled level
constexpr int PF_SMEM = PF_STAGES * PF_ROWPAIR * 16 + PF_N1 * 16 + PF_STAGES * 8;

__global__ void __launch_bounds__(128, 2) pyramid_cl_in_fine_kernel(const float4* __restrict__ in, float* __restrict__ pyr,
                                                                    long long off1, long long off2, int nmaps) {
  constexpr int W = 31, H1 = 15, H2 = 7, C4 = 8, RP = 15;   // RP: row pairs per map (row 30 is never read)
  extern __shared__ __align__(128) float4 pf_smem[];
  float4* ring = pf_smem;                                   // [PF_STAGES][PF_ROWPAIR]
  float4* t1 = ring + PF_STAGES * PF_ROWPAIR;               // level 1 of the current map
  uint64_t* full = reinterpret_cast<uint64_t*>(t1 + PF_N1);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < PF_STAGES; ++i) bulk::mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int my_maps = (nmaps - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // maps blockIdx.x, +gridDim.x, ...
  const long long total = (long long)my_maps * RP;
  auto issue = [&](long long g) {   // row pair g of this CTA's sequence -> stage g % PF_STAGES
    const long long k = g / RP;
    const int rp = (int)(g - k * RP);
    const long long map = blockIdx.x + k * gridDim.x;
    const int st = (int)(g % PF_STAGES);
    bulk::mbar_expect_tx(&full[st], PF_ROWPAIR * 16);
    bulk::copy_g2s(ring + st * PF_ROWPAIR, in + (map * (W * W) + (long long)rp * 2 * W) * C4, PF_ROWPAIR * 16, &full[st]);
  };
  if (tid == 0)
    for (long long g = 0; g < PF_STAGES && g < total; ++g) issue(g);

  auto pool = [](const float4& a, const float4& b, const float4& c, const float4& d) {
    return make_float4(((a.x + b.x) + (c.x + d.x)) * 0.25f, ((a.y + b.y) + (c.y + d.y)) * 0.25f,
                       ((a.z + b.z) + (c.z + d.z)) * 0.25f, ((a.w + b.w) + (c.w + d.w)) * 0.25f);
  };
  const int xo = tid >> 3, c4 = tid & 7;    // this thread's level-1 item within a row (tid < 120)
  long long g = 0;
  for (int k = 0; k < my_maps; ++k) {
    const long long map = blockIdx.x + (long long)k * gridDim.x;
    float4* d1 = reinterpret_cast<float4*>(pyr + off1) + map * PF_N1;
    float4* d2 = reinterpret_cast<float4*>(pyr + off2) + map * PF_N2;
    for (int rp = 0; rp < RP; ++rp, ++g) {
      const int st = (int)(g % PF_STAGES);
      bulk::mbar_wait(&full[st], (uint32_t)((g / PF_STAGES) & 1));
      if (tid < H1 * C4) {
        const float4* s4 = ring + st * PF_ROWPAIR + (2 * xo) * C4 + c4;
        const float4 v = pool(s4[0], s4[C4], s4[W * C4], s4[(W + 1) * C4]);
        t1[rp * (H1 * C4) + tid] = v;
        d1[rp * (H1 * C4) + tid] = v;
      }
      __syncthreads();   // the stage is drained (and this level-1 row is visible)
      if (tid == 0 && g + PF_STAGES < total) issue(g + PF_STAGES);
    }
    for (int i = tid; i < PF_N2; i += 128) {
      const int pos = i >> 3, cc = i & 7;
      const int yo = pos / H2, xx = pos - yo * H2;
      const float4* s4 = t1 + ((2 * yo) * H1 + 2 * xx) * C4 + cc;
      d2[i] = pool(s4[0], s4[C4], s4[H1 * C4], s4[(H1 + 1) * C4]);
    }
    __syncthreads();     // level 2 has read t1 before the next map overwrites it
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

extern "C" int comet_pyramid_cl_f32(const float* fmaps, float* pyr, int BS, int C, int H, int W, int L,
                                    int fmaps_layout, comet_stream_t stream) {
  COMET_REQUIRE(BS >= 0 && C >= 1 && H >= 1 && W >= 1, "bad shape");
  COMET_REQUIRE(fmaps_layout == COMET_FMAPS_NCHW || fmaps_layout == COMET_FMAPS_CHANNEL_LAST, "bad fmaps_layout %d",
                fmaps_layout);
  if (fmaps_layout == COMET_FMAPS_CHANNEL_LAST) {
    COMET_REQUIRE(L >= 1 && L <= COMET_MAX_LEVELS, "num_levels must be in [1, %d] (got %d)", COMET_MAX_LEVELS, L);
    COMET_REQUIRE((H >> (L - 1)) >= 1 && (W >> (L - 1)) >= 1, "map %dx%d too small for %d levels", H, W, L);
    COMET_REQUIRE(C % 4 == 0, "channel-last input needs C %% 4 == 0 (got %d)", C);
    if (L == 1 || BS == 0) return COMET_OK;
    COMET_REQUIRE(fmaps && pyr, "null pointer");
    COMET_REQUIRE(((uintptr_t)fmaps % 16) == 0 && ((uintptr_t)pyr % 16) == 0, "channel-last pyramid needs 16-byte aligned buffers");
    Levels lv = make_levels(BS, C, H, W, L);
    if (C == 32 && H == 31 && W == 31 && L == 3) {
      int sms = device_sm_count_if_sm100();
      if (sms <= 0) sms = 148;
      const int grid = BS < 2 * sms ? BS : 2 * sms;
      static bool attr_set = false;
      if (!attr_set) {
        COMET_CUDA(cudaFuncSetAttribute(pyramid_cl_in_fine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM));
        attr_set = true;
      }
      pyramid_cl_in_fine_kernel<<<grid, 128, PF_SMEM, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(fmaps), pyr,
                                                                             lv.off[1], lv.off[2], BS);
      return launch_status("pyramid_cl_in_fine_kernel");
    }
    size_t items = 0;
    for (int l = 1; l < L; ++l) items += (size_t)lv.H[l] * lv.W[l];
    const size_t smem = items * C * sizeof(float);
    COMET_REQUIRE(smem <= 200 * 1024, "pooled levels of one map (%zu bytes) do not fit shared memory", smem);
    if (smem > 48 * 1024)
      COMET_CUDA(cudaFuncSetAttribute(pyramid_cl_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pyramid_cl_in_kernel<<<BS, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(fmaps), pyr, C / 4, H, W,
                                                                   L, lv);
    return launch_status("pyramid_cl_in_kernel");
  }
  COMET_REQUIRE(L >= 1 && L <= COMET_MAX_LEVELS, "num_levels must be in [1, %d] (got %d)", COMET_MAX_LEVELS, L);
  COMET_REQUIRE((H >> (L - 1)) >= 1 && (W >> (L - 1)) >= 1, "map %dx%d too small for %d levels", H, W, L);
  COMET_REQUIRE(W <= 32 && H <= 2 * CL_MAX_ROWS + 1, "channel-last pyramid needs W <= 32 and H <= 33");
  if (L == 1 || BS == 0) return COMET_OK;
  COMET_REQUIRE(fmaps && pyr, "null pointer");
  Levels lv = make_levels(BS, C, H, W, L);
  if (C == 32 && H == 31 && W == 31 && L == 3 && ((uintptr_t)pyr % 16) == 0) {
    pyramid_cl_fine_kernel<<<BS, 256, 0, (cudaStream_t)stream>>>(fmaps, pyr, lv.off[1], lv.off[2]);
    return launch_status("pyramid_cl_fine_kernel");
  }
  size_t rows = 0;
  for (int l = 1; l < L; ++l) rows += (size_t)lv.H[l] * lv.W[l];
  const size_t smem = rows * (C + 1) * sizeof(float);
  COMET_REQUIRE(smem <= 160 * 1024, "channel-last pyramid tile does not fit in shared memory (C=%d)", C);
  if (smem > 48 * 1024)
    COMET_CUDA(cudaFuncSetAttribute(pyramid_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = C >= 8 ? 256 : 32 * C;
  pyramid_cl_kernel<<<BS, threads, smem, (cudaStream_t)stream>>>(fmaps, pyr, C, H, W, L, lv);
  return launch_status("pyramid_cl_kernel");
}
