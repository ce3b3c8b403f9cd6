// Linear layers of the track-update transformer (EfficientUpdateFormer, comet/models/track_modules/blocks.py:205-348;
// AttnBlock / CrossAttnBlock / Mlp, comet/models/modules.py:119-154, :248-344) on the 5th-generation tensor cores:
//
//     Y[M,N] = act( X[M,K] . W[N,K]^T + bias[N] ) (+ residual[M,N])
//
// tcgen05.mma (kind::f16, bf16 operands, float32 accumulation in TMEM), operand tiles staged by TMA (2-D tiled loads,
// 128-byte swizzle, K-major) through an mbarrier ring, accumulator double-buffered in TMEM so that the epilogue of tile i
// overlaps the MMAs of tile i+1.  Persistent CTAs (one per SM), 10 warps: TMA producer | MMA issuer (one elected thread) |
// 8 epilogue warps (TMEM lane quarter = warp % 4, two warps per quarter, each owning 64 of the tile's 128 columns: the
// exact-erf GELU of fc1 costs ~25 instructions per element, and with one warp per quarter the epilogue took twice as long
// as the tile's MMAs).
//
// Precision.  Operands are "bf16 planes": a float32 tensor x is held as NP bf16 tensors p0 = bf16(x), p1 = bf16(x - p0),
// p2 = bf16(x - p0 - p1).  NP = 1 is what torch.autocast(bf16) gives the reference's nn.Linear (COMET's shipped
// mixed_precision: bf16); NP = 3 carries all 24 mantissa bits, and the product is accumulated over the six plane pairs
// (i, j), i + j <= 2 (the dropped pairs are below 2^-24 relative) -- float32-grade results from bf16 tensor-core passes,
// which the float32 parity mode of the tracker loop needs (rounding noise is amplified ~200x per refinement iteration).
// The tensor core adds each K=16 block product to its float32 accumulator with truncation, an error that grows linearly
// with the number of accumulation steps (measured 4e-9 * K relative with all six pairs in one accumulator), so the five
// low-order pairs -- 2^-8 and less of the result -- go to a SECOND accumulator and the epilogue adds the two in float32:
// the main accumulator sees K/16 steps instead of 6K/16 (float32 cuBLAS-level error, measured).
// The epilogue emits the result as float32 and / or directly as planes for the next GEMM.
#include "comet_common.cuh"

#include <cuda.h>

namespace comet {
namespace gemm {

constexpr int BM = 128, BN = 128, BK = 64;       // BN: widest tile (TMEM spacing); the tile actually used is Params::bn
                                                 // BK bf16 = one 128-byte swizzle row
constexpr int TILE_BYTES = BM * BK * 2;          // 16 KB (A tile == B tile)
constexpr int THREADS = 320;                     // 10 warps: TMA | MMA | 8 epilogue (float32-grade mode)
// Autocast mode (one MMA pass per K block) is bound by the epilogue's instruction issue -- GELU, conversion and staging are
// ~35 instructions per element against 1626 MMA clocks per 128x128x384 tile -- and two epilogue warps per scheduler leave
// half of the issue slots idle on dependency stalls: that mode runs SIXTEEN epilogue warps (four per TMEM lane quarter,
// one 32-column chunk each; <= 112 registers, so no register prefetch of the residual) on a 128 KB operand ring.
constexpr int THREADS_EW16 = 576;
constexpr int MAX_NP = 3;
constexpr int SMEM_BUDGET = 192 * 1024;          // operand ring (np = 3: two 96 KB stages; np = 1: six 32 KB stages)
constexpr int OUT_STAGE_BYTES = 4096;            // per epilogue warp: one 32 x 32 float32 chunk (or two 32 x 32 bf16 chunks)

struct Maps {
  CUtensorMap a[MAX_NP];   // X planes: {K, M} bf16, box {64, 128}
  CUtensorMap w[MAX_NP];   // W planes: {K, N} bf16, box {64, bn}
  CUtensorMap wh[MAX_NP];  // W planes with box {64, bn / 2}: the half tile a CTA of a pair loads and multicasts (Params::cluster)
  CUtensorMap o;           // float32 result: {N, M}, box {32, 32}, 128-byte swizzle (TMA store; valid if Params::tma_out)
  CUtensorMap op[MAX_NP];  // result planes: {N, M} bf16, box {32, 32}, 64-byte swizzle (valid if Params::tma_pl)
};

struct Params {
  int M, N, K;
  int np;                  // planes per operand (1 or 3)
  int nstage;              // ring depth: nstage * np * 32 KB <= SMEM_BUDGET
  int tiles_m, tiles_n, ktiles;
  int bk;                  // K elements per pipeline stage: 64 (128-byte swizzle rows) or 32 (64-byte rows: half-size
                           // stages, twice as many of them in flight -- the float32-grade mode moves 96 KB per K=64 block)
  int bn;                  // output tile width: 128, or 64 / 32 when 128 would leave most SMs without a tile (small M)
  const float* bias;       // [N] or null
  const float* resid;      // [M, resid_ld] float32 or null (added after the activation)
  long long resid_ld;
  float* out;              // [M, out_ld] float32 or null
  long long out_ld;
  __nv_bfloat16* outp[MAX_NP];   // planes of the result, [M, outp_ld] each, or null
  long long outp_ld;
  int out_np;
  int gelu;                // exact (erf) GELU, nn.GELU() default (modules.py:133)
  int tma_out, tma_pl;     // results leave through shared memory and bulk tensor stores (row pitches allow a tensor map)
  int cluster;             // launched as clusters of two CTAs that share every W tile (see the kernel)
  int pairs_m;             // ceil(tiles_m / 2)
};

// ------------------------------------------------------------------ PTX helpers (see corr_tc.cu for the rationale)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// bounded by WALL time (10 s), not by a poll count: a protocol bug traps instead of hanging the box
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
// shared -> global bulk tensor store of one box (coordinates {x = column, y = row}); rows / columns outside the tensor
// are clipped by the copy engine
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// the same load delivered to the same shared-memory offset (and signalling the mbarrier at the same offset) of every
// CTA of the cluster in `mask`
__device__ __forceinline__ void tma_load_2d_multicast(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(x), "r"(y), "h"(mask) : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `mask` once the MMAs issued so far have retired
__device__ __forceinline__ void tcgen05_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], both K-major
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// K-major swizzled operand tile: rows of `row_bytes`, 8-row groups 8 * row_bytes apart (SBO); LBO is not used by
// swizzled K-major layouts (canonical value 1).
// `row_bytes` = 128 (SWIZZLE_128B, layout type 2) or 64 (SWIZZLE_64B, layout type 4); SBO = 8 rows.
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t saddr, uint32_t row_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(((8 * row_bytes) >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;
  return d;
}
// kind::f16 instruction descriptor: D f32, A/B bf16, both K-major, M=128; N = Params::bn goes to bits 17-22 at run time
constexpr uint32_t IDESC_BASE = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32-byte global accesses (sm_100: ld / st .v8.f32): a lane of the epilogue owns one ROW of the tile, so every warp-wide
// access touches 32 different rows; with 16-byte accesses each touched 32-byte sector is only half used and the epilogue
// of a K = 384 GEMM (64 KB of float32 output per 1626 MMA clocks) is bound by sector operations in L1TEX.
__device__ __forceinline__ void st_global_v8(float* ptr, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void st_global_v8_b32(void* ptr, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const float* ptr, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]),
               "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(ptr));
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
// Same function with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7) on the special-function unit: used when the
// result is rounded to ONE bf16 plane anyway (autocast mode), where the epilogue, not the tensor pipe, bounds the fc1
// GEMM (ncu r02: tensor pipe 13 %, issue slots 50 % busy with erff's ~25 instructions per element).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  // rcp.approx (one special-function instruction, 1 ulp): __frcp_rn expands to a Newton step plus a branch to a
  // denormal slow path -- with BSSY / BSYNC a third of this function's instructions (ncu r02c); 1 + 0.33 z is in [1, inf)
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.f - poly * t * __expf(-z * z);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

// ------------------------------------------------------------------ the kernel
template <int EW>
__global__ void __launch_bounds__((2 + EW) * 32, 1) gemm_tc_kernel(const __grid_constant__ Maps maps, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // sixteen epilogue warps serve the one-plane (autocast) mode only: let the compiler drop the multi-plane code there
  const int NPK = EW == 16 ? 1 : p.np;
  const int ONP = EW == 16 ? (p.out_np ? 1 : 0) : p.out_np;
  const int a_tile_bytes = BM * p.bk * 2, b_tile_bytes = p.bn * p.bk * 2;
  const uint32_t row_bytes = (uint32_t)p.bk * 2;
  const int ksteps = p.bk / 16;
  const int stage_bytes = NPK * (a_tile_bytes + b_tile_bytes);   // [A planes][B planes]
  const uint32_t idesc = IDESC_BASE | ((uint32_t)(p.bn >> 3) << 17);
  uint8_t* out_stage = smem + ((p.nstage * stage_bytes + 1023) & ~1023);   // [EW epilogue warps][OUT_STAGE_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + EW * OUT_STAGE_BYTES);
  uint64_t* full = bars;                 // [nstage]  TMA -> MMA
  uint64_t* empty = full + p.nstage;     // [nstage]  MMA -> TMA
  uint64_t* acc_full = empty + p.nstage; // [2]       MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;    // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work list.  Alone: CTA b takes tiles b, b + grid, ... (m fastest).  As a PAIR (Params::cluster): cluster c takes units
  // c, c + clusters, ...; a unit is one column block and two neighbouring row blocks, one per CTA -- both CTAs need the
  // same W tile at the same time, so each loads HALF of it and multicasts it into both shared memories: the operand
  // traffic per CTA drops from A + W to A + W / 2 (the float32-grade mode is bound by that traffic).
  const int crank = p.cluster ? (int)cluster_ctarank() : 0;
  const int csz = p.cluster ? 2 : 1;
  const int ntiles = p.cluster ? p.pairs_m * p.tiles_n : p.tiles_m * p.tiles_n;
  const int t_first = blockIdx.x / csz, t_step = gridDim.x / csz, t_div = p.cluster ? p.pairs_m : p.tiles_m;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nstage; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], csz); }   // a stage is free when both CTAs are done with it
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(4 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (p.cluster) cluster_sync_all();     // the peer's barriers exist before anything is sent to them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int tile = t_first; tile < ntiles; tile += t_step) {
        const int mb = (tile % t_div) * csz + crank, nb = tile / t_div;
        for (int kb = 0; kb < p.ktiles; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* dst = smem + stage * stage_bytes;
          mbar_expect_tx(&full[stage], (uint32_t)stage_bytes);      // A from this CTA, W half from each CTA of the pair
          for (int i = 0; i < NPK; ++i) {
            tma_load_2d(&maps.a[i], &full[stage], dst + i * a_tile_bytes, kb * p.bk, mb * BM);
            if (p.cluster)
              tma_load_2d_multicast(&maps.wh[i], &full[stage], dst + NPK * a_tile_bytes + i * b_tile_bytes + crank * (b_tile_bytes / 2),
                                    kb * p.bk, nb * p.bn + crank * (p.bn / 2), (uint16_t)3);
            else
              tma_load_2d(&maps.w[i], &full[stage], dst + NPK * a_tile_bytes + i * b_tile_bytes, kb * p.bk, nb * p.bn);
          }
          if (++stage == (uint32_t)p.nstage) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int tile = t_first; tile < ntiles; tile += t_step) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      // TMEM columns of tile buffer `acc`: [main 128 | low-order pairs 128]
      const uint32_t d_main = tmem_base + acc * 2 * BN, d_low = d_main + BN;
      uint32_t accum_main = 0, accum_low = 0;
      for (int kb = 0; kb < p.ktiles; ++kb) {
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t a0 = smem_u32(smem + stage * stage_bytes), b0 = a0 + NPK * a_tile_bytes;
          // plane pairs (i, j), i + j < np: (0,0) into the main accumulator, the low-order ones into their own
          for (int sum = NPK - 1; sum >= 0; --sum) {
            for (int i = 0; i <= sum; ++i) {
              const int j = sum - i;
              const uint64_t ad = make_desc_kmajor(a0 + i * a_tile_bytes, row_bytes), bd = make_desc_kmajor(b0 + j * b_tile_bytes, row_bytes);
              if (sum == 0) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  if (k < ksteps) {
                    umma_bf16_ss(d_main, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, accum_main);   // +32 bytes per K=16
                    accum_main = 1;
                  }
                }
              } else {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  if (k < ksteps) {
                    umma_bf16_ss(d_low, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, accum_low);
                    accum_low = 1;
                  }
                }
              }
            }
          }
          if (p.cluster) tcgen05_commit_multicast(&empty[stage], (uint16_t)3);   // ... in both CTAs: the peer's W half lives here too
          else tcgen05_commit(&empty[stage]);                  // smem stage free once these MMAs retire
          if (kb + 1 == p.ktiles) tcgen05_commit(&acc_full[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == (uint32_t)p.nstage) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int wq = warp & 3;                        // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;              // which column group of the tile this warp owns (EW / 4 groups)
    const uint32_t lane_addr = ((uint32_t)(32 * wq) << 16);
    uint8_t* my_stage = out_stage + (warp - 2) * OUT_STAGE_BYTES;
    uint32_t acc = 0, acc_phase = 0;
    const bool vec_out = p.out && (p.out_ld % 8 == 0) && (((uintptr_t)p.out) % 32 == 0);
    const bool vec_res = p.resid && (p.resid_ld % 8 == 0) && (((uintptr_t)p.resid) % 32 == 0);
    const bool vec_pl = ONP > 0 && (p.outp_ld % 16 == 0) && (((uintptr_t)p.outp[0]) % 32 == 0) &&
                        (ONP < 2 || ((uintptr_t)p.outp[1]) % 32 == 0) && (ONP < 3 || ((uintptr_t)p.outp[2]) % 32 == 0);
    const bool vec_bias = p.bias && (((uintptr_t)p.bias) % 16 == 0) && (p.bn % 32 == 0);   // n0 is then a multiple of 32
    for (int tile = t_first; tile < ntiles; tile += t_step) {
      const int mb = (tile % t_div) * csz + crank, nb = tile / t_div;
      const int m = mb * BM + 32 * wq + lane;
      // The residual of a chunk does not depend on the accumulator: it is requested before the wait for it (first chunk
      // of the tile) / while the previous chunk's result is leaving (later chunks), so its DRAM latency -- a tile's
      // epilogue is otherwise a serial chain "accumulator, residual load, add, store" per chunk -- is hidden.  Only with 8
      // epilogue warps: 16 warps have 96 registers each.  (The bias, an L1 hit, is loaded where it is added: holding it
      // in registers across the wait made the 16-warp variant spill 174 bytes and cost fc1 4 us of 29.)
      const int cph = ((p.bn >> 5) + EW / 4 - 1) / (EW / 4);     // 32-column chunks per column group (bn = 96: 2 + 1 or 1 + 1 + 1)
      float rg[32];
      bool res_vec = false;
      auto prefetch = [&](int c) {
        const int n0 = nb * p.bn + 32 * c;
        res_vec = EW == 8 && vec_res && n0 + 32 <= p.N && m < p.M;
        if (res_vec) {
          const float* r = p.resid + (long long)m * p.resid_ld + n0;
#pragma unroll
          for (int i = 0; i < 4; ++i) ld_global_nc_v8(r + 8 * i, rg + 8 * i);
        }
      };
      if (32 * cph * chalf < p.bn && nb * p.bn + 32 * cph * chalf < p.N) prefetch(cph * chalf);
      mbar_wait(&acc_full[acc], acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = cph * chalf; c < cph * (chalf + 1) && 32 * c < p.bn; ++c) {
        const int n0 = nb * p.bn + 32 * c;
        if (n0 >= p.N) break;                      // warp-uniform
        float v[32];
        const bool full_chunk = n0 + 32 <= p.N;
        tmem_ld32(tmem_base + lane_addr + acc * 2 * BN + 32 * c, v);
        if (NPK > 1) {
          float lo[32];
          tmem_ld32(tmem_base + lane_addr + acc * 2 * BN + BN + 32 * c, lo);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += lo[i];
        } else {
          tmem_ld_wait();
        }
        const bool row_ok = m < p.M;
        if (row_ok) {
          if (vec_bias && full_chunk) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + i);
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
          } else if (p.bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += (full_chunk || n0 + i < p.N) ? __ldg(p.bias + n0 + i) : 0.f;
          }
          if (p.gelu) {
            if (NPK == 1 && !p.out) {     // autocast mode, planes-only output (fc1): bf16 rounding dominates
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_erf_fast(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
            }
          }
          if (res_vec) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += rg[i];
          } else if (p.resid) {
            const float* r = p.resid + (long long)m * p.resid_ld + n0;
            if (EW == 16 && vec_res && full_chunk) {     // (8 epilogue warps: this case was prefetched into rg)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float g[8];
                ld_global_nc_v8(r + 8 * i, g);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[8 * i + e] += g[e];
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (n0 + i < p.N) v[i] += __ldg(r + i);
            }
          }
        }
        // operands of the next chunk on their way while this one is converted and stored
        if (c + 1 < cph * (chalf + 1) && 32 * (c + 1) < p.bn && n0 + 32 < p.N) prefetch(c + 1);
        // ---- float32 result
        if (p.tma_out) {
          // the warp's 32 x 32 chunk goes to shared memory in the box layout of the tensor map (rows of 128 bytes, 16-byte
          // pieces XOR-swizzled with the row number: conflict-free for a lane-per-row writer) and leaves as ONE bulk
          // tensor store of full 128-byte lines; rows >= M and columns >= N are clipped by the copy engine
          if (lane == 0) bulk_wait_read();           // the previous box has been read out of the staging buffer
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(my_stage + lane * 128 + ((i ^ (lane & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(&maps.o, my_stage, n0, mb * BM + 32 * wq); bulk_commit(); }
        } else if (p.out && row_ok) {
          float* o = p.out + (long long)m * p.out_ld + n0;
          if (vec_out && full_chunk) {
#pragma unroll
            for (int i = 0; i < 4; ++i) st_global_v8(o + 8 * i, v + 8 * i);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (n0 + i < p.N) o[i] = v[i];
          }
        }
        // ---- result planes
        if (p.tma_pl) {
          for (int pl = 0; pl < ONP; ++pl) {
            uint8_t* st = my_stage + (pl & 1) * 2048;          // two 32 x 32 bf16 boxes fit: planes alternate
            if (pl != 1 || p.tma_out) {                        // (plane 1 follows plane 0 into the other half: no wait)
              if (lane == 0) bulk_wait_read();
              __syncwarp();
            }
            const bool last = pl + 1 == ONP;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t w[4];
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[8 * i + 2 * h], v[8 * i + 2 * h + 1]);
                w[h] = *reinterpret_cast<const uint32_t*>(&b2);
                if (!last) {
                  v[8 * i + 2 * h] -= __low2float(b2);          // residual for the next plane (exact in float32)
                  v[8 * i + 2 * h + 1] -= __high2float(b2);
                }
              }
              *reinterpret_cast<uint4*>(st + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&maps.op[pl], st, n0, mb * BM + 32 * wq); bulk_commit(); }
          }
        } else if (row_ok) {
          for (int pl = 0; pl < ONP; ++pl) {
            __nv_bfloat16* o = p.outp[pl] + (long long)m * p.outp_ld + n0;
            if (vec_pl && full_chunk && pl + 1 == ONP) {
              // last (or only: autocast) plane: no residual to carry
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                uint32_t w[8];
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                  const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[16 * i + 2 * h], v[16 * i + 2 * h + 1]);
                  w[h] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                st_global_v8_b32(o + 16 * i, w);
              }
            } else if (vec_pl && full_chunk) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                uint32_t w[8];
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                  const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[16 * i + 2 * h], v[16 * i + 2 * h + 1]);
                  w[h] = *reinterpret_cast<const uint32_t*>(&b2);
                  v[16 * i + 2 * h] -= __low2float(b2);          // residual for the next plane (exact in float32)
                  v[16 * i + 2 * h + 1] -= __high2float(b2);
                }
                st_global_v8_b32(o + 16 * i, w);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const __nv_bfloat16 b = __float2bfloat16_rn(v[i]);
                if (n0 + i < p.N) o[i] = b;
                v[i] -= __bfloat162float(b);
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_all();   // the bulk stores of this warp have been written before the CTA retires
  }

  tcgen05_fence_before();
  __syncthreads();
  if (p.cluster) cluster_sync_all();     // no CTA retires while its peer may still arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(4 * BN));
  }
}

// 2-D tensor map over a row-major bf16 matrix [rows, ld] (K contiguous): dims {K, rows}, box {64, 128}, 128-byte swizzle,
// zero fill outside (K tails and row tails of the last tile read as 0).
static int encode_kmajor(CUtensorMap* tm, const void* base, long long rows, long long K, long long ld, int box_rows, int bk) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return fail(COMET_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1, 1};
  CUresult cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(COMET_ERR_CUDA, "cuTensorMapEncodeTiled (gemm operand) failed with %d", (int)cr);
  return COMET_OK;
}

// 2-D tensor map over a row-major result matrix [rows, ld]: dims {cols, rows}, box {32, 32}; rows of the box are 128
// bytes (float32, 128-byte swizzle) or 64 bytes (bf16, 64-byte swizzle) -- the layouts the epilogue writes.
static int encode_result(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, bool f32) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return fail(COMET_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estride[2] = {1, 1};
  CUresult cr = enc(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim,
                    gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(COMET_ERR_CUDA, "cuTensorMapEncodeTiled (gemm result) failed with %d", (int)cr);
  return COMET_OK;
}

}  // namespace gemm

// ---- float32 -> bf16 planes (the A operand of a GEMM whose producer is not one of the kernels of this file) -------
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, long long x_ld, __nv_bfloat16* __restrict__ p0,
                                                            __nv_bfloat16* __restrict__ p1, __nv_bfloat16* __restrict__ p2,
                                                            long long p_ld, long long rows, int cols, int np) {
  const long long total = rows * (long long)cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const int c = (int)(i - r * cols);
    float v = __ldg(x + r * x_ld + c);
    const __nv_bfloat16 a = __float2bfloat16_rn(v);
    p0[r * p_ld + c] = a;
    if (np > 1) {
      v -= __bfloat162float(a);
      const __nv_bfloat16 b = __float2bfloat16_rn(v);
      p1[r * p_ld + c] = b;
      v -= __bfloat162float(b);
      p2[r * p_ld + c] = __float2bfloat16_rn(v);
    }
  }
}

}  // namespace comet

using namespace comet;

extern "C" int comet_split_planes_f32(const float* x, long long x_ld, void* planes, long long plane_stride, long long p_ld,
                                      long long rows, int cols, int np, comet_stream_t stream) {
  COMET_REQUIRE(rows >= 0 && cols >= 0 && (np == 1 || np == 3), "bad shape (rows=%lld cols=%d np=%d)", rows, cols, np);
  if (rows * cols == 0) return COMET_OK;
  COMET_REQUIRE(x && planes, "null pointer");
  __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(planes);
  long long blocks = (rows * cols + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  split_planes_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, x_ld, p, p + plane_stride, p + 2 * plane_stride, p_ld,
                                                                          rows, cols, np);
  return launch_status("split_planes_kernel");
}

// x planes: np bf16 matrices [M, x_ld] spaced x_plane_stride elements apart; w planes likewise [N, w_ld].
// out (float32, optional), out planes (bf16, optional, out_np in {0, 1, 3}), bias [N] / resid [M, resid_ld] optional.
extern "C" int comet_linear_tc(const void* x_planes, long long x_plane_stride, long long x_ld, const void* w_planes,
                               long long w_plane_stride, long long w_ld, int np, const float* bias, const float* resid,
                               long long resid_ld, float* out, long long out_ld, void* out_planes,
                               long long out_plane_stride, long long outp_ld, int out_np, int gelu, long long M, int N, int K,
                               comet_stream_t stream) {
  COMET_REQUIRE(np == 1 || np == 3, "np must be 1 or 3 (got %d)", np);
  COMET_REQUIRE(out_np == 0 || out_np == 1 || out_np == 3, "out_np must be 0, 1 or 3 (got %d)", out_np);
  COMET_REQUIRE(M >= 0 && N >= 1 && K >= 1 && M < (1LL << 31), "bad shape (M=%lld N=%d K=%d)", M, N, K);
  if (M == 0) return COMET_OK;
  COMET_REQUIRE(x_planes && w_planes && (out || (out_planes && out_np > 0)), "null pointer");
  COMET_REQUIRE(x_ld % 8 == 0 && w_ld % 8 == 0 && ((uintptr_t)x_planes % 16) == 0 && ((uintptr_t)w_planes % 16) == 0 &&
                    x_plane_stride % 8 == 0 && w_plane_stride % 8 == 0,
                "operand planes must be 16-byte aligned with row pitches that are multiples of 8 elements");
  const int sms = device_sm_count_if_sm100();
  if (!sms || !tensor_map_encoder()) return fail(COMET_ERR_UNSUPPORTED, "comet_linear_tc needs an sm_100 device with TMA descriptors");
  gemm::Maps maps;
  memset(&maps, 0, sizeof(maps));
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x_planes);
  const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(w_planes);
  gemm::Params p{};
  p.M = (int)M; p.N = N; p.K = K; p.np = np;
  p.tiles_m = (int)((M + gemm::BM - 1) / gemm::BM);
  // Output tile width: the one that needs the fewest rounds of persistent CTAs x tile cost.  128 columns by default; 64 /
  // 32 when 128 would leave most SMs without a tile (the virtual-track GEMMs have M = 1024 rows: 24 tiles of 128x128 on
  // 148 SMs, each MMA-bound for its whole K -- measured 27.6 us against 7 us of work per SM); 96 where it divides N and
  // saves a round: the M = 9216 x N = 384 GEMMs (out-projection, fc2) are 216 tiles of 128 = 2 rounds of which the
  // second is 46 % full, but 288 tiles of 96 = 2 rounds of 3/4 the cost.  The +24 columns stand for the per-tile fixed
  // cost (A tile re-read per column block, accumulator hand-over).
  p.bk = (np == 3 && option(COMET_OPT_GEMM_BK32)) ? 32 : 64;
  if (np == 3) {
    // float32-grade mode: the tile is operand-feed-bound and a narrower one re-reads the A planes more often -- measured
    // slower at 96 and 64 columns (fc2: 71.8 us at 128, 80.3 at 96, 102.4 at 64): narrow only to fill the SMs
    p.bn = 128;
    while (p.bn > 32 && (long long)p.tiles_m * ((N + p.bn - 1) / p.bn) * 10 < 7LL * sms) p.bn >>= 1;
  } else {
    long long best_cost = -1;
    const int widths[4] = {128, 96, 64, 32};
    for (int i = 0; i < 4; ++i) {
      const int bn = widths[i];
      if (bn == 96 && (N % 96 != 0 || !option(COMET_OPT_GEMM_BN96))) continue;
      const long long tiles = (long long)p.tiles_m * ((N + bn - 1) / bn);
      const long long cost = ((tiles + sms - 1) / sms) * (bn + 24);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; p.bn = bn; }
    }
  }
  for (int i = 0; i < np; ++i) {
    int rc = gemm::encode_kmajor(&maps.a[i], xp + i * x_plane_stride, M, K, x_ld, gemm::BM, p.bk);
    if (rc != COMET_OK) return rc;
    rc = gemm::encode_kmajor(&maps.w[i], wp + i * w_plane_stride, N, K, w_ld, p.bn, p.bk);
    if (rc != COMET_OK) return rc;
  }
  const int stage_bytes = np * (gemm::BM + p.bn) * p.bk * 2;
  const int ew = (np == 1 && out_np <= 1 && option(COMET_OPT_GEMM_EW16)) ? 16 : 8;   // epilogue warps (see THREADS_EW16)
  p.nstage = (ew == 16 ? 128 * 1024 : gemm::SMEM_BUDGET) / stage_bytes;
  if (p.nstage > 12) p.nstage = 12;
  p.tiles_n = (N + p.bn - 1) / p.bn;
  p.ktiles = (K + p.bk - 1) / p.bk;
  p.bias = bias; p.resid = resid; p.resid_ld = resid_ld; p.out = out; p.out_ld = out_ld;
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out_planes);
  for (int i = 0; i < out_np; ++i) p.outp[i] = op + i * out_plane_stride;
  p.outp_ld = outp_ld; p.out_np = out_planes ? out_np : 0; p.gelu = gelu;
  // results through bulk tensor stores where the row pitches allow a tensor map (16-byte multiples, 16-byte aligned bases)
  const int tma_mode = option(COMET_OPT_GEMM_TMA_STORE);   // bit 0: float32 result, bit 1: result planes
  if (tma_mode) {
    if ((tma_mode & 1) && out && out_ld % 4 == 0 && ((uintptr_t)out % 16) == 0) {
      int rc = gemm::encode_result(&maps.o, out, M, N, out_ld, true);
      if (rc != COMET_OK) return rc;
      p.tma_out = 1;
    }
    if ((tma_mode & 2) && p.out_np > 0 && outp_ld % 8 == 0 && ((uintptr_t)out_planes % 16) == 0 && out_plane_stride % 8 == 0) {
      for (int i = 0; i < out_np; ++i) {
        int rc = gemm::encode_result(&maps.op[i], p.outp[i], M, N, outp_ld, false);
        if (rc != COMET_OK) return rc;
      }
      p.tma_pl = 1;
    }
  }
  const int smem = ((p.nstage * stage_bytes + 1023) & ~1023) + ew * gemm::OUT_STAGE_BYTES + (2 * p.nstage + 4) * 8 + 16;
  // CTA pairs sharing the W tile (COMET_OPT_GEMM_PAIR: 1 = float32-grade mode for K >= 1024, 2 = autocast mode, 4 = any K).
  // Measured at M = 9216 (scripts/gemm_time.py): the long-K GEMM of the float32-grade mode gains (fc2, K = 1536:
  // 76.2 -> 69.9 us) -- its 24 K blocks per tile are where the operand feed binds --; the K = 384 GEMMs do not
  // (43.3 -> 43.8, 23.0 -> 24.7, 64.3 -> 65.7 us: six K blocks per tile leave them bound by pipeline fill and epilogue, and
  // the pair runs in lock-step), nor does the autocast mode (20.6 -> 22.6, 22.7 -> 24.3 us).  Needs at least as many
  // (row-block pair, column block) units as clusters, and a half tile of whole swizzle atoms.
  p.pairs_m = (p.tiles_m + 1) / 2;
  const int pair_mode = option(COMET_OPT_GEMM_PAIR);
  p.cluster = ((np == 3 ? (pair_mode & 1) : (pair_mode & 2)) && (K >= 1024 || (pair_mode & 4)) && p.tiles_m >= 2 &&
               p.bn % 16 == 0 && (long long)p.pairs_m * p.tiles_n >= sms / 2) ? 1 : 0;
  if (p.cluster) {
    for (int i = 0; i < np; ++i) {
      int rc = gemm::encode_kmajor(&maps.wh[i], wp + i * w_plane_stride, N, K, w_ld, p.bn / 2, p.bk);
      if (rc != COMET_OK) return rc;
    }
  }
  const long long ntiles = p.cluster ? (long long)p.pairs_m * p.tiles_n : (long long)p.tiles_m * p.tiles_n;
  const int grid = p.cluster ? (int)(ntiles < sms / 2 ? ntiles : sms / 2) * 2 : (int)(ntiles < sms ? ntiles : sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(ew == 16 ? gemm::THREADS_EW16 : gemm::THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.cluster ? 1 : 0;
  if (ew == 16) {
    COMET_CUDA(cudaFuncSetAttribute(gemm::gemm_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    COMET_CUDA(cudaLaunchKernelEx(&cfg, gemm::gemm_tc_kernel<16>, maps, p));
  } else {
    COMET_CUDA(cudaFuncSetAttribute(gemm::gemm_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    COMET_CUDA(cudaLaunchKernelEx(&cfg, gemm::gemm_tc_kernel<8>, maps, p));
  }
  return launch_status("gemm_tc_kernel");
}
