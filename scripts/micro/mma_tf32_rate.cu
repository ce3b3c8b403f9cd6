// Issue-rate microbenchmark: legacy-path tensor instruction mma.sync.m16n8k8 (TF32) and m16n8k16 (BF16) on sm_100a,
// W warps per SM, 8 independent accumulators per warp.  Prints multiply-adds per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32_rate mma_tf32_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int kind = 0; kind < 2; ++kind)
    for (int warps : {4, 8, 16}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0) k<0><<<148, warps * 32>>>(out, iters); else k<1><<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double macs = (double)148 * warps * iters * 8 * (kind == 0 ? 16 * 8 * 8 : 16 * 8 * 16);
      printf("%s warps/SM %2d: %.3f ms, %.1f TMAC/s, %.0f MAC/clk/SM at %d MHz nominal\n", kind == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16",
             warps, ms, macs / ms * 1e-9, macs / (ms * 1e-3) / 148 / (clk_khz * 1e3), clk_khz / 1000);
    }
  return 0;
}
