// mma.sync.m16n8k16 (bf16 -> f32) issue rate on sm_100a: cycles per HMMA per SM sub-partition with W warps per CTA (one CTA per SM),
// each warp cycling over 8 independent accumulators (no dependency stalls), and the latency of a dependent chain.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void k(int iters, int chain, long long* out, float* sink) {
  uint32_t a[4] = {threadIdx.x, 1u, 2u, 3u};
  float d[8][4] = {};
  __syncthreads();
  long long t0 = clock64();
  if (chain) {
    for (int i = 0; i < iters; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) mma(d[0], a, 5u, 7u);
  } else {
    for (int i = 0; i < iters; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) mma(d[j], a, 5u, 7u);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < 8; j++) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  if (s == 123.f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
int main() {
  long long* d; cudaMalloc(&d, 8); float* s; cudaMalloc(&s, 4); long long h;
  const int iters = 2000;
  for (int chain = 0; chain < 2; chain++)
    for (int warps : {1, 4, 8, 16, 32}) {
      k<<<148, warps * 32>>>(iters, chain, d, s);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double per = (double)h / (iters * 8);
      printf("%s warps/CTA=%2d: %.2f clk per HMMA per warp -> %.2f clk per HMMA per SM sub-partition (%.0f dense bf16 FLOP/clk/SM)\n",
             chain ? "dependent chain  " : "8 independent acc", warps, per, per / ((warps + 3) / 4), 4096.0 * warps / per);
    }
  return 0;
}
