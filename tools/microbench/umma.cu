// tcgen05.mma issue-rate probe: cycles per UMMA (M=128, K=16, bf16, SS mode) at N = 64/128/256, accumulating into one
// TMEM tile or alternating between independent tiles; one CTA per SM, one issuing thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "../../candidate-reranking-cir_b200/csrc/tcgen05_ptx.cuh"
using namespace tc;
__global__ void __launch_bounds__(128, 1) k(int N, int nacc, int iters, int bmn, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_s;
  __shared__ uint32_t tptr;
  const uint32_t sbase = smem_u32(smem), bar = smem_u32(&bar_s);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) {
    if (elect_one()) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    tmem_alloc<512>(smem_u32(&tptr));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = tptr;
  if (threadIdx.x < 32 && elect_one()) {
    const uint64_t ad = make_smem_desc_sw128(sbase), bd = make_smem_desc_sw128(sbase + 32768);
    const uint32_t idesc = make_idesc_bf16(128, N) | (bmn ? (1u << 16) : 0u);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int kk = 0; kk < 4; kk++) umma_bf16(tm + (it & (nacc - 1)) * N, ad + kk * 2, bd + (bmn ? kk * 128 : kk * 2), idesc, 1u);
    }
    umma_commit(bar);
    long long t1 = clock64();
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tcgen05_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tcgen05_fence_after(); tmem_dealloc<512>(tm); }
}
int main() {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  const int iters = 2000;
  for (int bmn = 0; bmn < 2; bmn++)
    for (int N : {16, 64, 96, 128, 160, 192, 224, 256})
      for (int nacc : {1, 2}) {
        if (N * nacc > 512 || (nacc == 2 && (N & (N - 1)))) continue;
        for (int grid : {1, 148}) {
          k<<<grid, 128, 65536 + 1024>>>(N, nacc, iters, bmn, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("B %s N=%3d accumulators=%d grid=%3d: issue %.1f cyc/UMMA, complete %.1f cyc/UMMA (floor %d)\n", bmn ? "MN-major" : "K-major ", N, nacc, grid,
                 (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), 128 * N / 256);
        }
      }
  return 0;
}
