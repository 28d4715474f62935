// CUDA-core GEMM: C[b] = act(A[b] W[b]^T + bias[b]) (+ residual[b]).
// Used (a) as THE GEMM of the fp32 check mode (exact fp32 FMA accumulation, the 1e-4 score
// tolerance of BASELINE.json cannot be met with tf32 tensor cores) and (b) as an on-device
// cross-check for the tcgen05 kernel with bf16 operands.  Not a performance path.
#include "common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(THREADS)
gemm_simt_kernel(cir_gemm_args p) {
  __shared__ float sA[TK][TM + 4];
  __shared__ float sW[TK][TN + 4];
  const int b = blockIdx.z;
  const int64_t m0 = (int64_t)blockIdx.y * TM, n0 = (int64_t)blockIdx.x * TN;
  const T* A = (const T*)p.A + b * p.a_bstride;
  const T* W = (const T*)p.W + b * p.w_bstride;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;      // 16x16 threads, 4x4 outputs each
  float acc[4][4] = {};
  // loader mapping: 64 rows x 16 k = 1024 elements, 4 per thread, k fastest (coalesced along K)
  const int lr = tid / 4, lk = (tid % 4) * 4;
  for (int64_t k0 = 0; k0 < p.K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int64_t k = k0 + lk + i;
      int64_t m = m0 + lr, n = n0 + lr;
      sA[lk + i][lr] = (m < p.M && k < p.K) ? to_f32<T>(A[m * p.lda + k]) : 0.f;
      sW[lk + i][lr] = (n < p.N && k < p.K) ? to_f32<T>(W[n * p.ldw + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; k++) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { a[i] = sA[k][ty * 4 + i]; w[i] = sW[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float* bias = p.bias ? p.bias + b * p.bias_bstride : nullptr;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int64_t n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (p.act == CIR_ACT_GELU) v = gelu_erf(v);
      else if (p.act == CIR_ACT_RELU) v = fmaxf(v, 0.f);
      if (p.residual) {
        int64_t ro = b * p.res_bstride + m * p.ldres + n;
        v += p.res_f32 ? ((const float*)p.residual)[ro] : to_f32<T>(((const T*)p.residual)[ro]);
      }
      int64_t co = b * p.c_bstride + m * p.ldc + n;
      if (p.c_f32) ((float*)p.C)[co] = v;
      else ((T*)p.C)[co] = from_f32<T>(v);
    }
  }
}

}  // namespace

int cir_gemm_simt(cir_ctx* ctx, const cir_gemm_args* a) {
  if (a->M == 0 || a->N == 0 || a->batch == 0) return CIR_OK;
  dim3 grid((unsigned)((a->N + TN - 1) / TN), (unsigned)((a->M + TM - 1) / TM), (unsigned)a->batch);
  CIR_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "gemm_simt: M=%lld too large for grid.y", (long long)a->M);
  cir_gemm_args p = *a;
  if (ctx->dtype == CIR_DTYPE_F32) {
    p.c_f32 = 1; p.res_f32 = 1;
    gemm_simt_kernel<float><<<grid, THREADS, 0, ctx->stream>>>(p);
  } else {
    gemm_simt_kernel<bf16><<<grid, THREADS, 0, ctx->stream>>>(p);
  }
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}

// fp32 operands regardless of the context dtype (stage-I similarities stay true fp32 like the
// reference's fp32 matmul, src/validate.py:57,202).
int cir_gemm_simt_f32(cir_ctx* ctx, const cir_gemm_args* a) {
  if (a->M == 0 || a->N == 0 || a->batch == 0) return CIR_OK;
  dim3 grid((unsigned)((a->N + TN - 1) / TN), (unsigned)((a->M + TM - 1) / TM), (unsigned)a->batch);
  CIR_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "gemm_simt_f32: M=%lld too large for grid.y", (long long)a->M);
  cir_gemm_args p = *a;
  p.c_f32 = 1; p.res_f32 = 1;
  gemm_simt_kernel<float><<<grid, THREADS, 0, ctx->stream>>>(p);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}
