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

// ---- register-blocked fp32 SGEMM for the stage-I similarity tiles: C = A W^T (no epilogue), A [M,K], W [N,K]
//      both K-major with K % 16 == 0 and 16 B aligned rows.  128x128x16 tile, 256 threads, 8x8 outputs per thread.
constexpr int SG_T = 128, SG_K = 16;
__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C, int64_t M, int64_t N, int64_t K,
                int64_t lda, int64_t ldw, int64_t ldc) {
  __shared__ __align__(16) float sA[2][SG_K][SG_T + 4];
  __shared__ __align__(16) float sW[2][SG_K][SG_T + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * SG_T, n0 = (int64_t)blockIdx.x * SG_T;
  // loader: 128 rows x 16 k = 512 float4; each thread loads 2 float4 of A and 2 of W
  const int lrow = tid >> 2, lk = (tid & 3) * 4;            // rows lrow and lrow+64
  const int tx = tid & 15, ty = tid >> 4;                    // 16 x 16 threads; thread owns rows ty*4+{0..3} (+64), cols tx*4+{0..3} (+64)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
  float4 ra[2], rw[2];
  auto gload = [&](int64_t k0) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int64_t m = m0 + lrow + h * 64, n = n0 + lrow + h * 64;
      ra[h] = m < M ? *reinterpret_cast<const float4*>(A + m * lda + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[h] = n < N ? *reinterpret_cast<const float4*>(W + n * ldw + k0 + lk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int r = lrow + h * 64;
      sA[buf][lk + 0][r] = ra[h].x; sA[buf][lk + 1][r] = ra[h].y; sA[buf][lk + 2][r] = ra[h].z; sA[buf][lk + 3][r] = ra[h].w;
      sW[buf][lk + 0][r] = rw[h].x; sW[buf][lk + 1][r] = rw[h].y; sW[buf][lk + 2][r] = rw[h].z; sW[buf][lk + 3][r] = rw[h].w;
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int64_t k0 = 0; k0 < K; k0 += SG_K) {
    const bool more = k0 + SG_K < K;
    if (more) gload(k0 + SG_K);
#pragma unroll
    for (int k = 0; k < SG_K; k++) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[buf][k][64 + ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&sW[buf][k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&sW[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; jh++) {
      const int64_t n = n0 + jh * 64 + tx * 4;
      float* cp = C + m * ldc + n;
      if (n + 3 < N && (ldc & 3) == 0) {
        *reinterpret_cast<float4*>(cp) = make_float4(acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) if (n + j < N) cp[j] = acc[i][jh * 4 + j];
      }
    }
  }
}

// fp32 operands regardless of the context dtype (stage-I similarities stay true fp32 like the
// reference's fp32 matmul, src/validate.py:57,202).
int cir_gemm_simt_f32(cir_ctx* ctx, const cir_gemm_args* a) {
  if (a->M == 0 || a->N == 0 || a->batch == 0) return CIR_OK;
  if (a->batch == 1 && !a->bias && !a->residual && a->act == CIR_ACT_NONE && (a->K % SG_K) == 0 && (a->lda % 4) == 0 &&
      (a->ldw % 4) == 0 && ((uintptr_t)a->A & 15) == 0 && ((uintptr_t)a->W & 15) == 0 && ((uintptr_t)a->C & 15) == 0) {
    dim3 g((unsigned)((a->N + SG_T - 1) / SG_T), (unsigned)((a->M + SG_T - 1) / SG_T));
    if (g.y <= 65535) {
      sgemm_nt_kernel<<<g, 256, 0, ctx->stream>>>((const float*)a->A, (const float*)a->W, (float*)a->C, a->M, a->N, a->K, a->lda, a->ldw, a->ldc);
      CIR_LAUNCH_CHECK(ctx);
      return CIR_OK;
    }
  }
  dim3 grid((unsigned)((a->N + TN - 1) / TN), (unsigned)((a->M + TM - 1) / TM), (unsigned)a->batch);
  CIR_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "gemm_simt_f32: M=%lld too large for grid.y", (long long)a->M);
  cir_gemm_args p = *a;
  p.c_f32 = 1; p.res_f32 = 1;
  gemm_simt_kernel<float><<<grid, THREADS, 0, ctx->stream>>>(p);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}
