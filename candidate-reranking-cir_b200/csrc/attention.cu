// softmax(Q K^T * scale + mask) V per (batch, head), head dim 64.
//
// attention_simt_kernel: generic CUDA-core kernel (fp32 math, T = float | bf16 storage) used by the
// fp32 check mode for every attention, and by the bf16 mode where no tensor-core specialisation
// exists yet.  One CTA = one (batch, head, 32-query tile); scores for the whole key range live in
// shared memory (Lk <= 1024), so softmax is exact two-pass like the reference
// (src/nlvr_encoder.py:193-199, src/vit.py:74-75).
#include "common.cuh"

namespace {

constexpr int QT = 32;      // queries per CTA
constexpr int KC = 64;      // keys per shared-memory chunk
constexpr int DH = CIR_HEAD_DIM;
constexpr int THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(THREADS)
attention_simt_kernel(cir_attn_args p) {
  extern __shared__ float smem[];
  const int lk_pad = (p.Lk + 3) & ~3;
  float* sQ = smem;                         // [QT][DH]
  float* sKV = sQ + QT * DH;                // [KC][DH+1]
  float* sS = sKV + KC * (DH + 1);          // [QT][lk_pad+1]
  const int ss = lk_pad + 1;

  const int b = blockIdx.x, h = blockIdx.y, q0 = blockIdx.z * QT;
  const int tid = threadIdx.x;
  const int kvb = p.kv_index ? p.kv_index[b] : b;
  const T* Q = (const T*)p.q + (int64_t)b * p.q_bs + h * DH;
  const T* K = (const T*)p.k + (int64_t)kvb * p.k_bs + h * DH;
  const T* V = (const T*)p.v + (int64_t)kvb * p.v_bs + h * DH;
  T* O = (T*)p.o + (int64_t)b * p.o_bs + h * DH;
  const int32_t* mask = p.key_mask ? p.key_mask + (int64_t)(p.mask_index ? p.mask_index[b] : b) * p.Lk : nullptr;

  for (int i = tid; i < QT * DH; i += THREADS) {
    const int q = i / DH, d = i % DH;
    sQ[i] = (q0 + q < p.Lq) ? to_f32<T>(Q[(int64_t)(q0 + q) * p.q_rs + d]) : 0.f;
  }
  const int q = tid / 8, sub = tid % 8;
  // ---- phase 1: S = Q K^T * scale + mask
  for (int k0 = 0; k0 < p.Lk; k0 += KC) {
    __syncthreads();
    for (int i = tid; i < KC * DH; i += THREADS) {
      const int kk = i / DH, d = i % DH;
      sKV[kk * (DH + 1) + d] = (k0 + kk < p.Lk) ? to_f32<T>(K[(int64_t)(k0 + kk) * p.k_rs + d]) : 0.f;
    }
    __syncthreads();
    float acc[8] = {};
#pragma unroll 8
    for (int d = 0; d < DH; d++) {
      const float qv = sQ[q * DH + d];
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] = fmaf(qv, sKV[(sub + 8 * j) * (DH + 1) + d], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int kk = k0 + sub + 8 * j;
      if (kk < p.Lk) {
        float s = acc[j] * p.scale;
        if (mask) s += (1.0f - (float)mask[kk]) * -10000.0f;
        sS[q * ss + kk] = s;
      }
    }
  }
  __syncthreads();
  // ---- phase 2: exact softmax per row (one warp handles 4 rows)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp * 4; r < warp * 4 + 4; r++) {
      float m = -INFINITY;
      for (int k = lane; k < p.Lk; k += 32) m = fmaxf(m, sS[r * ss + k]);
      m = warp_max(m);
      float sum = 0.f;
      for (int k = lane; k < p.Lk; k += 32) { float e = expf(sS[r * ss + k] - m); sS[r * ss + k] = e; sum += e; }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int k = lane; k < p.Lk; k += 32) sS[r * ss + k] *= inv;
    }
  }
  // ---- phase 3: O = P V
  float o[8] = {};
  const int d0 = sub * 8;
  for (int k0 = 0; k0 < p.Lk; k0 += KC) {
    __syncthreads();
    for (int i = tid; i < KC * DH; i += THREADS) {
      const int kk = i / DH, d = i % DH;
      sKV[kk * (DH + 1) + d] = (k0 + kk < p.Lk) ? to_f32<T>(V[(int64_t)(k0 + kk) * p.v_rs + d]) : 0.f;
    }
    __syncthreads();
    const int kn = min(KC, p.Lk - k0);
    for (int kk = 0; kk < kn; kk++) {
      const float pv = sS[q * ss + k0 + kk];
#pragma unroll
      for (int j = 0; j < 8; j++) o[j] = fmaf(pv, sKV[kk * (DH + 1) + d0 + j], o[j]);
    }
  }
  if (q0 + q < p.Lq) {
    T* op = O + (int64_t)(q0 + q) * p.o_rs + d0;
#pragma unroll
    for (int j = 0; j < 8; j++) op[j] = from_f32<T>(o[j]);
  }
}

size_t simt_smem_bytes(int Lk) {
  const int lk_pad = (Lk + 3) & ~3;
  return sizeof(float) * (size_t)(QT * DH + KC * (DH + 1) + QT * (lk_pad + 1));
}

}  // namespace

extern "C" int cir_attention(cir_ctx* ctx, const cir_attn_args* a) {
  if (a->B == 0 || a->Lq == 0) return CIR_OK;
  CIR_CHECK_ARG(a->Lk >= 1 && a->Lk <= 1024, "attention: Lk=%d out of range [1,1024]", a->Lk);
  CIR_CHECK_ARG(a->H >= 1 && a->H <= 65535, "attention: bad head count %d", a->H);
  const size_t smem = simt_smem_bytes(a->Lk);
  dim3 grid((unsigned)a->B, (unsigned)a->H, (unsigned)((a->Lq + QT - 1) / QT));
  if (ctx->dtype == CIR_DTYPE_F32) {
    CIR_CUDA(cudaFuncSetAttribute(attention_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attention_simt_kernel<float><<<grid, THREADS, smem, ctx->stream>>>(*a);
  } else {
    CIR_CUDA(cudaFuncSetAttribute(attention_simt_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attention_simt_kernel<bf16><<<grid, THREADS, smem, ctx->stream>>>(*a);
  }
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}
