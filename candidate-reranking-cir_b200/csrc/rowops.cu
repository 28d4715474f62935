// Bandwidth-bound row kernels: LayerNorm (+residual, twin gains), embeddings, gathers, casts,
// L2-normalise, ViT patchify/assemble, score-head dot.  One warp per 768-wide row, lane-interleaved
// 16-byte vector accesses (a warp instruction covers 512 contiguous bytes), fp32 statistics.
#include "common.cuh"

namespace {

constexpr int D = CIR_HIDDEN;          // 768
constexpr int PER_LANE = D / 32;       // 24 elements per lane (three lane-interleaved groups of 8)
constexpr int WARPS = 8;

// Lane-interleaved row mapping: lane l owns the three 8-element groups g = l, l + 32, l + 64 of a 768-wide row (columns 8g..8g+7),
// v[8i + j] <-> column 8 (l + 32 i) + j.  A warp-wide 16 B bf16 access then covers 512 contiguous bytes (every 32 B sector fully
// used by one instruction); fp32 rows take two 16 B accesses per group (32 contiguous bytes per lane).
template <typename T> __device__ __forceinline__ void load_row(const T* row, int lane, float (&v)[PER_LANE]);
template <> __device__ __forceinline__ void load_row<float>(const float* row, int lane, float (&v)[PER_LANE]) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float4* p = reinterpret_cast<const float4*>(row + (lane + 32 * i) * 8);
    const float4 a = p[0], b = p[1];
    v[8 * i] = a.x; v[8 * i + 1] = a.y; v[8 * i + 2] = a.z; v[8 * i + 3] = a.w;
    v[8 * i + 4] = b.x; v[8 * i + 5] = b.y; v[8 * i + 6] = b.z; v[8 * i + 7] = b.w;
  }
}
template <> __device__ __forceinline__ void load_row<bf16>(const bf16* row, int lane, float (&v)[PER_LANE]) {
  uint4 t[3];
#pragma unroll
  for (int i = 0; i < 3; i++) t[i] = *reinterpret_cast<const uint4*>(row + (lane + 32 * i) * 8);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t[i]);
#pragma unroll
    for (int q = 0; q < 4; q++) { float2 f = __bfloat1622float2(h[q]); v[8 * i + 2 * q] = f.x; v[8 * i + 2 * q + 1] = f.y; }
  }
}
template <typename T> __device__ __forceinline__ void store_row(T* row, int lane, const float (&v)[PER_LANE]);
template <> __device__ __forceinline__ void store_row<float>(float* row, int lane, const float (&v)[PER_LANE]) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float4* p = reinterpret_cast<float4*>(row + (lane + 32 * i) * 8);
    p[0] = make_float4(v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3]);
    p[1] = make_float4(v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]);
  }
}
template <> __device__ __forceinline__ void store_row<bf16>(bf16* row, int lane, const float (&v)[PER_LANE]) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int q = 0; q < 4; q++) h[q] = __floats2bfloat162_rn(v[8 * i + 2 * q], v[8 * i + 2 * q + 1]);
    *reinterpret_cast<uint4*>(row + (lane + 32 * i) * 8) = t;
  }
}

__device__ __forceinline__ void layernorm24(float (&v)[PER_LANE], const float* gamma, const float* beta, float eps, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; i++) s += v[i];
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; i++) { float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  float g[PER_LANE], b[PER_LANE];
  load_row<float>(gamma, lane, g);
  load_row<float>(beta, lane, b);
#pragma unroll
  for (int i = 0; i < PER_LANE; i++) v[i] = fmaf((v[i] - mean) * rstd, g[i], b[i]);
}

// packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(uint64_t v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi2(uint64_t v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// The raw 24 elements of one row held by this lane, loaded now and unpacked later (so that the next row's loads are in
// flight while the current row is reduced).
template <typename T> struct RawRow;
template <> struct RawRow<bf16> {
  uint4 q[3];
  __device__ __forceinline__ void load(const bf16* row, int lane) {
#pragma unroll
    for (int i = 0; i < 3; i++) q[i] = *reinterpret_cast<const uint4*>(row + (lane + 32 * i) * 8);
  }
  __device__ __forceinline__ void unpack(float (&v)[PER_LANE]) const {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q[i]);
#pragma unroll
      for (int k = 0; k < 4; k++) { float2 f = __bfloat1622float2(h[k]); v[8 * i + 2 * k] = f.x; v[8 * i + 2 * k + 1] = f.y; }
    }
  }
};
template <> struct RawRow<float> {
  float4 q[6];
  __device__ __forceinline__ void load(const float* row, int lane) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float4* p = reinterpret_cast<const float4*>(row + (lane + 32 * i) * 8);
      q[2 * i] = p[0]; q[2 * i + 1] = p[1];
    }
  }
  __device__ __forceinline__ void unpack(float (&v)[PER_LANE]) const {
#pragma unroll
    for (int i = 0; i < 6; i++) { v[4 * i] = q[i].x; v[4 * i + 1] = q[i].y; v[4 * i + 2] = q[i].z; v[4 * i + 3] = q[i].w; }
  }
};

// y[r] = LayerNorm(x[r % x_rows] + res[r]) * gamma[g] + beta[g].  One warp walks LN_ROWS consecutive rows: gamma / beta stay
// in registers (they change only at a group edge), and the loads of row i+1 are issued before row i is reduced, so a warp
// always has a row of HBM reads in flight.  The single-row-per-warp version re-read gamma/beta through L1 for every row
// (twice the payload bytes) and measured 4.3 TB/s; the arithmetic (operation order) is unchanged.
constexpr int LN_ROWS = 4;
template <typename TX, typename TR, typename TY>
__global__ void __launch_bounds__(WARPS * 32)
add_layernorm_kernel(const TX* __restrict__ x, int64_t x_rows, const TR* __restrict__ res, const float* __restrict__ gamma,
                     const float* __restrict__ beta, int64_t rows_per_group, TY* __restrict__ y, int64_t rows, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t r0 = ((int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5)) * LN_ROWS;
  if (r0 >= rows) return;
  const int n = (int)(rows - r0 < LN_ROWS ? rows - r0 : LN_ROWS);
  RawRow<TX> xr;
  RawRow<TR> rr;
  // row -> x row (r % x_rows) and parameter group (r / rows_per_group) advance incrementally: one 64-bit division each per warp,
  // not per row (a 64-bit divide is ~100 instructions; two per row were a third of the kernel's instruction stream)
  // (32-bit: the launcher guarantees rows < 2^31; the common cases -- no broadcast, one group -- need no division at all)
  const uint32_t r32 = (uint32_t)r0;
  int64_t xi = (uint64_t)x_rows > (uint64_t)r0 ? r0 : (int64_t)(r32 % (uint32_t)x_rows);
  int64_t gr = (uint64_t)rows_per_group > (uint64_t)r0 ? 0 : (int64_t)(r32 / (uint32_t)rows_per_group);
  int64_t gleft = (gr + 1) * rows_per_group - r0;            // rows left in the current group, this one included
  xr.load(x + xi * D, lane);
  if (res) rr.load(res + r0 * D, lane);
  float g[PER_LANE], b[PER_LANE];
  bool fresh = true;                                          // gamma / beta must be (re)loaded
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    const int64_t r = r0 + i;
    float v[PER_LANE];
    xr.unpack(v);
    if (res) {
      float t[PER_LANE];
      rr.unpack(t);
#pragma unroll
      for (int k = 0; k < PER_LANE; k++) v[k] += t[k];
    }
    if (i + 1 < n) {                                          // next row's reads go out before this row's reductions
      xi = xi + 1 == x_rows ? 0 : xi + 1;
      xr.load(x + xi * D, lane);
      if (res) rr.load(res + (r + 1) * D, lane);
    }
    if (fresh) {                                              // warp-uniform
      fresh = false;
      load_row<float>(gamma + gr * D, lane, g);
      load_row<float>(beta + gr * D, lane, b);
    }
    if (--gleft == 0) { gr++; gleft = rows_per_group; fresh = true; }
    // statistics and normalisation on packed fp32 pairs (FADD2 / FFMA2 / FMUL2: one issue slot for two IEEE fp32 operations; the
    // kernel is paced by its ~280 instructions per row at 16 warps per SM, not by HBM -- profiles/r02_layernorm_selfattn_full.txt);
    // per-element arithmetic is unchanged, the row sums are two interleaved partial sums
    uint64_t v2[PER_LANE / 2];
#pragma unroll
    for (int k = 0; k < PER_LANE / 2; k++) v2[k] = pk2(v[2 * k], v[2 * k + 1]);
    uint64_t s2 = pk2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < PER_LANE / 2; k++) s2 = add2(s2, v2[k]);
    const float mean = warp_sum(lo2(s2) + hi2(s2)) * (1.0f / D);
    const uint64_t nmean2 = pk2(-mean, -mean);
    uint64_t q2 = pk2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < PER_LANE / 2; k++) { v2[k] = add2(v2[k], nmean2); q2 = fma2(v2[k], v2[k], q2); }
    const float rstd = rsqrtf(warp_sum(lo2(q2) + hi2(q2)) * (1.0f / D) + eps);
    const uint64_t rstd2 = pk2(rstd, rstd);
#pragma unroll
    for (int k = 0; k < PER_LANE / 2; k++) {
      const uint64_t o2 = fma2(mul2(v2[k], rstd2), pk2(g[2 * k], g[2 * k + 1]), pk2(b[2 * k], b[2 * k + 1]));
      v[2 * k] = lo2(o2); v[2 * k + 1] = hi2(o2);
    }
    store_row<TY>(y + r * D, lane, v);
  }
}

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
bert_embeddings_kernel(const int32_t* __restrict__ ids, int64_t rows, int64_t L, const float* __restrict__ word,
                       const float* __restrict__ pos, const float* __restrict__ gamma, const float* __restrict__ beta,
                       T* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int64_t l = r % L;
  float v[PER_LANE], t[PER_LANE];
  load_row<float>(word + (int64_t)ids[r] * D, lane, v);
  load_row<float>(pos + l * D, lane, t);
#pragma unroll
  for (int i = 0; i < PER_LANE; i++) v[i] += t[i];
  layernorm24(v, gamma, beta, 1e-12f, lane);
  store_row<T>(out + r * D, lane, v);
}

// rows of `vecs` 16-byte vectors; grid-stride over (row, vec)
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ index, uint4* __restrict__ dst,
                                   int64_t rows, int64_t vecs) {
  const int64_t total = rows * vecs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vecs, c = i - r * vecs;
    const int64_t s = index ? (int64_t)index[r] : r;
    dst[i] = src[s * vecs + c];
  }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = from_f32<TD>(to_f32<TS>(src[i]));
}

__global__ void l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int64_t dim) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (r >= rows) return;
  float s = 0.f;
  for (int64_t i = lane; i < dim; i += 32) { float v = x[r * dim + i]; s = fmaf(v, v, s); }
  const float nrm = fmaxf(sqrtf(warp_sum(s)), 1e-12f);        // F.normalize: x / max(||x||, eps)
  for (int64_t i = lane; i < dim; i += 32) y[r * dim + i] = x[r * dim + i] / nrm;
}

// images fp32 [B,3,S,S] -> patches [B*P, 768], column = c*256 + ky*16 + kx (Conv2d weight flatten order)
template <typename T>
__global__ void im2col16_kernel(const float* __restrict__ img, T* __restrict__ out, int64_t B, int S) {
  const int G = S / 16;
  const int64_t total = B * G * G * 192;          // one thread per 4 contiguous pixels (kx..kx+3)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % 192);
    const int64_t pr = i / 192;                   // patch row index b*P + py*G + px
    const int px = (int)(pr % G), py = (int)((pr / G) % G);
    const int64_t b = pr / (G * G);
    const int c = v / 64, ky = (v % 64) / 4, kx = (v % 4) * 4;
    const float4 t = *reinterpret_cast<const float4*>(img + ((b * 3 + c) * S + (py * 16 + ky)) * (int64_t)S + px * 16 + kx);
    T* o = out + pr * 768 + c * 256 + ky * 16 + kx;
    o[0] = from_f32<T>(t.x); o[1] = from_f32<T>(t.y); o[2] = from_f32<T>(t.z); o[3] = from_f32<T>(t.w);
  }
}

// x[b,0,:] = cls + pos[0]; x[b,1+p,:] = patch[b*P+p,:] + pos[1+p]   (src/vit.py:182-188); x fp32
__global__ void vit_assemble_kernel(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos,
                                    float* __restrict__ x, int64_t B, int64_t N) {
  const int64_t total = B * N * (D / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (D / 4)) * 4;
    const int64_t r = i / (D / 4);
    const int64_t n = r % N, b = r / N;
    float4 a = (n == 0) ? *reinterpret_cast<const float4*>(cls + c)
                        : *reinterpret_cast<const float4*>(patch + (b * (N - 1) + n - 1) * D + c);
    const float4 p = *reinterpret_cast<const float4*>(pos + n * D + c);
    *reinterpret_cast<float4*>(x + r * D + c) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}

// feats[t] = cat(h0[t,0,:], h1[t,0,:]) (src/nlvr_encoder.py:909); h = [2][T*L][768]
template <typename T>
__global__ void gather_cls_kernel(const T* __restrict__ h, int64_t Tn, int64_t L, T* __restrict__ feats, float* __restrict__ feats_f32) {
  const int64_t total = Tn * 2 * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D);
    const int s = (int)((i / D) % 2);
    const int64_t t = i / (2 * D);
    const T v = h[((int64_t)s * Tn * L + t * L) * D + c];
    feats[i] = v;
    if (feats_f32) feats_f32[i] = to_f32<T>(v);
  }
}

// scores[t] = hidden[t,:] . w + b    (row 0 of cls_head.2, src/blip_stage2.py:53,136)
__global__ void __launch_bounds__(WARPS * 32)
head_dot_kernel(const float* __restrict__ hidden, const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ scores, int64_t rows) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (r >= rows) return;
  float hv[PER_LANE], wv[PER_LANE];
  load_row<float>(hidden + r * D, lane, hv);
  load_row<float>(w, lane, wv);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; i++) s = fmaf(hv[i], wv[i], s);
  s = warp_sum(s);
  if (lane == 0) scores[r] = s + b[0];
}

inline unsigned row_blocks(int64_t rows) { return (unsigned)((rows + WARPS - 1) / WARPS); }
inline unsigned flat_blocks(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  return (unsigned)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int cir_add_layernorm(cir_ctx* ctx, const void* x, int x_f32, int64_t x_rows, const void* res,
                                 const float* gamma, const float* beta, int64_t rows_per_group,
                                 void* y, int y_f32, int64_t rows, float eps) {
  CIR_ENTER(ctx);
  if (rows == 0) return CIR_OK;
  CIR_CHECK_ARG(x && gamma && beta && y && x_rows > 0 && rows_per_group > 0, "add_layernorm: null/zero argument");
  CIR_CHECK_ARG(rows < (1ll << 31), "add_layernorm: rows=%lld does not fit 32-bit row arithmetic", (long long)rows);
  const bool f32 = ctx->dtype == CIR_DTYPE_F32;
  dim3 grid(row_blocks((rows + LN_ROWS - 1) / LN_ROWS)), block(WARPS * 32);
  {
    const double es = f32 ? 4.0 : 2.0;
    cir_prof_begin(ctx, CIR_PROF_LAYERNORM, (double)rows * D * ((x_f32 ? 4.0 : es) + (res ? es : 0.0) + (y_f32 ? 4.0 : es)));
  }
#define LN_LAUNCH(TX, TR, TY) \
  add_layernorm_kernel<TX, TR, TY><<<grid, block, 0, ctx->stream>>>((const TX*)x, x_rows, (const TR*)res, gamma, beta, rows_per_group, (TY*)y, rows, eps)
  if (f32) LN_LAUNCH(float, float, float);
  else if (x_f32 && y_f32) LN_LAUNCH(float, bf16, float);
  else if (x_f32) LN_LAUNCH(float, bf16, bf16);
  else if (y_f32) LN_LAUNCH(bf16, bf16, float);
  else LN_LAUNCH(bf16, bf16, bf16);
#undef LN_LAUNCH
  cir_prof_end(ctx);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}

extern "C" int cir_bert_embeddings(cir_ctx* ctx, const int32_t* ids, int64_t Q, int64_t L, const float* word_emb,
                                   const float* pos_emb, const float* gamma, const float* beta, void* out) {
  CIR_ENTER(ctx);
  const int64_t rows = Q * L;
  if (rows == 0) return CIR_OK;
  CIR_CHECK_ARG(L <= 512, "bert_embeddings: L=%lld exceeds max_position_embeddings 512", (long long)L);
  if (ctx->dtype == CIR_DTYPE_F32)
    bert_embeddings_kernel<float><<<row_blocks(rows), WARPS * 32, 0, ctx->stream>>>(ids, rows, L, word_emb, pos_emb, gamma, beta, (float*)out);
  else
    bert_embeddings_kernel<bf16><<<row_blocks(rows), WARPS * 32, 0, ctx->stream>>>(ids, rows, L, word_emb, pos_emb, gamma, beta, (bf16*)out);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}

extern "C" int cir_gather_rows(cir_ctx* ctx, const void* src, const int32_t* index, void* dst, int64_t rows, int64_t row_elems) {
  CIR_ENTER(ctx);
  if (rows == 0 || row_elems == 0) return CIR_OK;
  const int64_t row_bytes = row_elems * (int64_t)act_size(ctx);
  CIR_CHECK_ARG(row_bytes % 16 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "gather_rows: rows must be 16 B multiples and aligned");
  const int64_t vecs = row_bytes / 16;
  gather_rows_kernel<<<flat_blocks(rows * vecs), 256, 0, ctx->stream>>>((const uint4*)src, index, (uint4*)dst, rows, vecs);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}

extern "C" int cir_cast_f32_to_act(cir_ctx* ctx, const float* src, void* dst, int64_t n) {
  CIR_ENTER(ctx);
  if (n == 0) return CIR_OK;
  if (ctx->dtype == CIR_DTYPE_F32) cast_kernel<float, float><<<flat_blocks(n), 256, 0, ctx->stream>>>(src, (float*)dst, n);
  else cast_kernel<float, bf16><<<flat_blocks(n), 256, 0, ctx->stream>>>(src, (bf16*)dst, n);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}

extern "C" int cir_cast_act_to_f32(cir_ctx* ctx, const void* src, float* dst, int64_t n) {
  CIR_ENTER(ctx);
  if (n == 0) return CIR_OK;
  if (ctx->dtype == CIR_DTYPE_F32) cast_kernel<float, float><<<flat_blocks(n), 256, 0, ctx->stream>>>((const float*)src, dst, n);
  else cast_kernel<bf16, float><<<flat_blocks(n), 256, 0, ctx->stream>>>((const bf16*)src, dst, n);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}

extern "C" int cir_l2_normalize(cir_ctx* ctx, const float* x, float* y, int64_t rows, int64_t dim) {
  CIR_ENTER(ctx);
  if (rows == 0) return CIR_OK;
  l2_normalize_kernel<<<row_blocks(rows), WARPS * 32, 0, ctx->stream>>>(x, y, rows, dim);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}

// ---- internal (pipeline.cu) -------------------------------------------------------------
int cir_im2col16(cir_ctx* ctx, const float* img, void* out, int64_t B, int S) {
  const int64_t n = B * (S / 16) * (S / 16) * 192;
  if (ctx->dtype == CIR_DTYPE_F32) im2col16_kernel<float><<<flat_blocks(n), 256, 0, ctx->stream>>>(img, (float*)out, B, S);
  else im2col16_kernel<bf16><<<flat_blocks(n), 256, 0, ctx->stream>>>(img, (bf16*)out, B, S);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}
int cir_vit_assemble(cir_ctx* ctx, const float* patch, const float* cls, const float* pos, float* x, int64_t B, int64_t N) {
  vit_assemble_kernel<<<flat_blocks(B * N * (D / 4)), 256, 0, ctx->stream>>>(patch, cls, pos, x, B, N);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}
int cir_gather_cls(cir_ctx* ctx, const void* h, int64_t T, int64_t L, void* feats, float* feats_f32) {
  if (T == 0) return CIR_OK;
  if (ctx->dtype == CIR_DTYPE_F32) gather_cls_kernel<float><<<flat_blocks(T * 2 * D), 256, 0, ctx->stream>>>((const float*)h, T, L, (float*)feats, feats_f32);
  else gather_cls_kernel<bf16><<<flat_blocks(T * 2 * D), 256, 0, ctx->stream>>>((const bf16*)h, T, L, (bf16*)feats, feats_f32);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}
int cir_head_dot(cir_ctx* ctx, const float* hidden, const float* w, const float* b, float* scores, int64_t rows) {
  if (rows == 0) return CIR_OK;
  head_dot_kernel<<<row_blocks(rows), WARPS * 32, 0, ctx->stream>>>(hidden, w, b, scores, rows);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}
