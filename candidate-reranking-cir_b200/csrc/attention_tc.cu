// tcgen05 / TMEM flash attention (bf16, head dim 64, no key mask): the cross-attention of the
// dual-stream encoder (32 text rows x 577 image keys per (triplet, stream, head),
// src/nlvr_encoder.py:175-217) and the ViT self-attention (src/vit.py:74-82).
//
// One CTA = one 128-row query tile of one head.  A tile is (128/RB) batches x RB rows that all attend
// the SAME K/V batch (candidate-major triplets: 4 triplets x 32 rows share one candidate image), so the
// 128-row UMMA shape is filled even though a single triplet has only 32 query rows.
//
//   warps 0-3 (128 threads)  softmax: thread r owns query row r == TMEM lane r.  Per 128-key chunk:
//                            tcgen05.ld S -> row max -> exp2 -> bf16 P written to 128B-swizzled shared
//                            memory (the A operand of the PV product) -> O_j merged into registers.
//   warp 4                   one elected thread: TMA loads of K/V chunks (2-deep rings) and all
//                            tcgen05.mma issue:  S_j = Q K_j^T  (A=Q smem, B=K_j smem, both K-major)
//                                                O_j = P_j V_j  (A=P smem K-major, B=V_j smem MN-major)
//   TMEM: S 128 columns + O 64 columns (256 allocated) -> two CTAs per SM overlap each other's
//   MMA / softmax phases.  Scores and probabilities never touch HBM.
#include "common.cuh"
#include "tcgen05_ptx.cuh"

namespace fatc {

using namespace tc;

constexpr int KC = 128;                 // keys per chunk
constexpr int THREADS = 160;
constexpr int Q_BYTES = 128 * 128;      // 128 rows x 64 bf16
constexpr int KV_BYTES = KC * 128;      // 128 keys x 64 bf16
constexpr int P_BYTES = 2 * 128 * 128;  // two 64-key atoms of [128 rows x 128 B]
constexpr int OFF_Q = 0, OFF_K = Q_BYTES, OFF_V = OFF_K + 2 * KV_BYTES, OFF_P = OFF_V + 2 * KV_BYTES;
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, O_COL = 128;

// instruction descriptor with an MN-major B operand (V is [key][dh], dh contiguous): bit 16
__host__ __device__ constexpr uint32_t make_idesc_bf16_bmn(int m, int n) { return make_idesc_bf16(m, n) | (1u << 16); }

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {       // one MUFU.EX2; exp2(-inf) = 0, denormal results flush to 0
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__global__ void __launch_bounds__(THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, const cir_attn_args p,
                    int tiles_per_batch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y;
  // ---- tile descriptor
  int batch0, nb, row0, RB;
  if (p.tiles) {
    const int4 t = reinterpret_cast<const int4*>(p.tiles)[blockIdx.x];
    batch0 = t.x; nb = t.y; row0 = t.z; RB = t.w;
  } else {
    batch0 = blockIdx.x / tiles_per_batch; nb = 1; row0 = (blockIdx.x % tiles_per_batch) * 128; RB = 128;
  }
  const int kvb = p.kv_index ? p.kv_index[batch0] : batch0;
  const int nch = (p.Lk + KC - 1) / KC;

  // barriers: kfull[2], vfull[2], vfree[2], sfull, sfree, pfull, ofull, ofree, qfull, tmem_ptr
  const uint32_t bar = sbase + OFF_BAR;
  auto kfull = [&](int b) { return bar + 8u * b; };
  auto vfull = [&](int b) { return bar + 16u + 8u * b; };
  auto vfree = [&](int b) { return bar + 32u + 8u * b; };
  const uint32_t sfull = bar + 48, sfree = bar + 56, pfull = bar + 64, ofull = bar + 72, ofree = bar + 80, qfull = bar + 88;
  const uint32_t tmem_ptr_smem = bar + 96;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 96);

  if ((sbase & 1023u) != 0) { if (tid == 0) printf("cir: attention_tc smem misaligned\n"); __trap(); }
  if (warp == 4) {
    if (elect_one()) {
      tma_prefetch_desc(&map_k);
      tma_prefetch_desc(&map_v);
      for (int b = 0; b < 2; b++) { mbar_init(kfull(b), 1); mbar_init(vfull(b), 1); mbar_init(vfree(b), 1); }
      mbar_init(sfull, 1); mbar_init(sfree, 4); mbar_init(pfull, 4); mbar_init(ofull, 1); mbar_init(ofree, 4); mbar_init(qfull, 4);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 4) {
    // ===================== control: TMA + MMA issue =====================
    if (elect_one()) {
      const int32_t krow0 = kvb * p.Lk;
      const int32_t col = h * 64;
      const int pre = nch < 2 ? nch : 2;
      for (int j = 0; j < pre; j++) {
        mbar_expect_tx(kfull(j), KV_BYTES);
        tma_load_2d(sbase + OFF_K + j * KV_BYTES, &map_k, kfull(j), col, krow0 + j * KC);
        mbar_expect_tx(vfull(j), KV_BYTES);
        tma_load_2d(sbase + OFF_V + j * KV_BYTES, &map_v, vfull(j), col, krow0 + j * KC);
      }
      mbar_wait(qfull, 0);                                  // Q tile written (generic proxy) + proxy fence by the softmax warps
      tcgen05_fence_after();
      const uint64_t qdesc = make_smem_desc_sw128(sbase + OFF_Q);
      for (int j = 0; j < nch; j++) {
        const int b = j & 1;
        const uint32_t par2 = (uint32_t)((j >> 1) & 1);
        const int keys = (p.Lk - j * KC) < KC ? (p.Lk - j * KC) : KC;
        const int n_pad = (keys + 15) & ~15;                // UMMA N (multiple of 16); padded keys are masked in the softmax
        // ---- S_j = Q K_j^T
        mbar_wait(kfull(b), par2);
        if (j > 0) mbar_wait(sfree, (uint32_t)((j - 1) & 1));
        tcgen05_fence_after();
        {
          const uint64_t kdesc = make_smem_desc_sw128(sbase + OFF_K + b * KV_BYTES);
          const uint32_t idesc = make_idesc_bf16(128, n_pad);
#pragma unroll
          for (int k = 0; k < 4; k++) umma_bf16(tmem_base + S_COL, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc, (uint32_t)(k != 0));
          umma_commit(sfull);
        }
        // ---- V ring: chunk j+1 reuses the buffer of chunk j-1 once PV_{j-1} retired
        if (j >= 1 && j + 1 < nch) {
          const int bb = (j + 1) & 1;
          mbar_wait(vfree(bb), (uint32_t)(((j - 1) >> 1) & 1));
          mbar_expect_tx(vfull(bb), KV_BYTES);
          tma_load_2d(sbase + OFF_V + bb * KV_BYTES, &map_v, vfull(bb), col, krow0 + (j + 1) * KC);
        }
        // ---- O_j = P_j V_j
        mbar_wait(pfull, (uint32_t)(j & 1));                 // P_j in shared memory; S_j fully consumed (so K_j is free too)
        if (j + 2 < nch) {                                   // K ring: two chunks ahead
          mbar_expect_tx(kfull(b), KV_BYTES);
          tma_load_2d(sbase + OFF_K + b * KV_BYTES, &map_k, kfull(b), col, krow0 + (j + 2) * KC);
        }
        mbar_wait(vfull(b), par2);
        if (j > 0) mbar_wait(ofree, (uint32_t)((j - 1) & 1));
        tcgen05_fence_after();
        {
          const uint32_t idesc = make_idesc_bf16_bmn(128, 64);
          const int ksteps = n_pad >> 4;
          for (int k = 0; k < ksteps; k++) {
            // A: P atom (k/4) of [128 rows x 64 keys], +32 B per 16 keys inside the atom.  B: V rows 16k.. (16 x 128 B)
            const uint64_t pdesc = make_smem_desc_sw128(sbase + OFF_P + (k >> 2) * (128 * 128)) + (uint64_t)((k & 3) * 2);
            const uint64_t vdesc = make_smem_desc_sw128(sbase + OFF_V + b * KV_BYTES + k * 2048);
            umma_bf16(tmem_base + O_COL, pdesc, vdesc, idesc, (uint32_t)(k != 0));
          }
          umma_commit(ofull);
          umma_commit(vfree(b));
        }
      }
    }
  } else {
    // ===================== softmax warps: thread = query row = TMEM lane =====================
    const int r = tid;                                       // 0..127
    const int bi = r / RB, qi = row0 + r % RB;
    const bool valid = bi < nb && qi < p.Lq;
    const int b = batch0 + (bi < nb ? bi : 0);
    // ---- Q row -> swizzled shared memory (K-major SWIZZLE_128B: 16 B piece c of row r lives at c ^ (r & 7))
    {
      const uint4* qsrc = reinterpret_cast<const uint4*>((const bf16*)p.q + (int64_t)b * p.q_bs + (int64_t)qi * p.q_rs + h * 64);
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const uint4 v = valid ? qsrc[c] : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(smem + OFF_Q + r * 128 + ((c ^ (r & 7)) << 4)) = v;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(qfull);
    }
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; i++) o[i] = 0.f;
    float m = -INFINITY, l = 0.f, a_pending = 0.f;
    for (int j = 0; j < nch; j++) {
      const int keys = (p.Lk - j * KC) < KC ? (p.Lk - j * KC) : KC;
      const int n_pad = (keys + 15) & ~15;
      const int npieces = (n_pad + 31) >> 5;
      mbar_wait(sfull, (uint32_t)(j & 1));
      tcgen05_fence_after();
      // ---- pass 1: row max over the valid keys of this chunk (3-input FMNMX3; full chunks need no key test)
      float cmax = -INFINITY;
      const bool full_chunk = keys == KC;
      for (int pc = 0; pc < npieces; pc++) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32b_x32(lane_addr + S_COL + pc * 32, v);
        tmem_ld_wait();
        if (full_chunk) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) cmax = max3(cmax, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) if (pc * 32 + i < keys) cmax = fmaxf(cmax, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m, cmax * sl2);              // scale > 0: max commutes with the scaling
      const float alpha = ex2_approx(m - m_new);             // first chunk: exp2(-inf) = 0
      // ---- merge O_{j-1} (computed against the previous max) while the tensor core is idle anyway
      if (j > 0) {
        mbar_wait(ofull, (uint32_t)((j - 1) & 1));
        tcgen05_fence_after();
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld_32x32b_x32(lane_addr + O_COL + hh * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) o[hh * 32 + i] = fmaf(o[hh * 32 + i], a_pending, __uint_as_float(v[i]));
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ofree);
      }
      a_pending = alpha;
      l *= alpha;
      m = m_new;
      // ---- pass 2: P = exp2(S*c - m) as bf16 into the swizzled A-operand tile; padded keys -> 0
      for (int pc = 0; pc < npieces; pc++) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32b_x32(lane_addr + S_COL + pc * 32, v);
        tmem_ld_wait();
        float pr[32];
        if (full_chunk) {
#pragma unroll
          for (int i = 0; i < 32; i++) pr[i] = ex2_approx(fmaf(__uint_as_float(v[i]), sl2, -m));
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const float e = ex2_approx(fmaf(__uint_as_float(v[i]), sl2, -m));
            pr[i] = (pc * 32 + i < keys) ? e : 0.f;
          }
        }
        float ls = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i++) ls += pr[i];
        l += ls;
        // keys pc*32 .. pc*32+31 -> atom (pc >> 1), 16 B pieces ((pc & 1) * 4 + 0..3) of row r
        uint8_t* prow = smem + OFF_P + (pc >> 1) * (128 * 128) + r * 128;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          uint4 w;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
          for (int q = 0; q < 4; q++) h2[q] = __floats2bfloat162_rn(pr[c * 8 + 2 * q], pr[c * 8 + 2 * q + 1]);
          const int piece = (pc & 1) * 4 + c;
          *reinterpret_cast<uint4*>(prow + ((piece ^ (r & 7)) << 4)) = w;
        }
      }
      fence_proxy_async();                                   // generic-proxy P writes -> visible to the tensor core (async proxy)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(pfull); mbar_arrive(sfree); }
    }
    // ---- last O_j, normalise, store
    mbar_wait(ofull, (uint32_t)((nch - 1) & 1));
    tcgen05_fence_after();
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      uint32_t v[32];
      __syncwarp();
      tmem_ld_32x32b_x32(lane_addr + O_COL + hh * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i++) o[hh * 32 + i] = fmaf(o[hh * 32 + i], a_pending, __uint_as_float(v[i]));
    }
    if (valid) {
      const float inv = 1.0f / l;
      uint4* dst = reinterpret_cast<uint4*>((bf16*)p.o + (int64_t)b * p.o_bs + (int64_t)qi * p.o_rs + h * 64);
#pragma unroll
      for (int c = 0; c < 8; c++) {
        uint4 w;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
        for (int q = 0; q < 4; q++) h2[q] = __floats2bfloat162_rn(o[c * 8 + 2 * q] * inv, o[c * 8 + 2 * q + 1] * inv);
        dst[c] = w;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace fatc

int cir_attention_tc(cir_ctx* ctx, const cir_attn_args* a) {
  // eligibility (the caller falls back to the mma.sync kernel on CIR_EUNSUPPORTED)
  if (ctx->dtype != CIR_DTYPE_BF16 || a->key_mask != nullptr || a->Lk < 16 || a->kv_batches <= 0 || a->H * 64 > 65536 ||
      a->k_bs != (int64_t)a->Lk * a->k_rs || a->v_bs != (int64_t)a->Lk * a->v_rs || (a->k_rs % 8) || (a->v_rs % 8) ||
      ((uintptr_t)a->k & 15) || ((uintptr_t)a->v & 15) || ((uintptr_t)a->q & 15) || ((uintptr_t)a->o & 15) ||
      (a->q_rs % 8) || (a->q_bs % 8) || (a->o_rs % 8) || (a->o_bs % 8))
    return CIR_EUNSUPPORTED;
  if (!a->tiles && a->Lq <= 64 && a->B > 1) return CIR_EUNSUPPORTED;      // unshared short queries would waste 1/2..3/4 of every tile
  const int64_t kv_rows = (int64_t)a->kv_batches * a->Lk;
  if (kv_rows >= (1ll << 31)) return CIR_EUNSUPPORTED;
  CUtensorMap mk, mv;
  CIR_TRY(cir_make_map_2d(ctx, &mk, a->k, kv_rows, (int64_t)a->H * 64, a->k_rs, fatc::KC));
  CIR_TRY(cir_make_map_2d(ctx, &mv, a->v, kv_rows, (int64_t)a->H * 64, a->v_rs, fatc::KC));
  static bool attr_set = false;
  if (!attr_set) {
    CIR_CUDA(cudaFuncSetAttribute(fatc::attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fatc::SMEM_BYTES));
    attr_set = true;
  }
  const int tpb = (a->Lq + 127) / 128;
  const unsigned gx = a->tiles ? (unsigned)a->num_tiles : (unsigned)(a->B * tpb);
  if (gx == 0) return CIR_OK;
  fatc::attention_tc_kernel<<<dim3(gx, (unsigned)a->H), fatc::THREADS, fatc::SMEM_BYTES, ctx->stream>>>(mk, mv, *a, tpb);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}
