// tcgen05 / TMEM flash attention (bf16, head dim 64, no key mask): the cross-attention of the
// dual-stream encoder (32 text rows x 577 image keys per (triplet, stream, head),
// src/nlvr_encoder.py:175-217) and the ViT self-attention (src/vit.py:74-82).
//
// One CTA = one 128-row query tile of one head; two CTAs are resident per SM.  A tile is (128/RB) batches
// x RB rows that all attend the SAME K/V batch (candidate-major triplets: 4 triplets x 32 rows share one
// candidate image), so the 128-row UMMA shape is filled although a triplet has only 32 query rows.
//
//   warps 0-3   softmax: thread r owns query row r == TMEM lane r.  Per 64-key chunk ONE tcgen05.ld sweep
//               brings the row's scores into registers, row max -> exp2 -> bf16 P written to a 128B-swizzled
//               shared-memory atom (the A operand of the PV product).  S and P are double-buffered, so the
//               tensor core computes S_{j+1}, S_{j+2} while the softmax works on chunk j.  O accumulates in
//               TMEM across chunks; the reference max moves lazily (only when the row max grew by more than
//               2^8), so the O rescale (tcgen05.ld / tcgen05.st) is a rare path.
//   warp 4      one elected thread: TMA loads of K/V chunks (3-deep rings) and all tcgen05.mma issue:
//                 S_j = Q K_j^T  (A=Q smem, B=K_j smem, both K-major)
//                 O  += P_j V_j  (A=P smem K-major, B=V_j smem MN-major)
//   TMEM: S0, S1 (64 columns each) + O (64).  Scores and probabilities never touch HBM.
#include <type_traits>
#include "common.cuh"
#include "tcgen05_ptx.cuh"

namespace fatc {

using namespace tc;

constexpr int KC = 64;                  // keys per chunk (one 128-byte swizzle atom of P per chunk)
constexpr int KS = 3;                   // K / V ring depth
constexpr int THREADS = 160;            // 4 softmax warps + 1 control warp
constexpr int Q_BYTES = 128 * 128;      // 128 rows x 64 bf16
constexpr int KV_BYTES = KC * 128;      // 64 keys x 64 bf16
constexpr int P_BYTES = 128 * 128;      // [128 rows x 64 keys] bf16
constexpr int OFF_Q = 0, OFF_K = Q_BYTES, OFF_V = OFF_K + KS * KV_BYTES, OFF_P = OFF_V + KS * KV_BYTES;
constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int TMEM_COLS = 256;
constexpr int S_COL = 0, O_COL = 128;   // S buffers at 0 and 64, O at 128

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

// O[row, 0:64] *= alpha in TMEM (whole warp, each lane its own row/alpha); rare path, kept out of line
__device__ __noinline__ void rescale_o(uint32_t o_addr, float alpha) {
#pragma unroll
  for (int hh = 0; hh < 2; hh++) {
    uint32_t ov[32];
    __syncwarp();
    tmem_ld_32x32b_x32(o_addr + hh * 32, ov);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i++) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
    tmem_st_32x32b_x32(o_addr + hh * 32, ov);
  }
  tmem_st_wait();
}

__global__ void __launch_bounds__(THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, const cir_attn_args p,
                    int ctas_per_batch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y;
  // ---- CTA descriptor: 128 query rows r -> batch batch0 + r / RB, query row row0 + r % RB
  int batch0, nb, row0, RB;
  if (p.tiles) {
    const int4 t = reinterpret_cast<const int4*>(p.tiles)[blockIdx.x];
    batch0 = t.x; nb = t.y; row0 = t.z; RB = t.w;
  } else {
    batch0 = blockIdx.x / ctas_per_batch; nb = 1; row0 = (blockIdx.x % ctas_per_batch) * 128; RB = 128;
  }
  const int kvb = p.kv_index ? p.kv_index[batch0] : batch0;
  const int nch = (p.Lk + KC - 1) / KC;

  // barriers (all indexed by chunk parity so a waiter is never more than one phase behind):
  //   kfull[KS] vfull[KS] | sfull[2] pfull[2] pvdone[2] | qfull | tmem_ptr
  const uint32_t bar = sbase + OFF_BAR;
  auto kfull = [&](int s) { return bar + 8u * s; };
  auto vfull = [&](int s) { return bar + 8u * (KS + s); };
  auto sfull = [&](int b) { return bar + 8u * (2 * KS + b); };
  auto pfull = [&](int b) { return bar + 8u * (2 * KS + 2 + b); };
  auto pvdone = [&](int b) { return bar + 8u * (2 * KS + 4 + b); };
  const uint32_t qfull = bar + 8u * (2 * KS + 6);
  const uint32_t tmem_ptr_smem = qfull + 8;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 8 * (2 * KS + 6) + 8);

  if ((sbase & 1023u) != 0) { if (tid == 0) printf("cir: attention_tc smem misaligned\n"); __trap(); }
  if (warp == 4) {
    if (elect_one()) {
      tma_prefetch_desc(&map_k);
      tma_prefetch_desc(&map_v);
      for (int s = 0; s < KS; s++) { mbar_init(kfull(s), 1); mbar_init(vfull(s), 1); }
      for (int b = 0; b < 2; b++) { mbar_init(sfull(b), 1); mbar_init(pfull(b), 4); mbar_init(pvdone(b), 1); }
      mbar_init(qfull, 4);
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
      auto load_k = [&](int j) {
        const int s = j % KS;
        mbar_expect_tx(kfull(s), KV_BYTES);
        tma_load_2d(sbase + OFF_K + s * KV_BYTES, &map_k, kfull(s), col, krow0 + j * KC);
      };
      auto load_v = [&](int j) {
        const int s = j % KS;
        mbar_expect_tx(vfull(s), KV_BYTES);
        tma_load_2d(sbase + OFF_V + s * KV_BYTES, &map_v, vfull(s), col, krow0 + j * KC);
      };
      auto n_pad_of = [&](int j) {
        const int keys = (p.Lk - j * KC) < KC ? (p.Lk - j * KC) : KC;
        return (keys + 15) & ~15;                            // UMMA N (multiple of 16); padded keys are masked in the softmax
      };
      const uint64_t qdesc = make_smem_desc_sw128(sbase + OFF_Q);
      auto issue_s = [&](int j) {                           // S[j&1] = Q K_j^T
        mbar_wait(kfull(j % KS), (uint32_t)((j / KS) & 1));
        tcgen05_fence_after();
        const uint64_t kdesc = make_smem_desc_sw128(sbase + OFF_K + (j % KS) * KV_BYTES);
        const uint32_t idesc = make_idesc_bf16(128, n_pad_of(j));
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_bf16(tmem_base + S_COL + (j & 1) * 64, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc, (uint32_t)(k != 0));
        umma_commit(sfull(j & 1));
      };
      const int pre = nch < KS ? nch : KS;
      for (int j = 0; j < pre; j++) { load_k(j); load_v(j); }
      mbar_wait(qfull, 0);                                  // Q tile written (generic proxy) + proxy fence by the softmax warps
      issue_s(0);
      if (nch > 1) issue_s(1);
      for (int j = 0; j < nch; j++) {
        mbar_wait(pfull(j & 1), (uint32_t)((j >> 1) & 1));   // P_j in shared memory, S_j consumed, O rescaled if needed
        mbar_wait(vfull(j % KS), (uint32_t)((j / KS) & 1));
        tcgen05_fence_after();
        {                                                    // O (+)= P_j V_j, accumulating in TMEM across chunks
          const uint32_t idesc = make_idesc_bf16_bmn(128, 64);
          const int ksteps = n_pad_of(j) >> 4;
          for (int k = 0; k < ksteps; k++) {
            const uint64_t pdesc = make_smem_desc_sw128(sbase + OFF_P + (j & 1) * P_BYTES) + (uint64_t)(k * 2);    // +32 B per 16 keys
            const uint64_t vdesc = make_smem_desc_sw128(sbase + OFF_V + (j % KS) * KV_BYTES + k * 2048);         // 16 key rows x 128 B
            umma_bf16(tmem_base + O_COL, pdesc, vdesc, idesc, (uint32_t)((j | k) != 0));
          }
          umma_commit(pvdone(j & 1));
        }
        if (j + 2 < nch) issue_s(j + 2);                     // S buffer j&1 was drained before pfull(j)
        if (j + KS < nch) load_k(j + KS);                    // S_j retired long ago -> K_j's stage is free
        if (j >= 1 && j - 1 + KS < nch) {                    // V_{j-1}'s stage once PV_{j-1} retired
          mbar_wait(pvdone((j - 1) & 1), (uint32_t)(((j - 1) >> 1) & 1));
          load_v(j - 1 + KS);
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
    const uint32_t o_addr = lane_addr + O_COL;
    const float sl2 = p.scale * 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;                            // m: lazily updated reference max (scaled log2 domain)
    for (int j = 0; j < nch; j++) {
      const int pb = j & 1;
      const int keys = (p.Lk - j * KC) < KC ? (p.Lk - j * KC) : KC;
      const bool full_chunk = keys == KC;
      const bool two = keys > 32;                            // second 32-column piece needed?
      mbar_wait(sfull(pb), (uint32_t)((j >> 1) & 1));
      tcgen05_fence_after();
      // ---- the row's 64 scores -> registers in one sweep
      uint32_t v0[32], v1[32];
      __syncwarp();
      tmem_ld_32x32b_x32(lane_addr + S_COL + pb * 64, v0);
      if (two) tmem_ld_32x32b_x32(lane_addr + S_COL + pb * 64 + 32, v1);
      tmem_ld_wait();
      // ---- row max of the chunk
      float cmax = -INFINITY;
      if (full_chunk) {                                      // four independent FMNMX3 chains (8 deep instead of 32)
        float c0 = -INFINITY, c1 = -INFINITY, c2 = -INFINITY, c3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          c0 = max3(c0, __uint_as_float(v0[i]), __uint_as_float(v0[i + 1]));
          c1 = max3(c1, __uint_as_float(v0[16 + i]), __uint_as_float(v0[17 + i]));
          c2 = max3(c2, __uint_as_float(v1[i]), __uint_as_float(v1[i + 1]));
          c3 = max3(c3, __uint_as_float(v1[16 + i]), __uint_as_float(v1[17 + i]));
        }
        cmax = fmaxf(max3(c0, c1, c2), c3);
      } else {
#pragma unroll
        for (int i = 0; i < 32; i++) if (i < keys) cmax = fmaxf(cmax, __uint_as_float(v0[i]));
        if (two) {
#pragma unroll
          for (int i = 0; i < 32; i++) if (32 + i < keys) cmax = fmaxf(cmax, __uint_as_float(v1[i]));
        }
      }
      cmax *= sl2;                                           // scale > 0: max commutes with the scaling
      // ---- lazy reference max: move it only when the row max grew by more than 2^8; the (rare) move rescales
      //      l and the O accumulator in TMEM.  Otherwise exp2(s - m) <= 256: harmless in fp32 / bf16.
      const bool grow = cmax > m + 8.0f;
      if (j == 0) {
        m = cmax;
      } else if (__any_sync(0xffffffffu, grow)) {
        mbar_wait(pvdone((j - 1) & 1), (uint32_t)(((j - 1) >> 1) & 1));      // every PV issued so far has retired
        tcgen05_fence_after();
        const float m_new = grow ? cmax : m;
        const float alpha = ex2_approx(m - m_new);           // 1 for rows that keep their reference
        l *= alpha;
        m = m_new;
        rescale_o(o_addr, alpha);
      }
      if (j >= 2) mbar_wait(pvdone(pb), (uint32_t)(((j - 2) >> 1) & 1));     // P buffer pb: PV_{j-2} has read it
      // ---- P = exp2(S*c - m) as bf16 into the swizzled A-operand tile (one 128 B row per thread); padded keys -> 0.
      //      Two instantiations: full chunks carry no per-element masking; the tail chunk only touches the
      //      padded-to-16 keys the PV product reads.
      uint8_t* prow = smem + OFF_P + pb * P_BYTES + r * 128;
      auto exp_store = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int npad = (keys + 15) & ~15;
#pragma unroll
        for (int c = 0; c < 8; c++) {
          if (!FULL && c * 8 >= npad) continue;
          float e[8];
#pragma unroll
          for (int q = 0; q < 8; q++) {
            const int i = c * 8 + q;
            const float sv = __uint_as_float(i < 32 ? v0[i & 31] : v1[i & 31]);
            e[q] = ex2_approx(fmaf(sv, sl2, -m));
            if (!FULL && i >= keys) e[q] = 0.f;
          }
          l += ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
          uint4 w;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
          for (int q = 0; q < 4; q++) h2[q] = __floats2bfloat162_rn(e[2 * q], e[2 * q + 1]);
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) << 4)) = w;
        }
      };
      if (full_chunk) exp_store(std::true_type{}); else exp_store(std::false_type{});
      fence_proxy_async();                                   // generic-proxy P writes -> visible to the tensor core (async proxy)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull(pb));
    }
    // ---- O, normalise, store
    mbar_wait(pvdone((nch - 1) & 1), (uint32_t)(((nch - 1) >> 1) & 1));
    tcgen05_fence_after();
    float o[64];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      uint32_t ov[32];
      __syncwarp();
      tmem_ld_32x32b_x32(o_addr + hh * 32, ov);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i++) o[hh * 32 + i] = __uint_as_float(ov[i]);
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
  if (!a->tiles && a->Lq <= 64 && a->B > 1) return CIR_EUNSUPPORTED;      // unshared short queries would waste most of every tile
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
  const int cpb = (a->Lq + 127) / 128;
  const unsigned gx = a->tiles ? (unsigned)a->num_tiles : (unsigned)(a->B * cpb);
  if (gx == 0) return CIR_OK;
  fatc::attention_tc_kernel<<<dim3(gx, (unsigned)a->H), fatc::THREADS, fatc::SMEM_BYTES, ctx->stream>>>(mk, mv, *a, cpb);
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}
