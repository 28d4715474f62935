// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[b] = act(A[b] W[b]^T + bias[b]) (+ residual[b])
//
//   A [rows, K] bf16 K-major, W [N, K] bf16 K-major (PyTorch Linear layout) -> both operands are
//   "K-major" UMMA operands, loaded by TMA into 128B-swizzled shared-memory tiles.
//   One persistent CTA per SM, warp-specialised:
//     warp 0      TMA producer            (one elected lane)
//     warp 1      TMEM allocator + tcgen05.mma issuer (one elected lane)
//     warps 2..9  epilogue: tcgen05.ld TMEM -> registers -> bias/activation/residual -> global
//   Pipelines: smem ring full/empty mbarriers (TMA <-> MMA), 2 TMEM accumulator stages with
//   tmem_full/tmem_empty mbarriers (MMA <-> epilogue), static persistent tile scheduler.
//   Tile 128 x BN x 64 (BN = 256 or 128), UMMA 128 x BN x 16, fp32 accumulation in TMEM.
//
// Batched operands are folded into the row coordinate of 2-D tensor maps: batch b of A starts at
// row b*a_rows_per_batch (0 = shared A), of W at row b*w_rows_per_batch.  Tiles that cross a
// batch edge load neighbouring (or zero-filled OOB) rows and are masked in the epilogue.
#include "common.cuh"

namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = 128 B = one SWIZZLE_128B atom row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int THREADS = 64 + NUM_EPI_WARPS * 32;   // 320
constexpr int ACC_STAGES = 2;

template <int BN> struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;            // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;            // 32 / 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = ACC_STAGES * BN;      // 512 / 256
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 192 /*barriers*/ + ACC_STAGES * BN * 4 /*bias*/;
};

struct Params {
  void* C; const float* bias; const void* residual;
  int64_t M, N, K;
  int64_t ldc, ldres;
  int64_t c_bstride, bias_bstride, res_bstride;
  int64_t a_rows_per_batch, w_rows_per_batch;
  int32_t batch, act, c_f32, res_f32;
  int32_t m_blocks, n_blocks, k_blocks, num_tiles;
};

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a broken pipeline traps (surfacing as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint64_t t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((it & 0x3ff) == 0x3ff) {
      uint64_t t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) { printf("cir: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once all previously issued tcgen05.mma of this thread completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (unused for swizzled K-major; 1)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row core-matrix groups
//   [46,48) version = 1 (Blackwell)   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=BF16 [7,10),
// b=BF16 [10,13), a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Abramowitz-Stegun 7.1.26 erf (|err| < 1.5e-7, far below bf16 resolution): ~12 instructions
// instead of erff's ~40, so the GELU epilogue keeps up with the MMA pipe.
__device__ __forceinline__ float gelu_fast(float x) {
  float z = fabsf(x) * 0.70710678118654752440f;
  float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  float erf_abs = 1.0f - poly * __expf(-z * z);
  float erf_v = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_v);
}

// ----------------------------------------------------------------------------- the kernel
template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const Params p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B needs 1024 B alignment
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + C::STAGES * C::A_BYTES;
  const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
  // barrier layout: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::STAGES + ACC_STAGES + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * C::STAGES + 2 * ACC_STAGES);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_gen + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 2 * ACC_STAGES));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < C::STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < ACC_STAGES; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), NUM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  const int tiles_per_batch = p.m_blocks * p.n_blocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_batch;
        const int r = tile - b * tiles_per_batch;
        const int m_blk = r / p.n_blocks, n_blk = r - m_blk * p.n_blocks;
        const int32_t a_row = (int32_t)(b * p.a_rows_per_batch + (int64_t)m_blk * BM);
        const int32_t w_row = (int32_t)(b * p.w_rows_per_batch + (int64_t)n_blk * BN);
        for (int kb = 0; kb < p.k_blocks; kb++) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
          tma_load_2d(smem_a + stage * C::A_BYTES, &map_a, full_bar(stage), kb * BK, a_row);
          tma_load_2d(smem_b + stage * C::B_BYTES, &map_w, full_bar(stage), kb * BK, w_row);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);          // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.k_blocks; kb++) {
          mbar_wait(full_bar(stage), phase);                // TMA bytes landed
          tcgen05_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_a + stage * C::A_BYTES);
          const uint64_t bdesc = make_smem_desc_sw128(smem_b + stage * C::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) {
            // advance 32 B (16 bf16) along K inside the 128 B swizzle atom: +2 in the >>4 address field
            umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(empty_bar(stage));                    // frees the smem slot when the MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));                        // accumulator ready for the epilogue
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    // Latency matters more than bandwidth here (8 warps, dependent chains), so: the tile's bias row is
    // staged in shared memory BEFORE the accumulator is waited for, bf16 residual rows are prefetched
    // into registers ahead of the TMEM load, and two 32-column chunks are in flight per tcgen05.wait.
    const int ew = warp - 2;                 // 0..7
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access (warp_id % 4)
    const int half = ew >> 2;                // which half of the BN columns
    constexpr int COLS_PER_WARP = BN / 2;
    constexpr int NCH = COLS_PER_WARP / 32;  // 4 or 2 (even)
    float* sbias = reinterpret_cast<float*>(smem_gen + C::STAGES * C::STAGE_BYTES + 192);   // [ACC_STAGES][BN]
    const int etid = threadIdx.x - 64;       // 0..255
    const bool res_bf16_fast = p.residual && !p.res_f32 && (p.ldres & 7) == 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_batch;
      const int r = tile - b * tiles_per_batch;
      const int m_blk = r / p.n_blocks, n_blk = r - m_blk * p.n_blocks;
      const int64_t row = (int64_t)m_blk * BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const int64_t ntile0 = (int64_t)n_blk * BN;
      if (etid < BN) {
        const int64_t n = ntile0 + etid;
        sbias[acc * BN + etid] = (p.bias && n < p.N) ? __ldg(p.bias + b * p.bias_bstride + n) : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");       // epilogue warps only
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();

      auto finish = [&](int c, uint32_t (&v)[32], const uint4 (&rq)[4], bool have_rq) {
        const int col0 = half * COLS_PER_WARP + c * 32;
        const int64_t n0 = ntile0 + col0;
        if (n0 >= p.N || !row_ok) return;              // N tail chunk / M tail row: nothing to store
        const bool full = (n0 + 32 <= p.N);
        float f[32];
        const float4* sb = reinterpret_cast<const float4*>(sbias + acc * BN + col0);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = sb[j >> 2];
          f[j] = __uint_as_float(v[j]) + bv.x; f[j + 1] = __uint_as_float(v[j + 1]) + bv.y;
          f[j + 2] = __uint_as_float(v[j + 2]) + bv.z; f[j + 3] = __uint_as_float(v[j + 3]) + bv.w;
        }
        if (p.act == CIR_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = gelu_fast(f[j]);
        } else if (p.act == CIR_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; j++) f[j] = fmaxf(f[j], 0.f);
        }
        if (p.residual) {
          const int64_t ro = b * p.res_bstride + row * p.ldres + n0;
          if (have_rq) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rq[j]);
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const float2 t = __bfloat1622float2(h2[q]);
                f[j * 8 + 2 * q] += t.x; f[j * 8 + 2 * q + 1] += t.y;
              }
            }
          } else if (p.res_f32) {
            const float* rp = (const float*)p.residual + ro;
            if (full && (p.ldres & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 rv = *reinterpret_cast<const float4*>(rp + j);
                f[j] += rv.x; f[j + 1] += rv.y; f[j + 2] += rv.z; f[j + 3] += rv.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++) if (n0 + j < p.N) f[j] += rp[j];
            }
          } else {
            const bf16* rp = (const bf16*)p.residual + ro;
#pragma unroll
            for (int j = 0; j < 32; j++) if (n0 + j < p.N) f[j] += __bfloat162float(rp[j]);
          }
        }
        const int64_t co = b * p.c_bstride + row * p.ldc + n0;
        if (p.c_f32) {
          float* cp = (float*)p.C + co;
          if (full && (p.ldc & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) if (n0 + j < p.N) cp[j] = f[j];
          }
        } else {
          bf16* cp = (bf16*)p.C + co;
          if (full && (p.ldc & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 ov;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&ov);
#pragma unroll
              for (int q = 0; q < 4; q++) h2[q] = __floats2bfloat162_rn(f[j + 2 * q], f[j + 2 * q + 1]);
              *reinterpret_cast<uint4*>(cp + j) = ov;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) if (n0 + j < p.N) cp[j] = __float2bfloat16_rn(f[j]);
          }
        }
      };

#pragma unroll 1
      for (int c = 0; c < NCH; c += 2) {
        uint4 rq0[4], rq1[4];
        bool have0 = false, have1 = false;
        if (res_bf16_fast && row_ok) {                 // residual rows do not depend on the accumulator: fetch first
          const int64_t n00 = ntile0 + half * COLS_PER_WARP + c * 32;
          const bf16* rp = (const bf16*)p.residual + b * p.res_bstride + row * p.ldres + n00;
          have0 = n00 + 32 <= p.N;
          have1 = n00 + 64 <= p.N;
          if (have0) {
#pragma unroll
            for (int j = 0; j < 4; j++) rq0[j] = *reinterpret_cast<const uint4*>(rp + j * 8);
          }
          if (have1) {
#pragma unroll
            for (int j = 0; j < 4; j++) rq1[j] = *reinterpret_cast<const uint4*>(rp + 32 + j * 8);
          }
        }
        uint32_t v0[32], v1[32];
        __syncwarp();                                  // tcgen05.ld is .sync.aligned: reconverge first
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * COLS_PER_WARP + c * 32);
        tmem_ld_32x32b_x32(taddr, v0);
        tmem_ld_32x32b_x32(taddr + 32, v1);
        tmem_ld_wait();
        finish(c, v0, rq0, have0);
        finish(c + 1, v1, rq1, have1);
      }
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator back
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

}  // namespace tc

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(cir_ctx* ctx, PFN_encodeTiled* fn) {
  if (!ctx->encode_tiled) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CIR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p) {
      cir_set_error("cuTensorMapEncodeTiled not available from the driver");
      return CIR_EUNSUPPORTED;
    }
    ctx->encode_tiled = p;
  }
  *fn = (PFN_encodeTiled)ctx->encode_tiled;
  return CIR_OK;
}

// 2-D bf16 tensor map over a [rows, K] K-major matrix with row stride ld (elements); box = [box_rows, 64]
static int make_map_2d(cir_ctx* ctx, CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  PFN_encodeTiled enc;
  CIR_TRY(get_encode_fn(ctx, &enc));
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    cir_set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%lld K=%lld ld=%lld box_rows=%d", (int)r, base,
                  (long long)rows, (long long)K, (long long)ld, box_rows);
    return CIR_ECUDA;
  }
  return CIR_OK;
}

template <int BN>
static int launch_tc(cir_ctx* ctx, const cir_gemm_args* a, const tc::Params& p, const CUtensorMap& ma, const CUtensorMap& mw) {
  using C = tc::Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CIR_CUDA(cudaFuncSetAttribute(tc::gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  int grid = p.num_tiles < ctx->num_sms ? p.num_tiles : ctx->num_sms;
  cir_prof_gemm_begin(ctx, 2.0 * (double)p.M * (double)p.N * (double)p.K * (double)p.batch);
  tc::gemm_tcgen05_kernel<BN><<<grid, tc::THREADS, C::SMEM_BYTES, ctx->stream>>>(ma, mw, p);
  cir_prof_gemm_end(ctx);
  CIR_LAUNCH_CHECK(ctx);
  (void)a;
  return CIR_OK;
}

int cir_gemm_tcgen05(cir_ctx* ctx, const cir_gemm_args* a) {
  if (a->M == 0 || a->N == 0 || a->batch == 0) return CIR_OK;
  CIR_CHECK_ARG(ctx->dtype == CIR_DTYPE_BF16, "tcgen05 GEMM needs a bf16 context");
  CIR_CHECK_ARG(a->K % 8 == 0 && a->lda % 8 == 0 && a->ldw % 8 == 0, "tcgen05 GEMM: K, lda, ldw must be multiples of 8 (TMA 16 B strides)");
  CIR_CHECK_ARG(((uintptr_t)a->A & 15) == 0 && ((uintptr_t)a->W & 15) == 0, "tcgen05 GEMM: A and W must be 16 B aligned");
  CIR_CHECK_ARG(a->batch == 1 || (a->a_bstride % a->lda == 0 && a->w_bstride % a->ldw == 0),
                "tcgen05 GEMM: batch strides must be whole rows");
  tc::Params p;
  p.C = a->C; p.bias = a->bias; p.residual = a->residual;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.ldc = a->ldc; p.ldres = a->ldres;
  p.c_bstride = a->c_bstride; p.bias_bstride = a->bias_bstride; p.res_bstride = a->res_bstride;
  p.batch = a->batch; p.act = a->act; p.c_f32 = a->c_f32; p.res_f32 = a->res_f32;
  p.a_rows_per_batch = a->batch > 1 ? a->a_bstride / a->lda : 0;
  p.w_rows_per_batch = a->batch > 1 ? a->w_bstride / a->ldw : 0;
  const int64_t a_rows = p.a_rows_per_batch * (a->batch - 1) + a->M;
  const int64_t w_rows = p.w_rows_per_batch * (a->batch - 1) + a->N;
  CIR_CHECK_ARG(a_rows < (1ll << 31) && w_rows < (1ll << 31), "tcgen05 GEMM: too many rows for a 32-bit TMA coordinate");
  // small problems: narrower N tile so the persistent grid still covers the SMs
  const int64_t tiles256 = ((a->M + tc::BM - 1) / tc::BM) * ((a->N + 255) / 256) * a->batch;
  const bool use128 = (a->N <= 128) || (tiles256 < ctx->num_sms);
  const int BN = use128 ? 128 : 256;
  p.m_blocks = (int32_t)((a->M + tc::BM - 1) / tc::BM);
  p.n_blocks = (int32_t)((a->N + BN - 1) / BN);
  p.k_blocks = (int32_t)((a->K + tc::BK - 1) / tc::BK);
  const int64_t nt = (int64_t)p.m_blocks * p.n_blocks * a->batch;
  CIR_CHECK_ARG(nt < (1ll << 31), "tcgen05 GEMM: too many tiles");
  p.num_tiles = (int32_t)nt;
  CUtensorMap ma, mw;
  CIR_TRY(make_map_2d(ctx, &ma, a->A, a_rows, a->K, a->lda, tc::BM));
  CIR_TRY(make_map_2d(ctx, &mw, a->W, w_rows, a->K, a->ldw, BN));
  if (use128) return launch_tc<128>(ctx, a, p, ma, mw);
  return launch_tc<256>(ctx, a, p, ma, mw);
}
