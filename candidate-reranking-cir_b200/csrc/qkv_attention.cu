// Fused QKV projection + masked text self-attention for sm_100a (bf16):
//
//   ctx[b][rows, h*64 .. h*64+63] = softmax(Q_h K_h^T / 8 + mask) V_h,   [Q|K|V] = A[b] W[b]^T + bias[b]
//
// replaces the `query/key/value` Linears plus the score / softmax / context products of the twin text self-attention
// (src/nlvr_encoder.py:140-222 with encoder_hidden_states=None; src/med.py:112-216 in stage I) for captions of L = 16 or
// 32 tokens.  The unfused path writes the [rows, 2304] projection to HBM and reads it back in a second kernel that is
// purely bandwidth bound (8 % of a stage-II step for 0.5 % of its FLOPs); here the projection never leaves the SM.
//
// One tcgen05 CTA-pair tile = 256 rows (8 captions of 32 tokens) x 192 columns = [Q_h | K_h | V_h] of ONE head: the
// W tile is gathered by TMA as six 32-row slabs (rows h*64.. of the Q, K and V blocks of the stacked [2304, 768] weight;
// no re-packing of the checkpoint layout), UMMA 256 x 192 x 16 (cta_group::2), fp32 accumulators in TMEM (2 stages).  The
// epilogue warps round the biased tile to bf16 into a padded shared-memory tile (exactly the values the unfused path
// would have written to HBM), hand the accumulator back to the MMA warp, and finish the attention of their 16 query rows
// with mma.sync (16 + 16 HMMA per warp), so only the [rows, 64] context leaves the SM -- through coalesced 128-byte row
// segments.  The arithmetic (operation order included) is that of attention_small_kernel, so fused == unfused bit for bit.
#include "common.cuh"
#include "tcgen05_ptx.cuh"
#include <stdlib.h>

namespace qkvattn {
using namespace tc;

constexpr int BM = 128;                 // rows per CTA (two CTAs per tile)
constexpr int BK = 64;
constexpr int BN = 192;                 // Q_h | K_h | V_h
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int THREADS = 64 + NUM_EPI_WARPS * 32;
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;           // TMEM column stride between the accumulator stages
constexpr int TMEM_COLS = 512;
constexpr int B_ROWS = BN / 2;          // W rows staged per CTA: three 32-row slabs
constexpr int SLAB_ROWS = 32;
constexpr int SLAB_BYTES = SLAB_ROWS * BK * 2;     // 4 KB
constexpr int A_BYTES = BM * BK * 2;    // 16 KB
constexpr int B_BYTES = B_ROWS * BK * 2;           // 12 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGES = 6;
constexpr int PITCH = 400;              // bytes per staged row (192 bf16 = 384 B + 16 B pad: ldmatrix / row-per-lane stores conflict-free)
constexpr int STAGING_BYTES = (BM + 8) * PITCH;   // + 8 rows: the 16-row fragment loads of an 8-token caption's unit run past row 127
constexpr int BAR_BYTES = 192;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + STAGING_BYTES + BAR_BYTES + ACC_STAGES * BN * 4;
constexpr int DH = CIR_HEAD_DIM;        // 64
constexpr int HEADS = CIR_HEADS;        // 12
constexpr int DM = CIR_HIDDEN;          // 768

struct Params {
  void* out; const float* bias;
  const int32_t* key_mask; const int32_t* mask_index;
  int64_t M;                            // rows per batch (captions * L)
  int64_t out_rs, out_bs, bias_bs;
  int64_t a_rows_per_batch, w_rows_per_batch;
  int32_t batch, L, m_blocks, k_blocks, num_tiles;
  int32_t drain;                        // the MMA thread lets the UMMA queue run empty every `drain` k-blocks (0 = never); see the kernel comment
  float scale;
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// tile -> (batch, 256-row block, head); heads innermost so the twelve tiles that share an A block run side by side
__device__ __forceinline__ bool next_tile(const Params& p, int worker, int num_workers, int it, int& b, int& m_blk, int& h) {
  const int tile = worker + it * num_workers;
  if (tile >= p.num_tiles) return false;
  const int unit = tile / HEADS;
  h = tile - unit * HEADS;
  b = unit / p.m_blocks;
  m_blk = unit - b * p.m_blocks;
  return true;
}

// NT = key n-tiles of 8 (L / 8): 1..4.  A CTA's 128 accumulator rows hold CAPS = 128 / L whole captions (ROWS = CAPS * L rows: 128
// for L = 8 / 16 / 32, 120 for L = 24 -- the tile then advances by 240 rows and the last 8 rows of each CTA are dead); the attention is
// done in units of (caption, 16-query-row block), UNITS = CAPS * ceil(L / 16) of them spread over the 8 epilogue warps.
template <int NT>
__global__ void __launch_bounds__(THREADS, 1)
qkv_attention_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const Params p) {
  constexpr int L = NT * 8;
  constexpr int CAPS = BM / L, ROWS = CAPS * L, TILE_ROWS = 2 * ROWS;
  constexpr int MT = (L + 15) / 16, UNITS = CAPS * MT, KS = (NT + 1) / 2;
  const uint32_t rank = cluster_ctarank();
  const int worker = (int)(blockIdx.x >> 1);
  const int num_workers = (int)(gridDim.x >> 1);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * A_BYTES;
  const uint32_t stg = smem_base + STAGES * STAGE_BYTES;                 // [128 rows][PITCH]: Q | K | V of this CTA's rows, bf16
  const uint32_t bar_base = stg + STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
  const uint32_t drain_bar = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES + 1);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_gen + STAGES * STAGE_BYTES + STAGING_BYTES + 8 * (2 * STAGES + 2 * ACC_STAGES));
  float* sbias = reinterpret_cast<float*>(smem_gen + STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES);        // [ACC_STAGES][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < ACC_STAGES; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), NUM_EPI_WARPS * 2); }
    mbar_init(drain_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc_pair<TMEM_COLS>(tmem_ptr_smem);
  tcgen05_fence_before();
  __syncwarp();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      int b, m_blk, h;
      for (int it = 0; next_tile(p, worker, num_workers, it, b, m_blk, h); ++it) {
        const int32_t a_row = (int32_t)(b * p.a_rows_per_batch + (int64_t)m_blk * TILE_ROWS + rank * ROWS);
        int32_t w_row[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const int slab = 3 * (int)rank + j;            // slabs 0,1 = Q_h; 2,3 = K_h; 4,5 = V_h (32 rows each)
          w_row[j] = (int32_t)(b * p.w_rows_per_batch + (int64_t)(slab >> 1) * DM + h * DH + (slab & 1) * SLAB_ROWS);
        }
        for (int kb = 0; kb < p.k_blocks; kb++) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
          tma_load_2d_pair(smem_a + stage * A_BYTES, &map_a, full_bar(stage), kb * BK, a_row);
#pragma unroll
          for (int j = 0; j < 3; j++)
            tma_load_2d_pair(smem_b + stage * B_BYTES + j * SLAB_BYTES, &map_w, full_bar(stage), kb * BK, w_row[j]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      uint32_t drain_phase = 0;
      int b_, m_, h_;
      for (int it = 0; next_tile(p, worker, num_workers, it, b_, m_, h_); ++it) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_COLS);
        for (int kb = 0; kb < p.k_blocks; kb++) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_a + stage * A_BYTES);
          const uint64_t bdesc = make_smem_desc_sw128(smem_b + stage * B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++)
            umma_bf16_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)((kb | k) != 0));
          umma_commit_pair(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          // The epilogue warps' mma.sync HMMAs share the tensor pipe with this UMMA stream and are only dispatched when the
          // UMMA queue is empty (ncu: 48 % of their samples were stall_math on the first HMMAs of a burst).  With the ring
          // full the queue never empties inside a tile, so the attention of tile i would wait for the end of tile i+2's
          // UMMAs: let the queue run dry once in the middle of every tile (-7 % kernel time).
          if (p.drain && (kb + 1) % p.drain == 0 && kb + 1 < p.k_blocks) {
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(drain_bar) : "memory");
            mbar_wait(drain_bar, drain_phase);
            drain_phase ^= 1;
          }
        }
        umma_commit_pair(tfull_bar(acc));
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: projection tile -> shared memory -> attention -> context =====================
    const int ew = warp - 2;                 // 0..7: attention unit = query rows [16 ew, 16 ew + 16) of this CTA's 128 rows
    const int quarter = warp & 3;            // TMEM lane quarter this warp may read
    const int half = ew >> 2;                // which 96 of the 192 accumulator columns this warp stages
    const int etid = threadIdx.x - 64;
    const int g = lane >> 2, t = lane & 3;
    const float sl2 = p.scale * 1.4426950408889634f;
    // the 8 pad rows behind the staged tile are only ever READ (fragment loads that run past row 127 and meet probability 0): keep them finite
    if (etid < 8 * PITCH / 16) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(stg + (uint32_t)(BM * PITCH) + (uint32_t)etid * 16), "r"(0u) : "memory");
    int acc = 0; uint32_t acc_phase = 0;
    int b, m_blk, h;
    for (int it = 0; next_tile(p, worker, num_workers, it, b, m_blk, h); ++it) {
      if (etid < BN) {
        const int part = etid >> 6, c = etid & 63;           // Q / K / V block of the stacked bias
        sbias[acc * BN + etid] = p.bias ? __ldg(p.bias + b * p.bias_bs + part * DM + h * DH + c) : 0.f;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      // bias row visible; every warp has finished reading the previous tile's staged rows
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // ---- accumulator (+ bias) -> bf16 -> staged row `quarter*32 + lane`, columns half*96 .. +95
      {
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + half * 96);
        const uint32_t rowaddr = stg + (uint32_t)(quarter * 32 + lane) * PITCH + (uint32_t)(half * 96) * 2;
        const float* sb = sbias + acc * BN + half * 96;
        uint32_t v[3][32];
        __syncwarp();
        tmem_ld_32x32b_x32(taddr, v[0]);
        tmem_ld_32x32b_x32(taddr + 32, v[1]);
        tmem_ld_32x32b_x32(taddr + 64, v[2]);
        tmem_ld_wait();
        // all TMEM reads of this warp are complete: hand the accumulator back before the attention math
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
#pragma unroll
        for (int c = 0; c < 3; c++) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const float4 b0 = *reinterpret_cast<const float4*>(sb + c * 32 + j);
            const float4 b1 = *reinterpret_cast<const float4*>(sb + c * 32 + j + 4);
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v[c][j]) + b0.x, __uint_as_float(v[c][j + 1]) + b0.y);
            o.y = pack_bf16(__uint_as_float(v[c][j + 2]) + b0.z, __uint_as_float(v[c][j + 3]) + b0.w);
            o.z = pack_bf16(__uint_as_float(v[c][j + 4]) + b1.x, __uint_as_float(v[c][j + 5]) + b1.y);
            o.w = pack_bf16(__uint_as_float(v[c][j + 6]) + b1.z, __uint_as_float(v[c][j + 7]) + b1.w);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)(c * 32 + j) * 2), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
        }
      }
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      asm volatile("bar.sync 2, 256;" ::: "memory");             // the 128 x 192 tile is staged
      // ---- attention, one unit = 16 query rows of one caption against the caption's L keys
      const int64_t tile_row0 = (int64_t)m_blk * TILE_ROWS + rank * ROWS;            // first row of this CTA's captions within the batch
      auto unit = [&](const int u) {
      const int cl = u / MT, mt = u - cl * MT;                   // caption within the CTA, 16-row block within the caption
      const int krow0 = cl * L, qrow0 = krow0 + mt * 16;         // staged rows of the caption's keys / of this unit's queries
      const int64_t row_lo = tile_row0 + qrow0;                  // row within the batch
      if (tile_row0 + krow0 >= p.M) return;                      // tail tile: captions beyond the batch (warp-uniform)
      const int nvalid = (L % 16 == 0) ? 16 : (L - mt * 16 < 16 ? L - mt * 16 : 16);    // 16, or 8 in the last block of a 24- / 8-token caption
      const int64_t cap = (tile_row0 + krow0) / L;               // caption (= triplet) index within the batch
      const int32_t* mask = p.key_mask ? p.key_mask + (int64_t)(p.mask_index ? __ldg(p.mask_index + cap) : cap) * L : nullptr;
      float madd[NT][2];
#pragma unroll
      for (int j = 0; j < NT; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) madd[j][e] = (mask && __ldg(mask + j * 8 + 2 * t + e) == 0) ? -10000.0f * 1.4426950408889634f : 0.f;
      // Q fragments (A operand, 16 rows x 64): matrix i of an x4 load = rows (i&1)*8.., columns (i>>1)*8..  (rows past the caption
      // belong to the next one or are dead: their results are never stored)
      uint32_t qa[4][4];
      {
        const uint32_t qaddr = stg + (uint32_t)(qrow0 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (uint32_t)((lane >> 4) * 8) * 2;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) ldmatrix_x4(qaddr + (uint32_t)(kk * 16) * 2, qa[kk]);
      }
      // S = Q K^T: n-tile j = keys 8j..8j+7; one x4 load = B fragments of two k-steps
      float sc[NT][4];
#pragma unroll
      for (int j = 0; j < NT; j++) {
        sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
        const uint32_t kaddr = stg + (uint32_t)(krow0 + j * 8 + (lane & 7)) * PITCH + (uint32_t)(DH + (lane >> 3) * 8) * 2;
        uint32_t kb0[4], kb1[4];
        ldmatrix_x4(kaddr, kb0);                                 // head dims 0..31
        ldmatrix_x4(kaddr + 64, kb1);                            // head dims 32..63
        mma_bf16_16816(sc[j], qa[0], kb0[0], kb0[1]);
        mma_bf16_16816(sc[j], qa[1], kb0[2], kb0[3]);
        mma_bf16_16816(sc[j], qa[2], kb1[0], kb1[1]);
        mma_bf16_16816(sc[j], qa[3], kb1[2], kb1[3]);
      }
      // scores / 8 + mask (-10000 on padded keys, nlvr_encoder.py:193,774), exact softmax in the log2 domain
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < NT; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          sc[j][e] = fmaf(sc[j][e], sl2, madd[j][e]);
          sc[j][2 + e] = fmaf(sc[j][2 + e], sl2, madd[j][e]);
          m0 = fmaxf(m0, sc[j][e]);
          m1 = fmaxf(m1, sc[j][2 + e]);
        }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      float l0 = 0.f, l1 = 0.f;
      uint32_t pa[KS][4];
#pragma unroll
      for (int j = 0; j < KS; j++) { pa[j][0] = pa[j][1] = pa[j][2] = pa[j][3] = 0u; }      // keys past L (odd NT) carry probability 0
#pragma unroll
      for (int j = 0; j < NT; j++) {
        const float p00 = exp2f(sc[j][0] - m0), p01 = exp2f(sc[j][1] - m0);
        const float p10 = exp2f(sc[j][2] - m1), p11 = exp2f(sc[j][3] - m1);
        l0 += p00 + p01; l1 += p10 + p11;
        pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p00, p01);
        pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p10, p11);
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      // O = P V: V rows are keys (k), columns head dims (n) -> transposed ldmatrix (rows past L meet probability 0)
      float o[8][4];
#pragma unroll
      for (int j = 0; j < 8; j++) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < KS; kk++) {
        const int vrow = krow0 + kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const uint32_t vaddr = stg + (uint32_t)vrow * PITCH + (uint32_t)(2 * DH + (lane >> 4) * 8) * 2;
#pragma unroll
        for (int jp = 0; jp < 4; jp++) {
          uint32_t vb[4];
          ldmatrix_x4_trans(vaddr + (uint32_t)(jp * 16) * 2, vb);
          mma_bf16_16816(o[2 * jp], pa[kk], vb[0], vb[1]);
          mma_bf16_16816(o[2 * jp + 1], pa[kk], vb[2], vb[3]);
        }
      }
      const float i0 = 1.0f / l0, i1 = 1.0f / l1;
      // context rows over this unit's own (valid) Q rows -- nobody else reads them --, then out as whole 128-byte row segments
      __syncwarp();
      {
        const uint32_t o0 = stg + (uint32_t)(qrow0 + g) * PITCH + (uint32_t)(2 * t) * 2;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(o0 + (uint32_t)(j * 8) * 2), "r"(pack_bf16(o[j][0] * i0, o[j][1] * i0)) : "memory");
          if (nvalid == 16)
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(o0 + 8u * PITCH + (uint32_t)(j * 8) * 2), "r"(pack_bf16(o[j][2] * i1, o[j][3] * i1)) : "memory");
        }
      }
      __syncwarp();
      {
        bf16* op = (bf16*)p.out + b * p.out_bs + h * DH + (lane & 7) * 8;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int r = i * 4 + (lane >> 3);
          if (r < nvalid) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(stg + (uint32_t)(qrow0 + r) * PITCH + (uint32_t)(lane & 7) * 16) : "memory");
            *reinterpret_cast<uint4*>(op + (row_lo + r) * p.out_rs) = v;
          }
        }
      }
      __syncwarp();
      };
      if constexpr (UNITS == NUM_EPI_WARPS) {
        unit(ew);                                                // L = 16 / 32: exactly one unit per warp
      } else {
#pragma unroll 1
        for (int u = ew; u < UNITS; u += NUM_EPI_WARPS) unit(u);
      }
    }
  }

  tcgen05_fence_before();
  __syncwarp();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_pair<TMEM_COLS>(tmem_base);
  }
}

template <int NT>
static int launch(cir_ctx* ctx, const Params& p, const CUtensorMap& ma, const CUtensorMap& mw, double work) {
  const unsigned bit = NT == 4 ? 1u << 24 : (NT == 2 ? 1u << 25 : (NT == 3 ? 1u << 30 : 1u << 31));
  if (!(ctx->func_attr_mask & bit)) {
    CIR_CUDA(cudaFuncSetAttribute(qkv_attention_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ctx->func_attr_mask |= bit;
  }
  const int slots = ctx->num_sms / 2;
  const int workers = p.num_tiles < slots ? p.num_tiles : slots;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * workers);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cir_prof_begin(ctx, CIR_PROF_QKV_ATTN, work);
  cudaError_t e = cudaLaunchKernelEx(&cfg, qkv_attention_kernel<NT>, ma, mw, p);
  cir_prof_end(ctx);
  if (e != cudaSuccess) { cir_set_error("fused QKV + self-attention launch failed: %s", cudaGetErrorString(e)); return CIR_ECUDA; }
  CIR_LAUNCH_CHECK(ctx);
  return CIR_OK;
}

}  // namespace qkvattn

bool cir_qkv_attention_supported(const cir_ctx* ctx, int64_t L) {
  return ctx->dtype == CIR_DTYPE_BF16 && ctx->gemm_impl != CIR_GEMM_SIMT && ctx->attn_impl == 0 && ctx->gemm_pair && (L == 8 || L == 16 || L == 24 || L == 32);
}

extern "C" int cir_qkv_attention(cir_ctx* ctx, const cir_qkv_attn_args* a) {
  CIR_ENTER(ctx);
  CIR_CHECK_ARG(a && a->x && a->w && a->out, "qkv_attention: null operand");
  if (a->captions == 0 || a->batch == 0) return CIR_OK;
  CIR_CHECK_ARG(cir_qkv_attention_supported(ctx, a->L), "qkv_attention: needs a bf16 tcgen05 context and L in {8, 16, 24, 32} (got L=%d)", a->L);
  CIR_CHECK_ARG(a->batch >= 1 && a->captions > 0, "qkv_attention: bad shape");
  CIR_CHECK_ARG(((uintptr_t)a->x & 15) == 0 && ((uintptr_t)a->w & 15) == 0 && ((uintptr_t)a->out & 15) == 0 && (a->out_rs % 8) == 0 &&
                (a->out_bs % 8) == 0, "qkv_attention: operands must be 16 B aligned");
  using namespace qkvattn;
  Params p{};
  p.out = a->out; p.bias = a->bias; p.key_mask = a->key_mask; p.mask_index = a->mask_index;
  p.M = a->captions * (int64_t)a->L;
  p.out_rs = a->out_rs; p.out_bs = a->out_bs; p.bias_bs = 3 * DM;
  p.batch = a->batch; p.L = a->L; p.scale = a->scale;
  CIR_CHECK_ARG(a->batch == 1 || a->x_bs % DM == 0, "qkv_attention: batch stride must be whole rows");
  p.a_rows_per_batch = a->batch > 1 ? a->x_bs / DM : 0;
  p.w_rows_per_batch = 3 * DM;
  const int64_t a_rows = p.a_rows_per_batch * (a->batch - 1) + p.M;
  const int64_t w_rows = (int64_t)3 * DM * a->batch;
  CIR_CHECK_ARG(a_rows < (1ll << 31), "qkv_attention: too many rows for a 32-bit TMA coordinate");
  const int64_t tile_rows = 2 * (BM / a->L) * a->L;           // 256, or 240 for L = 24
  p.m_blocks = (int32_t)((p.M + tile_rows - 1) / tile_rows);
  p.k_blocks = DM / BK;
  const int64_t nt = (int64_t)p.m_blocks * HEADS * a->batch;
  CIR_CHECK_ARG(nt < (1ll << 31), "qkv_attention: too many tiles");
  p.num_tiles = (int32_t)nt;
  // UMMA-queue drain period in k-blocks (see the MMA warp): mid-tile by default; CIR_QKV_DRAIN overrides it for A/B measurements
  // (sustained: 0 = never 0.960 ms, 6 = mid-tile 0.885, 4 -> 0.901; every k-block 0.998 cold)
  { const char* d = getenv("CIR_QKV_DRAIN"); p.drain = d ? atoi(d) : p.k_blocks / 2; }
  CUtensorMap ma, mw;
  CIR_TRY(cir_make_map_2d(ctx, &ma, a->x, a_rows, DM, DM, BM));
  CIR_TRY(cir_make_map_2d(ctx, &mw, a->w, w_rows, DM, DM, SLAB_ROWS));
  const double work = 2.0 * (double)p.M * 3.0 * DM * DM * a->batch + 4.0 * (double)a->captions * a->batch * HEADS * (double)a->L * a->L * DH;
  switch (a->L) {
    case 32: return launch<4>(ctx, p, ma, mw, work);
    case 24: return launch<3>(ctx, p, ma, mw, work);
    case 16: return launch<2>(ctx, p, ma, mw, work);
    default: return launch<1>(ctx, p, ma, mw, work);
  }
}
