// Weight packing: fp32 tensors of the reference state_dicts (device pointers, reference key layout) -> the packed structs the
// pipelines take (include/cir_b200.h: cir_vit_weights, cir_stage1_weights, cir_stage2_weights).  GEMM weights are cast to the
// context's activation dtype and keep PyTorch's [out, in] layout (= K-major operands for A W^T); biases, LayerNorm parameters and
// embedding tables stay fp32; twin-stream / q;k;v tensors are stacked; the cross-attention output projections and the stream
// merge of the dual-stream encoder are composed into one [768, 1536] matrix per layer in fp64 (src/nlvr_encoder.py:250-258).
#include "common.cuh"

namespace {

constexpr int64_t D = CIR_HIDDEN, F = CIR_FFN, EM = CIR_EMBED;

__global__ void cast_to_act_kernel(const float* __restrict__ src, void* __restrict__ dst, int64_t n, int bf) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (bf) ((bf16*)dst)[i] = __float2bfloat16_rn(src[i]);
    else ((float*)dst)[i] = src[i];
  }
}

// out[n, col0 + k] = (act) (float) sum_j Wm[n, moff + j] * Wd[j, k]   (fp64 accumulation, j ascending); out row stride 1536
__global__ void fold_merge_kernel(const float* __restrict__ Wm, int moff, const float* __restrict__ Wd, void* __restrict__ out, int col0, int bf) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;      // 0..767
  const int n = blockIdx.y;                                 // 0..767
  if (k >= D) return;
  double acc = 0.0;
  for (int j = 0; j < D; j++) acc += (double)Wm[(int64_t)n * 2 * D + moff + j] * (double)Wd[(int64_t)j * D + k];
  const float v = (float)acc;
  if (bf) ((bf16*)out)[(int64_t)n * 2 * D + col0 + k] = __float2bfloat16_rn(v);
  else ((float*)out)[(int64_t)n * 2 * D + col0 + k] = v;
}
// bc[n] = (float)( sum_j Wm[n, j] b0[j] + sum_j Wm[n, 768 + j] b1[j] + bm[n] )
__global__ void fold_merge_bias_kernel(const float* __restrict__ Wm, const float* __restrict__ b0, const float* __restrict__ b1,
                                       const float* __restrict__ bm, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= D) return;
  double a = 0.0, b = 0.0;
  for (int j = 0; j < D; j++) a += (double)Wm[(int64_t)n * 2 * D + j] * (double)b0[j];
  for (int j = 0; j < D; j++) b += (double)Wm[(int64_t)n * 2 * D + D + j] * (double)b1[j];
  out[n] = (float)(a + b + (double)bm[n]);
}
// average merge: out[n, s*768 + k] = 0.5 * W_s[n, k];  bias 0.5 (b0 + b1)   (exact in fp32 and after the cast)
__global__ void half_kernel(const float* __restrict__ src, void* __restrict__ dst, int64_t rows, int64_t cols, int64_t ld, int64_t col0, int bf) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const float v = (float)(0.5 * (double)src[i]);
    if (bf) ((bf16*)dst)[r * ld + col0 + c] = __float2bfloat16_rn(v);
    else ((float*)dst)[r * ld + col0 + c] = v;
  }
}
__global__ void half_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)(0.5 * ((double)a[i] + (double)b[i]));
}

struct Packer {
  cir_ctx* ctx; char* base; size_t off, cap; bool dry, ok; int bf; size_t es; int rc;
  Packer(cir_ctx* c, void* blob, size_t bytes) : ctx(c), base((char*)blob), off(0), cap(bytes), dry(blob == nullptr), ok(true), rc(CIR_OK) {
    bf = c->dtype == CIR_DTYPE_BF16; es = bf ? 2 : 4;
  }
  void* take(size_t bytes) {
    off = align_up(off, 256);
    void* r = dry ? nullptr : base + off;
    off += bytes;
    if (!dry && off > cap) ok = false;
    return r;
  }
  unsigned blocks(int64_t n) const { int64_t b = (n + 255) / 256; return (unsigned)(b > 4096 ? 4096 : (b < 1 ? 1 : b)); }
  void fail(int r) { if (rc == CIR_OK) rc = r; }
  // GEMM weight made of `parts` fp32 tensors of `elems` elements each, stacked
  const void* W(const float* const* parts, int nparts, int64_t elems, const char* what) {
    char* dst = (char*)take((size_t)nparts * elems * es);
    if (dry || !ok) return dst;
    for (int i = 0; i < nparts; i++) {
      if (!parts[i]) { if (rc == CIR_OK) cir_set_error("pack: missing tensor %s (part %d)", what, i); fail(CIR_EINVAL); return dst; }
      cast_to_act_kernel<<<blocks(elems), 256, 0, ctx->stream>>>(parts[i], dst + (size_t)i * elems * es, elems, bf);
      ctx->launches++;
    }
    return dst;
  }
  const void* W1(const float* p, int64_t elems, const char* what) { const float* a[1] = {p}; return W(a, 1, elems, what); }
  // fp32 parameter made of stacked parts
  const float* P(const float* const* parts, int nparts, int64_t elems, const char* what) {
    float* dst = (float*)take((size_t)nparts * elems * 4);
    if (dry || !ok) return dst;
    for (int i = 0; i < nparts; i++) {
      if (!parts[i]) { if (rc == CIR_OK) cir_set_error("pack: missing tensor %s (part %d)", what, i); fail(CIR_EINVAL); return dst; }
      if (cudaMemcpyAsync(dst + (size_t)i * elems, parts[i], (size_t)elems * 4, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) {
        if (rc == CIR_OK) cir_set_error("pack: copy of %s failed: %s", what, cudaGetErrorString(cudaGetLastError()));
        fail(CIR_ECUDA);
      }
    }
    return dst;
  }
  const float* P1(const float* p, int64_t elems, const char* what) { const float* a[1] = {p}; return P(a, 1, elems, what); }
  int finish(const char* who) {
    if (dry) return CIR_OK;
    if (!ok) { cir_set_error("%s: blob of %zu bytes is too small (need %zu)", who, cap, align_up(off, 256)); return CIR_EWORKSPACE; }
    if (rc != CIR_OK) return rc;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cir_set_error("%s: kernel launch failed: %s", who, cudaGetErrorString(e)); return CIR_ECUDA; }
    return CIR_OK;
  }
};

void pack_embeddings(Packer& p, const cir_text_embed_state& e, const float*& word, const float*& pos, const float*& g, const float*& b) {
  word = p.P1(e.word_emb, e.vocab_rows * D, "embeddings.word_embeddings.weight");
  pos = p.P1(e.pos_emb, e.pos_rows * D, "embeddings.position_embeddings.weight");
  g = p.P1(e.ln_g, D, "embeddings.LayerNorm.weight");
  b = p.P1(e.ln_b, D, "embeddings.LayerNorm.bias");
}

int pack_vit(cir_ctx* ctx, const cir_vit_state* sd, void* blob, size_t bytes, cir_vit_weights* w, size_t* need) {
  Packer p(ctx, blob, bytes);
  cir_vit_weights o{};
  o.patch_w = p.W1(sd->patch_w, D * 3 * 16 * 16, "patch_embed.proj.weight");
  o.patch_b = p.P1(sd->patch_b, D, "patch_embed.proj.bias");
  o.cls_token = p.P1(sd->cls_token, D, "cls_token");
  o.pos_embed = p.P1(sd->pos_embed, sd->num_tokens * D, "pos_embed");
  for (int i = 0; i < CIR_LAYERS; i++) {
    o.norm1_g[i] = p.P1(sd->norm1_g[i], D, "norm1.weight"); o.norm1_b[i] = p.P1(sd->norm1_b[i], D, "norm1.bias");
    o.qkv_w[i] = p.W1(sd->qkv_w[i], 3 * D * D, "attn.qkv.weight"); o.qkv_b[i] = p.P1(sd->qkv_b[i], 3 * D, "attn.qkv.bias");
    o.proj_w[i] = p.W1(sd->proj_w[i], D * D, "attn.proj.weight"); o.proj_b[i] = p.P1(sd->proj_b[i], D, "attn.proj.bias");
    o.norm2_g[i] = p.P1(sd->norm2_g[i], D, "norm2.weight"); o.norm2_b[i] = p.P1(sd->norm2_b[i], D, "norm2.bias");
    o.fc1_w[i] = p.W1(sd->fc1_w[i], F * D, "mlp.fc1.weight"); o.fc1_b[i] = p.P1(sd->fc1_b[i], F, "mlp.fc1.bias");
    o.fc2_w[i] = p.W1(sd->fc2_w[i], D * F, "mlp.fc2.weight"); o.fc2_b[i] = p.P1(sd->fc2_b[i], D, "mlp.fc2.bias");
  }
  o.norm_g = p.P1(sd->norm_g, D, "norm.weight"); o.norm_b = p.P1(sd->norm_b, D, "norm.bias");
  if (need) *need = align_up(p.off, 256);
  if (w && !p.dry) *w = o;
  return p.finish("pack_vit");
}

int pack_stage1(cir_ctx* ctx, const cir_stage1_state* sd, void* blob, size_t bytes, cir_stage1_weights* w, size_t* need) {
  Packer p(ctx, blob, bytes);
  cir_stage1_weights o{};
  pack_embeddings(p, sd->emb, o.word_emb, o.pos_emb, o.emb_ln_g, o.emb_ln_b);
  for (int i = 0; i < CIR_LAYERS; i++) {
    const float* qkv_w[3] = {sd->self_q_w[i], sd->self_k_w[i], sd->self_v_w[i]};
    const float* qkv_b[3] = {sd->self_q_b[i], sd->self_k_b[i], sd->self_v_b[i]};
    o.self_qkv_w[i] = p.W(qkv_w, 3, D * D, "attention.self.{query,key,value}.weight");
    o.self_qkv_b[i] = p.P(qkv_b, 3, D, "attention.self.{query,key,value}.bias");
    o.self_out_w[i] = p.W1(sd->self_out_w[i], D * D, "attention.output.dense.weight"); o.self_out_b[i] = p.P1(sd->self_out_b[i], D, "attention.output.dense.bias");
    o.self_ln_g[i] = p.P1(sd->self_ln_g[i], D, "attention.output.LayerNorm.weight"); o.self_ln_b[i] = p.P1(sd->self_ln_b[i], D, "attention.output.LayerNorm.bias");
    o.cross_q_w[i] = p.W1(sd->cross_q_w[i], D * D, "crossattention.self.query.weight"); o.cross_q_b[i] = p.P1(sd->cross_q_b[i], D, "crossattention.self.query.bias");
    const float* kv_w[2] = {sd->cross_k_w[i], sd->cross_v_w[i]};
    const float* kv_b[2] = {sd->cross_k_b[i], sd->cross_v_b[i]};
    o.cross_kv_w[i] = p.W(kv_w, 2, D * D, "crossattention.self.{key,value}.weight");
    o.cross_kv_b[i] = p.P(kv_b, 2, D, "crossattention.self.{key,value}.bias");
    o.cross_out_w[i] = p.W1(sd->cross_out_w[i], D * D, "crossattention.output.dense.weight"); o.cross_out_b[i] = p.P1(sd->cross_out_b[i], D, "crossattention.output.dense.bias");
    o.cross_ln_g[i] = p.P1(sd->cross_ln_g[i], D, "crossattention.output.LayerNorm.weight"); o.cross_ln_b[i] = p.P1(sd->cross_ln_b[i], D, "crossattention.output.LayerNorm.bias");
    o.ffn1_w[i] = p.W1(sd->ffn1_w[i], F * D, "intermediate.dense.weight"); o.ffn1_b[i] = p.P1(sd->ffn1_b[i], F, "intermediate.dense.bias");
    o.ffn2_w[i] = p.W1(sd->ffn2_w[i], D * F, "output.dense.weight"); o.ffn2_b[i] = p.P1(sd->ffn2_b[i], D, "output.dense.bias");
    o.ffn_ln_g[i] = p.P1(sd->ffn_ln_g[i], D, "output.LayerNorm.weight"); o.ffn_ln_b[i] = p.P1(sd->ffn_ln_b[i], D, "output.LayerNorm.bias");
  }
  o.text_proj_w = p.W1(sd->text_proj_w, EM * D, "text_proj.weight"); o.text_proj_b = p.P1(sd->text_proj_b, EM, "text_proj.bias");
  o.vision_proj_w = p.W1(sd->vision_proj_w, EM * D, "vision_proj.weight"); o.vision_proj_b = p.P1(sd->vision_proj_b, EM, "vision_proj.bias");
  if (need) *need = align_up(p.off, 256);
  if (w && !p.dry) *w = o;
  return p.finish("pack_stage1");
}

int pack_stage2(cir_ctx* ctx, const cir_stage2_state* sd, void* blob, size_t bytes, cir_stage2_weights* w, size_t* need) {
  Packer p(ctx, blob, bytes);
  cir_stage2_weights o{};
  pack_embeddings(p, sd->emb, o.word_emb, o.pos_emb, o.emb_ln_g, o.emb_ln_b);
  for (int i = 0; i < CIR_LAYERS; i++) {
    // twin self-attention: stream-major q;k;v (self0 then self1)
    const float* qkv_w[6] = {sd->self_q_w[0][i], sd->self_k_w[0][i], sd->self_v_w[0][i], sd->self_q_w[1][i], sd->self_k_w[1][i], sd->self_v_w[1][i]};
    const float* qkv_b[6] = {sd->self_q_b[0][i], sd->self_k_b[0][i], sd->self_v_b[0][i], sd->self_q_b[1][i], sd->self_k_b[1][i], sd->self_v_b[1][i]};
    o.self_qkv_w[i] = p.W(qkv_w, 6, D * D, "attention.self{0,1}.{query,key,value}.weight");
    o.self_qkv_b[i] = p.P(qkv_b, 6, D, "attention.self{0,1}.{query,key,value}.bias");
    const float* so_w[2] = {sd->self_out_w[0][i], sd->self_out_w[1][i]};
    const float* so_b[2] = {sd->self_out_b[0][i], sd->self_out_b[1][i]};
    o.self_out_w[i] = p.W(so_w, 2, D * D, "attention.output.dense{0,1}.weight"); o.self_out_b[i] = p.P(so_b, 2, D, "attention.output.dense{0,1}.bias");
    const float* sg[2] = {sd->self_ln_g[0][i], sd->self_ln_g[1][i]};
    const float* sb[2] = {sd->self_ln_b[0][i], sd->self_ln_b[1][i]};
    o.self_ln_g[i] = p.P(sg, 2, D, "attention.output.LayerNorm{A,B}.weight"); o.self_ln_b[i] = p.P(sb, 2, D, "attention.output.LayerNorm{A,B}.bias");
    const float* cq_w[2] = {sd->cross_q_w[0][i], sd->cross_q_w[1][i]};
    const float* cq_b[2] = {sd->cross_q_b[0][i], sd->cross_q_b[1][i]};
    o.cross_q_w[i] = p.W(cq_w, 2, D * D, "crossattention.self{0,1}.query.weight"); o.cross_q_b[i] = p.P(cq_b, 2, D, "crossattention.self{0,1}.query.bias");
    // K0;V0;K1;V1: both streams project the SAME candidate tokens (src/nlvr_encoder.py:158-159)
    const float* kv_w[4] = {sd->cross_k_w[0][i], sd->cross_v_w[0][i], sd->cross_k_w[1][i], sd->cross_v_w[1][i]};
    const float* kv_b[4] = {sd->cross_k_b[0][i], sd->cross_v_b[0][i], sd->cross_k_b[1][i], sd->cross_v_b[1][i]};
    o.cross_kv_w[i] = p.W(kv_w, 4, D * D, "crossattention.self{0,1}.{key,value}.weight");
    o.cross_kv_b[i] = p.P(kv_b, 4, D, "crossattention.self{0,1}.{key,value}.bias");
    // merged output projection [768, 1536] = [ . dense0 | . dense1 ]
    char* cw = (char*)p.take((size_t)D * 2 * D * p.es);
    float* cb = (float*)p.take((size_t)D * 4);
    o.cross_out_w[i] = cw; o.cross_out_b[i] = cb;
    if (!p.dry && p.ok) {
      const float *W0 = sd->cross_out_w[0][i], *W1 = sd->cross_out_w[1][i], *b0 = sd->cross_out_b[0][i], *b1 = sd->cross_out_b[1][i];
      if (!W0 || !W1 || !b0 || !b1) { if (p.rc == CIR_OK) cir_set_error("pack_stage2: missing crossattention.output.dense{0,1} of layer %d", i); p.fail(CIR_EINVAL); }
      else if (i >= 6) {                                      // mergeMLP: merge_layer(cat[dense0(c0), dense1(c1)]), no activation (:252-254)
        if (!sd->merge_w[i] || !sd->merge_b[i]) {
          if (p.rc == CIR_OK) cir_set_error("pack_stage2: layer %d has no crossattention.output.merge_layer.*: a BLIP base checkpoint carries no trained merge "
                        "layers (the reference would leave them randomly initialised, src/blip_stage2.py:188-191); stage II needs a "
                        "fine-tuned BLIP_NLVR checkpoint", i);
          p.fail(CIR_EINVAL);
        } else {
          fold_merge_kernel<<<dim3(3, (unsigned)D), 256, 0, ctx->stream>>>(sd->merge_w[i], 0, W0, cw, 0, p.bf);
          fold_merge_kernel<<<dim3(3, (unsigned)D), 256, 0, ctx->stream>>>(sd->merge_w[i], (int)D, W1, cw, (int)D, p.bf);
          fold_merge_bias_kernel<<<3, 256, 0, ctx->stream>>>(sd->merge_w[i], b0, b1, sd->merge_b[i], cb);
          ctx->launches += 3;
        }
      } else {                                                // mergeAvg: (dense0(c0) + dense1(c1)) / 2 (:257-258)
        half_kernel<<<p.blocks(D * D), 256, 0, ctx->stream>>>(W0, cw, D, D, 2 * D, 0, p.bf);
        half_kernel<<<p.blocks(D * D), 256, 0, ctx->stream>>>(W1, cw, D, D, 2 * D, D, p.bf);
        half_sum_kernel<<<3, 256, 0, ctx->stream>>>(b0, b1, cb, (int)D);
        ctx->launches += 3;
      }
    }
    const float* cg[2] = {sd->cross_ln_g[0][i], sd->cross_ln_g[1][i]};
    const float* cbv[2] = {sd->cross_ln_b[0][i], sd->cross_ln_b[1][i]};
    o.cross_ln_g[i] = p.P(cg, 2, D, "crossattention.output.LayerNorm{A,B}.weight"); o.cross_ln_b[i] = p.P(cbv, 2, D, "crossattention.output.LayerNorm{A,B}.bias");
    o.ffn1_w[i] = p.W1(sd->ffn1_w[i], F * D, "intermediate.dense.weight"); o.ffn1_b[i] = p.P1(sd->ffn1_b[i], F, "intermediate.dense.bias");
    o.ffn2_w[i] = p.W1(sd->ffn2_w[i], D * F, "output.dense.weight"); o.ffn2_b[i] = p.P1(sd->ffn2_b[i], D, "output.dense.bias");
    o.ffn_ln_g[i] = p.P1(sd->ffn_ln_g[i], D, "output.LayerNorm.weight"); o.ffn_ln_b[i] = p.P1(sd->ffn_ln_b[i], D, "output.LayerNorm.bias");
  }
  o.cls0_w = p.W1(sd->cls0_w, D * 2 * D, "cls_head.0.weight"); o.cls0_b = p.P1(sd->cls0_b, D, "cls_head.0.bias");
  o.cls2_w = p.P1(sd->cls2_w, D, "cls_head.2.weight");        // class-0 row only (src/blip_stage2.py:136)
  o.cls2_b = p.P1(sd->cls2_b, 1, "cls_head.2.bias");
  if (need) *need = align_up(p.off, 256);
  if (w && !p.dry) *w = o;
  return p.finish("pack_stage2");
}

}  // namespace

extern "C" size_t cir_pack_vit_bytes(const cir_ctx* ctx, int64_t num_tokens) {
  cir_vit_state sd{}; sd.num_tokens = num_tokens; size_t need = 0;
  pack_vit(const_cast<cir_ctx*>(ctx), &sd, nullptr, 0, nullptr, &need);
  return need;
}
extern "C" int cir_pack_vit_weights(cir_ctx* ctx, const cir_vit_state* sd, void* blob, size_t blob_bytes, cir_vit_weights* out) {
  CIR_ENTER(ctx);
  CIR_CHECK_ARG(sd && blob && out && sd->num_tokens > 0, "pack_vit: null argument");
  return pack_vit(ctx, sd, blob, blob_bytes, out, nullptr);
}
extern "C" size_t cir_pack_stage1_bytes(const cir_ctx* ctx, int64_t vocab_rows, int64_t pos_rows) {
  cir_stage1_state sd{}; sd.emb.vocab_rows = vocab_rows; sd.emb.pos_rows = pos_rows; size_t need = 0;
  pack_stage1(const_cast<cir_ctx*>(ctx), &sd, nullptr, 0, nullptr, &need);
  return need;
}
extern "C" int cir_pack_stage1_weights(cir_ctx* ctx, const cir_stage1_state* sd, void* blob, size_t blob_bytes, cir_stage1_weights* out) {
  CIR_ENTER(ctx);
  CIR_CHECK_ARG(sd && blob && out && sd->emb.vocab_rows > 0 && sd->emb.pos_rows > 0, "pack_stage1: null argument");
  return pack_stage1(ctx, sd, blob, blob_bytes, out, nullptr);
}
extern "C" size_t cir_pack_stage2_bytes(const cir_ctx* ctx, int64_t vocab_rows, int64_t pos_rows) {
  cir_stage2_state sd{}; sd.emb.vocab_rows = vocab_rows; sd.emb.pos_rows = pos_rows; size_t need = 0;
  pack_stage2(const_cast<cir_ctx*>(ctx), &sd, nullptr, 0, nullptr, &need);
  return need;
}
extern "C" int cir_pack_stage2_weights(cir_ctx* ctx, const cir_stage2_state* sd, void* blob, size_t blob_bytes, cir_stage2_weights* out) {
  CIR_ENTER(ctx);
  CIR_CHECK_ARG(sd && blob && out && sd->emb.vocab_rows > 0 && sd->emb.pos_rows > 0, "pack_stage2: null argument");
  return pack_stage2(ctx, sd, blob, blob_bytes, out, nullptr);
}
