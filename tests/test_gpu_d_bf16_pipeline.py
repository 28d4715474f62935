"""Production bf16 path (tcgen05 GEMMs) end to end against the reference's golden outputs.
Tolerance: |score - reference| <= 2e-2 absolute (BASELINE.json north_star), plus a tighter relative
check on the 1536-d pre-head features."""
import numpy as np
import pytest
import torch

import cir_b200 as cir
from helpers import golden_weights, load_golden
from oracle import cir_oracle as O

pytestmark = pytest.mark.gpu
syn = cir.synthetic
SCORE_TOL = 2e-2


@pytest.fixture(scope="module")
def small():
    g = load_golden("pipeline_small.npz")
    sd1, sd2 = golden_weights(g)
    m1 = cir.blip_stage1.blip_stage1(image_size=384, state_dict=sd1, precision="bf16")
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="bf16")
    images = syn.make_images(int(g["G"]), 384, seed=1)
    tokens2 = m2.img_embed(images)
    return g, m1, m2, images, tokens2


def test_vit_tokens_bf16(small):
    g, m1, m2, images, tokens2 = small
    assert tokens2.dtype == torch.bfloat16
    t = tokens2.float().cpu()
    err = np.abs(t[:, ::48, ::16].numpy() - g["tokens2_sample"])
    assert err.max() < 0.15 and err.mean() < 0.02, (err.max(), err.mean())      # tokens are LayerNorm outputs, |x| up to ~5


def test_pipeline_scores_bf16(small):
    g, m1, m2, images, tokens2 = small
    tb = syn.TokenBatch(input_ids=torch.tensor(g["ids"]), attention_mask=torch.tensor(g["mask"]))
    ref_idx = torch.tensor(g["ref_idx"])
    z_t, _ = m1.encode_queries(tokens2, ref_idx, tb.input_ids, tb.attention_mask, want_z=True, want_emb=False)
    zerr = np.abs(z_t.float().cpu().numpy() - g["z_t"])
    assert zerr.max() < 0.25 and zerr.mean() < 0.03, (zerr.max(), zerr.mean())
    s = m2.score_triplets(z_t, tb.input_ids, tb.attention_mask, tokens2, g["cand_idx"])
    err = np.abs(s.cpu().numpy() - g["scores"])
    assert err.max() <= SCORE_TOL, (err.max(), s.cpu().numpy(), g["scores"])
    # drop-in per-query call gives the same numbers as the batched candidate-major path
    for q in range(int(g["Q"])):
        tbq = syn.TokenBatch(input_ids=tb.input_ids[q:q + 1], attention_mask=tb.attention_mask[q:q + 1])
        z = m1.img_txt_fusion(tokens2[int(ref_idx[q])][None], None, tbq, train=False, return_raw=True)
        sq = m2.img_txt_fusion_val(z, tokens2[torch.tensor(g["cand_idx"][q]).long().cuda()], tbq)
        assert np.abs(sq.cpu().numpy() - g["scores"][q]).max() <= SCORE_TOL


def test_features_relative_error_bf16(small):
    g, m1, m2, images, tokens2 = small
    ids, mask = torch.tensor(g["ids"]), torch.tensor(g["mask"])
    z_t = torch.tensor(g["z_t"]).cuda().bfloat16()
    ch = cir.schedule.plan_chunks(g["cand_idx"])[0]
    ql = torch.from_numpy(ch.query_list.astype(np.int64)).cuda()
    s, f = m2.engine.stage2_score_chunk(m2._w, tokens2, ch.cand_list, z_t[ql].contiguous(), ids.cuda()[ql], mask.cuda()[ql],
                                        ch.trip_query, ch.trip_slot, want_feats=True)
    want = torch.tensor(g["feats"]).reshape(-1, 1536)[torch.from_numpy(ch.flat_pos)]
    rel = (f.cpu() - want).norm() / want.norm()
    assert rel < 3e-2, rel


def test_L32_reference_init_bf16():
    g = load_golden("stage2_L32.npz")
    sd1, sd2 = golden_weights(g)
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="bf16")
    tokens2 = m2.img_embed(syn.make_images(int(g["G"]), 384, seed=1))
    s = m2.score_triplets(torch.tensor(g["z_t"]).cuda().bfloat16(), torch.tensor(g["ids"]), torch.tensor(g["mask"]), tokens2, g["cand_idx"])
    assert np.abs(s.cpu().numpy() - g["scores"]).max() <= SCORE_TOL


def test_validate_drivers_on_synthetic_dataset(small):
    """validate_stage2-shaped run: Recall@K of the CUDA path == recall computed by the oracle's
    sort/recall arithmetic from the same scores (identical on synthetic data)."""
    g, m1, m2, images, tokens2 = small
    G = int(g["G"])
    names = syn.index_names_for(G)
    Q, K = 5, 4
    ref, tgt, ids, mask = syn.make_queries(Q, G, 10, seed=9, min_len=6)
    cand, labels = syn.make_random_topk(Q, G, K, ref, tgt, seed=10, hit_rate=0.8)
    groups = syn.make_group_members(ref, tgt, G, seed=11)
    tb = syn.TokenBatch(input_ids=ids, attention_mask=mask)
    ds = syn.SyntheticRelativeDataset(names, ref, tgt, ["x"] * Q, cand.numpy(), kind="cirr", group_idx=groups.numpy(), token_batch=tb)
    V2 = cir.validate_stage2
    logits, glogits, rn, tn, gm = V2.generate_cirr_val_predictions(m2, m1, ds, names, tokens2)
    assert logits.shape == (Q, K) and glogits.shape == (Q, 5)
    inactive = ~ds.K_labels.any(1)
    assert torch.all(logits[torch.from_numpy(inactive).cuda()] == -99999.99)
    # length-bucketed driver == one padded batch (padding only adds masked keys)
    z_all, _ = m1.encode_queries(tokens2, ref.int(), ids.clone().index_fill_(1, torch.tensor([0]), 30523), mask, want_z=True, want_emb=False)
    direct = m2.score_triplets(z_all, ids.clone().index_fill_(1, torch.tensor([0]), 30523), mask, tokens2, cand.numpy(), ds.K_labels.any(1))
    assert (direct - logits).abs().max() < 2e-3
    got = V2.compute_cirr_val_metrics(ds, m2, m1, tokens2, names)
    gt = np.array(gm) == np.array(tn)[:, None]
    want = O.cirr_metrics(logits.cpu(), ds.K_labels, glogits.cpu(), gt)
    assert got == want
    r10, r50 = V2.compute_fiq_val_metrics(ds, m2, m1, tokens2, names)
    assert (r10, r50) == O.fiq_metrics(logits.cpu(), ds.K_labels)


def test_last_layer_pruning_bf16(small):
    g, m1, m2, images, tokens2 = small
    ids, mask = torch.tensor(g["ids"]), torch.tensor(g["mask"])
    z_t = torch.tensor(g["z_t"]).cuda().bfloat16()
    eng = m2.engine
    a = m2.score_triplets(z_t, ids, mask, tokens2, g["cand_idx"])
    eng.set_prune_last_layer(False)
    try:
        b = m2.score_triplets(z_t, ids, mask, tokens2, g["cand_idx"])
    finally:
        eng.set_prune_last_layer(True)
    assert (a - b).abs().max() < 5e-3          # different tile shapes / kernels on the last layer, same math
    assert np.abs(a.cpu().numpy() - g["scores"]).max() <= SCORE_TOL


def test_stage1_driver_on_synthetic_dataset(small):
    """validate.py-shaped run: stage-I top-K (reference excluded), labels and recalls vs the oracle on the
    CUDA path's own embeddings."""
    g, m1, m2, images, tokens2 = small
    G = int(g["G"])
    names = syn.index_names_for(G)
    tokens1, g_emb = m1.img_embed(images, return_pool_and_normalized=True)
    Q, K = 5, 4
    ref, tgt, ids, mask = syn.make_queries(Q, G, 10, seed=9, min_len=6)
    tb = syn.TokenBatch(input_ids=ids, attention_mask=mask)
    ds = syn.SyntheticRelativeDataset(names, ref, tgt, ["x"] * Q, np.zeros((Q, K), int), kind="cirr", token_batch=tb)
    V1 = cir.validate
    td, ti = V1.retrieve_topk(m1, ds, tokens1, g_emb, names, K, cirr=True)
    assert ti.shape == (Q, K)
    for q in range(Q):
        assert int(ref[q]) not in ti[q].tolist()
    _, q_emb = m1.encode_queries(tokens1, ref.int(), ids.clone().index_fill_(1, torch.tensor([0]), 30523), mask,
                                 want_z=False, want_emb=True, normalize_twice=True)
    wd, wi = O.stage1_topk(q_emb.cpu(), g_emb.cpu(), ref, K)
    assert (td.cpu() - wd).abs().max() < 2e-3
    (r10, r50), topk = V1.fiq_val_topk(ds, m1, tokens1, g_emb, names, k=K)      # CIRR 7-tuple + group_labels: test_gpu_h_parity_round2.py
    assert topk["sorted_index_names"].shape[0] == Q and topk["labels"].shape[0] == Q
    assert 0.0 <= r10 <= r50 <= 100.0


def test_large_chunk_vs_oracle():
    """A chunk big enough for the cta_group::2 pair tiles, and full 128-row attention
    tiles (T = 104 triplets x 32 rows) against the CPU oracle: |score diff| <= 2e-2."""
    syn_ = cir.synthetic
    g0 = load_golden("pipeline_small.npz")
    sd1, sd2 = golden_weights(g0)
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="bf16")
    g = torch.Generator().manual_seed(6)
    G, Q, K, L = 3, 13, 8, 32
    tokens = torch.randn(G, 577, 768, generator=g).cuda().bfloat16()
    ids, mask = syn_.make_token_ids(Q, L, seed=3, min_len=18)
    ids[:, 0] = syn_.ENC_TOKEN_ID
    z_t = torch.randn(Q, L, 768, generator=g).cuda().bfloat16()
    cand = torch.stack([torch.randint(0, G, (K,), generator=g) for _ in range(Q)]).int()      # heavy candidate reuse
    s = m2.score_triplets(z_t, ids, mask, tokens, cand.numpy())
    tok_ref, z_ref = tokens.float().cpu(), z_t.float().cpu()
    with torch.no_grad():
        want = torch.stack([O.stage2_score(sd2, z_ref[q:q + 1], ids[q:q + 1], mask[q:q + 1], tok_ref[cand[q].long()]) for q in range(Q)])
    err = (s.cpu() - want).abs()
    assert err.max() <= 2e-2, (err.max(), err.mean())


def test_bf16_vs_fp32_check_mode_at_scale():
    """Same weights, same inputs, 1,600 triplets in full-size chunks: the bf16 production path against the fp32 check
    mode of this library (itself within 1e-4 of the reference): the whole error distribution stays inside 2e-2."""
    syn_ = cir.synthetic
    g0 = load_golden("pipeline_small.npz")
    sd1, sd2 = golden_weights(g0)
    m16 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="bf16")
    m32 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="fp32")
    g = torch.Generator().manual_seed(8)
    G, Q, K, L = 24, 32, 50, 32
    tokens = torch.randn(G, 577, 768, generator=g).cuda()
    tokens16 = tokens.bfloat16()
    tokens32 = tokens16.float()                                # identical (bf16-representable) inputs for both modes
    ids, mask = syn_.make_token_ids(Q, L, seed=4, min_len=12)
    ids[:, 0] = syn_.ENC_TOKEN_ID
    z16 = torch.randn(Q, L, 768, generator=g).cuda().bfloat16()
    cand = torch.randint(0, G, (Q, K), generator=g).int().numpy()          # K > G: heavy candidate reuse, duplicates allowed
    a = m16.score_triplets(z16, ids, mask, tokens16, cand)
    b = m32.score_triplets(z16.float(), ids, mask, tokens32, cand)
    err = (a - b).abs()
    assert err.max() <= 2e-2 and err.mean() < 5e-3, (err.max().item(), err.mean().item())


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_first_layer_dedup_is_exact(precision):
    """cir_set_dedup_first_layer: layer 0's query-only work (QKV, masked self-attention, dense + LayerNorm, cross query
    projection) once per unique query and expanded over the triplets == the same work done per triplet."""
    syn_ = cir.synthetic
    sd1, sd2 = golden_weights(load_golden("pipeline_small.npz"))
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision=precision)
    eng = m2.engine
    g = torch.Generator().manual_seed(21)
    G, Q, K, L = 5, 9, 6, 20
    tokens = torch.randn(G, 577, 768, generator=g).cuda().to(eng.act_dtype)
    ids, mask = syn_.make_token_ids(Q, L, seed=7, min_len=9)
    ids[:, 0] = syn_.ENC_TOKEN_ID
    z_t = torch.randn(Q, L, 768, generator=g).cuda().to(eng.act_dtype)
    cand = torch.stack([torch.randint(0, G, (K,), generator=g) for _ in range(Q)]).int().numpy()
    on = m2.score_triplets(z_t, ids, mask, tokens, cand)
    eng.set_dedup_first_layer(False)
    try:
        off = m2.score_triplets(z_t, ids, mask, tokens, cand)
    finally:
        eng.set_dedup_first_layer(True)
    assert torch.isfinite(on).all()
    assert (on - off).abs().max() <= (1e-6 if precision == "fp32" else 1e-3), (on - off).abs().max()


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_query_prefix_is_exact(precision):
    """cir_stage2_prefix + cir_stage2_score_prefixed (layer 0's query-only part once per query set, shared by all chunks)
    == per-chunk computation, over several chunks."""
    syn_ = cir.synthetic
    sd1, sd2 = golden_weights(load_golden("pipeline_small.npz"))
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision=precision)
    eng = m2.engine
    g = torch.Generator().manual_seed(22)
    G, Q, K, L = 7, 10, 6, 16
    tokens = torch.randn(G, 577, 768, generator=g).cuda().to(eng.act_dtype)
    ids, mask = syn_.make_token_ids(Q, L, seed=8, min_len=9)
    ids[:, 0] = syn_.ENC_TOKEN_ID
    z_t = torch.randn(Q, L, 768, generator=g).cuda().to(eng.act_dtype)
    cand = torch.stack([torch.randperm(G, generator=g)[:K] for _ in range(Q)]).int().numpy()
    row_active = np.ones(Q, bool)
    row_active[3] = False
    old = (eng.max_triplets, eng.max_candidates, eng.query_prefix, eng.prefix_batch)
    eng.max_triplets, eng.max_candidates, eng.prefix_batch = 16, 2, 4   # several small chunks, prefix in 3 calls
    try:
        eng.query_prefix = True
        a = m2.score_triplets(z_t, ids, mask, tokens, cand, row_active)
        eng.query_prefix = False
        b = m2.score_triplets(z_t, ids, mask, tokens, cand, row_active)
    finally:
        eng.max_triplets, eng.max_candidates, eng.query_prefix, eng.prefix_batch = old
    assert (a[3] == cir.engine.NEG_FILL).all() and torch.isfinite(a).all()
    assert (a - b).abs().max() <= (1e-6 if precision == "fp32" else 1e-3), (a - b).abs().max()
