"""fp32 check mode end to end through the drop-in module surface, against the golden fixtures the
unmodified reference produced (tolerance 1e-4 on scores, BASELINE.json) and the CPU oracle."""
import numpy as np
import pytest
import torch

import cir_b200 as cir
from helpers import golden_weights, load_golden
from oracle import cir_oracle as O

pytestmark = pytest.mark.gpu
syn = cir.synthetic


@pytest.fixture(scope="module")
def small():
    g = load_golden("pipeline_small.npz")
    sd1, sd2 = golden_weights(g)
    m1 = cir.blip_stage1.blip_stage1(image_size=384, state_dict=sd1, precision="fp32")
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="fp32")
    images = syn.make_images(int(g["G"]), 384, seed=1)
    tokens2 = m2.img_embed(images)
    return g, m1, m2, images, tokens2


def test_vit_tokens(small):
    g, m1, m2, images, tokens2 = small
    t = tokens2.float().cpu()
    assert t.shape == (int(g["G"]), 577, 768)
    assert np.abs(t[:, ::48, ::16].numpy() - g["tokens2_sample"]).max() < 2e-4
    assert np.abs(t[:, 0, :].numpy() - g["tokens2_cls"]).max() < 2e-4
    emb, atts = m2.img_embed(images[:1], atts=True)
    assert atts.shape == (1, 577) and atts.dtype == torch.long


def test_stage1_embeddings_and_topk(small):
    g, m1, m2, images, tokens2 = small
    tokens1, g_emb = m1.img_embed(images, return_pool_and_normalized=True)
    assert np.abs(tokens1[:, 0, :].float().cpu().numpy() - g["tokens1_cls"]).max() < 2e-4
    assert np.abs(g_emb.cpu().numpy() - g["g_emb"]).max() < 2e-5
    tb = syn.TokenBatch(input_ids=torch.tensor(g["ids"]), attention_mask=torch.tensor(g["mask"]))
    ref_idx = torch.tensor(g["ref_idx"])
    q_emb = m1.img_txt_fusion(tokens1[ref_idx.cuda()], None, tb, train=False)
    q2 = torch.nn.functional.normalize(q_emb.cpu())            # host-side re-normalise only to compare
    assert np.abs(q2.numpy() - g["q_emb"]).max() < 2e-5
    _, q_emb2 = m1.encode_queries(tokens1, ref_idx, tb.input_ids.clone().index_fill_(1, torch.tensor([0]), 30523),
                                  tb.attention_mask, want_z=False, want_emb=True, normalize_twice=True)
    assert np.abs(q_emb2.cpu().numpy() - g["q_emb"]).max() < 2e-5
    # top-K from the reference's own fp32 embeddings: indices bit-exact
    td, ti = m1.engine.stage1_topk(torch.tensor(g["q_emb"]), torch.tensor(g["g_emb"]), int(g["K"]), exclude=ref_idx)
    assert np.array_equal(ti.cpu().numpy(), g["cand_idx"])
    assert np.abs(td.cpu().numpy() - g["distances_topk"]).max() < 1e-6


def test_stage1_in_batch_logits(small):
    """BLIP_Retrieval.img_txt_fusion(train=True): predicted @ targets.T / temp (src/blip_stage1.py:88-91), forward only."""
    g, m1, m2, images, tokens2 = small
    tokens1, g_emb = m1.img_embed(images, return_pool_and_normalized=True)
    tb = syn.TokenBatch(input_ids=torch.tensor(g["ids"]), attention_mask=torch.tensor(g["mask"]))
    ref_idx = torch.tensor(g["ref_idx"])
    logits = m1.img_txt_fusion(tokens1[ref_idx.cuda()], g_emb, tb, train=True)
    want = g["q_emb"].astype(np.float64) @ g["g_emb"].astype(np.float64).T / m1.temp
    assert logits.shape == (int(g["Q"]), g["g_emb"].shape[0]) and logits.dtype == torch.float32
    assert np.abs(logits.cpu().numpy() - want).max() < 5e-4


def test_z_t_and_stage2_scores_drop_in(small):
    g, m1, m2, images, tokens2 = small
    Q = int(g["Q"])
    for q in range(Q):
        tb = syn.TokenBatch(input_ids=torch.tensor(g["ids"][q:q + 1]), attention_mask=torch.tensor(g["mask"][q:q + 1]))
        r = tokens2[int(g["ref_idx"][q])][None]
        z = m1.img_txt_fusion(r, None, tb, train=False, return_raw=True)          # validate_stage2.py:243-244
        assert np.abs(z.last_hidden_state.float().cpu().numpy()[0] - g["z_t"][q]).max() < 2e-4
        cand = tokens2[torch.tensor(g["cand_idx"][q]).long().cuda()]                # :251
        s = m2.img_txt_fusion_val(z, cand, tb)                                      # :254
        assert s.shape == (int(g["K"]),) and s.dtype == torch.float32
        assert np.abs(s.cpu().numpy() - g["scores"][q]).max() < 1e-4, (s.cpu().numpy(), g["scores"][q])


def test_batched_candidate_major_matches_reference(small):
    g, m1, m2, images, tokens2 = small
    ids, mask = torch.tensor(g["ids"]), torch.tensor(g["mask"])
    z_t = torch.tensor(g["z_t"]).cuda()
    for (mt, mc) in ((2048, 48), (5, 2), (1, 1)):
        m2.engine.max_triplets, m2.engine.max_candidates = mt, mc
        s = m2.score_triplets(z_t, ids, mask, tokens2, g["cand_idx"])
        assert np.abs(s.cpu().numpy() - g["scores"]).max() < 1e-4
    m2.engine.max_triplets, m2.engine.max_candidates = 4096, 64
    # features (cat of the two CLS vectors) are a stronger check than the 2-way head
    ch = cir.schedule.plan_chunks(g["cand_idx"])[0]
    eng = m2.engine
    ql = torch.from_numpy(ch.query_list.astype(np.int64)).cuda()
    s, f = eng.stage2_score_chunk(m2._w, tokens2, ch.cand_list, z_t[ql], ids.cuda()[ql], mask.cuda()[ql], ch.trip_query,
                                  ch.trip_slot, want_feats=True)
    want = torch.tensor(g["feats"]).reshape(-1, 1536)[torch.from_numpy(ch.flat_pos)]
    assert (f.cpu() - want).abs().max() < 2e-4
    # rows with no positive are filled, not scored
    act = np.array([True, False, True])
    s = m2.score_triplets(z_t, ids, mask, tokens2, g["cand_idx"], row_active=act)
    assert torch.all(s[1] == -99999.99) and np.abs(s[0].cpu().numpy() - g["scores"][0]).max() < 1e-4


def test_rerank_and_recall_from_reference_scores(small):
    g, m1, m2, *_ = small
    eng = m2.engine
    order = eng.rerank_sort(torch.tensor(g["scores"]))
    assert np.array_equal(order.cpu().numpy(), g["order"])
    hits = eng.recall_counts(torch.tensor(g["k_labels"]), order, (1, 2, 3, 4))
    rec = [(torch.tensor(h) / int(g["Q"])).item() * 100 for h in hits]
    assert rec == list(g["recalls"])


def test_training_shape_forward_matches_oracle(small):
    g, m1, m2, images, tokens2 = small
    sd1, sd2 = golden_weights(g)
    ids, mask = torch.tensor(g["ids"]), torch.tensor(g["mask"])
    z_t = torch.tensor(g["z_t"])
    tb = syn.TokenBatch(input_ids=ids, attention_mask=mask)
    out = m2.img_txt_fusion(cir.blip.EncoderOutput(last_hidden_state=z_t.cuda()), tokens2[:3], tb)
    assert out.shape == (3, 3)
    tok_cpu = tokens2[:3].float().cpu()
    with torch.no_grad():
        for i in range(3):
            want = O.stage2_score(sd2, z_t[i:i + 1], ids[i:i + 1], mask[i:i + 1], tok_cpu)
            assert (out[i].cpu() - want).abs().max() < 2e-4


def test_training_shape_forward_matches_reference(small):
    """Both train=True forwards against the outputs of the unmodified reference (tests/golden/training_forward.npz)."""
    g0, m1, m2, images, tokens2 = small
    g = load_golden("training_forward.npz")
    tb = syn.TokenBatch(input_ids=torch.tensor(g["ids"]), attention_mask=torch.tensor(g["mask"]))
    ref_idx, target_idx = torch.tensor(g["ref_idx"]).cuda(), torch.tensor(g["target_idx"]).cuda()
    tokens1, g_emb = m1.img_embed(images, return_pool_and_normalized=True)
    s1 = m1.img_txt_fusion(tokens1[ref_idx], g_emb[target_idx], tb, train=True)                 # src/blip_stage1.py:88-91
    assert np.abs(s1.cpu().numpy() - g["s1_logits"]).max() < 1e-3                               # logits are O(1 / temp)
    z = m1.img_txt_fusion(tokens2[ref_idx], None, tb, train=False, return_raw=True)
    assert np.abs(z.last_hidden_state.float().cpu().numpy() - g["z_t"]).max() < 2e-4
    s2 = m2.img_txt_fusion(z, tokens2[target_idx], tb, train=True)                             # src/blip_stage2.py:65-99
    assert s2.shape == (int(g["B"]), int(g["B"]))
    assert np.abs(s2.cpu().numpy() - g["s2_logits"]).max() < 2e-4


def test_L32_reference_init(small):
    g = load_golden("stage2_L32.npz")
    sd1, sd2 = golden_weights(g)
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="fp32")
    tokens2 = m2.img_embed(syn.make_images(int(g["G"]), 384, seed=1))
    s = m2.score_triplets(torch.tensor(g["z_t"]).cuda(), torch.tensor(g["ids"]), torch.tensor(g["mask"]), tokens2, g["cand_idx"])
    assert np.abs(s.cpu().numpy() - g["scores"]).max() < 1e-4


def test_last_layer_pruning_is_exact(small):
    """Computing layer 11 for the CLS rows only must not change the scores (fp32: same arithmetic per row)."""
    g, m1, m2, images, tokens2 = small
    ids, mask = torch.tensor(g["ids"]), torch.tensor(g["mask"])
    z_t = torch.tensor(g["z_t"]).cuda()
    eng = m2.engine
    a = m2.score_triplets(z_t, ids, mask, tokens2, g["cand_idx"])
    eng.set_prune_last_layer(False)
    try:
        b = m2.score_triplets(z_t, ids, mask, tokens2, g["cand_idx"])
    finally:
        eng.set_prune_last_layer(True)
    assert (a - b).abs().max() < 2e-6
    assert np.abs(b.cpu().numpy() - g["scores"]).max() < 1e-4
