"""Shape / edge-case sweep of the bf16 production path against the CPU oracle on random tokens:
caption lengths that are not multiples of 16, L > 32 and > 64, K = 1, ragged masks, several candidate
runs per chunk.  (ViT is skipped: gallery tokens are random LayerNorm-like rows.)"""
import numpy as np
import pytest
import torch

import cir_b200 as cir
from helpers import weights
from oracle import cir_oracle as O

pytestmark = pytest.mark.gpu
syn = cir.synthetic


@pytest.fixture(scope="module")
def models():
    sd1, sd2 = weights(0, "dense", 1.0)
    m1 = cir.blip_stage1.blip_stage1(image_size=384, state_dict=sd1, precision="bf16")
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="bf16")
    g = torch.Generator().manual_seed(11)
    tokens = torch.randn(7, 577, 768, generator=g)
    return sd1, sd2, m1, m2, tokens


@pytest.mark.parametrize("Q,K,L,min_len", [(2, 1, 5, None), (3, 5, 13, 7), (2, 3, 40, 33), (1, 2, 70, None), (4, 6, 32, 9)])
def test_stage2_vs_oracle_shapes(models, Q, K, L, min_len):
    sd1, sd2, m1, m2, tokens = models
    G = tokens.shape[0]
    ref, tgt, ids, mask = syn.make_queries(Q, G, L, seed=100 + L, min_len=min_len)
    g = torch.Generator().manual_seed(L)
    cand = torch.stack([torch.randperm(G, generator=g)[:K] for _ in range(Q)]).int()
    tok_d = tokens.cuda().bfloat16()
    tok_ref = tok_d.float().cpu()                      # the oracle sees the same (bf16-rounded) tokens
    z_t, _ = m1.encode_queries(tok_d, ref.int(), ids, mask, want_z=True, want_emb=False)
    s = m2.score_triplets(z_t, ids, mask, tok_d, cand.numpy())
    with torch.no_grad():
        want = O.stage2_predictions(sd1, sd2, tok_ref, ref, ids, mask, cand)
    err = (s.cpu() - want).abs().max().item()
    assert err <= 2e-2, (err, s.cpu(), want)
    # drop-in single-query call (K == 1 included: the reference special-cases it at validate_stage2.py:110-112)
    tb = syn.TokenBatch(input_ids=ids[:1], attention_mask=mask[:1])
    z = m1.img_txt_fusion(tok_d[int(ref[0])][None], None, tb, train=False, return_raw=True)
    s0 = m2.img_txt_fusion_val(z, tok_d[cand[0].long().cuda()], tb)
    assert s0.shape == (K,) and (s0.cpu() - want[0]).abs().max() <= 2e-2


def test_empty_inputs(models):
    sd1, sd2, m1, m2, tokens = models
    tok_d = tokens.cuda().bfloat16()
    ref, tgt, ids, mask = syn.make_queries(3, tokens.shape[0], 8, seed=1)
    z_t, _ = m1.encode_queries(tok_d, ref.int(), ids, mask, want_z=True, want_emb=False)
    s = m2.score_triplets(z_t, ids, mask, tok_d, np.zeros((3, 4), np.int32), row_active=np.zeros(3, bool))
    assert torch.all(s == -99999.99)
    eng = m2.engine
    assert eng.rerank_sort(torch.zeros(0, 5)).shape == (0, 5)
    td, ti = eng.stage1_topk(torch.zeros(0, 256), torch.randn(10, 256), 3)
    assert td.shape == (0, 3)
    td, ti = eng.stage1_topk(torch.nn.functional.normalize(torch.randn(2, 256), dim=-1), torch.nn.functional.normalize(torch.randn(3, 256), dim=-1), 5)
    assert (ti[:, 3:] == -1).all() and torch.isinf(td[:, 3:]).all()        # fewer gallery rows than K


def test_max_sizes_sort_and_topk():
    eng = cir.engine.get_engine(precision="bf16")
    g = torch.Generator().manual_seed(21)
    s = torch.randn(3, 2048, generator=g)
    s[:, 100:200] = s[:, :1]                                  # a block of ties
    assert torch.equal(eng.rerank_sort(s).cpu().long(), O.rerank_order(s))
    q = torch.nn.functional.normalize(torch.randn(4, 256, generator=g), dim=-1)
    gal = torch.nn.functional.normalize(torch.randn(6000, 256, generator=g), dim=-1)
    dist = 1 - q @ gal.T
    td, ti = eng.topk_from_dist(dist, 1024)
    wd, wi = O.stage1_topk(q, gal, None, 1024)
    assert torch.equal(ti.cpu().long(), wi) and torch.equal(td.cpu(), wd)


def test_vit_other_image_size():
    """224 px (197 tokens): the ViT path is not specialised to 384 px."""
    sd = syn.make_stage2_state_dict(3, 224, "dense")
    m = cir.blip_stage2.blip_stage2(image_size=224, state_dict=sd, precision="fp32")
    images = syn.make_images(2, 224, seed=5)
    tok = m.img_embed(images)
    assert tok.shape == (2, 197, 768)
    with torch.no_grad():
        want = O.vit_forward(sd, images)
    assert (tok.cpu() - want).abs().max() < 2e-4
    mb = cir.blip_stage2.blip_stage2(image_size=224, state_dict=sd, precision="bf16")
    tb = mb.img_embed(images).float().cpu()
    assert (tb - want).abs().mean() < 0.02
