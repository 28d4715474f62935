"""CPU oracle for the stage-II re-ranker hot path (and the stage-I pieces feeding it).

TEST INFRASTRUCTURE ONLY.  This file is a plain fp32 restatement, on the CPU, of the
reference's PyTorch arithmetic for the path SURVEY.md section 8 names.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it, and only as the checker or the timed CPU baseline -- never as a product
path (the product fails loudly when the CUDA library is missing).

Parity pin: the reference ships no tests or golden vectors, so the pin is
manufactured: ``tests/golden/make_golden.py`` imports the UNMODIFIED reference modules from
/root/reference behind import shims, loads the same seeded ``state_dict`` and stores the
reference's outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against those files (observed max |diff| is recorded in the fixture metadata).

Every function cites the reference file:line it follows (paths relative to /root/reference).
All functions take a ``state_dict`` with the reference's key names.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
HEADS = 12                      # configs/med_config.json:13 ; src/blip.py:199-200
BERT_EPS = 1e-12                # configs/med_config.json:11
VIT_EPS = 1e-6                  # src/vit.py:142


def _lin(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(sd: SD, name: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _split_heads(x: torch.Tensor, heads: int) -> torch.Tensor:
    # transpose_for_scores: src/nlvr_encoder.py:135-138, src/med.py:153-156
    b, n, d = x.shape
    return x.view(b, n, heads, d // heads).permute(0, 2, 1, 3)


def bert_attention_core(sd: SD, prefix: str, hidden: torch.Tensor, kv_src: torch.Tensor,
                        add_mask: Optional[torch.Tensor], heads: int = HEADS) -> torch.Tensor:
    """BertSelfAttention.forward, src/nlvr_encoder.py:140-222 == src/med.py:158-240.
    q from ``hidden``; k/v from ``kv_src`` (== hidden for self-attention, the image
    tokens for cross-attention); scores divided by sqrt(dh) AFTER QK^T (:193), additive
    mask (:196), softmax (:199), @V (:213), heads merged (:215-217)."""
    q = _split_heads(_lin(sd, prefix + ".query", hidden), heads)
    k = _split_heads(_lin(sd, prefix + ".key", kv_src), heads)
    v = _split_heads(_lin(sd, prefix + ".value", kv_src), heads)
    s = torch.matmul(q, k.transpose(-1, -2))
    s = s / math.sqrt(q.shape[-1])
    if add_mask is not None:
        s = s + add_mask
    p = torch.softmax(s, dim=-1)
    c = torch.matmul(p, v)
    c = c.permute(0, 2, 1, 3).contiguous()
    return c.view(c.shape[0], c.shape[1], -1)


def self_attention_mask(attention_mask: torch.Tensor) -> torch.Tensor:
    """get_extended_attention_mask (encoder branch): src/nlvr_encoder.py:760,770-775."""
    m = attention_mask[:, None, None, :].to(torch.float32)
    return (1.0 - m) * -10000.0


def bert_embeddings(sd: SD, ids: torch.Tensor, prefix: str = "text_encoder.embeddings") -> torch.Tensor:
    """BertEmbeddings.forward: src/nlvr_encoder.py:68-91 == src/med.py:86-110."""
    L = ids.shape[1]
    e = sd[prefix + ".word_embeddings.weight"][ids] + sd[prefix + ".position_embeddings.weight"][:L][None]
    return _ln(sd, prefix + ".LayerNorm", e, BERT_EPS)


# --------------------------------------------------------------------------- stage II

def stage2_layer(sd: SD, i: int, h: List[torch.Tensor], smask: torch.Tensor, cand: torch.Tensor) -> List[torch.Tensor]:
    """One dual-stream BertLayer: src/nlvr_encoder.py:414-476 with BertAttention
    :310-368 and BertSelfOutput :247-270."""
    p = f"text_encoder.encoder.layer.{i}."
    # twin self-attention, no merge (:288-289, :261-264)
    a = []
    for s, ln in ((0, "LayerNormA"), (1, "LayerNormB")):
        c = bert_attention_core(sd, f"{p}attention.self{s}", h[s], h[s], smask)
        d = _lin(sd, f"{p}attention.output.dense{s}", c)
        a.append(_ln(sd, f"{p}attention.output.{ln}", d + h[s], BERT_EPS))
    # twin cross-attention onto the same candidate tokens; cross mask is all ones -> +0 (:158-160,:866-869)
    o = []
    for s in (0, 1):
        c = bert_attention_core(sd, f"{p}crossattention.self{s}", a[s], cand, None)
        o.append(_lin(sd, f"{p}crossattention.output.dense{s}", c))
    if i >= 6:      # mergeMLP, activation commented out (:252-254, :286)
        m = _lin(sd, f"{p}crossattention.output.merge_layer", torch.cat([o[0], o[1]], dim=-1))
    else:           # mergeAvg (:257-258)
        m = (o[0] + o[1]) / 2
    x = [_ln(sd, f"{p}crossattention.output.LayerNormA", m + a[0], BERT_EPS),
         _ln(sd, f"{p}crossattention.output.LayerNormB", m + a[1], BERT_EPS)]      # (:256,:260)
    # FFN with weights shared by both streams (:469-476, :381-382, :394-396); erf GELU
    out = []
    for s in (0, 1):
        t = F.gelu(_lin(sd, p + "intermediate.dense", x[s]))
        t = _lin(sd, p + "output.dense", t)
        out.append(_ln(sd, p + "output.LayerNorm", t + x[s], BERT_EPS))
    return out


def stage2_features(sd: SD, z_t: torch.Tensor, ids: torch.Tensor, amask: torch.Tensor,
                    cand: torch.Tensor) -> torch.Tensor:
    """nlvr_encoder.BertModel.forward (src/nlvr_encoder.py:777-909) as called by
    BLIP_NLVR.img_txt_fusion_val (src/blip_stage2.py:101-130).
    z_t [1,L,768]; ids/amask [1,L]; cand [K,N,768] -> [K,1536]."""
    K = cand.shape[0]
    emb = bert_embeddings(sd, ids)                                  # :880-886
    assert z_t.shape == emb.shape                                   # :891
    h = [z_t.expand(K, -1, -1), emb.expand(K, -1, -1)]              # :892 ; blip_stage2.py:118-124
    smask = self_attention_mask(amask)                              # :849
    for i in range(12):                                             # :506
        h = stage2_layer(sd, i, h, smask, cand)
    return torch.cat((h[0][:, 0, :], h[1][:, 0, :]), dim=-1)        # :906-909


def stage2_head(sd: SD, feats: torch.Tensor) -> torch.Tensor:
    """cls_head + class-0 logit: src/blip_stage2.py:50-54,134-136."""
    return _lin(sd, "cls_head.2", F.relu(_lin(sd, "cls_head.0", feats)))[:, 0]


def stage2_score(sd: SD, z_t: torch.Tensor, ids: torch.Tensor, amask: torch.Tensor,
                 cand: torch.Tensor) -> torch.Tensor:
    """BLIP_NLVR.img_txt_fusion_val: src/blip_stage2.py:101-136 -> [K] logits."""
    return stage2_head(sd, stage2_features(sd, z_t, ids, amask, cand))


# --------------------------------------------------------------------------- stage I

def stage1_hidden(sd: SD, ref_tokens: torch.Tensor, ids: torch.Tensor, amask: torch.Tensor) -> torch.Tensor:
    """med.BertModel.forward multimodal mode (src/med.py:685-821) with BertLayer
    :348-398 as called by BLIP_Retrieval.img_txt_fusion (src/blip_stage1.py:67-80).
    ref_tokens [B,N,768]; ids/amask [B,L] -> last_hidden_state [B,L,768] (== z_t)."""
    h = bert_embeddings(sd, ids)
    smask = self_attention_mask(amask)
    for i in range(12):
        p = f"text_encoder.encoder.layer.{i}."
        c = bert_attention_core(sd, p + "attention.self", h, h, smask)
        a = _ln(sd, p + "attention.output.LayerNorm", _lin(sd, p + "attention.output.dense", c) + h, BERT_EPS)
        c = bert_attention_core(sd, p + "crossattention.self", a, ref_tokens, None)
        x = _ln(sd, p + "crossattention.output.LayerNorm", _lin(sd, p + "crossattention.output.dense", c) + a, BERT_EPS)
        t = _lin(sd, p + "output.dense", F.gelu(_lin(sd, p + "intermediate.dense", x)))
        h = _ln(sd, p + "output.LayerNorm", t + x, BERT_EPS)
    return h


def stage1_query_embedding(sd: SD, hidden: torch.Tensor) -> torch.Tensor:
    """normalize(text_proj(CLS)): src/blip_stage1.py:83."""
    return F.normalize(_lin(sd, "text_proj", hidden[:, 0, :]), dim=-1)


def stage1_gallery_embedding(sd: SD, tokens: torch.Tensor) -> torch.Tensor:
    """normalize(vision_proj(CLS)): src/blip_stage1.py:57."""
    return F.normalize(_lin(sd, "vision_proj", tokens[:, 0, :]), dim=-1)


# --------------------------------------------------------------------------- ViT

def vit_forward(sd: SD, images: torch.Tensor, prefix: str = "visual_encoder.", heads: int = HEADS) -> torch.Tensor:
    """VisionTransformer.forward (src/vit.py:180-194); Block :107-110; Attention
    :70-86; Mlp :35-41; timm 0.4.12 PatchEmbed = Conv2d(3,768,16,16) then
    flatten(2).transpose(1,2) (call site src/vit.py:144-145,182)."""
    B = images.shape[0]
    x = F.conv2d(images, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"], stride=16)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((sd[prefix + "cls_token"].expand(B, -1, -1), x), dim=1)
    x = x + sd[prefix + "pos_embed"][:, : x.shape[1], :]
    D = x.shape[-1]
    dh = D // heads
    for i in range(12):
        b = f"{prefix}blocks.{i}."
        y = _ln(sd, b + "norm1", x, VIT_EPS)
        N = y.shape[1]
        qkv = _lin(sd, b + "attn.qkv", y).reshape(B, N, 3, heads, dh).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * (dh ** -0.5)              # scale BEFORE softmax as a multiply (:74)
        attn = attn.softmax(dim=-1)
        y = (attn @ v).transpose(1, 2).reshape(B, N, D)
        x = x + _lin(sd, b + "attn.proj", y)
        y = _ln(sd, b + "norm2", x, VIT_EPS)
        y = _lin(sd, b + "mlp.fc2", F.gelu(_lin(sd, b + "mlp.fc1", y)))
        x = x + y
    return _ln(sd, prefix + "norm", x, VIT_EPS)


# --------------------------------------------------------------------------- stage-I similarity / top-K

def stage1_topk(q_emb: torch.Tensor, g_emb: torch.Tensor, exclude: Optional[torch.Tensor], k: int
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """src/validate.py:57-58 (FIQ) / :202-210 (CIRR): distances = 1 - q @ G.T, ascending
    argsort over the whole gallery, CIRR drops the query's own reference image, keep [:, :K].
    Ties (the reference's argsort is unstable, so unpinned there) break lowest index first.
    Returns (distances[Q,K] fp32, indices[Q,K] int64)."""
    dist = 1 - q_emb.float() @ g_emb.float().T
    order = torch.sort(dist, dim=-1, stable=True).indices
    rows = []
    for qi in range(order.shape[0]):
        o = order[qi]
        if exclude is not None and int(exclude[qi]) >= 0:
            o = o[o != int(exclude[qi])]
        rows.append(o[:k])
    idx = torch.stack(rows)
    return torch.gather(dist, 1, idx), idx


def cirr_stage1_lists(q_emb: torch.Tensor, g_emb: torch.Tensor, ref_idx: Sequence[int], target_idx: Sequence[int],
                      group_noref: Sequence[Sequence[int]]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The index bookkeeping of compute_cirr_val_metrics, src/validate.py:202-218, on gallery ROW indices instead of names:
    full ascending ranking, reference removed (:207-210), labels (:213-214), ``group_labels = labels[group_mask]`` (:216-218).
    -> (sorted rows [Q,G-1] int64, labels bool [Q,G-1], group_labels bool [Q,P])."""
    dist = 1 - q_emb.float() @ g_emb.float().T
    order = torch.sort(dist, dim=-1, stable=True).indices
    Q, G = order.shape
    ref = torch.as_tensor(np.asarray(ref_idx), dtype=torch.int64)
    tgt = torch.as_tensor(np.asarray(target_idx), dtype=torch.int64)
    noref = order[order != ref[:, None]].reshape(Q, G - 1)
    labels = noref == tgt[:, None]
    members = torch.as_tensor(np.asarray(group_noref), dtype=torch.int64)
    group_mask = (noref[..., None] == members[:, None, :]).sum(-1).bool()
    group_labels = labels[group_mask].reshape(Q, -1)
    return noref, labels, group_labels


# --------------------------------------------------------------------------- re-sort + recall

def rerank_order(scores: torch.Tensor) -> torch.Tensor:
    """argsort(logits, descending): src/validate_stage2.py:53,174,190. Ties -> lowest index first."""
    return torch.sort(scores.float(), dim=-1, descending=True, stable=True).indices


def sorted_labels(scores: torch.Tensor, k_labels: np.ndarray) -> torch.Tensor:
    """np.take_along_axis(K_labels, sorted_indices): src/validate_stage2.py:56-57,178-179."""
    order = rerank_order(scores).numpy()
    return torch.tensor(np.take_along_axis(np.asarray(k_labels), order, axis=1))


def recall_at(labels: torch.Tensor, ks: Sequence[int]) -> List[float]:
    """(sum(labels[:, :k]) / len(labels)).item() * 100: src/validate_stage2.py:60-62,196-203."""
    return [(torch.sum(labels[:, :k]) / len(labels)).item() * 100 for k in ks]


def cirr_metrics(scores: torch.Tensor, k_labels: np.ndarray, group_scores: torch.Tensor,
                 group_is_target: np.ndarray) -> Tuple[float, ...]:
    """compute_cirr_val_metrics: src/validate_stage2.py:153-206. ``group_is_target`` [Q,5]
    bool marks which group member is the target (names compared at :192-193)."""
    labels = sorted_labels(scores, k_labels)
    glabels = sorted_labels(group_scores, group_is_target)
    r1, r5, r10, r50 = recall_at(labels, (1, 5, 10, 50))
    g1, g2, g3 = recall_at(glabels, (1, 2, 3))
    return g1, g2, g3, r1, r5, r10, r50


def fiq_metrics(scores: torch.Tensor, k_labels: np.ndarray) -> Tuple[float, float]:
    """compute_fiq_val_metrics: src/validate_stage2.py:33-66 -> (R@10, R@50)."""
    labels = sorted_labels(scores, k_labels)
    r10, r50 = recall_at(labels, (10, 50))
    return r10, r50


# --------------------------------------------------------------------------- whole-pipeline driver (per-query loop)

NEG_FILL = -99999.99            # src/validate_stage2.py:123,258


def stage2_predictions(sd1: SD, sd2: SD, gallery_tokens: torch.Tensor, ref_idx: torch.Tensor,
                       ids: torch.Tensor, amask: torch.Tensor, cand_idx: torch.Tensor,
                       k_labels: Optional[np.ndarray] = None) -> torch.Tensor:
    """generate_{cirr,fiq}_val_predictions main branch: src/validate_stage2.py:94-125,235-258.
    Per query: z_t from stage I on the reference image's (stage-II-ViT) tokens, gather the K
    candidates' tokens, score; rows with no positive in K_labels are filled with -99999.99."""
    Q, K = cand_idx.shape
    out = torch.empty(Q, K)
    for q in range(Q):
        if k_labels is not None and not bool(np.asarray(k_labels[q]).any()):
            out[q] = NEG_FILL
            continue
        r = gallery_tokens[int(ref_idx[q])][None]
        z_t = stage1_hidden(sd1, r, ids[q:q + 1], amask[q:q + 1])
        cand = gallery_tokens[cand_idx[q].long()]
        out[q] = stage2_score(sd2, z_t, ids[q:q + 1], amask[q:q + 1], cand)
    return out
