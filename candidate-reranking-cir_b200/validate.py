"""Stage-I evaluation / candidate-filtering drivers (reference: src/validate.py:33-99,176-268 and
src/utils.py:25-72).  The reference materialises ``1 - q @ G.T`` and fully argsorts it on the
device, then maps names and removes the reference image on the host (src/validate.py:202-210);
here the similarity tiles and the per-query top-K (with the reference index excluded in-kernel)
are fused, so nothing of size [Q, G] is ever stored."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from .blip import tokenize
from .validate_stage2 import LENGTH_BUCKET, _fiq_captions, _length_buckets, _name_index, _percent, _tokens


def extract_index_features(images: torch.Tensor, index_names: List[str], blip_model, blip_stage2=False, blip_stage1=False,
                           batch: int = 32):
    """src/utils.py:25-72 without the JPEG DataLoader: ``images`` is the already pre-processed
    [G,3,S,S] tensor (host or device).  blip_stage2 -> (tokens [G,577,768], names);
    blip_stage1 -> (tokens, pooled+normalised [G,256], names)."""
    assert blip_stage1 != blip_stage2, "only one condition shall be selected"      # src/utils.py:44,57
    eng = blip_model.engine
    tokens = eng.vit_forward(blip_model._vit, images, batch=batch)
    if blip_stage2:
        return tokens, list(index_names)
    return tokens, eng.stage1_gallery_embed(blip_model._w, tokens), list(index_names)


def retrieve_topk(blip_model, dataset, index_features, index_features_normed_pooled, index_names, k: int, cirr: bool):
    """Query embeddings (src/validate.py:305-311 / :140-146) + fused distance/top-K
    (src/validate.py:57-58,202-210,257).  Returns (top_dist fp32 [Q,k], top_idx int32 [Q,k]) on device;
    CIRR excludes each query's own reference image and normalises the query embedding twice."""
    eng = blip_model.engine
    n2i = _name_index(index_names)
    ref_idx = np.array([n2i[n] for n in dataset.reference_names], dtype=np.int32)
    caps = list(dataset.captions) if cirr else _fiq_captions(dataset.captions)
    ids, mask = _tokens(blip_model, dataset, caps)
    gallery = eng.to_act(index_features)
    q_emb = torch.empty(len(ref_idx), 256, dtype=torch.float32, device=eng.device)
    for rows, L in _length_buckets(mask, LENGTH_BUCKET):                            # each query at (about) its own length
        r_t = torch.from_numpy(rows).to(eng.device)
        _, qe = blip_model.encode_queries(gallery, ref_idx[rows], ids[r_t, :L].contiguous(), mask[r_t, :L].contiguous(),
                                          want_z=False, want_emb=True, normalize_twice=cirr)
        q_emb[r_t] = qe
    g_emb = index_features_normed_pooled.float()                                    # "already normed" (:55,:199)
    return eng.stage1_topk(q_emb, g_emb, k, exclude=ref_idx if cirr else None)


def _labels_and_recalls(eng, top_idx, target_idx, ks):
    labels = top_idx.to(torch.int64) == torch.as_tensor(target_idx, device=top_idx.device)[:, None]
    ident = torch.arange(top_idx.shape[1], dtype=torch.int32, device=top_idx.device).expand_as(top_idx).contiguous()
    hits = eng.recall_counts(labels, ident, ks)
    return labels, [_percent(h, len(labels)) for h in hits]


def compute_fiq_val_metrics(relative_val_dataset, blip_model, index_features, index_features_normed_pooled, index_names,
                            k: int = 100) -> Tuple[float, float, Dict]:
    """src/validate.py:33-99 -> (recall@10, recall@50, topk dict shaped like the saved file :87-94)."""
    eng = blip_model.engine
    k = min(max(k, 50), len(index_names))
    top_dist, top_idx = retrieve_topk(blip_model, relative_val_dataset, index_features, index_features_normed_pooled, index_names, k, cirr=False)
    n2i = _name_index(index_names)
    tgt = np.array([n2i[n] for n in relative_val_dataset.target_names])
    labels, (r10, r50) = _labels_and_recalls(eng, top_idx, tgt, (10, 50))
    names = np.array(index_names)[top_idx.cpu().numpy()]
    topk = {"sorted_index_names": names, "target_names": list(relative_val_dataset.target_names), "index_names": list(index_names),
            "labels": labels.cpu(), "split": relative_val_dataset.split, "dress_types": ",".join(relative_val_dataset.dress_types)}
    return r10, r50, topk


def compute_cirr_val_metrics(relative_val_dataset, blip_model, index_features, index_features_normed_pooled, index_names,
                             k: int = 50):
    """src/validate.py:176-268 -> (recall@1, recall@5, recall@10, recall@50, topk dict :256-263).
    (The subset/group recalls of the stage-I driver need the full ranking of the 5 group members; they
    are produced by the stage-II driver, src/validate_stage2.py:186-203.)"""
    eng = blip_model.engine
    k = min(max(k, 50), len(index_names) - 1)          # the reference image is removed from every row (:207-210)
    top_dist, top_idx = retrieve_topk(blip_model, relative_val_dataset, index_features, index_features_normed_pooled, index_names, k, cirr=True)
    n2i = _name_index(index_names)
    tgt = np.array([n2i[n] for n in relative_val_dataset.target_names])
    labels, (r1, r5, r10, r50) = _labels_and_recalls(eng, top_idx, tgt, (1, 5, 10, 50))
    names = np.array(index_names)[top_idx.cpu().numpy()]
    topk = {"sorted_index_names": names, "target_names": list(relative_val_dataset.target_names), "index_names": list(index_names),
            "labels": labels.cpu(), "split": relative_val_dataset.split}
    return r1, r5, r10, r50, topk
