"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It loads the seeded synthetic ``state_dict``s (cir_b200.synthetic) into the reference's own
``BLIP_Retrieval`` / ``BLIP_NLVR`` modules (strict key match), runs the reference methods the
hot path is made of, and stores inputs + outputs as small ``.npz`` files:

  * ``pipeline_small.npz``  -- G=6 images @384, Q=3 ragged 12-token captions, K=4:
      ViT tokens (strided sample + per-image moments), stage-I gallery/query embeddings,
      stage-I top-K, z_t, stage-II 1536-d features and scores, sorted labels, recalls.
  * ``stage2_L32.npz``      -- reference-style init, Q=1, L=32 full mask, K=3 (BASELINE shape).
  * ``training_forward.npz`` -- the in-batch B x B forward of both stages (train=True paths), B=3.

The few lines of ``validate.py`` / ``validate_stage2.py`` that need datasets are restated
inline with their file:line (they are index bookkeeping around the model calls).
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cir_b200 as cir  # noqa: E402
from ref_shim import build_models  # noqa: E402

syn = cir.synthetic
torch.manual_seed(0)
torch.set_num_threads(os.cpu_count())


def run(name, *, seed, style, G, Q, K, L, min_len, head_gain):
    t0 = time.time()
    sd1 = syn.make_stage1_state_dict(seed, 384, style)
    sd2 = syn.make_stage2_state_dict(seed, 384, style, head_gain=head_gain)
    m1, m2, tok, _ = build_models(sd1, sd2)
    images = syn.make_images(G, 384, seed=1)
    ref_idx, target_idx, ids, mask = syn.make_queries(Q, G, L, seed=3, min_len=min_len)
    feats_in = []
    m2.cls_head.register_forward_hook(lambda mod, inp, out: feats_in.append(inp[0].detach().clone()))
    with torch.no_grad():
        # utils.extract_index_features: src/utils.py:43-55 (stage II) and :56-70 (stage I)
        tokens2 = m2.img_embed(images)
        tokens1, g_emb = m1.img_embed(images, return_pool_and_normalized=True)
        # validate.generate_cirr_val_predictions: src/validate.py:305-311 (normalised twice)
        tok.push(ids, mask)
        q_emb = m1.img_txt_fusion(tokens1[ref_idx], None, ["x"] * Q, train=False)
        q_emb = F.normalize(q_emb)
        # validate.compute_cirr_val_metrics: src/validate.py:202-210
        distances = 1 - q_emb @ g_emb.float().T
        sorted_indices = torch.sort(distances, dim=-1, stable=True).indices
        keep = sorted_indices != ref_idx[:, None]
        sorted_noref = sorted_indices[keep].reshape(Q, G - 1)
        cand_idx = sorted_noref[:, :K]
        k_labels = (cand_idx == target_idx[:, None]).numpy()
        # validate_stage2.generate_cirr_val_predictions: src/validate_stage2.py:235-258
        z_all, scores = [], []
        for q in range(Q):
            tok.push(ids[q:q + 1], mask[q:q + 1])
            z = m1.img_txt_fusion(tokens2[ref_idx[q]][None], None, ["x"], train=False, return_raw=True)
            z_all.append(z.last_hidden_state[0])
            tok.push(ids[q:q + 1], mask[q:q + 1])
            scores.append(m2.img_txt_fusion_val(z, tokens2[cand_idx[q]], ["x"]))
        scores = torch.stack(scores)
        feats = torch.stack(feats_in)
        # validate_stage2.compute_cirr_val_metrics: src/validate_stage2.py:174-179,196-199
        order = torch.sort(scores, dim=-1, descending=True, stable=True).indices
        labels = np.take_along_axis(k_labels, order.numpy(), axis=1)
        lab_t = torch.tensor(labels)
        recalls = [(torch.sum(lab_t[:, :k]) / len(lab_t)).item() * 100 for k in (1, 2, 3, 4)]
    out = dict(
        seed=seed, style=style, G=G, Q=Q, K=K, L=L, min_len=-1 if min_len is None else min_len,
        head_gain=head_gain,
        ref_idx=ref_idx.numpy(), target_idx=target_idx.numpy(), ids=ids.numpy(), mask=mask.numpy(),
        tokens2_sample=tokens2[:, ::48, ::16].numpy(), tokens2_mean=tokens2.mean((1, 2)).numpy(),
        tokens2_std=tokens2.std((1, 2)).numpy(), tokens2_cls=tokens2[:, 0, :].numpy(),
        tokens1_cls=tokens1[:, 0, :].numpy(),
        g_emb=g_emb.numpy(), q_emb=q_emb.numpy(), distances_topk=torch.gather(distances, 1, cand_idx).numpy(),
        cand_idx=cand_idx.numpy().astype(np.int32), k_labels=k_labels,
        z_t=torch.stack(z_all).numpy(), feats=feats.numpy(), scores=scores.numpy(),
        order=order.numpy().astype(np.int32), sorted_labels=labels, recalls=np.array(recalls),
    )
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(f"{name}: {time.time() - t0:.1f}s  scores={scores.numpy().round(4).tolist()}")


def run_training_forward(name, *, seed, style, G, B, L, min_len):
    """In-batch (B x B) forward of both stages: BLIP_Retrieval.img_txt_fusion(train=True) (src/blip_stage1.py:67-92,
    as called by src/stage1_train.py) and BLIP_NLVR.img_txt_fusion (src/blip_stage2.py:65-99, as called by
    src/stage2_train.py:207,468).  Same seeded weights / images / queries as ``pipeline_small.npz``."""
    t0 = time.time()
    sd1 = syn.make_stage1_state_dict(seed, 384, style)
    sd2 = syn.make_stage2_state_dict(seed, 384, style, head_gain=1.0)
    m1, m2, tok, _ = build_models(sd1, sd2)
    images = syn.make_images(G, 384, seed=1)
    ref_idx, target_idx, ids, mask = syn.make_queries(B, G, L, seed=3, min_len=min_len)
    with torch.no_grad():
        tokens2 = m2.img_embed(images)
        tokens1, g_emb = m1.img_embed(images, return_pool_and_normalized=True)
        tok.push(ids, mask)
        s1_logits = m1.img_txt_fusion(tokens1[ref_idx], g_emb[target_idx], ["x"] * B, train=True)          # [B,B] / temp
        tok.push(ids, mask)
        z = m1.img_txt_fusion(tokens2[ref_idx], None, ["x"] * B, train=False, return_raw=True)           # batched z_t
        tok.push(ids, mask)
        s2_logits = m2.img_txt_fusion(z, tokens2[target_idx], ["x"] * B, train=True)                     # [B,B]
    np.savez_compressed(os.path.join(HERE, name), seed=seed, style=style, G=G, B=B, L=L, min_len=min_len,
                        ref_idx=ref_idx.numpy(), target_idx=target_idx.numpy(), ids=ids.numpy(), mask=mask.numpy(),
                        temp=float(m1.temp), s1_logits=s1_logits.numpy(), s2_logits=s2_logits.numpy(),
                        z_t=z.last_hidden_state.numpy())
    print(f"{name}: {time.time() - t0:.1f}s  s2_logits={s2_logits.numpy().round(4).tolist()}")


if __name__ == "__main__":
    run("pipeline_small.npz", seed=0, style="dense", G=6, Q=3, K=4, L=12, min_len=8, head_gain=1.0)
    run("stage2_L32.npz", seed=1, style="reference", G=4, Q=1, K=3, L=32, min_len=None, head_gain=1.0)
    run_training_forward("training_forward.npz", seed=0, style="dense", G=6, B=3, L=12, min_len=8)
