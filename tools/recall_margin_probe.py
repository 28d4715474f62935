"""How large must planted score margins be for Recall@K parity in bf16?  `cpu` mode: fp32 oracle scores of a config-1
shaped job (8 x 50, L=32, ViT tokens) for several cross_gain values -> tools/_tmp/margin_*.npz; `gpu` mode: the same
jobs through the CUDA bf16 path, printing score spread, top gaps and the in-row ranking error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cir_b200 as cir
syn = cir.synthetic
TMP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_tmp")
G, Q, K, L, SEED = 56, 8, 50, 32, 2


def job(xg):
    sd1 = syn.make_stage1_state_dict(SEED, 384, "reference")
    sd2 = syn.make_stage2_state_dict(SEED, 384, "reference", head_gain=1.0, cross_gain=xg)
    images = syn.make_diverse_images(G, 384, seed=1)
    ref_idx, _, ids, mask = syn.make_queries(Q, G, L, seed=3, min_len=None)
    g = torch.Generator().manual_seed(9)
    cand = torch.stack([(lambda p: p[p != ref_idx[q]][:K])(torch.randperm(G, generator=g)) for q in range(Q)])
    return sd1, sd2, images, ref_idx, ids, mask, cand


if sys.argv[1] == "cpu":
    from oracle import cir_oracle as O
    os.makedirs(TMP, exist_ok=True)
    torch.set_num_threads(8)
    for xg in [float(x) for x in sys.argv[2:]]:
        sd1, sd2, images, ref_idx, ids, mask, cand = job(xg)
        with torch.no_grad():
            tok = O.vit_forward(sd2, images)
            sc = O.stage2_predictions(sd1, sd2, tok, ref_idx, ids, mask, cand)
        np.savez(os.path.join(TMP, f"margin_{xg:g}.npz"), scores=sc.numpy())
        print(xg, "std", float(sc.std()), flush=True)
else:
    for f in sorted(os.listdir(TMP)):
        xg = float(f[len("margin_"):-4])
        want = np.load(os.path.join(TMP, f))["scores"]
        sd1, sd2, images, ref_idx, ids, mask, cand = job(xg)
        m1 = cir.blip_stage1.blip_stage1(image_size=384, state_dict=sd1, precision="bf16")
        m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="bf16")
        tokens = m2.img_embed(images)
        z, _ = m1.encode_queries(tokens, ref_idx.int(), ids, mask, want_z=True, want_emb=False)
        got = m2.score_triplets(z, ids, mask, tokens, cand.int().numpy()).cpu().numpy()
        d = got - want
        ss = -np.sort(-want, 1)
        print(f"cross_gain={xg:g}: score std {want.std():.4f} max|d|={np.abs(d).max():.2e} in-row ranking err={np.abs(d - d.mean(1, keepdims=True)).max():.2e} "
              f"s0-s1 {np.round(ss[:, 0] - ss[:, 1], 3).tolist()} s4-s10 {np.round(ss[:, 4] - ss[:, 10], 3).tolist()} "
              f"spearman-ish: order equal rows {(np.argsort(-got, 1) == np.argsort(-want, 1)).all(1).sum()}/{Q}", flush=True)
