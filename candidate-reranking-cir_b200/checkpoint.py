"""Checkpoint ingestion: turn the files the reference reads into the ``state_dict`` the model
wrappers bind.  Host-side, one-off; nothing here is on the hot path.

  * fine-tuned checkpoints written by ``save_model`` (src/utils.py:135-150):
      {'epoch', '<ClassName>': state_dict, 'optimizer_state_dict'}  -> key 'BLIP_Retrieval' / 'BLIP_NLVR'
      (read at src/validate_stage2.py:347-348,359-360)
  * BLIP base checkpoints {'model': state_dict}: stage II duplicates every (cross)attention ``self`` /
    ``output.dense`` / ``output.LayerNorm`` tensor into the twin keys (src/blip_stage2.py:160-187) and both
    stages interpolate ``visual_encoder.pos_embed`` to the target image size (src/vit.py:281-305).
"""
from __future__ import annotations

from typing import Dict

import torch


def interpolate_pos_embed(pos_embed: torch.Tensor, num_patches: int, num_extra_tokens: int = 1) -> torch.Tensor:
    """src/vit.py:281-305: bicubic resize of the patch position grid, class token kept."""
    emb = pos_embed.shape[-1]
    orig = int((pos_embed.shape[-2] - num_extra_tokens) ** 0.5)
    new = int(num_patches ** 0.5)
    if orig == new:
        return pos_embed
    extra = pos_embed[:, :num_extra_tokens]
    tok = pos_embed[:, num_extra_tokens:].reshape(-1, orig, orig, emb).permute(0, 3, 1, 2)
    tok = torch.nn.functional.interpolate(tok, size=(new, new), mode="bicubic", align_corners=False)
    return torch.cat((extra, tok.permute(0, 2, 3, 1).flatten(1, 2)), dim=1)


def twin_stream_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """src/blip_stage2.py:160-187: self -> self0/self1, dense -> dense0/dense1, LayerNorm -> LayerNormA/B for
    every attention / crossattention block (the originals are kept, as the reference does)."""
    out = dict(sd)
    for key in list(sd.keys()):
        if "crossattention.self." in key or "attention.self." in key:
            out[key.replace("self", "self0")] = sd[key]
            out[key.replace("self", "self1")] = sd[key]
        elif "crossattention.output.dense." in key or "attention.output.dense." in key:
            out[key.replace("dense", "dense0")] = sd[key]
            out[key.replace("dense", "dense1")] = sd[key]
        if "output.LayerNorm" in key and "attention" in key:
            out[key.replace("LayerNorm", "LayerNormA")] = sd[key]
            out[key.replace("LayerNorm", "LayerNormB")] = sd[key]
    return out


def load_state_dict(path: str, kind: str, image_size: int = 384) -> Dict[str, torch.Tensor]:
    """kind: 'BLIP_Retrieval' (stage I) or 'BLIP_NLVR' (stage II)."""
    assert kind in ("BLIP_Retrieval", "BLIP_NLVR")
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if kind in ckpt:
        sd = ckpt[kind]
    else:
        sd = ckpt.get("model", ckpt)
        if kind == "BLIP_NLVR" and not any("self0" in k for k in sd):
            sd = twin_stream_keys(sd)
    n_patches = (image_size // 16) ** 2
    key = "visual_encoder.pos_embed"
    if key in sd and sd[key].shape[-2] != n_patches + 1:
        sd = dict(sd)
        sd[key] = interpolate_pos_embed(sd[key], n_patches)
    return sd
