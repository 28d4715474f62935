"""Stage-I model wrapper: same surface as the reference's ``BLIP_Retrieval`` (src/blip_stage1.py:15-101)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from . import native as N
from .blip import EncoderOutput, check_vit, init_tokenizer, tokenize
from .engine import Engine, get_engine


class BLIP_Retrieval(nn.Module):
    def __init__(self, med_config="configs/med_config.json", image_size=384, vit="base", vit_grad_ckpt=False,
                 vit_ckpt_layer=0, embed_dim=256, *, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 precision: str = "bf16", device=None, engine: Optional[Engine] = None,
                 synthetic_tokenizer: bool = False):
        super().__init__()
        check_vit(vit)
        assert embed_dim == 256, "embed_dim is fixed at 256 (src/blip_stage1.py:22)"
        self.image_size = image_size
        self.engine = engine or get_engine(device, precision)
        self.tokenizer = init_tokenizer(synthetic_tokenizer)
        self._from_checkpoint = False
        self.temp = 0.07
        self._vit = self._w = None
        self._keep = []
        if state_dict is not None:
            self.load_state_dict(state_dict)

    def float(self):
        return self

    def load_state_dict(self, state_dict, strict: bool = True):
        e = self.engine
        self._vit, k1, n_tok = e.pack_vit(state_dict)
        assert n_tok == (self.image_size // 16) ** 2 + 1, "pos_embed does not match image_size"
        self._w, k2 = e.pack_stage1(state_dict)
        self._keep = [k1, k2]
        if "temp" in state_dict:
            self.temp = float(state_dict["temp"])
        return self

    def _need_weights(self):
        if self._w is None:
            raise N.CirError("BLIP_Retrieval has no weights: pass state_dict= or call load_state_dict()")

    def img_embed(self, image, atts=False, return_pool_and_normalized=False):
        """src/blip_stage1.py:48-64."""
        self._need_weights()
        image_embeds = self.engine.vit_forward(self._vit, image)
        out = (image_embeds,)
        if return_pool_and_normalized:
            out += (self.engine.stage1_gallery_embed(self._w, image_embeds),)
        if atts:
            out += (torch.ones(image_embeds.size()[:-1], dtype=torch.long, device=image_embeds.device),)
        return out[0] if len(out) == 1 else out

    def img_txt_fusion(self, r_image_embeds, t_image_embeds, text, train=True, return_raw=False):
        """src/blip_stage1.py:67-92.  ``train=False, return_raw=True`` -> object with
        ``.last_hidden_state`` [B,L,768] (z_t); ``train=False`` -> normalised [B,256]; ``train=True`` -> [B,Bt] logits
        ``predicted @ t_image_embeds.T / temp`` (forward only)."""
        self._need_weights()
        e = self.engine
        ref = e.to_act(r_image_embeds)
        B = ref.shape[0]
        ids, mask = tokenize(self.tokenizer, text, e.device)
        assert ids.shape[0] == B
        ar = torch.arange(B, dtype=torch.int32, device=e.device)
        if train:                                                # forward only: [B, Bt] in-batch logits (:88-91)
            _, emb = e.stage1_encode(self._w, ref, ar, ids, mask, want_z=False, want_emb=True)
            return e.stage1_logits(emb, t_image_embeds, self.temp)
        z, emb = e.stage1_encode(self._w, ref, ar, ids, mask, want_z=return_raw, want_emb=not return_raw)
        return EncoderOutput(last_hidden_state=z) if return_raw else emb

    # ---- batched gallery-resident paths used by validate / validate_stage2
    def encode_queries(self, gallery_tokens, ref_index, ids, mask, want_z=True, want_emb=False, normalize_twice=False):
        self._need_weights()
        return self.engine.stage1_encode(self._w, gallery_tokens, ref_index, ids, mask, want_z, want_emb, normalize_twice)

    def forward(self, *a, **k):
        raise NotImplementedError("use img_embed / img_txt_fusion (the reference defines no forward())")


def blip_stage1(pretrained="", **kwargs):
    """src/blip_stage1.py:95-101; checkpoint key 'BLIP_Retrieval' (src/validate_stage2.py:347-348)."""
    model = BLIP_Retrieval(**kwargs)
    if pretrained:
        from .checkpoint import load_state_dict
        from .synthetic import SyntheticTokenizer
        if isinstance(model.tokenizer, SyntheticTokenizer):
            raise N.CirError("a real checkpoint needs the real BERT WordPiece tokenizer: the synthetic hash tokenizer is for "
                             "synthetic weights only (assign model.tokenizer or pass pre-tokenised batches)")
        model.load_state_dict(load_state_dict(pretrained, "BLIP_Retrieval", kwargs.get("image_size", model.image_size)))
        model._from_checkpoint = True
    return model
