"""Boundary helpers shared by the two model wrappers (reference: src/blip.py:186-209)."""
from __future__ import annotations

from types import SimpleNamespace

from .synthetic import ENC_TOKEN_ID, SyntheticTokenizer, TokenBatch


def init_tokenizer():
    """src/blip.py:186-191: bert-base-uncased + '[DEC]' (bos) + '[ENC]', ``enc_token_id`` = id of
    '[ENC]' (30523).  The vocabulary is only used if it is already on disk (no network); otherwise
    the deterministic offline ``SyntheticTokenizer`` with the same call surface is returned.  Assign
    ``model.tokenizer = <real BertTokenizer>`` to override."""
    try:
        from transformers import BertTokenizer
        tok = BertTokenizer.from_pretrained("bert-base-uncased", local_files_only=True)
        tok.add_special_tokens({"bos_token": "[DEC]"})
        tok.add_special_tokens({"additional_special_tokens": ["[ENC]"]})
        tok.enc_token_id = tok.additional_special_tokens_ids[0]
        return tok
    except Exception:
        return SyntheticTokenizer()


def tokenize(tokenizer, text, device):
    """``text`` is a list of strings (reference behaviour: src/blip_stage2.py:113-114,
    src/blip_stage1.py:72-73) or an already tokenised batch exposing ``input_ids`` /
    ``attention_mask``.  Returns (ids, mask) int32 device tensors with ids[:,0] = enc_token_id."""
    if hasattr(text, "input_ids"):
        enc = text
    else:
        enc = tokenizer(text, padding="longest", return_tensors="pt")
    ids = enc.input_ids.to(device).clone()
    mask = enc.attention_mask.to(device)
    ids[:, 0] = tokenizer.enc_token_id
    return ids.int().contiguous(), mask.int().contiguous()


class EncoderOutput(SimpleNamespace):
    """Stands in for HF ``BaseModelOutputWithPoolingAndCrossAttentions`` (src/med.py:814-821): the
    callers only read ``.last_hidden_state`` (src/blip_stage2.py:66,106)."""


VIT_WIDTH = {"base": 768}


def check_vit(vit: str):
    assert vit in ("base", "large"), "vit parameter must be base or large"      # src/blip.py:196
    if vit != "base":
        raise NotImplementedError("only ViT-B/16 (vit='base') is built: every reference script and config "
                                  "uses it (configs/nlvr.yaml, configs/retrieval_coco.yaml)")
    return 768
