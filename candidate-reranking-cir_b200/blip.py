"""Boundary helpers shared by the two model wrappers (reference: src/blip.py:186-209)."""
from __future__ import annotations

import os
from types import SimpleNamespace

from .synthetic import ENC_TOKEN_ID, SyntheticTokenizer, TokenBatch


class MissingTokenizer:
    """Placeholder used when the bert-base-uncased vocabulary is not on disk: pre-tokenised batches (objects with
    ``input_ids`` / ``attention_mask``) still work, raw strings raise instead of silently getting non-WordPiece ids."""
    enc_token_id = ENC_TOKEN_ID

    def __init__(self, why: str):
        self.why = why

    def __call__(self, *a, **k):
        raise CirTokenizerError(
            "the BERT WordPiece vocabulary (bert-base-uncased) is not available offline (" + self.why + "): pass "
            "pre-tokenised batches, assign `model.tokenizer = <BertTokenizer>`, or opt in to the deterministic hash "
            "tokenizer with synthetic_tokenizer=True / CIR_SYNTHETIC_TOKENIZER=1 (synthetic weights only)")


class CirTokenizerError(RuntimeError):
    pass


def init_tokenizer(synthetic: bool = False):
    """src/blip.py:186-191: bert-base-uncased + '[DEC]' (bos) + '[ENC]', ``enc_token_id`` = id of '[ENC]' (30523).
    The vocabulary is only used if it is already on disk (no network).  When it is not, the result depends on an
    explicit opt-in: ``synthetic=True`` (or CIR_SYNTHETIC_TOKENIZER=1) returns the deterministic offline
    ``SyntheticTokenizer`` -- meaningful with synthetic weights only; otherwise a ``MissingTokenizer`` that accepts
    pre-tokenised batches and raises on raw strings (a real checkpoint must never see hash ids)."""
    try:
        from transformers import BertTokenizer
        tok = BertTokenizer.from_pretrained("bert-base-uncased", local_files_only=True)
        tok.add_special_tokens({"bos_token": "[DEC]"})
        tok.add_special_tokens({"additional_special_tokens": ["[ENC]"]})
        tok.enc_token_id = tok.additional_special_tokens_ids[0]
        return tok
    except Exception as ex:                                  # vocabulary not cached locally
        if synthetic or os.environ.get("CIR_SYNTHETIC_TOKENIZER") == "1":
            return SyntheticTokenizer()
        return MissingTokenizer(type(ex).__name__)


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
