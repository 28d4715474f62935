"""Seeded synthetic weights and inputs for the stage-I / stage-II hot path.

There are no datasets or checkpoints offline, so parity and benchmarks run on
random-init weights in the reference's own ``state_dict`` layout (key names and
shapes as produced by ``blip_stage1`` / ``blip_stage2`` in the reference:
``src/blip_stage1.py:15-45``, ``src/blip_stage2.py:19-54``, ``src/vit.py:113-161``,
``src/med.py:68-86``, ``src/nlvr_encoder.py:273-290,400-412``).

Two init styles:
  * ``"reference"`` -- what the reference constructors produce: Linear/Embedding
    N(0, 0.02), zero biases, LayerNorm 1/0 (``src/nlvr_encoder.py:663-673``,
    ``src/vit.py:163-174``).
  * ``"dense"`` (default) -- same scales but non-zero biases and non-trivial
    LayerNorm gains, so a parity test also exercises every bias / gain path.

Everything is generated on the CPU with an explicit ``torch.Generator`` so the
GPU box (same image, same torch) reproduces the tensors bit for bit.
"""
from __future__ import annotations

import hashlib
import re
from dataclasses import dataclass
from typing import Dict, List, Sequence

import torch

VOCAB_SIZE = 30524          # configs/med_config.json:18
ENC_TOKEN_ID = 30523        # '[ENC]' = 30522 base + '[DEC]' + '[ENC]' (src/blip.py:186-191)
MAX_POS = 512               # configs/med_config.json:12
HIDDEN = 768
HEADS = 12
FFN = 3072
LAYERS = 12
EMBED_DIM = 256             # src/blip_stage1.py:22
VIT_DEPTH = 12
PATCH = 16


def num_tokens(image_size: int) -> int:
    return (image_size // PATCH) ** 2 + 1


class _Init:
    def __init__(self, seed: int, style: str):
        assert style in ("reference", "dense")
        self.g = torch.Generator().manual_seed(seed)
        self.dense = style == "dense"

    def normal(self, *shape, std=0.02):
        return torch.empty(*shape, dtype=torch.float32).normal_(0.0, std, generator=self.g)

    def linear(self, sd: Dict[str, torch.Tensor], name: str, out_f: int, in_f: int, std=0.02):
        sd[name + ".weight"] = self.normal(out_f, in_f, std=std)
        sd[name + ".bias"] = self.normal(out_f, std=0.02) if self.dense else torch.zeros(out_f)

    def layernorm(self, sd: Dict[str, torch.Tensor], name: str, dim: int):
        if self.dense:
            sd[name + ".weight"] = 1.0 + self.normal(dim, std=0.05)
            sd[name + ".bias"] = self.normal(dim, std=0.02)
        else:
            sd[name + ".weight"] = torch.ones(dim)
            sd[name + ".bias"] = torch.zeros(dim)


def _vit(sd: Dict[str, torch.Tensor], ini: _Init, image_size: int, prefix="visual_encoder."):
    n = num_tokens(image_size)
    sd[prefix + "cls_token"] = ini.normal(1, 1, HIDDEN)
    sd[prefix + "pos_embed"] = ini.normal(1, n, HIDDEN)
    sd[prefix + "patch_embed.proj.weight"] = ini.normal(HIDDEN, 3, PATCH, PATCH)
    sd[prefix + "patch_embed.proj.bias"] = ini.normal(HIDDEN) if ini.dense else torch.zeros(HIDDEN)
    for i in range(VIT_DEPTH):
        b = f"{prefix}blocks.{i}."
        ini.layernorm(sd, b + "norm1", HIDDEN)
        ini.linear(sd, b + "attn.qkv", 3 * HIDDEN, HIDDEN)
        ini.linear(sd, b + "attn.proj", HIDDEN, HIDDEN)
        ini.layernorm(sd, b + "norm2", HIDDEN)
        ini.linear(sd, b + "mlp.fc1", FFN, HIDDEN)
        ini.linear(sd, b + "mlp.fc2", HIDDEN, FFN)
    ini.layernorm(sd, prefix + "norm", HIDDEN)


def _embeddings(sd: Dict[str, torch.Tensor], ini: _Init, prefix="text_encoder.embeddings."):
    sd[prefix + "position_ids"] = torch.arange(MAX_POS).expand((1, -1)).clone()
    sd[prefix + "word_embeddings.weight"] = ini.normal(VOCAB_SIZE, HIDDEN)
    sd[prefix + "position_embeddings.weight"] = ini.normal(MAX_POS, HIDDEN)
    ini.layernorm(sd, prefix + "LayerNorm", HIDDEN)


def make_stage1_state_dict(seed: int = 0, image_size: int = 384, style: str = "dense") -> Dict[str, torch.Tensor]:
    """``BLIP_Retrieval.state_dict()`` look-alike (src/blip_stage1.py:15-45)."""
    ini = _Init(seed * 2 + 11, style)
    sd: Dict[str, torch.Tensor] = {}
    sd["temp"] = torch.tensor(0.07)
    _vit(sd, ini, image_size)
    _embeddings(sd, ini)
    for i in range(LAYERS):
        p = f"text_encoder.encoder.layer.{i}."
        for att in ("attention", "crossattention"):
            for nm in ("query", "key", "value"):
                ini.linear(sd, f"{p}{att}.self.{nm}", HIDDEN, HIDDEN)
            ini.linear(sd, f"{p}{att}.output.dense", HIDDEN, HIDDEN)
            ini.layernorm(sd, f"{p}{att}.output.LayerNorm", HIDDEN)
        ini.linear(sd, p + "intermediate.dense", FFN, HIDDEN)
        ini.linear(sd, p + "output.dense", HIDDEN, FFN)
        ini.layernorm(sd, p + "output.LayerNorm", HIDDEN)
    ini.linear(sd, "vision_proj", EMBED_DIM, HIDDEN)
    ini.linear(sd, "text_proj", EMBED_DIM, HIDDEN)
    return sd


def make_stage2_state_dict(seed: int = 0, image_size: int = 384, style: str = "dense",
                           head_gain: float = 1.0, cross_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """``BLIP_NLVR.state_dict()`` look-alike (src/blip_stage2.py:19-54, twin keys
    as in src/nlvr_encoder.py:225-290). ``head_gain`` scales ``cls_head`` weights so
    random-init scores are not nearly tied.  ``cross_gain`` scales the cross-attention output
    projections (dense0 / dense1 weights, applied AFTER all random draws so every other tensor is unchanged): with the
    reference's N(0, 0.02) init the candidate image contributes only a few per cent of the residual stream and the
    scores of different candidates differ by less than a bf16 implementation's rounding noise; a gain of a few units
    makes rankings (and Recall@K) a meaningful parity check."""
    ini = _Init(seed * 2 + 12, style)
    sd: Dict[str, torch.Tensor] = {}
    _vit(sd, ini, image_size)
    _embeddings(sd, ini)
    for i in range(LAYERS):
        p = f"text_encoder.encoder.layer.{i}."
        for att in ("attention", "crossattention"):
            for s in (0, 1):
                for nm in ("query", "key", "value"):
                    ini.linear(sd, f"{p}{att}.self{s}.{nm}", HIDDEN, HIDDEN)
            ini.layernorm(sd, f"{p}{att}.output.LayerNormA", HIDDEN)
            ini.layernorm(sd, f"{p}{att}.output.LayerNormB", HIDDEN)
            ini.linear(sd, f"{p}{att}.output.dense0", HIDDEN, HIDDEN)
            ini.linear(sd, f"{p}{att}.output.dense1", HIDDEN, HIDDEN)
            if att == "crossattention" and i >= 6:          # src/nlvr_encoder.py:286
                ini.linear(sd, f"{p}{att}.output.merge_layer", HIDDEN, 2 * HIDDEN)
        ini.linear(sd, p + "intermediate.dense", FFN, HIDDEN)
        ini.linear(sd, p + "output.dense", HIDDEN, FFN)
        ini.layernorm(sd, p + "output.LayerNorm", HIDDEN)
    ini.linear(sd, "cls_head.0", HIDDEN, 2 * HIDDEN, std=0.02 * head_gain)
    ini.linear(sd, "cls_head.2", 2, HIDDEN, std=0.02 * head_gain)
    if cross_gain != 1.0:
        for i in range(LAYERS):
            for s in (0, 1):
                sd[f"text_encoder.encoder.layer.{i}.crossattention.output.dense{s}.weight"] *= cross_gain
    return sd


def make_images(n: int, image_size: int = 384, seed: int = 1) -> torch.Tensor:
    """Post-normalisation-like images (src/data_utils.py:99-100)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, image_size, image_size, generator=g)


def make_diverse_images(n: int, image_size: int = 384, seed: int = 1) -> torch.Tensor:
    """``make_images`` with a per-image contrast (log-uniform in [0.15, 6]) and brightness offset (uniform in [-3, 3]):
    a random-init ViT maps i.i.d. noise images to nearly identical token sets, so without this the candidates of a query are
    indistinguishable to the stage-II scorer (score spread below bf16 rounding noise)."""
    import math
    img = make_images(n, image_size, seed)
    g = torch.Generator().manual_seed(seed + 7700)
    c = torch.exp(torch.linspace(math.log(0.15), math.log(6.0), n))[torch.randperm(n, generator=g)]
    m = torch.linspace(-3.0, 3.0, n)[torch.randperm(n, generator=g)]
    return img * c[:, None, None, None] + m[:, None, None, None]


def make_token_ids(n: int, length: int = 32, seed: int = 2, min_len: int | None = None):
    """Token ids/mask shaped like BertTokenizer output: [CLS] w.. [SEP] [PAD]..; the
    caller overwrites ids[:,0] with ENC_TOKEN_ID as the reference does
    (src/blip_stage2.py:114). ``min_len`` < ``length`` makes ragged rows."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 30522, (n, length), generator=g)
    mask = torch.ones(n, length, dtype=torch.long)
    if min_len is not None and min_len < length:
        lens = torch.randint(min_len, length + 1, (n,), generator=g)
        for i, ln in enumerate(lens.tolist()):
            ids[i, ln:] = 0
            mask[i, ln:] = 0
            ids[i, ln - 1] = 102
    else:
        ids[:, -1] = 102
    ids[:, 0] = 101
    return ids, mask


class TokenBatch(dict):
    """Minimal stand-in for a HF ``BatchEncoding``: ``.input_ids``, ``.attention_mask``, ``.to()``."""

    def to(self, device, **_):
        return TokenBatch({k: v.to(device) for k, v in self.items()})

    __getattr__ = dict.__getitem__


class SyntheticTokenizer:
    """Deterministic offline tokenizer with BertTokenizer's call surface
    (src/blip.py:186-191). Words hash into [1000, 30522); [CLS]=101, [SEP]=102,
    [PAD]=0; ``enc_token_id`` = 30523. Not a WordPiece implementation: the real
    vocabulary is not available offline; inject a real tokenizer via
    ``model.tokenizer = ...`` when it is."""

    enc_token_id = ENC_TOKEN_ID
    pad_token_id = 0

    def __init__(self, max_length: int = 512, fixed_length: int | None = None):
        self.max_length = max_length
        self.fixed_length = fixed_length

    @staticmethod
    def _word_id(w: str) -> int:
        h = int.from_bytes(hashlib.blake2s(w.encode(), digest_size=4).digest(), "little")
        return 1000 + h % (30522 - 1000)

    def encode(self, text: str) -> List[int]:
        words = re.findall(r"[a-z0-9]+|[^\sa-z0-9]", text.lower())
        ids = [101] + [self._word_id(w) for w in words][: self.max_length - 2] + [102]
        return ids

    def __call__(self, text: Sequence[str], padding="longest", return_tensors="pt", **_):
        if isinstance(text, str):
            text = [text]
        rows = [self.encode(t) for t in text]
        L = self.fixed_length or max(len(r) for r in rows)
        ids = torch.zeros(len(rows), L, dtype=torch.long)
        mask = torch.zeros(len(rows), L, dtype=torch.long)
        for i, r in enumerate(rows):
            r = r[:L]
            if len(r) == L:
                r[-1] = 102
            ids[i, : len(r)] = torch.tensor(r)
            mask[i, : len(r)] = 1
        return TokenBatch(input_ids=ids, attention_mask=mask)


@dataclass
class SyntheticRetrievalSet:
    """A dataset-free stand-in for the reference's 'relative' val split + top-K file
    (src/data_utils.py:166-179,290-305): per query a reference image index, a
    caption, a target index and the stage-I top-K candidate list."""
    ref_idx: torch.Tensor        # [Q] int64
    target_idx: torch.Tensor     # [Q] int64
    ids: torch.Tensor            # [Q,L] int64 (ids[:,0] already ENC)
    mask: torch.Tensor           # [Q,L] int64
    cand_idx: torch.Tensor       # [Q,K] int32
    K_labels: torch.Tensor       # [Q,K] bool


def make_queries(num_queries: int, gallery: int, length: int = 32, seed: int = 3, min_len: int | None = None):
    g = torch.Generator().manual_seed(seed)
    ref_idx = torch.randint(0, gallery, (num_queries,), generator=g)
    off = torch.randint(1, gallery, (num_queries,), generator=g)
    target_idx = (ref_idx + off) % gallery          # never the reference itself
    ids, mask = make_token_ids(num_queries, length, seed=seed + 1000, min_len=min_len)
    ids[:, 0] = ENC_TOKEN_ID
    return ref_idx, target_idx, ids, mask


def make_random_topk(num_queries: int, gallery: int, k: int, ref_idx: torch.Tensor, target_idx: torch.Tensor,
                     seed: int = 4, hit_rate: float = 0.98):
    """Random stage-I-like candidate lists: K distinct gallery indices per query,
    never the reference; the target is planted in ~hit_rate of the rows."""
    g = torch.Generator().manual_seed(seed)
    cand = torch.empty(num_queries, k, dtype=torch.int32)
    for q in range(num_queries):
        perm = torch.randperm(gallery, generator=g)
        perm = perm[perm != ref_idx[q]]
        row = perm[:k].clone()
        t = int(target_idx[q])
        has = bool((row == t).any())
        want = bool(torch.rand((), generator=g) < hit_rate)
        if want and not has:
            row[int(torch.randint(0, k, (), generator=g))] = t
        elif not want and has:
            repl = perm[k] if perm.numel() > k else row[0]
            row[row == t] = repl
        cand[q] = row.to(torch.int32)
    labels = cand.to(torch.int64) == target_idx[:, None]
    return cand, labels


class SyntheticRelativeDataset:
    """Dataset-free stand-in for ``CIRRDataset``/``FashionIQDataset`` in 'relative' mode with a loaded
    top-K file (src/data_utils.py:166-179,290-305).  Exposes what the metric functions read:
    ``K``, ``K_labels`` (numpy bool [Q,K]), ``split``/``dress_types`` plus per-query arrays:
    ``reference_names``, ``target_names``, ``captions`` (str or [str,str] for Fashion-IQ),
    ``K_sorted_index_names`` [Q,K], and for CIRR ``group_members`` [Q,6] (reference first)."""

    def __init__(self, index_names, ref_idx, target_idx, captions, cand_idx, kind="cirr", group_idx=None,
                 split="val", dress_types=("dress",), token_batch=None):
        import numpy as np
        self.kind = kind
        self.split = split
        self.dress_types = list(dress_types)
        self.index_names = list(index_names)
        names = np.array(self.index_names)
        self.reference_names = names[np.asarray(ref_idx)].tolist()
        self.target_names = names[np.asarray(target_idx)].tolist()
        self.captions = list(captions)
        cand_idx = np.asarray(cand_idx)
        self.K = cand_idx.shape[1]
        self.K_sorted_index_names = names[cand_idx]
        self.K_labels = self.K_sorted_index_names == np.array(self.target_names)[:, None]
        self.group_members = None if group_idx is None else names[np.asarray(group_idx)]
        self.token_batch = token_batch      # optional pre-tokenised captions (TokenBatch)

    def __len__(self):
        return len(self.reference_names)


def index_names_for(n: int) -> List[str]:
    return [f"img_{i:07d}" for i in range(n)]


def make_group_members(ref_idx: torch.Tensor, target_idx: torch.Tensor, gallery: int, seed: int = 5) -> torch.Tensor:
    """CIRR 'img_set' members: [Q,6] = reference, target and 4 other distinct images."""
    g = torch.Generator().manual_seed(seed)
    Q = ref_idx.numel()
    out = torch.empty(Q, 6, dtype=torch.int64)
    for q in range(Q):
        r, t = int(ref_idx[q]), int(target_idx[q])
        perm = torch.randperm(gallery, generator=g)
        others = [int(x) for x in perm if int(x) not in (r, t)][:4]
        row = [t] + others
        order = torch.randperm(5, generator=g).tolist()
        out[q] = torch.tensor([r] + [row[i] for i in order])
    return out
