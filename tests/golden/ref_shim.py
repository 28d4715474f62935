"""Import the UNMODIFIED reference modules from /root/reference in this container.

Test-harness only (golden generation); never imported by the product or by tests that run on
the GPU box (/root/reference does not exist there).  The reference needs timm 0.4.12,
fairscale and transformers 4.25; none are installed, so the missing names are stubbed with
their pinned-version semantics:

  * timm PatchEmbed = Conv2d(3,768,16,16) + flatten(2).transpose(1,2); DropPath = identity in
    eval; trunc_normal_ = nn.init.trunc_normal_
  * fairscale checkpoint_wrapper = identity
  * transformers.modeling_utils: apply_chunking_to_forward / prune_linear_layer re-exported,
    get_head_mask -> [None]*n, init_weights -> self.apply(self._init_weights) (v4.25)
  * the BertTokenizer vocabulary is not on disk -> a tokenizer object is injected by the caller
"""
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE = os.environ.get("CIR_REFERENCE", "/root/reference")


def load_reference():
    import transformers  # noqa: F401  (must precede the fake timm)
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    def _mod(n):
        m = types.ModuleType(n)
        sys.modules[n] = m
        return m

    _mod("timm"); _mod("timm.models")
    tvt = _mod("timm.models.vision_transformer"); treg = _mod("timm.models.registry")
    tlay = _mod("timm.models.layers"); thelp = _mod("timm.models.helpers"); thub = _mod("timm.models.hub")

    class PatchEmbed(nn.Module):
        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None):
            super().__init__()
            self.img_size = (img_size,) * 2
            self.patch_size = (patch_size,) * 2
            self.num_patches = (img_size // patch_size) ** 2
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
            self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

        def forward(self, x):
            return self.norm(self.proj(x).flatten(2).transpose(1, 2))

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            assert not self.training or self.p == 0.0
            return x

    tvt._cfg = lambda **k: {}
    tvt.PatchEmbed = PatchEmbed
    treg.register_model = lambda f: f
    tlay.trunc_normal_ = nn.init.trunc_normal_
    tlay.DropPath = DropPath
    thelp.named_apply = thelp.adapt_input_conv = thub.download_cached_file = None
    _mod("fairscale"); _mod("fairscale.nn"); _mod("fairscale.nn.checkpoint")
    _mod("fairscale.nn.checkpoint.checkpoint_activations").checkpoint_wrapper = lambda m: m
    mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    mu.prune_linear_layer = pu.prune_linear_layer
    mu.find_pruneable_heads_and_indices = None
    mu.PreTrainedModel.get_head_mask = lambda self, hm, n, *a: [None] * n
    mu.PreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)

    sys.path.insert(0, os.path.join(REFERENCE, "src"))
    cwd = os.getcwd()
    os.chdir(REFERENCE)
    try:
        import blip
        import blip_stage1 as s1
        import blip_stage2 as s2
    finally:
        os.chdir(cwd)
    return blip, s1, s2


class FixedTokenizer:
    """Returns the (ids, mask) rows queued with ``push``; stands in for BertTokenizer."""
    enc_token_id = 30523

    def __init__(self):
        self.next = None

    def push(self, ids, mask):
        self.next = (ids.clone(), mask.clone())

    def __call__(self, text, padding="longest", return_tensors="pt"):
        ids, mask = self.next

        class Enc(dict):
            def to(self, d):
                return self
            __getattr__ = dict.__getitem__
        return Enc(input_ids=ids.clone(), attention_mask=mask.clone())


def build_models(sd1, sd2, image_size=384):
    blip, s1, s2 = load_reference()
    tok = FixedTokenizer()
    blip.init_tokenizer = s1.init_tokenizer = s2.init_tokenizer = lambda: tok
    cfg = os.path.join(REFERENCE, "configs/med_config.json")
    with torch.no_grad():
        m1 = s1.blip_stage1(pretrained="", image_size=image_size, vit="base", med_config=cfg).float().eval()
        m2 = s2.blip_stage2(pretrained="", image_size=image_size, vit="base", med_config=cfg).float().eval()
    msg1 = m1.load_state_dict(sd1, strict=True)
    msg2 = m2.load_state_dict(sd2, strict=True)
    return m1, m2, tok, (msg1, msg2)
