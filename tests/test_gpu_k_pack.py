"""cir_pack_{vit,stage1,stage2}_weights: the C-ABI packers that turn reference state_dict tensors into the packed structs
(stacking, casts, fp64 merge fold of src/nlvr_encoder.py:250-258) against a torch restatement of the same layout."""
import ctypes as C

import numpy as np
import pytest
import torch

import cir_b200 as cir

pytestmark = pytest.mark.gpu
syn = cir.synthetic
N = cir.native


def _view(blob, ptr, shape, dtype):
    """The packed tensor behind a struct pointer, as a view into the blob it was carved from."""
    esz = torch.empty((), dtype=dtype).element_size()
    off = int(ptr) - blob.data_ptr()
    n = int(np.prod(shape))
    assert 0 <= off and off + n * esz <= blob.numel()
    return blob[off:off + n * esz].view(dtype).view(*shape)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_stage2_packer_layout_and_merge_fold(precision):
    eng = cir.engine.get_engine(precision=precision)
    sd = syn.make_stage2_state_dict(3, 384, "dense")
    w, keep = eng.pack_stage2(sd)
    blob, act = keep[0], eng.act_dtype
    pre = "text_encoder.encoder.layer."
    for i in (0, 5, 6, 11):
        a, c = f"{pre}{i}.attention.", f"{pre}{i}.crossattention."
        qkv = torch.cat([sd[a + f"self{s}.{n}.weight"] for s in (0, 1) for n in ("query", "key", "value")]).to(act)
        assert torch.equal(_view(blob, w.self_qkv_w[i], (2, 2304, 768), act).cpu(), qkv.view(2, 2304, 768))
        kvb = torch.cat([sd[c + f"self{s}.{n}.bias"] for s in (0, 1) for n in ("key", "value")])
        assert torch.equal(_view(blob, w.cross_kv_b[i], (3072,), torch.float32).cpu(), kvb)
        lng = torch.cat([sd[c + "output.LayerNormA.weight"], sd[c + "output.LayerNormB.weight"]])
        assert torch.equal(_view(blob, w.cross_ln_g[i], (1536,), torch.float32).cpu(), lng)
        # folded output projection + merge, composed in float64
        W0, W1 = sd[c + "output.dense0.weight"].double(), sd[c + "output.dense1.weight"].double()
        b0, b1 = sd[c + "output.dense0.bias"].double(), sd[c + "output.dense1.bias"].double()
        if i >= 6:
            Wm, bm = sd[c + "output.merge_layer.weight"].double(), sd[c + "output.merge_layer.bias"].double()
            Wc = torch.cat([Wm[:, :768] @ W0, Wm[:, 768:] @ W1], dim=1)
            bc = Wm[:, :768] @ b0 + Wm[:, 768:] @ b1 + bm
        else:
            Wc, bc = 0.5 * torch.cat([W0, W1], dim=1), 0.5 * (b0 + b1)
        got_w = _view(blob, w.cross_out_w[i], (768, 1536), act).cpu()
        got_b = _view(blob, w.cross_out_b[i], (768,), torch.float32).cpu()
        want_w = Wc.float().to(act)
        # the fp64 sums run in another order than torch's dgemm: identical after rounding except (rarely) at a rounding tie
        diff = got_w.float() != want_w.float()
        assert diff.float().mean().item() < 1e-4
        assert (got_w.double() - Wc).abs().max() <= (2.0 ** -8 if precision == "bf16" else 2.0 ** -23) * Wc.abs().max()
        assert (got_b.double() - bc).abs().max() <= 1e-7 * max(1.0, bc.abs().max().item())
        if i < 6:
            assert torch.equal(got_w, want_w) and torch.equal(got_b, bc.float())
    assert torch.equal(_view(blob, w.cls2_w, (768,), torch.float32).cpu(), sd["cls_head.2.weight"][0])
    assert torch.equal(_view(blob, w.cls2_b, (1,), torch.float32).cpu(), sd["cls_head.2.bias"][:1])


def test_stage1_and_vit_packers_layout():
    eng = cir.engine.get_engine(precision="bf16")
    sd = syn.make_stage1_state_dict(4, 384, "dense")
    w, keep = eng.pack_stage1(sd)
    blob = keep[0]
    a = "text_encoder.encoder.layer.3."
    qkv = torch.cat([sd[a + f"attention.self.{n}.weight"] for n in ("query", "key", "value")]).bfloat16()
    assert torch.equal(_view(blob, w.self_qkv_w[3], (2304, 768), torch.bfloat16).cpu(), qkv)
    kv = torch.cat([sd[a + f"crossattention.self.{n}.weight"] for n in ("key", "value")]).bfloat16()
    assert torch.equal(_view(blob, w.cross_kv_w[3], (1536, 768), torch.bfloat16).cpu(), kv)
    assert torch.equal(_view(blob, w.text_proj_b, (256,), torch.float32).cpu(), sd["text_proj.bias"])
    assert torch.equal(_view(blob, w.word_emb, tuple(sd["text_encoder.embeddings.word_embeddings.weight"].shape), torch.float32).cpu(),
                       sd["text_encoder.embeddings.word_embeddings.weight"])
    v, vkeep, n_tok = eng.pack_vit(sd)
    assert n_tok == 577
    assert torch.equal(_view(vkeep[0], v.patch_w, (768, 768), torch.bfloat16).cpu(), sd["visual_encoder.patch_embed.proj.weight"].reshape(768, -1).bfloat16())
    assert torch.equal(_view(vkeep[0], v.pos_embed, (577, 768), torch.float32).cpu(), sd["visual_encoder.pos_embed"].reshape(577, 768))
    assert torch.equal(_view(vkeep[0], v.fc2_w[11], (768, 3072), torch.bfloat16).cpu(), sd["visual_encoder.blocks.11.mlp.fc2.weight"].bfloat16())


def test_packer_reports_a_small_blob_and_missing_tensors():
    eng = cir.engine.get_engine(precision="bf16")
    sd = syn.make_stage2_state_dict(3, 384, "dense")
    hold = []
    st = eng.stage2_state(sd, hold)
    need = eng._lib.cir_pack_stage2_bytes(eng.ctx, st.emb.vocab_rows, st.emb.pos_rows)
    blob = torch.empty(need // 2, dtype=torch.uint8, device=eng.device)
    w = N.Stage2Weights()
    rc = eng._lib.cir_pack_stage2_weights(eng.ctx, C.byref(st), N.ptr(blob), blob.numel(), C.byref(w))
    torch.cuda.synchronize()
    assert rc == -3 and b"too small" in N.lib().cir_last_error()          # CIR_EWORKSPACE
    del sd["text_encoder.encoder.layer.8.crossattention.output.merge_layer.weight"]
    with pytest.raises(N.CirError, match="merge_layer"):
        eng.pack_stage2(sd)
