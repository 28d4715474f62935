"""2-GPU NCCL run of the sharded paths: N-GPU result == 1-GPU result bit for bit on (score, idx).
Skipped on boxes with fewer than 2 GPUs (exercised with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cir_b200 as cir

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        syn, D = cir.synthetic, cir.distributed
        dev = torch.device("cuda", rank)
        sd1 = syn.make_stage1_state_dict(0, 384, "dense")
        sd2 = syn.make_stage2_state_dict(0, 384, "dense")
        m1 = cir.blip_stage1.blip_stage1(image_size=384, state_dict=sd1, precision="bf16", device=dev)
        m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="bf16", device=dev)
        eng = m2.engine
        G, Q, K, L = 9, 7, 4, 12
        tokens = m2.img_embed(syn.make_images(G, 384, seed=1))
        ref, tgt, ids, mask = syn.make_queries(Q, G, L, seed=3, min_len=8)
        cand, labels = syn.make_random_topk(Q, G, K, ref, tgt, seed=4)
        ref_d, ids_d, mask_d = ref.int().to(dev), ids.int().to(dev), mask.int().to(dev)
        # stage II: sharded == unsharded
        full = D.stage2_scores_gpu(m1, m2, tokens, ref_d, ids_d, mask_d, cand)
        z_t, _ = m1.encode_queries(tokens, ref_d, ids_d, mask_d, want_z=True, want_emb=False)
        single = m2.score_triplets(z_t, ids_d, mask_d, tokens, cand.numpy())
        assert torch.equal(full, single), (full - single).abs().max()
        # stage I: gallery sharded + merge == unsharded
        g = torch.Generator().manual_seed(7)
        q_emb = torch.nn.functional.normalize(torch.randn(33, 256, generator=g), dim=-1).to(dev)
        g_emb = torch.nn.functional.normalize(torch.randn(5001, 256, generator=g), dim=-1).to(dev)
        excl = torch.randint(0, 5001, (33,), generator=g)
        md, mi = D.stage1_topk_gpu(eng, q_emb, g_emb, 50, exclude=excl)
        sd_, si_ = eng.stage1_topk(q_emb, g_emb, 50, exclude=excl)
        assert torch.equal(mi, si_) and torch.equal(md, sd_)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        ret[rank] = f"{type(e).__name__}: {e}"
        raise
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharding_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
