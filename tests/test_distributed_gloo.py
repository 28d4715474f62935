"""world_size-2 gloo runs (CPU) of the multi-GPU host logic: sharding, padded all-gather, top-K merge.
The per-rank compute is the CPU oracle here; on GPUs the same functions are driven by the engine
(tests/test_gpu_e_multi.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cir_b200 as cir
from oracle import cir_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_cpu(ds, is_):
    """Reference merge: ascending distance, ties -> lowest global index (what cir_topk_merge does)."""
    P, Q, K = ds.shape
    d = ds.permute(1, 0, 2).reshape(Q, P * K)
    i = is_.permute(1, 0, 2).reshape(Q, P * K).long()
    key = torch.stack([torch.tensor(sorted(range(P * K), key=lambda c: (float(d[q, c]), int(i[q, c])))[:K]) for q in range(Q)])
    return torch.gather(d, 1, key), torch.gather(i, 1, key).int()


def _worker(rank, ws, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        D = cir.distributed
        g = torch.Generator().manual_seed(0)
        # ---- stage I: gallery sharded (uneven: 101 rows over 2 ranks), exclusion, merge
        Q, G, K = 7, 101, 10
        q = torch.nn.functional.normalize(torch.randn(Q, 256, generator=g), dim=-1)
        gal = torch.nn.functional.normalize(torch.randn(G, 256, generator=g), dim=-1)
        gal[60] = gal[3]                                       # tie across shards -> lowest global index wins
        excl = torch.randint(0, G, (Q,), generator=g)

        def local_topk(rows):
            d, i = O.stage1_topk(q, gal[rows], None, min(K + 1, rows.stop - rows.start))
            gi = i + rows.start
            out_d = torch.full((Q, K), float("inf")); out_i = torch.full((Q, K), -1, dtype=torch.int32)
            for r in range(Q):
                keep = gi[r] != excl[r]
                out_d[r, : min(K, int(keep.sum()))] = d[r][keep][:K]
                out_i[r, : min(K, int(keep.sum()))] = gi[r][keep][:K].int()
            return out_d, out_i
        md, mi = D.sharded_stage1_topk(local_topk, _merge_cpu, G)
        wd, wi = O.stage1_topk(q, gal, excl, K)
        assert torch.equal(mi.long(), wi) and torch.equal(md, wd)
        # ---- stage II: queries sharded (uneven: 5 rows over 2 ranks), padded all-gather
        Qs, Ks = 5, 6
        full = torch.randn(Qs, Ks, generator=g)
        got = D.sharded_stage2_scores(lambda rows: full[rows].clone(), Qs)
        assert torch.equal(got, full)
        order = O.rerank_order(got)
        assert torch.equal(order, O.rerank_order(full))
        # empty shard: 1 query over 2 ranks
        one = D.sharded_stage2_scores(lambda rows: full[:1][rows].clone(), 1)
        assert torch.equal(one, full[:1])
        # ---- stage II, candidate-range partition: each rank scores the triplets of its candidate range, ranks exchange
        #      (flat position, score) pairs, inactive rows keep the fill value
        S = cir.schedule
        Qc, Kc, Gc = 23, 7, 11
        rng = np.random.default_rng(5)
        cand = np.stack([rng.permutation(Gc)[:Kc] for _ in range(Qc)]).astype(np.int32)
        active = rng.random(Qc) < 0.8
        truth = torch.randn(Qc, Kc, generator=g)
        seen_cands = {}

        def pairs(part):
            assert part == (rank, ws)
            info = {}
            chunks = S.plan_chunks(cand, active, 16, 3, part=part, info=info)
            pos = np.concatenate([c.flat_pos for c in chunks]) if chunks else np.zeros(0, np.int64)
            seen_cands[rank] = set(np.concatenate([c.cand_list for c in chunks]).tolist()) if chunks else set()
            return torch.from_numpy(pos), truth.reshape(-1)[torch.from_numpy(pos)], info["part_sizes"]
        full2 = D.sharded_stage2_scores_by_candidate(pairs, Qc, Kc, -99999.99)
        want = truth.clone()
        want[~torch.from_numpy(active)] = -99999.99
        assert torch.equal(full2, want)
        # no gallery image is owned by two ranks
        mine = torch.zeros(Gc, dtype=torch.int64)
        mine[list(seen_cands[rank])] = 1
        both = [torch.zeros_like(mine) for _ in range(ws)]
        dist.all_gather(both, mine)
        assert int((both[0] * both[1]).sum()) == 0 and int((both[0] + both[1]).sum()) == len(set(cand[active].reshape(-1).tolist()))
        # ---- the name -> gallery-row join of the stage-II drivers, split over the ranks (uneven: 9 query rows over 2 ranks)
        V2 = cir.validate_stage2
        names = [f"img{i:04d}" for i in range(Gc)]
        mat = np.array(names)[np.stack([rng.permutation(Gc)[:Kc] for _ in range(9)])]

        class _Eng:
            device = torch.device("cpu")
        n2i = {n: i for i, n in enumerate(names)}
        got_rows = V2._names_to_rows_sharded(_Eng(), names, n2i, mat)
        assert np.array_equal(got_rows, V2._names_to_rows(names, n2i, mat)) and got_rows.dtype == np.int32
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        ret[rank] = f"{type(e).__name__}: {e}"
        raise
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_single_process_passthrough():
    D = cir.distributed
    x = torch.arange(12.0).view(4, 3)
    assert torch.equal(D.sharded_stage2_scores(lambda rows: x[rows], 4), x)
    d, i = D.sharded_stage1_topk(lambda rows: (x[:, :2], x[:, :2].int()), None, 10)
    assert torch.equal(d, x[:, :2])
