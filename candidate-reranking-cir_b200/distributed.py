"""Multi-GPU sharding of the hot path: one process per GPU (torchrun), ``torch.distributed`` for the
plumbing.  The path shards without any data-path collective: stage-II triplets are
independent, the stage-I gallery splits by rows; ranks exchange only (score, index) pairs:

  stage II : queries block-partitioned; each rank scores its [Q_r, K] block; one all-gather of the
             padded score blocks -> every rank holds the full [Q, K] matrix and re-sorts it.
  stage I  : gallery rows block-partitioned; each rank computes a local top-K with global column ids
             (``col_offset``); one all-gather of [Q, K] (distance, index) lists; ``cir_topk_merge``.

The compute steps are injected callables so the same code runs under ``gloo`` on CPU in the tests
(with the oracle standing in for the kernels) and under ``nccl`` on GPUs (with the engine).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from .schedule import shard_rows


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_rows(local: torch.Tensor, total_rows: int) -> torch.Tensor:
    """Concatenate per-rank row blocks produced by ``shard_rows`` (block sizes differ by at most one
    row: pad to the largest, gather, strip)."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_rows(total_rows, r, ws) for r in range(ws)]
    mx = max(s.stop - s.start for s in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    return torch.cat([o[: s.stop - s.start] for o, s in zip(out, sizes)], dim=0)


def sharded_stage2_scores(score_fn: Callable[[slice], torch.Tensor], num_queries: int) -> torch.Tensor:
    """``score_fn(rows)`` -> fp32 [rows, K] scores of this rank's query block; returns the full [Q, K]
    matrix on every rank (8 bytes per triplet cross NVLink once sorted indices are added)."""
    rank, ws = world()
    rows = shard_rows(num_queries, rank, ws)
    local = score_fn(rows)
    assert local.shape[0] == rows.stop - rows.start
    return all_gather_rows(local, num_queries)


def sharded_stage1_topk(local_topk_fn: Callable[[slice], Tuple[torch.Tensor, torch.Tensor]],
                        merge_fn: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                        gallery_rows: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """``local_topk_fn(rows)`` -> (dist [Q,K], idx [Q,K] with GLOBAL gallery ids) over this rank's gallery
    shard; ``merge_fn(dist [P,Q,K], idx [P,Q,K])`` -> merged lists.  Every rank gets the merged top-K."""
    rank, ws = world()
    rows = shard_rows(gallery_rows, rank, ws)
    d, i = local_topk_fn(rows)
    if ws == 1:
        return d, i
    ds = [torch.empty_like(d) for _ in range(ws)]
    is_ = [torch.empty_like(i) for _ in range(ws)]
    dist.all_gather(ds, d.contiguous())
    dist.all_gather(is_, i.contiguous())
    return merge_fn(torch.stack(ds), torch.stack(is_))


# ---- engine-backed convenience wrappers (GPU) ---------------------------------------------------------------

def stage2_scores_gpu(m1, m2, gallery_tokens, ref_idx, ids, mask, cand_idx, row_active=None) -> torch.Tensor:
    import numpy as np
    cand_np = cand_idx.cpu().numpy() if isinstance(cand_idx, torch.Tensor) else np.asarray(cand_idx)

    def score(rows: slice):
        z_t, _ = m1.encode_queries(gallery_tokens, ref_idx[rows], ids[rows], mask[rows], want_z=True, want_emb=False)
        act = None if row_active is None else np.asarray(row_active)[rows]
        return m2.score_triplets(z_t, ids[rows], mask[rows], gallery_tokens, cand_np[rows], act)
    return sharded_stage2_scores(score, cand_np.shape[0])


def stage1_topk_gpu(engine, q_emb, g_emb, k: int, exclude=None):
    def local(rows: slice):
        return engine.stage1_topk(q_emb, g_emb[rows], k, exclude=exclude, col_offset=rows.start)
    return sharded_stage1_topk(local, engine.topk_merge, g_emb.shape[0])
