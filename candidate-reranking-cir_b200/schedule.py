"""Candidate-major triplet scheduling (pure numpy host logic, no GPU needed).

The reference scores one query at a time (src/validate_stage2.py:94-125,235-275) and therefore
recomputes the cross-attention K/V projections of a gallery image for every top-K list that
names it (src/nlvr_encoder.py:158-159: 69 % of its FLOPs).  Triplets are independent
(src/blip_stage2.py:118-136), so here the Q*K (query, candidate) pairs are sorted by candidate
and cut into chunks; inside a chunk every unique candidate's K/V is computed once.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Optional

import numpy as np


@dataclass
class Chunk:
    flat_pos: np.ndarray     # [T] int64  position q*K+k of each triplet in the [Q,K] score matrix
    cand_list: np.ndarray    # [C] int32  unique gallery rows of the chunk (ascending)
    trip_slot: np.ndarray    # [T] int32  index into cand_list
    query_list: np.ndarray   # [Qc] int32 unique query rows of the chunk (ascending)
    trip_query: np.ndarray   # [T] int32  index into query_list


def _sorted_triplets(cand_idx: np.ndarray, row_active: Optional[np.ndarray]):
    """Active triplets of a [Q,K] candidate matrix sorted by candidate (stable: ties keep q*K+k order).
    -> (flat positions int64, candidate ids int64)."""
    Q, K = cand_idx.shape
    flat = np.arange(Q * K, dtype=np.int64)
    if row_active is not None:
        row_active = np.asarray(row_active, dtype=bool)
        assert row_active.shape == (Q,)
        flat = flat[np.repeat(row_active, K)]
    if flat.size == 0:
        return flat, flat
    cands = cand_idx.reshape(-1)[flat]
    assert cands.min() >= 0, "negative candidate index"
    # numpy's stable sort is a radix sort for 16-bit keys (10x faster than the merge sort it uses for wider ints)
    key = cands.astype(np.uint16) if cands.max() < 65536 else cands.astype(np.int64)
    order = np.argsort(key, kind="stable")
    return flat[order], cands[order].astype(np.int64)


def candidate_partition(cands_sorted: np.ndarray, world: int) -> np.ndarray:
    """Cut a candidate-sorted triplet list into ``world`` contiguous ranges of (nearly) equal triplet count, moving every
    cut to the nearest boundary between two candidates so that no candidate's K/V is computed on two ranks.
    -> int64 [world+1] offsets into the sorted list."""
    n = cands_sorted.size
    cuts = np.zeros(world + 1, np.int64)
    cuts[world] = n
    if n == 0:
        return cuts
    starts = np.flatnonzero(np.r_[True, cands_sorted[1:] != cands_sorted[:-1]])
    bounds = np.r_[starts, n]                                   # every legal cut position
    for r in range(1, world):
        t = (n * r) // world
        j = np.searchsorted(bounds, t)
        lo = bounds[max(j - 1, 0)]
        hi = bounds[min(j, bounds.size - 1)]
        cuts[r] = lo if t - lo <= hi - t else hi
    return np.maximum.accumulate(cuts)


def plan_chunks(cand_idx: np.ndarray, row_active: Optional[np.ndarray] = None, max_triplets: int = 2048,
                max_candidates: int = 64, part: Optional[tuple] = None, balance: bool = True,
                info: Optional[dict] = None) -> List[Chunk]:
    """cand_idx [Q,K] int -> chunks covering every triplet of the active rows exactly once.

    A chunk holds at most ``max_triplets`` triplets and ``max_candidates`` unique candidates; all
    triplets of one candidate are kept together unless a single candidate has more than
    ``max_triplets`` of them (then it is split, recomputing its K/V once per piece).
    ``part=(rank, world)``: only the chunks of this rank's candidate range (``candidate_partition``); the union over the
    ranks covers every triplet exactly once.  ``balance``: chunks of (nearly) equal size instead of greedily full chunks
    followed by a small tail chunk (small GEMMs run below the large-GEMM rate).  ``info``: optional dict that receives
    ``part_sizes`` (triplets per rank)."""
    cand_idx = np.asarray(cand_idx)
    assert cand_idx.ndim == 2
    Q, K = cand_idx.shape
    assert max_triplets >= 1 and max_candidates >= 1
    flat, cands = _sorted_triplets(cand_idx, row_active)
    if part is not None:
        rank, world = part
        assert 0 <= rank < world
        cuts = candidate_partition(cands, world)
        if info is not None:
            info["part_sizes"] = np.diff(cuts).tolist()          # triplets per rank: known everywhere without communication
        flat, cands = flat[cuts[rank]:cuts[rank + 1]], cands[cuts[rank]:cuts[rank + 1]]
    elif info is not None:
        info["part_sizes"] = [int(flat.size)]
    if flat.size == 0:
        return []
    # run boundaries of equal candidates
    starts = np.flatnonzero(np.r_[True, cands[1:] != cands[:-1]])
    ends = np.r_[starts[1:], cands.size]
    run_len = ends - starts
    if balance and run_len.max() <= max_triplets:
        # equal-size chunks: cut at the candidate boundary nearest to every multiple of n / nchunks; take the smallest
        # chunk count whose cuts respect both limits
        bounds = np.r_[starts, cands.size]
        n = cands.size
        nchunks = max(-(-n // max_triplets), -(-starts.size // max_candidates))
        for _ in range(64):
            cuts = candidate_partition(cands, nchunks) if nchunks > 1 else np.array([0, n], np.int64)
            sizes = np.diff(cuts)
            ncand = np.diff(np.searchsorted(bounds, cuts))
            if sizes.max() <= max_triplets and ncand.max() <= max_candidates:
                return [_make_chunk(flat[a:b], cands[a:b], K) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
            nchunks += 1
    chunks: List[Chunk] = []
    cur_lo = 0          # first triplet of the open chunk
    cur_hi = 0
    cur_c = 0
    def close(lo, hi):
        if hi > lo:
            chunks.append(_make_chunk(flat[lo:hi], cands[lo:hi], K))
    for s, e in zip(starts, ends):
        n = e - s
        if n > max_triplets:                      # oversized candidate: flush, then split it alone
            close(cur_lo, cur_hi)
            for p in range(s, e, max_triplets):
                close(p, min(p + max_triplets, e))
            cur_lo = cur_hi = e
            cur_c = 0
            continue
        if (cur_hi - cur_lo) + n > max_triplets or cur_c + 1 > max_candidates:
            close(cur_lo, cur_hi)
            cur_lo = s
            cur_c = 0
        cur_hi = e
        cur_c += 1
    close(cur_lo, cur_hi)
    return chunks


def _make_chunk(flat_pos: np.ndarray, cands: np.ndarray, K: int) -> Chunk:
    cand_list, trip_slot = np.unique(cands, return_inverse=True)
    queries = flat_pos // K
    query_list, trip_query = np.unique(queries, return_inverse=True)
    return Chunk(flat_pos=flat_pos.astype(np.int64), cand_list=cand_list.astype(np.int32),
                 trip_slot=trip_slot.astype(np.int32), query_list=query_list.astype(np.int32),
                 trip_query=trip_query.astype(np.int32))


def build_attn_work(trip_slot: np.ndarray, L: int, warps: int = 8) -> np.ndarray:
    """Cross-attention work list for ``cir_attn_args.work`` (include/cir_b200.h): one CTA of ``warps``
    16-row query tiles per entry, every entry inside one run of triplets that share a candidate.
    ``trip_slot`` must be sorted (candidate-major).  Returns int32 [W,4] = (first triplet of the run,
    first unit, units in the run, 0)."""
    trip_slot = np.asarray(trip_slot)
    if trip_slot.size == 0:
        return np.zeros((0, 4), np.int32)
    assert np.all(np.diff(trip_slot) >= 0), "trip_slot must be candidate-major (sorted)"
    mt = (L + 15) // 16
    starts = np.flatnonzero(np.r_[True, trip_slot[1:] != trip_slot[:-1]])
    counts = np.diff(np.r_[starts, trip_slot.size])
    units = counts * mt
    nctas = (units + warps - 1) // warps
    run_of = np.repeat(np.arange(starts.size), nctas)
    first = np.cumsum(nctas) - nctas
    unit0 = (np.arange(run_of.size) - first[run_of]) * warps
    out = np.zeros((run_of.size, 4), np.int32)
    out[:, 0] = starts[run_of]
    out[:, 1] = unit0
    out[:, 2] = units[run_of]
    return out


TILE_ROWS = 256          # query rows of one attention work item: two 128-row UMMA tiles sharing every K/V chunk


def build_attn_tiles(trip_slot: np.ndarray, L: int) -> np.ndarray:
    """Tile list for the tcgen05 attention kernel (``cir_attn_args.tiles``): int32 [W,4] =
    (first triplet, triplets in the tile, first query row, rows per triplet RB).  A tile is 256 query
    rows: 256/RB consecutive triplets of ONE candidate run, with RB = L rounded up to a power of two, or
    a 256-row slice of a single triplet when L > 256."""
    trip_slot = np.asarray(trip_slot)
    if trip_slot.size == 0:
        return np.zeros((0, 4), np.int32)
    assert np.all(np.diff(trip_slot) >= 0), "trip_slot must be candidate-major (sorted)"
    if L > TILE_ROWS:
        nslices = (L + TILE_ROWS - 1) // TILE_ROWS
        t = np.repeat(np.arange(trip_slot.size), nslices)
        out = np.zeros((t.size, 4), np.int32)
        out[:, 0] = t
        out[:, 1] = 1
        out[:, 2] = np.tile(np.arange(nslices) * TILE_ROWS, trip_slot.size)
        out[:, 3] = TILE_ROWS
        return out
    RB = 1
    while RB < L:
        RB *= 2                      # rows per triplet padded to a power of two (divides 256)
    G = TILE_ROWS // RB
    starts = np.flatnonzero(np.r_[True, trip_slot[1:] != trip_slot[:-1]])
    counts = np.diff(np.r_[starts, trip_slot.size])
    ntiles = (counts + G - 1) // G
    run_of = np.repeat(np.arange(starts.size), ntiles)
    first = np.cumsum(ntiles) - ntiles
    g0 = (np.arange(run_of.size) - first[run_of]) * G
    out = np.zeros((run_of.size, 4), np.int32)
    out[:, 0] = starts[run_of] + g0
    out[:, 1] = np.minimum(G, counts[run_of] - g0)
    out[:, 2] = 0
    out[:, 3] = RB
    return out


def shard_rows(num_rows: int, rank: int, world: int) -> slice:
    """Contiguous block partition of ``num_rows`` units over ``world`` ranks (first ranks get the
    remainder).  Used for queries (stage II) and gallery rows (stage I)."""
    base, rem = divmod(num_rows, world)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))
