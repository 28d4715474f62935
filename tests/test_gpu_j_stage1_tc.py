"""Stage-I top-K with the tensor-core candidate filter (stage1_topk_tc.cu: bf16 tcgen05 similarity tiles + rigorous error margin +
exact fp32 re-check) against the fp32 CUDA-core path it replaces for large galleries: (distance, index) lists must be identical
BIT FOR BIT -- ties, the excluded reference index, shard offsets, near-duplicate rows and the overflow fallback included -- and
equal to the CPU oracle's ranking (src/validate.py:57-58,202-210)."""
import numpy as np
import pytest
import torch

import cir_b200 as cir
from oracle import cir_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    return cir.engine.get_engine(precision="bf16")


def _unit(n, seed, dim=256):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=-1)


def _both(eng, q, g, k, **kw):
    d1, i1 = eng.stage1_topk(q, g, k, **kw)
    eng.set_stage1_tensor_cores(False)
    try:
        d0, i0 = eng.stage1_topk(q, g, k, **kw)
    finally:
        eng.set_stage1_tensor_cores(True)
    return (d1, i1), (d0, i0)


@pytest.mark.parametrize("Q,G,K", [(37, 20000, 50), (300, 70001, 200), (5, 16384, 1), (64, 33000, 1000)])
def test_tensor_core_filter_equals_fp32_path(eng, Q, G, K):
    q, g = _unit(Q, 1), _unit(G, 2)
    ex = torch.randint(0, G, (Q,), generator=torch.Generator().manual_seed(3))
    ex[::3] = -1
    (d1, i1), (d0, i0) = _both(eng, q, g, K, exclude=ex, col_offset=1000)
    assert torch.equal(i1, i0) and torch.equal(d1, d0)
    assert not (i1 == (ex[:, None].cuda() )).any()
    # against the CPU oracle (its fp32 matmul sums in another order: compare where its own distances are not within 1e-6 of a swap)
    if G <= 20000:
        want_d, want_i = O.stage1_topk(q, g, ex, K)
        got = i1.cpu().long() - 1000
        differ = got != want_i.long()
        if differ.any():
            gap = (want_d[:, 1:] - want_d[:, :-1]).abs()
            near = torch.zeros_like(differ)
            near[:, 1:] |= gap < 1e-6
            near[:, :-1] |= gap < 1e-6
            assert not (differ & ~near).any()
        assert (d1.cpu() - want_d).abs().max() < 1e-5


def test_ties_and_near_duplicates(eng):
    """Blocks of identical gallery rows (exact ties -> lowest index first) and rows that differ by one bf16 rounding step."""
    Q, G, K = 16, 40000, 100
    q, g = _unit(Q, 5), _unit(G, 6)
    g[1000:1040] = g[999]                      # 41 identical rows
    g[20000:20030] = g[7] * (1 + 1e-4)         # near-duplicates of one row, distinguishable only in fp32
    g[30000] = q[3]                            # an exact match for query 3
    (d1, i1), (d0, i0) = _both(eng, q, g, K)
    assert torch.equal(i1, i0) and torch.equal(d1, d0)
    assert int(i1[3, 0]) == 30000


def test_adversarial_order_falls_back_and_stays_exact(eng):
    """Gallery sorted by increasing similarity to one query: every super-block beats the running threshold, the candidate list of
    that query overflows and the call falls back to the fp32 path -- same results."""
    Q, G, K = 4, 60000, 50
    q = _unit(Q, 8)
    base = _unit(G, 9)
    sim = base @ q[0]
    g = base[torch.argsort(sim)]               # ascending similarity to query 0 -> its best rows come last
    (d1, i1), (d0, i0) = _both(eng, q, g, K)
    assert torch.equal(i1, i0) and torch.equal(d1, d0)
    assert int(i1[0, 0]) == G - 1


def test_unnormalised_rows_keep_the_margin_rigorous(eng):
    """Rows with norms far from 1 (the bound scales with ||q|| max||g||)."""
    Q, G, K = 32, 25000, 64
    q = _unit(Q, 11) * torch.linspace(0.1, 30, Q)[:, None]
    g = _unit(G, 12) * torch.rand(G, generator=torch.Generator().manual_seed(13))[:, None] * 5
    (d1, i1), (d0, i0) = _both(eng, q, g, K)
    assert torch.equal(i1, i0) and torch.equal(d1, d0)
