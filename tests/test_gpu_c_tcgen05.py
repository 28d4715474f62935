"""The tcgen05/TMEM/TMA GEMM against the CUDA-core GEMM on identical bf16 operands (same fp32
accumulation, so agreement is ~1e-5 relative before the output rounding) and against torch."""
import pytest
import torch

import cir_b200 as cir

pytestmark = pytest.mark.gpu
N_ = cir.native


@pytest.fixture(scope="module")
def e16():
    return cir.engine.get_engine(precision="bf16")


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def _both(e, *args, **kw):
    e.set_gemm_impl(N_.GEMM_TCGEN05)
    tc = e.gemm(*args, **kw)
    e.set_gemm_impl(N_.GEMM_SIMT)
    ref = e.gemm(*args, **kw)
    e.set_gemm_impl(N_.GEMM_AUTO)
    return tc, ref


SHAPES = [
    (128, 256, 64, 1),      # one tile, one k-block
    (128, 256, 768, 1),     # pipeline wrap-around (12 k-blocks > 4 stages)
    (300, 768, 768, 2),     # M tail, batch folding, 3 N tiles
    (1000, 2304, 768, 2),   # self QKV shape
    (64, 768, 1536, 1),     # merged cross-out shape, M < tile
    (2308, 3072, 768, 1),   # K/V projection shape (4 x 577 rows), many tiles per CTA
    (500, 768, 3072, 1),    # FFN2 shape, long K
    (96, 200, 128, 1),      # N tail not a multiple of 32
    (40, 100, 64, 3),       # tiny batched
    (3, 256, 768, 1),       # stage-I projection (M=3)
    (5000, 768, 768, 2),    # cta_group::2 pair tiles (>= 74 pair tiles), batch folding, M tail inside a pair
    (4100, 2304, 768, 1),   # pair tiles, odd number of 128-row blocks (peer CTA fully out of range on the last tile)
    (19000, 3072, 768, 1),  # pair tiles, several tiles per CTA pair (persistent loop + accumulator double buffering)
    (9500, 768, 3072, 1),   # pair tiles, long K (ring wraps many times)
]


@pytest.mark.parametrize("M,N,K,batch", SHAPES)
def test_tcgen05_matches_simt(e16, M, N, K, batch):
    A = _rand(batch, M, K, seed=1).bfloat16()
    W = _rand(batch, N, K, seed=2, scale=0.05).bfloat16()
    bias = _rand(batch, N, seed=3)
    tc, ref = _both(e16, A, W, bias, out_f32=True)
    assert torch.isfinite(tc).all()
    err = (tc - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"tcgen05 vs simt max err {err}"
    gold = torch.einsum("bmk,bnk->bmn", A.double(), W.double()) + bias.double()[:, None, :]
    assert (tc.double() - gold).abs().max() < 1e-3 * max(1.0, gold.abs().max().item())


def test_tcgen05_epilogues(e16):
    M, N, K = 260, 768, 768
    A = _rand(2, M, K, seed=1).bfloat16()
    W = _rand(2, N, K, seed=2, scale=0.05).bfloat16()
    bias = _rand(2, N, seed=3)
    res32 = _rand(2, M, N, seed=4)
    res16 = res32.bfloat16()
    gold = torch.einsum("bmk,bnk->bmn", A.double(), W.double()) + bias.double()[:, None, :]
    for kw, want in (
        (dict(act=N_.ACT_GELU, out_f32=True), torch.nn.functional.gelu(gold)),
        (dict(act=N_.ACT_RELU, out_f32=True), gold.clamp_min(0)),
        (dict(residual=res32, out_f32=True), gold + res32.double()),
        (dict(residual=res16, out_f32=True), gold + res16.double()),
        (dict(act=N_.ACT_GELU), torch.nn.functional.gelu(gold)),
        (dict(), gold),
    ):
        e16.set_gemm_impl(N_.GEMM_TCGEN05)
        out = e16.gemm(A, W, bias, **kw)
        e16.set_gemm_impl(N_.GEMM_AUTO)
        tol = 2e-3 if out.dtype == torch.float32 else 3e-2
        assert (out.double() - want).abs().max() < tol * max(1.0, want.abs().max().item()), kw


def test_pair_tile_matches_single_cta_and_epilogues(e16):
    M, N, K = 3300, 768, 768
    A = _rand(2, M, K, seed=1).bfloat16()
    W = _rand(2, N, K, seed=2, scale=0.05).bfloat16()
    bias = _rand(2, N, seed=3)
    res16 = _rand(2, M, N, seed=4).bfloat16()
    for kw in (dict(out_f32=True), dict(act=N_.ACT_GELU), dict(residual=res16, out_f32=True), dict(residual=res16)):
        e16.set_gemm_impl(N_.GEMM_TCGEN05)
        pair = e16.gemm(A, W, bias, **kw)
        e16.set_gemm_impl(N_.GEMM_TCGEN05_1CTA)
        single = e16.gemm(A, W, bias, **kw)
        e16.set_gemm_impl(N_.GEMM_AUTO)
        assert torch.equal(pair, single), kw          # same MMA order along K -> bit-identical


def test_tcgen05_shared_a_and_strided_rows(e16):
    """Stage-I projection reads CLS rows with lda = L*768; exercised through the C-ABI directly."""
    import ctypes as C
    L, Q = 12, 37
    h = _rand(Q, L, 768, seed=5).bfloat16()
    W = _rand(256, 768, seed=6, scale=0.05).bfloat16()
    out = torch.empty(Q, 256, device="cuda")
    g = N_.GemmArgs()
    g.A, g.W, g.C = N_.ptr(h), N_.ptr(W), N_.ptr(out)
    g.M, g.N, g.K = Q, 256, 768
    g.lda, g.ldw, g.ldc = L * 768, 768, 256
    g.batch, g.act, g.c_f32 = 1, 0, 1
    e16._sync_stream()
    N_.check(e16._lib.cir_gemm(e16.ctx, C.byref(g)))
    want = h[:, 0, :].double() @ W.double().T
    assert (out.double() - want).abs().max() < 2e-3


def test_tma_store_epilogue_is_bit_identical(e16):
    """bf16 outputs through cp.async.bulk.tensor stores vs. per-lane stores: same bytes, M tails clipped by the map,
    rows of neighbouring batches untouched."""
    for (M, N, K, b) in ((3300, 768, 768, 2), (130, 2304, 768, 2), (577 * 3, 3072, 768, 1), (31, 64, 64, 3)):
        A = _rand(b, M, K, seed=1).bfloat16()
        W = _rand(b, N, K, seed=2, scale=0.05).bfloat16()
        bias = _rand(b, N, seed=3)
        res16 = _rand(b, M, N, seed=4).bfloat16()
        for kw in (dict(), dict(act=N_.ACT_GELU), dict(residual=res16)):
            e16.set_gemm_impl(N_.GEMM_TCGEN05)
            try:
                e16.set_gemm_tma_store(True)
                tma = e16.gemm(A, W, bias, **kw)
                e16.set_gemm_tma_store(False)
                lane = e16.gemm(A, W, bias, **kw)
            finally:
                e16.set_gemm_tma_store(True)
                e16.set_gemm_impl(N_.GEMM_AUTO)
            assert torch.equal(tma, lane), (M, N, K, b, kw)
