"""Quick on-box performance probe (not a bench): GEMM TF/s at the path's shapes, attention time,
stage-II chunk throughput.  Prints one line per measurement."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cir_b200 as cir

N_ = cir.native
e = cir.engine.get_engine(precision="bf16")


def timeit(fn, warm=2, it=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    t.record()
    torch.cuda.synchronize()
    return s.elapsed_time(t) / it


def gemm_probe():
    for (M, Nn, K, b, act) in [(8192, 8192, 8192, 1, 0), (65536, 768, 768, 2, 0), (65536, 2304, 768, 2, 0), (36928, 3072, 768, 1, 0),
                               (131072, 3072, 768, 1, 1), (131072, 768, 3072, 1, 0), (65536, 768, 1536, 1, 0), (4096, 768, 768, 2, 0)]:
        A = torch.randn(b, M, K, device="cuda").bfloat16()
        W = (torch.randn(b, Nn, K, device="cuda") * 0.05).bfloat16()
        bias = torch.randn(b, Nn, device="cuda")
        ms = timeit(lambda: e.gemm(A, W, bias, act=act))
        tf = 2.0 * M * Nn * K * b / ms / 1e9
        e.set_gemm_tma_store(False)
        ms_l = timeit(lambda: e.gemm(A, W, bias, act=act))
        e.set_gemm_tma_store(True)
        print(f"   (per-lane stores: {ms_l:.3f} ms {2.0 * M * Nn * K * b / ms_l / 1e9:.0f} TF/s)", flush=True)
        ms_t = timeit(lambda: torch.matmul(A, W.transpose(1, 2)))
        tf_t = 2.0 * M * Nn * K * b / ms_t / 1e9
        print(f"gemm M={M} N={Nn} K={K} batch={b} act={act}: {ms:.3f} ms {tf:.0f} TF/s | cuBLAS {ms_t:.3f} ms {tf_t:.0f} TF/s", flush=True)
        del A, W


def stage2_probe(T_per=2048, C=48, L=32, reps=3, configs=((1024, 32), (2048, 48), (4096, 64), (8192, 128)), Q=512):
    syn = cir.synthetic
    sd2 = syn.make_stage2_state_dict(0, 384, "reference")
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="bf16")
    G = int(os.environ.get("CIR_PROBE_G", "256"))
    tokens = torch.randn(G, 577, 768, device="cuda").bfloat16()
    K = int(os.environ.get("CIR_PROBE_K", "50"))
    ids, mask = syn.make_token_ids(Q, L, seed=2)
    ids[:, 0] = 30523
    z_t = torch.randn(Q, L, 768, device="cuda").bfloat16()
    g = torch.Generator().manual_seed(0)
    cand = torch.stack([torch.randperm(G, generator=g)[:K] for _ in range(Q)]).int().numpy()
    for (mt, mc) in configs:
        m2.engine.max_triplets, m2.engine.max_candidates = mt, mc
        e.launch_count(reset=True)
        ms = timeit(lambda: m2.score_triplets(z_t, ids, mask, tokens, cand), warm=1, it=reps)
        nl = e.launch_count() / (reps + 1)
        print(f"stage2 Q={Q} K={K} L={L} G={G} chunk(T<={mt},C<={mc}): {ms:.1f} ms -> {Q*K/ms*1e3:.0f} triplets/s, {nl:.0f} launches/pass", flush=True)
        if os.environ.get("CIR_PROBE_DEDUP"):
            e.set_dedup_first_layer(False)
            ms = timeit(lambda: m2.score_triplets(z_t, ids, mask, tokens, cand), warm=1, it=reps)
            e.set_dedup_first_layer(True)
            print(f"   without first-layer dedup: {ms:.1f} ms -> {Q*K/ms*1e3:.0f} triplets/s", flush=True)


def stage1_probe():
    """stage-I candidate filtering: fused 1 - q @ G^T + per-query top-K (validate.py:57-58,202-210)"""
    e32 = e
    for (Q, G, K) in ((4181, 2297, 100), (4181, 100000, 100), (4181, 1000000, 200)):
        g = torch.Generator().manual_seed(0)
        q = torch.nn.functional.normalize(torch.randn(Q, 256, generator=g), dim=-1).cuda()
        gal = torch.nn.functional.normalize(torch.randn(G, 256, device="cuda"), dim=-1)
        excl = torch.randint(0, G, (Q,), generator=g)
        ms = timeit(lambda: e32.stage1_topk(q, gal, K, exclude=excl), warm=1, it=3)
        print(f"stage1 top-{K}: Q={Q} G={G}: {ms:.1f} ms -> {Q/ms*1e3:.0f} queries/s ({2.0*Q*G*256/ms/1e9:.1f} TF/s fp32 sims)", flush=True)


def vit_probe():
    """gallery token extraction: ViT-B/16 @384 (src/vit.py:180-194), 110.967 GF per image"""
    syn = cir.synthetic
    sd2 = syn.make_stage2_state_dict(0, 384, "reference")
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="bf16")
    for B in (64, 128):
        img = torch.randn(B, 3, 384, 384, device="cuda")
        ms = timeit(lambda: m2.engine.vit_forward(m2._vit, img, batch=B), warm=1, it=3)
        print(f"vit B={B}: {ms:.1f} ms -> {B/ms*1e3:.0f} images/s, {B*110.967/ms:.0f} TF/s algorithmic", flush=True)


def attn_probe():
    """cross-attention shape of one stage-II chunk: T triplets x 32 rows vs 577 keys, candidate runs of ~91"""
    T, L, Lk, C = 4096, 32, 577, 45
    slot = torch.sort(torch.arange(T) % C).values.int()
    q = torch.randn(T, L, 768, device="cuda").bfloat16()
    kv = torch.randn(C, Lk, 3072, device="cuda").bfloat16()
    o = torch.empty(T, L, 1536, device="cuda", dtype=torch.bfloat16)
    import ctypes as C_
    sched = cir.schedule
    tiles = e._i32(sched.build_attn_tiles(slot.numpy(), L))
    work = e._i32(sched.build_attn_work(slot.numpy(), L))
    slot_d = slot.cuda()
    flops = 4.0 * T * L * Lk * 64 * 12

    def call(impl, use_tiles):
        a = N_.AttnArgs()
        a.q, a.k, a.v, a.o = N_.ptr(q), N_.ptr(kv), N_.vp(kv.data_ptr() + 768 * 2), N_.ptr(o)
        a.q_bs, a.q_rs = L * 768, 768
        a.k_bs = a.v_bs = Lk * 3072
        a.k_rs = a.v_rs = 3072
        a.o_bs, a.o_rs = L * 1536, 1536
        a.kv_index = N_.ptr(slot_d)
        a.key_mask = a.mask_index = N_.vp(0)
        a.work, a.num_work = (N_.ptr(work), work.shape[0]) if not use_tiles else (N_.vp(0), 0)
        a.tiles, a.num_tiles = (N_.ptr(tiles), tiles.shape[0]) if use_tiles else (N_.vp(0), 0)
        a.kv_batches = C
        a.B, a.H, a.Lq, a.Lk, a.scale = T, 12, L, Lk, 0.125
        e.set_attention_impl(impl)
        e._sync_stream()
        N_.check(e._lib.cir_attention(e.ctx, C_.byref(a)))
        e.set_attention_impl(0)
    for name, impl, ut in (("tcgen05", 0, True), ("mma.sync", 2, False)):
        ms = timeit(lambda: call(impl, ut), warm=2, it=10)
        print(f"cross-attn T={T} L={L} Lk={Lk} runs of ~{T//C}: {name}: {ms:.3f} ms  {flops/ms/1e9:.0f} TF/s (algorithmic)", flush=True)


def qkv_probe(T=4096, L=32):
    """fused QKV + self-attention vs QKV GEMM + attention kernel at the stage-II chunk shape"""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, T, L, 768, generator=g).cuda().bfloat16()
    w = (torch.randn(2, 2304, 768, generator=g) * 0.04).cuda().bfloat16()
    bias = torch.randn(2, 2304, generator=g).cuda()
    mask = torch.ones(T, L, dtype=torch.int32).cuda()
    flops = 2 * 2 * T * L * 2304 * 768 + 4 * 2 * T * 12 * L * L * 64
    ms = timeit(lambda: e.qkv_attention(x, w, bias, key_mask=mask), warm=2, it=10)
    print(f"fused qkv+attention T={T} L={L}: {ms:.3f} ms  {flops/ms/1e9:.0f} TF/s  (CIR_QKV_DRAIN={os.environ.get('CIR_QKV_DRAIN')})", flush=True)
    x2 = x.reshape(2, T * L, 768)
    ms_g = timeit(lambda: e.gemm(x2, w, bias), warm=2, it=10)
    qkv = e.gemm(x2, w, bias)[0].reshape(T, L, 2304)
    q, k, v = qkv[..., :768], qkv[..., 768:1536], qkv[..., 1536:]
    ms_a = timeit(lambda: e.attention(q, k, v, key_mask=mask), warm=2, it=10)
    print(f"unfused: QKV GEMM {ms_g:.3f} ms + 2 x attention {ms_a:.3f} ms = {ms_g + 2 * ms_a:.3f} ms", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "stage2"]
    if "gemm" in which:
        gemm_probe()
    if "stage2" in which:
        stage2_probe()
    if "vit" in which:
        vit_probe()
    if "stage1" in which:
        stage1_probe()
    if "attn" in which:
        attn_probe()
    if "qkv" in which:
        qkv_probe()
    if "stage2_profile" in which:      # one short pass for an ncu launch list
        stage2_probe(reps=1, configs=((4096, 64),), Q=int(os.environ.get("CIR_PROBE_Q", "96")))


def power_probe(seconds=4.0):
    """Which kernels run into the board's power cap?  Each kernel class of the stage-II step is looped alone for a few seconds while
    nvidia-smi samples SM clock and power draw."""
    import subprocess, threading
    T, L, C = 4096, 32, 46
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, T, L, 768, generator=g).cuda().bfloat16()
    w = (torch.randn(2, 2304, 768, generator=g) * 0.04).cuda().bfloat16()
    bias = torch.randn(2, 2304, generator=g).cuda()
    mask = torch.ones(T, L, dtype=torch.int32).cuda()
    A = torch.randn(2 * T * L, 768, device="cuda").bfloat16()
    W1 = (torch.randn(3072, 768, device="cuda") * 0.04).bfloat16()
    gam, bet = torch.ones(768, device="cuda"), torch.zeros(768, device="cuda")
    slot = torch.sort(torch.arange(T) % C).values.int()
    q = torch.randn(T, L, 768, device="cuda").bfloat16()
    kv = torch.randn(C, 577, 3072, device="cuda").bfloat16()
    o = torch.empty(T, L, 1536, device="cuda", dtype=torch.bfloat16)
    import ctypes as C_
    tiles = e._i32(cir.schedule.build_attn_tiles(slot.numpy(), L))
    slot_d = slot.cuda()

    def attn():
        a = N_.AttnArgs()
        a.q, a.k, a.v, a.o = N_.ptr(q), N_.ptr(kv), N_.vp(kv.data_ptr() + 768 * 2), N_.ptr(o)
        a.q_bs, a.q_rs = L * 768, 768
        a.k_bs = a.v_bs = 577 * 3072
        a.k_rs = a.v_rs = 3072
        a.o_bs, a.o_rs = L * 1536, 1536
        a.kv_index = N_.ptr(slot_d)
        a.key_mask = a.mask_index = N_.vp(0)
        a.work, a.num_work = N_.vp(0), 0
        a.tiles, a.num_tiles = N_.ptr(tiles), tiles.shape[0]
        a.kv_batches = C
        a.B, a.H, a.Lq, a.Lk, a.scale = T, 12, L, 577, 0.125
        e._sync_stream()
        N_.check(e._lib.cir_attention(e.ctx, C_.byref(a)))
    kernels = {"gemm FFN1 (262144 x 3072 x 768)": lambda: e.gemm(A, W1, None, act=1),
               "fused QKV + self-attention": lambda: e.qkv_attention(x, w, bias, key_mask=mask),
               "tcgen05 cross-attention": attn,
               "LayerNorm (262144 rows)": lambda: e.add_layernorm(A, gam, bet)}
    for name, fn in kernels.items():
        rows, stop = [], [False]

        def sample():
            while not stop[0]:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", "0"],
                                     capture_output=True, text=True).stdout.strip()
                if out:
                    rows.append([c.strip() for c in out.split(",")])
                time.sleep(0.1)
        th = threading.Thread(target=sample, daemon=True)
        fn(); torch.cuda.synchronize()
        th.start()
        t0 = time.time(); n = 0
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        while time.time() - t0 < seconds:
            for _ in range(20):
                fn()
            n += 20
            torch.cuda.synchronize()
        t.record(); torch.cuda.synchronize()
        stop[0] = True; th.join()
        rows = rows[len(rows) // 3:]                         # steady state
        clk = sorted(int(r[0]) for r in rows)[len(rows) // 2] if rows else -1
        pw = sorted(float(r[1]) for r in rows)[len(rows) // 2] if rows else -1
        cap = sum(r[2].lower().startswith("active") for r in rows) / max(1, len(rows))
        print(f"{name:40s} {s.elapsed_time(t) / n:.3f} ms/launch  SM clock {clk} MHz  power {pw:.0f} W  sw_power_cap active in {100 * cap:.0f} % of samples", flush=True)


if "power" in sys.argv[1:]:
    power_probe()
