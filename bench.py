"""bench.py -- stage-II re-ranked triplets/s (BASELINE.json metric) on N B200s of one node.

A "step" is one pass of the hot path over one batch of synthetic input: for every query of a
Fashion-IQ-val-shaped category (Q queries x top-K=100 candidates) compute z_t (stage-I encoder on the
reference image's tokens), score all Q*K triplets with the dual-stream stage-II encoder
(candidate-major, K/V once per unique candidate), re-sort each row and count Recall@{10,50}.
Gallery ViT tokens are the resident cache the reference also keeps on the device
(src/utils.py:43-55); they are produced once, outside the timed region, by the ViT kernel path.

  value : whole-job triplets/s with the step's inputs already resident in HBM
  e2e   : the same metric through the public validate_stage2.compute_fiq_val_metrics call with HOST
          (pinned) token ids / masks / candidate lists / labels copied H2D inside the timed region and
          the order + scores read back D2H
  roofline : the dominant kernel (tcgen05 GEMM): algorithmic FLOPs / CUDA-event duration per launch,
          against the measured sustained bf16 peak in MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (a port of the reference's PyTorch arithmetic) on the host cores, on a
          bounded sample of the same workload

N > 1 (torchrun): one process per GPU, weak scaling (every rank re-ranks its own Q x K block, gallery
tokens replicated), ranks exchange only (score, order) rows with one NCCL all-gather per step.
`--impl reference` times the reference's CPU implementation of the path (oracle port) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

F_REF_GF = 47.247          # algorithmic GFLOP per triplet as the reference executes them (L=32)
METRIC = "stage2_reranked_triplets_per_s"
UNIT = "triplets/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=2017)      # Fashion-IQ dress val (analysis_plot/fiq_stageII_labels_val_dress.pt)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--gallery", type=int, default=2297)      # CIRR-val-sized token gallery (BASELINE configs[1])
    ap.add_argument("--length", type=int, default=32)
    ap.add_argument("--cpu-sample-triplets", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons}


def synth_workload(args, rank):
    import cir_b200 as cir
    syn = cir.synthetic
    Q, K, G, L = args.queries, args.k, args.gallery, args.length
    ref, tgt, ids, mask = syn.make_queries(Q, G, L, seed=300 + rank)
    g = torch.Generator().manual_seed(400 + rank)
    # stage-I-like candidate lists: K distinct gallery rows per query, never the reference; the target planted in ~98 %
    scores = torch.rand(Q, G, generator=g)
    scores[torch.arange(Q), ref] = -1.0
    hit = torch.rand(Q, generator=g) < 0.98
    scores[torch.arange(Q)[hit], tgt[hit]] = 2.0
    scores[torch.arange(Q)[~hit], tgt[~hit]] = -1.0
    cand = scores.topk(K, dim=1).indices
    cand = cand[:, torch.randperm(K, generator=g)].to(torch.int32)
    labels = cand.long() == tgt[:, None]
    return ref.int(), tgt, ids.int(), mask.int(), cand, labels


def run_reference(args, rank, world, emit):
    """Reference arm: the reference's CPU arithmetic (oracle port; the Python reference itself cannot
    travel to the box) on all host threads, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import cir_oracle as O
    import cir_b200 as cir
    syn = cir.synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd1 = syn.make_stage1_state_dict(0, 384, "reference")
    sd2 = syn.make_stage2_state_dict(0, 384, "reference")
    L, K = args.length, args.k
    n_trip = max(10, min(args.cpu_sample_triplets, 50))          # per step; ~4 s at ~12 triplets/s
    g = torch.Generator().manual_seed(0)
    tokens = torch.randn(n_trip + 1, 577, 768, generator=g)      # LayerNorm-like statistics (mean 0, std 1)
    ids, mask = syn.make_token_ids(1, L, seed=2)
    ids[:, 0] = syn.ENC_TOKEN_ID

    def step():
        with torch.no_grad():
            z = O.stage1_hidden(sd1, tokens[:1], ids, mask)
            s = O.stage2_score(sd2, z, ids, mask, tokens[1:])
            O.rerank_order(s[None])
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = n_trip / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"stage2_rerank_fiq_shape: Q={args.queries} queries x K={K} candidates per GPU, L={L} tokens, "
                                   f"G={args.gallery} gallery images (577 ViT-B/16 tokens each, resident), z_t + stage-II + re-sort + recall",
                       "sample": f"each step = 1 query x {n_trip} candidates of that workload on the host CPU"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"1 query x {n_trip} candidates per step (z_t + stage-II + sort), fp32, torch CPU {cores} threads"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def cpu_baseline(args):
    from oracle import cir_oracle as O
    import cir_b200 as cir
    syn = cir.synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd1 = syn.make_stage1_state_dict(0, 384, "reference")
    sd2 = syn.make_stage2_state_dict(0, 384, "reference")
    L = args.length
    per_q = 50
    nq = max(1, args.cpu_sample_triplets // per_q)
    g = torch.Generator().manual_seed(0)
    tokens = torch.randn(per_q + 1, 577, 768, generator=g)
    ids, mask = syn.make_token_ids(nq, L, seed=2)
    ids[:, 0] = syn.ENC_TOKEN_ID
    with torch.no_grad():                                        # warm-up
        O.stage2_score(sd2, O.stage1_hidden(sd1, tokens[:1], ids[:1], mask[:1]), ids[:1], mask[:1], tokens[1:9])
        t0 = time.perf_counter()
        for q in range(nq):
            z = O.stage1_hidden(sd1, tokens[:1], ids[q:q + 1], mask[q:q + 1])
            s = O.stage2_score(sd2, z, ids[q:q + 1], mask[q:q + 1], tokens[1:])
            O.rerank_order(s[None])
        dt = time.perf_counter() - t0
    return {"value": nq * per_q / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{nq} queries x {per_q} candidates (z_t + stage-II + sort per query, as src/validate_stage2.py:94-125), fp32, torch CPU {cores} threads"}


def main():
    # stdout carries exactly ONE JSON line: everything else a library prints there (e.g. "NCCL version ..." on the
    # first collective) is routed to stderr by pointing fd 1 at fd 2 until the line is ready
    sys.stdout.flush()
    _stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return
    import cir_b200 as cir
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    syn = cir.synthetic
    Q, K, G, L = args.queries, args.k, args.gallery, args.length

    # ---- models (random-init, reference init style) and the resident gallery token cache
    sd1 = syn.make_stage1_state_dict(0, 384, "reference")
    sd2 = syn.make_stage2_state_dict(0, 384, "reference")
    m1 = cir.blip_stage1.blip_stage1(image_size=384, state_dict=sd1, precision="bf16", device=dev)
    m2 = cir.blip_stage2.blip_stage2(image_size=384, state_dict=sd2, precision="bf16", device=dev)
    eng = m2.engine
    del sd1, sd2
    tokens = torch.empty(G, 577, 768, dtype=torch.bfloat16, device=dev)
    gi = torch.Generator().manual_seed(1)
    for g0 in range(0, G, 64):                                   # ViT-B/16 on synthetic 384x384 images (not timed)
        n = min(64, G - g0)
        tokens[g0:g0 + n] = m2.img_embed(torch.randn(n, 3, 384, 384, generator=gi))
    torch.cuda.synchronize()

    ref, tgt, ids, mask, cand, labels = synth_workload(args, rank)
    names = syn.index_names_for(G)
    ref_d, ids_d, mask_d = ref.to(dev), ids.to(dev), mask.to(dev)
    cand_np = cand.numpy()
    labels_d = labels.to(dev)
    row_active = labels.any(1).numpy()
    n_trip = int(row_active.sum()) * K                            # rows without a positive are filled, not scored (:95,:123)

    def step_resident():
        z_t, _ = m1.encode_queries(tokens, ref_d, ids_d, mask_d, want_z=True, want_emb=False)
        scores = m2.score_triplets(z_t, ids_d, mask_d, tokens, cand_np, row_active)
        order = eng.rerank_sort(scores)
        hits = eng.recall_counts(labels_d, order, (10, 50))
        if world > 1:
            out = [torch.empty_like(scores) for _ in range(world)]
            dist.all_gather(out, scores)
            outo = [torch.empty_like(order) for _ in range(world)]
            dist.all_gather(outo, order)
        return scores, order, hits

    # host-resident inputs for the e2e leg (pinned)
    tb = syn.TokenBatch(input_ids=ids.long().pin_memory(), attention_mask=mask.long().pin_memory())
    ds = syn.SyntheticRelativeDataset(names, ref, tgt, ["x"] * Q, cand_np, kind="fiq", token_batch=tb)

    V2 = cir.validate_stage2

    def step_e2e():
        # the reference's own entry point for this path (src/validate_stage2.py:33-66) on a dataset object whose
        # token ids / masks / candidate names / labels live in HOST memory: every step re-tokenises from the host
        # batch, maps names to gallery rows, uploads ids, masks, per-chunk triplet lists and labels (H2D), scores,
        # re-sorts, and reads the recall counters back (D2H)
        r10, r50 = V2.compute_fiq_val_metrics(ds, m2, m1, tokens, names)
        torch.cuda.synchronize()
        return r10, r50
    # bytes per step, counted from the tensors the call copies: int64 ids+mask, int32 reference rows, bool labels,
    # per triplet flat position (8) + query row (4) + candidate slot (4), per-chunk candidate/query lists and
    # attention work lists (16 B per 8 query tiles + 16 B per 128-row tile + 16 B per 128 CLS rows)
    h2d = 2 * Q * L * 8 + Q * 4 + Q * K + n_trip * 16 + n_trip * (16 // 4 + 16 // 4 + 1) + 2 * (Q + G) * 4
    d2h = 2 * 8
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    eng.launch_count(reset=True)
    eng.profile_gemm(True)
    ms_step = timed(step_resident, args.steps)
    eng.profile_gemm(False)
    launches = eng.launch_count() + m1.engine.launch_count() * 0   # same engine object for both models
    gemm_ms, gemm_flops, gemm_n = eng.profile_gemm_read()
    step_e2e()
    ms_e2e = timed(step_e2e, max(1, args.steps))
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- secondary figures of the same path (outside the timed region): stage-I candidate filtering
    #      (BASELINE configs[1]: cosine + top-100 over the 2.3k gallery, Q = 4,181) and ViT token extraction
    def _time(fn, it=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / it
    extras = {}
    if rank == 0:
        gq = torch.Generator().manual_seed(5)
        qe = torch.nn.functional.normalize(torch.randn(4181, 256, generator=gq), dim=-1).to(dev)
        ge = m1.engine.stage1_gallery_embed(m1._w, tokens)
        ex = torch.randint(0, G, (4181,), generator=gq)
        ms1 = _time(lambda: eng.stage1_topk(qe, ge, 100, exclude=ex))
        img = torch.randn(64, 3, 384, 384, device=dev)
        msv = _time(lambda: eng.vit_forward(m2._vit, img, batch=64), it=2)
        extras = {"stage1_topk": {"queries_per_s": 4181 / (ms1 / 1e3), "Q": 4181, "G": G, "K": 100, "ms": ms1,
                                  "note": "fused fp32 1 - q @ G^T + per-query top-100 with the reference index excluded"},
                  "vit_b16_384": {"images_per_s": 64 / (msv / 1e3), "batch": 64, "ms": msv}}

        # the reference's per-query formulation on stock PyTorch (ATen / cuBLAS) on this same B200: the oracle's torch
        # code with weights and inputs moved to the GPU -- the library baseline SURVEY 8(d) asks for, since the
        # reference ships no Blackwell kernel.  Bounded sample: 4 queries x K candidates, fp32 and bf16 autocast.
        try:
            from oracle import cir_oracle as O
            sd1c = {k: v.to(dev) for k, v in syn.make_stage1_state_dict(0, 384, "reference").items()}
            sd2c = {k: v.to(dev) for k, v in syn.make_stage2_state_dict(0, 384, "reference").items()}
            tok32 = tokens[: K + 1].float()
            ids_l, mask_l = ids_d[:4].long(), mask_d[:4].long()

            def torch_pass():
                with torch.no_grad():
                    for q in range(4):
                        z = O.stage1_hidden(sd1c, tok32[:1], ids_l[q:q + 1], mask_l[q:q + 1])
                        sc = O.stage2_score(sd2c, z, ids_l[q:q + 1], mask_l[q:q + 1], tok32[1:])
                        torch.argsort(sc, descending=True)
            ms32 = _time(torch_pass, it=2)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ms16 = _time(torch_pass, it=2)
            extras["stock_pytorch_b200"] = {"fp32_triplets_per_s": 4 * K / (ms32 / 1e3), "bf16_autocast_triplets_per_s": 4 * K / (ms16 / 1e3),
                                            "sample": f"4 queries x {K} candidates, per-query loop as in src/validate_stage2.py:94-125 (z_t + "
                                                      "img_txt_fusion_val + argsort), oracle torch code on cuda"}
            del sd1c, sd2c, tok32
        except Exception as ex:                                  # never let a baseline leg break the bench line
            extras["stock_pytorch_b200"] = {"error": repr(ex)[:200]}

    total_trip = n_trip
    if world > 1:
        t = torch.tensor([float(n_trip)], device=dev)
        dist.all_reduce(t)
        total_trip = float(t.item())
    value = total_trip / (ms_step / 1e3)
    e2e_value = total_trip / (ms_e2e / 1e3)
    peak, how = measured_peaks()
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    if rank == 0:
        cpu = None if args.no_cpu_baseline else cpu_baseline(args)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"stage2_rerank_fiq_shape: Q={Q} queries x K={K} candidates per GPU, L={L} tokens, "
                                   f"G={G} gallery images (577 ViT-B/16 tokens each, resident), z_t + stage-II + re-sort + recall",
                       "triplets_scored_per_gpu_step": n_trip, "chunk_triplets": eng.max_triplets, "chunk_candidates": eng.max_candidates,
                       "l2_note": "per-step working set (gallery tokens 2.0 GB + activations) exceeds the 126 MB L2; no flush needed",
                       "parallelism": f"dp{world} (queries sharded, weights+gallery replicated, NCCL all-gather of scores/order)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "tc::gemm_tcgen05_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "peak_source": how, "traffic": None,
                         "launches": int(gemm_n), "gflop_per_launch": gemm_flops / max(1, gemm_n) / 1e9,
                         "avg_launch_ms": gemm_ms / max(1, gemm_n), "gemm_share_of_step": gemm_ms / (ms_step * args.steps),
                         "effective_tflops_at_F_ref": value / world * F_REF_GF / 1e3,
                         "effective_frac_of_peak": value / world * F_REF_GF / 1e3 / peak if peak else None},
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "extras": extras,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
