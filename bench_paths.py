#!/usr/bin/env python
"""bench_paths.py -- the other hot-path rows of SURVEY.md section 8 (BASELINE.json configs C1, C2, C4, C5).

`bench.py` stays the headline (C3) and calls `run_extra` / the latency helpers below so that these rows
appear in the driver-run line under `extra`; the command line runs one of them alone:

    python bench_paths.py --workload c2        # 1M x 768 fp32 exact DEWI-re-ranked top-10, B = 1..4096
    python bench_paths.py --workload c4        # fit_stats + score over 100M Signals rows
    python bench_paths.py --workload c5        # redundancy self-join, 512-d, cosine threshold (bounded size)
    python bench_paths.py --workload latency   # README quick-start flow (10K docs) + 1M rows, single-query latency

Every row carries `roofline` (algorithmic bytes or flops / CUDA-event time vs MEASURED_PEAKS.json), `clocks`
sampled while it ran, and `cpu_baseline` (the reference's code path on the host cores, bench_ref.py).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def timed(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


# ---- single-query latency through the reference-facing call ------------------------------------------------
def single_query_latency(torch, index, dim, k, eta, pref, n_queries, label):
    """p50 / p99 of `index.search(query[dim], k, eta, entropy_pref)` -- host query in, host `(doc_id, score,
    Payload)` tuples out, the 1-D contract of index.py:77-93 -- after 10 warm-up queries."""
    rng = np.random.RandomState(5)
    qs = rng.standard_normal((n_queries + 10, dim)).astype(np.float32)
    for q in qs[:10]:
        index.search(q, k=k, eta=eta, entropy_pref=pref)
    ms = []
    for q in qs[10:]:
        t0 = time.perf_counter()
        index.search(q, k=k, eta=eta, entropy_pref=pref)
        ms.append((time.perf_counter() - t0) * 1e3)
    ms = np.asarray(ms)
    return {"corpus": label, "queries": int(n_queries), "p50": float(np.percentile(ms, 50)), "p99": float(np.percentile(ms, 99)),
            "mean": float(ms.mean()), "min": float(ms.min()),
            "api": "search(query[dim]) -> [(doc_id, score, Payload)], host in / host out"}


def quickstart_flow(torch, dewi_b200, n_docs=10_000, dim=768, seed=0):
    """BASELINE.json config 1 -- the README quick-start (README.md:67-110) at 10K documents, as ONE flow through
    the public API: Signals rows -> DewiScorer.fit_stats -> score -> Payload(dewi=..., **signals) ->
    DewiIndex.add -> build.  Returns (index, signal rows, embeddings, dewi scores)."""
    from dewi_b200 import DewiIndex, DewiScorer, Payload, Signals, Weights

    rng = np.random.RandomState(seed)
    emb = rng.rand(n_docs, dim).astype(np.float32)                       # README: np.random.rand(768)
    hi = (10, 15, 5, 8, 1, 1, 0.2)                                       # README.md:83-91 ranges
    cols = [rng.uniform(0, h, n_docs).astype(np.float32) for h in hi]
    rows = [Signals(*(float(c[i]) for c in cols)) for i in range(n_docs)]
    scorer = DewiScorer(Weights())
    scorer.fit_stats(rows)                                               # README.md:101-102
    dewi = scorer.score_batch(np.stack(cols), out_dtype="float64").cpu().numpy()
    index = DewiIndex(dim=dim, space="cosine")
    for i, sig in enumerate(rows):
        index.add(f"doc{i}", emb[i], Payload(dewi=float(dewi[i]), **sig.__dict__))   # README.md:93-98,104-110
    index.build()
    return index, rows, emb, dewi


def quickstart_latency(torch, dewi_b200, k=10):
    """C1 (10K x 768, the README flow) and C2-sized (1M x 768 fp32) single-query latency, eta = 0.3, pref = 0.5."""
    out = {}
    index, _, _, _ = quickstart_flow(torch, dewi_b200)
    out["c1"] = single_query_latency(torch, index, 768, k, 0.3, 0.5, 200, "README quick-start, 10000 x 768 fp32")
    out["c1"]["reference_cpu_ms"] = "0.35 (8-core survey host, BASELINE.md section 2); measured on this box in cpu_baseline.points"
    del index
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    n = 1_000_000
    ix = dewi_b200.DewiIndex(dim=768, space="cosine", dtype="fp32")
    pay = torch.rand((n, 8), generator=g, device=dev).cpu().numpy()
    ix.add_batch(None, torch.randn((n, 768), generator=g, device=dev), payload_columns=pay)
    ix.build()
    out["c2"] = single_query_latency(torch, ix, 768, k, 0.3, 0.5, 200, "1000000 x 768 fp32")
    del ix
    torch.cuda.empty_cache()
    return out


# ---- C2: 1M x 768 fp32 exact search ------------------------------------------------------------------------
def run_c2(rows, torch, dewi_b200, peaks):
    n, d, k = rows or 1_000_000, 768, 10
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    emb = torch.randn((n, d), generator=g, device=dev)
    ix = dewi_b200.CudaIndex(d, dtype="fp32", device=dev.index)
    ix.add_batch(None, emb, normalized=False)
    ix.set_payload_columns(torch.rand(n, generator=g, device=dev), torch.rand(n, generator=g, device=dev) * 3)
    ix.build()
    del emb
    out = []
    ridge = peaks["bf16_sustained"] * 1e12 * 4 / (2 * peaks["hbm_gbs"] * 1e9)
    for b in (1, 8, 64, 256, 1024, 4096):
        q = torch.randn((b, d), generator=g, device=dev)
        steps = 20 if b <= 256 else 5
        ix.set_profiling(True)
        ms = timed(torch, lambda: ix.search_batch(q, k=k, eta=0.3, entropy_pref=0.5), steps, 3)
        kms = float(np.mean([ix.sweep_ms(i)[0] for i in range(steps)]))
        kind = ix.sweep_ms(0)[1]
        ix.set_profiling(False)
        if b <= ridge:
            ach, peak, unit, bound = n * d * 4 / (kms / 1e3) / 1e9, peaks["hbm_gbs"], "GB/s", "hbm"
        else:
            ach, peak, unit, bound = 2.0 * b * n * d / (kms / 1e3) / 1e12, peaks["bf16_sustained"], "TFLOP/s", "tensor"
        out.append({"batch": b, "value": b / (ms / 1e3), "ms_per_step": ms, "kernel_ms": kms, "kernel": kind, "bound": bound,
                    "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak})
    del ix
    torch.cuda.empty_cache()
    head = out[2]
    return {"metric": "queries/sec (1M x 768 fp32 exact, k=10, DEWI re-rank)", "unit": "queries/s", "dtype": "f32 (bf16 hi/lo planes)",
            "value": head["value"],
            "config": {"workload": f"C2: {n} x {d} fp32 exact search, 1 B200, value at B=64", "l2": "corpus 3 GB >> L2"},
            "roofline": {"bound": head["bound"], "achieved": head["achieved"], "peak": head["peak"], "unit": head["unit"],
                         "frac": head["frac"], "kernel_ms": head["kernel_ms"], "algorithmic_bytes_per_launch": n * d * 4},
            "batches": out}


# ---- C4: fit_stats + score over 100M Signals rows ------------------------------------------------------------
def run_c4(rows, torch, dewi_b200, peaks):
    n = rows or 100_000_000
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(21)
    hi = torch.tensor([10, 15, 5, 8, 1, 1, 0.2], device=dev).view(7, 1)
    sig = torch.rand((7, n), generator=g, device=dev) * hi  # README.md:83-91 ranges
    s = dewi_b200.DewiScorer()
    fit_ms = timed(torch, lambda: s.fit_stats_columns(sig), 3, 1)
    score_ms = timed(torch, lambda: s.score_batch(sig), 5, 2)
    fit_bytes, score_bytes = 2 * n * 7 * 4, n * (7 * 4 + 4)
    del sig
    torch.cuda.empty_cache()

    def roof(bytes_, ms):
        ach = bytes_ / (ms / 1e3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "algorithmic_bytes": bytes_}

    return {
        "metric": "rows/sec (fit_stats + score over Signals rows)", "unit": "rows/s", "dtype": "f32 data, f64 arithmetic",
        "value": n / ((fit_ms + score_ms) / 1e3),
        "config": {"workload": f"C4: {n} Signals rows x 7 columns, 1 B200", "l2": "2.8 GB columns >> L2"},
        "fit_stats": {"ms": fit_ms, "rows_per_s": n / (fit_ms / 1e3), "roofline": roof(fit_bytes, fit_ms)},
        "score": {"ms": score_ms, "rows_per_s": n / (score_ms / 1e3), "roofline": roof(score_bytes, score_ms)},
        "roofline": roof(fit_bytes + score_bytes, fit_ms + score_ms),
    }


# ---- C5: redundancy self-join ----------------------------------------------------------------------------------
def run_c5(rows, torch, dewi_b200, peaks, dist_ready=False):
    """1 process: bounded self-join, bf16 and hi/lo planes.  Under torchrun: the row-sharded self-join
    (each rank owns a block of rows, all-gather once, its share of the symmetric block grid), bf16 planes."""
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    own_pg = world > 1 and not dist_ready
    if own_pg:
        dist.init_process_group("nccl", device_id=dev)
    n, d, tau = rows or (1_000_000 if world == 1 else 10_000_000), 512, 0.9
    lo, hi = dewi_b200.shard_range(n, world, rank, align=256)
    g = torch.Generator(device=dev)
    g.manual_seed(44 + rank)
    x = torch.randn((hi - lo, d), generator=g, device=dev)
    m = (hi - lo) // 100  # planted 1 % near-duplicates (within the shard)
    dup = torch.randperm(hi - lo, generator=g, device=dev)[:m]
    src = torch.randperm(hi - lo, generator=g, device=dev)[:m]
    x[dup] = x[src] + 0.05 * torch.randn((m, d), generator=g, device=dev)
    runs = []
    configs = (("bf16", n), ("fp32", max(n // 4, 1))) if world == 1 else (("bf16", n),)
    for precision, nrows in configs:
        res = {}
        xs = x[: nrows if world == 1 else hi - lo]

        def go():
            if world == 1:
                res["out"] = dewi_b200.redundancy_join(xs, tau=tau, pair_cap=1 << 22, precision=precision)
            else:
                res["out"] = dewi_b200.sharded_self_join(xs, tau=tau, pair_cap=1 << 22, precision=precision)

        if world > 1:
            go()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            go()
            torch.cuda.synchronize()
            ms_t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
            ms = float(ms_t.item())
            found = torch.tensor([res["out"]["n_pairs"]], dtype=torch.int64, device=dev)
            dist.all_reduce(found)
            n_found = int(found.item())
        else:
            ms = timed(torch, go, 2, 1)
            n_found = res["out"]["n_pairs"]
        # algorithmic work of a self-join: N(N-1)/2 pair-dots of 2D flops each.  The kernel multiplies each
        # unordered pair of 256-row blocks once (circulant half of the block grid + the diagonal blocks):
        # executed = (T/2 + 1) / (T/2) of that on whole tiles; hi/lo planes ("fp32") issue three MMAs per tile
        alg = float(nrows) * (nrows - 1) * d
        tiles = -(-nrows // 256)
        executed = 2.0 * 256 * 256 * d * tiles * (1 + (tiles - 1) / 2 + (0.5 if tiles % 2 == 0 else 0.0))
        # planted pairs whose two rows both lie inside the rows this run joins (the fp32 run takes a prefix)
        planted = int(((dup < nrows) & (src < nrows)).sum().item()) if world == 1 else n // 100
        runs.append({"precision": precision, "rows": nrows, "ms": ms, "pairs_found": n_found, "pairs_planted": planted,
                     "pair_dots_per_s": nrows * (nrows - 1) / 2 / (ms / 1e3), "algorithmic_tflops": alg / (ms / 1e3) / 1e12,
                     "executed_tflops": executed * (3 if precision == "fp32" else 1) / (ms / 1e3) / 1e12})
        res.clear()
    del x
    torch.cuda.empty_cache()
    main = runs[0]
    line = {
        "metric": "pair-dots/sec (redundancy self-join, cosine threshold)", "unit": "pairs/s", "dtype": "bf16 planes, f32 accumulate",
        "value": main["pair_dots_per_s"], "n_gpus": world, "runs": runs,
        "config": {"workload": f"C5: {n} x {d} self-join, tau={tau}, row-sharded over {world} B200"
                               + (" (bounded; full size is 10M rows over 8 GPUs)" if world == 1 else ""),
                   "timing": "wall clock around one call incl. all-gather and row normalisation, max over ranks" if world > 1
                   else "CUDA events around the call (incl. row normalisation and plane split)",
                   "parity": "join reductions are PARITY UNPINNED (the reference defines only the dense product, redundancy.py:36-38)"},
        "roofline": {"bound": "tensor", "achieved": main["algorithmic_tflops"] / world, "peak": peaks["bf16_sustained"],
                     "unit": "TFLOP/s per GPU", "frac": main["algorithmic_tflops"] / world / peaks["bf16_sustained"],
                     "executed_frac": main["executed_tflops"] / world / peaks["bf16_sustained"]},
    }
    if own_pg:
        dist.barrier()
        dist.destroy_process_group()
    return line if rank == 0 or dist_ready else None


def run_extra(name, torch, dewi_b200, peaks, local_rank=0, dist_ready=False, rows=0):
    """One secondary row with clocks and its CPU baseline (bench.py's `extra[name]`)."""
    import bench_ref
    from bench import ClockSampler

    try:
        with ClockSampler(local_rank) as clocks:
            if name == "c2":
                line = run_c2(rows, torch, dewi_b200, peaks)
            elif name == "c4":
                line = run_c4(rows, torch, dewi_b200, peaks)
            else:
                line = run_c5(rows, torch, dewi_b200, peaks, dist_ready=dist_ready)
        if line is None:
            return None
        line["clocks"] = clocks.summary()
        if int(os.environ.get("RANK", "0")) == 0:
            threads = bench_ref.pin_blas(bench_ref.host_threads())
            if name == "c4":
                line["cpu_baseline"] = bench_ref.scorer_baseline(np, threads, bench_ref.load_reference())
            elif name == "c5":
                line["cpu_baseline"] = bench_ref.join_baseline(np, threads)
            # (c2's CPU baseline is the 1M-row point of the reference arm; bench.py fills it in)
        return line
    except Exception as exc:  # a secondary row must never take the headline down with it
        return {"error": repr(exc)[:500]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=["c2", "c4", "c5", "latency"])
    ap.add_argument("--rows", type=int, default=0)
    args = ap.parse_args()
    import torch

    import dewi_b200
    from bench import load_peaks

    peaks = load_peaks()
    real_stdout = os.dup(1)  # stdout carries exactly one JSON line (NCCL prints a banner there)
    os.dup2(2, 1)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if args.workload == "latency":
        line = quickstart_latency(torch, dewi_b200)
    else:
        line = run_extra(args.workload, torch, dewi_b200, peaks, local_rank, rows=args.rows)
    if line is None:
        return
    line["peak_source"] = peaks["source"]
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
