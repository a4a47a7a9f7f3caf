#!/usr/bin/env python
"""bench_paths.py -- the other hot-path rows of SURVEY.md section 8 (BASELINE.json configs C2, C4, C5), one JSON
line per workload.  `bench.py` stays the headline (C3); this script is how DESIGN.md's numbers for the
fp32 exact search, the scorer and the redundancy join are produced.

    python bench_paths.py --workload c2        # 1M x 768 fp32 exact DEWI-re-ranked top-10, B = 1..4096
    python bench_paths.py --workload c4        # fit_stats + score over 100M Signals rows
    python bench_paths.py --workload c5        # redundancy self-join, 512-d, cosine threshold (bounded size)
"""

from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from bench import ClockSampler, load_peaks  # noqa: E402


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


def run_c2(args, torch, dewi_b200, peaks):
    n, d, k = args.rows or 1_000_000, 768, 10
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    emb = torch.randn((n, d), generator=g, device=dev)
    ix = dewi_b200.CudaIndex(d, dtype="fp32", device=0)
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
        ix.set_profiling(False)
        if b <= ridge:
            ach, peak, unit, bound = n * d * 4 / (kms / 1e3) / 1e9, peaks["hbm_gbs"], "GB/s", "hbm"
        else:
            ach, peak, unit, bound = 2.0 * b * n * d / (kms / 1e3) / 1e12, peaks["bf16_sustained"], "TFLOP/s", "tensor"
        out.append({"batch": b, "value": b / (ms / 1e3), "ms_per_step": ms, "kernel_ms": kms, "bound": bound, "achieved": ach,
                    "peak": peak, "unit": unit, "frac": ach / peak})
    return {"metric": "queries/sec (1M x 768 fp32 exact, k=10, DEWI re-rank)", "unit": "queries/s", "dtype": "f32 (bf16 hi/lo planes)",
            "config": {"workload": f"C2: {n} x {d} fp32 exact search, 1 B200", "l2": "corpus 3 GB >> L2"}, "batches": out}


def run_c4(args, torch, dewi_b200, peaks):
    from oracle import scorer as oscorer

    n = args.rows or 100_000_000
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(21)
    hi = torch.tensor([10, 15, 5, 8, 1, 1, 0.2], device=dev).view(7, 1)
    sig = torch.rand((7, n), generator=g, device=dev) * hi  # README.md:83-91 ranges
    s = dewi_b200.DewiScorer()
    fit_ms = timed(torch, lambda: s.fit_stats_columns(sig), 3, 1)
    score_ms = timed(torch, lambda: s.score_batch(sig), 5, 2)
    fit_bytes, score_bytes = 2 * n * 7 * 4, n * (7 * 4 + 4)
    # CPU: the reference's own per-row path is a Python loop (6-10 us/row); time the vectorised oracle port
    m = min(n, 2_000_000)
    cols = {k: sig[i, :m].cpu().numpy() for i, k in enumerate(oscorer.SIGNAL_KEYS)}
    t0 = time.perf_counter()
    med, mad = oscorer.robust_fit(cols)
    t1 = time.perf_counter()
    oscorer.score_rows(cols, med, mad)
    t2 = time.perf_counter()
    return {
        "metric": "rows/sec (fit_stats + score over Signals rows)", "unit": "rows/s", "dtype": "f32 data, f64 arithmetic",
        "value": n / ((fit_ms + score_ms) / 1e3),
        "config": {"workload": f"C4: {n} Signals rows x 7 columns, 1 B200", "l2": "2.8 GB columns >> L2"},
        "fit_stats": {"ms": fit_ms, "rows_per_s": n / (fit_ms / 1e3),
                      "roofline": {"bound": "hbm", "achieved": fit_bytes / (fit_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
                                   "unit": "GB/s", "frac": fit_bytes / (fit_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                                   "algorithmic_bytes": fit_bytes}},
        "score": {"ms": score_ms, "rows_per_s": n / (score_ms / 1e3),
                  "roofline": {"bound": "hbm", "achieved": score_bytes / (score_ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
                               "unit": "GB/s", "frac": score_bytes / (score_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
                               "algorithmic_bytes": score_bytes}},
        "cpu_baseline": {"kind": "port", "sample": f"vectorised numpy oracle on {m} rows", "fit_rows_per_s": m / (t1 - t0),
                         "score_rows_per_s": m / (t2 - t1)},
    }


def run_c5(args, torch, dewi_b200, peaks):
    """1 process: bounded self-join, bf16 and hi/lo planes.  Under torchrun: the row-sharded self-join
    (each rank owns a block of rows, all-gather once, join its rows against all rows), bf16 planes."""
    import os

    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d, tau = args.rows or (1_000_000 if world == 1 else 10_000_000), 512, 0.9
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
    for precision, rows in configs:
        res = {}
        xs = x[: rows if world == 1 else hi - lo]

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
        alg = float(rows) * (rows - 1) * d
        tiles = -(-rows // 256)
        executed = 2.0 * 256 * 256 * d * tiles * (1 + (tiles - 1) / 2 + (0.5 if tiles % 2 == 0 else 0.0))
        runs.append({"precision": precision, "rows": rows, "ms": ms, "pairs_found": n_found,
                     "pair_dots_per_s": rows * (rows - 1) / 2 / (ms / 1e3), "algorithmic_tflops": alg / (ms / 1e3) / 1e12,
                     "executed_tflops": executed * (3 if precision == "fp32" else 1) / (ms / 1e3) / 1e12})
    main = runs[0]
    line = {
        "metric": "pair-dots/sec (redundancy self-join, cosine threshold)", "unit": "pairs/s", "dtype": "bf16 planes, f32 accumulate",
        "value": main["pair_dots_per_s"], "n_gpus": world, "runs": runs,
        "config": {"workload": f"C5: {n} x {d} self-join, tau={tau}, row-sharded over {world} B200"
                               + (" (bounded; full size is 10M rows over 8 GPUs)" if world == 1 else ""),
                   "timing": "wall clock around one call incl. all-gather and row normalisation, max over ranks" if world > 1
                   else "CUDA events"},
        "roofline": {"bound": "tensor", "achieved": main["algorithmic_tflops"] / world, "peak": peaks["bf16_sustained"],
                     "unit": "TFLOP/s per GPU", "frac": main["algorithmic_tflops"] / world / peaks["bf16_sustained"],
                     "executed_frac": main["executed_tflops"] / world / peaks["bf16_sustained"]},
    }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return None
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True, choices=["c2", "c4", "c5"])
    ap.add_argument("--rows", type=int, default=0)
    args = ap.parse_args()
    import torch

    import dewi_b200

    peaks = load_peaks()
    import os

    real_stdout = os.dup(1)  # stdout carries exactly one JSON line (NCCL prints a banner there)
    os.dup2(2, 1)

    with ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) as clocks:
        line = {"c2": run_c2, "c4": run_c4, "c5": run_c5}[args.workload](args, torch, dewi_b200, peaks)
    if line is None:
        return
    line["clocks"] = clocks.summary()
    line["peak_source"] = peaks["source"]
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
