#!/usr/bin/env python
"""bench.py -- queries/sec of DEWI-re-ranked top-10 search over a 100M x 768 bf16 corpus (BASELINE.json).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # the reference's CPU path (oracle port)

A step = one pass of the hot path (sweep -> candidate select -> [all-gather] -> DEWI re-rank -> top-k)
over one batch of B synthetic queries.  The corpus (N rows in total, fixed as the GPU count grows:
strong scaling) is row-sharded across the ranks.  Rank 0 prints ONE JSON line.

value      device-timed throughput, queries and corpus already resident in HBM
e2e        same metric through the public API with HOST (pinned) query buffers and HOST results
roofline   the sweep kernel: algorithmic corpus bytes per launch / its CUDA-event duration vs the
           measured HBM peak (B <= ridge) or algorithmic flops vs the measured bf16 peak (B > ridge)
cpu_baseline  the oracle port of ExactIndex.search (numpy/OpenBLAS) on a bounded row sample, rank 0
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CHUNK_ROWS = 500_000  # corpus rows are generated in globally numbered chunks -> same corpus for every GPU count
METRIC = "queries/sec (100M x 768 bf16, k=10, DEWI re-rank)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--eta", type=float, default=0.3)
    ap.add_argument("--entropy-pref", type=float, default=0.5)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--sweep", default="1,8,256,1024,4096", help="extra batch sizes reported in `batch_sweep` ('' = none)")
    ap.add_argument("--cpu-sample-rows", type=int, default=500_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "push", "nccl"],
                    help="multi-GPU candidate exchange: fused peer stores (symmetric memory) or one NCCL all-gather per batch")
    return ap.parse_args()


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ---- synthetic inputs (scripts/profile_index.py:34-72) -------------------------
def gen_chunk(torch, chunk: int, rows: int, dim: int, device):
    """Rows [chunk*CHUNK_ROWS, +rows) of the corpus: N(0,1), unnormalised fp32, + payload columns."""
    g = torch.Generator(device=device)
    g.manual_seed(42_000_003 + chunk)
    emb = torch.randn((rows, dim), generator=g, device=device, dtype=torch.float32)
    u = torch.rand((4, rows), generator=g, device=device, dtype=torch.float32)
    # dewi ~ Beta(2,2) as the median of three uniforms; entropies ~ Gamma(2, .) as sums of two exponentials
    dewi = u[:3].median(dim=0).values
    e = -torch.log(torch.rand((4, rows), generator=g, device=device, dtype=torch.float32).clamp_min(1e-12))
    ht_mean = (e[0] + e[1]) * 0.5
    hi_mean = (e[2] + e[3]) * 0.3
    ent = ((ht_mean.double() + hi_mean.double()) * 0.5).float()  # backends.py:458
    return emb, dewi.contiguous(), ent.contiguous()


def gen_queries(torch, seed: int, b: int, dim: int):
    g = torch.Generator()
    g.manual_seed(43_000_000 + seed)
    return torch.randn((b, dim), generator=g, dtype=torch.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 5.0:  # sampling is live before the timed region starts
                time.sleep(0.01)
            self.lines.clear()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.03)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---- the reference's CPU path (oracle port of ExactIndex.search) --------------------------------------
def cpu_reference_qps(args, steps: int, warmup: int, queries_per_step: int):
    """Times oracle.search.exact_search (numpy sgemv + argpartition + blend, backends.py:414-481) on a
    bounded row sample and scales queries/s linearly to the full corpus (the path is one linear
    stream over N rows; 100M x 768 fp32 = 307 GB cannot be held on the host)."""
    from oracle import search as osearch

    try:
        from threadpoolctl import threadpool_info

        threads = max([p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        threads = os.cpu_count() or 1
    import torch

    n, d = args.cpu_sample_rows, args.dim
    g = torch.Generator()
    g.manual_seed(7)
    emb = torch.randn((n, d), generator=g, dtype=torch.float32)
    emb = torch.nn.functional.normalize(emb, dim=1).numpy()
    rng = np.random.RandomState(8)
    dewi = rng.beta(2, 2, n)
    ent = rng.gamma(2, 0.4, n)
    qs = gen_queries(torch, 99, queries_per_step * (steps + warmup), d).numpy()
    qi = 0
    for _ in range(warmup):
        for _ in range(queries_per_step):
            osearch.exact_search(emb, dewi, ent, qs[qi], args.k, args.eta, args.entropy_pref, True)
            qi += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        for _ in range(queries_per_step):
            osearch.exact_search(emb, dewi, ent, qs[qi], args.k, args.eta, args.entropy_pref, True)
            qi += 1
    dt = time.perf_counter() - t0
    qps_sample = steps * queries_per_step / dt
    scale = n / args.rows
    return {
        "value": qps_sample * scale, "unit": "queries/s", "cores": int(threads), "kind": "port",
        "sample": (f"oracle port of ExactIndex.search (numpy/OpenBLAS sgemv, fp32) on a {n}-row x {d} sample, "
                   f"{steps * queries_per_step} queries at {1e3 / qps_sample:.2f} ms/query; scaled x{scale:.4g} to {args.rows} "
                   f"rows (linear extrapolation); host has {os.cpu_count()} logical cores"),
        "ms_per_step_sample": dt / steps * 1e3,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    qps = 8
    res = cpu_reference_qps(args, args.steps, args.warmup, qps)
    value = res["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res.pop("ms_per_step_sample"), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus, args.rows),
        "cpu_baseline": res,
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, rows):
    return {
        "workload": f"C3: {rows} x {args.dim} {args.dtype} corpus row-sharded over {world} B200, top-{args.k} DEWI re-rank, "
                    f"query batch {args.batch}",
        "rows": rows, "dim": args.dim, "k": args.k, "batch": args.batch, "eta": args.eta, "entropy_pref": args.entropy_pref,
        "shards": world, "l2": "corpus shard >> 126 MB L2 (streamed once per step); no flush needed",
    }


# ---- our arm -------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import dewi_b200
    from dewi_b200 import shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: anything libraries print there meanwhile (NCCL's version
    # banner at communicator creation, for one) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    peaks = load_peaks()

    # ---- corpus: fit the requested rows into HBM (state it if it had to shrink) ----------------------
    rows = args.rows
    bytes_per_row = args.dim * (2 if args.dtype == "bf16" else 8) + 8
    free, total = torch.cuda.mem_get_info(device)
    budget = free - (6 << 30)  # staging chunk, workspaces, NCCL
    lo, hi = shard_range(rows, world, rank, align=CHUNK_ROWS)
    shrunk = None
    if (hi - lo) * bytes_per_row > budget:
        per_rank = int(budget // bytes_per_row // CHUNK_ROWS * CHUNK_ROWS)
        shrunk = f"requested {rows} rows do not fit {world} x {total >> 30} GiB; using {per_rank * world}"
        rows = per_rank * world
        lo, hi = shard_range(rows, world, rank, align=CHUNK_ROWS)
    if world > 1:
        agree = torch.tensor([rows], dtype=torch.int64, device=device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if int(agree.item()) != rows:
            rows = int(agree.item())
            lo, hi = shard_range(rows, world, rank, align=CHUNK_ROWS)

    t_build = time.perf_counter()
    if world > 1:
        index = dewi_b200.ShardedDewiIndex(args.dim, dtype=args.dtype, device=local_rank, exchange=args.exchange)
        local = index.local
    else:
        index = local = dewi_b200.CudaIndex(args.dim, dtype=args.dtype, device=local_rank)
    local.reserve(hi - lo)
    done = lo
    while done < hi:
        m = min(CHUNK_ROWS, hi - done)
        emb, dewi_c, ent_c = gen_chunk(torch, done // CHUNK_ROWS, m, args.dim, device)
        local.add_batch(None, emb, normalized=False)  # device-side normalise + bf16 cast (prep kernel)
        local.set_payload_columns(dewi_c, ent_c, offset=done - lo)
        done += m
        del emb, dewi_c, ent_c
    index.build()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build

    n_qsets = 4
    q_host = [gen_queries(torch, s, args.batch, args.dim).pin_memory() for s in range(n_qsets)]
    q_dev = [q.to(device) for q in q_host]

    def step_device(i, b_queries=None):
        q = b_queries if b_queries is not None else q_dev[i % n_qsets]
        return index.search_batch(q, k=args.k, eta=args.eta, entropy_pref=args.entropy_pref)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- value: device-resident queries --------------------------------------------------------------
    # The library brackets each sweep launch with an event pair on the same stream (ring of 64, no
    # synchronisation added), so the roofline figure comes from the very steps that are timed.
    local.set_profiling(True)
    with ClockSampler(local_rank) as clocks:
        total_ms = timed(step_device, args.steps, args.warmup)
    ms_per_step = total_ms / args.steps
    value = args.batch * args.steps / (total_ms / 1e3)
    launches_per_step = local.last_launches() + (1 if world > 1 else 0)

    # ---- roofline: the sweep kernel alone, live over the timed region ----------------------------------
    n_prof = min(args.steps, 64)
    sweep_ms = [local.sweep_ms(i) for i in range(n_prof)]
    kind = sweep_ms[0][1]
    sweep = float(np.mean([m for m, _ in sweep_ms]))
    local.set_profiling(False)
    elem = 2 if args.dtype == "bf16" else 4
    shard_bytes = (hi - lo) * args.dim * elem
    flops = 2.0 * args.batch * (hi - lo) * args.dim
    ridge = peaks["bf16_sustained"] * 1e12 * elem / (2 * peaks["hbm_gbs"] * 1e9)

    def roofline_for(batch, ms, flops_):
        if batch <= ridge:
            ach = shard_bytes / (ms / 1e3) / 1e9
            return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]}
        ach = flops_ / (ms / 1e3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_sustained"]}

    roofline = roofline_for(args.batch, sweep, flops)
    # dram__bytes_read+write of the main sweep launch from the committed `ncu --set full` capture
    # (profiles/r1_v7b_prof_b64_12M5rows_raw.csv: 19.2006 GB read + 6.3 MB written for a 19.2 GB shard),
    # scaled to this shard; only the single-query-block kernel (B <= 128) was captured at that ratio
    traffic = shard_bytes * 1.00036 if (args.batch <= 128 and kind == "tcgen05") else None
    roofline.update({"traffic": traffic, "traffic_source": "ncu capture of the same kernel at 12.5M rows, scaled by shard size"
                     if traffic else None, "kernel": f"search_tc_kernel ({kind}) incl. its sample pre-pass", "kernel_ms": sweep,
                     "algorithmic_bytes_per_launch": shard_bytes, "peak_source": peaks["source"],
                     "kernel_share_of_step": sweep / ms_per_step})

    # ---- e2e: host (pinned) queries in, host results out, through the public API ---------------------
    def step_e2e(i):
        qh = q_host[i % n_qsets]
        if world > 1:
            ids, sc = index.search_batch(qh.to(device, non_blocking=True), k=args.k, eta=args.eta, entropy_pref=args.entropy_pref)
            return ids.cpu(), sc.cpu()
        return index.search_batch(qh.numpy(), k=args.k, eta=args.eta, entropy_pref=args.entropy_pref)

    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(i)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e = {"value": args.batch * args.steps / float(e2e_s.item()), "unit": "queries/s",
           "h2d_bytes_per_step": args.batch * args.dim * 4, "d2h_bytes_per_step": args.batch * args.k * 12,
           "ms_per_step": float(e2e_s.item()) / args.steps * 1e3}

    # ---- other batch sizes (device-timed, few steps) ---------------------------------------------------
    batch_sweep = []
    for b in [int(x) for x in args.sweep.split(",") if x]:
        if b == args.batch:
            continue
        qb = gen_queries(torch, 100 + b, b, args.dim).to(device)
        steps_b = 3 if b >= 1024 else 5
        local.set_profiling(True)
        ms_b = timed(lambda i: step_device(i, qb), steps_b, 2) / steps_b
        kms = float(np.mean([local.sweep_ms(i)[0] for i in range(steps_b)]))
        local.set_profiling(False)
        r = roofline_for(b, kms, 2.0 * b * (hi - lo) * args.dim)
        batch_sweep.append({"batch": b, "value": b / (ms_b / 1e3), "ms_per_step": ms_b, "kernel_ms": kms, "bound": r["bound"],
                            "achieved": r["achieved"], "frac": r["frac"], "unit": r["unit"]})

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_reference_qps(args, steps=4, warmup=1, queries_per_step=8)
            cpu.pop("ms_per_step_sample", None)
        cfg = workload_config(args, world, rows)
        cfg.update({"build_s": round(t_build, 2), "parallelism": f"row-shard x{world}", "note": shrunk})
        if world > 1:
            cfg["exchange"] = ("fused peer stores over NVLink (symmetric memory), no collective on the search path"
                               if index.exchange == "push" else "one NCCL all_gather_into_tensor per batch")
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": cfg, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks.summary(), "batch_sweep": batch_sweep,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
