"""CPU legs of bench.py: the reference's own implementation of the hot path, timed on the box's host cores.

Protocol of BASELINE.md section 3 / SURVEY.md section 8(d):
  * what runs is the UNMODIFIED reference (`dewi.index.DewiIndex(use_ann=False)` -> `ExactIndex.search`,
    src/dewi/backends.py:414-481; `dewi.scorer.DewiScorer`, src/dewi/scorer.py:18-89) installed into
    `baseline/_ref` by scripts/vendor_reference.sh (`kind: "reference"`); when that directory is absent the
    numpy restatement under oracle/ is timed instead (`kind: "port"`);
  * the index is filled in bulk the way backends.py:403-411 leaves it (row-normalised fp32 matrix assigned to
    `_embeddings`; ids and payloads behind lazy containers so that 10M documents cost no 10M Python objects) --
    the `search` body that executes is the reference's, verbatim;
  * sizes N = 10K (C1), 1M (C2) and the largest N that fits host RAM (target 10M = 30.7 GB fp32); 10 warm-up
    queries, then median / p10 / p90 of >= 30 single queries (the reference has no batch API);
  * the 100M-row figure is a LINEAR EXTRAPOLATION from the largest measured N and is labelled so;
  * the BLAS pool is pinned explicitly (threadpoolctl + OMP/OPENBLAS_NUM_THREADS) so that a launcher which
    exports OMP_NUM_THREADS=1 (torch.distributed.run does) cannot shrink it; `cores` = threads actually used;
  * hnswlib / FAISS-CPU are attempted by import and reported "unavailable -- not installed" when absent.

Only bench.py imports this module; like oracle/ it is measurement infrastructure, not product code.
"""

from __future__ import annotations

import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REF_DIR = ROOT / "baseline" / "_ref"


DEFAULT_MAX_THREADS = 16


def host_threads(requested: int = 0) -> int:
    """Threads the CPU legs use: the cores this process is allowed on, capped at 16 unless `--cpu-threads` says
    otherwise.  The cap keeps `cores` identical across the 1/2/4/8-GPU runs (the pool hands a job more host cores
    with more GPUs: 16 at N = 1, 24 at N = 2) and costs the reference nothing: its hot statement is a memory-bound
    sgemv that saturates well before 16 threads (measured on the B200 host, 10M x 768 rows: 207 ms/query with 16
    threads, 244 ms with 24)."""
    if requested and requested > 0:
        return int(requested)
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, DEFAULT_MAX_THREADS))


def pin_env_threads(threads: int) -> None:
    """Must run BEFORE numpy is imported: OpenBLAS sizes its pool from these at load time."""
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(threads)


def pin_blas(threads: int) -> int:
    """Set the live BLAS pool and return the thread count it actually reports."""
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=threads, user_api="blas")
        got = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return int(max(got)) if got else 1
    except Exception:
        return 1


def cpu_model() -> str:
    try:
        for ln in Path("/proc/cpuinfo").read_text().splitlines():
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def ann_baselines() -> dict:
    """hnswlib / FAISS-CPU are the reference's ANN backends (backends.py:18-30, pyproject.toml:60-64)."""
    out = {}
    for name in ("hnswlib", "faiss"):
        try:
            __import__(name)
            out[name] = "importable"
        except Exception as exc:  # ImportError here; anything else is just as unusable
            out[name] = f"unavailable -- not installed ({type(exc).__name__})"
    return out


def load_reference():
    """The unmodified reference package from baseline/_ref, or None."""
    if not (REF_DIR / "dewi" / "index.py").exists():
        return None
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    import logging

    logging.getLogger("dewi.backends").setLevel(logging.ERROR)  # "HNSW not available" warnings -> ann_baselines()
    try:
        import dewi.index  # noqa: F401
        import dewi.scorer  # noqa: F401
        import dewi.types  # noqa: F401

        return sys.modules["dewi"]
    except Exception:
        return None


# ---- synthetic corpus on the host (N(0,1) rows, row-normalised in fp32) -------
def gen_corpus(np, n: int, dim: int, threads: int, seed: int = 7):
    """[n, dim] float32 unit rows, generated and normalised by `threads` workers (numpy's Generator releases
    the GIL), straight into one preallocated matrix."""
    emb = np.empty((n, dim), dtype=np.float32)
    step = 50_000
    starts = list(range(0, n, step))
    seeds = np.random.SeedSequence(seed).spawn(len(starts))

    def fill(i):
        lo = starts[i]
        hi = min(lo + step, n)
        rng = np.random.Generator(np.random.SFC64(seeds[i]))
        blk = emb[lo:hi]
        rng.standard_normal(out=blk, dtype=np.float32)
        nrm = np.sqrt(np.einsum("ij,ij->i", blk, blk))
        blk /= nrm[:, None]  # profile_index.py:55-56

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(fill, range(len(starts))))
    return emb


class LazyIds:
    """`doc_{i:08d}` (scripts/profile_index.py:52) without materialising N strings."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return f"doc_{int(i):08d}"


class LazyPayloads:
    """id -> Payload built on demand from the payload columns (ExactIndex.search touches 2k + k of them)."""

    def __init__(self, payload_cls, dewi, ht_mean, hi_mean):
        self.cls, self.dewi, self.ht, self.hi = payload_cls, dewi, ht_mean, hi_mean

    def __getitem__(self, doc_id):
        i = int(doc_id[4:])
        return self.cls(dewi=float(self.dewi[i]), ht_mean=float(self.ht[i]), hi_mean=float(self.hi[i]))


class ExactSearcher:
    """One CPU index over `emb`: the reference's DewiIndex(use_ann=False) when installed, else the oracle port."""

    def __init__(self, np, emb, k, eta, pref, ref_pkg):
        n = emb.shape[0]
        rng = np.random.RandomState(8)
        self.dewi = rng.beta(2, 2, n).astype(np.float32)
        self.ht = rng.gamma(2, 0.5, n).astype(np.float32)
        self.hi = rng.gamma(2, 0.3, n).astype(np.float32)
        self.k, self.eta, self.pref = k, eta, pref
        self.emb = emb
        self.kind = "reference" if ref_pkg is not None else "port"
        if ref_pkg is not None:
            from dewi.index import DewiIndex
            from dewi.types import Payload

            idx = DewiIndex(dim=emb.shape[1], space="cosine", use_ann=False)  # -> ExactIndex (index.py:50-51)
            be = idx._backend
            be._embeddings = emb                     # what build() leaves behind (backends.py:411)
            be._doc_ids = LazyIds(n)
            be._payloads = LazyPayloads(Payload, self.dewi, self.ht, self.hi)
            be._is_trained = True
            idx._built = True
            self.index = idx
        else:
            from oracle import search as osearch

            self.osearch = osearch
            self.ent = (self.ht.astype(np.float64) + self.hi.astype(np.float64)) * 0.5

    def search(self, q):
        if self.kind == "reference":
            return self.index.search(q, k=self.k, eta=self.eta, entropy_pref=self.pref)
        return self.osearch.exact_search(self.emb, self.dewi, self.ent, q, self.k, self.eta, self.pref, True)


def time_queries(np, searcher, queries, warmup: int, timed: int):
    """Per-query wall times (ms) of `timed` single queries after `warmup` untimed ones."""
    qi = 0
    for _ in range(warmup):
        searcher.search(queries[qi % len(queries)])
        qi += 1
    out = []
    for _ in range(timed):
        q = queries[qi % len(queries)]
        qi += 1
        t0 = time.perf_counter()
        searcher.search(q)
        out.append((time.perf_counter() - t0) * 1e3)
    return np.asarray(out)


def largest_rows(dim: int, want: int) -> int:
    """Largest corpus (multiple of 1M rows, at most `want`) whose fp32 matrix fits in half of the free host RAM."""
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 16 << 30
    fit = int(avail * 0.5 // (dim * 4))
    n = min(want, fit)
    return max(1_000_000, n // 1_000_000 * 1_000_000)


def search_points(np, dim, k, eta, pref, threads, sizes, ref_pkg, keep_last=False):
    """one measurement point per corpus size."""
    rng = np.random.RandomState(99)
    queries = rng.standard_normal((64, dim)).astype(np.float32)
    points, last = [], None
    for n in sizes:
        t0 = time.perf_counter()
        emb = gen_corpus(np, n, dim, threads)
        gen_s = time.perf_counter() - t0
        s = ExactSearcher(np, emb, k, eta, pref, ref_pkg)
        # bound the point to ~15 s of queries: 10 warm-up + 30 timed unless a query takes longer than 0.4 s
        probe = time_queries(np, s, queries, 2, 1)[0]
        timed = 30 if probe < 400 else max(10, int(12_000 / probe))
        ms = time_queries(np, s, queries, 10 if probe < 400 else 3, timed)
        points.append({"rows": int(n), "ms_median": float(np.median(ms)), "ms_p10": float(np.percentile(ms, 10)),
                       "ms_p90": float(np.percentile(ms, 90)), "queries": int(len(ms)), "gen_s": round(gen_s, 2),
                       "stream_gbs": n * dim * 4 / (float(np.median(ms)) / 1e3) / 1e9})
        last = s
        if not (keep_last and n == sizes[-1]):
            del emb, s
            last = None
    return points, last, queries


def reference_search_line(args, metric: str, workload_config):
    """The `--impl reference` line (rank 0 only; other ranks exit without work)."""
    import numpy as np

    threads = pin_blas(host_threads(args.cpu_threads))
    ref_pkg = load_reference()
    n_big = largest_rows(args.dim, args.cpu_max_rows)
    sizes = [s for s in (10_000, 1_000_000) if s < n_big] + [n_big]
    points, searcher, queries = search_points(np, args.dim, args.k, args.eta, args.entropy_pref, threads, sizes, ref_pkg,
                                              keep_last=True)
    big = points[-1]
    # steps: a bounded sample each -- queries_per_step single queries on the largest corpus, sized so that
    # `steps + warmup` steps take about a minute
    per_q = big["ms_median"] / 1e3
    qps_step = int(max(1, min(8, 60.0 / max(1e-9, (args.steps + args.warmup) * per_q))))
    qi = 0
    for _ in range(args.warmup * qps_step):
        searcher.search(queries[qi % 64])
        qi += 1
    t0 = time.perf_counter()
    for _ in range(args.steps * qps_step):
        searcher.search(queries[qi % 64])
        qi += 1
    dt = time.perf_counter() - t0
    qps_sample = args.steps * qps_step / dt
    scale = n_big / args.rows
    value = qps_sample * scale
    # sanity: the port and the reference agree on the result ids (same inputs)
    agree = None
    if searcher.kind == "reference":
        from oracle import search as osearch

        ent = (searcher.ht.astype(np.float64) + searcher.hi.astype(np.float64)) * 0.5
        got = [int(r[0][4:]) for r in searcher.search(queries[0])]
        want, _ = osearch.exact_search(searcher.emb, searcher.dewi, ent, queries[0], args.k, args.eta, args.entropy_pref, True)
        agree = bool(got == want.tolist())
    cpu = {
        "value": value, "unit": "queries/s", "cores": threads, "kind": searcher.kind,
        "sample": (f"{'unmodified reference DewiIndex(use_ann=False).search' if searcher.kind == 'reference' else 'oracle port of ExactIndex.search'}"
                   f" (numpy/OpenBLAS sgemv fp32, {threads} BLAS threads, {cpu_model()}, {os.cpu_count()} logical cores) on a "
                   f"{n_big}-row x {args.dim} corpus: {args.steps * qps_step} single queries at {1e3 / qps_sample:.1f} ms/query; "
                   f"EXTRAPOLATED linearly x{scale:.4g} to {args.rows} rows (100M x 768 fp32 = 307 GB cannot be held on the host)"),
        "points": points, "extrapolated": True, "ann_baselines": ann_baselines(),
        "port_matches_reference_ids": agree,
    }
    return {
        "impl": "reference", "metric": metric, "value": value, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, args.gpus, args.rows),
        "cpu_baseline": cpu, "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ---- secondary rows (C4 scorer, C5 redundancy product): CPU baselines for bench.py's `extra` keys -------
def scorer_baseline(np, threads: int, ref_pkg, n_ref: int = 200_000, n_vec: int = 2_000_000):
    """`DewiScorer.fit_stats` + per-row `score` (scorer.py:18-75; a Python loop in the reference, README.md:101-110)
    on `n_ref` rows, and the vectorised numpy port on `n_vec` rows."""
    from oracle import scorer as oscorer

    rng = np.random.RandomState(21)
    hi = np.array([10, 15, 5, 8, 1, 1, 0.2], dtype=np.float32)
    sig = (rng.random_sample((7, n_vec)).astype(np.float32) * hi[:, None])
    out = {"cores": 1, "threads_note": "the reference scorer is single-threaded Python / numpy median"}
    cols = {k: sig[i] for i, k in enumerate(oscorer.SIGNAL_KEYS)}
    t0 = time.perf_counter()
    med, mad = oscorer.robust_fit(cols)
    t1 = time.perf_counter()
    oscorer.score_rows(cols, med, mad)
    t2 = time.perf_counter()
    out["port"] = {"rows": n_vec, "fit_rows_per_s": n_vec / (t1 - t0), "score_rows_per_s": n_vec / (t2 - t1),
                   "fit_plus_score_rows_per_s": n_vec / (t2 - t0)}
    if ref_pkg is not None:
        from dewi.scorer import DewiScorer

        rows = [{k: float(sig[i, r]) for i, k in enumerate(oscorer.SIGNAL_KEYS)} for r in range(n_ref)]
        s = DewiScorer()
        t0 = time.perf_counter()
        s.fit_stats(rows)
        t1 = time.perf_counter()
        for r in rows:
            s.score(r)
        t2 = time.perf_counter()
        out["reference"] = {"rows": n_ref, "fit_rows_per_s": n_ref / (t1 - t0), "score_rows_per_s": n_ref / (t2 - t1),
                            "fit_plus_score_rows_per_s": n_ref / (t2 - t0)}
        out["kind"], out["value"] = "reference", out["reference"]["fit_plus_score_rows_per_s"]
    else:
        out["kind"], out["value"] = "port", out["port"]["fit_plus_score_rows_per_s"]
    out["unit"] = "rows/s"
    out["sample"] = (f"{'unmodified reference DewiScorer.fit_stats + per-row score on ' + str(n_ref) + ' rows; ' if ref_pkg else ''}"
                     f"vectorised numpy port on {n_vec} rows")
    return out


def join_baseline(np, threads: int, rows: int = 20_000, dim: int = 512):
    """`F.normalize(T) @ F.normalize(I).T` (redundancy.py:36-38) as one fp32 block product: pair-dots/s."""
    from oracle import redundancy as ored

    rng = np.random.RandomState(44)
    x = rng.standard_normal((rows, dim)).astype(np.float32)
    ored.cross_modal_similarity(x[:2000], x[:2000])
    t0 = time.perf_counter()
    s = ored.cross_modal_similarity(x, x)
    dt = time.perf_counter() - t0
    del s
    return {"value": rows * rows / dt, "unit": "pair-dots/s", "cores": threads, "kind": "port",
            "sample": f"numpy restatement of normalize(T) @ normalize(I).T on a {rows} x {rows} x {dim} fp32 block "
                      f"({threads} BLAS threads); the reference's estimator needs CLIP weights (not available offline)"}
