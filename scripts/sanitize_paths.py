"""Every kernel family once, at sizes compute-sanitizer finishes in seconds (scripts/sanitize.sh runs this under
`--tool memcheck`, `synccheck` and, for the kernels without asynchronous-proxy traffic, `racecheck`).

Not a parity test (tests/ hold those): results are only sanity-checked against a plain numpy cosine top-k so that a
kernel that silently did nothing cannot pass.  Cases: rows-on-M sweep (B <= 64), queries-on-M sweep (B <= 128),
CTA-pair sweep (B > 128, incl. the staged form), certified fp32 sweep and the full hi/lo product, CUDA-core sweep
(l2 space), full-scope blend, fused tail and separate re-rank, fit_stats (windowed and plain radix select), score,
local weights, dense similarity, tensor-core and CUDA-core joins, cluster labelling, bulk export."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dewi_b200  # noqa: E402
from dewi_b200 import _native  # noqa: E402

rng = np.random.RandomState(3)
done, failed = [], []


def case(fn):
    """Run one case; a failure is recorded and the remaining cases still run."""
    def wrapped(name, *a, **kw):
        try:
            fn(name, *a, **kw)
            done.append(name)
        except Exception as exc:  # noqa: BLE001
            failed.append(f"{name}: {type(exc).__name__}: {exc}")
    return wrapped


def cosine_topk(emb, q, k):
    e = emb / np.linalg.norm(emb, axis=1, keepdims=True)
    qq = q / np.linalg.norm(q, axis=1, keepdims=True)
    s = qq @ e.T
    idx = np.argsort(-s, axis=1)[:, :k]
    return idx, np.take_along_axis(s, idx, axis=1)


@case
def search_case(name, n, d, b, dtype, flags=0, space="cosine", scope="candidates", k=10):
    emb = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((b, d)).astype(np.float32)
    ix = dewi_b200.CudaIndex(d, space=space, dtype=dtype, device=0, rerank_scope=scope)
    ix.add_batch(None, emb, normalized=False)
    ix.set_payload_columns(rng.rand(n).astype(np.float32), rng.rand(n).astype(np.float32))
    ix.build()
    ids, sc = ix.search_batch(q, k=k, eta=0.0, entropy_pref=0.0, flags=flags)       # host in / host out
    ids_d, sc_d = ix.search_batch(torch.from_numpy(q).cuda(), k=k, eta=0.3, entropy_pref=0.5, flags=flags)
    torch.cuda.synchronize()
    assert ids.shape == (b, k) and int(ids.min()) >= 0 and int(ids.max()) < n and np.all(np.diff(sc, axis=1) <= 0)
    assert int(ids_d.min()) >= 0 and int(ids_d.max()) < n
    if space == "cosine":
        ridx, rsc = cosine_topk(emb, q, k)
        tol = 2e-2 if dtype == "bf16" else 1e-5
        assert np.max(np.abs(sc - rsc)) <= tol, (name, np.max(np.abs(sc - rsc)))
        if dtype == "fp32":
            assert np.mean(ids == ridx) > 0.99, name
    rows = ix.export_rows(0, min(n, 64)) if hasattr(ix, "export_rows") else None
    del ix, rows


TC, SIMT, NO_PAIR = _native.FLAG_FORCE_TC, _native.FLAG_FORCE_SIMT, getattr(_native, "FLAG_NO_PAIR", 1 << 6)
NO_CERT, FORCE_CERT = getattr(_native, "FLAG_NO_CERT", 1 << 9), getattr(_native, "FLAG_FORCE_CERT", 1 << 10)
NO_ROWS = getattr(_native, "FLAG_NO_ROWS_ON_M", 1 << 11)

search_case("rows-on-M bf16 B=1", 6000, 128, 1, "bf16", TC)
search_case("rows-on-M bf16 B=33", 5000, 192, 33, "bf16", TC)
search_case("queries-on-M bf16 B=40", 6000, 128, 40, "bf16", TC | NO_ROWS)
search_case("queries-on-M bf16 B=100", 7001, 128, 100, "bf16", TC)
search_case("pair sweep bf16 B=300", 9000, 128, 300, "bf16", TC)
search_case("staged pair sweep bf16 B=1100", 40000, 64, 1100, "bf16", TC)
search_case("certified fp32 B=8", 6000, 128, 8, "fp32", TC | FORCE_CERT)
search_case("certified fp32 pair B=260", 8000, 128, 260, "fp32", TC | FORCE_CERT)
search_case("hi/lo product fp32 B=8", 6000, 128, 8, "fp32", TC | NO_CERT)
search_case("hi/lo product fp32 pair B=200", 6000, 128, 200, "fp32", TC | NO_CERT)
search_case("CUDA-core fp32 odd dim", 3000, 100, 5, "fp32", 0)
search_case("CUDA-core l2", 3000, 64, 5, "fp32", SIMT, space="l2")
search_case("full scope bf16 B=16", 6000, 128, 16, "bf16", 0, scope="full")
search_case("full scope fp32 B=4", 4000, 128, 4, "fp32", 0, scope="full")
search_case("k=64 bf16 B=3", 6000, 128, 3, "bf16", 0, k=64)

# scorer: windowed selection (large n), plain radix select (small n), ties, score, local weights
@case
def scorer_case(name, n):
    sig = np.stack([rng.uniform(0, hi, n) for hi in (10, 15, 5, 8, 1, 1, 0.2)]).astype(np.float32)
    if n == 5000:
        sig[2] = np.round(sig[2])  # heavy ties
    sc = dewi_b200.DewiScorer()
    sc.fit_stats_columns(sig)
    med = {k: float(np.median(sig[i])) for i, k in enumerate(sc.stats.medians)}
    assert all(sc.stats.medians[k] == med[k] for k in med), (n, sc.stats.medians, med)
    out = sc.score_batch(sig).cpu().numpy()
    assert out.shape == (n,) and np.all((out > 0) & (out < 1))


@case
def local_weights_case(name):
    w = dewi_b200.local_weights_from_surprisal(rng.gamma(2.0, 1.0, 4096).astype(np.float32))
    assert w.shape == (4096,) and np.all(np.isfinite(w)) and np.all(w > 0)


# redundancy: dense product, tensor-core self-join, cross join, CUDA-core join, clusters
@case
def dense_case(name):
    t, i = rng.standard_normal((300, 96)).astype(np.float32), rng.standard_normal((200, 96)).astype(np.float32)
    sim = np.asarray(dewi_b200.cross_modal_similarity(t, i))
    ref = (t / np.linalg.norm(t, axis=1, keepdims=True)) @ (i / np.linalg.norm(i, axis=1, keepdims=True)).T
    assert np.max(np.abs(sim - ref)) <= 1e-5


x = rng.standard_normal((1700, 128)).astype(np.float32)
x[5::40] = x[4::40][: len(x[5::40])] * 1.2 + 0.02 * rng.standard_normal((len(x[5::40]), 128)).astype(np.float32)


@case
def join_case(name, force, prec):
    out = dewi_b200.redundancy_join(x, tau=0.92, force=force, precision=prec)
    assert out["n_pairs"] >= 40, out["n_pairs"]
    labels = dewi_b200.cluster_pairs(out["pairs_i"], out["pairs_j"], x.shape[0])
    assert len(dewi_b200.clusters_from_labels(labels)) > 0


@case
def cross_join_case(name):
    out = dewi_b200.redundancy_join(x[:900], x[700:], tau=0.92, force="tc")
    assert out["n_pairs"] >= 200, out["n_pairs"]


for n in (300_000, 5000, 3):
    scorer_case(f"fit_stats/score n={n}", n)
local_weights_case("local weights")
dense_case("dense similarity")
for force, prec in (("tc", "fp32"), ("tc", "bf16"), ("simt", "fp32")):
    join_case(f"self-join {force}/{prec} + clusters", force, prec)
cross_join_case("cross join")
torch.cuda.synchronize()
print(f"sanitize_paths: {len(done)} cases ok: " + "; ".join(done))
for f in failed:
    print("FAILED", f)
sys.exit(1 if failed else 0)
