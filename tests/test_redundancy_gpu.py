"""Redundancy similarity pass: dense product vs torch's F.normalize + matmul fixtures
(redundancy.py:36-38); thresholded join vs the reductions of the reference's own dense matrix on the fixtures, and vs this
repository's CPU restatement at larger sizes (the reductions themselves have no reference definition: parity unpinned)."""

import numpy as np
import pytest

import dewi_b200
from oracle import redundancy as ored

from _util import GOLD

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["t37_i53_d512", "t64_i64_d64"])
def test_golden_dense_similarity(name):
    g = np.load(GOLD / f"redundancy_{name}.npz")
    sim = dewi_b200.cross_modal_similarity(g["tfeat"], g["ifeat"])
    assert isinstance(sim, np.ndarray) and sim.shape == g["sim"].shape and sim.dtype == np.float32
    np.testing.assert_allclose(sim, g["sim"], atol=1e-6, rtol=0)


@pytest.mark.parametrize("force", ["simt", "tc"])
@pytest.mark.parametrize("name", ["t37_i53_d512", "t64_i64_d64"])
def test_join_reductions_of_the_references_own_matrix(name, force):
    """The thresholded join has no reference definition, but what it reduces does: the dense matrix the reference's
    `F.normalize(T) @ F.normalize(I).T` produced for the fixture (redundancy.py:36-38, written by make_golden.py from
    torch).  Row maximum, argmax, count >= tau and the pair list must be exactly the reductions of THAT matrix, up to
    elements within rounding of the threshold."""
    g = np.load(GOLD / f"redundancy_{name}.npz")
    sim, tau = g["sim"], 0.9
    tol = 2e-6 if force == "simt" else 1e-5
    out = dewi_b200.redundancy_join(g["tfeat"], g["ifeat"], tau=tau, force=force)
    t = sim.shape[0]
    np.testing.assert_allclose(out["max_sim"].cpu().numpy(), sim.max(axis=1), atol=tol)
    am = out["argmax"].cpu().numpy()
    assert np.all(sim[np.arange(t), am] >= sim.max(axis=1) - tol)
    edge = (np.abs(sim - tau) <= tol).sum(axis=1)
    assert np.all(np.abs(out["count"].cpu().numpy() - (sim >= tau).sum(axis=1)) <= edge)
    got = set(zip(out["pairs_i"].cpu().tolist(), out["pairs_j"].cpu().tolist()))
    ref = set(zip(*np.nonzero(sim >= tau)))
    ref = {(int(i), int(j)) for i, j in ref}
    assert ref and all(abs(sim[i, j] - tau) <= tol for i, j in got ^ ref)
    for i, j, sv in zip(out["pairs_i"].cpu().tolist(), out["pairs_j"].cpu().tolist(), out["pairs_sim"].cpu().tolist()):
        assert abs(sv - sim[i, j]) <= tol


def planted(n, d, seed, frac=0.02):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    src = rng.choice(n, int(n * frac), replace=False)
    dst = rng.choice(n, int(n * frac), replace=False)
    x[dst] = x[src] * 1.3 + 0.05 * rng.standard_normal((len(src), d)).astype(np.float32)
    return x


@pytest.mark.parametrize("force", ["simt", "tc"])
@pytest.mark.parametrize("m,n,d", [(1000, 777, 512), (130, 4100, 64)])
def test_cross_join_vs_oracle(m, n, d, force):
    """Both kernels: the fp32 CUDA-core tiles and the tcgen05 CTA-pair sweep with the join epilogue
    (hi+lo bf16 planes, three MMAs)."""
    a, b = planted(m, d, 1), planted(n, d, 2)
    b[: m // 10] = a[: m // 10] * 0.7
    tau = 0.9
    tol = 2e-6 if force == "simt" else 1e-5  # hi/lo planes drop the lo.lo term: up to ~6e-6 when sim -> 1
    mx, am, cnt, pairs = ored.join_rowstats(a, b, tau)
    out = dewi_b200.redundancy_join(a, b, tau=tau, force=force)
    np.testing.assert_allclose(out["max_sim"].cpu().numpy(), mx, atol=tol)
    sim = ored.cross_modal_similarity(a, b)
    got_am = out["argmax"].cpu().numpy()
    assert np.all(sim[np.arange(m), got_am] >= mx - tol)
    # counts may differ only for similarities within rounding of tau
    edge = (np.abs(sim - tau) <= tol).sum(axis=1)
    assert np.all(np.abs(out["count"].cpu().numpy() - cnt) <= edge)
    got = set(zip(out["pairs_i"].cpu().tolist(), out["pairs_j"].cpu().tolist()))
    ref = {(i, j) for i, j, _ in pairs}
    assert all(abs(sim[i, j] - tau) <= tol for i, j in got ^ ref)
    assert out["n_pairs"] == len(got)


def _check_self_join(out, a, tau, mx, cnt, pairs, tol=1e-5):
    np.testing.assert_allclose(out["max_sim"].cpu().numpy(), mx, atol=tol)
    sim = ored.cross_modal_similarity(a, a)
    np.fill_diagonal(sim, -np.inf)
    am = out["argmax"].cpu().numpy()
    assert np.all((am >= 0) & (am < len(a)) & (am != np.arange(len(a))))
    assert np.all(sim[np.arange(len(a)), am] >= mx - tol)  # an index that attains the maximum
    np.testing.assert_array_equal(out["count"].cpu().numpy(), cnt)
    got = list(zip(out["pairs_i"].cpu().tolist(), out["pairs_j"].cpu().tolist()))
    assert len(got) == len(set(got)) == out["n_pairs"]
    assert set(got) == {(i, j) for i, j, _ in pairs} and all(i < j for i, j in got)


@pytest.mark.parametrize("force,symmetric", [("simt", True), ("tc", True), ("tc", False)])
@pytest.mark.parametrize("n", [3000, 2700, 300, 200])  # 12 / 11 / 2 / 1 tiles of 256 rows: even, odd, tiny grids
def test_self_join_vs_oracle(force, symmetric, n):
    """The tensor-core self-join multiplies each unordered pair of 256-row blocks once (circulant half,
    column-direction statistics); `symmetric=False` is the full product.  Both must equal the oracle."""
    a = planted(n, 128, 5, frac=0.05)
    tau = 0.92
    mx, am, cnt, pairs = ored.join_rowstats(a, a, tau, self_join=True)
    out = dewi_b200.redundancy_join(a, tau=tau, force=force, symmetric=symmetric)
    _check_self_join(out, a, tau, mx, cnt, pairs)
    assert out["n_pairs"] == len(pairs) > 0


@pytest.mark.parametrize("n,cuts", [(2500, (0, 768, 1792, 2500)), (5000, (0, 1280, 2560, 3840, 5000)), (700, (0, 256, 700))])
def test_range_joins_add_up_to_the_self_join(n, cuts):
    """The sharded symmetric join on one device: each 'rank' evaluates its row range against the circulant
    half of the block grid; statistics combine with max / sum, pair lists unite without duplicates."""
    a = planted(n, 128, 17, frac=0.05)
    tau = 0.92
    mx, am, cnt, pairs = ored.join_rowstats(a, a, tau, self_join=True)
    parts = [dewi_b200.self_join_range(a, lo, hi, tau=tau) for lo, hi in zip(cuts[:-1], cuts[1:])]
    cmx, carg, ccnt = dewi_b200.combine_range_stats(parts)
    import torch
    merged = {"max_sim": cmx, "argmax": carg, "count": ccnt,
              "pairs_i": torch.cat([p["pairs_i"] for p in parts]), "pairs_j": torch.cat([p["pairs_j"] for p in parts]),
              "n_pairs": sum(p["n_pairs"] for p in parts)}
    _check_self_join(merged, a, tau, mx, cnt, pairs)


def test_symmetric_join_duplicate_heavy_rows():
    """Many exact duplicates: every row's best match is 1.0 with ties everywhere, counts are large and the
    column direction fires on every tile."""
    rng = np.random.RandomState(3)
    base = rng.standard_normal((40, 64)).astype(np.float32)
    a = base[rng.randint(0, 40, size=1500)]
    tau = 0.99
    mx, am, cnt, pairs = ored.join_rowstats(a, a, tau, self_join=True)
    out = dewi_b200.redundancy_join(a, tau=tau, force="tc", pair_cap=1 << 17)
    _check_self_join(out, a, tau, mx, cnt, pairs, tol=2e-5)


def test_range_join_rejects_unaligned_ranges():
    a = planted(1000, 64, 1)
    with pytest.raises(ValueError):
        dewi_b200.self_join_range(a, 100, 612)
    with pytest.raises(ValueError):
        dewi_b200.self_join_range(a, 0, 500)


def test_pair_cap_overflow_is_reported():
    a = np.ones((64, 16), np.float32)
    out = dewi_b200.redundancy_join(a, tau=0.5, pair_cap=10)
    assert out["n_pairs"] == 64 * 63 // 2 and len(out["pairs_i"]) == 10


def test_bf16_join_finds_the_planted_duplicates():
    """One-plane mode (the 10M-row configuration): similarities carry bf16 rounding, so only pairs well
    clear of the threshold are compared."""
    a = planted(20_000, 512, 9, frac=0.01)
    tau = 0.9
    mx, am, cnt, pairs = ored.join_rowstats(a, a, tau, self_join=True)
    out = dewi_b200.redundancy_join(a, tau=tau, precision="bf16")
    np.testing.assert_allclose(out["max_sim"].cpu().numpy(), mx, atol=4e-3)
    got = set(zip(out["pairs_i"].cpu().tolist(), out["pairs_j"].cpu().tolist()))
    clear = {(i, j) for i, j, s in pairs if s >= tau + 0.01}
    assert clear and clear <= got
    # ... and nothing is reported that the exact product does not place within the bf16 tolerance of the
    # threshold: got must be a subset of {pairs with sim >= tau - 4e-3}
    _, _, _, loose = ored.join_rowstats(a, a, tau - 4e-3, self_join=True)
    assert got <= {(i, j) for i, j, _ in loose}
    sims = {(i, j): s for i, j, s in loose}
    for i, j, s in zip(out["pairs_i"].cpu().tolist(), out["pairs_j"].cpu().tolist(), out["pairs_sim"].cpu().tolist()):
        assert abs(s - sims[(i, j)]) <= 4e-3


@pytest.mark.parametrize("force", ["simt", "tc"])
def test_slice_joins_add_up_to_the_self_join(force):
    """Row-sharded self-join on one device: each 'rank' joins its slice of rows against all rows
    (`a_offset`); statistics concatenate and the pair lists unite to the single-call self-join."""
    a = planted(2500, 128, 15, frac=0.05)
    tau = 0.92
    whole = dewi_b200.redundancy_join(a, tau=tau, force=force)
    pairs, mx, cnt = set(), [], []
    for lo, hi in ((0, 700), (700, 1800), (1800, 2500)):
        part = dewi_b200.redundancy_join(a[lo:hi], a, tau=tau, force=force, a_offset=lo)
        got = list(zip(part["pairs_i"].cpu().tolist(), part["pairs_j"].cpu().tolist()))
        assert all(lo <= i < hi and j > i for i, j in got)
        pairs |= set(got)
        mx.append(part["max_sim"].cpu().numpy())
        cnt.append(part["count"].cpu().numpy())
    assert pairs == set(zip(whole["pairs_i"].cpu().tolist(), whole["pairs_j"].cpu().tolist()))
    np.testing.assert_allclose(np.concatenate(mx), whole["max_sim"].cpu().numpy(), atol=1e-6)
    np.testing.assert_array_equal(np.concatenate(cnt), whole["count"].cpu().numpy())
