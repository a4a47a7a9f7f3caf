"""Redundancy similarity pass: dense product vs torch's F.normalize + matmul fixtures
(redundancy.py:36-38); thresholded join vs this repository's CPU restatement (parity unpinned)."""

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


def planted(n, d, seed, frac=0.02):
    rng = np.random.RandomState(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    src = rng.choice(n, int(n * frac), replace=False)
    dst = rng.choice(n, int(n * frac), replace=False)
    x[dst] = x[src] * 1.3 + 0.05 * rng.standard_normal((len(src), d)).astype(np.float32)
    return x


@pytest.mark.parametrize("m,n,d", [(1000, 777, 512), (130, 4100, 64)])
def test_cross_join_vs_oracle(m, n, d):
    a, b = planted(m, d, 1), planted(n, d, 2)
    b[: m // 10] = a[: m // 10] * 0.7
    tau = 0.9
    mx, am, cnt, pairs = ored.join_rowstats(a, b, tau)
    out = dewi_b200.redundancy_join(a, b, tau=tau)
    np.testing.assert_allclose(out["max_sim"].cpu().numpy(), mx, atol=2e-6)
    sim = ored.cross_modal_similarity(a, b)
    got_am = out["argmax"].cpu().numpy()
    assert np.all(sim[np.arange(m), got_am] >= mx - 2e-6)
    # counts may differ only for similarities within rounding of tau
    edge = (np.abs(sim - tau) <= 2e-6).sum(axis=1)
    assert np.all(np.abs(out["count"].cpu().numpy() - cnt) <= edge)
    got = set(zip(out["pairs_i"].cpu().tolist(), out["pairs_j"].cpu().tolist()))
    ref = {(i, j) for i, j, _ in pairs}
    assert all(abs(sim[i, j] - tau) <= 2e-6 for i, j in got ^ ref)
    assert out["n_pairs"] == len(got)


def test_self_join_vs_oracle():
    a = planted(3000, 128, 5, frac=0.05)
    tau = 0.92
    mx, am, cnt, pairs = ored.join_rowstats(a, a, tau, self_join=True)
    out = dewi_b200.redundancy_join(a, tau=tau)
    np.testing.assert_allclose(out["max_sim"].cpu().numpy(), mx, atol=2e-6)
    np.testing.assert_array_equal(out["count"].cpu().numpy(), cnt)
    got = set(zip(out["pairs_i"].cpu().tolist(), out["pairs_j"].cpu().tolist()))
    assert got == {(i, j) for i, j, _ in pairs} and all(i < j for i, j in got)
    assert out["n_pairs"] == len(pairs) > 0


def test_pair_cap_overflow_is_reported():
    a = np.ones((64, 16), np.float32)
    out = dewi_b200.redundancy_join(a, tau=0.5, pair_cap=10)
    assert out["n_pairs"] == 64 * 63 // 2 and len(out["pairs_i"]) == 10
