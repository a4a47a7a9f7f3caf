"""The host-side launch planners, walked over the shape space on the CPU (no GPU needed: `dewi_plan_probe` only runs
the planning code of csrc/search_tc.cu, search_tcr.cu and search_tc2.cu).

Every plan a planner hands out must be launchable: its (mode, tile, query rows, residency) combination has to be one
of the kernels the library actually instantiates -- read here from the shared library's own symbol table, an
independent source -- and fit the SM's shared memory, ring depth and grid limits.  Two holes of exactly this kind were
found at the end of round 2 (an M = 64 sweep planned on 128-row tiles, for which no kernel exists; a staged first round
longer than the seed selection's window) and are pinned below."""

import ctypes
import re
import shutil
import subprocess

import numpy as np
import pytest

SMEM_MAX = 227 * 1024
SEED_WINDOW = 2048   # csrc/internal.h: kSeedWindow


@pytest.fixture(scope="module")
def lib(lib_path):
    from dewi_b200 import _native

    return _native.load_library()


@pytest.fixture(scope="module")
def kernels(lib_path):
    """Instantiations of the sweep kernels, from the host stubs in the library's symbol table."""
    if shutil.which("nm") is None:
        pytest.skip("binutils nm is not available")
    out = subprocess.run(["nm", "-C", str(lib_path)], capture_output=True, text=True, check=True).stdout
    found = set(re.findall(r"(search_t\w*_kernel<[^>]*>)", out))
    assert any(k.startswith("search_tc_kernel<") for k in found) and any(k.startswith("search_tcr_kernel<") for k in found)
    return found


def probe(lib, which, mode, dim, n_rows, n_qb, kc, sms, q_rows, opt):
    out = (ctypes.c_int * 10)()
    rc = lib.dewi_plan_probe(which, mode, dim, int(n_rows), n_qb, kc, sms, q_rows, opt, out)
    return rc, list(out)


def check_single(kernels, plan, mode, dim, n_rows, n_qb, sms, what):
    rows_on_m, q_rows, n_tile, stages, q_stages, q_res, chunks, grid, smem, pmode = plan
    assert smem <= SMEM_MAX, what
    assert 1 <= grid <= sms and chunks >= 1, what
    assert chunks <= -(-n_rows // n_tile), what
    if rows_on_m:
        assert f"search_tcr_kernel<{q_rows}>" in kernels, what
        assert stages >= 3 and q_res == 1 and q_stages == dim // 64 and pmode == 0 and n_qb == 1, what
    else:
        assert f"search_tc_kernel<{pmode}, {n_tile}, {q_rows}, {q_res}>" in kernels, f"{what}: no such kernel"
        assert stages >= 2 and pmode == mode, what
        assert grid <= chunks * n_qb, what


def test_single_cta_plans_are_launchable(lib, kernels):
    rng = np.random.RandomState(11)
    dims = [64, 128, 192, 256, 384, 512, 768, 1024, 2048, 4096, 8192]
    rows = [1, 200, 2047, 2048, 2049, 40_000, 1_000_000, 12_500_000, 100_000_000]
    seen = set()
    for trial in range(6000):
        mode = int(rng.randint(3))
        dim, n_rows = int(rng.choice(dims)), int(rng.choice(rows))
        kc = int(rng.randint(1, 430))
        sms = int(rng.choice([148, 132, 74]))
        if rng.rand() < 0.5:    # at most 64 queries: M = 64, optionally the rows-on-M sweep (api.cu: q_rows / swap_b)
            n_qb, q_rows, opt = 1, 64, int(rng.choice([0, 1, 8, 16, 17, 33, 64]))
        else:
            n_qb, q_rows, opt = int(rng.choice([1, 2, 3, 7])), 128, 0
        rc, plan = probe(lib, 0, mode, dim, n_rows, n_qb, kc, sms, q_rows, opt)
        if rc == 0:
            check_single(kernels, plan, mode, dim, n_rows, n_qb, sms,
                         f"mode={mode} dim={dim} n={n_rows} n_qb={n_qb} kc={kc} q_rows={q_rows} opt={opt}: {plan}")
            seen.add((plan[0], plan[9], plan[2], plan[1], plan[5]))
    # the walk reached every family: rows-on-M at each width, M = 64 resident and streamed, M = 128 on both tile sizes
    assert {(1, 0, 256, 16, 1), (1, 0, 256, 32, 1), (1, 0, 256, 64, 1)} <= seen
    assert any(s[0] == 0 and s[3] == 64 and s[4] == 1 for s in seen) and any(s[0] == 0 and s[3] == 64 and s[4] == 0 for s in seen)
    assert any(s[0] == 0 and s[3] == 128 and s[2] == 128 for s in seen) and any(s[0] == 0 and s[3] == 128 and s[2] == 256 for s in seen)


def test_m64_sweep_is_never_planned_on_128_row_tiles(lib):
    """k ~ 105-187 with at most 64 queries: lists of 226-390 entries leave room for 128-row tiles only, and the M = 64
    sweeps of one / two planes exist for 256-row tiles -- the planner must say no (the caller then takes the CUDA-core
    sweep) instead of handing out a plan `tc_launch` refuses."""
    for mode in (0, 1):
        for kc in range(200, 420, 5):
            rc, plan = probe(lib, 0, mode, 768, 1_000_000, 1, kc, 148, 64, 0)
            assert rc != 0 or plan[2] == 256, (mode, kc, plan)
    rc, plan = probe(lib, 0, 0, 128, 5000, 1, 316, 148, 64, 0)     # k = 150: kc = 2k + 16
    assert rc != 0


def test_pair_plans_and_staged_first_rounds(lib, kernels):
    rng = np.random.RandomState(12)
    staged_seen = 0
    for trial in range(6000):
        mode = int(rng.randint(3))
        dim = int(rng.choice([64, 128, 256, 512, 768, 1024, 4096]))
        n_rows = int(rng.choice([300, 2048, 40_000, 700_000, 1_000_000, 4_000_000, 12_500_000, 100_000_000]))
        n_qb = 2 * int(rng.randint(1, 40))
        kc = int(rng.randint(1, 200))
        sms = int(rng.choice([148, 132]))
        staged = int(rng.randint(2))
        rc, plan = probe(lib, 1, mode, dim, n_rows, n_qb, kc, sms, 128, staged)
        if rc != 0:
            continue
        pmode, stages, _, chunks, grid, first, smem, tile_rows, _, _ = plan
        what = f"mode={mode} dim={dim} n={n_rows} n_qb={n_qb} kc={kc} staged={staged}: {plan}"
        assert f"search_tc2_kernel<{pmode}, 0, 0>" in kernels and pmode == mode, what
        assert 2 <= stages and smem <= SMEM_MAX, what
        assert grid % 2 == 0 and 2 <= grid <= 2 * (sms // 2), what
        n_tiles, n_qpairs, clusters = -(-n_rows // tile_rows), n_qb // 2, grid // 2
        assert 1 <= chunks <= n_tiles and clusters <= chunks * n_qpairs, what
        if not staged:
            assert first == 0, what
        if first:
            staged_seen += 1
            a = first // n_qpairs
            assert first % n_qpairs == 0 and 1 <= a < chunks, what                     # whole chunks, and a second launch remains
            assert first <= clusters, what                                              # one round of clusters
            assert a * kc <= SEED_WINDOW, what                                          # what seed_from_partials can rank
            assert ((chunks - a) * n_qpairs) % clusters == 0, what                      # the rest is a whole number of rounds
    assert staged_seen > 50


def test_staged_first_round_respects_the_seed_window(lib):
    """B = 1024 (4 query pairs), k = 60 (kc = 136): one round of 74 clusters would hold 18 chunks x 136 = 2448 list
    entries per query -- more than the 2048 the seed selection ranks; the plan must stay a single launch."""
    rc, plan = probe(lib, 1, 0, 64, 120_000, 8, 136, 148, 128, 1)
    assert rc == 0 and plan[5] == 0, plan
    rc, plan = probe(lib, 1, 0, 64, 700_000, 8, 36, 148, 128, 1)      # k = 10: staged, 18 chunks in the first round
    assert rc == 0 and plan[5] == 18 * 4, plan


def test_plans_of_the_benchmark_configurations(lib):
    """The launch geometry the measured numbers in DESIGN.md section 6 were taken with (148 SMs): a regression here
    changes what `bench.py` times."""
    # C3 100M x 768 bf16, k = 10 (kc = 36): B <= 64 -> rows on M, one CTA per SM, resident queries (12 k-blocks)
    rc, p = probe(lib, 0, 0, 768, 100_000_000, 1, 36, 148, 64, 64)
    assert rc == 0 and p[:8] == [1, 64, 256, 3, 12, 1, 148, 148], p          # QN = 64: three ring stages beside the queries
    rc, p = probe(lib, 0, 0, 768, 100_000_000, 1, 36, 148, 64, 1)
    assert rc == 0 and p[:8] == [1, 16, 256, 4, 12, 1, 148, 148], p          # QN = 16: four
    rc, p = probe(lib, 0, 0, 768, 100_000_000, 1, 36, 148, 128, 0)
    assert rc == 0 and p[:8] == [0, 128, 256, 4, 2, 0, 148, 148], p          # one block of 65..128 queries: queries on M
    # B = 4096 (32 query blocks): CTA pairs, four ring stages, items a multiple of the 74 clusters
    rc, p = probe(lib, 1, 0, 768, 100_000_000, 32, 36, 148, 128, 0)
    assert rc == 0 and p[1] == 4 and p[4] == 148 and (p[3] * 16) % 74 == 0 and p[5] == 0, p
    # C2 1M x 768 fp32, certified single-plane sweep (kc = 48), staged: 4 of 41 chunks first at B = 4096, 18 of 55 at B = 1024
    rc, p = probe(lib, 1, 0, 768, 1_000_000, 32, 48, 148, 128, 1)
    assert rc == 0 and (p[3], p[5]) == (41, 4 * 16), p
    rc, p = probe(lib, 1, 0, 768, 1_000_000, 8, 48, 148, 128, 1)
    assert rc == 0 and (p[3], p[5]) == (55, 18 * 4), p
