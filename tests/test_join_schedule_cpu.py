"""Host-checkable invariants of the symmetric self-join (no GPU): the circulant block schedule that
`search_tc2.cu: item_span` walks, restated here, covers every unordered pair of 256-row blocks exactly once
for any tile count and any split into row ranges / chunks / rotations; and the workspace size reported by the
C ABI covers the operand planes."""

import ctypes
import itertools

import pytest


def block_row_length(ig: int, T: int) -> int:
    """Tiles block row `ig` meets: itself, the next (T-1)//2 blocks (mod T) and, for even T, the block at
    distance T/2 when ig is in the first half (search_tc2.cu, item_span)."""
    return 1 + (T - 1) // 2 + (1 if (T % 2 == 0 and ig < T // 2) else 0)


def walk(ig: int, T: int, n_chunks: int, chunk: int, rot_seed: int):
    """(tile, is_diagonal) sequence of one work item, in the rotated order the kernel uses."""
    L = block_row_length(ig, T)
    o0, o1 = chunk * L // n_chunks, (chunk + 1) * L // n_chunks
    ln = o1 - o0
    rot = rot_seed % ln if ln > 0 else 0
    for s in range(ln):
        o = o0 + s + rot
        if o >= o0 + ln:
            o -= ln
        t = ig + o
        yield (t - T if t >= T else t), o == 0


@pytest.mark.parametrize("T", list(range(1, 34)) + [97, 128])
def test_every_block_pair_is_multiplied_exactly_once(T):
    for n_chunks, cuts in itertools.product((1, 3, 8), (1, 2, 5)):
        seen = {}
        bounds = [T * i // cuts for i in range(cuts + 1)]            # row ranges of `cuts` ranks
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            for ig in range(lo, hi):
                for chunk in range(n_chunks):
                    for t, diag in walk(ig, T, n_chunks, chunk, rot_seed=73 - (ig - lo) % 74):
                        assert diag == (t == ig)
                        key = (min(ig, t), max(ig, t))
                        seen[key] = seen.get(key, 0) + 1
        want = {(i, j) for i in range(T) for j in range(i, T)}
        assert set(seen) == want, f"T={T}: missing {sorted(want - set(seen))[:5]}"
        assert all(v == 1 for v in seen.values()), f"T={T}: a block pair is multiplied twice"


def test_block_rows_have_equal_length():
    """Contiguous row ranges of equal size carry equal work (that is what the sharded join relies on)."""
    for T in (7, 8, 39063):
        lens = [block_row_length(i, T) for i in range(T)]
        assert max(lens) - min(lens) <= 1
        assert sum(lens) == T * (T + 1) // 2


def test_join_workspace_bytes(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    f = lib.dewi_join_workspace_bytes
    f.restype = ctypes.c_int64
    f.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    BF16, SIMT, TC = 1, 2, 4
    n, d = 1_000_000, 512
    planes = n * d * 2
    self_bf16 = f(n, n, d, 1, BF16)
    self_hilo = f(n, n, d, 1, 0)
    cross_hilo = f(n, n, d, 0, 0)
    assert planes + n * 8 <= self_bf16 <= planes + n * 8 + (1 << 20)          # one plane + packed row statistics
    assert 2 * planes <= self_hilo <= 2 * planes + n * 8 + (1 << 20)          # hi + lo
    assert 4 * planes <= cross_hilo <= 4 * planes + n * 8 + (1 << 20)         # both sides
    assert f(1000, 1000, 48, 1, 0) >= 1000 * 48 * 4                            # d % 64 != 0: fp32 CUDA-core path
    assert f(1000, 1000, 64, 1, SIMT) >= 1000 * 64 * 4 and f(1000, 1000, 64, 1, TC) >= 1024 * 64 * 2 * 2
    assert f(0, 0, 64, 1, 0) == 0
