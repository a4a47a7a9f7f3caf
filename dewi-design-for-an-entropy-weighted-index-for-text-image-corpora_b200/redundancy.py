"""Redundancy similarity pass on a B200.

`cross_modal_similarity(tfeat, ifeat)` is the arithmetic of
`RedundancyEstimator.compute_cross_modal_similarity` after the CLIP forward
(src/dewi/signals/redundancy.py:36-38): L2-normalise both feature matrices, `T @ I.T`, float32.
`redundancy_join` is the thresholded form that never materialises `[M, N]` (this repository's
definition -- the reference has none).
"""

from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _native


def _torch():
    import torch

    return torch


def _as_cuda(x, dev):
    torch = _torch()
    return torch.as_tensor(x, dtype=torch.float32).to(torch.device("cuda", dev)).contiguous()


def cross_modal_similarity(tfeat, ifeat, device: Optional[int] = None):
    """`normalize(tfeat) @ normalize(ifeat).T` -> numpy `[T, I]` float32 (redundancy.py:36-38)."""
    torch = _torch()
    dev = int(device) if device is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    _native.require_device(dev)
    a, b = _as_cuda(tfeat, dev), _as_cuda(ifeat, dev)
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[1]:
        raise ValueError("feature matrices must be [T, D] and [I, D]")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    lib = _native.load_library()
    with torch.cuda.device(dev):
        rc = lib.dewi_similarity_dense(ctypes.c_void_p(a.data_ptr()), a.shape[0], ctypes.c_void_p(b.data_ptr()),
                                       b.shape[0], a.shape[1], ctypes.c_void_p(out.data_ptr()), dev, _native.stream_ptr())
    _native.check(rc)
    return out.cpu().numpy()  # the reference returns `.cpu().numpy()` (redundancy.py:38)


def redundancy_join(a, b=None, tau: float = 0.9, pair_cap: int = 1 << 20, device: Optional[int] = None,
                    precision: str = "fp32", force: Optional[str] = None, a_offset: int = -1, symmetric: bool = True):
    """Thresholded similarity join.  `b=None` is the self-join (diagonal excluded, pairs j > i).

    symmetric=True (default): a tensor-core self-join multiplies each unordered pair of 256-row blocks once
    (off-diagonal tiles update the statistics of their rows AND columns); False evaluates the full product.

    precision "fp32": similarities good to ~1e-5 (bf16 hi+lo planes on the tensor cores, or the fp32
    CUDA-core kernel for small inputs / d % 64 != 0); "bf16": one bf16 plane, ~1e-3, 3x fewer MMAs.
    force: None | "simt" | "tc" pins the kernel (tests).
    a_offset >= 0: `a` is rows `[a_offset, a_offset + len(a))` of `b` (one shard of a row-sharded
    self-join): each row's own column is excluded, pairs are kept for j > global i and carry global i.

    Returns dict(max_sim [M] f32, argmax [M] i64, count [M] i32, pairs_i, pairs_j, pairs_sim, n_pairs)
    of CUDA tensors; `n_pairs` may exceed `pair_cap`, in which case only `pair_cap` pairs were kept."""
    torch = _torch()
    dev = int(device) if device is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    _native.require_device(dev)
    self_join = b is None
    ta = _as_cuda(a, dev)
    tb = ta if self_join else _as_cuda(b, dev)
    if ta.ndim != 2 or tb.ndim != 2 or ta.shape[1] != tb.shape[1]:
        raise ValueError("feature matrices must be [M, D] and [N, D]")
    m = ta.shape[0]
    tdev = ta.device
    row_max = torch.empty(m, dtype=torch.float32, device=tdev)
    row_arg = torch.empty(m, dtype=torch.int64, device=tdev)
    row_cnt = torch.empty(m, dtype=torch.int32, device=tdev)
    pi = torch.empty(pair_cap, dtype=torch.int64, device=tdev)
    pj = torch.empty(pair_cap, dtype=torch.int64, device=tdev)
    ps = torch.empty(pair_cap, dtype=torch.float32, device=tdev)
    cnt = ctypes.c_int64(0)
    flags = (_native.JOIN_BF16 if precision in ("bf16", "bfloat16") else 0) | \
        {None: 0, "simt": _native.JOIN_FORCE_SIMT, "tc": _native.JOIN_FORCE_TC}[force] | \
        (0 if symmetric else _native.JOIN_NO_SYMMETRY)
    lib = _native.load_library()
    with torch.cuda.device(dev):
        # scratch (operand planes, packed statistics) comes from torch's caching allocator: no cudaMalloc / cudaFree per call
        ws_bytes = int(lib.dewi_join_workspace_bytes(m, tb.shape[0], ta.shape[1], int(self_join), flags))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=tdev)
        rc = lib.dewi_join(ctypes.c_void_p(ta.data_ptr()), m, ctypes.c_void_p(tb.data_ptr()), tb.shape[0], ta.shape[1],
                           float(tau), int(self_join), int(a_offset), flags, ctypes.c_void_p(row_max.data_ptr()),
                           ctypes.c_void_p(row_arg.data_ptr()), ctypes.c_void_p(row_cnt.data_ptr()),
                           ctypes.c_void_p(pi.data_ptr()), ctypes.c_void_p(pj.data_ptr()), ctypes.c_void_p(ps.data_ptr()),
                           int(pair_cap), ctypes.byref(cnt), ctypes.c_void_p(ws.data_ptr()), ws_bytes, dev, _native.stream_ptr())
        del ws
    _native.check(rc)
    kept = min(int(cnt.value), pair_cap)
    return {"max_sim": row_max, "argmax": row_arg, "count": row_cnt, "pairs_i": pi[:kept], "pairs_j": pj[:kept],
            "pairs_sim": ps[:kept], "n_pairs": int(cnt.value)}


def self_join_range(x, row_lo: int, row_hi: int, tau: float = 0.9, pair_cap: int = 1 << 20, device: Optional[int] = None,
                    precision: str = "fp32"):
    """One shard's share of the symmetric self-join of `x` (all rows present): row blocks
    `[row_lo, row_hi)` (multiples of 256, or ending at the last row) against their circulant half of the
    block grid (`dewi_self_join_range`).  The statistics cover ALL rows of `x` but hold this range's
    contribution only; combine shards with max / sum.  Pairs are emitted once each, as (min, max)."""
    torch = _torch()
    dev = int(device) if device is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    _native.require_device(dev)
    tx = _as_cuda(x, dev)
    if tx.ndim != 2:
        raise ValueError("feature matrix must be [N, D]")
    n = tx.shape[0]
    if not (0 <= row_lo <= row_hi <= n) or row_lo % 256 or (row_hi % 256 and row_hi != n):
        raise ValueError("row range must lie inside the matrix and start / end on multiples of 256 (or at the last row)")
    if tx.shape[1] % 64:
        raise ValueError("the symmetric range join runs on the tensor cores and needs d % 64 == 0")
    tdev = tx.device
    row_max = torch.empty(n, dtype=torch.float32, device=tdev)
    row_arg = torch.empty(n, dtype=torch.int64, device=tdev)
    row_cnt = torch.empty(n, dtype=torch.int32, device=tdev)
    pi = torch.empty(pair_cap, dtype=torch.int64, device=tdev)
    pj = torch.empty(pair_cap, dtype=torch.int64, device=tdev)
    ps = torch.empty(pair_cap, dtype=torch.float32, device=tdev)
    cnt = ctypes.c_int64(0)
    flags = _native.JOIN_BF16 if precision in ("bf16", "bfloat16") else 0
    lib = _native.load_library()
    with torch.cuda.device(dev):
        ws_bytes = int(lib.dewi_join_workspace_bytes(n, n, tx.shape[1], 1, flags | _native.JOIN_FORCE_TC))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=tdev)
        rc = lib.dewi_self_join_range(ctypes.c_void_p(tx.data_ptr()), n, tx.shape[1], float(tau), int(row_lo), int(row_hi),
                                      flags, ctypes.c_void_p(row_max.data_ptr()), ctypes.c_void_p(row_arg.data_ptr()),
                                      ctypes.c_void_p(row_cnt.data_ptr()), ctypes.c_void_p(pi.data_ptr()),
                                      ctypes.c_void_p(pj.data_ptr()), ctypes.c_void_p(ps.data_ptr()), int(pair_cap),
                                      ctypes.byref(cnt), ctypes.c_void_p(ws.data_ptr()), ws_bytes, dev, _native.stream_ptr())
        del ws
    _native.check(rc)
    kept = min(int(cnt.value), pair_cap)
    return {"max_sim": row_max, "argmax": row_arg, "count": row_cnt, "pairs_i": pi[:kept], "pairs_j": pj[:kept],
            "pairs_sim": ps[:kept], "n_pairs": int(cnt.value)}


def combine_range_stats(parts):
    """Merge the row statistics of several `self_join_range` results over the same matrix (what the
    sharded join does with all-reduces): max of the best similarities, an argmax that attains it, sum of
    the counts."""
    torch = _torch()
    mx = torch.stack([p["max_sim"] for p in parts]).max(dim=0).values
    big = torch.iinfo(torch.int64).max
    arg = torch.stack([torch.where((p["max_sim"] == mx) & (p["argmax"] >= 0), p["argmax"], big) for p in parts]).min(dim=0).values
    arg = torch.where(arg == big, torch.full_like(arg, -1), arg)
    cnt = torch.stack([p["count"] for p in parts]).sum(dim=0).to(torch.int32)
    return mx, arg, cnt


def sharded_self_join(local_rows, tau: float = 0.9, pair_cap: int = 1 << 20, group=None, precision: str = "bf16"):
    """Row-sharded near-duplicate self-join over the ranks of `group` (one process per GPU, SURVEY.md
    section 8e): every rank passes its contiguous block of rows (rank order = row order); the blocks are
    all-gathered once over NCCL (10M x 512 fp32 = 20 GB, fits every B200).

    When every shard boundary is a multiple of 256 rows and the rows run on the tensor cores
    (d % 64 == 0) the join is SYMMETRIC: rank g multiplies its row blocks against the circulant half of
    the block grid (`self_join_range`), so each unordered pair of blocks is evaluated once across the
    box; the per-row statistics (all N rows, partial per rank) are then combined with three all-reduces
    (max of best similarity, min of the argmax candidates attaining it, sum of counts) and every rank
    returns the statistics of ITS rows.  Otherwise each rank joins its rows against all rows
    (`a_offset` mode, twice the multiplies).  Pairs carry global indices with i < j and the union of the
    ranks' pair lists is the full pair set without duplicates."""
    torch = _torch()
    import torch.distributed as dist

    rows = local_rows.contiguous()
    if not rows.is_cuda:
        raise ValueError("local_rows must live on this rank's GPU")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return redundancy_join(rows, tau=tau, pair_cap=pair_cap, device=rows.device.index, precision=precision)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = torch.zeros(world, dtype=torch.int64, device=rows.device)
    dist.all_gather_into_tensor(counts, torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device), group=group)
    counts = counts.cpu().tolist()
    if len(set(counts)) == 1:
        everything = torch.empty((world * counts[0], rows.shape[1]), dtype=rows.dtype, device=rows.device)
        dist.all_gather_into_tensor(everything, rows, group=group)
    else:
        parts = [torch.empty((c, rows.shape[1]), dtype=rows.dtype, device=rows.device) for c in counts]
        dist.all_gather(parts, rows, group=group)
        everything = torch.cat(parts)
    lo = int(sum(counts[:rank]))
    hi = lo + counts[rank]
    bounds = [int(sum(counts[:r])) for r in range(world)]
    if all(b % 256 == 0 for b in bounds) and rows.shape[1] % 64 == 0:
        out = self_join_range(everything, lo, hi, tau=tau, pair_cap=pair_cap, device=rows.device.index, precision=precision)
        mx = out["max_sim"].clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        big = torch.iinfo(torch.int64).max
        arg = torch.where((out["max_sim"] == mx) & (out["argmax"] >= 0), out["argmax"], big)
        dist.all_reduce(arg, op=dist.ReduceOp.MIN, group=group)
        cnt = out["count"]
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
        out["max_sim"] = mx[lo:hi]
        out["argmax"] = torch.where(arg == big, torch.full_like(arg, -1), arg)[lo:hi]
        out["count"] = cnt[lo:hi]
    else:
        out = redundancy_join(rows, everything, tau=tau, pair_cap=pair_cap, device=rows.device.index, precision=precision,
                              a_offset=lo)
    out["row_offset"] = lo
    return out
