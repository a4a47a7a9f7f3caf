"""`DewiScorer` / `RobustStats` with the reference's API (src/dewi/scorer.py:11-89), computed on a B200.

`fit_stats(rows)` / `score(sig)` / `score_conditional(sig)` / `is_fitted()` keep their signatures.
The arithmetic runs in libdewi_b200.so: exact median/MAD by radix selection (K3) and the float64
z-score -> clip -> sigmoid kernel (K4).  Batch forms (`fit_stats_columns`, `score_batch`) are the
extension that makes 10^8 rows practical; the per-row methods call the same kernels with n = 1.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np

from . import _native
from .types import SIGNAL_FIELDS, Weights


def _torch():
    import torch

    return torch


def _device_index(device: Optional[int]) -> int:
    if device is not None:
        return int(device)
    torch = _torch()
    return torch.cuda.current_device() if torch.cuda.is_available() else 0


MAX_FIT_COLUMNS = 32  # columns one dewi_fit_stats call accepts (csrc/scorer.cu: kMaxCols)


def _fit_columns(cols, device: int, zero_mad_as: float = 1e-8):
    """cols: CUDA float32 tensor [F, N] (row c = one signal column).  Returns (med[F], mad[F]) as float64
    arrays.  The library reports a zero MAD as 1e-8 (scorer.py:24); `zero_mad_as` maps it back for callers
    with a different rule (robust.py:8-10 adds 1e-8 instead)."""
    torch = _torch()
    lib = _native.load_library()
    f, n = cols.shape
    med_arr = np.empty(f, dtype=np.float64)
    mad_arr = np.empty(f, dtype=np.float64)
    # (the library selects the device itself; the stream is the current one of THAT device.)  One call fits up to
    # MAX_FIT_COLUMNS columns; the reference fits however many keys a row has (scorer.py:19-25), so wider inputs take
    # several calls -- the seven DEWI signals take one.
    for c0 in range(0, f, MAX_FIT_COLUMNS):
        fc = min(MAX_FIT_COLUMNS, f - c0)
        med = (ctypes.c_double * fc)()
        mad = (ctypes.c_double * fc)()
        rc = lib.dewi_fit_stats(ctypes.c_void_p(cols.data_ptr() + c0 * cols.stride(0) * cols.element_size()), n, fc,
                                cols.stride(0), med, mad, device,
                                ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        _native.check(rc)
        med_arr[c0:c0 + fc] = med[:]
        mad_arr[c0:c0 + fc] = mad[:]
    if zero_mad_as != 1e-8:
        mad_arr[mad_arr == 1e-8] = zero_mad_as  # a float32 MAD is never 1e-8 exactly unless it was zero
    return med_arr, mad_arr


@dataclass
class RobustStats:
    """Median and MAD per signal (scorer.py:11-31)."""

    medians: Dict[str, float]
    mads: Dict[str, float]

    @classmethod
    def fit(cls, rows: Sequence[Mapping[str, float]], device: Optional[int] = None) -> "RobustStats":
        """Keys come from `rows[0]`; every key present is fitted (scorer.py:19-25)."""
        keys = list(rows[0].keys())
        host = np.asarray([[r[k] for r in rows] for k in keys], dtype=np.float32)  # scorer.py:21
        return cls.fit_columns(dict(zip(keys, host)), device)

    @classmethod
    def fit_columns(cls, columns: Mapping[str, "np.ndarray"], device: Optional[int] = None) -> "RobustStats":
        """columns: name -> 1-D array / tensor of equal length (host or device)."""
        torch = _torch()
        dev = _device_index(device)
        _native.require_device(dev)
        keys = list(columns.keys())
        tdev = torch.device("cuda", dev)
        cols = torch.stack([torch.as_tensor(columns[k], dtype=torch.float32).to(tdev) for k in keys]).contiguous()
        return cls.fit_matrix(keys, cols, dev)

    @classmethod
    def fit_matrix(cls, keys: Sequence[str], matrix, device: Optional[int] = None) -> "RobustStats":
        """`matrix`: `[len(keys), N]` array / tensor, row c = the column named keys[c].  A float32 CUDA
        tensor with unit stride along N is used in place (no copy)."""
        torch = _torch()
        dev = _device_index(device)
        _native.require_device(dev)
        cols = torch.as_tensor(matrix)
        if cols.ndim != 2 or cols.shape[0] != len(keys):
            raise ValueError("expected one row per key")
        cols = cols.to(torch.device("cuda", dev), torch.float32)
        if cols.stride(1) != 1:
            cols = cols.contiguous()
        med, mad = _fit_columns(cols, dev)
        return cls(medians={k: float(m) for k, m in zip(keys, med)}, mads={k: float(m) for k, m in zip(keys, mad)})

    def z(self, name: str, val: float) -> float:
        return float((val - self.medians[name]) / (1.4826 * self.mads[name]))  # scorer.py:28-31


class DewiScorer:
    """Robust DEWI scorer, standard and conditional modes (scorer.py:34-89)."""

    def __init__(self, weights: Optional[Weights] = None, delta: float = 3.0, device: Optional[int] = None):
        self.weights = weights or Weights()
        self.weights.delta = delta  # the constructor argument wins, as in the reference (scorer.py:37-39)
        self.stats: Optional[RobustStats] = None
        self.device = device

    # ---- fitting ---------------------------------------------------------------------------------
    def fit_stats(self, rows: List[Mapping[str, float]]) -> None:
        self.stats = RobustStats.fit(rows, self.device)

    def fit_stats_columns(self, columns) -> None:
        """Batch form: a mapping name -> column, or a `[7, N]` array/tensor in `SIGNAL_FIELDS` order."""
        if not isinstance(columns, Mapping):
            self.stats = RobustStats.fit_matrix(SIGNAL_FIELDS, columns, self.device)
        else:
            self.stats = RobustStats.fit_columns(columns, self.device)

    def is_fitted(self) -> bool:
        return self.stats is not None

    # ---- scoring ---------------------------------------------------------------------------------
    def _w6(self):
        w = self.weights
        return (ctypes.c_double * 6)(w.alpha_t, w.alpha_i, w.alpha_m, w.alpha_r, w.alpha_n, w.delta)

    def score_batch(self, columns, conditional: bool = False, out_dtype: str = "float32"):
        """Scores for N rows at once.  `columns`: mapping name -> column, or `[7, N]` in `SIGNAL_FIELDS`
        order (numpy or torch, host or device).  Returns a CUDA tensor `[N]` (float32, or float64)."""
        assert self.stats is not None, "Call fit_stats() before scoring."
        torch = _torch()
        dev = _device_index(self.device)
        _native.require_device(dev)
        tdev = torch.device("cuda", dev)
        if isinstance(columns, Mapping):
            parts = [torch.as_tensor(columns[k]) for k in SIGNAL_FIELDS]
            in_f64 = all(p.dtype == torch.float64 for p in parts)
            cols = torch.stack([p.to(tdev, torch.float64 if in_f64 else torch.float32) for p in parts])
        else:
            cols = torch.as_tensor(columns)
            in_f64 = cols.dtype == torch.float64  # float64 rows are scored un-rounded, as the reference does
            cols = cols.to(tdev, torch.float64 if in_f64 else torch.float32)
        cols = cols.contiguous()
        if cols.ndim != 2 or cols.shape[0] != 7:
            raise ValueError("expected seven signal columns")
        n = cols.shape[1]
        f64 = out_dtype in ("float64", "f64", torch.float64)
        out = torch.empty(n, dtype=torch.float64 if f64 else torch.float32, device=tdev)
        med = (ctypes.c_double * 7)(*[self.stats.medians[k] for k in SIGNAL_FIELDS])
        mad = (ctypes.c_double * 7)(*[self.stats.mads[k] for k in SIGNAL_FIELDS])
        lib = _native.load_library()
        with torch.cuda.device(dev):
            rc = lib.dewi_score(ctypes.c_void_p(cols.data_ptr()), int(in_f64), n, cols.stride(0), med, mad, self._w6(),
                                int(bool(conditional)), ctypes.c_void_p(out.data_ptr()), int(f64), dev,
                                _native.stream_ptr())
        _native.check(rc)
        return out

    def _score_one(self, sig: Mapping[str, float], conditional: bool) -> float:
        assert self.stats is not None, "Call fit_stats() before scoring."
        # the reference scores the caller's un-rounded Python floats (scorer.py:28-31): float64 in
        cols = np.asarray([[sig[k]] for k in SIGNAL_FIELDS], dtype=np.float64)
        return float(self.score_batch(cols, conditional, "float64")[0].item())

    def score(self, sig: Mapping[str, float]) -> float:
        return self._score_one(sig, False)

    def score_conditional(self, sig: Mapping[str, float]) -> float:
        return self._score_one(sig, True)
