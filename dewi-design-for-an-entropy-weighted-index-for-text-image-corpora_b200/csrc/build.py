#!/usr/bin/env python
"""Build libdewi_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""

from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent
PKG = CSRC.parent
SOURCES = ["api.cu", "prep.cu", "search_tc.cu", "search_tcr.cu", "search_tc2.cu", "search_simt.cu", "select.cu", "scorer.cu", "join.cu", "extras.cu"]
HEADERS = ["internal.h", "ptx.cuh", "sweep_epilogue.cuh", "join_epilogue.cuh", "../../include/dewi_b200.h"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]
LIB = PKG / "libdewi_b200.so"


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)
    hdrs = [CSRC / h for h in HEADERS] + [Path(__file__)]

    def compile_one(src: str) -> Path:
        obj = objdir / (Path(src).stem + ".o")
        if force or _stale(obj, [CSRC / src] + hdrs):
            cmd = [NVCC, *FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
