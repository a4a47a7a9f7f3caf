#!/usr/bin/env python
"""Write tests/golden/cuda_saved_index_{fp32,bf16}: directories saved by `dewi_b200.DewiIndex.save` on a B200,
plus the searches the CUDA backend answered on them.  Run on the GPU box (gpurun), copy the output from
gpurun_out/ into tests/golden/.  tests/test_reference_interop_cpu.py then has the UNMODIFIED reference load
those directories (`dewi.index.DewiIndex.load`, index.py:143-166 -> ExactIndex.load, backends.py:517-556) and
reproduce the stored results -- SURVEY.md section 8f row N1, the direction the CUDA backend cannot test itself."""

import shutil
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import dewi_b200  # noqa: E402
from _util import PAYLOAD_FIELDS, synth_payload_columns  # noqa: E402

out_root = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "golden"
out_root.mkdir(parents=True, exist_ok=True)
for dtype in ("fp32", "bf16"):
    rng = np.random.RandomState(71 if dtype == "fp32" else 72)
    n, d, k = 48, 32, 5
    emb = rng.randn(n, d).astype(np.float32)
    pay = synth_payload_columns(rng, n, "profile")
    ix = dewi_b200.DewiIndex(dim=d, space="cosine", dtype=dtype, rerank_eta=0.3, entropy_pref=0.5)
    for i in range(n):
        ix.add(f"doc_{i:03d}", emb[i], dewi_b200.Payload(**{f: float(pay[i, j]) for j, f in enumerate(PAYLOAD_FIELDS)}),
               meta={"source": f"file_{i}.txt"})
    ix.build()
    dst = out_root / f"cuda_saved_index_{dtype}"
    shutil.rmtree(dst, ignore_errors=True)
    ix.save(dst)
    q = rng.randn(4, d).astype(np.float32)
    res = [ix.search(q[i], k=k) for i in range(4)]
    np.savez_compressed(out_root / f"cuda_saved_index_{dtype}_queries.npz", queries=q,
                        ids=np.array([[r[0] for r in rr] for rr in res]), scores=np.array([[r[1] for r in rr] for rr in res]))
    print("wrote", dst)
