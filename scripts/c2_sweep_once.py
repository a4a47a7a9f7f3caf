"""Kernel time of the fp32 search sweep at a few batch sizes (experiments with DEWI_CERT_* switches)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import dewi_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
batches = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "64,1024,4096").split(",")]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(42)
ix = dewi_b200.CudaIndex(768, dtype="fp32", device=0)
ix.add_batch(None, torch.randn((n, 768), generator=g, device=dev), normalized=False)
ix.set_payload_columns(torch.rand(n, generator=g, device=dev), torch.rand(n, generator=g, device=dev))
ix.build()
for b in batches:
    q = torch.randn((b, 768), generator=g, device=dev)
    ix.set_profiling(True)
    for _ in range(6):
        ix.search_batch(q, k=10, eta=0.3, entropy_pref=0.5)
    torch.cuda.synchronize()
    ms = float(np.mean([ix.sweep_ms(i)[0] for i in range(4)]))
    ix.set_profiling(False)
    print(f"B={b}: sweep {ms:.3f} ms  ({2.0 * b * n * 768 / ms / 1e9:.0f} TFLOP/s algorithmic)  cert (used, failed) = {ix.cert_stats()}")
