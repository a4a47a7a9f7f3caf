"""Profiling helper: bf16 index of N random 768-d rows (argv[1]), B queries (argv[2]); prints the mean
step and sweep time of `reps` searches.  Used for A/B runs with the DEWI_TC2_* experiment switches."""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
import dewi_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(7)
ix = dewi_b200.CudaIndex(768, dtype="bf16", device=0)
ix.reserve(n)
done = 0
while done < n:
    m = min(1_000_000, n - done)
    ix.add_batch(None, torch.randn((m, 768), generator=g, device=dev), normalized=False)
    done += m
ix.set_payload_columns(torch.rand(n, generator=g, device=dev), torch.rand(n, generator=g, device=dev))
ix.build()
q = torch.randn((b, 768), generator=g, device=dev)
ix.set_profiling(True)
for _ in range(2):
    ix.search_batch(q, k=10, eta=0.3, entropy_pref=0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ids, sc = ix.search_batch(q, k=10, eta=0.3, entropy_pref=0.5)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
kms = float(np.mean([ix.sweep_ms(i)[0] for i in range(reps)]))
print(f"n={n} B={b} step_ms={ms:.3f} sweep_ms={kms:.3f} tflops={2.0 * b * n * 768 / kms / 1e9:.0f} checksum={int(ids.sum())}")
