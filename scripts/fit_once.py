"""One fit_stats call sequence on synthetic Signals columns (profiling / stage timing: DEWI_FIT_TIMING=1)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import dewi_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(21)
hi = torch.tensor([10, 15, 5, 8, 1, 1, 0.2], device=dev).view(7, 1)
sig = torch.rand((7, n), generator=g, device=dev) * hi
s = dewi_b200.DewiScorer()
for _ in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.fit_stats_columns(sig)
    print(f"n={n}: fit_stats wall {(time.perf_counter() - t0) * 1e3:.3f} ms", file=sys.stderr)
