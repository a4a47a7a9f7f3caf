"""Profiling helper: one bf16-plane self-join of N random 512-d rows (argv[1], default 400000); prints its
wall time.  Used under ncu / with the DEWI_JOIN_* experiment switches."""
import sys
import time

import torch

sys.path.insert(0, "/root/repo")
import dewi_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator(device="cuda")
g.manual_seed(1)
x = torch.randn((n, 512), generator=g, device="cuda")
for r in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = dewi_b200.redundancy_join(x, tau=0.9, precision="bf16", pair_cap=1 << 20)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    print(f"n={n} rep={r} ms={ms:.1f} alg_tflops={n * (n - 1) * 512 / ms / 1e9:.0f} pairs={out['n_pairs']} "
          f"mean_max={float(out['max_sim'].mean()):.4f}")
