"""A/B of the two B <= 64 sweeps on a bf16 corpus of N random 768-d rows (argv[1]): rows-on-M (search_tcr.cu) against
queries-on-M (search_tc.cu, DEWI_FLAG_NO_ROWS_ON_M), for each batch size in argv[2] (comma list).  Prints step / sweep
time, the fraction of the measured copy peak and the median SM clock sampled while the loop runs."""
import json
import statistics
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
import dewi_b200  # noqa: E402
from dewi_b200 import _native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
batches = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,8,16,32,64").split(",")]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
peak = json.load(open("/root/repo/MEASURED_PEAKS.json")).get("hbm_gbs", 6500.6)
try:
    import pynvml
    pynvml.nvmlInit()
    nv = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:  # noqa: BLE001
    nv = None
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
ix.set_profiling(True)
for b in batches:
    q = torch.randn((b, 768), generator=g, device=dev)
    res = {}
    for name, flags in (("rows-on-M", 0), ("queries-on-M", _native.FLAG_NO_ROWS_ON_M), ("rows-on-M", 0), ("queries-on-M", _native.FLAG_NO_ROWS_ON_M)):
        for _ in range(3):
            ix.search_batch(q, k=10, eta=0.3, entropy_pref=0.5, flags=flags)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ids, sc = ix.search_batch(q, k=10, eta=0.3, entropy_pref=0.5, flags=flags)
        e1.record()
        clocks = []
        while not e1.query():
            if nv is not None:
                clocks.append(pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM))
            time.sleep(0.01)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        kms = float(np.mean([ix.sweep_ms(i)[0] for i in range(min(reps, 32))]))
        kind = ix.sweep_ms(0)[1]
        frac = n * 768 * 2 / (kms * 1e-3) / 1e9 / peak
        mhz = statistics.median(clocks) if clocks else 0
        print(f"n={n} B={b:3d} {name:13s} kind={kind:13s} step_ms={ms:.3f} sweep_ms={kms:.3f} frac={frac:.3f} sm_mhz={mhz:.0f} "
              f"checksum={int(ids.sum())}", flush=True)
