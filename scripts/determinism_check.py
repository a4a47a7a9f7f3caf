"""Run one fp32 batch repeatedly: results must be identical, the certificate must hold every time."""
import sys
import time

import torch

sys.path.insert(0, ".")
import dewi_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(42)
ix = dewi_b200.CudaIndex(768, dtype="fp32", device=0)
ix.add_batch(None, torch.randn((n, 768), generator=g, device=dev), normalized=False)
ix.set_payload_columns(torch.rand(n, generator=g, device=dev), torch.rand(n, generator=g, device=dev))
ix.build()
q = torch.randn((b, 768), generator=g, device=dev)
ref = None
for it in range(12):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ids, sc = ix.search_batch(q, k=10, eta=0.3, entropy_pref=0.5)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    if ref is None:
        ref = (ids.clone(), sc.clone())
    same = bool(torch.equal(ids, ref[0]) and torch.equal(sc, ref[1]))
    print(f"it {it}: {dt:.3f} ms  same={same}  cert={ix.cert_stats()}  launches={ix.last_launches()}")
