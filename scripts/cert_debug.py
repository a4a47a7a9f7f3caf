import sys
import torch
n, b = 1_000_000, 1024
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(42)
e = torch.randn((n, 768), generator=g, device=dev)
_ = torch.rand(n, generator=g, device=dev); _ = torch.rand(n, generator=g, device=dev)
q = torch.randn((b, 768), generator=g, device=dev)
e = e / e.norm(dim=1, keepdim=True); q = q / q.norm(dim=1, keepdim=True)
eh = e.half().float(); qh = q.half().float()
row_err = (e - eh).norm(dim=1); row_hi = eh.norm(dim=1)
D, E = row_err.max().item(), row_hi.max().item()
qe = (q - qh).norm(dim=1)
gamma = 768 * 2.0 ** -22
eps = 1.0 * D + qe * E + gamma * (1 + qe) * E
S = qh @ eh.T                                    # [b, n] swept scores (fp32 accumulate)
top = S.topk(128, dim=1).values
bar = top[:, 19] - 2 * eps
cnt = (S >= bar[:, None]).sum(dim=1)
print("D", D, "E", E, "eps mean", eps.mean().item(), "count in margin: mean", cnt.float().mean().item(), "max", cnt.max().item(),
      "queries >= 80:", int((cnt >= 80).sum()), ">= 96:", int((cnt >= 96).sum()), ">=64:", int((cnt >= 64).sum()))
exact = q @ e.T
print("max |S - exact|", (S - exact).abs().max().item(), " vs eps", eps.min().item())
