"""One launch set of the RQ-VAE residual argmin at N = 2^20 and N = 12,101 (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import genrec_b200.functional as Fn
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
cbs = torch.stack([(torch.rand(256, 32, generator=g) - 0.5) / 2 ** l for l in range(3)]).to(dev)
for N in (1 << 20, 12101):
    x = torch.randn(N, 32, generator=g).to(dev)
    for _ in range(2):
        Fn.rq_residual_argmin(x, cbs, 0.25, want_aux=False)
        Fn.rq_residual_argmin(x, cbs, 0.25, want_aux=True)
torch.cuda.synchronize()
