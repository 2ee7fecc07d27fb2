"""One HSTU layer forward + backward at cfg-3 geometry (B=16, L=2048, D=256, H=8) for an ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from genrec_b200.hstu import HSTULayer

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, L, D, H = 16, 2048, 256, 8
layer = HSTULayer(D, H, 0.0, 32, 64, 128, True).to(dev).train()
ts = (1_300_000_000 + torch.cumsum(torch.randint(1, 3 * 86400, (B, L), generator=g), 1)).to(dev)
pad = torch.zeros(B, L, dtype=torch.bool, device=dev)
x = torch.randn(B, L, D, generator=g).to(dev).requires_grad_(True)
dy = torch.randn(B, L, D, generator=g).to(dev)
for _ in range(3):
    y = layer(x, None, pad, ts)
    y.backward(dy)
torch.cuda.synchronize()
print("done")
