"""One forward + one backward of the tcgen05 attention core at a given geometry (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genrec_b200.functional as Fn  # noqa: E402
from genrec_b200.hstu import RelativePositionBias, _thresholds_on  # noqa: E402

B, L, D, H = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (128, 200, 128, 4)))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
zp = (0.7 * torch.randn(B, L, 4 * D, generator=g)).to(torch.bfloat16).to(dev)
P = torch.nn.functional.silu(zp.float()).to(torch.bfloat16)
dO = (torch.randn(B, L, D, generator=g) / L ** 0.5).to(torch.bfloat16).to(dev)
ts = (1_300_000_000 + torch.cumsum(torch.randint(1, 3 * 86400, (B, L), generator=g), 1)).to(dev)
pad = torch.zeros(B, L, dtype=torch.uint8, device=dev)
rpb = RelativePositionBias(32, 128, H)
meta = Fn.SeqMeta(pad, ts, rpb.bucket_of_delta(L, dev), _thresholds_on(dev), 64, 32, rpb.uniform_of(L, dev))
wpos = (0.3 * torch.randn(32, H, generator=g)).to(dev)
wtime = (0.5 * torch.randn(64, H, generator=g)).to(dev)
for _ in range(3):
    O = Fn.hstu_attention_fwd(P, meta, H, wpos, wtime)
    Fn.hstu_attention_bwd(P, zp, dO, meta, H, wpos, wtime)
torch.cuda.synchronize()
