"""BASELINE configs[2] geometry (seq_len 2048, d=256, h=8) through the same kernels: one layer, forward + backward, against the
CPU oracle (B=1: the oracle materialises 8 x 2048 x 2048 fp32 score tensors)."""
import pytest
import torch

from tests.util import frob_relerr, relerr

pytestmark = pytest.mark.gpu


def test_hstu_layer_L2048_d256_h8_vs_oracle():
    from genrec_b200.hstu import HSTULayer
    from oracle import hstu as oh
    dev = torch.device("cuda:0")
    torch.manual_seed(2048)
    B, L, D, H = 1, 2048, 256, 8
    layer = HSTULayer(D, H, 0.0, 32, 64, 128, True)
    with torch.no_grad():
        for n, p in layer.named_parameters():
            if "attention_bias" in n:
                p.normal_(0, 0.3)
            elif n.endswith("bias"):
                p.normal_(0, 0.05)
            elif "norm" not in n:
                p.normal_(0, 0.03)
    g = torch.Generator().manual_seed(1)
    ts = 1_300_000_000 + torch.cumsum(torch.randint(1, 3 * 86400, (B, L), generator=g), 1)
    pad = torch.zeros(B, L, dtype=torch.bool); pad[0, :100] = True; ts[pad] = 0
    x = torch.randn(B, L, D, generator=g)
    dy = torch.randn(B, L, D, generator=g) / L ** 0.5
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in layer.state_dict().items()}
    xo = x.clone().requires_grad_(True)
    yo = oh.hstu_layer_forward(xo, pad, ts, sd, "", H)
    yo.backward(dy)
    layer = layer.to(dev).train()
    xg = x.to(dev).requires_grad_(True)
    yg = layer(xg, None, pad.to(dev), ts.to(dev))
    yg.backward(dy.to(dev))
    assert frob_relerr(yg, yo) < 1e-2 and relerr(yg, yo) < 5e-2, (frob_relerr(yg, yo), relerr(yg, yo))
    assert frob_relerr(xg.grad, xo.grad) < 2e-2, frob_relerr(xg.grad, xo.grad)
    for n, p in layer.named_parameters():
        assert frob_relerr(p.grad, sd[n].grad) < 3e-2, (n, frob_relerr(p.grad, sd[n].grad))
