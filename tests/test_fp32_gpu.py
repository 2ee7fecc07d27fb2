"""fp32-exact forward path (north_star: "1e-5 (fp32)"): HSTU blocks and the tied head against the golden fixtures of the UNMODIFIED
reference run in fp32, and against the fp32 oracle (in fp64 where that removes the oracle's own rounding) on tile-boundary shapes.
Tolerance, written here: max-norm relative error <= 1e-5."""
import pytest
import torch

from tests.util import make_batch, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _randomise(layer):
    with torch.no_grad():
        for n, p in layer.named_parameters():
            if "attention_bias" in n:
                p.normal_(0, 0.5)
            elif n.endswith("bias"):
                p.normal_(0, 0.1)
            elif "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
            else:
                p.normal_(0, 0.08)


@pytest.mark.parametrize("B,L,D,H", [(2, 7, 64, 2), (3, 50, 128, 4), (2, 200, 128, 4), (2, 257, 128, 2), (2, 130, 256, 8)])
def test_layer_fp32_vs_oracle(B, L, D, H):
    from genrec_b200.hstu import HSTULayer
    from oracle import hstu as oh
    dev = torch.device("cuda:0")
    torch.manual_seed(L * 131 + D)
    layer = HSTULayer(D, H, 0.0, 32, 64, 128, True)
    _randomise(layer)
    ids, ts, _ = make_batch(max(B, 3), L, 50, seed=L)
    ids, ts = ids[:B], ts[:B]
    x = torch.randn(B, L, D)
    sd64 = {k: v.detach().double() for k, v in layer.state_dict().items()}
    with torch.no_grad():
        yo = oh.hstu_layer_forward(x.double(), ids == 0, ts, sd64, "", H).float()
    layer = layer.to(dev).eval()
    layer.precision = "fp32"
    with torch.no_grad():
        yg = layer(x.to(dev), None, (ids == 0).to(dev), ts.to(dev))
    assert yg.dtype == torch.float32
    assert relerr(yg, yo) < TOL, relerr(yg, yo)
    # and the bf16 path on the same inputs sits where bf16 sits - the two modes are really different code
    layer.precision = "bf16"
    with torch.no_grad():
        yb = layer(x.to(dev), None, (ids == 0).to(dev), ts.to(dev))
    assert 1e-4 < relerr(yb, yo) < 3e-2


@pytest.mark.parametrize("name", ["hstu_model_d64h2.pt", "hstu_model_d128h4_nots.pt", "hstu_model_notime.pt"])
def test_model_fp32_vs_reference_golden(golden, name):
    """Logits of the unmodified reference (fp32, CPU) reproduced to 1e-5, top-10 identical."""
    from genrec_b200.hstu import HSTU
    g = golden(name)
    cfg = g["cfg"]
    dev = torch.device("cuda:0")
    m = HSTU(cfg["num_items"], 64, cfg["embed_dim"], cfg["num_heads"], cfg["num_blocks"], dropout=0.0,
             use_temporal_bias=cfg["use_temporal_bias"])
    m.load_state_dict(g["state_dict"])
    m = m.to(dev).eval().set_precision("fp32")
    ids = g["input_ids"].to(dev)
    ts = g["timestamps"].to(dev) if cfg["pass_ts"] else None
    with torch.no_grad():
        logits, loss = m(ids, ts)
    assert loss is None and logits.shape == g["logits"].shape
    assert relerr(logits, g["logits"]) < TOL, relerr(logits, g["logits"])
    with torch.no_grad():
        last = m.last_logits(ids, ts)
    assert relerr(last, g["logits"][:, -1]) < TOL
    top = m.predict(ids, ts, top_k=10)
    assert torch.equal(top.cpu(), g["top10"])


def test_fp32_path_is_forward_only():
    from genrec_b200.hstu import HSTU
    dev = torch.device("cuda:0")
    m = HSTU(50, 16, 64, 2, 1, dropout=0.0).to(dev).eval().set_precision("fp32")
    ids, ts, tg = make_batch(2, 16, 50, seed=1)
    with pytest.raises(RuntimeError):
        m(ids.to(dev), ts.to(dev))                       # grad enabled, parameters require grad
    with torch.no_grad(), pytest.raises(RuntimeError):
        m(ids.to(dev), ts.to(dev), tg.to(dev))           # no loss in this mode
    m.set_precision("bf16")
    logits, _ = m(ids.to(dev), ts.to(dev))                # back on the default path
    assert logits.shape == (2, 16, 51)


def test_linear_f32x3_bias_residual_ragged_n():
    """The generalised split-bf16 GEMM: bias, residual, SiLU, and an N that is no multiple of 4 (the tied head has N = V + 1)."""
    from genrec_b200 import functional as Fn
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(300, 128, generator=g); w = 0.1 * torch.randn(1001, 128, generator=g); b = torch.randn(1001, generator=g)
    y = Fn.linear_f32x3_bias(Fn.split3(x.to(dev), 0), Fn.split3(w.to(dev), 1), b.to(dev), None, 1).cpu()
    ref = torch.nn.functional.silu(x.double() @ w.double().t() + b.double()).float()
    assert y.shape == ref.shape and relerr(y, ref) < 2e-6
    w2 = 0.1 * torch.randn(128, 512, generator=g); h = torch.randn(300, 512, generator=g); r = torch.randn(300, 128, generator=g)
    y2 = Fn.linear_f32x3_bias(Fn.split3(h.to(dev), 0), Fn.split3(w2.to(dev), 1), b[:128].to(dev), r.to(dev), 0).cpu()
    ref2 = (h.double() @ w2.double().t() + b[:128].double() + r.double()).float()
    assert relerr(y2, ref2) < 2e-6
