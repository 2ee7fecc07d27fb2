"""GEMM building blocks through the C ABI (tcgen05/TMA path by default; GRB_GEMM=mma selects the mma.sync path) vs torch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_linear(x, w, b):
    return x.float() @ w.float().T + b


@pytest.mark.parametrize("T,N,K", [(128, 128, 64), (210, 512, 128), (25600, 512, 128), (1000, 128, 512), (77, 64, 64), (300, 256, 1024)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_linear_forward(T, N, K, act):
    import genrec_b200.functional as Fn
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(T + N + K)
    x = torch.randn(T, K, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.1).to(dev).bfloat16()
    b = torch.randn(N, generator=g).to(dev)
    z, a = Fn.linear_fwd(x, w, b, act)
    ref = _ref_linear(x, w, b)
    torch.testing.assert_close(z.float(), ref, rtol=1e-2, atol=2e-2)
    if act == 1:
        torch.testing.assert_close(a.float(), torch.nn.functional.silu(z.float()), rtol=1e-2, atol=1e-2)
    if act == 2:
        torch.testing.assert_close(a.float(), torch.relu(z.float()), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("T,N,K", [(128, 128, 128), (210, 512, 128), (25600, 512, 128), (25600, 128, 512), (77, 64, 64), (1000, 12102, 128)])
def test_linear_backward(T, N, K):
    import genrec_b200.functional as Fn
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(T + N + K + 1)
    dy = (torch.randn(T, N, generator=g) * 0.1).to(dev).bfloat16()
    if N % 8:
        pytest.skip("N must be a multiple of 8 for a contiguous bf16 operand")
    x = torch.randn(T, K, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.1).to(dev).bfloat16()
    res = torch.randn(T, K, generator=g).to(dev)
    dx, dw, db = Fn.linear_bwd(dy, w, x, dx_residual=res)
    torch.testing.assert_close(dx, res + dy.float() @ w.float(), rtol=1e-2, atol=1e-2 * (N ** 0.5) * 0.1)
    ref_dw = dy.float().T @ x.float()
    assert ((dw - ref_dw).abs().max() / ref_dw.abs().max()).item() < 2e-3
    ref_db = dy.float().sum(0)
    assert ((db - ref_db).abs().max() / ref_db.abs().max()).item() < 2e-3


def test_linear_residual_and_rowscale():
    import genrec_b200.functional as Fn
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    T, N, K = 333, 128, 512
    x = torch.randn(T, K, generator=g).to(dev).bfloat16()
    w = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    b = torch.randn(N, generator=g).to(dev)
    res = torch.randn(T, N, generator=g).to(dev)
    rs = (torch.rand(T, generator=g) > 0.3).float().to(dev)
    y = Fn.linear_residual_fwd(x, w, b, res, rs)
    torch.testing.assert_close(y, (res + _ref_linear(x, w, b)) * rs[:, None], rtol=1e-2, atol=2e-2)
