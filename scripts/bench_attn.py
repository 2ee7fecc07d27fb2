"""Micro-benchmark of the attention kernels (CUDA events, warm, L2-sized working sets): the tcgen05 attention core alone, and one
whole HSTU block forward + backward with the tcgen05 vs the mma.sync attention kernels.  Prints JSON lines."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import genrec_b200.functional as Fn  # noqa: E402
from genrec_b200.hstu import HSTULayer, RelativePositionBias, _thresholds_on  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    dev = torch.device("cuda:0")
    for name, B, L, D, H in (("cfg2", 128, 200, 128, 4), ("cfg3", 16, 2048, 256, 8), ("ref-default", 128, 50, 64, 2)):
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
        t_f = timeit(lambda: Fn.hstu_attention_fwd(P, meta, H, wpos, wtime))
        t_b = timeit(lambda: Fn.hstu_attention_bwd(P, zp, dO, meta, H, wpos, wtime))
        flops_f = 2 * D * L * (L + 1) * B
        print(json.dumps(dict(shape=name, B=B, L=L, D=D, H=H, kernel="attn_tc", fwd_us=round(t_f, 1), bwd_us=round(t_b, 1),
                              fwd_tflops=round(flops_f / t_f / 1e6, 1), bwd_tflops=round(2 * flops_f / t_b / 1e6, 1))), flush=True)
        layer = HSTULayer(D, H, 0.0, 32, 64, 128, True).to(dev).train()
        x = torch.randn(B, L, D, device=dev)
        dy = torch.randn(B, L, D, device=dev)
        for mode in ("tc", "mma"):
            os.environ["GRB_ATTN"] = mode

            def step():
                xi = x.clone().requires_grad_(True)
                y = layer(xi, None, pad.bool(), ts)
                y.backward(dy)

            t = timeit(step, iters=10, warm=3)
            F = (72 * L * D * D + 6 * D * L * (L + 1)) * B
            print(json.dumps(dict(shape=name, kernel=f"layer_fwd_bwd[{mode}]", us=round(t, 1), alg_tflops=round(F / t / 1e6, 1))), flush=True)
        os.environ.pop("GRB_ATTN", None)


if __name__ == "__main__":
    main()
