"""Fixed cost vs per-tile cost of the tcgen05 GEMM kernels: time act(x W^T + b) at growing M with the launches
captured in a CUDA graph (no host gaps) and fit  t = a + b * tiles_per_SM.  One JSON line per shape."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import genrec_b200.functional as Fn


def graph_time(fn, reps=20, iters=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)   # us per call


def main():
    dev = torch.device("cuda:0")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    gen = torch.Generator().manual_seed(0)
    for name, N, K in (("bias_act_silu N=512 K=128", 512, 128), ("bias_residual N=128 K=512", 128, 512), ("dact N=512 K=128", 512, 128)):
        pts = []
        for t in (1, 2, 4, 8, 16):
            M = sms * 128 * t // (N // 128)
            x = torch.randn(M, K, generator=gen).to(dev).bfloat16()
            w = (torch.randn(N, K, generator=gen) * 0.1).to(dev).bfloat16()
            b = torch.randn(N, generator=gen).to(dev)
            if name.startswith("bias_act"):
                z = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
                fn = lambda: Fn.linear_fwd(x, w, b, 1, p=0.2, seed=1, site=1)
            elif name.startswith("bias_res"):
                res = torch.randn(M, N, generator=gen).to(dev)
                fn = lambda: Fn.linear_residual_fwd(x, w, b, res, p=0.2, seed=1, site=1)
            else:
                dy = torch.randn(M, K, generator=gen).to(dev).bfloat16()       # [M, 128]
                wt = (torch.randn(K, N, generator=gen) * 0.1).to(dev).bfloat16()   # [128, 512]
                zz = torch.randn(M, N, generator=gen).to(dev).bfloat16()
                fn = lambda: Fn.linear_dact_bwd(dy, wt, zz, 1, p=0.2, seed=1, site=1)
            us = graph_time(fn)
            pts.append((t, us))
        (t0, u0), (t1, u1) = pts[0], pts[-1]
        slope = (u1 - u0) / (t1 - t0)
        print(json.dumps(dict(kernel=name, us_by_tiles_per_sm={str(t): round(u, 2) for t, u in pts}, us_per_tile=round(slope, 3),
                              fixed_us=round(u0 - slope * t0, 2))))


if __name__ == "__main__":
    main()
