"""Repeat the deferred-vs-inline comparison of tests/test_cfg2_parity_gpu.py and report WHERE the flat gradient differs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_batch
from genrec_b200.hstu import HSTU
from genrec_b200.optim import FlatAdam
import genrec_b200.functional as Fn

dev = torch.device("cuda:0")
ids, ts, tg = make_batch(6, 70, 300, seed=1, pad=True, device=dev)
bad = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    res = []
    for defer in (False, True):
        torch.manual_seed(0)
        m = HSTU(300, 70, 64, 2, 2, dropout=0.0).to(dev).train()
        opt = FlatAdam(m, lr=1e-3, unit_loss_grad=True, defer_weight_grads=defer)
        Fn.set_defer_weight_grads(defer)
        gs = []
        for _ in range(2):
            _, loss = m(ids, ts, tg)
            loss.backward()
            opt.sync_grads()
            gs.append(opt.grad.clone())
            opt.step()
        torch.cuda.synchronize()
        res.append((gs, opt.flat.clone(), opt))
    Fn.set_defer_weight_grads(False)
    Fn.join_deferred(dev)
    for it in range(2):
        a, b = res[0][0][it], res[1][0][it]
        d = (a - b).abs()
        tol = 1e-5 * a.abs().max().item() + 1e-3 * a.abs()
        if (d > tol).any():
            bad += 1
            idx = (d > tol).nonzero().flatten()
            opt = res[0][2]
            names = [n for n, q in m.named_parameters() if q.requires_grad]
            offs = opt.buffers.offsets
            hit = {}
            for i in idx.tolist():
                k = max(j for j, o in enumerate(offs) if o <= i)
                hit[names[k]] = hit.get(names[k], 0) + 1
            print(f"rep {rep} iter {it}: {idx.numel()} elements differ, max diff {d.max().item():.3e} (max |g| {a.abs().max().item():.3e}) in {hit}")
            break
print("bad runs:", bad)
