"""Find which part of the training step invalidates CUDA graph capture."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genrec_b200.hstu import HSTU
from genrec_b200.optim import FlatAdam
from bench import synth_batch, CFG

dev = torch.device("cuda:0")
B, L = int(os.environ.get("DBG_B", "32")), 200
torch.manual_seed(0)
model = HSTU(**CFG).to(dev).train()
opt = FlatAdam(model, lr=1e-3)
ids, ts, tg = (t.to(dev) for t in synth_batch(B, L, CFG["num_items"], 0))

def fwd():
    return model(ids, ts, tg)[1]
def fwd_bwd():
    l = fwd(); l.backward(); return l
def full():
    l = fwd_bwd(); opt.step(); return l
def fwd_nograd():
    with torch.no_grad():
        return model(ids, ts, tg)[1]

for mode in ("global", "thread_local", "relaxed"):
    for name, fn in (("fwd_nograd", fwd_nograd), ("fwd", fwd), ("fwd_bwd", fwd_bwd), ("full", full)):
        try:
            s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2): fn()
            torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode=mode):
                out = fn()
            g.replay(); torch.cuda.synchronize()
            print(f"[{mode}] {name}: OK loss={float(out):.4f}", flush=True)
        except Exception as e:
            print(f"[{mode}] {name}: FAIL {type(e).__name__}: {str(e).splitlines()[0]}", flush=True)
            try: torch.cuda.synchronize()
            except Exception as e2: print("  sync:", e2)
