"""SyncBN kernels (f-4) on one GPU: achieved HBM bandwidth of the statistics / apply / backward passes at a PSPNet
feature-map shape.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from dml_b200.anomaly.lib.nn import SynchronizedBatchNorm2d

B, C, H, W = 8, 512, 90, 160
x = torch.randn(B, C, H, W, device="cuda", requires_grad=True)
gy = torch.randn(B, C, H, W, device="cuda")
m = SynchronizedBatchNorm2d(C, always_sync=True).cuda().train()
ref = torch.nn.BatchNorm2d(C).cuda().train()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fwd(mod):
    with torch.no_grad():
        # training-mode forward without autograd bookkeeping
        return mod(x)


def fwd_bwd(mod):
    x.grad = None
    mod(x).backward(gy)


n = x.numel()
peak = 6552.3
out = {"shape": [B, C, H, W]}
for name, mod in (("ours", m), ("torch_batch_norm", ref)):
    tf = timed(lambda: fwd(mod))
    tb = timed(lambda: fwd_bwd(mod))
    out[name] = {"forward_ms": tf, "forward_GBps": n * 12 / (tf * 1e-3) / 1e9, "forward_frac": n * 12 / (tf * 1e-3) / 1e9 / peak,
                 "forward_backward_ms": tb, "forward_backward_GBps": n * 32 / (tb * 1e-3) / 1e9,
                 "forward_backward_frac": n * 32 / (tb * 1e-3) / 1e9 / peak}
out["bytes_model"] = "forward: read x (stats) + read x + write y = 12 B/elem; + backward: read x, dy (stats) + read x, dy + write dx = 20 B/elem"
print(json.dumps(out))
