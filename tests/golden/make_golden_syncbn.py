#!/usr/bin/env python
"""Golden fixture for synchronised batch norm, produced by RUNNING THE UNMODIFIED REFERENCE:
anomaly/lib/nn/modules/batchnorm.py's SynchronizedBatchNorm2d on its parallel-training branch (forward lines 64-88,
_compute_mean_std lines 121-139), with the master pipe short-circuited to the module's own _compute_mean_std so that one
"device" holds the whole batch -- the global-batch semantics every multi-rank run must reproduce.  Two training steps
(the moving averages evolve), autograd gradients of sum(y * g).  Run in the build container only:

    python tests/golden/make_golden_syncbn.py        ->  tests/golden/syncbn.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_loader import reference  # noqa: E402


def main():
    torch.manual_seed(20261017)
    B, C, H, W = 4, 6, 5, 7
    # (|mean| / std kept moderate: the reference forms the variance as ssum - sum * mean in fp32, which cancels catastrophically
    #  otherwise and would make the fixture a record of its rounding noise)
    scale = torch.tensor([1.0, 0.1, 5.0, 0.05, 2.0, 1.0]).view(1, C, 1, 1)
    shift = torch.tensor([0.0, 0.3, -2.0, 0.2, 3.0, 0.5]).view(1, C, 1, 1)
    xs = [torch.randn(B, C, H, W) * scale + shift for _ in range(2)]
    for x in xs:
        x[:, 5] = 0.5                                   # constant channel: variance 0 -> the clamp at eps is active
    gs = [torch.randn(B, C, H, W) for _ in range(2)]
    weight = torch.rand(C) + 0.5
    bias = torch.randn(C)
    out = {"weight": weight.numpy(), "bias": bias.numpy(), "eps": np.float32(1e-5), "momentum": np.float32(0.001)}
    with reference("anomaly"):
        from lib.nn import SynchronizedBatchNorm2d
        m = SynchronizedBatchNorm2d(C)
        with torch.no_grad():
            m.weight.copy_(weight)
            m.bias.copy_(bias)
        m.train()
        m._is_parallel = True
        m._parallel_id = 0
        m._sync_master.run_master = lambda msg: m._compute_mean_std(msg.sum, msg.ssum, msg.sum_size)
        for step, (x, g) in enumerate(zip(xs, gs)):
            x = x.clone().requires_grad_(True)
            m.zero_grad()
            y = m(x)
            (y * g).sum().backward()
            out.update({f"x{step}": x.detach().numpy(), f"g{step}": g.numpy(), f"y{step}": y.detach().numpy(),
                        f"dx{step}": x.grad.numpy(), f"dw{step}": m.weight.grad.numpy().copy(), f"db{step}": m.bias.grad.numpy().copy(),
                        f"running_mean{step}": m.running_mean.numpy().copy(), f"running_var{step}": m.running_var.numpy().copy(),
                        f"tmp_running_mean{step}": m._tmp_running_mean.numpy().copy(),
                        f"tmp_running_var{step}": m._tmp_running_var.numpy().copy(), f"running_iter{step}": m._running_iter.numpy().copy()})
        m.eval()
        out["y_eval"] = m(xs[0]).detach().numpy()
    path = os.path.join(HERE, "syncbn.npz")
    np.savez_compressed(path, **out)
    print("wrote syncbn.npz", sorted(out), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
