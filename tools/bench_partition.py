#!/usr/bin/env python
"""Times the local steps of the multi-GPU pooled metric on ONE GPU: range partition (G buckets) vs the radix sort it
replaces, count_positive and the range scan.  CUDA events, 3 warm-ups, keys larger than L2.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from dml_b200.distributed import CudaOps
    n = int(os.environ.get("PART_N", str(400_000_000)))
    dev = torch.device("cuda", 0)
    ops = CudaOps(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    conf = torch.rand(n, generator=g, device=dev)
    conf[::17] = 1.0
    gt = (torch.rand(n, generator=g, device=dev) < 0.01).to(torch.uint8) * 13
    keys, stats = ops.make_keys(conf, gt, (13,), 0x80000000)
    keys = keys.clone()
    del conf, gt

    def timeit(fn, reps=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    out = {"n_keys": n}
    smp = ops.sample_unsorted(keys, 16384).cpu().numpy().view(np.uint32)
    smp = np.sort(smp)
    for G in (2, 4, 8):
        inner = (smp[[(r * smp.size) // G for r in range(1, G)]] & ~np.uint32(1)).view(np.int32).copy()
        b = torch.from_numpy(inner)
        t = timeit(lambda: ops.partition(keys, b))
        _, cnt = ops.partition(keys, b)
        out[f"partition_G{G}_ms"] = t
        out[f"partition_G{G}_GBps"] = n * 12 / t / 1e6
        out[f"partition_G{G}_balance"] = float(cnt.max().item() * G / n)
    work = keys.clone()

    def do_sort():
        work.copy_(keys)
        ops.sort(work, "b")
    t_copy = timeit(lambda: work.copy_(keys))
    out["sort_ms"] = timeit(do_sort) - t_copy
    srt = ops.sort(work, "b")
    out["count_positive_ms"] = timeit(lambda: ops.count_positive(srt))
    info = torch.tensor([0, 0, int(stats[0].item()), n], dtype=torch.int64, device=dev)
    out["scan_range_ms"] = timeit(lambda: ops.scan_range(srt, info, 0.95))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
