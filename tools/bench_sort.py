#!/usr/bin/env python
"""This repo's radix sort (csrc/ood_sort.cu through dml_ood_sort) on the key counts of tools/cub_sort_yardstick.cu --
50 x 921 600 keys as one segment / as 50 segments, and 1.3824 G keys -- so that the two can be read side by side.
Prints one JSON line."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import dml_b200
    from dml_b200._lib import check, lib, ptr, stream_ptr
    dev = torch.device("cuda", 0)
    seg = 921600
    out = {}
    for name, n_seg, seg_len in (("one_segment_46M", 1, 50 * seg), ("50_segments_x_921600", 50, seg), ("one_segment_1382M", 1, 1500 * seg)):
        n = n_seg * seg_len
        g = torch.Generator(device=dev).manual_seed(1)
        conf = (torch.rand(n, generator=g, device=dev) * 0.75 + 0.25)
        keys0 = ((conf.view(torch.int32) - 0x3e800000 + 0x3e800000) << 1).contiguous()
        del conf
        ws = torch.empty(lib().dml_ood_workspace_bytes(n_seg, seg_len), dtype=torch.uint8, device=dev)
        keys = torch.empty_like(keys0)
        best = 1e9
        for _ in range(4):
            keys.copy_(keys0)
            res = C.c_void_p()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib().dml_ood_sort(ptr(keys), n_seg, seg_len, 0, 32, ptr(ws), ws.numel(), C.byref(res), stream_ptr(dev)), "dml_ood_sort")
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name] = {"keys": n, "ms": best, "Gkeys_per_s": n / best * 1e-6}
        del keys, keys0, ws
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
