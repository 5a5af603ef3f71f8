#!/usr/bin/env python
"""Device timings of kernel (b) -- fused DCE + VL (+ Inter) loss forward / backward -- and kernel (c) -- masked
per-class sums -- on the few-shot shape of BASELINE.json configs[3] (crop 768, K = 17 / D = 16), against their
algorithmic bytes (DESIGN.md sections 3.5 / 3.6) and the measured HBM bandwidth.  Prints one JSON line.

    python tools/bench_loss_protos.py [--batch 20] [--iters 20]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=20)     # 20 x 17 x 768^2 fp32 = 802 MB >> the 126 MB L2
    ap.add_argument("--size", type=int, default=768)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    import dml_b200
    from dml_b200 import prototypes
    dev = torch.device("cuda", 0)
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    g = torch.Generator(device=dev).manual_seed(4)
    B, S = a.batch, a.size
    px = B * S * S
    out = {"peak_GBps": peak, "batch": B, "size": S}
    for D in (17, 16):
        x = torch.randn(B, D, S, S, device=dev, generator=g)
        # spatially coherent labels (64 x 64 constant tiles, like bench.py's synthetic segmentation maps) with 10 % of the
        # pixels ignored; `_iid` = independent labels per pixel, the worst case of the per-warp class grouping in kernel (c)
        tl = (S + 63) // 64
        t64 = torch.randint(0, D, (B, tl, tl), device=dev, generator=g).repeat_interleave(64, 1).repeat_interleave(64, 2)
        t64 = t64[:, :S, :S].contiguous()
        t64[torch.rand(B, S, S, device=dev, generator=g) < 0.1] = 255
        t8 = t64.to(torch.uint8)
        t8_iid = torch.randint(0, D, (B, S, S), device=dev, generator=g).to(torch.uint8)
        xg = x.clone().requires_grad_(True)

        def fwd():
            return dml_b200.dml_loss(xg, t64, alpha=0.01, beta=0.01 / 80, ignore_index=255)

        def fwd_bwd():
            xg.grad = None
            fwd().backward()

        with torch.no_grad():
            ms_f = timed(lambda: dml_b200.dml_loss(x, t64, alpha=0.01, beta=0.01 / 80, ignore_index=255), a.iters)
        ms_fb = timed(fwd_bwd, a.iters)
        bpp_f, bpp_fb = 4 * D + 8, 12 * D + 16            # int64 targets: L_i = 8
        out[f"loss_D{D}"] = {"fwd_ms": ms_f, "fwd_GBps": px * bpp_f / ms_f / 1e6, "fwd_frac": px * bpp_f / ms_f / 1e6 / peak,
                             "fwd_bwd_ms": ms_fb, "fwd_bwd_GBps": px * bpp_fb / ms_fb / 1e6,
                             "fwd_bwd_frac": px * bpp_fb / ms_fb / 1e6 / peak, "bytes_per_pixel": [bpp_f, bpp_fb],
                             "Mpixel_per_s_fwd_bwd": px / ms_fb / 1e3}
        ms_c = timed(lambda: prototypes.class_sums(x, t8, 19), a.iters)
        bpp_c = 4 * D + 1
        out[f"class_sums_D{D}"] = {"ms": ms_c, "GBps": px * bpp_c / ms_c / 1e6, "frac": px * bpp_c / ms_c / 1e6 / peak,
                                   "bytes_per_pixel": bpp_c, "Mpixel_per_s": px / ms_c / 1e3}
        ms_i = timed(lambda: prototypes.class_sums(x, t8_iid, 19), a.iters)
        out[f"class_sums_D{D}_iid_labels"] = {"ms": ms_i, "GBps": px * bpp_c / ms_i / 1e6, "frac": px * bpp_c / ms_i / 1e6 / peak}
        del x, xg, t64, t8, t8_iid
    print(json.dumps(out))


if __name__ == "__main__":
    main()
