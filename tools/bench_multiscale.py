#!/usr/bin/env python
"""Times the anomaly sub-project's score path at the StreetHazards shape (5 stride-8 scales -> 720x1280):

  fused   : stride-8 distance head x5 (dml_head_forward) + ONE dml_multiscale_head_forward launch
            (upsample + average + labels + EDS + MSP + min/max + confusion; nothing full-resolution is read)
  unfused : the same kernels fed by torch's CUDA F.interpolate / `/5` / `+` replay of
            anomaly/models/models.py:659-661 + anomaly/eval_ood_traditional.py:192-208 (what the reference
            executes on a GPU), then dml_head_forward(input_is_logits)

Prints one JSON line.  CUDA events on the current stream, warm-up first, inputs rotate through a ring larger
than L2 only for the unfused path's full-resolution tensors (the fused path has no large input)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SH_SCALES = [(38, 67), (47, 84), (57, 100), (66, 117), (71, 125)]


def main():
    import dml_b200
    from dml_b200 import head as H
    B = int(os.environ.get("MS_BATCH", "50"))
    K, size = 13, (720, 1280)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    embs = [torch.randn(B, K, h, w, generator=g, device=dev) * 0.7 + 1.0 for (h, w) in SH_SCALES]
    # spatially coherent ground truth (64x64 blocks), like real label maps and like bench.py's synthetic data
    gt = torch.randint(0, K + 1, (B, (size[0] + 63) // 64, (size[1] + 63) // 64), generator=g, device=dev)
    gt = gt.repeat_interleave(64, 1).repeat_interleave(64, 2)[:, :size[0], :size[1]].contiguous().to(torch.uint8)
    out = H.HeadOutput()
    lows = [H.HeadOutput() for _ in embs]
    conf = torch.zeros(K + 1, K, dtype=torch.int64, device=dev)

    def fused():
        z = [H.dml_head(e, want_logits=True, label_dtype=None, out=o).logits for e, o in zip(embs, lows)]
        H.dml_multiscale_head(z, size, label_dtype=torch.uint8, want_eds=True, eds_clamp=400.0, want_msp=True,
                              want_minmax=True, gt=gt, confusion=conf, out=out)

    out_u = H.HeadOutput()

    def unfused(nb):
        for s in range(0, B, nb):
            scores = torch.zeros(nb, K, *size, device=dev)
            for e, o in zip(embs, lows):
                z = H.dml_head(e[s:s + nb], want_logits=True, label_dtype=None).logits
                scores = scores + F.interpolate(z, size=size, mode="bilinear", align_corners=False) / len(embs)
            H.dml_head(scores, input_is_logits=True, label_dtype=torch.uint8, want_eds=True, eds_clamp=400.0,
                       want_msp=True, want_minmax=True, gt=gt[s:s + nb], confusion=conf)

    def timeit(fn, reps):
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

    l0 = dml_b200.load_library().dml_kernel_launches()
    fused()
    launches = dml_b200.load_library().dml_kernel_launches() - l0
    if os.environ.get("MS_ONLY_FUSED"):      # profiling runs (ncu): just the fused path, a few launches
        fused()
        torch.cuda.synchronize()
        return
    t_f = timeit(fused, 20)
    t_u = timeit(lambda: unfused(1), 3)       # the reference's batch size (1 image per forward)
    t_u10 = timeit(lambda: unfused(10), 3)
    px = B * size[0] * size[1]
    print(json.dumps({"workload": f"{B} x 5 scales -> 720x1280, K=13", "fused_ms": t_f, "fused_Mpixel_s": px / t_f / 1e3,
                      "fused_launches": int(launches), "fused_write_GBps": px * 10 / t_f / 1e6,
                      "torch_replay_batch1_ms": t_u, "torch_replay_batch10_ms": t_u10,
                      "speedup_vs_batch1": t_u / t_f, "speedup_vs_batch10": t_u10 / t_f}))


if __name__ == "__main__":
    main()
