"""Validation input pipeline (f-4): five resizes + normalisations of a 720 x 1280 image, GPU (CUDA events) vs the
reference's CPU route (PIL + NumPy + torch) on one host core.  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from PIL import Image
from dml_b200.anomaly import dataset as D

rng = np.random.default_rng(0)
B = 16
imgs = rng.integers(0, 256, (B, 720, 1280, 3), dtype=np.uint8)
sizes = D.val_target_sizes(720, 1280, (300, 375, 450, 525, 600), 1000, 8)
dev = torch.from_numpy(imgs).cuda()


def gpu_step():
    return [D.imresize_normalize(dev, (tw, th)) for th, tw in sizes]


for _ in range(3):
    gpu_step()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
ev0.record()
for _ in range(reps):
    gpu_step()
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / reps
out_bytes = sum(3 * th * tw * 4 for th, tw in sizes) * B
in_bytes = imgs.nbytes
mean = torch.tensor(D.MEAN).view(3, 1, 1)
std = torch.tensor(D.STD).view(3, 1, 1)
t0 = time.perf_counter()
n_cpu = 4
for b in range(n_cpu):
    im = Image.fromarray(imgs[b])
    for th, tw in sizes:
        a = np.float32(np.array(im.resize((tw, th), Image.BILINEAR))) / 255.
        torch.from_numpy(a.transpose((2, 0, 1)).copy()).sub_(mean).div_(std)
cpu_ms = (time.perf_counter() - t0) / n_cpu * 1e3
print(json.dumps({"workload": f"{B} x 720x1280 RGB -> 5 scales {sizes}", "gpu_ms_per_image": ms / B,
                  "gpu_images_per_s": B / (ms * 1e-3), "algorithmic_GBps": (out_bytes + 5 * in_bytes) / (ms * 1e-3) / 1e9,
                  "bytes_per_image": (out_bytes + 5 * in_bytes) // B, "cpu_ms_per_image_1core": cpu_ms,
                  "speedup_vs_1core": cpu_ms / (ms / B)}))
