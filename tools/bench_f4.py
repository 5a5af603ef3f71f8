"""One pass of the f-4 kernels for an ncu capture: SyncBN forward + backward at a PSPNet feature-map shape, the five
resizes of one 720 x 1280 image."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from dml_b200.anomaly import dataset as D
from dml_b200.anomaly.lib.nn import SynchronizedBatchNorm2d

x = torch.randn(8, 512, 90, 160, device="cuda", requires_grad=True)
m = SynchronizedBatchNorm2d(512, always_sync=True).cuda().train()
m(x).backward(torch.randn_like(x))
img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (720, 1280, 3), dtype=np.uint8)).cuda()
D.val_image_pyramid(img, (300, 375, 450, 525, 600), 1000, 8)
torch.cuda.synchronize()
