#!/usr/bin/env python
"""Golden fixture for the validation input pipeline, produced by RUNNING THE UNMODIFIED
REFERENCE: anomaly/dataset.py's ValDataset.__getitem__ on a seeded synthetic image + label map written to a temp
directory as PNG files (StreetHazards' 720 x 1280 and the reference's imgSizes / imgMaxSize / padding_constant, all
divided by 8).  Run in the build container only:

    python tests/golden/make_golden_resize.py        ->  tests/golden/resize_val.npz
"""
import os
import sys
import tempfile
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _ref_loader import reference  # noqa: E402


def main():
    from PIL import Image
    import PIL
    rng = np.random.default_rng(20261017)
    h, w = 90, 160
    # smooth blobs + noise + a few saturated patches: exercises rounding, clipping and the box filter of the down-scales
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(127 + 120 * np.sin(xx / (5.0 + c) + yy / (9.0 - c))) for c in range(3)], -1)
    img = np.clip(img + rng.normal(0, 25, img.shape), 0, 255).astype(np.uint8)
    img[10:20, 30:60] = 255
    img[50:70, 100:130] = 0
    segm = rng.integers(0, 14, (h, w)).astype(np.uint8)
    opt = SimpleNamespace(imgSizes=(38, 47, 56, 66, 75), imgMaxSize=125, padding_constant=8)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "images/test/t0"))
        os.makedirs(os.path.join(tmp, "annotations/test/t0"))
        Image.fromarray(img).save(os.path.join(tmp, "images/test/t0/7.png"))
        Image.fromarray(segm, mode="L").save(os.path.join(tmp, "annotations/test/t0/7.png"))
        rec = [{"fpath_img": "images/test/t0/7.png", "fpath_segm": "annotations/test/t0/7.png", "width": w, "height": h}]
        with reference("anomaly"):
            import dataset as ref_dataset
            ds = ref_dataset.ValDataset(tmp, rec, opt)
            out = ds[0]
    arrays = {"img": img, "segm": segm, "seg_label": out["seg_label"].numpy(), "img_ori": out["img_ori"],
              "img_sizes": np.asarray(opt.imgSizes), "img_max_size": np.asarray(opt.imgMaxSize),
              "padding_constant": np.asarray(opt.padding_constant), "pillow": np.asarray(PIL.__version__)}
    for i, t in enumerate(out["img_data"]):
        arrays[f"img_data_{i}"] = t.numpy()
    path = os.path.join(HERE, "resize_val.npz")
    np.savez_compressed(path, **arrays)
    print("wrote resize_val.npz", {k: getattr(v, "shape", None) for k, v in arrays.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
