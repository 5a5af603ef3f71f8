"""GPU: dml_resize_bilinear_normalize / anomaly.dataset against the reference's ValDataset output (golden), the oracle
and Pillow -- bit for bit (anomaly/dataset.py:11-21,65-70,281-297)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import resize_oracle as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "resize_val.npz")


def _bits(t):
    return t.detach().cpu().numpy().view(np.uint32)


def test_pyramid_equals_the_reference_val_dataset():
    from dml_b200.anomaly import dataset as D
    g = np.load(GOLD)
    img = torch.from_numpy(g["img"]).cuda()
    pyr = D.val_image_pyramid(img, tuple(int(v) for v in g["img_sizes"]), int(g["img_max_size"]), int(g["padding_constant"]))
    assert len(pyr) == 5
    for i, t in enumerate(pyr):
        ref = g[f"img_data_{i}"]
        assert tuple(t.shape) == ref.shape and t.dtype == torch.float32 and t.is_cuda
        np.testing.assert_array_equal(_bits(t), ref.view(np.uint32))


@pytest.mark.parametrize("shape", [(72, 128, 30, 53), (72, 128, 72, 128), (50, 70, 80, 100), (37, 41, 8, 8), (16, 16, 1, 1),
                                   (20, 30, 57, 29), (3, 5, 9, 2), (180, 320, 76, 134), (33, 1000, 5, 65)])
def test_resize_equals_oracle(shape):
    from dml_b200.anomaly import dataset as D
    H, W, oh, ow = shape
    rng = np.random.default_rng(H * 1000 + ow)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    img[: H // 3] = 255
    got = D.imresize_normalize(torch.from_numpy(img).cuda(), (ow, oh))
    ref = R.img_transform(R.resize_bilinear_u8(img, oh, ow))[None]
    np.testing.assert_array_equal(_bits(got), ref.view(np.uint32))


def test_full_shape_pyramid_equals_pillow_and_torchvision():
    """720 x 1280 at the reference's five scales: PIL resize + the reference's img_transform arithmetic, live"""
    Image = pytest.importorskip("PIL.Image")
    from dml_b200.anomaly import dataset as D
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:720, 0:1280]
    img = np.stack([127 + 120 * np.sin(xx / (37.0 + 5 * c) + yy / (61.0 - 7 * c)) for c in range(3)], -1)
    img = np.clip(img + rng.normal(0, 20, img.shape), 0, 255).astype(np.uint8)
    pyr = D.val_image_pyramid(torch.from_numpy(img).cuda(), (300, 375, 450, 525, 600), 1000, 8)
    sizes = [(304, 536), (376, 672), (456, 800), (528, 936), (568, 1000)]
    mean = torch.tensor(R.MEAN).view(3, 1, 1)
    std = torch.tensor(R.STD).view(3, 1, 1)
    for t, (th, tw) in zip(pyr, sizes):
        assert tuple(t.shape) == (1, 3, th, tw)
        pil = np.array(Image.fromarray(img).resize((tw, th), Image.BILINEAR))
        x = torch.from_numpy((np.float32(pil) / 255.).transpose((2, 0, 1)).copy())
        ref = x.sub_(mean).div_(std)                                  # torchvision.transforms.Normalize
        np.testing.assert_array_equal(_bits(t[0]), ref.numpy().view(np.uint32))


def test_batched_and_out_argument():
    from dml_b200.anomaly import dataset as D
    rng = np.random.default_rng(9)
    imgs = rng.integers(0, 256, (3, 40, 56, 3), dtype=np.uint8)
    out = torch.empty(3, 3, 24, 32, device="cuda")
    got = D.imresize_normalize(torch.from_numpy(imgs).cuda(), (32, 24), out=out)
    assert got.data_ptr() == out.data_ptr()
    for b in range(3):
        ref = R.img_transform(R.resize_bilinear_u8(imgs[b], 24, 32))
        np.testing.assert_array_equal(_bits(got[b]), ref.view(np.uint32))


def test_val_dataset_on_files(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    from dml_b200.anomaly import dataset as D
    g = np.load(GOLD)
    os.makedirs(tmp_path / "images/test/t0")
    os.makedirs(tmp_path / "annotations/test/t0")
    Image.fromarray(g["img"]).save(tmp_path / "images/test/t0/7.png")
    Image.fromarray(g["segm"], mode="L").save(tmp_path / "annotations/test/t0/7.png")
    rec = [{"fpath_img": "images/test/t0/7.png", "fpath_segm": "annotations/test/t0/7.png", "width": 160, "height": 90}]
    opt = SimpleNamespace(imgSizes=tuple(int(v) for v in g["img_sizes"]), imgMaxSize=int(g["img_max_size"]),
                          padding_constant=int(g["padding_constant"]))
    ds = D.ValDataset(str(tmp_path), rec, opt)
    assert len(ds) == 1
    out = ds[0]
    assert set(out) == {"img_ori", "img_data", "seg_label", "info", "name"}
    np.testing.assert_array_equal(out["img_ori"], g["img_ori"])
    np.testing.assert_array_equal(out["seg_label"].numpy(), g["seg_label"])
    assert out["info"] == "images/test/t0/7.png" and out["name"] == os.path.join("t0", "7.png")
    for i, t in enumerate(out["img_data"]):
        np.testing.assert_array_equal(_bits(t), g[f"img_data_{i}"].view(np.uint32))


def test_error_behaviour():
    from dml_b200 import DmlError
    from dml_b200.anomaly import dataset as D
    with pytest.raises(DmlError):
        D.imresize_normalize(torch.zeros(4, 4, 3, dtype=torch.uint8), (2, 2))          # CPU tensor: no fallback
    with pytest.raises(ValueError):
        D.imresize_normalize(torch.zeros(4, 4, 3, device="cuda"), (2, 2))               # float image
    with pytest.raises(ValueError):
        D.imresize_normalize(torch.zeros(4, 4, 3, dtype=torch.uint8, device="cuda"), (0, 2))
