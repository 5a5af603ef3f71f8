"""CPU: the resize oracle against the reference's own ValDataset output (golden), against Pillow itself, and the
library's host-side coefficient tables against the oracle's."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import resize_oracle as R

GOLD = os.path.join(os.path.dirname(__file__), "golden", "resize_val.npz")


def test_oracle_reproduces_the_reference_val_dataset():
    g = np.load(GOLD)
    pyr = R.val_image_pyramid(g["img"], tuple(int(v) for v in g["img_sizes"]), int(g["img_max_size"]), int(g["padding_constant"]))
    assert len(pyr) == 5
    for i, t in enumerate(pyr):
        ref = g[f"img_data_{i}"]
        assert t.shape == ref.shape and t.dtype == np.float32
        np.testing.assert_array_equal(t.view(np.uint32), ref.view(np.uint32))          # bit for bit
    np.testing.assert_array_equal(g["seg_label"][0], g["segm"].astype(np.int64) - 1)
    np.testing.assert_array_equal(g["img_ori"], g["img"])


@pytest.mark.parametrize("shape", [(72, 128, 30, 53), (72, 128, 72, 128), (50, 70, 80, 100), (37, 41, 8, 8), (16, 16, 1, 1),
                                   (20, 30, 57, 29), (3, 5, 9, 2), (180, 320, 76, 134)])
def test_oracle_equals_pillow(shape):
    Image = pytest.importorskip("PIL.Image")
    H, W, oh, ow = shape
    rng = np.random.default_rng(H * 1000 + ow)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    img[: H // 3] = 255                                               # saturated band: the clip at 255
    ref = np.array(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
    np.testing.assert_array_equal(R.resize_bilinear_u8(img, oh, ow), ref)


def test_target_sizes_streethazards():
    # anomaly/config/*.yaml: imgSizes (300, 375, 450, 525, 600), imgMaxSize 1000, padding_constant 8 on 720 x 1280
    assert R.val_target_sizes(720, 1280, (300, 375, 450, 525, 600), 1000, 8) == [(304, 536), (376, 672), (456, 800), (528, 936), (568, 1000)]


@pytest.mark.parametrize("sizes", [(1280, 536), (720, 304), (90, 90), (64, 200), (1000, 7), (5, 1)])
def test_library_coefficient_tables_equal_the_oracle(sizes):
    """dml_resize_coeffs is host code: callable without a GPU"""
    from dml_b200 import load_library
    lib = load_library()
    n_in, n_out = sizes
    b_ref, k_ref = R.coeffs(n_in, n_out)
    ks = lib.dml_resize_ksize(n_in, n_out)
    assert ks == k_ref.shape[1]
    b = np.empty((n_out, 2), np.int32)
    k = np.empty((n_out, ks), np.int32)
    assert lib.dml_resize_coeffs(n_in, n_out, b.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), ks) == 0
    np.testing.assert_array_equal(b, b_ref)
    np.testing.assert_array_equal(k, k_ref)
    assert lib.dml_resize_coeffs(n_in, n_out, b.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), ks + 1) != 0


def test_product_target_sizes_equal_the_oracle():
    from dml_b200.anomaly import dataset as D
    for (h, w) in ((720, 1280), (1024, 2048), (90, 160), (375, 500), (500, 375)):
        for sizes, mx, pad in (((300, 375, 450, 525, 600), 1000, 8), ((38, 47, 56, 66, 75), 125, 8), ((512,), 2048, 32)):
            assert D.val_target_sizes(h, w, sizes, mx, pad) == R.val_target_sizes(h, w, sizes, mx, pad)
    assert D.round2nearest_multiple(533, 8) == 536 and D.round2nearest_multiple(536, 8) == 536
