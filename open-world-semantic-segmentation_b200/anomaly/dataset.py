"""GPU form of the reference's validation input pipeline:

  anomaly/dataset.py:11-21      imresize(im, size, 'bilinear')       -> PIL.Image.resize(size, Image.BILINEAR)
  anomaly/dataset.py:65-70      BaseDataset.img_transform            -> float32 / 255, HWC -> CHW, Normalize(mean, std)
  anomaly/dataset.py:242-322    ValDataset.__getitem__               -> one resized + normalised tensor per entry of imgSizes

``ValDataset`` keeps the reference's constructor and the keys of the dict it returns; decoding the files stays on the CPU
(PIL, untouched), the five resizes + normalisations of an image run as five launches of ``dml_resize_bilinear_normalize`` on
the decoded uint8 image and are bit-identical to what PIL + NumPy + torchvision produce (Pillow's 8-bit two-pass
resampling with integer coefficients; tests/test_gpu_resize.py).  No CPU fallback: the tensors must live on a CUDA device."""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .._lib import check, lib, ptr, require_cuda, stream_ptr

MEAN = (0.485, 0.456, 0.406)          # anomaly/dataset.py:35-37
STD = (0.229, 0.224, 0.225)
_TILE_ROWS = 8                         # output rows per CTA tile of the kernel (csrc/dml_resize.cu RS_TY)


class _AxisPlan:
    """tap windows + integer weights of one axis (host: dml_resize_coeffs; device copies cached per device)"""

    def __init__(self, in_size: int, out_size: int):
        self.in_size, self.out_size = int(in_size), int(out_size)
        self.ksize = int(lib().dml_resize_ksize(self.in_size, self.out_size))
        if self.ksize <= 0:
            raise ValueError(f"resize: invalid sizes {in_size} -> {out_size}")
        self.bounds = np.empty((self.out_size, 2), np.int32)
        self.coeffs = np.empty((self.out_size, self.ksize), np.int32)
        check(lib().dml_resize_coeffs(self.in_size, self.out_size, self.bounds.ctypes.data_as(C.c_void_p),
                                      self.coeffs.ctypes.data_as(C.c_void_p), self.ksize), "dml_resize_coeffs")
        lo, cnt = self.bounds[:, 0].astype(np.int64), self.bounds[:, 1].astype(np.int64)
        last = np.minimum(np.arange(self.out_size) + _TILE_ROWS - 1, self.out_size - 1)
        self.max_rows_per_tile = int((lo[last] + cnt[last] - lo)[::_TILE_ROWS].max())
        self._dev: Dict[torch.device, Tuple[torch.Tensor, torch.Tensor]] = {}

    def on(self, device) -> Tuple[torch.Tensor, torch.Tensor]:
        device = torch.device(device)
        if device not in self._dev:
            self._dev[device] = (torch.from_numpy(self.bounds).to(device), torch.from_numpy(self.coeffs).to(device))
        return self._dev[device]


_PLANS: Dict[Tuple[int, int], _AxisPlan] = {}


def _plan(in_size: int, out_size: int) -> _AxisPlan:
    key = (int(in_size), int(out_size))
    if key not in _PLANS:
        _PLANS[key] = _AxisPlan(*key)
    return _PLANS[key]


def imresize_normalize(img: torch.Tensor, size: Tuple[int, int], mean: Sequence[float] = MEAN, std: Sequence[float] = STD,
                       out: torch.Tensor = None) -> torch.Tensor:
    """``img_transform(imresize(img, size, 'bilinear'))`` of the reference for a decoded image on the GPU.
    img: uint8 CUDA tensor [H, W, 3] or [B, H, W, 3] (RGB, as ``np.array(PIL image)``); size = (target_width,
    target_height) like ``imresize``.  Returns float32 [B, 3, target_height, target_width] (B = 1 for a single image)."""
    require_cuda(img, "img")
    if img.dtype != torch.uint8 or img.dim() not in (3, 4) or img.shape[-1] != 3:
        raise ValueError("img must be a uint8 tensor [H, W, 3] or [B, H, W, 3]")
    x = img.contiguous()
    if x.dim() == 3:
        x = x.unsqueeze(0)
    B, H, W, _ = x.shape
    tw, th = int(size[0]), int(size[1])
    if tw < 1 or th < 1:
        raise ValueError("target size must be positive")
    px, py = _plan(W, tw), _plan(H, th)
    bx, kx = px.on(x.device)
    by, ky = py.on(x.device)
    if out is None:
        out = torch.empty(B, 3, th, tw, dtype=torch.float32, device=x.device)
    elif out.shape != (B, 3, th, tw) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
        raise ValueError("out must be a contiguous float32 [B, 3, target_height, target_width] tensor on img's device")
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    with torch.cuda.device(x.device):
        check(lib().dml_resize_bilinear_normalize(ptr(x), B, H, W, ptr(bx), ptr(kx), px.ksize, ptr(by), ptr(ky), py.ksize, th, tw,
                                                  py.max_rows_per_tile, C.cast(m, C.c_void_p), C.cast(s, C.c_void_p), ptr(out),
                                                  stream_ptr(x.device)), "dml_resize_bilinear_normalize")
    return out


def round2nearest_multiple(x: int, p: int) -> int:
    """anomaly/dataset.py:96-97"""
    return ((x - 1) // p + 1) * p


def val_target_sizes(ori_height: int, ori_width: int, img_sizes: Sequence[int], img_max_size: int,
                     padding_constant: int) -> List[Tuple[int, int]]:
    """[(target_height, target_width)] per entry of ``imgSizes`` (anomaly/dataset.py:281-289)"""
    out = []
    for this_short_size in img_sizes:
        scale = min(this_short_size / float(min(ori_height, ori_width)), img_max_size / float(max(ori_height, ori_width)))
        th, tw = int(ori_height * scale), int(ori_width * scale)
        out.append((round2nearest_multiple(th, padding_constant), round2nearest_multiple(tw, padding_constant)))
    return out


def val_image_pyramid(img: torch.Tensor, img_sizes: Sequence[int], img_max_size: int, padding_constant: int) -> List[torch.Tensor]:
    """``img_resized_list`` of ValDataset.__getitem__ for a decoded uint8 CUDA image [H, W, 3]: float32 [1, 3, h_s, w_s]"""
    H, W = int(img.shape[-3]), int(img.shape[-2])
    return [imresize_normalize(img, (tw, th)) for th, tw in val_target_sizes(H, W, img_sizes, img_max_size, padding_constant)]


class ValDataset(torch.utils.data.Dataset):
    """anomaly/dataset.py:242-322 with the resize / normalise loop on ``device``.  Same constructor (``opt`` needs
    imgSizes, imgMaxSize, padding_constant), same dict: img_ori (uint8 array), img_data (list of [1, 3, h, w] float32, on
    the GPU), seg_label ([1, H, W] int64 = label - 1), info, name."""

    def __init__(self, root_dataset, odgt, opt, rec_dataset=None, device="cuda", max_sample=-1, start_idx=-1, end_idx=-1):
        self.imgSizes = opt.imgSizes
        self.imgMaxSize = opt.imgMaxSize
        self.padding_constant = opt.padding_constant
        self.root_dataset = root_dataset
        self.rec_dataset = rec_dataset
        self.device = torch.device(device)
        # anomaly/dataset.py:39-61 (parse_input_list)
        if isinstance(odgt, list):
            self.list_sample = odgt
        elif isinstance(odgt, str):
            self.list_sample = [json.loads(x.rstrip()) for x in open(odgt, "r")][0]
        else:
            raise TypeError("odgt must be a list of records or the path of an .odgt file")
        if max_sample > 0:
            self.list_sample = self.list_sample[0:max_sample]
        if start_idx >= 0 and end_idx >= 0:
            self.list_sample = self.list_sample[start_idx:end_idx]
        self.num_sample = len(self.list_sample)
        assert self.num_sample > 0

    def __getitem__(self, index):
        from PIL import Image                      # decoding stays on the CPU (I/O: out of the hot path)
        rec = self.list_sample[index]
        if self.rec_dataset:
            folder_name, image_name = rec["fpath_img"].split("/")[-2:]
            image_path = os.path.join(self.rec_dataset, folder_name, image_name)
        else:
            image_path = os.path.join(self.root_dataset, rec["fpath_img"])
        segm = Image.open(os.path.join(self.root_dataset, rec["fpath_segm"]))
        img = Image.open(image_path).convert("RGB")
        if self.rec_dataset:
            img = img.resize(segm.size, Image.NEAREST)
        assert segm.mode == "L"
        assert img.size == segm.size
        img_ori = np.array(img)
        dev_img = torch.from_numpy(img_ori).to(self.device, non_blocking=True)
        out = dict()
        out["img_ori"] = img_ori
        out["img_data"] = val_image_pyramid(dev_img, self.imgSizes, self.imgMaxSize, self.padding_constant)
        out["seg_label"] = (torch.from_numpy(np.array(segm)).long() - 1).unsqueeze(0).contiguous()     # dataset.py:72-76
        out["info"] = rec["fpath_img"]
        out["name"] = os.path.join(*rec["fpath_img"].split("/")[-2:])
        return out

    def __len__(self):
        return self.num_sample
