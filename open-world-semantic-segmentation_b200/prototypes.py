"""Few-shot novel-prototype generation (kernel (c)): masked per-class feature means.

Mirrors the recipe of DeepLabV3Plus-Pytorch/test_embedding.py:413-425 (per support image, the
mean feature over the pixels of the novel class when it covers more than 5 % of the image) and the
prototype averaging of :245-258.  The segmented reduction runs in libdml_b200.so
(``dml_class_sums``); only the tiny [n_cls, D] results visit the host.
"""
from __future__ import annotations

import json
from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import check, lib, ptr, require_cuda, stream_ptr


def class_sums(x: torch.Tensor, labels: torch.Tensor, n_cls: int, nhwc: bool = False):
    """Per-image, per-class float64 feature sums and pixel counts.

    x: [B,D,H,W] (``nhwc=False``) or [B,H,W,D] (``nhwc=True``) fp32 CUDA; labels [B,H,W] uint8/int64
    (labels outside [0, n_cls) are skipped).  Returns (sums [B,n_cls,D] float64, counts [B,n_cls] int64)."""
    require_cuda(x, "x")
    require_cuda(labels, "labels")
    if x.dim() != 4 or not x.is_floating_point():
        raise ValueError("x must be a 4-D floating-point tensor")
    x = x.float().contiguous()               # the kernel reads const float*
    labels = labels.contiguous()
    if labels.dtype not in (torch.uint8, torch.int64):
        labels = labels.to(torch.int64)
    if nhwc:
        B, H, W, D = x.shape
    else:
        B, D, H, W = x.shape
    if tuple(labels.shape) != (B, H, W):
        raise ValueError("labels must be [B,H,W]")
    dev = x.device
    hw = H * W
    sums = torch.empty(B, n_cls, D, dtype=torch.float64, device=dev)
    counts = torch.empty(B, n_cls, dtype=torch.int64, device=dev)
    ws = torch.empty(lib().dml_class_sums_workspace_bytes(B, D, n_cls, hw), dtype=torch.uint8, device=dev)
    u8 = labels.dtype == torch.uint8
    with torch.cuda.device(dev):
        check(lib().dml_class_sums(ptr(x), 1 if nhwc else 0, ptr(labels) if u8 else None, None if u8 else ptr(labels),
                                   B, D, hw, n_cls, ptr(ws), ptr(sums), ptr(counts), stream_ptr(dev)), "dml_class_sums")
    return sums, counts


def class_means(x, labels, n_cls, nhwc=False):
    sums, counts = class_sums(x, labels, n_cls, nhwc)
    return sums / counts.clamp_min(1).unsqueeze(-1).to(torch.float64), counts


def novel_prototypes(features: torch.Tensor, labels: torch.Tensor, cls: int, n_cls: int = 19, min_frac: float = 0.05,
                     nhwc: bool = True) -> List[Optional[np.ndarray]]:
    """One entry per support image: the mean feature of class ``cls`` if it covers more than
    ``min_frac`` of ALL pixels of the image (np.unique counts every label value, test_embedding.py:413-415),
    else None."""
    sums, counts = class_sums(features, labels, n_cls, nhwc)
    total = labels[0].numel()
    s, c = sums[:, cls].cpu().numpy(), counts[:, cls].cpu().numpy()
    out = []
    for i in range(len(c)):
        if c[i] > 0 and c[i] / total > min_frac:
            out.append(s[i] / c[i])
        else:
            out.append(None)
    return out


def prototype_mean(prototypes: Sequence) -> np.ndarray:
    """float64 mean of the stored support prototypes (test_embedding.py:255-258)."""
    acc = np.zeros((len(prototypes[0]),))
    for p in prototypes:
        acc += np.array(p)
    acc /= len(prototypes)
    return acc


def save_prototypes(path: str, prototypes: Sequence):
    """JSON list-of-lists, the reference's on-disk format (test_embedding.py:421-425)."""
    with open(path, "w") as fh:
        json.dump([np.asarray(p).tolist() for p in prototypes], fh)


def load_prototype(path: str) -> np.ndarray:
    with open(path, "r") as fh:
        return prototype_mean(json.load(fh))
