"""GPU drop-in for the counting helpers of the reference's ``anomaly/utils.py``
(``accuracy`` :128-133, ``intersectionAndUnion`` :136-156).  Both are views of one confusion
matrix that the kernel accumulates in a single pass over (gt, pred)."""
from __future__ import annotations

import numpy as np
import torch

from ..head import confusion_counts


def _cuda_labels(a) -> torch.Tensor:
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    if t.dtype not in (torch.uint8, torch.int64):
        t = t.to(torch.int64)
    return t if t.is_cuda else t.cuda(non_blocking=True)


def confusion_matrix(preds, label, num_class: int, n_gt: int | None = None) -> torch.Tensor:
    """int64 [n_gt, num_class] counts on the device; gt < 0 (unlabelled) is skipped, rows up to
    ``n_gt`` (default num_class + 1, so that the OOD label ``num_class`` keeps its own row)."""
    n_gt = n_gt or num_class + 1
    return confusion_counts(_cuda_labels(label), _cuda_labels(preds), n_gt, num_class)


def accuracy_from_confusion(conf: np.ndarray):
    k = conf.shape[1]
    acc_sum = np.trace(conf[:k, :k])
    valid_sum = conf.sum()
    return float(acc_sum) / (valid_sum + 1e-10), valid_sum


def intersection_union_from_confusion(conf: np.ndarray):
    k = conf.shape[1]
    inter = np.diag(conf[:k, :k]).copy()
    area_pred = conf.sum(axis=0)            # predictions on every labelled pixel (incl. rows >= k)
    area_lab = conf[:k].sum(axis=1)
    return inter, area_pred + area_lab - inter


def accuracy(preds, label):
    """anomaly/utils.py:128-133 -> (acc, valid_pixel_count)."""
    lab = _cuda_labels(label)
    n_gt = int(lab.max().item()) + 1 if lab.numel() else 1
    pr = _cuda_labels(preds)
    n_pr = int(pr.max().item()) + 1 if pr.numel() else 1
    n = max(n_gt, n_pr, 1)
    conf = confusion_counts(lab, pr, n, n).cpu().numpy()
    acc_sum = np.trace(conf)
    valid_sum = conf.sum()
    return float(acc_sum) / (valid_sum + 1e-10), valid_sum


def intersectionAndUnion(imPred, imLab, numClass):
    """anomaly/utils.py:136-156 -> (area_intersection, area_union), each int64 [numClass]."""
    lab = _cuda_labels(imLab)
    n_gt = max(int(lab.max().item()) + 1 if lab.numel() else 1, numClass)
    conf = confusion_counts(lab, _cuda_labels(imPred), n_gt, numClass).cpu().numpy()
    return intersection_union_from_confusion(conf)
