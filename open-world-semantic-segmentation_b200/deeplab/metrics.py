"""GPU drop-in for ``DeepLabV3Plus-Pytorch/metrics/stream_metrics.py``: the running confusion matrix
is accumulated on the device by ``dml_confusion`` (or fused into the head pass) instead of
``np.bincount`` on host copies; the derived scores (:57-81) are computed from the 19x19 counts."""
from __future__ import annotations

import numpy as np
import torch

from ..head import confusion_counts


def _labels_cuda(a, device=None):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    if t.dtype not in (torch.uint8, torch.int64):
        t = t.to(torch.int64)
    if not t.is_cuda:
        t = t.cuda(device, non_blocking=True)
    return t


class StreamSegMetrics:
    """Same interface as the reference: ``update(label_trues, label_preds)``, ``get_results()`` ->
    dict(Overall Acc, Mean Acc, FreqW Acc, Mean IoU, Class IoU), ``to_str``, ``reset``.
    Like the reference, ``n_classes`` is pinned to 19 regardless of the constructor argument (:30)."""

    def __init__(self, n_classes, device=None):
        self.n_classes = 19
        self.device = device
        self._conf = None
        self.confusion_matrix = np.zeros((n_classes, n_classes))

    def _device_conf(self, dev):
        if self._conf is None or self._conf.device != dev:
            self._conf = torch.zeros(self.n_classes, self.n_classes, dtype=torch.int64, device=dev)
        return self._conf

    def update(self, label_trues, label_preds):
        lt = _labels_cuda(label_trues, self.device)
        lp = _labels_cuda(label_preds, lt.device)
        confusion_counts(lt, lp, self.n_classes, self.n_classes, out=self._device_conf(lt.device))

    def accumulator(self, device) -> torch.Tensor:
        """int64 [19,19] device accumulator, for fusing the counts into ``dml_head(gt=..., confusion=...)``."""
        return self._device_conf(torch.device(device))

    def sync(self):
        if self._conf is not None:
            self.confusion_matrix = self._conf.cpu().numpy().astype(np.float64)
        return self.confusion_matrix

    @staticmethod
    def to_str(results):
        """one ``name: value`` line per scalar entry (the per-class dict is skipped), stream_metrics.py:36-47"""
        return "\n" + "".join("%s: %f\n" % (k, v) for k, v in results.items() if k != "Class IoU")

    def get_results(self):
        """Overall / mean-class / frequency-weighted accuracy and IoU from the 19 x 19 counts (rows = ground truth);
        classes that never occur give NaN and are skipped by the nan-means, as in stream_metrics.py:57-81."""
        hist = self.sync()
        tp = np.diag(hist)
        gt_n, pred_n, total = hist.sum(axis=1), hist.sum(axis=0), hist.sum()
        with np.errstate(divide="ignore", invalid="ignore"):
            iu = tp / (gt_n + pred_n - tp)
            res = {"Overall Acc": tp.sum() / total, "Mean Acc": np.nanmean(tp / gt_n)}
            freq = gt_n / total
            seen = freq > 0
            res["FreqW Acc"] = (freq[seen] * iu[seen]).sum()
            res["Mean IoU"] = np.nanmean(iu)
        print(iu)                                  # the reference prints the per-class IoU vector (:76)
        res["Class IoU"] = dict(zip(range(self.n_classes), iu))
        return res

    def reset(self):
        self.confusion_matrix = np.zeros((self.n_classes, self.n_classes))
        if self._conf is not None:
            self._conf.zero_()
