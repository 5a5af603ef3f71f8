"""GPU drop-in for ``DeepLabV3Plus-Pytorch/metrics/stream_metrics.py``: the running confusion matrix
is accumulated on the device by ``dml_confusion`` (or fused into the head pass) instead of
``np.bincount`` on host copies; the derived scores (:57-81) are computed from the 19x19 counts."""
from __future__ import annotations

import numpy as np
import torch

from ..head import confusion_counts


class _StreamMetrics(object):
    def __init__(self):
        raise NotImplementedError()

    def update(self, gt, pred):
        raise NotImplementedError()

    def get_results(self):
        raise NotImplementedError()

    def to_str(self, metrics):
        raise NotImplementedError()

    def reset(self):
        raise NotImplementedError()


def _labels_cuda(a, device=None):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    if t.dtype not in (torch.uint8, torch.int64):
        t = t.to(torch.int64)
    if not t.is_cuda:
        t = t.cuda(device, non_blocking=True)
    return t


class StreamSegMetrics(_StreamMetrics):
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
        string = "\n"
        for k, v in results.items():
            if k != "Class IoU":
                string += "%s: %f\n" % (k, v)
        return string

    def get_results(self):
        hist = self.sync()
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.diag(hist).sum() / hist.sum()
            acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
            iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
            mean_iu = np.nanmean(iu)
            freq = hist.sum(axis=1) / hist.sum()
            fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
        print(iu)
        cls_iu = dict(zip(range(self.n_classes), iu))
        return {"Overall Acc": acc, "Mean Acc": acc_cls, "FreqW Acc": fwavacc, "Mean IoU": mean_iu, "Class IoU": cls_iu}

    def reset(self):
        self.confusion_matrix = np.zeros((self.n_classes, self.n_classes))
        if self._conf is not None:
            self._conf.zero_()


class AverageMeter(object):
    """metrics/stream_metrics.py:86-111 (host bookkeeping)."""

    def __init__(self):
        self.book = dict()

    def reset_all(self):
        self.book.clear()

    def reset(self, id):
        item = self.book.get(id, None)
        if item is not None:
            item[0] = 0
            item[1] = 0

    def update(self, id, val):
        record = self.book.get(id, None)
        if record is None:
            self.book[id] = [val, 1]
        else:
            record[0] += val
            record[1] += 1

    def get_results(self, id):
        record = self.book.get(id, None)
        assert record is not None
        return record[0] / record[1]
