"""NPM / PLM evaluation steps of the reference's DeepLab scripts on the GPU.

* ``npm_eval_batch``   -- DeepLabV3Plus-Pytorch/test_embedding.py:337-451,564 for a whole batch: argmax,
  max-softmax score, EDS score (clamp 1000, normalised, complemented), float64 novel-prototype
  override, Cityscapes label remap and the 19x19 confusion update, fused into one head pass.
* ``plm_eval_batch``   -- test_self_distillation.py:292-297,351-354,378: merge of the per-head argmax
  labels and confusion update.
* ``plm_pseudo_labels``-- test_self_distillation.py:558-570 training pseudo-label fill.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .. import head as H

_REMAP_LUT = {}


def cityscapes_remap_lut(device) -> torch.Tensor:
    """uint8 LUT of the callers' in-place label remap (test_embedding.py:448-451):
    13 -> 16 (held-out class becomes the novel id), 14.. -> shifted down by one, 254 -> 255."""
    key = str(device)
    if key not in _REMAP_LUT:
        lut = torch.arange(256, dtype=torch.int64)
        lab = lut.clone()
        lab[lab == 13] = -1
        lab[lab >= 14] -= 1
        lab[lab == -1] = 16
        lab[lab == 254] = 255
        _REMAP_LUT[key] = lab.to(torch.uint8).to(device)
    return _REMAP_LUT[key]


def remap_labels(labels: torch.Tensor) -> torch.Tensor:
    """Remapped uint8 copy of ``labels`` (values 0..255)."""
    lut = cityscapes_remap_lut(labels.device)
    return lut[labels.long()]


def npm_eval_batch(x: torch.Tensor, labels: torch.Tensor, prototype: torch.Tensor, *, confusion: torch.Tensor,
                   novel_label: int = 16, novel_thr: float = H.NOVEL_THRESHOLD, clamp: float = H.CLAMP_DEEPLAB,
                   magnitude: float = H.DEFAULT_MAGNITUDE, want_scores: bool = True, remap: bool = True):
    """x: upsampled classifier output [B,16,H,W]; labels: raw Cityscapes train ids [B,H,W];
    prototype: float64 [16] (mean of the support prototypes).  Updates ``confusion`` (int64 [19,19])
    in place and returns dict(preds uint8 [B,H,W], scores_auc_softmax, scores_auc_dis, targets)."""
    targets = remap_labels(labels) if remap else labels
    out = H.dml_head(x, magnitude=magnitude, want_logits=False, label_dtype=torch.uint8, want_eds=want_scores,
                     eds_clamp=clamp, want_msp=want_scores, want_minmax=want_scores,
                     novel=prototype.view(1, -1), novel_label_base=novel_label, novel_thr=novel_thr,
                     gt=targets, confusion=confusion)
    res = {"preds": out.label, "targets": targets}
    if want_scores:
        # scores_auc_dis = 1 - Normalization(dis_sum) (test_embedding.py:365,370); softmax score = 1 - max softmax (:341)
        eds_c, _, _ = H.finalize_scores(out.eds, None, out.minmax, want_eds=True, complement=True)
        res["scores_auc_dis"] = eds_c
        res["scores_auc_softmax"] = 1.0 - out.msp
    return res


def plm_eval_batch(xs: Sequence[torch.Tensor], labels: Optional[torch.Tensor] = None, *,
                   confusion: Optional[torch.Tensor] = None, base: int = 16,
                   magnitude: float = H.DEFAULT_MAGNITUDE, remap: bool = True) -> torch.Tensor:
    """xs[i]: upsampled output of classifier i ([B,16+i,H,W]).  Returns the merged uint8 prediction
    ``preds[argmax(head_{i+1}) == base+i] = base+i`` and (optionally) updates ``confusion``."""
    preds = H.dml_head(xs[0], magnitude=magnitude, want_logits=False, label_dtype=torch.uint8).label
    for i, x in enumerate(xs[1:]):
        lab_i = H.dml_head(x, magnitude=magnitude, want_logits=False, label_dtype=torch.uint8).label
        H.plm_merge(preds, lab_i, base + i)
    if confusion is not None and labels is not None:
        targets = remap_labels(labels) if remap else labels
        H.confusion_counts(targets, preds, confusion.shape[0], confusion.shape[1], out=confusion)
    return preds


def plm_pseudo_labels(labels: torch.Tensor, xs: Sequence[torch.Tensor], novel_cls: int = 1, base: int = 16,
                      magnitude: float = H.DEFAULT_MAGNITUDE) -> torch.Tensor:
    """test_self_distillation.py:558-570: ``labels[labels==0] = base+novel_cls-1``; ignored pixels (255)
    take the base head's argmax; earlier novel heads overwrite their own class."""
    labels = labels.clone()
    labels[labels == 0] = base + novel_cls - 1
    lab0 = H.dml_head(xs[0], magnitude=magnitude, want_logits=False, label_dtype=torch.int64).label
    fill = labels == 255
    labels[fill] = lab0[fill].to(labels.dtype)
    for i in range(novel_cls - 1):
        lab_i = H.dml_head(xs[i + 1], magnitude=magnitude, want_logits=False, label_dtype=torch.int64).label
        m = lab_i == (base + i)
        labels[m] = base + i
    return labels
