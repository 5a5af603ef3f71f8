"""Softmax-baseline OOD evaluation of the reference's DeepLab scripts on the GPU:
DeepLabV3Plus-Pytorch/test.py:179-248 -- ``1 - max softmax`` anomaly scores and, per image, scikit-learn's
``roc_auc_score`` / ``average_precision_score`` plus ``fpr95 = fpr[tpr >= 0.95][0]`` on ``roc_curve``.

The ranking runs on the same radix sort + tie-aware scan as the DML metrics (``dml_ood_eval_segments``); the
``roc_curve`` FPR convention (first KEPT point of the drop_intermediate ROC curve whose recall reaches the level)
is ``dml_ood_roc_fpr``.  No CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import head as H
from .. import ood


def softmax_scores(outputs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """test.py:181-183: ``preds = outputs.max(1)[1]`` (int64) and ``scores = 1 - softmax(outputs).max(1)[0]``
    for logits ``outputs`` [B,K,H,W], one fused pass (no [B,K,H,W] softmax is materialised)."""
    out = H.dml_head(outputs, input_is_logits=True, label_dtype=torch.int64, want_msp=True)
    return out.label, 1.0 - out.msp


def roc_measures(scores: torch.Tensor, labels: torch.Tensor, labels_true: Optional[torch.Tensor] = None,
                 ood_label: int = 255, recall_level: float = 0.95,
                 workspace: Optional[ood.OodWorkspace] = None):
    """test.py:205-248 for one image (the script runs at batch size 1): pixels with ``labels_true != 255`` are
    evaluated, positives are ``labels == 255`` (the held-out classes after the dataset's remap).
    Returns ``(auc, aupr, fpr95)`` or ``None`` when the image has no positive pixel (``if 1 in instance``)."""
    s = scores.reshape(-1)
    lab = labels.reshape(-1)
    if labels_true is not None:
        keep = labels_true.reshape(-1) != 255
        s, lab = s[keep], lab[keep]
    pos = lab == ood_label
    if s.numel() == 0 or not bool(pos.any()):
        return None
    if bool(pos.all()):
        # every evaluated pixel is positive: sklearn's roc_auc_score (test.py:241) raises here
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    return ood.measures_from_scores(s.contiguous(), pos, recall_level=recall_level, workspace=workspace,
                                    fpr_convention="roc_curve")
