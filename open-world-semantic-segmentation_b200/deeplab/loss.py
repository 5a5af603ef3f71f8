"""Drop-ins for ``DeepLabV3Plus-Pytorch/utils/loss.py``.

``CrossEntropyLoss.forward(logit, target, features_in)`` keeps the reference signature.  In the
shipped reference an early ``return CE_loss / n`` (:41-42) makes the VL / Inter / Center terms dead
code; ``shipped_early_return=True`` (default) reproduces exactly that, ``False`` evaluates the
intended line-79 form ``(CE + alpha*VL + beta*Inter + gamma*Center) / n`` with the fused kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..autograd import dml_loss
from ..prototypes import class_sums


class FocalLoss(nn.Module):
    """utils/loss.py:7-23 (softmax baseline, plain torch; not on the DML path)."""

    def __init__(self, alpha=1, gamma=0, size_average=True, ignore_index=255):
        super().__init__()
        self.alpha, self.gamma, self.ignore_index, self.size_average = alpha, gamma, ignore_index, size_average

    def forward(self, inputs, targets):
        ce_loss = F.cross_entropy(inputs, targets, reduction='none', ignore_index=self.ignore_index)
        pt = torch.exp(-ce_loss)
        focal_loss = self.alpha * (1 - pt) ** self.gamma * ce_loss
        return focal_loss.mean() if self.size_average else focal_loss.sum()


def center_term(features_in: torch.Tensor, target: torch.Tensor, ignore_index: int, n_cls: int) -> torch.Tensor:
    """Center = sum_i (1/T_i) sum_c sum_{p in i, y_p = c} ||f_p - mean_c f||^2 (utils/loss.py:65-68 intent;
    the literal code indexes [H,W,C] features with flat pixel ids).  Per-class sums
    come from the segmented-reduction kernel; sum ||f||^2 - ||sum f||^2 / n_c closes the form."""
    b = features_in.shape[0]
    d = features_in.shape[-1]
    f = features_in.reshape(b, -1, d)
    t = target.reshape(b, -1)
    sums, counts = class_sums(features_in.detach().reshape(b, 1, -1, d).contiguous(), t.reshape(b, 1, -1), n_cls, nhwc=True)
    means = (sums / counts.clamp_min(1).unsqueeze(-1).double()).to(f.dtype)        # [B, n_cls, D]
    valid = (t != ignore_index) & (t >= 0) & (t < n_cls)
    idx = torch.where(valid, t, torch.zeros_like(t)).long()
    mu = torch.gather(means, 1, idx.unsqueeze(-1).expand(-1, -1, d))
    sq = ((f - mu) ** 2).sum(-1) * valid.to(f.dtype)
    return sq.sum() / float(t.shape[1])


class CrossEntropyLoss(nn.Module):
    """utils/loss.py:25-82."""

    def __init__(self, alpha=0, beta=0, gamma=0, size_average=True, ignore_index=255, shipped_early_return=True):
        super().__init__()
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.ignore_index = ignore_index
        self.size_average = size_average
        self.shipped_early_return = shipped_early_return

    def _ce_over_n(self, logit, target):
        """``nn.CrossEntropyLoss(ignore_index, size_average)(logit, target) / n`` (utils/loss.py:39-42): the kernel
        averages over the valid pixels; ``size_average=False`` (a sum in the reference) is that mean times the valid count,
        taken from the same kernel's float64 parts vector."""
        if self.size_average:
            return dml_loss(logit, target, alpha=0.0, beta=0.0, ignore_index=self.ignore_index, input_is_logits=True)
        loss, parts = dml_loss(logit, target, alpha=0.0, beta=0.0, ignore_index=self.ignore_index, input_is_logits=True,
                               return_parts=True)
        return loss * parts[4].to(loss.dtype)

    def forward(self, logit, target, features_in=None):
        if self.shipped_early_return:
            return self._ce_over_n(logit, target)
        if not self.size_average:
            raise NotImplementedError("size_average=False is only defined for the shipped CE / n form")
        loss = dml_loss(logit, target, alpha=self.alpha, beta=self.beta, ignore_index=self.ignore_index,
                        input_is_logits=True)
        if self.gamma != 0 and features_in is not None:
            loss = loss + self.gamma * center_term(features_in, target, self.ignore_index, logit.shape[1]) / logit.shape[0]
        return loss


class CrossEntropyLoss_dis(nn.Module):
    """utils/loss.py:84-122: shipped state is ``CE / n`` (:102); the distillation term
    (0.01 * mean squared feature drift on non-novel pixels, :104-118) is available with
    ``shipped_early_return=False``."""

    def __init__(self, alpha=0, beta=0, gamma=0, size_average=True, ignore_index=255, shipped_early_return=True,
                 novel_label=16):
        super().__init__()
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.ignore_index, self.size_average = ignore_index, size_average
        self.shipped_early_return = shipped_early_return
        self.novel_label = novel_label

    def forward(self, logit, target, features_1=None, features_2=None):
        n = logit.shape[0]
        ce_n = CrossEntropyLoss._ce_over_n(self, logit, target)
        if self.shipped_early_return:
            return ce_n
        pad = torch.zeros(*features_1.shape[:3], 1, device=features_1.device, dtype=features_1.dtype)
        f1 = torch.cat((features_1, pad), dim=3)
        dis = logit.new_zeros(())
        for i in range(n):
            keep = target[i] != self.novel_label
            diff = features_2[i][keep] - f1[i][keep]
            dis = dis + torch.sum(diff ** 2) / diff.shape[0]
        return ce_n + 0.01 * dis / n
