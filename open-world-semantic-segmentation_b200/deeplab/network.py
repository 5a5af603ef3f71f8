"""Drop-ins for the DML model wrappers of ``DeepLabV3Plus-Pytorch/network/utils.py``.

The backbone and the DeepLabV3+ classifier heads are the caller's (cuDNN) modules and stay
untouched; what is replaced is everything after ``F.interpolate`` in ``forward`` (:89-118,
:158-193): the NHWC copy, the materialised [B,HW,K,D] difference tensor, square, sum and the
permute back -- one pass of the fused CUDA head instead.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import head as H
from ..autograd import distance_logits


def _head(x: torch.Tensor, magnitude: float):
    """(logits NCHW, centers [K,K] on x's device, features NHWC) -- the reference's return triple."""
    k = x.shape[1]
    if x.requires_grad and torch.is_grad_enabled():
        logits = distance_logits(x, magnitude=magnitude)
        feats = x.permute(0, 2, 3, 1).contiguous()
    else:
        out = H.dml_head(x, magnitude=magnitude, want_logits=True, label_dtype=None, want_features=True)
        logits, feats = out.logits, out.features
    centers = torch.eye(k, device=x.device, dtype=x.dtype) * magnitude
    return logits, centers, feats


class _SimpleSegmentationModel_embedding(nn.Module):
    """network/utils.py:56-118.  ``forward(x) -> (logits[B,K,H,W], centers[K,K], features[B,H,W,K])``."""

    def __init__(self, backbone, classifier, magnitude: float = H.DEFAULT_MAGNITUDE):
        super().__init__()
        self.backbone = backbone
        self.classifier = classifier
        self.magnitude = magnitude
        self.centers = torch.eye(17) * magnitude   # plain attribute, refreshed per forward like the reference (:103-106)

    def forward(self, x):
        input_shape = x.shape[-2:]
        features = self.backbone(x)
        x = self.classifier(features)
        x = F.interpolate(x, size=input_shape, mode='bilinear', align_corners=False)
        logits, centers, feats = _head(x, self.magnitude)
        self.centers = torch.eye(x.shape[1]) * self.magnitude
        return logits, centers, feats


class _SimpleSegmentationModel_embedding_self_distillation(nn.Module):
    """network/utils.py:120-193 (PLM): one backbone pass, one distance head per classifier
    (base ``classifier`` with 16 classes, ``classifier_<i>`` with 16+i).  Returns three LISTS
    (logits, centers, features).  The reference builds its ``DeepLabHeadV3Plus`` heads inside the
    constructor; here they are passed in (``classifiers[0]`` is the base head) so that the untouched
    reference / torchvision heads can be used as they are."""

    def __init__(self, backbone, classifiers, magnitude: float = H.DEFAULT_MAGNITUDE):
        super().__init__()
        self.backbone = backbone
        self.cls_novel = len(classifiers) - 1
        self.classifier_list = ['classifier'] + ['classifier_' + str(i + 1) for i in range(self.cls_novel)]
        for name, mod in zip(self.classifier_list, classifiers):
            self.__setattr__(name, mod)
        self.magnitude = magnitude
        self.centers = torch.zeros(17, 17)

    def forward_single(self, classifier, features, input_shape):
        x = classifier(features)
        x = F.interpolate(x, size=input_shape, mode='bilinear', align_corners=False)
        return _head(x, self.magnitude)

    def forward(self, x):
        input_shape = x.shape[-2:]
        features = self.backbone(x)
        logits, centers, feats = [], [], []
        for name in self.classifier_list:
            l, c, f = self.forward_single(self.__getattr__(name), features, input_shape)
            logits.append(l)
            centers.append(c)
            feats.append(f)
        return logits, centers, feats
