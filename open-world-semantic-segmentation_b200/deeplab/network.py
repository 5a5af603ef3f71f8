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


def _reference_head_factory():
    """``DeepLabHeadV3Plus`` of the caller's reference tree (network/utils.py:9-53, an untouched cuDNN module): the
    reference's self-distillation wrapper builds its heads from it inside the constructor (network/utils.py:121-135)."""
    import importlib
    for mod in ("network.utils", "network._deeplab"):
        try:
            return getattr(importlib.import_module(mod), "DeepLabHeadV3Plus")
        except Exception:
            continue
    raise ImportError("DeepLabHeadV3Plus not importable: construct the model inside the reference's DeepLabV3Plus-Pytorch tree, "
                      "or pass `classifiers=[...]` / `head_factory=` explicitly")


class _SimpleSegmentationModel_embedding_self_distillation(nn.Module):
    """network/utils.py:120-193 (PLM): one backbone pass, one distance head per classifier
    (base ``classifier`` with 16 classes, ``classifier_<i>`` with 16+i).  Returns three LISTS
    (logits, centers, features).

    Constructor parity: ``cls(backbone)`` like the reference (network/utils.py:121, called from
    network/modeling.py:40) -- the ``DeepLabHeadV3Plus`` heads (2048 / 256 input planes, ASPP rates 6-12-18,
    16 and 16 + i classes) are then built from the reference's own head class (``head_factory``, default: looked up
    in the caller's ``network`` package).  ``classifiers=[base, novel_1, ...]`` passes ready-made heads instead."""

    def __init__(self, backbone, classifiers=None, magnitude: float = H.DEFAULT_MAGNITUDE, head_factory=None, cls_novel: int = 1,
                 num_classes: int = 16, inplanes: int = 2048, low_level_planes: int = 256, aspp_dilate=(6, 12, 18)):
        super().__init__()
        self.backbone = backbone
        if classifiers is None:
            factory = head_factory or _reference_head_factory()
            classifiers = [factory(inplanes, low_level_planes, num_classes + i, list(aspp_dilate)) for i in range(cls_novel + 1)]
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


# ---- factories of network/modeling.py:140-158 -------------------------------------------------------------------
def _reference_pieces():
    import importlib
    net = importlib.import_module("network")
    util = importlib.import_module("network.utils")
    resnet = importlib.import_module("network.backbone.resnet")
    return net, util, resnet


def _segm_resnet_embedding(name, backbone_name, num_classes, output_stride, pretrained_backbone):
    """network/modeling.py:7-44 for the two embedding variants: the reference's ResNet backbone, ``IntermediateLayerGetter``
    and ``DeepLabHeadV3Plus`` (all untouched cuDNN modules, imported from the caller's reference tree) under the fused
    DML wrappers of this file."""
    _, util, resnet = _reference_pieces()
    if output_stride == 8:
        replace, aspp_dilate = [False, True, True], [12, 24, 36]
    else:
        replace, aspp_dilate = [False, False, True], [6, 12, 18]
    backbone = resnet.__dict__[backbone_name](pretrained=pretrained_backbone, replace_stride_with_dilation=replace)
    backbone = util.IntermediateLayerGetter(backbone, return_layers={'layer4': 'out', 'layer1': 'low_level'})
    if name == 'deeplabv3plus_embedding':
        return _SimpleSegmentationModel_embedding(backbone, util.DeepLabHeadV3Plus(2048, 256, num_classes, aspp_dilate))
    # (the reference ignores num_classes / output_stride for the heads of this variant: 16 (+ i) classes, rates 6-12-18)
    return _SimpleSegmentationModel_embedding_self_distillation(backbone, head_factory=util.DeepLabHeadV3Plus)


def deeplabv3plus_embedding_resnet101(num_classes=21, output_stride=8, pretrained_backbone=True):
    """network/modeling.py:140-148"""
    return _segm_resnet_embedding('deeplabv3plus_embedding', 'resnet101', num_classes, output_stride, pretrained_backbone)


def deeplabv3plus_embedding_self_distillation_resnet101(num_classes=21, output_stride=8, pretrained_backbone=True):
    """network/modeling.py:150-158"""
    return _segm_resnet_embedding('deeplabv3plus_embedding_self_distillation', 'resnet101', num_classes, output_stride,
                                  pretrained_backbone)
