"""Drop-in for the reference's embedding decoder ``PPMDeepsup_embedding``
(anomaly/models/models.py:586-687).  The cuDNN layers (PPM pooling branches, ``conv_last``,
deep-supervision branch) are ordinary ``torch.nn`` modules with the reference's parameter
names, so its checkpoints load unchanged; the distance block (:636-657) is the CUDA head.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import head as H
from ..autograd import distance_logits

BatchNorm2d = nn.BatchNorm2d  # the reference aliases its vendored SynchronizedBatchNorm2d (same state_dict keys)


def conv3x3_bn_relu(in_planes, out_planes, stride=1):
    """anomaly/models/models.py conv3x3_bn_relu helper (3x3 conv, BN, ReLU)."""
    return nn.Sequential(nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False),
                         BatchNorm2d(out_planes), nn.ReLU(inplace=True))


class PPMDeepsup_embedding(nn.Module):
    """Same constructor / forward contract as the reference (:587-588,620):

    eval (``use_softmax=True``): ``forward(conv_out, segSize, output_ft=True)`` ->
        ``(logits[B,K,H,W], embedding_upsampled[B,K,H,W])`` or just ``logits``;
    train: ``((logits_lowres, deepsup), ft)`` or ``(logits_lowres, deepsup)``.
    Prototypes are ``3 * I`` sized from ``num_class`` (the reference hard-codes 13, :614)."""

    def __init__(self, num_class=150, fc_dim=4096, use_softmax=False, pool_scales=(1, 2, 3, 6), magnitude=3.0):
        super().__init__()
        self.use_softmax = use_softmax
        self.ppm = nn.ModuleList([
            nn.Sequential(nn.AdaptiveAvgPool2d(scale), nn.Conv2d(fc_dim, 512, kernel_size=1, bias=False),
                          BatchNorm2d(512), nn.ReLU(inplace=True)) for scale in pool_scales])
        self.cbr_deepsup = conv3x3_bn_relu(fc_dim // 2, fc_dim // 4, 1)
        self.conv_last = nn.Sequential(
            nn.Conv2d(fc_dim + len(pool_scales) * 512, 512, kernel_size=3, padding=1, bias=False),
            BatchNorm2d(512), nn.ReLU(inplace=True), nn.Dropout2d(0.1), nn.Conv2d(512, num_class, kernel_size=1))
        self.conv_last_deepsup = nn.Conv2d(fc_dim // 4, num_class, 1, 1, 0)
        self.dropout_deepsup = nn.Dropout2d(0.1)
        self.magnitude = magnitude
        self.centers = torch.eye(num_class) * magnitude   # plain attribute like the reference (:614)

    def forward_lowres(self, conv_out):
        """Stride-8 ``(logits, embedding)`` of :620-657, i.e. the eval branch WITHOUT the two full-resolution
        ``F.interpolate`` calls of :659-669.  Feed the per-scale results to
        ``dml_b200.anomaly.eval_ood.multiscale_scores`` / ``MultiScaleEvaluator``, which upsample, average
        and score inside one kernel."""
        conv5 = conv_out[-1]
        size = conv5.shape[2:]
        ppm_out = torch.cat([conv5] + [nn.functional.interpolate(p(conv5), size, mode='bilinear', align_corners=False)
                                       for p in self.ppm], 1)
        if not torch.is_grad_enabled() and ppm_out.is_cuda and ppm_out.dtype == torch.float32 and not self.training:
            # inference: the final 1x1 conv (:609) is fused into the distance head -- the 512-channel
            # feature is read once, embedding and logits leave the same kernel
            feat = self.conv_last[:4](ppm_out)
            emb, z = H.conv1x1_head(feat, self.conv_last[4].weight, self.conv_last[4].bias, self.magnitude)
            return z, emb
        emb = self.conv_last(ppm_out)
        return distance_logits(emb, magnitude=self.magnitude), emb

    def forward(self, conv_out, segSize=None, output_ft=True):
        conv5 = conv_out[-1]
        size = conv5.shape[2:]
        ppm_out = [conv5] + [nn.functional.interpolate(p(conv5), size, mode='bilinear', align_corners=False)
                             for p in self.ppm]
        ppm_out = torch.cat(ppm_out, 1)
        ft = ppm_out
        emb = self.conv_last(ppm_out)
        x = distance_logits(emb, magnitude=self.magnitude)          # CUDA head (:636-657)
        if self.use_softmax:  # inference
            x = nn.functional.interpolate(x, size=segSize, mode='bilinear', align_corners=False)
            if output_ft:
                emb_up = nn.functional.interpolate(emb, size=segSize, mode='bilinear', align_corners=False)
                return x, emb_up
            return x
        conv4 = conv_out[-2]
        d = self.conv_last_deepsup(self.dropout_deepsup(self.cbr_deepsup(conv4)))
        if output_ft:
            return (x, d), ft.clone()
        return (x, d)
