"""GPU drop-in for the score + metric block of the reference's ``evaluate()``
(anomaly/eval_ood_traditional.py:212-305,434-450,128-148,566-597), plus the batched pipeline
the B200 path is built around: one fused head pass -> key generation with the per-image
normalisation folded in -> segmented radix sort -> tie-aware scan, all images of a batch at once.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from .. import head as H
from .. import ood
from . import anom_utils
from .utils import accuracy_from_confusion, intersection_union_from_confusion

OOD_MODES = ("msp", "maxlogit", "dissum", "mmsp", "mix", "background")


def score_map(scores: torch.Tensor, mode: str = "dissum", exclude_back: bool = False, clamp: float = H.CLAMP_ANOMALY,
              lam: float = 50.0, thr: float = 0.2):
    """``conf`` map(s) [B,H,W] from (multi-scale averaged) logits ``scores`` [B,K,H,W], as the reference
    computes them for ``cfg.OOD.ood`` (eval_ood_traditional.py:276-305,434-450,468-470), and the
    argmax prediction (:218).  ``dissum`` returns the normalised EDS (what the shipped script
    evaluates, :450); ``mmsp`` / ``mix`` expose the intermediate maps of :434-448.
    Returns (pred int64 [B,H,W], conf fp32 [B,H,W])."""
    if mode not in OOD_MODES:
        raise ValueError(f"unsupported OOD.ood '{mode}' (crf / knn are outside the DML hot path)")
    if mode == "background":
        out = H.dml_head(scores, input_is_logits=True, label_dtype=torch.int64)
        first = 1 if exclude_back else 0
        return out.label, scores[:, first].contiguous()
    want_eds = mode in ("dissum", "mix")
    want_msp = mode in ("msp", "mmsp", "mix")
    out = H.dml_head(scores, input_is_logits=True, label_dtype=torch.int64, want_maxlogit=(mode == "maxlogit"),
                     want_eds=want_eds, eds_clamp=clamp, want_msp=want_msp,
                     want_minmax=mode in ("dissum", "mmsp", "mix"), exclude_back=exclude_back)
    if mode == "msp":
        return out.label, out.msp
    if mode == "maxlogit":
        return out.label, out.maxlogit
    eds_n, msp_n, mix = H.finalize_scores(out.eds, out.msp, out.minmax, want_eds=want_eds, want_msp=(mode == "mmsp"),
                                          want_mix=(mode == "mix"), lam=lam, thr=thr)
    return out.label, {"dissum": eds_n, "mmsp": msp_n, "mix": mix}[mode]


def multiscale_scores(z_list, segSize, *, want_ft_from=None, reciprocal_average: bool = False):
    """The multi-scale loop of ``evaluate()`` (anomaly/eval_ood_traditional.py:192-210) on stride-8 inputs:
    ``scores = sum_s F.interpolate(z_s, segSize) / S`` (and ``ft1`` likewise from the low-resolution
    embeddings ``want_ft_from``), each in ONE kernel that reads only the low-resolution maps.
    Returns ``scores`` [B,K,H,W] or ``(scores, ft1)``."""
    scores = H.multiscale_average(z_list, segSize, reciprocal_average=reciprocal_average)
    if want_ft_from is None:
        return scores
    return scores, H.multiscale_average(want_ft_from, segSize, reciprocal_average=reciprocal_average)


def multiscale_score_map(z_list, segSize, mode: str = "dissum", exclude_back: bool = False,
                         clamp: float = H.CLAMP_ANOMALY, lam: float = 50.0, thr: float = 0.2,
                         reciprocal_average: bool = False):
    """``score_map(multiscale_scores(z_list, segSize), mode)`` without materialising ``scores``: the
    upsample + average (:198-208) is fused into the score head (:212-305,434-450).
    Returns (pred int64 [B,H,W], conf fp32 [B,H,W])."""
    if mode not in OOD_MODES:
        raise ValueError(f"unsupported OOD.ood '{mode}' (crf / knn are outside the DML hot path)")
    if mode == "background":
        scores = H.multiscale_average(z_list, segSize, reciprocal_average=reciprocal_average)
        return score_map(scores, mode, exclude_back)
    want_eds = mode in ("dissum", "mix")
    want_msp = mode in ("msp", "mmsp", "mix")
    out = H.dml_multiscale_head(z_list, segSize, reciprocal_average=reciprocal_average, label_dtype=torch.int64,
                                want_maxlogit=(mode == "maxlogit"), want_eds=want_eds, eds_clamp=clamp,
                                want_msp=want_msp, want_minmax=mode in ("dissum", "mmsp", "mix"),
                                exclude_back=exclude_back)
    if mode == "msp":
        return out.label, out.msp
    if mode == "maxlogit":
        return out.label, out.maxlogit
    eds_n, msp_n, mix = H.finalize_scores(out.eds, out.msp, out.minmax, want_eds=want_eds, want_msp=(mode == "mmsp"),
                                          want_mix=(mode == "mix"), lam=lam, thr=thr)
    return out.label, {"dissum": eds_n, "mmsp": msp_n, "mix": mix}[mode]


def evaluate_image(segmentation_module, batch_data, cfg, reciprocal_average: bool = False):
    """One image of the reference's ``evaluate()`` loop (anomaly/eval_ood_traditional.py:183-218,276-305,434-450) on the
    fused path, with the reference's own objects: ``segmentation_module`` (``.encoder(img, return_feature_maps=True)``,
    ``.decoder`` = this repo's ``PPMDeepsup_embedding``), ``batch_data`` (the loader item: ``img_data`` = list of
    resized images, ``seg_label``) and ``cfg`` (``OOD.ood``, ``OOD.exclude_back``).  Per scale only the stride-8 logits
    are produced (``decoder.forward_lowres``: final 1x1 conv fused into the distance head); ONE kernel then upsamples,
    averages and scores -- the ten full-resolution ``F.interpolate`` / accumulate passes of the reference loop never run.
    Returns ``(pred [H,W] int64, conf [H,W] fp32)`` device tensors, i.e. what :218 / :450 hand to the metric code."""
    seg = batch_data["seg_label"][0]
    seg_size = (int(seg.shape[0]), int(seg.shape[1]))
    dev = next(segmentation_module.parameters()).device
    z_list = []
    with torch.no_grad():
        for img in batch_data["img_data"]:
            conv_out = segmentation_module.encoder(img.to(dev, non_blocking=True), return_feature_maps=True)
            z, _ = segmentation_module.decoder.forward_lowres(conv_out)
            z_list.append(z)
        pred, conf = multiscale_score_map(z_list, seg_size, cfg.OOD.ood, cfg.OOD.exclude_back,
                                          reciprocal_average=reciprocal_average)
    return pred[0], conf[0]


def eval_ood_measure(conf, seg_label, cfg, mask=None):
    """anomaly/eval_ood_traditional.py:128-148 (``cfg.OOD.out_labels``; ``mask`` filters the labels
    only -- like the reference, ``conf`` must then already be masked)."""
    out_labels = cfg.OOD.out_labels
    if mask is not None:
        seg_label = seg_label[mask]
    res = anom_utils.eval_conf_map(conf, seg_label, tuple(out_labels))
    if res is None:
        print("This image does not contain any OOD pixels or is only OOD.")
    return res


@dataclass
class BatchEval:
    """Device-resident results of one batch (no host synchronisation until ``.host()``).
    ``per_image`` / ``stats`` are private copies; ``label`` / ``conf`` / ``msp`` / ``confusion`` are the evaluator's
    reused output buffers and are only valid until its next call (clone them to keep a batch)."""
    label: torch.Tensor                 # [B,H,W] uint8 argmax prediction
    conf: Optional[torch.Tensor]        # [B,H,W] normalised EDS (the map the reference ranks), if requested
    msp: Optional[torch.Tensor]         # [B,H,W] raw max-softmax, if requested
    confusion: Optional[torch.Tensor]   # [K+1,K] int64
    per_image: torch.Tensor             # [B,7] float64 rows (auroc, aupr, fpr, n_pos, n_neg, n_nan, n_groups)
    stats: torch.Tensor                 # [B,4] int64 (n_pos, n_nan, n_out_of_window, 0)

    def host(self):
        vals, counts = ood.results_to_host(self.per_image, self.stats)
        return vals, counts


class EmbeddingEvaluator:
    """Full-resolution embedding -> labels, EDS / MMSP maps, confusion and exact per-image
    AUROC / AUPR / FPR@95 for a batch of images (BASELINE.json config 2).

    Kernel sequence per batch (all on the current stream, buffers reused):
      1. dml_head_forward : x -> label, raw EDS (clamped), raw MSP, per-image min/max, confusion
      2. dml_ood_keygen   : raw EDS -> normalised conf (optional store) -> packed keys
      3. dml_ood_eval_segments : segmented sort + group scan + finalize, one segment per image
    """

    def __init__(self, num_class: int = 13, out_labels: Sequence[int] = (13,), clamp: float = H.CLAMP_ANOMALY,
                 magnitude: float = H.DEFAULT_MAGNITUDE, want_msp: bool = True, store_conf: bool = True,
                 recall_level: float = 0.95, method: str = "sort", pos_capacity: int = ood.POS_CAPACITY_DEFAULT):
        """``method`` / ``pos_capacity``: see ``ood.eval_segments`` ("rank": the minority-rank path for images whose
        OOD pixels are few; "auto": rank with a checked fall-back to the sort path)."""
        self.K = num_class
        self.method = method
        self.pos_capacity = pos_capacity
        self.out_labels = tuple(out_labels)
        self.clamp = clamp
        self.magnitude = magnitude
        self.want_msp = want_msp
        self.store_conf = store_conf
        self.recall_level = recall_level
        self._out = None
        self._ws = None
        self._conf = None

    def __call__(self, x: torch.Tensor, gt: torch.Tensor, confusion: Optional[torch.Tensor] = None) -> BatchEval:
        B, D, Hh, Ww = x.shape
        if self._out is not None and (self._out.label.shape != (B, Hh, Ww) or self._out.label.device != x.device):
            self._out = None
        if self._ws is None or self._ws.device != x.device:
            self._ws = ood.OodWorkspace(x.device)
        out = H.dml_head(x, magnitude=self.magnitude, want_logits=False, label_dtype=torch.uint8, want_eds=True,
                         eds_clamp=self.clamp, want_msp=self.want_msp, want_minmax=True, gt=gt,
                         confusion=confusion, confusion_shape=(self.K + 1, self.K), out=self._out)
        self._out = out
        conf = None
        if self.store_conf:
            if self._conf is None or self._conf.shape != out.eds.shape or self._conf.device != x.device:
                self._conf = torch.empty_like(out.eds)
            conf = self._conf
        res, stats = ood.eval_segments(out.eds, B, Hh * Ww, gt=gt, out_labels=self.out_labels, score_kind=0,
                                       minmax=out.minmax, minmax_slot=0, conf_out=conf,
                                       recall_level=self.recall_level, workspace=self._ws, method=self.method,
                                       pos_capacity=self.pos_capacity)
        return BatchEval(out.label, conf, out.msp, out.confusion, res.clone(), stats.clone())


class MultiScaleEvaluator(EmbeddingEvaluator):
    """The anomaly sub-project's real data flow (SURVEY.md section 8 rows a1 + a2): per scale the stride-8
    embedding [B,K,h_s,w_s] of ``conv_last`` -> stride-8 distance logits (``dml_head_forward``, tiny) ->
    ONE fused kernel that bilinearly upsamples every scale to ``segSize``, averages them in the reference's
    order and emits labels, raw EDS / MSP, per-image min/max and confusion counts
    (``dml_multiscale_head_forward``) -> key-gen with the normalisation folded in -> segmented sort -> scan.
    HBM traffic per output pixel: ~10 B written (label, EDS, MSP) instead of the ~1 KB the reference moves
    for its ten full-resolution interpolations and read-modify-write accumulations."""

    def __call__(self, emb_list, gt: torch.Tensor, confusion: Optional[torch.Tensor] = None,
                 inputs_are_logits: bool = False, reciprocal_average: bool = False,
                 reference_order: bool = False) -> BatchEval:
        """``reference_order``: round the stride-8 distance logits exactly like the reference's torch-CPU ops
        (``dml_head(..., reference_order=True)``; K < 16): with the bit-exact upsample / average / EDS arithmetic of
        the fused kernel, pred and conf are then bit-identical to the CPU reference's on identical embeddings."""
        B, Hh, Ww = gt.shape
        dev = gt.device
        if inputs_are_logits:
            z_list = emb_list
        else:
            z_list = [H.dml_head(e, magnitude=self.magnitude, want_logits=True, label_dtype=None,
                                 reference_order=reference_order).logits for e in emb_list]
        if self._out is not None and (self._out.label.shape != (B, Hh, Ww) or self._out.label.device != dev):
            self._out = None
        if self._ws is None or self._ws.device != dev:
            self._ws = ood.OodWorkspace(dev)
        out = H.dml_multiscale_head(z_list, (Hh, Ww), reciprocal_average=reciprocal_average, label_dtype=torch.uint8,
                                    want_eds=True, eds_clamp=self.clamp, want_msp=self.want_msp, want_minmax=True,
                                    gt=gt, confusion=confusion, confusion_shape=(self.K + 1, self.K), out=self._out)
        self._out = out
        conf = None
        if self.store_conf:
            if self._conf is None or self._conf.shape != out.eds.shape or self._conf.device != dev:
                self._conf = torch.empty_like(out.eds)
            conf = self._conf
        res, stats = ood.eval_segments(out.eds, B, Hh * Ww, gt=gt, out_labels=self.out_labels, score_kind=0,
                                       minmax=out.minmax, minmax_slot=0, conf_out=conf,
                                       recall_level=self.recall_level, workspace=self._ws, method=self.method,
                                       pos_capacity=self.pos_capacity)
        return BatchEval(out.label, conf, out.msp, out.confusion, res.clone(), stats.clone())


def summarize(confusion: np.ndarray, per_image_vals: np.ndarray):
    """The numbers of the reference's eval summary (eval_ood_traditional.py:634-641): per-class IoU,
    mean IoU, pixel accuracy, and the MEAN over images of the per-image AUROC / AUPR / FPR
    (images with a single class are skipped, :566-572)."""
    inter, union = intersection_union_from_confusion(confusion)
    iou = inter / (union + 1e-10)
    acc, _ = accuracy_from_confusion(confusion)
    ok = ~np.isnan(per_image_vals[:, 0])
    m = per_image_vals[ok].mean(axis=0) if ok.any() else np.full(3, np.nan)
    return {"iou": iou, "mean_iou": float(iou.mean()), "accuracy": acc, "mean_auroc": float(m[0]),
            "mean_aupr": float(m[1]), "mean_fpr": float(m[2]), "n_images_scored": int(ok.sum())}
