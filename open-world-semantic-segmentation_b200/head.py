"""Fused distance + score head (kernel (a)) -- functional API over ``dml_head_forward``.

Host-side mirror of what the reference computes after the last 1x1 conv of its decoders:
anomaly/models/models.py:636-657, DeepLabV3Plus-Pytorch/network/utils.py:89-118 and the
score lines of anomaly/eval_ood_traditional.py:218,276-305,434 /
DeepLabV3Plus-Pytorch/test_embedding.py:339-350,428-445.  PyTorch here only owns memory
and streams; all arithmetic happens in libdml_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import HeadParams, MultiscaleParams, check, lib, ptr, require_cuda, stream_ptr

DEFAULT_MAGNITUDE = 3.0          # anomaly/models/models.py:615, network/utils.py:104
CLAMP_ANOMALY = 400.0            # anomaly/eval_ood_traditional.py:304
CLAMP_DEEPLAB = 1000.0           # DeepLabV3Plus-Pytorch/test_embedding.py:350
NOVEL_THRESHOLD = -1.5           # DeepLabV3Plus-Pytorch/test_embedding.py:445


@dataclass
class HeadOutput:
    logits: Optional[torch.Tensor] = None        # [B,K,H,W] fp32
    label: Optional[torch.Tensor] = None         # [B,H,W] uint8 or int64
    maxlogit: Optional[torch.Tensor] = None      # [B,H,W]
    eds: Optional[torch.Tensor] = None           # [B,H,W] raw (clamped) distance sum
    msp: Optional[torch.Tensor] = None           # [B,H,W] max softmax
    features: Optional[torch.Tensor] = None      # [B,H,W,D] NHWC copy
    novel_dist: Optional[torch.Tensor] = None    # [n_novel,B,H,W] float64
    minmax: Optional[torch.Tensor] = None        # [B,4] (eds_min, eds_max, msp_min, msp_max)
    confusion: Optional[torch.Tensor] = None     # [rows, cols] int64 counts


def scaled_identity_magnitude(centers: torch.Tensor) -> Optional[float]:
    """If ``centers`` (a HOST tensor) is ``m * I`` return m, else None.  Device tensors are
    never inspected (that would synchronise): they take the dense-prototype path."""
    if centers.is_cuda or centers.dim() != 2 or centers.shape[0] != centers.shape[1]:
        return None
    k = centers.shape[0]
    m = float(centers[0, 0])
    if torch.equal(centers, torch.eye(k, dtype=centers.dtype) * centers[0, 0]):
        return m
    return None


def dml_head(x: torch.Tensor, centers: Optional[torch.Tensor] = None, magnitude: float = DEFAULT_MAGNITUDE, *,
             input_is_logits: bool = False, want_logits: bool = True, label_dtype: Optional[torch.dtype] = torch.uint8,
             want_maxlogit: bool = False, want_eds: bool = False, eds_clamp: float = 0.0,
             want_msp: bool = False, want_features: bool = False, want_minmax: bool = False,
             exclude_back: bool = False, novel: Optional[torch.Tensor] = None, novel_label_base: int = 16,
             novel_thr: float = NOVEL_THRESHOLD, want_novel_dist: bool = False,
             gt: Optional[torch.Tensor] = None, confusion: Optional[torch.Tensor] = None,
             confusion_shape: Optional[tuple] = None, out: Optional[HeadOutput] = None,
             reference_order: bool = False) -> HeadOutput:
    """One pass over ``x`` [B,D,H,W] (fp32, CUDA, contiguous NCHW).

    input_is_logits: ``x`` already holds the logits z [B,K,H,W] (anomaly path: stride-8 distances
             upsampled and averaged by the caller); only labels / scores are produced.
    centers: None -> ``magnitude * I`` (K = D, cancellation-free fast path);
             host tensor equal to m*I -> same fast path; anything else -> dense [K,D] prototypes.
    novel:   [n_novel, D] float64 novel prototypes (NPM); labels are overridden with
             ``novel_label_base + j`` where z_nov > novel_thr and z_nov > max_k z_k.
    gt / confusion: fused confusion counts ``confusion[gt, label] += 1`` (gt outside
             [0, rows) ignored); ``confusion`` is an int64 [rows, cols] accumulator (created
             from ``confusion_shape`` when None).
    ``out`` lets callers reuse output buffers (CUDA-graph friendly).
    reference_order: parity mode (``centers`` = m*I, D < 16, logits requested): every logit is rounded exactly like
             the reference's torch-CPU op sequence (anomaly/models/models.py:649-651), so the logits -- and everything
             derived from them -- are bit-identical to the reference's on identical embeddings.
    """
    require_cuda(x, "x")
    if x.dtype != torch.float32 or x.dim() != 4:
        raise ValueError("x must be a float32 [B,D,H,W] tensor")
    x = x.contiguous()
    B, D, H, W = x.shape
    dev = x.device
    mu = None
    if centers is not None:
        m = scaled_identity_magnitude(centers)
        if m is not None and centers.shape[0] == D:
            magnitude = m
        else:
            mu = centers.to(device=dev, dtype=torch.float32).contiguous()
            if mu.dim() != 2 or mu.shape[1] != D:
                raise ValueError(f"centers must be [K,{D}]")
    K = D if mu is None else mu.shape[0]
    o = out or HeadOutput()

    def buf(cur, shape, dtype):
        if cur is not None:
            if tuple(cur.shape) != tuple(shape) or cur.dtype != dtype or cur.device != dev:
                raise ValueError("preallocated output has the wrong shape/dtype/device")
            return cur
        return torch.empty(shape, dtype=dtype, device=dev)

    p = HeadParams()
    p.struct_bytes = C.sizeof(HeadParams)
    p.B, p.D, p.K, p.H, p.W = B, D, K, H, W
    p.x = x.data_ptr()
    p.mu = mu.data_ptr() if mu is not None else None
    p.diag_m = magnitude
    p.input_is_logits = 1 if input_is_logits else 0
    if input_is_logits:
        want_logits = False
    p.score_first_class = 1 if exclude_back else 0
    p.eds_clamp = eds_clamp
    if reference_order:
        if mu is not None or input_is_logits or not want_logits or D >= 16:
            raise ValueError("reference_order needs m*I prototypes, D < 16 and want_logits=True")
        p.reference_order = 1
    if novel is not None:
        novel = novel.to(device=dev, dtype=torch.float64).contiguous().view(-1, D)
        p.mu_novel, p.n_novel = novel.data_ptr(), novel.shape[0]
        p.novel_label_base, p.novel_thr = novel_label_base, novel_thr
        if want_novel_dist:
            o.novel_dist = buf(o.novel_dist, (novel.shape[0], B, H, W), torch.float64)
            p.novel_dist = o.novel_dist.data_ptr()
    if want_logits:
        o.logits = buf(o.logits, (B, K, H, W), torch.float32)
        p.logits = o.logits.data_ptr()
    if label_dtype is not None:
        if label_dtype not in (torch.uint8, torch.int64):
            raise ValueError("label_dtype must be torch.uint8, torch.int64 or None")
        o.label = buf(o.label, (B, H, W), label_dtype)
        if label_dtype == torch.uint8:
            p.label_u8 = o.label.data_ptr()
        else:
            p.label_i64 = o.label.data_ptr()
    if want_maxlogit:
        o.maxlogit = buf(o.maxlogit, (B, H, W), torch.float32)
        p.maxlogit = o.maxlogit.data_ptr()
    if want_eds:
        o.eds = buf(o.eds, (B, H, W), torch.float32)
        p.eds = o.eds.data_ptr()
    if want_msp:
        o.msp = buf(o.msp, (B, H, W), torch.float32)
        p.msp = o.msp.data_ptr()
    if want_features:
        o.features = buf(o.features, (B, H, W, D), torch.float32)
        p.features_nhwc = o.features.data_ptr()
    if want_minmax:
        o.minmax = buf(o.minmax, (B, 4), torch.float32)
        p.minmax = o.minmax.data_ptr()
        p.want_eds_minmax = 1 if want_eds else 0
        p.want_msp_minmax = 1 if want_msp else 0
    if gt is not None:
        require_cuda(gt, "gt")
        gt = gt.contiguous()
        if tuple(gt.shape) != (B, H, W) or gt.dtype not in (torch.uint8, torch.int64):
            raise ValueError("gt must be a [B,H,W] uint8 or int64 tensor")
        if confusion is None:
            if o.confusion is not None:
                confusion = o.confusion
            else:
                rows, cols = confusion_shape or (K + 1, K if novel is None else max(K, novel_label_base + novel.shape[0]))
                confusion = torch.zeros(rows, cols, dtype=torch.int64, device=dev)
        if confusion.dtype != torch.int64 or confusion.dim() != 2 or not confusion.is_contiguous():
            raise ValueError("confusion must be a contiguous int64 [rows, cols] tensor")
        o.confusion = confusion
        p.confusion = confusion.data_ptr()
        p.conf_rows, p.conf_cols = confusion.shape
        if gt.dtype == torch.uint8:
            p.gt_u8 = gt.data_ptr()
        else:
            p.gt_i64 = gt.data_ptr()
    with torch.cuda.device(dev):
        check(lib().dml_head_forward(C.byref(p), stream_ptr(dev)), "dml_head_forward")
    return o


def conv1x1_head(features: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 magnitude: float = DEFAULT_MAGNITUDE, want_embedding: bool = True, want_logits: bool = True):
    """Final 1x1 classifier conv + distance block in one kernel (``dml_conv1x1_head_forward``):
    ``features`` [B,C,h,w] fp32 CUDA (the input of ``conv_last[4]`` / ``classifier[3]``), ``weight`` [K,C] or
    [K,C,1,1], ``bias`` [K] or None.  Returns ``(embedding [B,K,h,w] or None, logits [B,K,h,w] or None)`` --
    anomaly/models/models.py:609,636-657; inference only (no autograd)."""
    require_cuda(features, "features")
    if features.dtype != torch.float32 or features.dim() != 4:
        raise ValueError("features must be a float32 [B,C,h,w] tensor")
    features = features.contiguous()
    B, Cc, Hh, Ww = features.shape
    dev = features.device
    w = weight.detach().to(device=dev, dtype=torch.float32).reshape(weight.shape[0], -1).contiguous()
    if w.shape[1] != Cc:
        raise ValueError(f"weight must be [K,{Cc}] (or [K,{Cc},1,1])")
    K = w.shape[0]
    bvec = None if bias is None else bias.detach().to(device=dev, dtype=torch.float32).contiguous()
    if bvec is not None and bvec.numel() != K:
        raise ValueError("bias must have K entries")
    emb = torch.empty(B, K, Hh, Ww, dtype=torch.float32, device=dev) if want_embedding else None
    logits = torch.empty(B, K, Hh, Ww, dtype=torch.float32, device=dev) if want_logits else None
    with torch.cuda.device(dev):
        check(lib().dml_conv1x1_head_forward(ptr(features), ptr(w), ptr(bvec), magnitude, B, Cc, K, Hh, Ww, ptr(emb), ptr(logits),
                                             stream_ptr(dev)), "dml_conv1x1_head_forward")
    return emb, logits


def dml_multiscale_head(z_list, size, *, reciprocal_average: bool = False, want_scores: bool = False,
                        label_dtype: Optional[torch.dtype] = torch.uint8, want_maxlogit: bool = False,
                        want_eds: bool = False, eds_clamp: float = 0.0, want_msp: bool = False,
                        want_minmax: bool = False, exclude_back: bool = False, gt: Optional[torch.Tensor] = None,
                        confusion: Optional[torch.Tensor] = None, confusion_shape: Optional[tuple] = None,
                        out: Optional[HeadOutput] = None) -> HeadOutput:
    """Fused multi-scale upsample + average + score head (``dml_multiscale_head_forward``).

    ``z_list``: per scale the stride-8 logits [B,K,h_s,w_s] (fp32, CUDA); ``size`` = (H, W) = ``segSize``.
    Equivalent to the reference loop ``scores += F.interpolate(z_s, segSize, 'bilinear',
    align_corners=False) / len(z_list)`` (anomaly/models/models.py:659-661,
    anomaly/eval_ood_traditional.py:192-208) followed by the score lines (:212-218,276-305,434), but no
    full-resolution [B,K,H,W] tensor is read or -- unless ``want_scores`` -- written.
    ``reciprocal_average``: multiply by fl(1/S) like torch's CUDA division-by-scalar kernel instead of
    the correctly rounded division torch performs on the CPU.
    ``out.logits`` holds the averaged ``scores`` when ``want_scores``."""
    if not 1 <= len(z_list) <= _lib.MAX_SCALES:
        raise ValueError(f"between 1 and {_lib.MAX_SCALES} scales are supported")
    zs = []
    for z in z_list:
        require_cuda(z, "z")
        if z.dtype != torch.float32 or z.dim() != 4:
            raise ValueError("every scale must be a float32 [B,K,h,w] tensor")
        zs.append(z.contiguous())
    B, K = zs[0].shape[:2]
    dev = zs[0].device
    for z in zs:
        if z.shape[0] != B or z.shape[1] != K or z.device != dev:
            raise ValueError("scales differ in batch size, class count or device")
    H, W = int(size[0]), int(size[1])
    o = out or HeadOutput()

    def buf(cur, shape, dtype):
        if cur is not None:
            if tuple(cur.shape) != tuple(shape) or cur.dtype != dtype or cur.device != dev:
                raise ValueError("preallocated output has the wrong shape/dtype/device")
            return cur
        return torch.empty(shape, dtype=dtype, device=dev)

    p = MultiscaleParams()
    p.struct_bytes = C.sizeof(MultiscaleParams)
    p.B, p.K, p.H, p.W = B, K, H, W
    p.n_scales = len(zs)
    for s, z in enumerate(zs):
        p.z[s] = z.data_ptr()
        p.h[s], p.w[s] = z.shape[2], z.shape[3]
    p.reciprocal_average = 1 if reciprocal_average else 0
    p.score_first_class = 1 if exclude_back else 0
    p.eds_clamp = eds_clamp
    if want_scores:
        o.logits = buf(o.logits, (B, K, H, W), torch.float32)
        p.scores = o.logits.data_ptr()
    if label_dtype is not None:
        if label_dtype not in (torch.uint8, torch.int64):
            raise ValueError("label_dtype must be torch.uint8, torch.int64 or None")
        o.label = buf(o.label, (B, H, W), label_dtype)
        if label_dtype == torch.uint8:
            p.label_u8 = o.label.data_ptr()
        else:
            p.label_i64 = o.label.data_ptr()
    if want_maxlogit:
        o.maxlogit = buf(o.maxlogit, (B, H, W), torch.float32)
        p.maxlogit = o.maxlogit.data_ptr()
    if want_eds:
        o.eds = buf(o.eds, (B, H, W), torch.float32)
        p.eds = o.eds.data_ptr()
    if want_msp:
        o.msp = buf(o.msp, (B, H, W), torch.float32)
        p.msp = o.msp.data_ptr()
    if want_minmax:
        o.minmax = buf(o.minmax, (B, 4), torch.float32)
        p.minmax = o.minmax.data_ptr()
        p.want_eds_minmax = 1 if want_eds else 0
        p.want_msp_minmax = 1 if want_msp else 0
    if gt is not None:
        require_cuda(gt, "gt")
        gt = gt.contiguous()
        if tuple(gt.shape) != (B, H, W) or gt.dtype not in (torch.uint8, torch.int64):
            raise ValueError("gt must be a [B,H,W] uint8 or int64 tensor")
        if confusion is None:
            if o.confusion is not None:
                confusion = o.confusion
            else:
                rows, cols = confusion_shape or (K + 1, K)
                confusion = torch.zeros(rows, cols, dtype=torch.int64, device=dev)
        if confusion.dtype != torch.int64 or confusion.dim() != 2 or not confusion.is_contiguous():
            raise ValueError("confusion must be a contiguous int64 [rows, cols] tensor")
        o.confusion = confusion
        p.confusion = confusion.data_ptr()
        p.conf_rows, p.conf_cols = confusion.shape
        if gt.dtype == torch.uint8:
            p.gt_u8 = gt.data_ptr()
        else:
            p.gt_i64 = gt.data_ptr()
    with torch.cuda.device(dev):
        check(lib().dml_multiscale_head_forward(C.byref(p), stream_ptr(dev)), "dml_multiscale_head_forward")
    return o


def multiscale_average(x_list, size, *, reciprocal_average: bool = False, out: Optional[torch.Tensor] = None):
    """``sum_s F.interpolate(x_s, size, 'bilinear', align_corners=False) / S`` in the reference's order
    (the ``ft1`` accumulator of anomaly/eval_ood_traditional.py:194-196,209-210): [B,C,H,W]."""
    o = HeadOutput(logits=out)
    return dml_multiscale_head(x_list, size, reciprocal_average=reciprocal_average, want_scores=True,
                               label_dtype=None, out=o).logits


def finalize_scores(eds: Optional[torch.Tensor], msp: Optional[torch.Tensor], minmax: torch.Tensor, *,
                    want_eds: bool = True, want_msp: bool = False, want_mix: bool = False, lam: float = 50.0,
                    thr: float = 0.2, complement: bool = False, out_eds: Optional[torch.Tensor] = None,
                    out_msp: Optional[torch.Tensor] = None, out_mix: Optional[torch.Tensor] = None):
    """Per-image min-max normalisation (+ EDS/MMSP mix) of raw score maps [B,H,W]
    (anomaly/eval_ood_traditional.py:101-106,305,435,447-448).  Returns (eds_n, msp_n, mix)."""
    ref = eds if eds is not None else msp
    require_cuda(ref, "scores")
    B = ref.shape[0]
    hw = ref[0].numel()
    dev = ref.device
    eds_n = (out_eds if out_eds is not None else torch.empty_like(eds)) if (want_eds and eds is not None) else None
    msp_n = (out_msp if out_msp is not None else torch.empty_like(msp)) if (want_msp and msp is not None) else None
    mix = (out_mix if out_mix is not None else torch.empty_like(eds)) if want_mix else None
    with torch.cuda.device(dev):
        check(lib().dml_scores_finalize(ptr(eds), ptr(msp), ptr(minmax), B, hw, lam, thr, 1 if complement else 0,
                                        ptr(eds_n), ptr(msp_n), ptr(mix), stream_ptr(dev)), "dml_scores_finalize")
    return eds_n, msp_n, mix


def confusion_counts(gt: torch.Tensor, pred: torch.Tensor, rows: int, cols: int,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[g, p] += #pixels`` for gt in [0, rows), pred in [0, cols)
    (DeepLabV3Plus-Pytorch/metrics/stream_metrics.py:49-55; anomaly/utils.py:128-156)."""
    require_cuda(gt, "gt")
    require_cuda(pred, "pred")
    gt, pred = gt.contiguous(), pred.contiguous()
    if gt.numel() != pred.numel():
        raise ValueError("gt and pred differ in size")
    for t in (gt, pred):
        if t.dtype not in (torch.uint8, torch.int64):
            raise ValueError("labels must be uint8 or int64")
    dev = gt.device
    if out is None:
        out = torch.zeros(rows, cols, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib().dml_confusion(ptr(gt) if gt.dtype == torch.uint8 else None, ptr(gt) if gt.dtype == torch.int64 else None,
                                  ptr(pred) if pred.dtype == torch.uint8 else None,
                                  ptr(pred) if pred.dtype == torch.int64 else None,
                                  gt.numel(), rows, cols, ptr(out), stream_ptr(dev)), "dml_confusion")
    return out


def plm_merge(base: torch.Tensor, head: torch.Tensor, novel_label: int) -> torch.Tensor:
    """In place ``base[head == novel_label] = novel_label``
    (DeepLabV3Plus-Pytorch/test_self_distillation.py:292-297)."""
    require_cuda(base, "base")
    if base.dtype != head.dtype or base.dtype not in (torch.uint8, torch.int64) or base.numel() != head.numel():
        raise ValueError("base/head must be same-size uint8 or int64 tensors")
    if not base.is_contiguous():
        raise ValueError("base must be contiguous (it is updated in place)")
    head = head.contiguous()
    dev = base.device
    u8 = base.dtype == torch.uint8
    with torch.cuda.device(dev):
        check(lib().dml_plm_merge(ptr(base) if u8 else None, None if u8 else ptr(base), ptr(head) if u8 else None,
                                  None if u8 else ptr(head), base.numel(), novel_label, stream_ptr(dev)), "dml_plm_merge")
    return base
