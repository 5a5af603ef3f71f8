"""CPU oracle for the DMLNet per-pixel metric-learning hot path.

TEST INFRASTRUCTURE ONLY.  This module is a CPU restatement (torch-CPU / numpy) of
the reference's algorithm for the hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker (or the timed CPU baseline) -- never
as part of the product path.  The product (``dml_b200``) never imports it.

Parity pin: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF, generated in the build container by ``tests/golden/make_golden.py``
(which imports the unmodified reference from /root/reference) and committed
under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every
function here against them.  AUROC / AUPR come from scikit-learn in the
reference (``requirements.txt:111`` pins 0.24.1; the container has 1.9.0);
the restatement below follows the published ``_binary_clf_curve`` /
``roc_curve`` / ``average_precision_score`` algorithm and is validated against
the installed scikit-learn on the golden vectors.

Citations ``path:line`` are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

RECALL_LEVEL_DEFAULT = 0.95  # anomaly/anom_utils.py:4


# --------------------------------------------------------------------------- #
# (a) distance head
# --------------------------------------------------------------------------- #
def make_centers(num_classes: int, magnitude: float = 3.0) -> torch.Tensor:
    """Fixed prototypes ``magnitude * I_K``.

    anomaly/models/models.py:614-618 (13x13), DeepLabV3Plus-Pytorch/network/utils.py:103-106
    (rebuilt per forward from the channel count).
    """
    c = torch.zeros(num_classes, num_classes)
    for i in range(num_classes):
        c[i][i] = magnitude
    return c


def distance_logits(x: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """``z[b,k,h,w] = -sum_d (x[b,d,h,w] - centers[k,d])**2`` in the reference's op order.

    anomaly/models/models.py:636-657 and DeepLabV3Plus-Pytorch/network/utils.py:89-117:
    NCHW -> NHWC copy, broadcast against the prototypes, subtract, square, sum over
    the embedding dim, negate, permute back to NCHW.  Materialises [B,HW,K,D] exactly
    like the reference does (that cost is part of the CPU baseline).
    """
    b, d, h, w = x.shape
    k = centers.shape[0]
    feats = x.permute(0, 2, 3, 1).contiguous().view(b, h * w, d)
    feats = feats.unsqueeze(2).expand(b, h * w, k, d)
    diff = feats - centers
    z = -torch.sum(diff ** 2, 3)
    return z.permute(0, 2, 1).contiguous().view(b, k, h, w)


def distance_logits_f64(x: torch.Tensor, centers: torch.Tensor) -> torch.Tensor:
    """Float64 ground truth of the same quantity (chunk-friendly einsum-free form)."""
    x64 = x.double()
    c64 = centers.double()
    out = torch.empty(x.shape[0], c64.shape[0], x.shape[2], x.shape[3], dtype=torch.float64)
    for k in range(c64.shape[0]):
        out[:, k] = -((x64 - c64[k].view(1, -1, 1, 1)) ** 2).sum(1)
    return out


def features_nhwc(x: torch.Tensor) -> torch.Tensor:
    """The contiguous NHWC copy the DeepLab heads return as ``features``
    (DeepLabV3Plus-Pytorch/network/utils.py:92-93,118)."""
    return x.permute(0, 2, 3, 1).contiguous()


def ppm_head_eval(x_low: torch.Tensor, centers: torch.Tensor, seg_size):
    """Eval branch of ``PPMDeepsup_embedding.forward`` after ``conv_last``
    (anomaly/models/models.py:636-669): stride-8 distance logits, then bilinear
    (align_corners=False) upsampling of logits AND raw embedding to ``seg_size``."""
    z = distance_logits(x_low, centers)
    z_up = F.interpolate(z, size=seg_size, mode="bilinear", align_corners=False)
    f_up = F.interpolate(x_low, size=seg_size, mode="bilinear", align_corners=False)
    return z_up, f_up


def multiscale_scores(x_lows, centers: torch.Tensor, seg_size):
    """Multi-scale accumulation ``scores += z_up / n_scales``
    (anomaly/eval_ood_traditional.py:192-210)."""
    n = len(x_lows)
    k = centers.shape[0]
    b = x_lows[0].shape[0]
    scores = torch.zeros(b, k, seg_size[0], seg_size[1])
    ft = torch.zeros(b, x_lows[0].shape[1], seg_size[0], seg_size[1])
    for x_low in x_lows:
        z_up, f_up = ppm_head_eval(x_low, centers, seg_size)
        scores = scores + z_up / n
        f_up = F.interpolate(f_up, size=ft.shape[2:], mode="bilinear", align_corners=False)
        ft = ft + f_up / n
    return scores, ft


def argmax_label(scores: torch.Tensor) -> np.ndarray:
    """``pred = argmax_k scores`` (first maximal index)
    (anomaly/eval_ood_traditional.py:218-219, DeepLabV3Plus-Pytorch/test_embedding.py:339)."""
    return scores.max(dim=1)[1].cpu().numpy()


# --------------------------------------------------------------------------- #
# (a) OOD scores
# --------------------------------------------------------------------------- #
def normalization(x: np.ndarray) -> np.ndarray:
    """Per-image min-max normalise, dtype preserved (fp32 in the reference).
    anomaly/eval_ood_traditional.py:101-102, DeepLabV3Plus-Pytorch/test_embedding.py:150-153."""
    lo = np.min(x)
    hi = np.max(x)
    return (x - lo) / (hi - lo)


def coefficient_map(x: np.ndarray, thre: float, lamda: float = 50) -> np.ndarray:
    """Sigmoid gate ``1 / (1 + exp(lamda (x - thre)))`` (anomaly/eval_ood_traditional.py:104-106)."""
    return 1 / (1 + np.exp(lamda * (x - thre)))


def _tmp_scores(scores: torch.Tensor, exclude_back: bool) -> torch.Tensor:
    # anomaly/eval_ood_traditional.py:212-214
    return scores[:, 1:] if exclude_back else scores


def score_msp(scores: torch.Tensor, exclude_back: bool = False) -> np.ndarray:
    """``--ood msp`` (anomaly/eval_ood_traditional.py:276-278)."""
    conf, _ = torch.max(F.softmax(_tmp_scores(scores, exclude_back), dim=1), dim=1)
    return conf.squeeze(0).cpu().numpy()


def score_maxlogit(scores: torch.Tensor, exclude_back: bool = False) -> np.ndarray:
    """``--ood maxlogit`` (anomaly/eval_ood_traditional.py:288-290)."""
    conf, _ = torch.max(_tmp_scores(scores, exclude_back), dim=1)
    return conf.squeeze(0).cpu().numpy()


def score_background(scores: torch.Tensor, exclude_back: bool = False) -> np.ndarray:
    """``--ood background`` (anomaly/eval_ood_traditional.py:468-470)."""
    return _tmp_scores(scores, exclude_back)[:, 0].squeeze(0).cpu().numpy()


def score_dissum(scores: torch.Tensor, clamp: float = 400.0, exclude_back: bool = False) -> np.ndarray:
    """EDS, ``--ood dissum`` (anomaly/eval_ood_traditional.py:301-305): negative sum of
    the logits over classes, clamp at 400, per-image min-max normalise."""
    tmp = _tmp_scores(scores, exclude_back)
    dis_sum = torch.sum(tmp, dim=1)
    dis_sum = -dis_sum.squeeze(0).cpu().numpy()
    dis_sum[dis_sum >= clamp] = clamp
    return normalization(dis_sum)


def score_mmsp(scores: torch.Tensor, exclude_back: bool = False) -> np.ndarray:
    """MMSP (anomaly/eval_ood_traditional.py:434-435): normalised max-softmax."""
    tmp = _tmp_scores(scores, exclude_back)
    prob_map = np.max(F.softmax(tmp, dim=1).squeeze().cpu().numpy(), axis=0)
    return normalization(prob_map)


def score_mix(dis_sum: np.ndarray, prob_map: np.ndarray, thre: float = 0.2) -> np.ndarray:
    """EDS/MMSP mix (anomaly/eval_ood_traditional.py:447-448).  The shipped script then
    overwrites it with ``conf = dis_sum`` (:450); callers choose."""
    c = coefficient_map(dis_sum, thre)
    return c * dis_sum + (1 - c) * prob_map


def deeplab_scores(outputs: torch.Tensor, clamp: float = 1000.0):
    """Score block of DeepLab ``validate`` for one image
    (DeepLabV3Plus-Pytorch/test_embedding.py:339-351): returns
    (preds[1,H,W] int64, scores_auc_softmax[H,W], dis_sum_map_norm[H,W], scores_auc_dis[H,W])."""
    preds = outputs.detach().max(dim=1)[1].cpu().numpy()
    soft = F.softmax(outputs, dim=1)
    scores_auc_softmax = (1 - soft.detach().max(dim=1)[0].cpu().numpy()).squeeze()
    dis_sum_map = -np.sum(outputs.squeeze().cpu().numpy(), axis=0)
    dis_sum_map[dis_sum_map > clamp] = clamp
    dis_norm = normalization(dis_sum_map)
    return preds, scores_auc_softmax, dis_norm, 1 - dis_norm


# --------------------------------------------------------------------------- #
# (a) NPM / PLM
# --------------------------------------------------------------------------- #
def novel_prototype(prototypes) -> np.ndarray:
    """Mean of the stored support prototypes, float64
    (DeepLabV3Plus-Pytorch/test_embedding.py:255-258)."""
    protos = [np.array(p) for p in prototypes]
    acc = np.zeros((len(protos[0]),))
    for p in protos:
        acc += p
    acc /= len(protos)
    return acc


def npm_override(preds: np.ndarray, outputs: torch.Tensor, features: torch.Tensor,
                 prototype: np.ndarray, novel_label: int = 16, thr: float = -1.5):
    """Novel-prototype override for one image (batch 1)
    (DeepLabV3Plus-Pytorch/test_embedding.py:428-433,445).  ``features`` is the NHWC
    tensor the head returned; the novel distance is float64 (numpy promotion).
    Returns (preds (modified copy), dis_novel[H,W] float64)."""
    b, h, w, c = features.shape
    f = features.view(b, h * w, c).squeeze().cpu().numpy()
    dis = -np.sum((f - prototype) ** 2, axis=1)
    dis = dis.reshape(h, w)
    preds = preds.copy()
    maxlogit = outputs.detach().max(dim=1)[0].squeeze().cpu().numpy()
    preds[0][np.logical_and(dis > thr, dis > maxlogit)] = novel_label
    return preds, dis


def remap_labels_cityscapes(labels: torch.Tensor) -> torch.Tensor:
    """In-place-style label remap done by the DeepLab callers
    (DeepLabV3Plus-Pytorch/test_embedding.py:448-451)."""
    labels = labels.clone()
    labels[labels == 13] = -1
    labels[labels >= 14] -= 1
    labels[labels == -1] = 16
    labels[labels == 254] = 255
    return labels


def plm_merge(outputs_list, novel_cls: int = 1, base: int = 16) -> torch.Tensor:
    """PLM eval merge (DeepLabV3Plus-Pytorch/test_self_distillation.py:292-297)."""
    preds_base = outputs_list[0].detach().max(dim=1)[1]
    for i in range(novel_cls):
        labels_i = outputs_list[i + 1].detach().max(dim=1)[1]
        preds_base[labels_i == (base + i)] = base + i
    return preds_base


def plm_pseudo_labels(labels: torch.Tensor, outputs_list, novel_cls: int = 1, base: int = 16) -> torch.Tensor:
    """PLM training pseudo-label fill (DeepLabV3Plus-Pytorch/test_self_distillation.py:558-570)."""
    labels = labels.clone()
    labels[labels == 0] = base + novel_cls - 1
    labels_base = outputs_list[0].detach().max(dim=1)[1]
    labels[labels == 255] = labels_base[labels == 255]
    for i in range(novel_cls - 1):
        labels_base = outputs_list[i + 1].detach().max(dim=1)[1]
        labels[labels_base == (base + i)] = labels_base[labels_base == (base + i)]
    return labels


def masked_class_mean(features: np.ndarray, labels: np.ndarray, cls: int, min_frac: float = 0.05):
    """Few-shot prototype generation for one support image
    (DeepLabV3Plus-Pytorch/test_embedding.py:413-419 recipe): if class ``cls`` covers
    more than ``min_frac`` of the pixels, the mean feature over its pixels, else None.
    ``features`` [H,W,D], ``labels`` [H,W]."""
    inst, counts = np.unique(labels, False, False, True)
    if cls in inst:
        if counts[np.where(inst == cls)] / np.sum(counts) > min_frac:
            return np.mean(features[labels == cls], axis=0)
    return None


def per_class_sums(features: np.ndarray, labels: np.ndarray, n_cls: int):
    """Float64 per-class feature sums and counts (ground truth for the segmented reduce).
    ``features`` [N,D], ``labels`` [N]."""
    d = features.shape[1]
    sums = np.zeros((n_cls, d), dtype=np.float64)
    cnt = np.zeros((n_cls,), dtype=np.int64)
    for c in range(n_cls):
        m = labels == c
        cnt[c] = int(m.sum())
        if cnt[c]:
            sums[c] = features[m].astype(np.float64).sum(0)
    return sums, cnt


# --------------------------------------------------------------------------- #
# (b) loss
# --------------------------------------------------------------------------- #
def dml_loss(logit: torch.Tensor, target: torch.Tensor, features_in=None, *, alpha=0.0, beta=0.0,
             gamma=0.0, ignore_index=255, shipped_early_return=False) -> torch.Tensor:
    """DCE + VL (+ Inter + Center) loss, restated without the per-(image,class) host loop.

    Full intended form: DeepLabV3Plus-Pytorch/utils/loss.py:34-82 (line 79),
    anomaly form anomaly/models/models.py:42-78 is ``beta = gamma = 0, ignore_index = -1``.
    ``shipped_early_return=True`` reproduces the shipped DeepLab state (``return CE/n``,
    utils/loss.py:41-42).

      CE    = mean over valid pixels of (logsumexp_k z - z_y)
      VL    = sum_i (1/T_i) sum_{valid p in i} (-z_{p,y_p})
      Inter = sum_i (1/T_i) sum_{valid p in i} sum_{k != y_p} z_{p,k}
      Center= sum_i (1/T_i) sum_c sum_{p in i, y_p = c} ||f_p - mean_c f||^2
      T_i   = all pixels of image i INCLUDING ignored ones (np.unique counts, loss.py:55-57)
      loss  = (CE + alpha VL + beta Inter + gamma Center) / n
    """
    n, c, h, w = logit.shape
    ce = F.cross_entropy(logit, target.long(), ignore_index=ignore_index, reduction="mean")
    if shipped_early_return:
        return ce / n
    t_i = float(h * w)
    valid = target != ignore_index
    tgt = torch.where(valid, target, torch.zeros_like(target)).long()
    z_y = torch.gather(logit, 1, tgt.unsqueeze(1)).squeeze(1)
    vf = valid.to(logit.dtype)
    var_loss = (-(z_y) * vf).sum() / t_i
    inter = ((logit.sum(1) - z_y) * vf).sum() / t_i
    center = logit.new_zeros(())
    if gamma != 0 and features_in is not None:
        for i in range(n):
            f = features_in[i].reshape(h * w, -1)
            lab = target[i].reshape(-1)
            for cls in torch.unique(lab).tolist():
                if cls == ignore_index:
                    continue
                fc = f[lab == cls]
                center = center + ((fc - fc.mean(0)) ** 2).sum() / t_i
    return (ce + alpha * var_loss + beta * inter + gamma * center) / n


# --------------------------------------------------------------------------- #
# (d) exact ranking metrics
# --------------------------------------------------------------------------- #
def stable_cumsum(arr, rtol=1e-05, atol=1e-08):
    """float64 cumsum with a drift check (anomaly/anom_utils.py:7-23)."""
    out = np.cumsum(arr, dtype=np.float64)
    expected = np.sum(arr, dtype=np.float64)
    if not np.allclose(out[-1], expected, rtol=rtol, atol=atol):
        raise RuntimeError("cumsum was found to be unstable: its last element does not correspond to sum")
    return out


def binary_clf_curve(y_true: np.ndarray, y_score: np.ndarray):
    """Distinct-threshold cumulative counts (the scheme shared by
    anomaly/anom_utils.py:39-53 and scikit-learn's ``_binary_clf_curve``): stable sort
    descending by score, thresholds at the last index of each run of equal scores.
    Returns (fps, tps, thresholds) as float64/float64/score dtype."""
    order = np.argsort(y_score, kind="mergesort")[::-1]
    s = y_score[order]
    y = y_true[order]
    distinct = np.where(np.diff(s))[0]
    idx = np.r_[distinct, y.size - 1]
    tps = stable_cumsum(y)[idx]
    fps = 1 + idx - tps
    return fps, tps, s[idx]


def fpr_and_fdr_at_recall(y_true, y_score, recall_level=RECALL_LEVEL_DEFAULT, pos_label=None):
    """FPR at the threshold whose recall is closest to ``recall_level``
    (anomaly/anom_utils.py:25-65), scanning from the lowest threshold with full recall
    downwards in index and taking the first minimum."""
    classes = np.unique(y_true)
    ok = any(np.array_equal(classes, c) for c in ([0, 1], [-1, 1], [0], [-1], [1]))
    if pos_label is None and not ok:
        raise ValueError("Data is not binary and pos_label is not specified")
    elif pos_label is None:
        pos_label = 1.0
    y_bool = (y_true == pos_label)
    fps, tps, thr = binary_clf_curve(y_bool, y_score)
    recall = tps / tps[-1]
    last = tps.searchsorted(tps[-1])
    sl = slice(last, None, -1)
    recall, fps = np.r_[recall[sl], 1], np.r_[fps[sl], 0]
    cutoff = np.argmin(np.abs(recall - recall_level))
    return fps[cutoff] / (np.sum(np.logical_not(y_bool)))


def roc_auc(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """Restatement of ``sklearn.metrics.roc_auc_score`` (binary): trapezoid of
    tpr over fpr on the distinct-threshold ROC with a (0,0) origin
    (call site anomaly/anom_utils.py:74)."""
    if len(np.unique(y_true)) != 2:
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    fps, tps, _ = binary_clf_curve(y_true == 1, y_score)
    tps = np.r_[0.0, tps]
    fps = np.r_[0.0, fps]
    fpr = fps / fps[-1]
    tpr = tps / tps[-1]
    return float(np.trapezoid(tpr, fpr))


def average_precision(y_true: np.ndarray, y_score: np.ndarray) -> float:
    """Restatement of ``sklearn.metrics.average_precision_score`` (binary):
    ``sum_g (R_g - R_{g-1}) * P_g`` (call site anomaly/anom_utils.py:75)."""
    fps, tps, _ = binary_clf_curve(y_true == 1, y_score)
    ps = tps + fps
    precision = np.where(ps != 0, tps / np.where(ps != 0, ps, 1), 0.0)
    recall = tps / tps[-1]
    precision = np.r_[precision[::-1], 1.0]
    recall = np.r_[recall[::-1], 0.0]
    return float(max(0.0, -np.sum(np.diff(recall) * precision[:-1])))


def roc_curve_drop_intermediate(y_true: np.ndarray, y_score: np.ndarray):
    """``sklearn.metrics.roc_curve(y_true, y_score)`` with its default ``drop_intermediate=True`` as the
    softmax-baseline evaluator calls it (DeepLabV3Plus-Pytorch/test.py:242): points of the distinct-threshold
    curve whose second differences of fps AND tps vanish (collinear with both neighbours) are dropped, the
    first and last are always kept, then (0, 0) is prepended.  scikit-learn is not vendored in the reference
    (pinned 0.24.1, requirements.txt:111; same rule in the installed 1.9.0, _ranking.py roc_curve).
    Returns (fpr, tpr) float64."""
    fps, tps, _ = binary_clf_curve(y_true, y_score)
    if len(fps) > 2:
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        fps, tps = fps[keep], tps[keep]
    tps = np.r_[0, tps]
    fps = np.r_[0, fps]
    fpr = fps / fps[-1] if fps[-1] > 0 else np.repeat(np.nan, fps.shape)
    tpr = tps / tps[-1] if tps[-1] > 0 else np.repeat(np.nan, tps.shape)
    return fpr, tpr


def baseline_roc_measures(msk_auc: np.ndarray, scores_auc: np.ndarray, recall_level: float = 0.95):
    """(auc, aupr, fpr95) of the softmax-baseline evaluator, DeepLabV3Plus-Pytorch/test.py:241-244:
    ``roc_auc_score``, ``average_precision_score`` and ``fpr95 = fpr[tpr >= 0.95][0]`` on ``roc_curve``
    -- the FIRST kept ROC point whose recall reaches the level (not the closest-recall rule of
    anomaly/anom_utils.py:57-65)."""
    y = np.asarray(msk_auc).astype(np.int64).ravel()
    s = np.asarray(scores_auc).ravel()
    fpr, tpr = roc_curve_drop_intermediate(y, s)
    return roc_auc(y, s), average_precision(y, s), float(fpr[tpr >= recall_level][0])


def _check_finite(a: np.ndarray):
    if not np.all(np.isfinite(a)):
        raise ValueError("Input contains NaN or infinity.")  # sklearn's check_array behaviour


def get_measures(_pos, _neg, recall_level=RECALL_LEVEL_DEFAULT, use_sklearn=False):
    """(AUROC, AUPR, FPR@recall) for positive / negative score lists
    (anomaly/anom_utils.py:67-78).  ``use_sklearn`` routes AUROC/AUPR through the
    installed scikit-learn exactly as the reference does (used for the timed CPU
    baseline and to validate the numpy restatement)."""
    pos = np.array(_pos[:]).reshape((-1, 1))
    neg = np.array(_neg[:]).reshape((-1, 1))
    examples = np.squeeze(np.vstack((pos, neg)))
    labels = np.zeros(len(examples), dtype=np.int32)
    labels[:len(pos)] += 1
    if use_sklearn:
        import sklearn.metrics as sk
        auroc = sk.roc_auc_score(labels, examples)
        aupr = sk.average_precision_score(labels, examples)
    else:
        _check_finite(examples)
        auroc = roc_auc(labels, examples)
        aupr = average_precision(labels, examples)
    fpr = fpr_and_fdr_at_recall(labels, examples, recall_level)
    return auroc, aupr, fpr


def get_and_print_results(out_score, in_score, num_to_avg=1, use_sklearn=False):
    """anomaly/anom_utils.py:95-104 (the mean over a single measurement)."""
    m = get_measures(out_score, in_score, use_sklearn=use_sklearn)
    return float(np.mean([m[0]])), float(np.mean([m[1]])), float(np.mean([m[2]]))


def eval_ood_measure(conf: np.ndarray, seg_label: np.ndarray, out_labels=(13,), mask=None, use_sklearn=False):
    """Script-level OOD evaluation for one image (anomaly/eval_ood_traditional.py:128-148):
    positives are the pixels whose gt is in ``out_labels``; the ranked score is ``-conf``.
    Returns None when either class is empty."""
    if mask is not None:
        seg_label = seg_label[mask]
    out = seg_label == out_labels[0]
    for lab in out_labels:
        out = np.logical_or(out, seg_label == lab)
    in_scores = -conf[np.logical_not(out)]
    out_scores = -conf[out]
    if (len(out_scores) != 0) and (len(in_scores) != 0):
        return get_and_print_results(out_scores, in_scores, use_sklearn=use_sklearn)
    return None


def eval_ood_measure_anom_utils(conf: np.ndarray, seg_label: np.ndarray, out_label=13, use_sklearn=False):
    """anomaly/anom_utils.py:106-116 variant (single out label; names swapped in the
    reference but positives are still the OOD pixels)."""
    pos = -conf[np.where(seg_label == out_label)]
    neg = -conf[np.where(seg_label != out_label)]
    if (len(neg) != 0) and (len(pos) != 0):
        return get_and_print_results(pos, neg, use_sklearn=use_sklearn)
    return None


# --------------------------------------------------------------------------- #
# segmentation counts
# --------------------------------------------------------------------------- #
def accuracy(preds: np.ndarray, label: np.ndarray):
    """anomaly/utils.py:128-133."""
    valid = (label >= 0)
    acc_sum = (valid * (preds == label)).sum()
    valid_sum = valid.sum()
    return float(acc_sum) / (valid_sum + 1e-10), valid_sum


def intersection_and_union(im_pred: np.ndarray, im_lab: np.ndarray, num_class: int):
    """anomaly/utils.py:136-156."""
    im_pred = np.asarray(im_pred).copy() + 1
    im_lab = np.asarray(im_lab).copy() + 1
    im_pred = im_pred * (im_lab > 0)
    inter = im_pred * (im_pred == im_lab)
    area_i, _ = np.histogram(inter, bins=num_class, range=(1, num_class))
    area_p, _ = np.histogram(im_pred, bins=num_class, range=(1, num_class))
    area_l, _ = np.histogram(im_lab, bins=num_class, range=(1, num_class))
    return area_i, area_p + area_l - area_i


def fast_hist(label_true: np.ndarray, label_pred: np.ndarray, n_classes: int = 19) -> np.ndarray:
    """DeepLabV3Plus-Pytorch/metrics/stream_metrics.py:49-55 (n_classes is fixed at 19, :30)."""
    mask = (label_true >= 0) & (label_true < n_classes)
    return np.bincount(n_classes * label_true[mask].astype(int) + label_pred[mask],
                       minlength=n_classes ** 2).reshape(n_classes, n_classes)


def seg_results(hist: np.ndarray) -> dict:
    """DeepLabV3Plus-Pytorch/metrics/stream_metrics.py:57-81."""
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(hist).sum() / hist.sum()
        acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
        mean_iu = np.nanmean(iu)
        freq = hist.sum(axis=1) / hist.sum()
        fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
    return {"Overall Acc": acc, "Mean Acc": acc_cls, "FreqW Acc": fwavacc, "Mean IoU": mean_iu,
            "Class IoU": dict(zip(range(hist.shape[0]), iu))}
