"""GPU drop-in for the reference's ``anomaly/anom_utils.py`` (same names, argument meaning,
return values and error behaviour).  Inputs may be NumPy arrays / lists (copied to the
current CUDA device) or CUDA tensors (used in place); the ranking itself -- sort, tie-aware
cumulative counts, AUROC / AUPR / FPR reductions -- runs in libdml_b200.so.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ood

recall_level_default = 0.95  # anomaly/anom_utils.py:4

_workspaces = {}


def _ws(dev):
    key = (dev.type, dev.index)
    if key not in _workspaces:
        _workspaces[key] = ood.OodWorkspace(dev)
    return _workspaces[key]


def _to_cuda_f32(a) -> torch.Tensor:
    """Scores are RANKED IN FLOAT32 (the reference's conf maps are float32; sklearn ranks whatever dtype it is given).
    A float64 input whose distinct values collapse in float32 can gain ties the reference does not see: a lossy cast
    is reported with a warning instead of silently changing the ranking."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    if t.dtype == torch.float64 and t.numel() and not torch.equal(t.to(torch.float32).to(torch.float64), t):
        import warnings
        warnings.warn("float64 scores are ranked as float32 keys on the GPU: values that differ only beyond float32 precision "
                      "become ties (the reference ranks the float64 values)", RuntimeWarning, stacklevel=3)
    if not t.is_cuda:
        if not torch.cuda.is_available():
            ood.require_cuda(t, "scores")  # raises: there is no CPU fallback
        t = t.cuda(non_blocking=True)
    return t.reshape(-1).to(torch.float32)


def fpr_and_fdr_at_recall(y_true, y_score, recall_level=recall_level_default, pos_label=None):
    """anomaly/anom_utils.py:25-65: FPR at the threshold whose recall is closest to ``recall_level``."""
    yt = y_true if isinstance(y_true, torch.Tensor) else torch.from_numpy(np.asarray(y_true))
    if yt.is_cuda:
        classes = torch.unique(yt).cpu().numpy()
    else:
        classes = np.unique(yt.numpy())
    ok = any(np.array_equal(classes, c) for c in ([0, 1], [-1, 1], [0], [-1], [1]))
    if pos_label is None and not ok:
        raise ValueError("Data is not binary and pos_label is not specified")
    elif pos_label is None:
        pos_label = 1.
    score = _to_cuda_f32(y_score)
    positive = (yt.to(score.device).reshape(-1) == pos_label)
    _, _, fpr = ood.measures_from_scores(score, positive, recall_level, _ws(score.device))
    return fpr


def get_measures(_pos, _neg, recall_level=recall_level_default):
    """anomaly/anom_utils.py:67-78: (auroc, aupr, fpr) with ``_pos`` the positive-class scores."""
    pos = _to_cuda_f32(_pos[:])
    neg = _to_cuda_f32(_neg[:]).to(pos.device)
    if pos.numel() == 0 or neg.numel() == 0:
        # sklearn: "Only one class present in y_true. ROC AUC score is not defined in that case."
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    examples = torch.cat((pos, neg))
    positive = torch.zeros(examples.numel(), dtype=torch.uint8, device=pos.device)
    positive[: pos.numel()] = 1
    return ood.measures_from_scores(examples, positive, recall_level, _ws(pos.device))


def print_measures(auroc, aupr, fpr, method_name='Ours', recall_level=recall_level_default):
    print('\t\t\t\t' + method_name)
    print('FPR{:d}:\t\t\t{:.2f}'.format(int(100 * recall_level), 100 * fpr))
    print('AUROC: \t\t\t{:.2f}'.format(100 * auroc))
    print('AUPR:  \t\t\t{:.2f}'.format(100 * aupr))


def print_measures_with_std(aurocs, auprs, fprs, method_name='Ours', recall_level=recall_level_default):
    print('\t\t\t\t' + method_name)
    print('FPR{:d}:\t\t\t{:.2f}\t+/- {:.2f}'.format(int(100 * recall_level), 100 * np.mean(fprs), 100 * np.std(fprs)))
    print('AUROC: \t\t\t{:.2f}\t+/- {:.2f}'.format(100 * np.mean(aurocs), 100 * np.std(aurocs)))
    print('AUPR:  \t\t\t{:.2f}\t+/- {:.2f}'.format(100 * np.mean(auprs), 100 * np.std(auprs)))


def get_and_print_results(out_score, in_score, num_to_avg=1):
    """anomaly/anom_utils.py:95-104."""
    auroc, aupr, fpr = get_measures(out_score, in_score)
    return float(np.mean([auroc])), float(np.mean([aupr])), float(np.mean([fpr]))


def _as_cuda(a, dtype=None):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    if not t.is_cuda:
        t = t.cuda(non_blocking=True)
    return t if dtype is None else t.to(dtype)


def eval_ood_measure(conf, seg_label, out_label=13):
    """anomaly/anom_utils.py:106-116: positives are the pixels labelled ``out_label``; the ranked
    score is ``-conf``.  Returns ``None`` when either class is empty.  One fused key-gen + sort +
    scan over the whole map (no boolean-mask gathers)."""
    return eval_conf_map(conf, seg_label, (out_label,))


def eval_conf_map(conf, seg_label, out_labels, recall_level=recall_level_default):
    conf_t = _as_cuda(conf, torch.float32).reshape(-1)
    lab = _as_cuda(seg_label).reshape(-1).to(conf_t.device)
    if lab.numel() != conf_t.numel():
        raise IndexError("conf and seg_label differ in size")
    ws = _ws(conf_t.device)
    if lab.dtype in (torch.uint8, torch.int64) and all(0 <= int(l) < 64 for l in out_labels):
        # fast path (non-negative conf, e.g. any min-max normalised map): labels and conf go straight
        # into key generation, no mask / negation passes
        res, stats = ood.eval_segments(conf_t, 1, conf_t.numel(), gt=lab, out_labels=out_labels, score_kind=0,
                                       recall_level=recall_level, workspace=ws)
        st = stats.cpu().numpy()
        if st[0, 1] > 0:
            raise ValueError("Input contains NaN.")
        if st[0, 2] == 0:
            r = res.cpu().numpy()[0]
            return None if np.isnan(r[0]) else (float(r[0]), float(r[1]), float(r[2]))
    positive = torch.zeros(lab.numel(), dtype=torch.bool, device=lab.device)
    for l in out_labels:
        positive |= (lab == l)
    # conf is ranked as score = -conf
    a, p, f = ood.measures_from_scores(-conf_t, positive, recall_level, ws)
    if np.isnan(a):
        return None
    return a, p, f
