"""Exact, tie-aware AUROC / AUPR / FPR@recall on the GPU (kernel (d)) -- functional API.

Replaces anomaly/anom_utils.py:25-78 (``fpr_and_fdr_at_recall``, ``get_measures`` with its two
scikit-learn calls) and the per-image driver anomaly/eval_ood_traditional.py:128-148.
Everything between the score map and the three numbers runs in libdml_b200.so: key
generation, segmented radix sort, group scan, fixed-order reductions.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from ._lib import OOD_RESULT_WORDS, DmlError, check, lib, ptr, require_cuda, stream_ptr

RECALL_LEVEL_DEFAULT = 0.95       # anomaly/anom_utils.py:4
KEY_BASE_NONNEG = 0x80000000      # sortable image of +0.0: any conf >= 0 fits the 31-bit window
PARTIAL_WORDS = 10                # u64 auroc_num, f64 ap_sum, i64 a_idx,a_tps,a_fps, b_tps,b_idx,b_fps, n_groups, reserved
NO_B = (1 << 63) - 1
POS_CAPACITY_DEFAULT = 16384      # positives per segment of the minority-rank path (<= 32768: shared-memory sort)


class MinorityOverflow(DmlError):
    """A segment evaluated with ``method="rank"`` holds more positives than ``pos_capacity``: its result is NaN.
    Re-evaluate with ``method="sort"`` (no limit) or ``method="auto"`` (checks and falls back by itself)."""


def label_mask(out_labels: Iterable[int]) -> int:
    m = 0
    for l in out_labels:
        l = int(l)
        if not 0 <= l < 64:
            raise ValueError("out labels must lie in [0, 64)")
        m |= 1 << l
    return m


class OodWorkspace:
    """Grow-only device scratch (keys, sort/scan workspace, per-segment stats and results), so a
    streaming evaluation does no allocation after warm-up."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._bufs = {}

    def get(self, name: str, nbytes: int) -> torch.Tensor:
        cur = self._bufs.get(name)
        if cur is None or cur.numel() < nbytes:
            cur = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._bufs[name] = cur
        return cur


class KeyPool:
    """Pooled metric over the (score, label) pairs of many ``eval_segments`` calls -- e.g. the exact full-set
    AUROC / AUPR / FPR@95 next to the per-image mean -- without generating or counting the ranking keys twice.

    Pass the pool to every ``eval_segments(..., pool=pool)`` call of the evaluation: the call writes its packed keys
    into the pool's next slice (the per-segment sort then runs in place there), adds its digit histograms to the
    pooled histogram (``dml_ood_pool_histograms``) and its positive / NaN counts to the pooled stats.
    ``evaluate()`` finally sorts and scans the whole buffer as ONE segment.  All calls must use the same
    ``key_base`` / ``score_kind``; ``reset()`` starts the next evaluation (buffers are kept)."""

    def __init__(self, capacity: int, device, workspace: Optional["OodWorkspace"] = None, histograms: bool = True):
        self.device = torch.device(device)
        self.capacity = int(capacity)
        if not 0 < self.capacity < (1 << 32):
            raise ValueError("KeyPool capacity must lie in (0, 2^32)")
        self.ws = workspace or OodWorkspace(self.device)
        self.keys = torch.empty(self.capacity, dtype=torch.int32, device=self.device)   # u32 bit patterns
        # histograms=False: only the keys and counts are pooled (multi-GPU: the keys are re-partitioned across ranks
        # by distributed.pooled_measures(keys_and_stats=...), so a local pooled histogram would be of no use)
        self.scratch = self.ws.get("pool_scratch", lib().dml_ood_workspace_bytes(1, self.capacity)) if histograms else None
        self.stats = torch.zeros(1, 4, dtype=torch.int64, device=self.device)
        # score keys of the positives the minority-rank batches gathered anyway (method="rank"): the pooled rank
        # evaluation then skips its compaction pass over all keys.  Allocated by the first such batch.
        self.pos = None
        self.pos_count = torch.zeros(1, dtype=torch.int64, device=self.device)
        # multi-GPU, method="rank": a distributed.PositiveExchange that ships every batch's positives to all ranks at once
        # (set by the caller; every eval_segments call on this pool then publishes exactly one slot)
        self.exchange = None
        self.reset()

    def reset(self):
        self.n = 0
        self.signature = None
        self.hist_ok = True      # every contribution so far also left its digit histograms (method="sort" batches)
        self.stats.zero_()
        self.pos_count.zero_()
        if getattr(self, "exchange", None) is not None:
            self.exchange.begin()

    def _take(self, n: int, signature) -> torch.Tensor:
        if self.n + n > self.capacity:
            raise ValueError(f"KeyPool overflow: {self.n} + {n} > capacity {self.capacity}")
        if self.signature is None:
            self.signature = signature
        elif self.signature != signature:
            raise ValueError("all eval_segments calls of one KeyPool must share key_base and score_kind")
        view = self.keys[self.n:self.n + n]
        self.n += n
        return view

    def evaluate(self, recall_level: float = RECALL_LEVEL_DEFAULT, method: str = "sort"):
        """(results [1,7] float64, stats [1,4] int64) device tensors of the pooled segment, like ``eval_segments``.
        method="sort": radix sort of all pooled keys + scan; the pool must be full (``capacity`` keys): the pooled
        histogram slot belongs to that segment length.  method="rank": ``rank_keys`` -- only the positives are sorted,
        the negatives are grouped by bucket and located among them (two small host synchronisations)."""
        if method not in ("sort", "rank"):
            raise ValueError("method must be 'sort' or 'rank'")
        if method == "rank":
            key_base = self.signature[0] if self.signature is not None else KEY_BASE_NONNEG
            return rank_keys(self.keys[: self.n], self.stats, key_base, recall_level, self.ws, pos_keys=self.pos,
                             pos_count=self.pos_count), self.stats
        if self.scratch is None:
            raise ValueError("KeyPool(histograms=False) only collects keys; evaluate them with distributed.pooled_measures")
        if self.n != self.capacity:
            raise ValueError(f"KeyPool holds {self.n} keys, capacity is {self.capacity}: size the pool to the evaluation")
        results = self.ws.get("pool_results", 8 * OOD_RESULT_WORDS).view(torch.float64)[:OOD_RESULT_WORDS].view(1, OOD_RESULT_WORDS)
        with torch.cuda.device(self.device):
            check(lib().dml_ood_eval_segments(ptr(self.keys), ptr(self.stats), 1, self.capacity, recall_level,
                                              ptr(self.scratch), self.scratch.numel(), 1 if self.hist_ok else 0, ptr(results),
                                              stream_ptr(self.device)), "dml_ood_eval_segments")
        return results, self.stats


def _raise_on_bad_stats(stats: np.ndarray, what: str):
    if (stats[:, 1] > 0).any():
        raise ValueError("Input contains NaN.")           # sklearn's validation in the reference path
    if (stats[:, 2] > 0).any():
        raise ValueError(f"{what}: ranking keys outside the packed 31-bit window (wrong key_base)")
    if stats.shape[1] > 3 and (stats[:, 3] > 0).any():
        raise MinorityOverflow(f"{what}: {int((stats[:, 3] > 0).sum())} segment(s) hold more positives than pos_capacity "
                               "(method='rank'); use method='sort' or 'auto'")


def eval_segments(values: torch.Tensor, n_seg: int, seg_len: int, *, gt: Optional[torch.Tensor] = None,
                  out_labels: Sequence[int] = (13,), positive: Optional[torch.Tensor] = None, score_kind: int = 0,
                  key_base: int = KEY_BASE_NONNEG, minmax: Optional[torch.Tensor] = None, minmax_slot: int = 0,
                  conf_out: Optional[torch.Tensor] = None, recall_level: float = RECALL_LEVEL_DEFAULT,
                  workspace: Optional[OodWorkspace] = None, msp: Optional[torch.Tensor] = None,
                  msp_norm_out: Optional[torch.Tensor] = None, mix_out: Optional[torch.Tensor] = None,
                  lam: float = 50.0, thr: float = 0.2, pool: Optional[KeyPool] = None, method: str = "sort",
                  pos_capacity: int = POS_CAPACITY_DEFAULT):
    """Evaluate ``n_seg`` independent segments of ``seg_len`` (score, label) pairs each.

    values: flat fp32 CUDA tensor (n_seg*seg_len): a ``conf`` map ranked as score = -conf
            (``score_kind=0``; positives expected at low conf) or plain scores (``score_kind=1``).
    gt/out_labels: positives are pixels whose gt label is in ``out_labels``; or pass ``positive``
            (uint8, non-zero = positive) directly.
    minmax: optional [n_seg,4] per-segment (min,max) pairs; the kernel then ranks the min-max
            normalised value (slot 0: eds, 1: msp) and can store it to ``conf_out``.
    msp / msp_norm_out / mix_out: fused score maps (slot 0 only): MMSP = normalised ``msp`` and the
            EDS/MMSP mix (anomaly/eval_ood_traditional.py:434-435,447-448) written in the same pass.
    pool:   a ``KeyPool`` collecting the keys / digit histograms / counts of this call for a later pooled
            evaluation over all calls (``pool.evaluate()``).
    method: "sort" = radix sort of every pair + tie-aware scan (no limits); "rank" = the minority-rank path
            (``dml_ood_rank_segments``: only the positives are sorted, every negative is located among them in the
            one pass that also writes the maps; 3-4x less work when positives are rare) -- segments with more than
            ``pos_capacity`` positives come back NaN with ``stats[:, 3] = 1`` and ``results_to_host`` raises
            ``MinorityOverflow``; "auto" = "rank", then ONE host synchronisation to look at the overflow flags and a
            re-evaluation with "sort" if any is set (the inputs must still be intact, so not for in-place pipelines).
            AUROC / FPR are bit-identical between the methods, AUPR differs by float64 summation order;
            ``results[:, 6]`` (n_groups) is -1 on the rank path.
    Returns (results, stats): device tensors -- results float64 [n_seg,7] viewed as
    (auroc, aupr, fpr, n_pos, n_neg, n_nan, n_groups; the last four are int64 bit patterns),
    stats int64 [n_seg,4] = (n_pos, n_nan, n_out_of_window, 0).  No host synchronisation.
    """
    require_cuda(values, "values")
    if method not in ("sort", "rank", "auto"):
        raise ValueError("method must be 'sort', 'rank' or 'auto'")
    dev = values.device
    values = values.contiguous().view(-1)
    if values.dtype != torch.float32 or values.numel() != n_seg * seg_len:
        raise ValueError("values must be float32 with n_seg*seg_len elements")
    ws = workspace or OodWorkspace(dev)
    n = n_seg * seg_len
    if method != "sort" and n > 0:
        return _eval_segments_rank(values, n_seg, seg_len, gt, out_labels, positive, score_kind, key_base, minmax, minmax_slot,
                                   conf_out, recall_level, ws, msp, msp_norm_out, mix_out, lam, thr, pool, method, pos_capacity)
    keys = ws.get("keys", 4 * n) if pool is None else pool._take(n, (int(key_base), int(score_kind)))
    stats = ws.get("stats", 32 * max(n_seg, 1)).view(torch.int64)[: 4 * n_seg].view(n_seg, 4)
    results = ws.get("results", 8 * OOD_RESULT_WORDS * max(n_seg, 1)).view(torch.float64)[: OOD_RESULT_WORDS * n_seg]
    results = results.view(n_seg, OOD_RESULT_WORDS)
    nbytes = lib().dml_ood_workspace_bytes(n_seg, seg_len)
    scratch = ws.get("scratch", nbytes)
    gt_u8 = gt_i64 = pos_u8 = None
    if positive is not None:
        pos_u8 = positive.contiguous().view(-1)
        if pos_u8.dtype == torch.bool:
            pos_u8 = pos_u8.view(torch.uint8)
        if pos_u8.dtype != torch.uint8 or pos_u8.numel() != n:
            raise ValueError("positive must be uint8/bool with one entry per value")
    elif gt is not None:
        g = gt.contiguous().view(-1)
        if g.numel() != n:
            raise ValueError("gt must have one entry per value")
        if g.dtype == torch.uint8:
            gt_u8 = g
        elif g.dtype == torch.int64:
            gt_i64 = g
        else:
            raise ValueError("gt must be uint8 or int64")
    else:
        raise ValueError("need gt or positive")
    # batched segments (and every pooled evaluation): the sort's digit histograms are accumulated by the key-generation
    # kernel while the keys are in registers (B200: 264 us per 46 M keys against 231 us key-gen + 227 us separate counting
    # read); one long segment keeps the separate hist_kernel (few, long look-back chains: the fused form measured
    # slower there).  DML_FUSED_HIST=0 / 1 forces the separate / fused form.
    fh = os.environ.get("DML_FUSED_HIST", "")
    fused_hist = n > 0 and (fh == "1" or (fh != "0" and n_seg >= 4) or pool is not None)
    with torch.cuda.device(dev):
        s = stream_ptr(dev)
        check(lib().dml_ood_keygen(ptr(values), ptr(minmax), minmax_slot, ptr(conf_out), ptr(gt_u8), ptr(gt_i64),
                                   label_mask(out_labels) if positive is None else 0, ptr(pos_u8), score_kind, key_base,
                                   n_seg, seg_len, ptr(keys), ptr(stats), ptr(msp), ptr(msp_norm_out), ptr(mix_out),
                                   lam, thr, ptr(scratch) if fused_hist else None, scratch.numel() if fused_hist else 0, s),
              "dml_ood_keygen")
        if pool is not None and n > 0:
            # between key-gen and the sort the workspace holds the raw per-segment digit counts
            check(lib().dml_ood_pool_histograms(ptr(scratch), scratch.numel(), ptr(stats), n_seg, seg_len, ptr(pool.scratch),
                                                pool.scratch.numel() if pool.scratch is not None else 0, pool.capacity,
                                                ptr(pool.stats), 1 if pool.n == n else 0, s), "dml_ood_pool_histograms")
            if pool.exchange is not None:        # a sort-path batch keeps its positives: an empty slot keeps the ranks in step
                pool.exchange.publish_empty(pool.stats, pool.n)
        check(lib().dml_ood_eval_segments(ptr(keys), ptr(stats), n_seg, seg_len, recall_level, ptr(scratch),
                                          scratch.numel(), 1 if fused_hist else 0, ptr(results), s), "dml_ood_eval_segments")
    return results, stats


def _labels_of(gt, positive, n):
    gt_u8 = gt_i64 = pos_u8 = None
    if positive is not None:
        pos_u8 = positive.contiguous().view(-1)
        if pos_u8.dtype == torch.bool:
            pos_u8 = pos_u8.view(torch.uint8)
        if pos_u8.dtype != torch.uint8 or pos_u8.numel() != n:
            raise ValueError("positive must be uint8/bool with one entry per value")
    elif gt is not None:
        g = gt.contiguous().view(-1)
        if g.numel() != n:
            raise ValueError("gt must have one entry per value")
        if g.dtype == torch.uint8:
            gt_u8 = g
        elif g.dtype == torch.int64:
            gt_i64 = g
        else:
            raise ValueError("gt must be uint8 or int64")
    else:
        raise ValueError("need gt or positive")
    return gt_u8, gt_i64, pos_u8


def _eval_segments_rank(values, n_seg, seg_len, gt, out_labels, positive, score_kind, key_base, minmax, minmax_slot, conf_out,
                        recall_level, ws, msp, msp_norm_out, mix_out, lam, thr, pool, method, pos_capacity):
    """``eval_segments`` through ``dml_ood_rank_segments`` (see there)."""
    dev = values.device
    n = n_seg * seg_len
    pos_capacity = int(min(max(pos_capacity, 1), 32768))
    gt_u8, gt_i64, pos_u8 = _labels_of(gt, positive, n)
    stats = ws.get("stats", 32 * max(n_seg, 1)).view(torch.int64)[: 4 * n_seg].view(n_seg, 4)
    results = ws.get("results", 8 * OOD_RESULT_WORDS * max(n_seg, 1)).view(torch.float64)[: OOD_RESULT_WORDS * n_seg]
    results = results.view(n_seg, OOD_RESULT_WORDS)
    rws = ws.get("rank_ws", lib().dml_ood_rank_workspace_bytes(n_seg, pos_capacity))
    pool_mark = (pool.n, pool.signature, pool.hist_ok) if pool is not None else None
    keys = pool._take(n, (int(key_base), int(score_kind))) if pool is not None else None
    with torch.cuda.device(dev):
        s = stream_ptr(dev)
        check(lib().dml_ood_rank_segments(ptr(values), ptr(minmax), minmax_slot, ptr(conf_out), ptr(gt_u8), ptr(gt_i64),
                                          label_mask(out_labels) if positive is None else 0, ptr(pos_u8), score_kind,
                                          key_base, n_seg, seg_len, ptr(keys), ptr(stats), ptr(msp), ptr(msp_norm_out),
                                          ptr(mix_out), lam, thr, pos_capacity, recall_level, ptr(rws), rws.numel(),
                                          ptr(results), s), "dml_ood_rank_segments")
        if method == "auto" and bool((stats[:, 3] != 0).any().item()):       # the one host synchronisation of "auto"
            if pool is not None:
                pool.n, pool.signature, pool.hist_ok = pool_mark
            return eval_segments(values, n_seg, seg_len, gt=gt, out_labels=out_labels, positive=positive,
                                 score_kind=score_kind, key_base=key_base, minmax=minmax, minmax_slot=minmax_slot,
                                 conf_out=conf_out, recall_level=recall_level, workspace=ws, msp=msp,
                                 msp_norm_out=msp_norm_out, mix_out=mix_out, lam=lam, thr=thr, pool=pool, method="sort")
        if pool is not None:
            pool.hist_ok = False       # this batch left keys and counts, no digit histograms
            check(lib().dml_ood_pool_histograms(None, 0, ptr(stats), n_seg, seg_len, None, 0, pool.capacity, ptr(pool.stats),
                                                1 if pool.n == n else 0, s), "dml_ood_pool_histograms")
            if pool.exchange is not None:
                # multi-GPU: the batch's positives go into the exchange's next slot and travel to the other ranks while
                # the following batches are evaluated (distributed.PositiveExchange)
                slot_keys, slot_count = pool.exchange.next_slot(n_seg * pos_capacity)
                check(lib().dml_ood_rank_export_positives(ptr(rws), rws.numel(), n_seg, pos_capacity, ptr(slot_keys),
                                                          slot_keys.numel(), ptr(slot_count), s), "dml_ood_rank_export_positives")
                pool.exchange.publish(pool.stats, pool.n)
            else:
                if pool.pos is None:
                    pool.pos = torch.empty(max(1 << 20, pool.capacity // 16), dtype=torch.int32, device=dev)
                check(lib().dml_ood_rank_export_positives(ptr(rws), rws.numel(), n_seg, pos_capacity, ptr(pool.pos), pool.pos.numel(),
                                                          ptr(pool.pos_count), s), "dml_ood_rank_export_positives")
    return results, stats


MAX_RANK_GROUPS = 4096 * 12288    # distinct positive scores dml_ood_bucket_rank can bucket


def _i32(ws: "OodWorkspace", name: str, n: int) -> torch.Tensor:
    return ws.get(name, 4 * max(int(n), 1)).view(torch.int32)[: max(int(n), 0)]


def sorted_positive_keys(keys: torch.Tensor, n_pos: int, ws: "OodWorkspace", tag: str = "pr") -> torch.Tensor:
    """Score keys (key >> 1) of the ``n_pos`` positives of ``keys`` (packed, bit 0 = positive), sorted ascending.
    int32 tensor of u32 bit patterns (may alias workspace memory)."""
    dev = keys.device
    out = _i32(ws, tag + "_pos", n_pos)
    cnt = ws.get(tag + "_count", 8).view(torch.int64)[:1]
    with torch.cuda.device(dev):
        s = stream_ptr(dev)
        check(lib().dml_ood_pos_compact(ptr(keys), keys.numel(), ptr(out), n_pos, ptr(cnt), s), "dml_ood_pos_compact")
        return sort_keys(out, ws, tag + "_sort", end_bit=31)


def sort_keys(keys: torch.Tensor, ws: "OodWorkspace", tag: str, end_bit: int = 32) -> torch.Tensor:
    """radix sort of u32 bit patterns (int32 tensor); returns the sorted tensor (``keys`` itself or workspace memory)"""
    n = keys.numel()
    if n == 0:
        return keys
    dev = keys.device
    scratch = ws.get(tag, lib().dml_ood_workspace_bytes(1, n))
    out = C.c_void_p()
    with torch.cuda.device(dev):
        check(lib().dml_ood_sort(ptr(keys), 1, n, 0, end_bit, ptr(scratch), scratch.numel(), C.byref(out), stream_ptr(dev)),
              "dml_ood_sort")
    if out.value == keys.data_ptr():
        return keys
    off = out.value - scratch.data_ptr()
    return scratch[off: off + 4 * n].view(torch.int32)


def unique_groups(sorted_pos: torch.Tensor, ws: "OodWorkspace", tag: str = "pr"):
    """sorted score keys -> (S [G] distinct scores, pc [G] multiplicities, G).  One host synchronisation (G)."""
    dev = sorted_pos.device
    n = sorted_pos.numel()
    S, pc = _i32(ws, tag + "_S", n), _i32(ws, tag + "_pc", n)
    g_dev = ws.get(tag + "_G", 8).view(torch.int64)[:1]
    scratch = ws.get(tag + "_uniq", lib().dml_ood_unique_workspace_bytes(n))
    with torch.cuda.device(dev):
        check(lib().dml_ood_unique_counts(ptr(sorted_pos), n, ptr(S), ptr(pc), ptr(g_dev), ptr(scratch), scratch.numel(),
                                          stream_ptr(dev)), "dml_ood_unique_counts")
    G = int(g_dev.item())
    return S[:G], pc[:G], G


def bucket_rank_counters(keys: torch.Tensor, S: torch.Tensor, key_base: int, ws: "OodWorkspace", tag: str = "pr") -> torch.Tensor:
    """int64 [2G + 2] counters of the negatives of ``keys`` against the sorted distinct positive scores ``S``
    (``dml_ood_bucket_rank``); counters of disjoint key sets add (multi-GPU: one all_reduce)."""
    dev = keys.device
    n, G = keys.numel(), S.numel()
    cnt = ws.get(tag + "_cnt", 8 * (2 * G + 2)).view(torch.int64)[: 2 * G + 2]
    scratch = ws.get(tag + "_bucket", lib().dml_ood_bucket_rank_workspace_bytes(n, G))
    with torch.cuda.device(dev):
        check(lib().dml_ood_bucket_rank(ptr(keys), n, ptr(S), G, key_base, ptr(cnt), ptr(scratch), scratch.numel(),
                                        stream_ptr(dev)), "dml_ood_bucket_rank")
    return cnt


def pooled_scan(pc: torch.Tensor, cnt: torch.Tensor, total_pos: int, total_n: int, n_nan: int, recall_level: float,
                ws: "OodWorkspace", tag: str = "pr") -> torch.Tensor:
    """[1,7] float64 result row (auroc, aupr, fpr, n_pos, n_neg, n_nan, -1) from the group counts / counters"""
    dev = ws.device
    G = pc.numel()
    res = ws.get(tag + "_result", 8 * OOD_RESULT_WORDS).view(torch.float64)[:OOD_RESULT_WORDS].view(1, OOD_RESULT_WORDS)
    scratch = ws.get(tag + "_scan", lib().dml_ood_pooled_scan_workspace_bytes(G))
    with torch.cuda.device(dev):
        check(lib().dml_ood_pooled_scan(ptr(pc) if G else None, ptr(cnt) if G else None, G, total_pos, total_n, n_nan,
                                        recall_level, ptr(scratch), scratch.numel(), ptr(res), stream_ptr(dev)),
              "dml_ood_pooled_scan")
    return res


def rank_keys(keys: torch.Tensor, stats: torch.Tensor, key_base: int = KEY_BASE_NONNEG,
              recall_level: float = RECALL_LEVEL_DEFAULT, workspace: Optional["OodWorkspace"] = None,
              pos_keys: Optional[torch.Tensor] = None, pos_count: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Exact metrics of ONE ranking given as packed keys (int32 tensor of u32 bit patterns, any order; e.g. a
    ``KeyPool``) without sorting its negatives: positives compacted + sorted + de-duplicated, negatives bucketed and
    located among them, scan over the positive groups (csrc/ood_pool_rank.cu).  ``stats``: int64 [.., 4] device tensor
    whose first row starts with (n_pos, n_nan, ...).  Returns the [1,7] float64 device row of ``eval_segments``
    (n_groups = -1).  Two host synchronisations (n_pos, number of distinct positive scores); falls back to the sort
    path when the positives have more than 4096 * 12288 distinct scores.  ``pos_keys`` / ``pos_count``: score keys of
    the positives collected beforehand (``KeyPool.pos``); used when they are complete, else the keys are compacted."""
    require_cuda(keys, "keys")
    keys = keys.contiguous().view(-1)
    dev = keys.device
    ws = workspace or OodWorkspace(dev)
    n = keys.numel()
    head = stats.view(-1)[:2] if pos_count is None else torch.cat([stats.view(-1)[:2], pos_count.view(-1)[:1]])
    hv = [int(v) for v in head.tolist()]
    n_pos, n_nan = hv[0], hv[1]
    if n_pos <= 0 or n_pos >= n:
        return pooled_scan(keys[:0], keys[:0], n_pos, n, n_nan, recall_level, ws)
    if pos_keys is not None and pos_count is not None and hv[2] == n_pos and n_pos <= pos_keys.numel():
        srt = sort_keys(pos_keys[:n_pos], ws, "pr_sort", end_bit=31)
    else:
        srt = sorted_positive_keys(keys, n_pos, ws)
    S, pc, G = unique_groups(srt, ws)
    if G > MAX_RANK_GROUPS:
        res = ws.get("pr_result", 8 * OOD_RESULT_WORDS).view(torch.float64)[:OOD_RESULT_WORDS].view(1, OOD_RESULT_WORDS)
        scratch = ws.get("pr_fallback", lib().dml_ood_workspace_bytes(1, n))
        st = stats.view(-1)[:4].view(1, 4).contiguous()
        work = keys.clone()
        with torch.cuda.device(dev):
            check(lib().dml_ood_eval_segments(ptr(work), ptr(st), 1, n, recall_level, ptr(scratch), scratch.numel(), 0,
                                              ptr(res), stream_ptr(dev)), "dml_ood_eval_segments")
        return res
    cnt = bucket_rank_counters(keys, S, key_base, ws)
    return pooled_scan(pc, cnt, n_pos, n, n_nan, recall_level, ws)


def roc_fpr_after_eval(workspace: "OodWorkspace", n_seg: int, seg_len: int,
                       recall_level: float = RECALL_LEVEL_DEFAULT) -> torch.Tensor:
    """Second FPR@recall convention for the segments just evaluated by ``eval_segments(..., workspace=workspace)``:
    ``fpr[tpr >= recall_level][0]`` on ``sklearn.metrics.roc_curve`` with ``drop_intermediate=True``
    (DeepLabV3Plus-Pytorch/test.py:241-244).  Returns a float64 device tensor [n_seg] (NaN: single class)."""
    dev = workspace.device
    keys = workspace.get("keys", 4 * n_seg * seg_len)
    stats = workspace.get("stats", 32 * max(n_seg, 1)).view(torch.int64)[: 4 * n_seg].view(n_seg, 4)
    scratch = workspace.get("scratch", lib().dml_ood_workspace_bytes(n_seg, seg_len))
    out = torch.empty(n_seg, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib().dml_ood_roc_fpr(ptr(keys), ptr(stats), n_seg, seg_len, recall_level, ptr(scratch), scratch.numel(),
                                    ptr(out), stream_ptr(dev)), "dml_ood_roc_fpr")
    return out


def results_to_host(results: torch.Tensor, stats: torch.Tensor, what: str = "ood metrics"):
    """D2H copy + the reference's error behaviour: NaN scores -> ValueError (sklearn);
    single-class segments come back as NaN rows (callers map them to ``None``)."""
    r = results.cpu().numpy()
    st = stats.cpu().numpy()
    _raise_on_bad_stats(st, what)
    return r[:, :3].copy(), r[:, 3:].view(np.int64).copy()


# --------------------------------------------------------------------------------------------
# arbitrary scores (any sign / range): choose the key window from the data
# --------------------------------------------------------------------------------------------
def _scan_sorted_range(sorted_keys_ptr: int, n: int, pos_before: int, idx_before: int, total_pos: int, total_n: int,
                       recall_level: float, ws: OodWorkspace, dev) -> np.ndarray:
    info = torch.tensor([pos_before, idx_before, total_pos, total_n], dtype=torch.int64, device=dev)
    partial = torch.empty(PARTIAL_WORDS, dtype=torch.int64, device=dev)
    nbytes = lib().dml_ood_workspace_bytes(1, max(n, 1))
    scratch = ws.get("scan_scratch", nbytes)
    check(lib().dml_ood_scan_range(C.c_void_p(sorted_keys_ptr), n, ptr(info), recall_level, ptr(scratch), scratch.numel(),
                                   ptr(partial), stream_ptr(dev)), "dml_ood_scan_range")
    return partial.cpu().numpy()


def combine_partials(partials: Sequence[np.ndarray], total_pos: int, total_n: int,
                     recall_level: float = RECALL_LEVEL_DEFAULT):
    """Combine per-range partial tuples (see ``dml_ood_scan_range``) into (auroc, aupr, fpr, n_groups).
    Sums are exact integers / fixed-order float64 adds and the FPR choice compares two float64
    recalls, so every rank computes identical bits."""
    num = 0
    ap = 0.0
    a = (-1, 0, 0)          # idx, tps, fps
    b = (NO_B, -1, 0)       # tps, idx, fps
    groups = 0
    for p in partials:
        p = np.asarray(p, dtype=np.int64)
        num += int(p.view(np.uint64)[0])
        ap += float(p.view(np.float64)[1])
        if int(p[2]) > a[0]:
            a = (int(p[2]), int(p[3]), int(p[4]))
        if int(p[5]) < b[0] or (int(p[5]) == b[0] and int(p[6]) > b[1]):
            b = (int(p[5]), int(p[6]), int(p[7]))
        groups += int(p[8])
    n_neg = total_n - total_pos
    if total_pos == 0 or n_neg == 0:
        return float("nan"), float("nan"), float("nan"), groups
    P = float(total_pos)
    da = abs(a[1] / P - recall_level) if a[0] >= 0 else float("inf")
    db = abs(b[0] / P - recall_level) if b[0] != NO_B else float("inf")
    fps = b[2] if db <= da else a[2]
    return num / (2.0 * total_pos * n_neg), ap / total_pos, fps / n_neg, groups


def measures_from_scores(scores: torch.Tensor, positive: torch.Tensor, recall_level: float = RECALL_LEVEL_DEFAULT,
                         workspace: Optional[OodWorkspace] = None, fpr_convention: str = "closest"):
    """(auroc, aupr, fpr) for arbitrary fp32 scores (higher = more positive) and a 0/1 mask.
    ``fpr_convention``: "closest" = anomaly/anom_utils.py:57-65 (recall closest to the level);
    "roc_curve" = ``fpr[tpr >= level][0]`` on sklearn's ``roc_curve`` (DeepLabV3Plus-Pytorch/test.py:242-244).
    One segment; the packed-key window is chosen from the data (one extra read of the scores);
    mixed-sign inputs whose keys span more than 31 bits are ranked as two ranges (score > 0,
    then score <= 0) that are sorted separately and scanned with carried counts."""
    if fpr_convention not in ("closest", "roc_curve"):
        raise ValueError("fpr_convention must be 'closest' or 'roc_curve'")
    require_cuda(scores, "scores")
    dev = scores.device
    scores = scores.contiguous().view(-1).float()
    positive = positive.contiguous().view(-1)
    if positive.dtype == torch.bool:
        positive = positive.view(torch.uint8)
    n = scores.numel()
    ws = workspace or OodWorkspace(dev)
    with torch.cuda.device(dev):
        st = torch.empty(4, dtype=torch.int64, device=dev)
        check(lib().dml_ood_keystats(ptr(scores), 1, 1, n, ptr(st), stream_ptr(dev)), "dml_ood_keystats")
        kmin, kmax, n_nan, _ = st.cpu().tolist()
        if n_nan:
            raise ValueError("Input contains NaN.")
        if kmax - kmin < (1 << 31):
            res, stats = eval_segments(scores, 1, n, positive=positive, score_kind=1, key_base=kmin,
                                       recall_level=recall_level, workspace=ws)
            roc = roc_fpr_after_eval(ws, 1, n, recall_level) if fpr_convention == "roc_curve" else None
            vals, counts = results_to_host(res, stats)
            fpr = float(vals[0, 2]) if roc is None else float(roc.cpu()[0])
            return float(vals[0, 0]), float(vals[0, 1]), fpr
        if fpr_convention == "roc_curve":
            raise NotImplementedError("roc_curve FPR convention: scores must span less than 2^31 sortable key values "
                                      "(always true for same-sign scores such as 1 - max softmax)")
        # wide, mixed-sign range: keys of positive scores (negative ranking key) come first
        first = scores > 0
        parts = [(scores[first], positive[first]), (scores[~first], positive[~first])]
        total_pos = int(positive.sum().item())
        partials = []
        pos_before = idx_before = 0
        for sc, po in parts:
            m = sc.numel()
            if m == 0:
                continue
            check(lib().dml_ood_keystats(ptr(sc), 1, 1, m, ptr(st), stream_ptr(dev)), "dml_ood_keystats")
            base = int(st[0].item())
            keys = ws.get("keys", 4 * m)
            stats = ws.get("stats", 32).view(torch.int64)[:4].view(1, 4)
            check(lib().dml_ood_keygen(ptr(sc), None, 0, None, None, None, 0, ptr(po), 1, base, 1, m, ptr(keys),
                                       ptr(stats), None, None, None, 0.0, 0.0, None, 0, stream_ptr(dev)), "dml_ood_keygen")
            nbytes = lib().dml_ood_workspace_bytes(1, m)
            scratch = ws.get("scratch", nbytes)
            sorted_ptr = C.c_void_p()
            check(lib().dml_ood_sort(ptr(keys), 1, m, 0, 32, ptr(scratch), scratch.numel(), C.byref(sorted_ptr),
                                     stream_ptr(dev)), "dml_ood_sort")
            _raise_on_bad_stats(stats.cpu().numpy(), "measures_from_scores")
            partials.append(_scan_sorted_range(sorted_ptr.value, m, pos_before, idx_before, total_pos, n,
                                               recall_level, ws, dev))
            pos_before += int(po.sum().item())
            idx_before += m
        a, p, f, _ = combine_partials(partials, total_pos, n, recall_level)
        return a, p, f
