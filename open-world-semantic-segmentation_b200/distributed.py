"""Pooled (full-set) exact AUROC / AUPR / FPR@95 across ranks -- SURVEY.md section 8(e).

One process per GPU (``torch.distributed``, NCCL over NVLink).  Images shard per rank with no
data-path collective; only the pooled ranking needs an exchange:

  1. every rank turns its (conf, gt) pairs into packed keys and radix-sorts them locally;
  2. G-1 splitters are chosen from an all-gathered regular sample of the sorted shards
     (splitters have the positive bit cleared, so one score value never straddles two ranges);
  3. exchange of the locally SORTED shards:
       * ``mode="alltoall"``: rank r receives only key range r from every peer
         (NCCL all_to_all_single with uneven splits; each rank moves ~n/G keys);
       * ``mode="allgather"``: every rank receives every sorted shard (the contract form of the
         north star: "locally sorted shards merged through an NCCL allgather") and cuts its own
         key range out of each;
     or, ``mode="partition"`` (default, fastest): steps 1-3 without the local sort -- splitters come from
     a strided sample of the UNSORTED keys, ``dml_ood_partition`` scatters the keys into the G ranges
     (one 12 B/key pass instead of a 36 B/key sort) and range r travels to rank r by all-to-all;
  4. the G sorted runs of a range are merged by one more local radix sort, the positives /
     elements that precede the range are all-gathered (2 integers per rank) and the range is scanned
     with those carried counts (``dml_ood_scan_range``);
  5. the 48-byte partials are all-gathered and combined in rank order with exact integer /
     fixed-order float64 arithmetic, so every rank returns bit-identical (auroc, aupr, fpr).

The kernel-calling steps are isolated in ``CudaOps`` so that the exchange / carry / combine logic can
be exercised on CPU with the gloo backend (tests/test_distributed.py injects a NumPy stand-in).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import ood

SAMPLES_PER_RANK = 4096
PARTITION_SAMPLES_PER_RANK = 16384   # unsorted shards: a larger sample keeps the ranges balanced to ~1 %


class CudaOps:
    """The CUDA implementation of the local steps (the only product backend)."""

    def __init__(self, device, workspace: Optional[ood.OodWorkspace] = None):
        self.device = torch.device(device)
        self.ws = workspace or ood.OodWorkspace(self.device)

    # conf / gt -> packed keys (int32 view of the u32 keys), n_pos
    def make_keys(self, conf, gt, out_labels, key_base):
        from ._lib import check, lib, ptr, stream_ptr
        if conf.dtype != torch.float32:
            raise ValueError("conf must be float32 (the kernel reads const float*)")
        n = conf.numel()
        keys = self.ws.get("d_keys", 4 * max(n, 1)).view(torch.int32)[:n]
        stats = self.ws.get("d_stats", 32).view(torch.int64)[:4]
        g = gt.contiguous().view(-1)
        if g.numel() != n or g.dtype not in (torch.uint8, torch.int64):
            raise ValueError("gt must be uint8 or int64 with one label per conf value")
        with torch.cuda.device(self.device):
            check(lib().dml_ood_keygen(ptr(conf.contiguous().view(-1)), None, 0, None,
                                       ptr(g) if g.dtype == torch.uint8 else None, ptr(g) if g.dtype == torch.int64 else None,
                                       ood.label_mask(out_labels), None, 0, key_base, 1, n, ptr(keys), ptr(stats),
                                       None, None, None, 0.0, 0.0, None, 0, stream_ptr(self.device)), "dml_ood_keygen")
        return keys, stats

    def sort(self, keys, tag="a"):
        """returns a sorted int32 tensor (may alias ``keys`` or workspace memory)"""
        from ._lib import check, lib, ptr, stream_ptr
        n = keys.numel()
        if n == 0:
            return keys
        nbytes = lib().dml_ood_workspace_bytes(1, n)
        scratch = self.ws.get("d_sort_" + tag, nbytes)
        out = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().dml_ood_sort(ptr(keys), 1, n, 0, 32, ptr(scratch), scratch.numel(), C.byref(out),
                                     stream_ptr(self.device)), "dml_ood_sort")
        if out.value == keys.data_ptr():
            return keys
        off = out.value - scratch.data_ptr()
        return scratch[off: off + 4 * n].view(torch.int32)

    def lower_bound(self, sorted_keys, queries):
        from ._lib import check, lib, ptr, stream_ptr
        q = queries.to(self.device).contiguous()
        pos = torch.empty(q.numel(), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().dml_ood_lower_bound(ptr(sorted_keys), sorted_keys.numel(), ptr(q), q.numel(), ptr(pos),
                                            stream_ptr(self.device)), "dml_ood_lower_bound")
        return pos

    def count_positive(self, keys):
        from ._lib import check, lib, ptr, stream_ptr
        out = torch.zeros(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            check(lib().dml_ood_count_positive(ptr(keys), keys.numel(), ptr(out), stream_ptr(self.device)),
                  "dml_ood_count_positive")
        return out

    def sample(self, sorted_keys, n_samples):
        n = sorted_keys.numel()
        if n == 0:
            return torch.full((n_samples,), -1, dtype=torch.int32, device=self.device)
        idx = (torch.arange(n_samples, device=self.device, dtype=torch.float64) + 0.5) * (n / n_samples)
        return sorted_keys[idx.long().clamp_(max=n - 1)]

    def scan_range(self, sorted_keys, info, recall_level):
        from ._lib import check, lib, ptr, stream_ptr
        n = sorted_keys.numel()
        partial = torch.empty(ood.PARTIAL_WORDS, dtype=torch.int64, device=self.device)
        nbytes = lib().dml_ood_workspace_bytes(1, max(n, 1))
        scratch = self.ws.get("d_scan", nbytes)
        with torch.cuda.device(self.device):
            check(lib().dml_ood_scan_range(ptr(sorted_keys), n, ptr(info), recall_level, ptr(scratch), scratch.numel(),
                                           ptr(partial), stream_ptr(self.device)), "dml_ood_scan_range")
        return partial

    # ---- minority-rank building blocks (ood.rank_keys split at its collective points) ---------------------------
    def sorted_positive_keys(self, keys, n_pos):
        return ood.sorted_positive_keys(keys, n_pos, self.ws, "d_pr")

    def sort31(self, keys, tag):
        return ood.sort_keys(keys, self.ws, "d_" + tag, end_bit=31)

    def unique_groups(self, sorted_pos):
        return ood.unique_groups(sorted_pos, self.ws, "d_pr")

    def bucket_rank_counters(self, keys, S, key_base):
        if S.numel() > ood.MAX_RANK_GROUPS:
            raise ValueError("pooled_measures(mode='rank'): too many distinct positive scores; use mode='partition'")
        return ood.bucket_rank_counters(keys, S, key_base, self.ws, "d_pr")

    def pooled_scan(self, pc, cnt, total_pos, total_n, n_nan, recall_level):
        return ood.pooled_scan(pc, cnt, total_pos, total_n, n_nan, recall_level, self.ws, "d_pr")

    def empty_keys(self, n, tag):
        return self.ws.get("d_recv_" + tag, 4 * max(n, 1)).view(torch.int32)[:n]

    def slots_gather(self, records, hdr_words, slot_keys, total):
        """[n_records, stride] int32 slot records (``PositiveExchange``) -> their ``total`` keys, contiguous, in record order"""
        from ._lib import check, lib, ptr, stream_ptr
        out = self.empty_keys(total, "rk_all")
        with torch.cuda.device(self.device):
            check(lib().dml_ood_slots_gather(ptr(records), records.shape[0], records.shape[1], hdr_words, slot_keys, ptr(out),
                                             total, stream_ptr(self.device)), "dml_ood_slots_gather")
        return out

    def sample_unsorted(self, keys, n_samples):
        """strided sample of unsorted keys (int32 bit patterns; -1 marks an empty shard)"""
        n = keys.numel()
        if n == 0:
            return torch.full((n_samples,), -1, dtype=torch.int32, device=self.device)
        idx = (torch.arange(n_samples, device=self.device, dtype=torch.float64) + 0.5) * (n / n_samples)
        return keys[idx.long().clamp_(max=n - 1)]

    def partition(self, keys, inner_bounds):
        """keys (unsorted) -> (keys grouped by range, int64 counts [G]); inner_bounds: int32 bit patterns of the
        G-1 ascending range starts b_1..b_{G-1}"""
        from ._lib import check, lib, ptr, stream_ptr
        n = keys.numel()
        G = inner_bounds.numel() + 1
        out = self.ws.get("d_part", 4 * max(n, 1)).view(torch.int32)[:n]
        counts = torch.empty(G, dtype=torch.int64, device=self.device)
        wbytes = lib().dml_ood_partition_workspace_bytes(G)
        scratch = self.ws.get("d_part_ws", wbytes)
        b = inner_bounds.to(self.device).contiguous()
        with torch.cuda.device(self.device):
            check(lib().dml_ood_partition(ptr(keys), n, ptr(b) if G > 1 else None, G, ptr(out), ptr(counts), ptr(scratch),
                                          scratch.numel(), stream_ptr(self.device)), "dml_ood_partition")
        return out, counts


def _as_unsigned(t: torch.Tensor) -> torch.Tensor:
    """int32 bit patterns -> int64 values in [0, 2^32) (host-side ordering of a handful of samples)"""
    return t.to(torch.int64) & 0xFFFFFFFF


def choose_splitters(all_samples: torch.Tensor, world: int) -> torch.Tensor:
    """all_samples: int64 unsigned key values gathered from every rank (-1 entries = empty shard).
    Returns world+1 int64 boundaries b_0 = 0 < ... < b_world = 2^32 with the positive bit cleared."""
    s = all_samples[all_samples >= 0]
    s, _ = torch.sort(s)
    bounds = [0]
    for r in range(1, world):
        if s.numel() == 0:
            b = 0
        else:
            b = int(s[min(s.numel() - 1, (r * s.numel()) // world)].item()) & ~1
        bounds.append(max(b, bounds[-1]))
    bounds.append(1 << 32)
    return torch.tensor(bounds, dtype=torch.int64)


class PositiveExchange:
    """``pooled_measures(mode="rank")`` with the exchange of the positives hidden behind the per-image pass.

    Attach one to the ``ood.KeyPool`` of the evaluation (``pool.exchange = PositiveExchange(...)``): every
    ``eval_segments(..., pool=pool, method="rank")`` batch then exports its positives' score keys into the next
    fixed-capacity slot and all-gathers that slot right away -- asynchronously, on NCCL's stream, while the head and metric
    kernels of the following batches run (NVLink is otherwise idle during that phase).  When the pooled stage starts,
    every rank already holds every rank's positives: no count exchange, no local sort, no bulk all-gather on the critical
    path; one read of the slot headers sizes the merge.  Slot record (int32 words): [0:2] int64 keys in the slot,
    [2:10] the pool's running (n_pos, n_nan, n_out_of_window, -) as int64, [10:12] int64 keys pooled so far, then the
    keys.  Collective: every rank must publish the same number of slots per evaluation (``publish_empty`` pads)."""
    HDR = 12

    def __init__(self, device, slot_keys: int, max_slots: int, group=None):
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group)
        self.slot_keys = (int(slot_keys) + 1) & ~1
        self.stride = self.HDR + self.slot_keys
        self.max_slots = int(max_slots)
        self.local = torch.zeros(self.max_slots, self.stride, dtype=torch.int32, device=self.device)
        self.gathered = torch.zeros(self.max_slots, self.world, self.stride, dtype=torch.int32, device=self.device)
        self.comm = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.works = []
        self.n = 0
        self._open = False

    def begin(self):
        """start of an evaluation (all slots of the previous one must have been consumed by ``finish``)"""
        self.finish()
        self.n = 0

    def next_slot(self, need: int):
        """(keys int32 [slot_keys], count int64 [1]) of the next slot, count zeroed; ``need`` = keys the batch may export"""
        if self.n >= self.max_slots:
            raise ValueError(f"PositiveExchange: more than max_slots = {self.max_slots} batches in one evaluation")
        if need > self.slot_keys:
            raise ValueError(f"PositiveExchange: a batch may export {need} keys, slot capacity is {self.slot_keys}")
        slot = self.local[self.n]
        slot[:self.HDR].zero_()
        self._open = True
        return slot[self.HDR:], slot[:2].view(torch.int64)

    def publish(self, running_stats: torch.Tensor, keys_so_far: int):
        """header <- the pool's running counts; launch the all-gather of the slot (returns at once)"""
        if not self._open:
            raise ValueError("PositiveExchange.publish without next_slot")
        slot = self.local[self.n]
        slot[2:10].view(torch.int64).copy_(running_stats.view(-1)[:4])
        slot[10:12].view(torch.int64).fill_(int(keys_so_far))
        out = self.gathered[self.n]
        if self.comm is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                w = dist.all_gather_into_tensor(out.view(-1), slot, group=self.group, async_op=True)
        else:
            w = dist.all_gather(list(out.unbind(0)), slot, group=self.group, async_op=True)
        self.works.append(w)
        self.n += 1
        self._open = False

    def publish_empty(self, running_stats: torch.Tensor, keys_so_far: int):
        """a slot without keys (a batch that went another way, or padding to the common number of batches)"""
        self.next_slot(0)
        self.publish(running_stats, keys_so_far)

    def finish(self) -> torch.Tensor:
        """the current stream waits for the outstanding all-gathers; [n, world, stride] int32 records"""
        for w in self.works:
            w.wait()
        self.works = []
        return self.gathered[: self.n]


def pooled_measures(conf: torch.Tensor, gt: torch.Tensor, out_labels: Sequence[int] = (13,), *, group=None,
                    recall_level: float = ood.RECALL_LEVEL_DEFAULT, mode: str = "partition", ops=None,
                    workspace: Optional[ood.OodWorkspace] = None, key_base: int = ood.KEY_BASE_NONNEG,
                    timing: bool = False, keys_and_stats=None, exchange: Optional["PositiveExchange"] = None):
    """Exact pooled (auroc, aupr, fpr, info) over the (conf, gt) pairs of ALL ranks of ``group``.
    ``conf`` must be non-negative (normalised maps); positives are gt in ``out_labels``; the ranked
    score is -conf like anomaly/eval_ood_traditional.py:139-141.  Collective: every rank must call it.
    ``keys_and_stats`` = (packed keys int32 [n], stats int64 [>=3] = n_pos, n_nan, n_out_of_window), e.g. an
    ``ood.KeyPool``'s ``keys`` / ``stats[0]`` filled by the per-image evaluation: the rank's key generation is skipped
    (``conf`` / ``gt`` may then be None); the keys may be in any order.  ``exchange`` (mode="rank"): the
    ``PositiveExchange`` the per-image batches published their positives to."""
    if mode not in ("partition", "alltoall", "allgather", "rank"):
        raise ValueError("mode must be 'partition', 'alltoall', 'allgather' or 'rank'")
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = conf.device if keys_and_stats is None else keys_and_stats[0].device
    ops = ops or CudaOps(dev, workspace)
    if mode == "rank":
        return _pooled_measures_rank(conf, gt, out_labels, group, recall_level, ops, key_base, timing, keys_and_stats, exchange)
    if exchange is not None:
        raise ValueError("exchange belongs to mode='rank'")
    marks = []

    def mark(name):
        # CUDA-event phase boundaries on the current stream (``timing=True``; CUDA tensors only)
        if timing and dev.type == "cuda":
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(dev))
            marks.append((name, ev))

    mark("start")
    if keys_and_stats is None:
        keys, stats = ops.make_keys(conf, gt, out_labels, key_base)
    else:
        keys, stats = keys_and_stats
        keys, stats = keys.contiguous().view(-1), stats.view(-1)
    mark("keygen")
    partition = mode == "partition"
    n_local = keys.numel()
    # ---- splitters from a sample of every shard (regular sample of the sorted shard, or a strided sample of
    #      the unsorted one): ranges are balanced up to sampling error, correctness never depends on them --------
    if partition:
        smp = ops.sample_unsorted(keys, PARTITION_SAMPLES_PER_RANK)
    else:
        srt = ops.sort(keys, "local")
        smp = ops.sample(srt, SAMPLES_PER_RANK)
    gathered = [torch.empty_like(smp) for _ in range(world)]
    dist.all_gather(gathered, smp, group=group)
    all_s = torch.cat(gathered).cpu()
    # shards that are empty contribute the marker -1 (int32) -> drop them
    vals = _as_unsigned(all_s)
    vals[all_s == -1] = -1
    bounds = choose_splitters(vals, world)                       # int64 [world+1]
    q = (bounds[:-1] & 0xFFFFFFFF).to(torch.int64)
    queries = torch.where(q >= (1 << 31), q - (1 << 32), q).to(torch.int32)   # bit patterns
    mark("local_sort+sample+splitters")
    if partition:
        srt, cnt_dev = ops.partition(keys, queries[1:])           # grouped by range, not sorted
        mark("partition")
        send_counts = cnt_dev.cpu().tolist()
    else:
        cuts = ops.lower_bound(srt, queries)                      # [world] start of every range in my shard
        cuts = torch.cat([cuts.cpu(), torch.tensor([n_local])])
        send_counts = (cuts[1:] - cuts[:-1]).tolist()

    # totals + stats (n_pos, n_nan, n_oow) in one tiny all_reduce
    tot = torch.cat([stats[:3].to(torch.int64), torch.tensor([n_local], dtype=torch.int64, device=dev)])
    dist.all_reduce(tot, group=group)
    total_pos, n_nan, n_oow, total_n = [int(v) for v in tot.cpu().tolist()]
    if n_nan:
        raise ValueError("Input contains NaN.")
    if n_oow:
        raise ValueError("pooled_measures: conf must be non-negative (keys left the packed 31-bit window)")

    # ---- exchange of the sorted shards ---------------------------------------------------------------
    cnt = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    all_cnt = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(all_cnt, cnt, group=group)
    all_cnt = torch.stack(all_cnt).cpu()                          # [src, dst]
    recv_counts = all_cnt[:, rank].tolist()
    m = int(sum(recv_counts))
    mine = ops.empty_keys(m, "range")
    moved_bytes = 0
    mark("counts_exchange")
    if mode in ("alltoall", "partition"):
        dist.all_to_all_single(mine, srt, recv_counts, send_counts, group=group)
        moved_bytes = 4 * (m - recv_counts[rank])
    else:
        sizes = all_cnt.sum(dim=1).tolist()
        nmax = int(max(sizes)) if sizes else 0
        pad = ops.empty_keys(nmax, "pad")
        pad[:n_local].copy_(srt)
        shards = [ops.empty_keys(nmax, f"shard{r}") for r in range(world)]
        dist.all_gather(shards, pad, group=group)
        off = 0
        starts = torch.cat([torch.zeros(world, 1, dtype=torch.int64), torch.cumsum(all_cnt, dim=1)], dim=1)
        for r in range(world):
            a, c = int(starts[r, rank]), int(recv_counts[r])
            mine[off: off + c].copy_(shards[r][a: a + c])
            off += c
        moved_bytes = 4 * int(sum(sizes) - sizes[rank])

    mark("key_exchange")
    # ---- merge (re-sort) my range, carry the counts that precede it, scan -----------------------------------
    merged = ops.sort(mine, "merge")
    mark("range_sort")
    pos_r = ops.count_positive(merged)
    pr = torch.cat([pos_r.view(1), torch.tensor([m], dtype=torch.int64, device=dev)])
    all_pr = [torch.empty_like(pr) for _ in range(world)]
    dist.all_gather(all_pr, pr, group=group)
    all_pr = torch.stack(all_pr).cpu()
    pos_before = int(all_pr[:rank, 0].sum())
    idx_before = int(all_pr[:rank, 1].sum())
    info = torch.tensor([pos_before, idx_before, total_pos, total_n], dtype=torch.int64, device=dev)
    mark("carry_exchange")
    partial = ops.scan_range(merged, info, recall_level)
    mark("range_scan")
    parts = [torch.empty_like(partial) for _ in range(world)]
    dist.all_gather(parts, partial, group=group)
    auroc, aupr, fpr, groups = ood.combine_partials([p.cpu().numpy() for p in parts], total_pos, total_n, recall_level)
    mark("combine")
    out = {"n_pos": total_pos, "n_neg": total_n - total_pos, "n_groups": groups, "range_keys": m,
           "exchanged_bytes": moved_bytes, "mode": mode}
    if marks:
        marks[-1][1].synchronize()
        out["phase_ms"] = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])}
    return auroc, aupr, fpr, out


def _pooled_measures_rank(conf, gt, out_labels, group, recall_level, ops, key_base, timing, keys_and_stats, exchange=None):
    """``pooled_measures(mode="rank")``: the minority-rank form of the pooled metric.  No negative ever leaves its GPU:

      1. every rank compacts and radix-sorts the score keys of ITS positives (~1 % of the pairs);
      2. the locally sorted shards are all-gathered (NCCL) and merged on every rank into the distinct positive scores
         S[g] with multiplicities pc[g] -- identical on all ranks;
      3. every rank buckets its negatives by S and locates them in shared memory (``dml_ood_bucket_rank``): counters
         bt[g] / eq[g] of its own pixels;
      4. one all-reduce (exact int64 sums) of the 2G + 2 counters, then the same scan over the G groups on every rank
         -> bit-identical (auroc, aupr, fpr) everywhere, equal to the single-GPU result.

    Exchanged per rank: the positives (4 B each) and 16 B per distinct positive score, instead of 4 B for every pair."""
    world = dist.get_world_size(group)
    dev = conf.device if keys_and_stats is None else keys_and_stats[0].device
    marks = []

    def mark(name):
        if timing and dev.type == "cuda":
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(dev))
            marks.append((name, ev))

    mark("start")
    pos_keys = pos_count = None
    if keys_and_stats is None:
        keys, stats = ops.make_keys(conf, gt, out_labels, key_base)
    else:
        keys, stats = keys_and_stats[:2]
        keys, stats = keys.contiguous().view(-1), stats.view(-1)
        if len(keys_and_stats) == 4:          # (.., KeyPool.pos, KeyPool.pos_count): positives gathered by the rank batches
            pos_keys, pos_count = keys_and_stats[2:]
    mark("keygen")
    n_local = keys.numel()
    me = dist.get_rank(group)
    allpos = hold = None
    if exchange is not None and exchange.n > 0:
        # the batches published their positives while the per-image pass ran: all that is left is to wait for the last
        # all-gathers and to read the slot headers (identical on every rank, so every rank takes the same branch below)
        rec = exchange.finish()                                          # [batches, world, stride] int32
        nb = rec.shape[0]
        hdr = rec[:, :, :exchange.HDR].contiguous().cpu().view(torch.int64)   # [batches, world, (count, n_pos, n_nan, n_oow, -, n)]
        last = hdr[-1]
        total_pos, n_nan, n_oow, total_n = int(last[:, 1].sum()), int(last[:, 2].sum()), int(last[:, 3].sum()), int(last[:, 5].sum())
        if n_nan:
            raise ValueError("Input contains NaN.")
        if n_oow:
            raise ValueError("pooled_measures: conf must be non-negative (keys left the packed 31-bit window)")
        mark("positive_exchange_wait")
        out = {"n_pos": total_pos, "n_neg": total_n - total_pos, "n_groups": -1, "mode": "rank",
               "exchanged_bytes": 4 * exchange.stride * (world - 1) * nb, "positives_from": "slots"}
        if total_pos == 0 or total_pos == total_n:
            return float("nan"), float("nan"), float("nan"), out
        complete = bool((hdr[:, :, 0].sum(0) == last[:, 1]).all()) and bool((hdr[:, :, 0] <= exchange.slot_keys).all()) \
            and bool((hdr[:, :, 0] >= 0).all())
        if not complete:  # some batch kept its positives to itself (e.g. an overflowed segment): compact + all-gather below
            hold = out
        else:
            allpos = ops.slots_gather(rec.reshape(nb * world, exchange.stride), exchange.HDR, exchange.slot_keys, total_pos)
            mark("slots_gather")
    if allpos is None:
        # local counts -> every rank (one small all_gather; the host needs them to size the exchange)
        have = pos_count.view(-1)[:1].to(torch.int64) if pos_count is not None else torch.full((1,), -1, dtype=torch.int64, device=dev)
        mine = torch.cat([stats[:3].to(torch.int64), torch.tensor([n_local], dtype=torch.int64, device=dev), have])
        allc = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine, group=group)
        allc = torch.stack(allc).cpu()                                      # [rank, (n_pos, n_nan, n_oow, n, positives on hand)]
        total_pos, n_nan, n_oow, total_n = [int(v) for v in allc[:, :4].sum(0).tolist()]
        if n_nan:
            raise ValueError("Input contains NaN.")
        if n_oow:
            raise ValueError("pooled_measures: conf must be non-negative (keys left the packed 31-bit window)")
        pos_counts = [int(v) for v in allc[:, 0].tolist()]
        mark("counts_exchange")
        if pos_keys is not None and int(allc[me, 4]) == pos_counts[me] and pos_counts[me] <= pos_keys.numel():
            srt = ops.sort31(pos_keys[: pos_counts[me]], "rk_local")                # the list the per-image pass left
        else:
            srt = ops.sorted_positive_keys(keys, pos_counts[me])                    # int32 [n_pos_local], ascending
        mark("local_positive_sort")
        pmax = max(pos_counts) if pos_counts else 0
        out = {"n_pos": total_pos, "n_neg": total_n - total_pos, "n_groups": -1, "mode": "rank", "positives_from": "allgather",
               "exchanged_bytes": 4 * (sum(pos_counts) - pos_counts[me]) + (hold["exchanged_bytes"] if hold else 0)}
        if total_pos == 0 or total_pos == total_n:
            return float("nan"), float("nan"), float("nan"), out
        pad = ops.empty_keys(pmax, "rk_pad")
        pad[: srt.numel()].copy_(srt)
        shards = [ops.empty_keys(pmax, f"rk_shard{r}") for r in range(world)]
        dist.all_gather(shards, pad, group=group)
        allpos = ops.empty_keys(total_pos, "rk_all")
        off = 0
        for r in range(world):
            allpos[off: off + pos_counts[r]].copy_(shards[r][: pos_counts[r]])
            off += pos_counts[r]
        mark("positive_allgather")
    merged = ops.sort31(allpos, "rk_merge")
    S, pc, G = ops.unique_groups(merged)
    mark("merge_positives")
    cnt = ops.bucket_rank_counters(keys, S, key_base)
    mark("bucket_rank")
    dist.all_reduce(cnt, group=group)
    out["exchanged_bytes"] += 8 * cnt.numel()
    mark("counter_allreduce")
    res = ops.pooled_scan(pc, cnt, total_pos, total_n, n_nan, recall_level)
    r = res.cpu().numpy()
    mark("scan")
    out["n_pos_groups"] = G
    if marks:
        marks[-1][1].synchronize()
        out["phase_ms"] = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])}
    return float(r[0, 0]), float(r[0, 1]), float(r[0, 2]), out


def mean_of_per_image(per_image_vals: torch.Tensor, group=None):
    """The reference's aggregation (eval_ood_traditional.py:569,641): mean over images of the per-image
    metrics, across ranks: one all_reduce of (sum_auroc, sum_aupr, sum_fpr, count).  ``per_image_vals``
    is the float64 [n,>=3] device tensor of ``ood.eval_segments`` (NaN rows = skipped images)."""
    v = per_image_vals[:, :3]
    ok = ~torch.isnan(v[:, 0])
    acc = torch.cat([torch.where(ok.unsqueeze(1), v, torch.zeros_like(v)).sum(0), ok.sum().to(v.dtype).view(1)])
    dist.all_reduce(acc, group=group)
    acc = acc.cpu()
    cnt = float(acc[3])
    return (float(acc[0]) / cnt, float(acc[1]) / cnt, float(acc[2]) / cnt, int(cnt)) if cnt else (float("nan"),) * 3 + (0,)


def allreduce_counts(confusion: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-rank confusion / class-count matrices (int64, exact)."""
    dist.all_reduce(confusion, group=group)
    return confusion
