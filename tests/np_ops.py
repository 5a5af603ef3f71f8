"""NumPy stand-in for ``dml_b200.distributed.CudaOps`` (TEST DOUBLE, CPU only): the same local steps
(pack keys, sort, cut, count, group scan) restated with NumPy so that the exchange / carry / combine logic of
``pooled_measures`` can run under the gloo backend without a GPU."""
import numpy as np
import torch


def _u(t):  # int32 bit patterns -> uint32 numpy
    return t.numpy().view(np.uint32)


class NumpyOps:
    def make_keys(self, conf, gt, out_labels, key_base):
        c = conf.contiguous().view(-1).numpy().astype(np.float32)
        g = gt.contiguous().view(-1).numpy()
        pos = np.isin(g, list(out_labels))
        f = np.where(c == 0, np.float32(0), c)
        bits = f.view(np.uint32)
        srt = np.where(bits & 0x80000000, ~bits, bits | np.uint32(0x80000000)).astype(np.uint32)
        n_nan = int(np.isnan(c).sum())
        rel = (srt.astype(np.int64) - key_base)
        oow = (rel < 0) | (rel >= (1 << 31))
        keys = ((np.clip(rel, 0, (1 << 31) - 1).astype(np.uint64) << 1) | pos).astype(np.uint32)
        stats = torch.tensor([int(pos.sum()), n_nan, int(oow.sum()), 0], dtype=torch.int64)
        return torch.from_numpy(keys.view(np.int32).copy()), stats

    def sort(self, keys, tag="a"):
        return torch.from_numpy(np.sort(_u(keys)).view(np.int32).copy())

    def lower_bound(self, sorted_keys, queries):
        return torch.from_numpy(np.searchsorted(_u(sorted_keys), _u(queries.contiguous()), side="left").astype(np.int64))

    def count_positive(self, keys):
        return torch.tensor([int((_u(keys) & 1).sum())], dtype=torch.int64)

    def sample(self, sorted_keys, n_samples):
        n = sorted_keys.numel()
        if n == 0:
            return torch.full((n_samples,), -1, dtype=torch.int32)
        idx = ((torch.arange(n_samples, dtype=torch.float64) + 0.5) * (n / n_samples)).long().clamp_(max=n - 1)
        return sorted_keys[idx]

    def scan_range(self, sorted_keys, info, recall_level):
        k = _u(sorted_keys).astype(np.int64)
        pos_before, idx_before, total_pos, total_n = [int(v) for v in info.tolist()]
        out = np.zeros(10, np.int64)
        out[2], out[5], out[6] = -1, (1 << 63) - 1, -1
        # T* = largest t with float64 t/P <= recall_level
        tstar = 0
        if total_pos > 0:
            tstar = min(total_pos, max(0, int(np.floor(recall_level * total_pos))))
            while tstar < total_pos and (tstar + 1) / total_pos <= recall_level:
                tstar += 1
            while tstar > 0 and tstar / total_pos > recall_level:
                tstar -= 1
        if k.size:
            score, lab = k >> 1, k & 1
            ends = np.r_[np.nonzero(np.diff(score))[0], k.size - 1]
            starts = np.r_[0, ends[:-1] + 1]
            cpos = np.cumsum(lab)
            tps = pos_before + cpos[ends]
            n_so_far = idx_before + ends + 1
            pos_g = cpos[ends] - np.r_[0, cpos[ends[:-1]]]
            neg_g = (ends - starts + 1) - pos_g
            num = int(sum(int(a) * (2 * int(b) - int(c)) for a, b, c in zip(neg_g, tps, pos_g)))
            ap = 0.0
            a = (-1, 0, 0)
            b = ((1 << 63) - 1, -1, 0)
            for e, t, pg, ns in zip(ends, tps, pos_g, n_so_far):
                t, pg, ns = int(t), int(pg), int(ns)
                if pg:
                    ap += float(pg) * (float(t) / float(ns))
                if t - pg < total_pos:
                    idx = idx_before + int(e)
                    if t <= tstar:
                        if idx > a[0]:
                            a = (idx, t, ns - t)
                    elif t < b[0] or (t == b[0] and idx > b[1]):
                        b = (t, idx, ns - t)
            out.view(np.uint64)[0] = num
            out.view(np.float64)[1] = ap
            out[2], out[3], out[4] = a
            out[5], out[6], out[7] = b
            out[8] = len(ends)
        return torch.from_numpy(out)

    def sample_unsorted(self, keys, n_samples):
        return self.sample(keys, n_samples)

    def partition(self, keys, inner_bounds):
        k = _u(keys)
        b = np.searchsorted(_u(inner_bounds.contiguous()), k, side="right") if inner_bounds.numel() else np.zeros(k.size, np.int64)
        order = np.argsort(b, kind="stable")
        counts = np.bincount(b, minlength=inner_bounds.numel() + 1).astype(np.int64)
        return torch.from_numpy(k[order].view(np.int32).copy()), torch.from_numpy(counts)

    def empty_keys(self, n, tag):
        return torch.empty(n, dtype=torch.int32)

    def slots_gather(self, records, hdr_words, slot_keys, total):
        out = []
        for r in records.numpy():
            c = int(r[:2].view(np.int64)[0])
            out.append(r[hdr_words: hdr_words + min(max(c, 0), slot_keys)])
        k = np.concatenate(out) if out else np.zeros(0, np.int32)
        assert k.size == total
        return torch.from_numpy(k.copy())

    # ---- minority-rank mode (distributed.pooled_measures(mode="rank")): NumPy restatement of csrc/ood_pool_rank.cu ----
    def sorted_positive_keys(self, keys, n_pos):
        k = _u(keys)
        return torch.from_numpy(np.sort(k[(k & 1) == 1] >> 1).astype(np.uint32).view(np.int32).copy())

    def sort31(self, keys, tag):
        return self.sort(keys, tag)

    def unique_groups(self, sorted_pos):
        s, c = np.unique(_u(sorted_pos), return_counts=True)
        return (torch.from_numpy(s.astype(np.uint32).view(np.int32).copy()),
                torch.from_numpy(c.astype(np.uint32).view(np.int32).copy()), int(s.size))

    def bucket_rank_counters(self, keys, S, key_base):
        k = _u(keys)
        s = _u(S)
        neg = (k[(k & 1) == 0] >> 1).astype(np.uint32)
        lb = np.searchsorted(s, neg, side="left")
        eq = (lb < s.size) & (s[np.minimum(lb, s.size - 1)] == neg)
        cnt = np.bincount(2 * lb + eq, minlength=2 * s.size + 2).astype(np.int64)
        return torch.from_numpy(cnt)

    def pooled_scan(self, pc, cnt, total_pos, total_n, n_nan, recall_level):
        pc = _u(pc).astype(np.int64)
        c = cnt.numpy().astype(np.int64)
        G, P, N = pc.size, int(total_pos), int(total_n - total_pos)
        row = np.zeros(7, np.float64)
        row.view(np.int64)[3:] = [P, N, n_nan, -1]
        if P <= 0 or N <= 0 or G == 0:
            row[:3] = np.nan
            return torch.from_numpy(row.reshape(1, 7))
        bt, eq, tail = c[0:2 * G:2], c[1:2 * G:2], int(c[2 * G])
        T = np.cumsum(pc)
        F = np.cumsum(bt + eq)
        au = sum(int(b) * 2 * int(t - p) + int(e) * (2 * int(t) - int(p)) for b, e, t, p in zip(bt, eq, T, pc)) + 2 * P * tail
        ap = 0.0
        for p, t, f in zip(pc, T, F):
            ap += float(p) * (float(t) / float(t + f))
        tstar = min(P, max(0, int(np.floor(recall_level * P))))
        while tstar < P and (tstar + 1) / P <= recall_level:
            tstar += 1
        while tstar > 0 and tstar / P > recall_level:
            tstar -= 1
        gs = int(np.searchsorted(T, tstar, side="right")) - 1          # last group with tps <= T*
        da = db = float("inf")
        a_fps = b_fps = 0
        Tg, Fg = (int(T[gs]), int(F[gs])) if gs >= 0 else (0, 0)
        trail = int(bt[gs + 1]) if gs + 1 <= G - 1 else 0
        if gs >= 0 or trail > 0:
            da, a_fps = abs(Tg / P - recall_level), Fg + trail
        if gs + 1 <= G - 1:
            trail = int(bt[gs + 2]) if gs + 2 <= G - 1 else 0
            db, b_fps = abs(int(T[gs + 1]) / P - recall_level), int(F[gs + 1]) + trail
        row[0] = au / (2.0 * P * N)
        row[1] = ap / P
        row[2] = (b_fps if db <= da else a_fps) / N
        return torch.from_numpy(row.reshape(1, 7))
