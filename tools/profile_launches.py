#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the
LAST step (everything after the middle head_kernel launch) and shares."""
import collections
import csv
import re
import sys


def main(path, last_fraction=0.5):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    head_idx = [i for i, n in enumerate(names) if "head_kernel" in n]
    start = head_idx[int(len(head_idx) * (1 - last_fraction))] - 1 if head_idx else 0
    agg = collections.OrderedDict()
    for r in rows[max(start, 0):]:
        n = r["Kernel Name"]
        m = re.search(r"(\w+_kernel)", n)
        key = m.group(1) if m else n[:48]
        if "unsigned long long" in n and "onesweep" in n:
            key += "<u64>"
        t = float(r["Metric Value"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':36s} {'n':>5s} {'total ms':>10s} {'share':>7s} {'avg us':>10s}")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:36s} {c:5d} {t / 1e6:10.3f} {100 * t / tot:6.1f}% {t / c / 1e3:10.1f}")
    print(f"{'total':36s} {'':5s} {tot / 1e6:10.3f}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
