#!/usr/bin/env python
"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into the tracked text table under profiles/:
the last `--last N` launches (one timed step of bench.py) with grid / block / duration, and the per-kernel totals."""
import argparse
import collections
import csv
import re


def short(name):
    name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.split("::")[-1].replace("void ", "").strip() or "?"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--last", type=int, default=42)
    ap.add_argument("--title", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hi], [r for r in rows[hi + 1:] if len(r) > 5]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    gi, bi = h.index("Grid Size"), h.index("Block Size")
    step = data[-a.last:]
    if a.title:
        print("# " + a.title)
    tot = collections.OrderedDict()
    total = 0.0
    for r in step:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        n = short(r[ki])
        print(f"{n:34s} grid={r[gi]:>16s} block={r[bi]:>14s} {v:9.1f} us")
        t = tot.setdefault(n, [0, 0.0])
        t[0] += 1
        t[1] += v
        total += v
    print(f"total {total:.1f} us\n")
    print(f"{'kernel':34s} {'n':>4s} {'total us':>10s} {'share':>7s}")
    for n, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:34s} {c:4d} {v:10.1f} {100 * v / total:6.1f}%")


if __name__ == "__main__":
    main()
